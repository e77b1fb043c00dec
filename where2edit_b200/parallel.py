"""Batch sharding of the synthesis path across the GPUs of one box.

Synthesis is embarrassingly parallel (each image depends only on its own latent, replicated weights and
the shared fixed noise buffers, models/stylegan2/model.py:492-494), so ranks never exchange data on the
path itself.  The only collective is the north-star's final all-gather of images / latents, the
equivalent of the reference's GatherLayer (utils.py:114-131) for the no-grad case.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous [begin, end) slice of `total` items owned by `rank`; the first `total % world`
    ranks take one extra item (ragged totals are legal, empty shards too)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard(tensor_or_list, rank=None, world=None):
    """This rank's slice of a batch: a tensor [B,...] or a stylespace list of tensors [B,1,C,1,1]."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if isinstance(tensor_or_list, (list, tuple)):
        return [shard(t, rank, world) for t in tensor_or_list]
    b, e = shard_bounds(tensor_or_list.shape[0], rank, world)
    return tensor_or_list[b:e]


def gather_images(local, total=None, group=None, out=None):
    """All-gather of per-rank image (or latent) batches along dim 0, in rank order.  Equal shard sizes use
    one `all_gather_into_tensor`; ragged shards (total given) are padded to the largest shard and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    if total is None or total % world == 0:
        shape = (world * local.shape[0],) + tuple(local.shape[1:])
        if out is None:
            out = torch.empty(shape, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    buf = torch.empty((world * biggest,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * biggest: r * biggest + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(parts, 0)
