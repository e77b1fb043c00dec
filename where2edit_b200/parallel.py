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


def synthesize_sharded(generate, latents, micro_batch=32, gather=True, group=None):
    """a global latent batch [B_total, ...] (the same tensor on every rank) is split
    contiguously by rank; each rank runs `generate(chunk) -> images` on micro-batches of at most `micro_batch`
    latents (a 1024^2 forward holds ~0.45 GB of activations per image, so the micro-batch bounds memory, not
    the batch) and concatenates the results in order.  `gather=True` returns the whole batch on every rank
    (one all-gather; ragged totals and empty shards are handled), otherwise this rank's shard.
    Because every kernel of the path is batch-invariant, the result equals `generate(latents)` on one GPU."""
    if micro_batch < 1:
        raise ValueError("micro_batch must be positive")
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    total = latents.shape[0]
    b, e = shard_bounds(total, rank, world)
    chunks = [generate(latents[i:min(i + micro_batch, e)]) for i in range(b, e, micro_batch)]
    local = torch.cat(chunks, 0) if len(chunks) > 1 else (chunks[0] if chunks else None)
    if world == 1 or not gather:
        return local if local is not None else latents.new_zeros((0,))
    if total == 0:
        return latents.new_zeros((0,))
    if total < world:   # some rank has nothing to run: it learns the image shape / dtype from its peers
        metas = [None] * world
        dist.all_gather_object(metas, None if local is None else (tuple(local.shape[1:]), local.dtype), group=group)
        shape, dtype = next(m for m in metas if m is not None)
        if local is None:
            local = torch.zeros((0,) + shape, dtype=dtype, device=latents.device)
    return gather_images(local, total=total, group=group)


class PeerGather:
    """All-gather of equal per-rank shards by COPY-ENGINE pushes into peer-mapped (symmetric-memory) buffers
    over NVLink: no SM is used, so the transfer overlaps the persistent convolution kernels, which occupy
    every SM (the NCCL all-gather kernels can only run in the gaps between them).

        pg = PeerGather(shard.shape, shard.dtype, device)          # once (collective)
        pg.push(shard, slot)                                       # every step, on any stream
        pg.barrier(); full = pg.result(slot)                       # before reading / before a slot is reused

    `slots` independent buffers let step i+1 be pushed while step i is still being read."""

    def __init__(self, shard_shape, dtype, device, group=None, slots=2):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.shape = (slots, self.world) + tuple(shard_shape)
        self.buf = symm.empty(self.shape, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.peers = [self.handle.get_buffer(r, self.shape, dtype) for r in range(self.world)]

    def push(self, local, slot):
        """Copy this rank's shard into slot `slot` of every rank's buffer (current stream, asynchronous)."""
        for r in range(self.world):
            self.peers[(self.rank + r) % self.world][slot, self.rank].copy_(local, non_blocking=True)

    def barrier(self):
        """All pushes issued before it (on every rank, in stream order) have landed when it returns on the stream."""
        self.handle.barrier()

    def result(self, slot):
        """[world * shard, ...] in rank order (valid after barrier())."""
        return self.buf[slot].flatten(0, 1)
