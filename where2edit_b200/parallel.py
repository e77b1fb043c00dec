"""Batch sharding of the synthesis path across the GPUs of one box.

Synthesis is embarrassingly parallel (each image depends only on its own latent, replicated weights and
the shared fixed noise buffers, models/stylegan2/model.py:492-494), so ranks never exchange data on the
path itself.  Collectives, as in the reference:
  * the final all-gather of images / latents (`gather_images`, `PeerGather`), and its autograd form
    `GatherLayer` / `gather_with_grad` (utils.py:114-131, used at attention/run_attention.py:1313-1314);
  * for the optimisation loop, the all-reduce of the shared mapper's gradients (`GradBucketReducer`: what
    DistributedDataParallel does for the mapper at run_attention.py:1022-1030), bucketed and issued on a side
    stream while the generator backward is still producing the gradients of earlier layers.
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


class GatherLayer(torch.autograd.Function):
    """utils.py:114-131: all-gather with autograd.  forward returns one tensor per rank (a tuple, in rank order);
    backward hands back ONLY the gradient of this rank's own slice (`grads[rank]`) -- the gradients that other
    ranks' losses would send to this rank's features are dropped, exactly as in the reference, whose
    DistributedDataParallel then averages the mapper gradients over ranks."""

    @staticmethod
    def forward(ctx, input, group=None):
        ctx.group = group
        world = dist.get_world_size(group)
        x = input.contiguous()
        buf = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(buf.view(world * x.shape[0], *x.shape[1:]) if x.ndim else buf, x, group=group)
        return tuple(buf.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        return grads[dist.get_rank(ctx.group)].clone(), None


def gather_with_grad(x, group=None):
    """`torch.cat(GatherLayer.apply(x), dim=0)` (run_attention.py:1313-1314); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return torch.cat(GatherLayer.apply(x, group), dim=0)


class GradBucketReducer:
    """All-reduce (average) of a shared mapper's gradients over the ranks, overlapped with the backward that
    produces them -- the job DistributedDataParallel(find_unused_parameters=True) does for the mapper in the
    reference (run_attention.py:1022-1030); the generator is frozen and replicated, so nothing else is reduced.

    All gradients live in ONE flat fp32 buffer (`p.grad` of every parameter is a view into it), cut into buckets
    of about `bucket_bytes` in REVERSE parameter order: in the edit loop the styles of the last generator layers
    receive their gradient first (the backward runs image -> ... -> 4x4), so their mapper branches finish first.
    A post-accumulate-grad hook counts a bucket's parameters down; when the last one lands, the bucket is
    all-reduced on a side stream (NCCL over NVLink) while the main stream keeps running the generator backward.
    `finish()` reduces the buckets that never completed (parameters without a gradient this step contribute
    zeros, as with find_unused_parameters=True), joins the side stream and divides by the world size.

        reducer = GradBucketReducer(mapper.parameters())
        for step in ...:
            reducer.zero_grad()                # instead of optimizer.zero_grad(): keeps the views
            loss.backward()                    # buckets are reduced as they fill
            reducer.finish()
            optimizer.step()
    """

    def __init__(self, params, group=None, bucket_bytes=25 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucketReducer: no trainable parameter")
        self.group = group
        self.average = average
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("GradBucketReducer: parameters must live on one device")
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []       # [begin, end) ranges of the flat buffer
        self._bucket_of = {}
        self._views = {}
        off = begin = 0
        count = 0
        for p in order:
            if p.dtype != torch.float32:
                raise ValueError("GradBucketReducer: fp32 parameters only (the reference trains the mapper in fp32)")
            self._views[p] = self.flat[off:off + p.numel()].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
            count += 1
            if (off - begin) * 4 >= bucket_bytes:
                self.buckets.append((begin, off, count))
                begin, count = off, 0
        if count:
            self.buckets.append((begin, off, count))
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._pending = []
        self._works = []
        self._launched = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.zero_grad()

    def zero_grad(self):
        self.flat.zero_()
        for p in self.params:
            p.grad = self._views[p]
        self._pending = [c for (_, _, c) in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []

    def _on_grad(self, p):
        view = self._views[p]
        if p.grad is not view and p.grad.data_ptr() != view.data_ptr():   # someone replaced .grad (set_to_none)
            view.copy_(p.grad)
            p.grad = view
        b = self._bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        if self._launched[b]:
            return
        self._launched[b] = True
        if self.world == 1:
            return
        begin, end, _ = self.buckets[b]
        chunk = self.flat[begin:end]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.stream):
                dist.all_reduce(chunk, group=self.group)
        else:
            self._works.append(dist.all_reduce(chunk, group=self.group, async_op=True))

    def finish(self):
        """Reduce what is still outstanding, wait for the collectives, average.  Returns the flat gradient."""
        for b in range(len(self.buckets)):
            self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        if self.average and self.world > 1:
            self.flat.div_(self.world)
        return self.flat

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_mapper_grads(module_or_params, group=None, bucket_bytes=25 << 20):
    """Convenience constructor: `reducer = allreduce_mapper_grads(mapper)`; see GradBucketReducer."""
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    return GradBucketReducer(params, group=group, bucket_bytes=bucket_bytes)


class PeerGather:
    """All-gather of equal per-rank shards by COPY-ENGINE pushes into peer-mapped (symmetric-memory) buffers
    over NVLink: no SM is used, so the transfer overlaps the persistent convolution kernels, which occupy
    every SM (the NCCL all-gather kernels can only run in the gaps between them).

        pg = PeerGather(shard.shape, shard.dtype, device)          # once (collective)
        pg.own(slot).copy_(image)                                  # producer writes its shard in place ...
        pg.push(None, slot)                                        # ... and pushes it to the peers, every step
        full = pg.result(slot)                                     # consumable on the same stream right after

    `push` is a complete per-step collective: a cross-rank barrier BEFORE the copies (every rank has stopped
    reading this slot: a rank enters it only after its own reads, which precede the push in stream order) and one
    AFTER them (every rank's shard has landed everywhere), both device-side on the pushing stream, so `result`
    may be read as soon as `push` returns (in stream order).  `slots` independent buffers let step i+1 be pushed
    while step i is still being read on another stream (make that stream's reads precede the next push of the
    same slot, e.g. `push_stream.wait_stream(reader_stream)`)."""

    def __init__(self, shard_shape, dtype, device, group=None, slots=2):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.shape = (slots, self.world) + tuple(shard_shape)
        self.buf = symm.empty(self.shape, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.peers = [self.handle.get_buffer(r, self.shape, dtype) for r in range(self.world)]

    def own(self, slot):
        """This rank's shard inside its own buffer: the producer can write there directly (no self-copy)."""
        return self.buf[slot, self.rank]

    def push(self, local, slot, sync=True):
        """Copy this rank's shard into slot `slot` of every rank's buffer (current stream, asynchronous).
        `local=None`: the shard already sits in `own(slot)` and only the peers are written.  `sync=False` skips the
        two barriers (an upper bound for measurements only: the result is then NOT consumable per step)."""
        if sync:
            self.handle.barrier(channel=slot)
        if local is not None:
            self.buf[slot, self.rank].copy_(local, non_blocking=True)
        src = self.buf[slot, self.rank]
        for r in range(1, self.world):
            self.peers[(self.rank + r) % self.world][slot, self.rank].copy_(src, non_blocking=True)
        if sync:
            self.handle.barrier(channel=self.shape[0] + slot)

    def barrier(self):
        """All pushes issued before it (on every rank, in stream order) have landed when it returns on the stream."""
        self.handle.barrier()

    def result(self, slot):
        """[world * shard, ...] in rank order (valid after barrier())."""
        return self.buf[slot].flatten(0, 1)
