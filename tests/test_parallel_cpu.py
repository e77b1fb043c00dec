"""Host logic of the multi-GPU path on CPU: world-size-2 gloo processes (no GPU needed)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from where2edit_b200 import parallel  # noqa: E402


def test_shard_bounds_cover_the_batch_exactly():
    for total in (0, 1, 5, 8, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _worker(rank, world, port, total, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 3 * 2 * 2, dtype=torch.float32).reshape(total, 3, 2, 2)
        mine = parallel.shard(full)
        b, e = parallel.shard_bounds(total, rank, world)
        assert torch.equal(mine, full[b:e])
        styles = [full[:, :1], full[:, 1:2]]
        assert all(torch.equal(s, t[b:e]) for s, t in zip(parallel.shard(styles), styles))
        # "synthesis" stand-in: a per-sample function, so shard-then-gather must equal the unsharded result
        got = parallel.gather_images(mine * 2 + 1, total=total)
        assert torch.equal(got, full * 2 + 1), (rank, got.shape)
        torch.save(got, os.path.join(result_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_shard_then_gather_equals_single_process(tmp_path, total):
    world, port = 2, 29500 + (os.getpid() % 2000) + total
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b) and a.shape[0] == total


def _mb_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for total, mb in [(0, 4), (1, 4), (5, 2), (8, 3), (7, 100)]:
            latents = torch.arange(total * 6, dtype=torch.float32).reshape(total, 2, 3)
            calls = []

            def generate(chunk):      # per-sample "synthesis" stand-in producing [n, 3, 2, 2] images
                calls.append(chunk.shape[0])
                return (chunk.sum((1, 2)).reshape(-1, 1, 1, 1) + torch.arange(12.).reshape(1, 3, 2, 2)) * 0.5

            full = parallel.synthesize_sharded(generate, latents, micro_batch=mb)
            b, e = parallel.shard_bounds(total, rank, world)
            assert calls == [min(mb, e - i) for i in range(b, e, mb)], (total, mb, calls)
            calls.clear()
            want = generate(latents) if total else torch.zeros(0)
            assert full.shape[0] == total and (total == 0 or torch.equal(full, want)), (rank, total, mb)
            mine = parallel.synthesize_sharded(generate, latents, micro_batch=mb, gather=False)
            assert mine.shape[0] == e - b and (e == b or torch.equal(mine, want[b:e]))
        torch.save(torch.ones(1), os.path.join(result_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_micro_batched_sharded_synthesis_equals_single_process(tmp_path):
    """Config 5's host logic: contiguous shards, micro-batches, ragged totals, an empty shard, final all-gather."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_mb_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok0.pt").exists() and (tmp_path / "ok1.pt").exists()


def test_micro_batched_synthesis_without_process_group():
    latents = torch.randn(5, 4)
    out = parallel.synthesize_sharded(lambda w: w * 2, latents, micro_batch=2)
    assert torch.equal(out, latents * 2)
    with pytest.raises(ValueError):
        parallel.synthesize_sharded(lambda w: w, latents, micro_batch=0)


def _make_mapper():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 5), torch.nn.Tanh(),
                               torch.nn.Linear(5, 4))


def _grad_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- GatherLayer (utils.py:114-131): forward = all-gather in rank order, backward = OWN slice only
        x = (torch.arange(6, dtype=torch.float32).reshape(2, 3) + 10 * rank).requires_grad_(True)
        parts = parallel.GatherLayer.apply(x)
        assert len(parts) == world and all(torch.equal(parts[r], torch.arange(6.).reshape(2, 3) + 10 * r) for r in range(world))
        full = parallel.gather_with_grad(x * 1.0)
        weight = torch.arange(1., 1. + full.numel()).reshape(full.shape) * (rank + 1)
        (full * weight).sum().backward()
        assert torch.equal(x.grad, weight[2 * rank:2 * rank + 2]), (rank, x.grad)     # no cross-rank term
        # ---- GradBucketReducer == average of the per-rank gradients, unused parameters included
        mapper = _make_mapper()
        unused = torch.nn.Parameter(torch.ones(3))
        params = list(mapper.parameters()) + [unused]
        red = parallel.GradBucketReducer(params, bucket_bytes=64)      # tiny buckets: several collectives
        assert len(red.buckets) > 2
        for step in range(2):
            red.zero_grad()
            inp = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 * step + rank))
            mapper(inp).square().sum().backward()
            flat = red.finish().clone()
            want = []
            for r in range(world):
                ref = _make_mapper()
                ri = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 * step + r))
                ref(ri).square().sum().backward()
                want.append([p.grad for p in ref.parameters()])
            for i, p in enumerate(mapper.parameters()):
                avg = sum(w[i] for w in want) / world
                assert torch.allclose(p.grad, avg, rtol=1e-6, atol=1e-7), (rank, step, i)
            assert torch.equal(unused.grad, torch.zeros(3))
        # optimizer.zero_grad(set_to_none=True) between steps must not break the views
        for p in params:
            p.grad = None
        red._pending = [c for (_, _, c) in red.buckets]; red._launched = [False] * len(red.buckets); red.flat.zero_()
        mapper(torch.ones(1, 6)).sum().backward()
        red.finish()
        ref = _make_mapper()
        ref(torch.ones(1, 6)).sum().backward()
        for p, q in zip(mapper.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-7)
        red.close()
        torch.save(torch.ones(1), os.path.join(result_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gather_layer_and_mapper_gradient_allreduce(tmp_path):
    """The two collectives of the optimisation loop: autograd all-gather with the own-slice
    backward, and the bucketed all-reduce of the shared mapper's gradients."""
    world, port = 2, 33500 + (os.getpid() % 2000)
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok0.pt").exists() and (tmp_path / "ok1.pt").exists()


def test_grad_bucket_reducer_without_process_group():
    m = _make_mapper()
    red = parallel.allreduce_mapper_grads(m, bucket_bytes=128)
    m(torch.ones(2, 6)).sum().backward()
    flat = red.finish()
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    ref = _make_mapper()
    ref(torch.ones(2, 6)).sum().backward()
    assert all(torch.equal(p.grad, q.grad) for p, q in zip(m.parameters(), ref.parameters()))
    assert parallel.gather_with_grad(flat) is flat
