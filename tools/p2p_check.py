"""2-GPU check of parallel.PeerGather against the NCCL all-gather (run under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from where2edit_b200 import parallel
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
shard = torch.full((4, 3, 64, 64), float(rank + 1), device=dev, dtype=torch.bfloat16) + torch.arange(64, device=dev).to(torch.bfloat16)
pg = parallel.PeerGather(shard.shape, shard.dtype, dev)
ref = parallel.gather_images(shard)
for it, slot in enumerate((0, 1, 0, 1)):
    if it < 2:
        pg.push(shard * (it + 1), slot)                      # copy-in + pushes, closed by the per-step barriers
    else:
        pg.own(slot).copy_(shard * (it + 1)); pg.push(None, slot)   # producer wrote its shard in place
    assert torch.equal(pg.result(slot), parallel.gather_images(shard * (it + 1))), "mismatch"   # consumable at once
# timing at image size
big = torch.randn(32, 3, 1024, 1024, device=dev).to(torch.bfloat16)
pg2 = parallel.PeerGather(big.shape, big.dtype, dev)
out = torch.empty((world * 32, 3, 1024, 1024), device=dev, dtype=torch.bfloat16)
for name, fn in (("p2p", lambda: pg2.push(big, 0)), ("nccl", lambda: parallel.gather_images(big, out=out))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, "ms per gather", e0.elapsed_time(e1) / 10, flush=True)
if rank == 0: print("PeerGather OK", flush=True)
dist.destroy_process_group()
