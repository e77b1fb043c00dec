"""Per-kernel time of one forward+backward (precision 'bf16' autograd path) at 1024^2, batch 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import where2edit_b200 as w2e
from where2edit_b200 import _native as N

dev = "cuda:0"
torch.manual_seed(0)
gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16").to(dev).eval()
for p in gen.parameters():
    p.requires_grad_(False)
wb = torch.randn(1, gen.n_latent, 512, device=dev, requires_grad=True)
gimg = torch.randn(1, 3, 1024, 1024, device=dev) / (3 * 1024 * 1024)
for rep in range(2):
    wb.grad = None
    N.STATS.trace = [] if rep == 1 else None
    img, _ = gen([wb], input_is_latent=True, randomize_noise=False)
    img.backward(gimg)
    torch.cuda.synchronize()
acc = {}
for call, note, e0, e1 in N.STATS.trace:
    a = acc.setdefault(call, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
N.STATS.trace = None
tot = sum(v[1] for v in acc.values())
for k, (n, t) in sorted(acc.items(), key=lambda x: -x[1][1]):
    print(f"{k:28s} {n:4d} launches {t:8.3f} ms {100 * t / tot:5.1f}%")
print("sum of libw2e kernels", round(tot, 2), "ms")
