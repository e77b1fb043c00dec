"""CUDA-graph capture of the whole latent-optimisation step (BASELINE config 4): mapper forward -> generator forward ->
loss -> backward to the mapper parameters, one graph launch per step.  The step is ~140 kernel launches whose
low-resolution half is shorter than the host time to enqueue it (train step 15.3 ms against 14.3 ms of kernels).
    python tools/graph_train_step.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import mappers  # noqa: E402


def loop_ms(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision="bf16").to(dev).eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    opts = type("O", (), {"no_coarse_mapper": False, "no_medium_mapper": False, "no_fine_mapper": False})()
    mapper = mappers.LevelsMapper(opts).to(dev).train()
    w = torch.randn(batch, gen.n_latent, 512, device=dev)
    gimg = (torch.randn(batch, 3, 1024, 1024, device=dev) / (3 * 1024 * 1024))

    def step(win):
        w_hat = win + 0.1 * mapper(win)
        img, _ = gen([w_hat], input_is_latent=True, randomize_noise=False)
        loss = (img * gimg).sum()
        loss.backward()
        return loss

    def eager():
        for p in mapper.parameters():
            p.grad = None
        return step(w)

    t_eager = loop_ms(eager)
    eager()
    torch.cuda.synchronize()
    ref = [p.grad.detach().clone() for p in mapper.parameters()]

    static_w = w.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for p in mapper.parameters():
                p.grad = None
            step(static_w)
    torch.cuda.current_stream().wait_stream(side)
    gen.assert_ok()
    for p in mapper.parameters():
        p.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = step(static_w)
    t_graph = loop_ms(graph.replay)
    static_w.copy_(w)
    graph.replay()
    torch.cuda.synchronize()
    got = [p.grad.detach().clone() for p in mapper.parameters()]
    same = all(torch.equal(a, b) for a, b in zip(got, ref))
    worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(got, ref))
    print(f"batch {batch}: eager {t_eager:.3f} ms/step ({batch / t_eager * 1e3:.0f} images/s), CUDA graph {t_graph:.3f} ms/step "
          f"({batch / t_graph * 1e3:.0f} images/s); gradients bit-identical: {same} (worst relative difference {worst:.2e}); "
          f"loss {float(static_loss):.6f}")


if __name__ == "__main__":
    main()
