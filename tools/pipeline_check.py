"""Full-size runs of the two caller-facing pipelines around the hot path (BASELINE.json configs 3 and 4):

  edit      original forward (features captured) + edited forward with per-region mask blend at layer 13
            (attention/attention_model.py:548-549), 1024^2, bf16 engine, stylespace inputs
  backward  forward + backward through modconv / upfirdn2d / fused_act (fp32 kernels, autograd), 1024^2,
            gradients w.r.t. the W+ latent from a seeded synthetic dL/dimage

    python tools/pipeline_check.py [edit_batch] [bwd_batch]
Prints one JSON line per pipeline (images/s, ms)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402


def timed(fn, reps):
    for _ in range(3):   # cached weight layouts + allocator warm-up
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    eb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    bb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision="bf16").to(dev).eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    w = torch.randn(eb, gen.n_latent, 512, device=dev)
    mask = torch.rand(eb, 1, 64, 64, device=dev)

    def edit():
        with torch.no_grad():
            img0, _, styles, feats = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
            edited = [s * 1.05 for s in styles]
            img1, _ = gen([edited], input_is_stylespace=True, randomize_noise=False, attention_layer=13,
                          attention_map=mask, feature_map=feats)
        return img0, img1

    ms = timed(edit, 3)
    img0, img1 = edit()
    print(json.dumps({"pipeline": "edit (original + blended edited forward, bf16)", "batch": eb, "ms": round(ms, 2),
                      "edited_images_per_s": round(eb / ms * 1e3, 1),
                      "finite": bool(torch.isfinite(img1).all()), "changed": float((img1 - img0).abs().mean())}))
    del img0, img1
    torch.cuda.empty_cache()

    wb = torch.randn(bb, gen.n_latent, 512, device=dev, requires_grad=True)
    gimg = torch.randn(bb, 3, 1024, 1024, device=dev) / (3 * 1024 * 1024)

    def fwd_bwd():
        wb.grad = None
        img, _ = gen([wb], input_is_latent=True, randomize_noise=False)
        img.backward(gimg)
        return wb.grad

    grads = {}
    for prec, what in (("fp32", "fp32 kernels"), ("bf16", "3x3 convolutions and their dgrad on the tensor cores")):
        gen.set_precision(prec)
        ms = timed(fwd_bwd, 2)
        g = fwd_bwd()
        grads[prec] = g.detach().double().flatten()
        print(json.dumps({"pipeline": f"forward + backward to W+ (autograd, {what})", "batch": bb, "ms": round(ms, 2),
                          "images_per_s": round(bb / ms * 1e3, 2), "grad_finite": bool(torch.isfinite(g).all()),
                          "grad_abs_mean": float(g.abs().mean())}))
    a, b = grads["bf16"], grads["fp32"]
    print(json.dumps({"bf16 vs fp32 gradient": {"cosine": float(torch.dot(a, b) / (a.norm() * b.norm())),
                                               "rel_l2": float((a - b).norm() / b.norm())}}))


if __name__ == "__main__":
    main()
