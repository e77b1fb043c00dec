"""CUDA-graph capture of the no-grad synthesis / edit step.

A 1024^2 forward is ~33 kernel launches plus the host work of building their tensor maps; at small batch that
host time, not the GPU, bounds the step.  `GraphedGenerator` captures one `Generator.forward` (bf16 engine, fixed
noise buffers, no host synchronisation anywhere on the path) with static input / output buffers and replays it:

    fast = GraphedGenerator(gen, [wplus], input_is_latent=True)            # capture (after two warm-up runs)
    image, _ = fast([new_wplus])                                           # copy-in, one graph launch

Inputs may be a W+ / z tensor list or a stylespace list; `attention_map` / `feature_map` (the blended edit forward,
attention/attention_model.py:546-549) are static buffers too and are refreshed by passing new values to the call.
The outputs are the graph's own buffers: they are overwritten by the next replay.

`GraphedStep` does the same for a whole optimisation step -- mapper forward, generator forward, loss, `backward()` to the
mapper parameters or latents (attention/run_attention.py:1233-1424, mapper/training/coach.py:81-92): ~140 launches at
1024^2 whose low-resolution half is shorter than the host time to enqueue it.
"""
import torch


def _clone_tree(x):
    if isinstance(x, torch.Tensor):
        return x.detach().clone()
    if isinstance(x, (list, tuple)):
        return [_clone_tree(v) for v in x]
    return x


def _copy_tree(dst, src):
    if isinstance(dst, torch.Tensor):
        if tuple(dst.shape) != tuple(src.shape):
            raise ValueError(f"GraphedGenerator was captured for shape {tuple(dst.shape)}, got {tuple(src.shape)}")
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, list):
        if len(dst) != len(src):
            raise ValueError("GraphedGenerator: input list length changed since capture")
        for d, s in zip(dst, src):
            _copy_tree(d, s)


class GraphedGenerator:
    def __init__(self, gen, styles, attention_map=None, feature_map=None, **kwargs):
        if gen.precision != "bf16":
            raise ValueError("GraphedGenerator captures the bf16 tensor-core engine: Generator(precision='bf16')")
        if kwargs.get("randomize_noise", False) or kwargs.get("noise") is not None:
            raise ValueError("GraphedGenerator needs the fixed noise buffers (randomize_noise=False, noise=None)")
        kwargs["randomize_noise"] = False
        self.gen = gen
        self.kwargs = kwargs
        self.styles = _clone_tree(list(styles))
        self.attention_map = _clone_tree(attention_map)
        self.feature_map = _clone_tree(feature_map)
        dev = gen.input.input.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):   # builds the cached weight layouts / style plan, warms the allocator
                self._forward()
        torch.cuda.current_stream(dev).wait_stream(side)
        gen.assert_ok()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs = self._forward()

    def _forward(self):
        return self.gen(self.styles, attention_map=self.attention_map, feature_map=self.feature_map, **self.kwargs)

    def __call__(self, styles, attention_map=None, feature_map=None):
        _copy_tree(self.styles, list(styles))
        if attention_map is not None:
            _copy_tree(self.attention_map, attention_map)
        if feature_map is not None:
            _copy_tree(self.feature_map, list(feature_map))
        self.graph.replay()
        return self.outputs


class GraphedStep:
    """One forward + backward of the optimisation loop as ONE CUDA-graph launch.

        def step(w):                                   # any callable that ends in backward(); returns tensors to keep
            img, _ = gen([w + 0.1 * mapper(w)], input_is_latent=True, randomize_noise=False)
            loss = criterion(img)
            loss.backward()
            return loss
        fast = GraphedStep(step, [w], params=mapper.parameters())
        loss = fast(w_new)                             # copy-in + graph launch; p.grad of every parameter is refreshed
        optimizer.step()

    `params`: the leaves whose `.grad` the step produces (empty for a no-grad callable, e.g. the cluster-style mapper's
    forward in an edit application: ~250 small launches, host-bound when run eagerly).  Their gradients are cleared before the capture, so
    `backward()` allocates them inside the graph's memory pool and every replay OVERWRITES them (PyTorch's whole-network
    capture recipe): zero_grad() between replays is neither needed nor allowed to free them (use set_to_none=False, or
    none at all).  The frozen bf16 engine path (precision="bf16", fixed noise buffers) has no host synchronisation, which
    is what makes the step capturable; gradients are bit-identical to the eager step (tests/test_engine_features_gpu.py).
    Inputs must keep their shapes; the returned tensors are the graph's own buffers.  One process, one device: a step that
    issues collectives (parallel.GradBucketReducer) is run eagerly -- all-reduce the refreshed `.grad`s after the replay."""

    def __init__(self, fn, example_inputs, params=(), warmup=3):
        self.fn = fn
        self.params = [p for p in params]
        self.inputs = _clone_tree(list(example_inputs))   # (may be empty: a step that closes over its leaves)

        def first_tensor(x):
            if isinstance(x, torch.Tensor):
                return x
            if isinstance(x, (list, tuple)):
                for v in x:
                    t = first_tensor(v)
                    if t is not None:
                        return t
            return None
        anchor = self.params[0] if self.params else first_tensor(self.inputs)
        if anchor is None:
            raise ValueError("GraphedStep: needs `params` or at least one tensor input (to find the device)")
        dev = anchor.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # (the warm-up runs on a side stream, as PyTorch's capture recipe asks; gradient accumulators created by earlier
        # eager steps live on the default stream, which autograd reports once per run -- not an error here)
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        try:
            with torch.cuda.stream(side):
                for _ in range(max(1, int(warmup))):   # cached layouts, allocator and library workspaces warm before the capture
                    self._clear()
                    self.fn(*self.inputs)
            torch.cuda.current_stream(dev).wait_stream(side)
            self._clear()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = self.fn(*self.inputs)
        finally:
            if quiet is not None:
                quiet(True)

    def _clear(self):
        for p in self.params:
            p.grad = None

    def __call__(self, *inputs):
        _copy_tree(self.inputs, list(inputs))
        self.graph.replay()
        return self.outputs
