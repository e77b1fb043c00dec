"""CUDA-graph capture of the no-grad synthesis / edit step.

A 1024^2 forward is ~33 kernel launches plus the host work of building their tensor maps; at small batch that
host time, not the GPU, bounds the step.  `GraphedGenerator` captures one `Generator.forward` (bf16 engine, fixed
noise buffers, no host synchronisation anywhere on the path) with static input / output buffers and replays it:

    fast = GraphedGenerator(gen, [wplus], input_is_latent=True)            # capture (after two warm-up runs)
    image, _ = fast([new_wplus])                                           # copy-in, one graph launch

Inputs may be a W+ / z tensor list or a stylespace list; `attention_map` / `feature_map` (the blended edit forward,
attention/attention_model.py:546-549) are static buffers too and are refreshed by passing new values to the call.
The outputs are the graph's own buffers: they are overwritten by the next replay.
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
