"""ctypes binding of libw2e.so (C ABI in include/w2e.h).

There is no CPU fallback and no alternative backend: if the library is missing or a call fails,
a RuntimeError is raised.  Tensors cross the boundary as raw device pointers (`data_ptr()`); the
stream is torch's current CUDA stream, so the calls compose with torch ops and CUDA graphs.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_LRELU = 0, 1

# name -> (restype, argtypes)   -- must mirror include/w2e.h exactly (tests check every symbol)
PROTOTYPES = {
    "w2e_version": (_I, []),
    "w2e_last_error_string": (ctypes.c_char_p, []),
    "w2e_device_info": (_I, [_P, _P, _P]),
    "w2e_upfirdn2d_fwd": (_I, [_P, _P, _P, _L] + [_I] * 12 + [_I, _P]),
    "w2e_upfirdn2d_bwd": (_I, [_P, _P, _P, _L] + [_I] * 12 + [_I, _P]),
    "w2e_bias_act_fwd": (_I, [_P, _P, _P, _P, _I, _P, _L, _I, _L, _F, _F, _I, _P]),
    "w2e_bias_act_bwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _L, _F, _F, _I, _P]),
    "w2e_bias_act_bwd_workspace": (_L, [_L, _I, _L]),
    "w2e_style_demod": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "w2e_style_mod_all": (_I, [_P, _L, _L, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "w2e_style_demod_all": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "w2e_conv_engine_f32": (_I, [_P] * 7 + [_I, _P] + [_I] * 13 + [_P, _I, _I, _I, _P]),
    "w2e_rowdot_f32": (_I, [_P, _P, _P, _P, _P, _L, _L, _P]),
    "w2e_rowdot_segments": (_L, [_L, _L]),
    "w2e_rowdot_seg_f32": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _I, _P]),
    "w2e_torgb_fwd": (_I, [_P] * 7 + [_I, _I, _I, _I, _I, _P]),
    "w2e_torgb_bwd": (_I, [_P] * 6 + [_I, _I, _I, _I, _P]),
    "w2e_torgb_bwd_seg": (_I, [_P] * 7 + [_I, _I, _I, _I, _I, _P]),
    "w2e_mask_blend_fwd": (_I, [_P, _P, _P, _P] + [_I] * 6 + [_I, _P]),
    "w2e_mask_blend_bwd": (_I, [_P] * 7 + [_I] * 6 + [_P]),
    "w2e_modconv_tc_supported": (_I, []),
    "w2e_modconv_tc2": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P] + [_I] * 7 + [_P, _P]),
    "w2e_modconv_tc2_rgb": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P] + [_I] * 6 + [_P] * 6 + [_I, _P, _P]),
    "w2e_modconv_tc2_rgb_pair": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "w2e_modconv_tc2_pair": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "w2e_modconv_tc2_upblur": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P] + [_I] * 6 + [_P, _P]),
    "w2e_modconv_tc2_tf32": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P] + [_I] * 7 + [_P, _P]),
    "w2e_attn_heads_fwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "w2e_attn_heads_bwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "w2e_modconv_tc2_dgrad_up": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "w2e_modconv_tc2_dgrad_up_k": (_I, [_P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "w2e_modconv_tc2_view": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _P, _P]),
    "w2e_grad_assemble_workspace": (_L, [_I, _L, _I]),
    "w2e_grad_assemble_nhwc": (_I, [_P] * 8 + [_I, _P, _P, _I, _P, _P, _P, _I, _L, _I, _P]),
    "w2e_rowdot_nhwc": (_I, [_P, _P, _I, _P, _P, _I, _L, _I, _P]),
    "w2e_sum4_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "w2e_skip_grad": (_I, [_P, _P, _P, _L, _I, _I, _P]),
    "w2e_nchw_to_nhwc_mod": (_I, [_P, _P, _P, _I, _I, _I, _L, _I, _P]),
    "w2e_nhwc_to_nchw_f32": (_I, [_P, _P, _I, _I, _L, _I, _P]),
    "w2e_nchw_class_to_nhwc_mod": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "w2e_nhwc_sum4_to_nchw_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "w2e_blur_act_nhwc": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P] + [_I] * 10 + [_P]),
    "w2e_torgb_nhwc": (_I, [_P] * 7 + [_I, _I, _I, _I, _P]),
    "w2e_blend_nhwc": (_I, [_P] * 6 + [_I] * 6 + [_P]),
    "w2e_cluster_assign": (_I, [_P, _P, _P, _P] + [_I] * 6 + [_P]),
    "w2e_region_mask_fwd": (_I, [_P] * 8 + [_I, _I, _I, _F, _F, _P]),
    "w2e_region_mask_bwd": (_I, [_P] * 8 + [_I, _I, _I, _F, _P]),
    "w2e_box_resample_fwd": (_I, [_P, _P, _L, _I, _I, _I, _I, _P]),
    "w2e_box_resample_bwd": (_I, [_P, _P, _L, _I, _I, _I, _I, _P]),
}

# entry points that enqueue no kernel (host queries)
_HOST_ONLY = {"w2e_version", "w2e_last_error_string", "w2e_device_info", "w2e_bias_act_bwd_workspace", "w2e_rowdot_segments",
              "w2e_modconv_tc_supported", "w2e_grad_assemble_workspace"}


class Tc2Config(ctypes.Structure):
    """include/w2e.h: w2e_tc2_config -- per-call tuning / A-B switches of the tcgen05 convolution (the library
    keeps no global tuning state).  Pass `tc2_cfg(obj)` as the `cfg` argument; None = defaults."""
    _fields_ = [("max_ctas", ctypes.c_int), ("ts_mode", ctypes.c_int), ("flags", ctypes.c_int),
                ("cluster_log2", ctypes.c_int), ("timeline", ctypes.c_void_p)]


def tc2_config(max_ctas=0, ts_mode=1, flags=0, cluster_log2=0, timeline=None):
    return Tc2Config(int(max_ctas), int(ts_mode), int(flags), int(cluster_log2),
                     None if timeline is None else timeline.data_ptr())


def tc2_cfg(cfg):
    return None if cfg is None else ctypes.byref(cfg)


class Stats:
    """Launch accounting of the C-ABI calls: `launches[name]` counts kernel-enqueueing calls; when
    `trace` is a list, every call appends (name, note, start_event, end_event) with CUDA events
    recorded on the current stream around the call (bench.py's per-kernel roofline pass)."""

    def __init__(self):
        self.launches = {}
        self.trace = None
        self.note = None

    def total(self):
        return sum(self.launches.values())

    def reset(self):
        self.launches = {}


STATS = Stats()


def note(**work):
    """Attach algorithmic work (flops=..., bytes=..., tag=...) to the next C-ABI call (tracing only)."""
    if STATS.trace is not None:
        STATS.note = work


_tls = threading.local()   # .device = ordinal of the CUDA tensors of the call being assembled (set by ptr())


def _wrap(name, fn):
    if name in _HOST_ONLY:
        return fn

    def launch(*args):
        st = STATS
        st.launches[name] = st.launches.get(name, 0) + 1
        if st.trace is None:
            return fn(*args)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        st.trace.append((name, st.note, e0, e1))
        st.note = None
        return rc

    def call(*args):
        # device guard: the kernels run on the device of their tensors (ptr() noted it while the arguments were
        # built, stream_ptr() took that device's current stream), whichever device is current in the caller
        dev = getattr(_tls, "device", None)
        _tls.device = None
        if dev is not None and dev != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return launch(*args)
        return launch(*args)

    return call


class _Lib:
    pass

_lib = None
_lock = threading.Lock()


def library_path():
    return _build.LIB


def load():
    """Load (building first if the in-tree library is stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        stale = os.path.exists(path) and os.path.exists(_build.STAMP) and not _build.is_current()
        alt = os.environ.get("W2E_LIB_PATH")   # kernel A/B only (tools/): another build of the same ABI
        if alt:
            path, stale = alt, False
        if not os.path.exists(path) or stale or (os.environ.get("W2E_REBUILD") == "1"):
            # a stale library would be called through mismatched prototypes: rebuild it, or refuse to load it
            try:
                path = _build.build(force=True)
            except RuntimeError as e:
                raise RuntimeError(f"where2edit_b200: {path} is missing or older than csrc/ and cannot be rebuilt "
                                   f"here ({e}); run `python -m where2edit_b200.build`") from e
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:  # fail loudly: there is no other implementation
            raise RuntimeError(f"where2edit_b200: cannot load the CUDA library {path}: {e}") from e
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                if alt:        # an older build under test (tools/): entry points added since are simply absent
                    continue
                raise RuntimeError(f"where2edit_b200: {path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        proxy = _Lib()
        proxy.cdll = lib
        for name in PROTOTYPES:
            if alt and not hasattr(lib, name):
                continue
            setattr(proxy, name, _wrap(name, getattr(lib, name)))
        _lib = proxy
    return _lib


def last_error():
    msg = load().w2e_last_error_string()
    return msg.decode("utf-8", "replace") if msg else ""


ERR_UNSUPPORTED = 3


def check(code, what):
    if code != 0:
        raise RuntimeError(f"libw2e {what} failed (code {code}): {last_error()}")


def stream_ptr(device=None):
    """torch's current stream on the device of the tensors passed through ptr() for this call (else `device`,
    else the current device)."""
    if device is None:
        device = getattr(_tls, "device", None)
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ErrorFlag:
    """Device-side pipeline-timeout flag of the tcgen05 kernels (tc_ptx.cuh mbar_wait: a hung barrier wait sets
    it instead of hanging the GPU) with a pinned host mirror, so that the PUBLIC path notices a failure without a
    per-call synchronisation: `publish()` queues a 4-byte device->host copy behind the kernels of a call,
    `poll()` (start of the next call / of a backward) raises once that copy has landed non-zero, `check()` is
    the synchronising variant.  A raised failure resets the flag."""

    def __init__(self, what):
        self.what = what
        self.dev = None
        self.host = None
        self.event = None

    def tensor(self, device):
        if self.dev is None or self.dev.device != device:
            self.dev = torch.zeros(1, dtype=torch.int32, device=device)
            self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self.event = None
        return self.dev

    def publish(self):
        if self.dev is None or torch.cuda.is_current_stream_capturing():
            return
        self.host.copy_(self.dev, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(self.dev.device))

    def _raise_if_set(self):
        if int(self.host[0]) != 0:
            self.host.zero_()
            self.dev.zero_()
            self.event = None
            raise RuntimeError(f"where2edit_b200: a tcgen05 pipeline wait timed out ({self.what}); the outputs of "
                               "the calls since the last successful check are invalid")

    def poll(self):
        if torch.cuda.is_current_stream_capturing():
            return
        if self.event is not None and self.event.query():
            self._raise_if_set()

    def check(self):
        if self.dev is None:
            return
        self.publish()   # a fresh copy behind everything queued so far
        if self.event is not None:
            self.event.synchronize()
        self._raise_if_set()


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("where2edit_b200 has no CPU path: tensor must live on a CUDA device")
    if not t.is_contiguous():
        raise RuntimeError("where2edit_b200: tensor must be contiguous")
    _tls.device = t.device.index
    return ctypes.c_void_p(t.data_ptr())


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.uint8:   # image outputs only (W2E_U8)
        return U8
    raise RuntimeError(f"where2edit_b200: unsupported dtype {t.dtype} (float32 or bfloat16)")


def host_floats(values):
    arr = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return arr


def host_ptrs(tensors):
    """Host array of device pointers (NULL for None) for the grouped entry points."""
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def host_ints(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "where2edit_b200 has no CPU path (the reference's CPU implementation is only kept as "
                "the test oracle): move the model and its inputs to a CUDA device")
