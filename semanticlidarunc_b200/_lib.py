"""ctypes binding of libslu.so (include/slu.h).  There is no CPU fallback.

`lib()` raises if the shared library has not been built (run `python -m semanticlidarunc_b200.build`
or `__graft_entry__.build()`); `require_cuda()` raises if no sm_100 device is visible.  Every wrapper
takes torch CUDA tensors purely as buffers (data_ptr + shape) and enqueues on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLU_LIB_PATH", os.path.join(_PKG, "libslu.so"))   # override: kernel experiments only

IN_LOGITS, IN_PROBS, IN_ALPHA = 0, 1, 2
CONF_RAW, CONF_RENORM = 0, 1
MAX_CLASSES, MAX_BINS, MAX_SCANS = 32, 64, 256

_p = C.c_void_p
_i, _i64, _f, _d = C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); mirrors include/slu.h line by line
SIGNATURES = {
    "slu_version": (_i, []),
    "slu_last_error": (C.c_char_p, []),
    "slu_launch_count": (_i64, []),
    "slu_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "slu_reduce_metrics": (_i, [_p, _p, _i, _i, _i, _i64, _i, _i, _f, _i, _i, _i64, _i, _p,
                                _p, _p, _p, _p, _p, _p, _p, _p]),
    "slu_reduce_metrics_direct": (_i, [_p, _p, _i, _i, _i, _i64, _i, _i, _f, _i, _i, _i64, _i, _p,
                                       _p, _p, _p, _p, _p, _p, _p, _p]),
    "slu_debug_reduce_no_single": (_i, [_i]),
    "slu_debug_reduce_no_private": (_i, [_i]),
    "slu_evidential_reduce": (_i, [_p, _p, _p, _i, _i, _i64, _f, _f, _f, _i, _i, _i64, _i, _p,
                                   _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "slu_dirichlet_loss": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _f, _f, _i, _i, _p, _p, _p, _p]),
    "slu_evidential_loss_fused": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _f, _f, _f, _f, _f, _f, _i, _p, _p, _p]),
    "slu_evidential_loss_step": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _f, _f, _f, _f, _f, _f, _i, _p, _p, _p, _p, _p]),
    "slu_debug_no_packed_loss": (_i, [_i]),
    "slu_debug_no_packed_evidential": (_i, [_i]),
    "slu_count_valid": (_i, [_p, _p, _i64, _p, _i, _p, _p]),
    "slu_peer_mailbox_create": (_i, [C.POINTER(C.c_void_p), _p]),
    "slu_peer_mailbox_open": (_i, [_p, C.POINTER(C.c_void_p)]),
    "slu_peer_mailbox_close": (_i, [_p]),
    "slu_peer_mailbox_destroy": (_i, [_p]),
    "slu_peer_mailbox_timeouts": (_i, [_p, C.POINTER(C.c_uint32)]),
    "slu_count_valid_exchange": (_i, [_p, _p, _i64, _p, _i, _p, _i, _i, _d, _p, _p]),
    "slu_peer_allreduce_i64": (_i, [_p, _i, _p, _i, _p, _i, _i, _d, _p, _p]),
    "slu_dirichlet_term": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _i, _f, _f, _p, _p, _p]),
    "slu_evidence_term": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _i, _p, _i, _p, _p, _p]),
    "slu_logit_regularizer": (_i, [_p, _p, _p, _i, _i, _i64, _p, _i, _i, _f, _p, _p, _p]),
    "slu_diag_special": (_i, [_p, _i64, _p, _p]),
    "slu_confusion_ece": (_i, [_p, _p, _p, _i64, _i, _i, _i64, _i, _p, _p, _p, _p]),
    "slu_confusion_ece_i32": (_i, [_p, _p, _p, _i64, _i, _i, _i64, _i, _p, _p, _p, _p]),
    "slu_debug_hist_generic": (_i, [_i]),
    "slu_score_hist": (_i, [_p, _p, _p, _i64, _i, _p, _i, _p, _p]),
    "slu_score_hist_hybrid": (_i, [_p, _p, _p, _i64, _p, _i, _p, _p]),
    "slu_class_score_hist": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p]),
    "slu_project_workspace_bytes": (_i64, [_i64, _i, _i64]),
    "slu_debug_project_exact": (_i, [_i]),
    "slu_diag_fast_atan2": (_i, [_p, _p, _i64, _p, _p]),
    "slu_project_batch": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _d, _d, _i, _p, _p,
                               _p, _p, _p, _p, _p, _p, _p]),
    "slu_project_points": (_i, [_p, _i64, _i, _i, _i, _i, _d, _d, _i, _p, _p, _p, _p, _p, _p, _p]),
    "slu_project_points_bins": (_i, [_p, _i64, _i, _i, _i, _i, _d, _d, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "slu_frame_normals": (_i, [_p, _i, _i, _i, _i64, _f, _p, _p]),
    "slu_organized_planes": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "slu_frame_tensors": (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _p, _p, _p, _p, _p, _p, _i, _p]),
    "slu_backproject": (_i, [_p, _p, _p, _i64, _i, _i64, _p, _p]),
    "slu_stager_create": (_i, [_i, _i64, _i, _i, C.POINTER(C.c_void_p)]),
    "slu_stager_submit": (_i, [_p, C.c_char_p, C.c_char_p, C.POINTER(_i64)]),
    "slu_stager_fetch": (_i, [_p, _i64, _p, _p, _i64, C.POINTER(_i64), C.POINTER(_i), _p]),
    "slu_stager_destroy": (_i, [_p]),
    "slu_diag_read_stream": (_i, [_p, _i64, _p, _p]),
}

_lib = None


class SluError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SluError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. "
                "Run `python -m semanticlidarunc_b200.build` (needs nvcc). There is no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise SluError("no CUDA device is visible; semanticlidarunc_b200 runs its hot path only on a B200 "
                       "(sm_100a) and has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise SluError(f"device {dev} is not a CUDA device")
    return dev


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = lib().slu_last_error().decode(errors="replace")
    if rc == -5:
        raise OSError(f"{what}: {msg}")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (slu error {rc})")
    raise SluError(f"{what}: CUDA error {rc}: {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    """Kernels libslu has launched in this process so far (slu_launch_count)."""
    return int(lib().slu_launch_count())


def device_guard(fn):
    """Run an ops entry point on the device its tensor arguments live on.

    The library launches on the CURRENT device and stream; a tensor on another GPU would otherwise be handed to a
    kernel on the wrong device.  All CUDA tensor arguments must share one device (ValueError otherwise); when that
    device is not the current one the call runs under `torch.cuda.device(dev)`, so `stream_ptr()` is that device's
    current stream."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in (*args, *kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise ValueError(f"{fn.__name__}: tensor arguments live on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def as_buffer(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    """A contiguous CUDA tensor of `dtype` sharing storage with `t` whenever `t` already is one."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise SluError(f"{name} must live on a CUDA device")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def edges_array(edges):
    arr = (C.c_float * len(edges))(*[float(e) for e in edges])
    return arr
