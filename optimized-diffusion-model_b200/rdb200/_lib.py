"""ctypes binding of librdb200.so (include/rdb200.h) -- the only route from Python to the kernels.

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RDB200_LIB selects a measurement build of the same library (rdb200/build.py `defines`); there is still no fallback
LIB_PATH = os.environ.get("RDB200_LIB") or os.path.join(_HERE, "librdb200.so")

_lib: Optional[C.CDLL] = None

c_f32p = C.c_void_p
c_vp = C.c_void_p

# every exported symbol of include/rdb200.h: name -> (restype, argtypes)
SIGNATURES = {
    "rd_last_error": (C.c_char_p, []),
    "rd_version": (C.c_int, []),
    "rd_device_cc": (C.c_int, []),
    "rd_reflect_f32": (C.c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "rd_inside_f32": (C.c_int, [c_vp, c_vp, C.c_size_t, C.c_size_t, c_vp]),
    "rd_score_hk_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_float, c_vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                  C.c_float, c_vp]),
    "rd_score_hk_workspace_bytes": (C.c_size_t, [C.c_size_t]),
    "rd_score_hk_ws_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_float, c_vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                     C.c_float, c_vp, C.c_size_t, c_vp]),
    "rd_philox_normal_f32": (C.c_int, [c_vp, C.c_size_t, C.c_uint64, C.c_uint32, c_vp]),
    "rd_pc_norms": (C.c_int, [c_vp, c_vp, c_vp, C.POINTER(C.c_int), C.c_size_t, C.c_size_t, C.c_uint64,
                              C.c_uint32, c_vp, C.c_size_t, c_vp]),
    "rd_pc_corrector_apply": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_float, c_vp, c_vp, c_vp, C.c_size_t,
                                        C.c_size_t, C.c_uint64, C.c_uint32, c_vp, C.c_size_t, c_vp]),
    "rd_pc_predictor_step": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_float, C.c_float, c_vp, c_vp, C.c_size_t,
                                       C.c_size_t, C.c_uint64, C.c_uint32, c_vp, C.c_size_t, C.c_int, C.c_int, c_vp]),
    "rd_cfg_combine_f32": (C.c_int, [c_vp, c_vp, C.c_float, c_vp, C.c_size_t, C.c_size_t, c_vp]),
    "rd_perturb_reflect_f32": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_size_t, C.c_size_t, c_vp]),
    "rd_dsm_reduce_f32": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_size_t, C.c_size_t, C.c_int, c_vp]),
    "rd_pf_drift_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_float, C.c_float, c_vp, C.c_size_t, C.c_size_t, c_vp]),
    "rd_gto_halo_decode_f32": (C.c_int, [c_vp, c_vp, C.c_size_t, C.c_size_t, c_vp, c_vp]),
    "rd_gto_halo_encode_f32": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_float, c_vp]),
    "rd_resblock": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp]),
    "rd_attn_block": (C.c_int, [c_vp, c_vp]),
    "rd_rk45_stage_f64": (C.c_int, [c_vp, c_vp, C.c_size_t, C.c_int, c_vp, C.c_double, c_vp, c_vp, c_vp]),
    "rd_rk45_error_f64": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t, c_vp, C.c_double, C.c_double, C.c_double, c_vp, C.c_int, c_vp,
                                    c_vp]),
    "rd_plan_create": (C.c_int, [C.POINTER(c_vp)]),
    "rd_plan_add": (C.c_int, [c_vp, c_vp]),
    "rd_plan_size": (C.c_int, [c_vp]),
    "rd_plan_run": (C.c_int, [c_vp, c_vp]),
    "rd_plan_run_range": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "rd_plan_destroy": (C.c_int, [c_vp]),
    "rd_conv_launch_info": (C.c_int, [c_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rd_checksum_f32": (C.c_int, [c_vp, C.c_int, c_vp, c_vp]),
    "rd_conv_geometry": (C.c_int, [c_vp, C.POINTER(C.c_int)]),
    "rd_sampler_create": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "rd_sampler_run": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "rd_sampler_launches_per_iter": (C.c_int, [c_vp]),
    "rd_sampler_destroy": (C.c_int, [c_vp]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
            if args and args[-1] is c_vp and name not in ("rd_plan_add", "rd_plan_create"):
                setattr(handle, name, _device_guarded(fn))   # entry points that launch on a stream
        _lib = handle
    return _lib


_tls = threading.local()


def _device_guarded(fn):
    """CUDA launches go to the calling thread's CURRENT device; the stream argument names the device the caller means.
    `stream_ptr(device)` (evaluated for the call's last argument) leaves that device in a thread-local, and the call
    runs under `torch.cuda.device(it)` when it is not the current one -- so one process may drive several GPUs (the
    library keeps its launch state per device ordinal)."""
    def call(*args):
        dev = getattr(_tls, "device", None)   # device of this thread's most recent stream_ptr() (callers may reuse the pointer)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args)
        with torch.cuda.device(dev):
            return fn(*args)
    call.__name__ = getattr(fn, "__name__", "rd_call")
    return call


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rd_last_error()
        raise RuntimeError(f"librdb200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def stream_ptr(device=None) -> int:
    if device is not None:
        device = torch.device(device)
        _tls.device = device if device.type == "cuda" else None
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """The kernels consume contiguous fp32 CUDA tensors; anything else is an error (no CPU path)."""
    if not torch.is_tensor(t):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the B200 path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
