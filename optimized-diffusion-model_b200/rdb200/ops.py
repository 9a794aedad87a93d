"""Tensor-level wrappers of the element-wise kernels (csrc/elementwise.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr


def reflect(x: torch.Tensor) -> torch.Tensor:
    x = require_cuda_f32(x, "x")
    out = torch.empty_like(x)
    check(lib().rd_reflect_f32(ptr(x), ptr(out), x.numel(), stream_ptr(x.device)), "rd_reflect_f32")
    return out


def inside(x: torch.Tensor) -> torch.Tensor:
    x = require_cuda_f32(x, "x")
    B = x.shape[0]
    D = x.numel() // max(B, 1)
    ok = torch.empty(B, dtype=torch.uint8, device=x.device)
    check(lib().rd_inside_f32(ptr(x), ptr(ok), B, D, stream_ptr(x.device)), "rd_inside_f32")
    return ok.bool()


def score_hk(x, x_orig, sigma, efs: int = 20, refls: int = 10, min_cutoff: float = 1e-2) -> torch.Tensor:
    x = require_cuda_f32(x, "x")
    x_orig = require_cuda_f32(x_orig, "x_orig")
    if x.shape != x_orig.shape:
        raise ValueError("x and x_orig must have the same shape")
    B = x.shape[0]
    D = x.numel() // max(B, 1)
    out = torch.empty_like(x)
    if torch.is_tensor(sigma):
        sig = require_cuda_f32(sigma.to(x.device), "sigma").reshape(-1)
        if sig.numel() != B:
            raise ValueError("sigma must have one entry per sample")
        sp, ss = ptr(sig), 0.0
    else:
        sp, ss = None, float(sigma)
    check(lib().rd_score_hk_f32(ptr(x), ptr(x_orig), sp, ss, ptr(out), B, D, int(efs), int(refls),
                                float(min_cutoff), stream_ptr(x.device)), "rd_score_hk_f32")
    return out


def philox_normal(shape, seed: int, draw: int, device) -> torch.Tensor:
    out = torch.empty(shape, dtype=torch.float32, device=device)
    check(lib().rd_philox_normal_f32(ptr(out), out.numel(), seed, draw, stream_ptr(out.device)),
          "rd_philox_normal_f32")
    return out


def cfg_combine(s2: torch.Tensor, weight) -> torch.Tensor:
    s2 = require_cuda_f32(s2, "scores")
    B = s2.shape[0] // 2
    D = s2.numel() // (2 * B)
    out = torch.empty((B,) + tuple(s2.shape[1:]), dtype=torch.float32, device=s2.device)
    if weight is None:
        wp, ws = None, 0.0
    elif isinstance(weight, (float, int)):
        wp, ws = None, float(weight)
    else:
        w = require_cuda_f32(weight.to(s2.device), "weight").reshape(-1)
        wp, ws = ptr(w), 0.0
        s2._rd_keep = w
    check(lib().rd_cfg_combine_f32(ptr(s2), wp, ws, ptr(out), B, D, stream_ptr(s2.device)), "rd_cfg_combine_f32")
    return out


def _scratch(x: torch.Tensor) -> torch.Tensor:
    B = x.shape[0]
    return torch.empty(2 * ((B + 7) // 8), dtype=torch.float32, device=x.device)


def corrector_step(x, grad, noise: Optional[torch.Tensor], snr: float, seed: int = 0, draw_base: int = 0,
                   want_mean: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """One Langevin corrector iteration (sampling.py:222-231).  Returns (x, x_mean, stats[gbar,nbar,eps])."""
    x = require_cuda_f32(x, "x")
    grad = require_cuda_f32(grad, "grad")
    if noise is not None:
        noise = require_cuda_f32(noise, "noise")
    B = x.shape[0]
    D = x.numel() // B
    partial = _scratch(x)
    nblk = C.c_int(0)
    st = stream_ptr(x.device)
    check(lib().rd_pc_norms(ptr(grad), ptr(noise), ptr(partial), C.byref(nblk), B, D, seed, draw_base, None, 0, st),
          "rd_pc_norms")
    x_out = torch.empty_like(x)
    x_mean = torch.empty_like(x) if want_mean else None
    stats = torch.empty(3, dtype=torch.float32, device=x.device)
    check(lib().rd_pc_corrector_apply(ptr(x), ptr(grad), ptr(noise), ptr(partial), nblk.value, float(snr),
                                      ptr(x_out), ptr(x_mean), ptr(stats), B, D, seed, draw_base, None, 0, st),
          "rd_pc_corrector_apply")
    return x_out, x_mean, stats


def predictor_step(x, score, z: Optional[torch.Tensor], g, N: int, seed: int = 0, draw_base: int = 0,
                   want_mean: bool = True):
    """One reflected Euler-Maruyama step (sampling.py:198-207).  `g`: python float (shared by the batch)
    or a [B] tensor of per-sample diffusion coefficients."""
    import numpy as np
    x = require_cuda_f32(x, "x")
    score = require_cuda_f32(score, "score")
    if z is not None:
        z = require_cuda_f32(z, "z")
    B = x.shape[0]
    D = x.numel() // B
    if torch.is_tensor(g):
        g_t = require_cuda_f32(g.reshape(-1).to(x.device), "g")
        if g_t.numel() != B:
            raise ValueError("g must have one entry per sample")
        per_sample = 1
    else:
        g_t = torch.tensor([float(g)], dtype=torch.float32).to(x.device)
        per_sample = 0
    dt = -1.0 / N
    x_out = torch.empty_like(x)
    x_mean = torch.empty_like(x) if want_mean else None
    check(lib().rd_pc_predictor_step(ptr(x), ptr(score), ptr(z), ptr(g_t), float(np.float32(dt)),
                                     float(np.float32(np.sqrt(-dt))), ptr(x_out), ptr(x_mean), B, D, seed,
                                     draw_base, None, 0, 0, per_sample, stream_ptr(x.device)), "rd_pc_predictor_step")
    return x_out, x_mean
