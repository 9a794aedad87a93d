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
    ws = _hk_workspace(x.device, B)
    check(lib().rd_score_hk_ws_f32(ptr(x), ptr(x_orig), sp, ss, ptr(out), B, D, int(efs), int(refls), float(min_cutoff),
                                   ptr(ws), ws.numel(), stream_ptr(x.device)), "rd_score_hk_ws_f32")
    return out


_HK_WS = {}


def _hk_workspace(device, B: int) -> torch.Tensor:
    """Per-device scratch for the streaming score_hk path (17 bytes per sample), grown on demand and reused: calls
    on one stream are ordered, so the pre-pass of a call cannot overtake the previous call's readers."""
    need = int(lib().rd_score_hk_workspace_bytes(B))
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _HK_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _HK_WS[key] = torch.empty((max(need, 1 << 16),), dtype=torch.uint8, device=device)
    return ws


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


def perturb_reflect(x0: torch.Tensor, z: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """reflect(x0 + std[:, None] * z) -- the forward perturbation of the DSM loss (losses.py:80-82)."""
    x0 = require_cuda_f32(x0, "x0")
    z = require_cuda_f32(z, "z")
    std = require_cuda_f32(std.to(x0.device), "std").reshape(-1)
    if z.shape != x0.shape or std.numel() != x0.shape[0]:
        raise ValueError("perturb_reflect: shape mismatch")
    B = x0.shape[0]
    out = torch.empty_like(x0)
    check(lib().rd_perturb_reflect_f32(ptr(x0), ptr(z), ptr(std), ptr(out), B, x0.numel() // max(B, 1),
                                       stream_ptr(x0.device)), "rd_perturb_reflect_f32")
    return out


def dsm_reduce(score: torch.Tensor, target: torch.Tensor, weight: torch.Tensor, reduce_mean: bool) -> torch.Tensor:
    """[B] per-sample losses: reduce_op(weight[b] * (score - target)^2) (losses.py:86-92)."""
    score = require_cuda_f32(score, "score")
    target = require_cuda_f32(target, "target")
    weight = require_cuda_f32(weight.to(score.device), "weight").reshape(-1)
    if score.shape != target.shape or weight.numel() != score.shape[0]:
        raise ValueError("dsm_reduce: shape mismatch")
    B = score.shape[0]
    out = torch.empty(B, dtype=torch.float32, device=score.device)
    check(lib().rd_dsm_reduce_f32(ptr(score), ptr(target), ptr(weight), ptr(out), B, score.numel() // max(B, 1),
                                  1 if reduce_mean else 0, stream_ptr(score.device)), "rd_dsm_reduce_f32")
    return out


def pf_drift(x: torch.Tensor, score: torch.Tensor, g, moll: float) -> torch.Tensor:
    """Probability-flow drift of the reflected VE SDE times the boundary mollifier (sampling.py:345-383)."""
    x = require_cuda_f32(x, "x")
    score = require_cuda_f32(score, "score")
    if x.shape != score.shape:
        raise ValueError("pf_drift: shape mismatch")
    B = x.shape[0]
    if torch.is_tensor(g):
        g_t = require_cuda_f32(g.reshape(-1).to(x.device), "g")
        if g_t.numel() != B:
            raise ValueError("g must have one entry per sample")
        gp, gs = ptr(g_t), 0.0
    else:
        gp, gs = None, float(g)
    out = torch.empty_like(x)
    check(lib().rd_pf_drift_f32(ptr(x), ptr(score), gp, gs, float(moll), ptr(out), B, x.numel() // max(B, 1),
                                stream_ptr(x.device)), "rd_pf_drift_f32")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# parameter change detection (csrc/weights.cu)
_SEG = 1 << 16


def checksum_segments(tensors):
    """Device segment table [(address, first flat index, count)] over `tensors` (contiguous fp32 CUDA) + result slot."""
    rows, first = [], 0
    for t in tensors:
        n, base = t.numel(), t.data_ptr()
        for o in range(0, n, _SEG):
            rows.append((base + 4 * o, first + o, min(_SEG, n - o)))
        first += n
    dev = tensors[0].device
    return (torch.tensor(rows, dtype=torch.int64).to(dev), torch.zeros((1,), dtype=torch.int64, device=dev))


def checksum(segs: torch.Tensor, out: torch.Tensor) -> int:
    check(lib().rd_checksum_f32(segs.data_ptr(), segs.shape[0], out.data_ptr(), stream_ptr(segs.device)), "rd_checksum_f32")
    return int(out.item())
