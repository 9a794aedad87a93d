"""Stand-alone execution of single NCSN++ layers on the B200 kernels: what `ResnetBlockDDPMpp.forward` and
`AttnBlockpp.forward` of the drop-in call (reference models/layerspp.py:198-214 and :80-96).

The sampler never comes through here (it runs the whole network as one plan, rdb200/engine.py); these entry points
exist so that code which uses the reference's layers on their own -- and the layer-level parity tests -- get the same
kernels through `rd_resblock` / `rd_attn_block` (include/rdb200.h).  Tensors are NCHW fp32 at the module boundary like
the reference's; the NHWC bf16 (or fp32) kernel layout is produced by torch memory-format plumbing around the call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import cdefs as D
from ._lib import check, lib, require_cuda_f32, stream_ptr
from .pack import n_slices, pack_1x1, pack_conv3x3, pad_rows

_PREC = {"bf16": (D.RD_PREC_BF16, torch.bfloat16), "fp32": (D.RD_PREC_F32X3, torch.float32)}


def _nhwc(x: torch.Tensor, dtype) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def _nchw(y: torch.Tensor) -> torch.Tensor:
    return y.permute(0, 3, 1, 2).float().contiguous()


def _conv_ops(keep, src, H, W, C_out, weight, bias, prec, dtype, *, ntaps, gn=None, tproj=None, residual=None, out_scale=1.0):
    """rd_op_conv list (one per channel slice) for a stride-1 layer on NHWC `src`; returns (ops, out tensor)."""
    B2, Cin, x3 = src.shape[0], src.shape[3], prec == D.RD_PREC_F32X3
    out = torch.empty((B2, H, W, C_out), dtype=dtype, device=src.device)
    esz = out.element_size()
    bias = bias.detach().float().contiguous()
    keep += [out, bias]
    ops = []
    for n0, n in n_slices(C_out):
        w = (pack_conv3x3(weight[n0:n0 + n].detach().float(), x3) if ntaps == 9 else pack_1x1(weight[:, n0:n0 + n].detach().float(), x3))
        keep.append(w)
        c = D.OpConv()
        c.nsrc = 1
        c.src[0].ptr, c.src[0].C, c.src[0].Hs, c.src[0].Ws = src.data_ptr(), Cin, H, W
        c.H_in, c.W_in, c.pad, c.stride, c.H_out, c.W_out = H, W, (1 if ntaps == 9 else 0), 1, H, W
        c.ntaps, c.C_out, c.out_stride, c.precision = ntaps, n, C_out, prec
        if gn is not None:
            g, b = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
            keep += [g, b]
            c.gn_groups, c.gn_silu, c.gn_eps, c.gn_gamma, c.gn_beta = gn.num_groups, 1, gn.eps, g.data_ptr(), b.data_ptr()
        c.w, c.bias = w.data_ptr(), bias.data_ptr() + 4 * n0
        if tproj is not None:
            c.tproj, c.tproj_stride, c.tproj_off, c.tproj_wrap = tproj.data_ptr(), tproj.shape[1], n0, 0
        if residual is not None:
            c.residual = residual.data_ptr() + esz * n0
        c.out_scale, c.out, c.B2 = out_scale, out.data_ptr() + esz * n0, B2
        ops.append(c)
    return ops, out


@torch.no_grad()
def resblock_forward(block, x: torch.Tensor, temb=None, precision: str = "bf16") -> torch.Tensor:
    """ResnetBlockDDPMpp.forward(x, temb) (layerspp.py:198-214, eval mode: Dropout_0 is the identity)."""
    x = require_cuda_f32(x, "x")
    prec, dtype = _PREC[precision]
    B, Cin, H, W = x.shape
    Cout = block.out_ch
    if Cin % 64 or Cout % 32:
        raise ValueError("the B200 conv kernel needs C_in % 64 == 0 and C_out % 32 == 0")
    keep = []
    xin = _nhwc(x, dtype)
    rs = float(1.0 / np.sqrt(2.0)) if block.skip_rescale else 1.0
    temb_op, tproj = None, None
    if temb is not None:
        temb = require_cuda_f32(temb, "temb")
        tproj = torch.empty((B, Cout), dtype=torch.float32, device=x.device)
        rows = torch.arange(B, dtype=torch.int32, device=x.device)
        dw, db = block.Dense_0.weight.detach().float().contiguous(), block.Dense_0.bias.detach().float().contiguous()
        keep += [tproj, rows, dw, db]
        temb_op = D.OpTemb()
        temb_op.time_table, temb_op.dense_w, temb_op.dense_b, temb_op.out = temb.data_ptr(), dw.data_ptr(), db.data_ptr(), tproj.data_ptr()
        temb_op.row_idx, temb_op.B2, temb_op.temb_dim, temb_op.num_classes, temb_op.n_out_total = rows.data_ptr(), B, temb.shape[1], 0, Cout
    short_ops, short = [], xin
    if Cin != Cout:
        short_ops, short = _conv_ops(keep, xin, H, W, Cout, block.NIN_0.W, block.NIN_0.b, prec, dtype, ntaps=1)
    ops0, h = _conv_ops(keep, xin, H, W, Cout, block.Conv_0.weight, block.Conv_0.bias, prec, dtype, ntaps=9, gn=block.GroupNorm_0,
                        tproj=tproj)
    ops1, out = _conv_ops(keep, h, H, W, Cout, block.Conv_1.weight, block.Conv_1.bias, prec, dtype, ntaps=9, gn=block.GroupNorm_1,
                          residual=short, out_scale=rs)
    n = len(ops0)
    arr = lambda ops: (D.OpConv * n)(*ops)  # noqa: E731
    a_short, a0, a1 = (arr(short_ops) if short_ops else None), arr(ops0), arr(ops1)
    check(lib().rd_resblock(C.byref(temb_op) if temb_op is not None else None, a_short, a0, a1, n, stream_ptr(x.device)),
          "rd_resblock")
    return _nchw(out)


@torch.no_grad()
def attnblock_forward(block, x: torch.Tensor) -> torch.Tensor:
    """AttnBlockpp.forward(x) (layerspp.py:80-96) on the fused kernel (C = 64, T = H*W <= 128)."""
    x = require_cuda_f32(x, "x")
    B, Cc, H, W = x.shape
    if Cc != 64 or H * W > 128:
        raise NotImplementedError("AttnBlockpp.forward stand-alone covers the fused kernel's shapes (C = 64, H*W <= 128); "
                                  "other shapes run inside the network plan (rdb200/engine.py)")
    xin = _nhwc(x, torch.bfloat16)
    out = torch.empty_like(xin)
    f32 = lambda t: t.detach().float()  # noqa: E731
    qkv = torch.cat([f32(block.NIN_0.W), f32(block.NIN_1.W), f32(block.NIN_2.W)], dim=1)
    wq, wp = pad_rows(qkv.t().contiguous(), 72), pad_rows(f32(block.NIN_3.W).t().contiguous(), 72)
    bq = torch.cat([f32(block.NIN_0.b), f32(block.NIN_1.b), f32(block.NIN_2.b)]).contiguous()
    bp = (f32(block.NIN_3.b) + f32(block.NIN_2.b) @ f32(block.NIN_3.W)).contiguous()   # value bias folded (csrc/attn_core.cu)
    g, b = f32(block.GroupNorm_0.weight).contiguous(), f32(block.GroupNorm_0.bias).contiguous()
    a = D.OpAttnBlock()
    a.x, a.out, a.wqkv_t, a.wproj_t = xin.data_ptr(), out.data_ptr(), wq.data_ptr(), wp.data_ptr()
    a.bqkv, a.bproj, a.gamma, a.beta = bq.data_ptr(), bp.data_ptr(), g.data_ptr(), b.data_ptr()
    a.B2, a.T, a.C, a.groups, a.eps = B, H * W, Cc, block.GroupNorm_0.num_groups, block.GroupNorm_0.eps
    a.out_scale = float(1.0 / np.sqrt(2.0)) if block.skip_rescale else 1.0
    check(lib().rd_attn_block(C.byref(a), stream_ptr(x.device)), "rd_attn_block")
    return _nchw(out)
