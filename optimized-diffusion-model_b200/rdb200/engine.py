"""Lowering of the NCSN++ topology to an op plan, and the per-(model, batch, H, W) execution engine.

`NetSpec` mirrors the constructor logic of the reference network (models/ncsnpp.py:42-224):
which ResBlocks exist, their channel counts, where attention / down / up-sampling sit.
`Engine` owns the device buffers (NHWC bf16 activations, fp32 state/score, tables) and the C-side
plan; `Engine.forward` is what `NCSNpp.forward` calls, `Engine.sample` runs the whole
predictor-corrector loop (sampling.py:292-339) natively with one CUDA graph per iteration.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import cdefs as D
from ._lib import check, lib, stream_ptr
from .pack import PackedWeights, n_slices

PRECISIONS = {"bf16": D.RD_PREC_BF16, "fp32": D.RD_PREC_F32X3}


@dataclass
class NetSpec:
    channels: int = 1
    image_size: int = 8
    nf: int = 64
    ch_mult: Sequence[int] = (1, 2, 2)
    num_res_blocks: int = 2
    attn_resolutions: Sequence[int] = (8,)
    num_classes: int = 1
    conditional: bool = True
    skip_rescale: bool = True
    scale_by_sigma: bool = False
    precision: str = "bf16"   # "bf16" | "fp32" (fp32-class: fp32 activations, split-bf16 tensor-core operands)

    @property
    def levels(self) -> int:
        return len(self.ch_mult)

    def has_attn(self, i: int) -> bool:
        # the reference gates attention on image_size (H) only (ncsnpp.py:144,178,206)
        return (self.image_size // (2 ** i)) in self.attn_resolutions

    @property
    def mid_attn(self) -> bool:
        return (self.image_size // (2 ** (self.levels - 1))) in self.attn_resolutions

    def res_blocks(self) -> List[str]:
        n_down = self.levels * self.num_res_blocks
        n_up = self.levels * (self.num_res_blocks + 1)
        return ([f"down_blocks.{i}" for i in range(n_down)] + ["mid_block1", "mid_block2"] +
                [f"up_blocks.{i}" for i in range(n_up)])

    def attn_blocks(self) -> List[str]:
        out, d = [], 0
        for i in range(self.levels):
            for _ in range(self.num_res_blocks):
                if self.has_attn(i):
                    out.append(f"down_attn.{d}")
                d += 1
        if self.mid_attn:
            out.append("mid_attn")
        u = 0
        for i in reversed(range(self.levels)):
            for _ in range(self.num_res_blocks + 1):
                if self.has_attn(i):
                    out.append(f"up_attn.{u}")
                u += 1
        return out


def precision_from_config(config) -> str:
    """`config.model.rd_precision` ("bf16" default | "fp32"); the environment variable RDB200_PRECISION overrides it
    (so an unmodified consumer script can be switched without touching its YAML)."""
    import os
    prec = os.environ.get("RDB200_PRECISION") or getattr(config.model, "rd_precision", "bf16")
    if prec not in PRECISIONS:
        raise ValueError(f"unknown rd_precision {prec!r}: expected one of {sorted(PRECISIONS)}")
    return prec


def spec_from_config(config) -> NetSpec:
    m = config.model
    return NetSpec(precision=precision_from_config(config), channels=m.channels, image_size=m.image_size, nf=m.nf, ch_mult=tuple(m.ch_mult),
                   num_res_blocks=m.num_res_blocks, attn_resolutions=tuple(m.attn_resolutions),
                   num_classes=getattr(m, "num_classes", 1), conditional=m.conditional,
                   skip_rescale=m.skip_rescale, scale_by_sigma=getattr(m, "scale_by_sigma", False))


class _Act:
    """An NHWC activation buffer [B2, H, W, C] (bf16, or fp32 in the fp32-class plan)."""

    def __init__(self, B2, H, W, Cc, device, dtype=torch.bfloat16):
        self.t = torch.empty((B2, H, W, Cc), dtype=dtype, device=device)
        self.H, self.W, self.C = H, W, Cc

    @property
    def ptr(self):
        return self.t.data_ptr()


def _down_hw(h: int, w: int) -> Tuple[int, int]:
    # Downsample: pad (0,1,0,1) then 3x3 stride 2 without padding (layerspp.py:157-159)
    return (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1


class Engine:
    """Plan + buffers for one (weights, B, H, W, cfg) combination."""

    def __init__(self, spec: NetSpec, weights: PackedWeights, B: int, H: int, W: int, cfg: bool, device):
        self.spec, self.w, self.B, self.H, self.W, self.cfg = spec, weights, B, H, W, cfg
        self.device = torch.device(device)
        self.prec = PRECISIONS[spec.precision]
        if bool(weights.x3) != (self.prec == D.RD_PREC_F32X3):
            raise RuntimeError("packed weights and engine precision disagree")
        self.act_dtype = torch.float32 if self.prec == D.RD_PREC_F32X3 else torch.bfloat16
        import os
        self.fuse_shortcut = os.environ.get("RD_FUSE_SHORTCUT", "1") != "0"  # A/B switch (measurement)
        # identity skips are fused (as a 1x1 with the identity matrix) at the image sizes where it measured faster than the
        # epilogue's residual read: RD_FUSE_IDENTITY = "all" | "none" | comma-separated pixel counts
        ident = os.environ.get("RD_FUSE_IDENTITY", "16")
        self.fuse_identity_px = (set(range(1, 1 << 12)) if ident == "all" else set() if ident == "none"
                                 else {int(v) for v in ident.split(",") if v})
        self.B2 = 2 * B if cfg else B
        nc = max(spec.num_classes, 1)
        dev = self.device
        self.x = torch.zeros((B, spec.channels, H, W), dtype=torch.float32, device=dev)
        self.score = torch.zeros((self.B2 if not cfg else B, spec.channels, H, W), dtype=torch.float32, device=dev)
        self.labels2 = torch.zeros((self.B2, nc), dtype=torch.float32, device=dev)
        self.cfg_w = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.step_ctr = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.row_idx = torch.arange(self.B2, dtype=torch.int32, device=dev)
        self.temb_dim = spec.nf * 4
        self.table_rows = 0
        self.time_table: Optional[torch.Tensor] = None
        # CFG: the unconditional half shares label 0 and (inside the sampler) sigma, hence one temb row (index B)
        self.temb_rows = self.B2
        self.tproj = torch.empty((self.B2, weights.n_dense_out), dtype=torch.float32, device=dev)
        self.acts: List[_Act] = []
        self._keep: List[object] = []
        self.plan = C.c_void_p()
        check(lib().rd_plan_create(C.byref(self.plan)), "rd_plan_create")
        self._temb_op_index = 0
        self.op_names: List[str] = []
        self.op_kinds: Dict[str, str] = {}
        self.op_info: Dict[str, dict] = {}  # conv launches: tile geometry + algorithmic flops (profiling feedback)
        self.conv_flops_per_sample = 0.0  # algorithmic 2*M*N*K of the tcgen05 launches, valid pixels only
        self.attn_flops_per_sample = 0.0  # QK^T + PV of the stand-alone attention core launches
        self.tensors: Dict[str, _Act] = {}
        self._sampler = None
        self._built_for_table = None

    # ------------------------------------------------------------------ plan construction
    def _act(self, H, W, Cc) -> _Act:
        a = _Act(self.B2, H, W, Cc, self.device, self.act_dtype)
        self.acts.append(a)
        return a

    def _add(self, op: D.Op, name: str):
        check(lib().rd_plan_add(self.plan, C.byref(op)), f"rd_plan_add({name})")
        self.op_names.append(name)
        if op.kind == D.RD_OP_CONV:
            geom = (C.c_int * 12)()
            check(lib().rd_conv_geometry(C.byref(op.u.conv), geom), "rd_conv_geometry")
            c = op.u.conv
            self.op_info[name] = dict(zip(("S", "n_tiles", "R", "n_groups", "a_stages", "w_resident", "w_stages", "acc_bufs",
                                           "xmode", "smem", "grid", "tmem_cols"), list(geom)),
                                      cin=sum(c.src[i].C for i in range(c.nsrc)), n=c.C_out, taps=c.ntaps, hw=(c.H_in, c.W_in),
                                      flops=2.0 * c.H_out * c.W_out * c.C_out * sum(c.src[i].C for i in range(c.nsrc)) * c.ntaps,
                                      res=bool(c.residual), gn=c.gn_groups > 0)
        self.op_kinds[name] = {D.RD_OP_CONV: "conv", D.RD_OP_ATTN_CORE: "attn", D.RD_OP_TEMB: "temb",
                               D.RD_OP_IN_CONV: "in_conv", D.RD_OP_OUT_HEAD: "out_head", D.RD_OP_ATTN_BLOCK: "attn_block"}[op.kind]

    def _conv(self, name, srcs, H_in, W_in, C_out, wname, bias, *, ntaps=9, pad=1, stride=1, gn=None, silu=1,
              tproj_off=None, residual=None, out_scale=1.0, shortcut_srcs=None) -> _Act:
        if stride == 2:
            Ho, Wo = _down_hw(H_in, W_in)
        else:
            Ho, Wo = H_in, W_in
        out = self._act(Ho, Wo, C_out)
        esz = out.t.element_size()
        # a layer wider than one launch is issued as channel slices: same inputs, the slice's own packed filter, and
        # out / residual / bias / temb offset advanced to the slice's first channel (include/rdb200.h `out_stride`)
        for si, (n0, n) in enumerate(n_slices(C_out)):
            op = D.Op()
            op.kind = D.RD_OP_CONV
            c = op.u.conv
            c.nsrc = len(srcs)
            for i, s in enumerate(srcs):
                c.src[i].ptr, c.src[i].C, c.src[i].Hs, c.src[i].Ws = s.ptr, s.C, s.H, s.W
            c.H_in, c.W_in, c.pad, c.stride, c.H_out, c.W_out = H_in, W_in, pad, stride, Ho, Wo
            c.ntaps, c.C_out, c.out_stride, c.precision = ntaps, n, C_out, self.prec
            if gn is not None:
                cin = sum(s.C for s in srcs)
                c.gn_groups, c.gn_silu, c.gn_eps = min(cin // 4, 32), silu, 1e-6
                c.gn_gamma, c.gn_beta = self.w.ptr(gn + ".weight"), self.w.ptr(gn + ".bias")
            c.w, c.bias = self.w.ptr(wname + (f".s{si}" if si else "")), self.w.ptr(bias) + 4 * n0
            if tproj_off is not None:
                c.tproj, c.tproj_stride, c.tproj_off = self.tproj.data_ptr(), self.w.n_dense_out, tproj_off + n0
                c.tproj_wrap = self.B if self.temb_rows == self.B + 1 else 0
            if residual is not None:
                c.residual = residual.ptr + esz * n0
            if shortcut_srcs is not None:  # raw 1x1 shortcut accumulated into the same tile (wname holds both filters)
                c.sc_nsrc = len(shortcut_srcs)
                for i, s in enumerate(shortcut_srcs):
                    c.sc_src[i].ptr, c.sc_src[i].C, c.sc_src[i].Hs, c.sc_src[i].Ws = s.ptr, s.C, s.H, s.W
            c.out_scale, c.out, c.B2, c.samples_per_cta = out_scale, out.ptr + esz * n0, self.B2, 0
            self._add(op, name if si == 0 else f"{name}#s{si}")
        self.conv_flops_per_sample += 2.0 * Ho * Wo * C_out * sum(s.C for s in srcs) * ntaps
        if shortcut_srcs is not None:
            self.conv_flops_per_sample += 2.0 * Ho * Wo * C_out * sum(s.C for s in shortcut_srcs)
        self.tensors[name] = out
        return out

    def _resblock(self, p: str, srcs: List[_Act], H, W, C_out) -> _Act:
        """ResnetBlockDDPMpp (layerspp.py:198-214) as [NIN shortcut] + conv0 + conv1 launches."""
        cin = sum(s.C for s in srcs)
        rs = float(1.0 / np.sqrt(2.0)) if self.spec.skip_rescale else 1.0
        fuse = ((p + ".Conv_1.wf") in self.w.t and self.fuse_shortcut and min(C_out // 4, 32) * 8 >= C_out
                and all(s.C % 64 == 0 for s in srcs[:-1]) and (cin != C_out or (H * W) in self.fuse_identity_px))
        if fuse:
            # NIN_0 rides in Conv_1's launch: extra raw K-chunks into the same accumulator (no launch, no residual read)
            h = self._conv(p + ".Conv_0", srcs, H, W, C_out, p + ".Conv_0.w", p + ".Conv_0.bias", gn=p + ".GroupNorm_0",
                           tproj_off=self.w.dense_offsets[p])
            try:
                return self._conv(p, [h], H, W, C_out, p + ".Conv_1.wf", p + ".Conv_1.bias_f", gn=p + ".GroupNorm_1",
                                  out_scale=rs, shortcut_srcs=srcs)
            except RuntimeError:  # no tile geometry for the fused form at this size: separate NIN launch + residual
                short = (self._conv(p + ".NIN_0", srcs, H, W, C_out, p + ".NIN_0.w", p + ".NIN_0.bias", ntaps=1, pad=0)
                         if cin != C_out else srcs[0])
                return self._conv(p, [h], H, W, C_out, p + ".Conv_1.w", p + ".Conv_1.bias", gn=p + ".GroupNorm_1",
                                  residual=short, out_scale=rs)
        if cin != C_out:
            short = self._conv(p + ".NIN_0", srcs, H, W, C_out, p + ".NIN_0.w", p + ".NIN_0.bias", ntaps=1, pad=0)
        else:
            assert len(srcs) == 1 and srcs[0].H == H and srcs[0].W == W
            short = srcs[0]
        h = self._conv(p + ".Conv_0", srcs, H, W, C_out, p + ".Conv_0.w", p + ".Conv_0.bias", gn=p + ".GroupNorm_0",
                       tproj_off=self.w.dense_offsets[p])
        return self._conv(p, [h], H, W, C_out, p + ".Conv_1.w", p + ".Conv_1.bias", gn=p + ".GroupNorm_1",
                          residual=short, out_scale=rs)

    def _attn(self, p: str, x: _Act) -> _Act:
        """AttnBlockpp (layerspp.py:80-96).  bf16 plan at the GTO-Halo shapes (C = 64, T <= 128): ONE fused kernel (GN,
        q/k/v, softmax(qk^T)v, projection, skip).  Otherwise (fp32-class plan, or C5's T = 256 / C = 256): GroupNorm +
        q|k|v projection on tcgen05, the flash-style attention core, output projection + skip on tcgen05."""
        rs = float(1.0 / np.sqrt(2.0)) if self.spec.skip_rescale else 1.0
        Cc = x.C
        if self.prec == D.RD_PREC_BF16 and Cc == 64 and x.H * x.W <= 128:
            out = self._act(x.H, x.W, Cc)
            op = D.Op()
            op.kind = D.RD_OP_ATTN_BLOCK
            a = op.u.attn_block
            a.x, a.out = x.ptr, out.ptr
            a.wqkv_t, a.wproj_t = self.w.ptr(p + ".wqkv_t"), self.w.ptr(p + ".wproj_t")
            a.bqkv, a.bproj = self.w.ptr(p + ".qkv.bias"), self.w.ptr(p + ".proj.bias_fused")
            a.gamma, a.beta = self.w.ptr(p + ".GroupNorm_0.weight"), self.w.ptr(p + ".GroupNorm_0.bias")
            a.B2, a.T, a.C, a.groups, a.eps, a.out_scale = self.B2, x.H * x.W, Cc, min(Cc // 4, 32), 1e-6, rs
            self._add(op, p)
            self.tensors[p] = out
            return out
        qkv = self._conv(p + ".qkv", [x], x.H, x.W, 3 * Cc, p + ".qkv.w", p + ".qkv.bias", ntaps=1, pad=0,
                         gn=p + ".GroupNorm_0", silu=0)
        a = self._act(x.H, x.W, Cc)
        op = D.Op()
        op.kind = D.RD_OP_ATTN_CORE
        op.u.attn.qkv, op.u.attn.out, op.u.attn.B2, op.u.attn.T, op.u.attn.C = qkv.ptr, a.ptr, self.B2, x.H * x.W, Cc
        op.u.attn.precision = self.prec
        self._add(op, p + ".core")
        self.tensors[p + ".core"] = a
        self.attn_flops_per_sample += 4.0 * (x.H * x.W) ** 2 * Cc
        return self._conv(p, [a], x.H, x.W, Cc, p + ".proj.w", p + ".proj.bias", ntaps=1, pad=0, residual=x,
                          out_scale=rs)

    def build(self):
        sp, w = self.spec, self.w
        nf, L = sp.nf, sp.levels
        # -- temb (ncsnpp.py:252-262 + every ResBlock's Dense_0)
        op = D.Op()
        op.kind = D.RD_OP_TEMB
        t = op.u.temb
        t.time_table = 0  # patched by _set_time_table
        t.label_w = w.ptr("label_emb.weight") if sp.conditional else None
        t.labels = self.labels2.data_ptr() if sp.conditional else None
        t.dense_w, t.dense_b, t.out = w.ptr("dense.weight"), w.ptr("dense.bias"), self.tproj.data_ptr()
        t.step_ctr, t.row_idx = self.step_ctr.data_ptr(), None
        t.B2, t.temb_dim, t.num_classes, t.n_out_total = self.temb_rows, self.temb_dim, (sp.num_classes if sp.conditional else 0), w.n_dense_out
        self._temb_op = op
        # -- input conv
        op_in = D.Op()
        op_in.kind = D.RD_OP_IN_CONV
        h = self._act(self.H, self.W, nf)
        ic = op_in.u.inconv
        ic.x, ic.w, ic.bias, ic.out = self.x.data_ptr(), w.ptr("input_conv.weight"), w.ptr("input_conv.bias"), h.ptr
        ic.B, ic.B2, ic.C_in, ic.C_out, ic.H, ic.W = self.B, self.B2, sp.channels, nf, self.H, self.W
        ic.precision = self.prec
        self._in_op = op_in
        self._pending = [("temb", self._temb_op), ("input_conv", op_in)]
        self.tensors["input_conv"] = h
        # ops are added lazily in finalize() because the temb op needs its table pointer first
        self._h0 = h

    def finalize(self, time_table: torch.Tensor, use_row_idx: bool):
        """Emit the plan (called once; the time table buffer must keep its address afterwards)."""
        sp, w = self.spec, self.w
        nf, L = sp.nf, sp.levels
        self.time_table = time_table
        self._temb_op.u.temb.time_table = time_table.data_ptr()
        self._temb_op.u.temb.row_idx = self.row_idx.data_ptr() if use_row_idx else None
        for name, op in self._pending:
            self._add(op, name)
        h = self._h0
        H, W = self.H, self.W
        hs: List[_Act] = [h]
        d = 0
        in_ch = nf
        for i in range(L):
            out_ch = nf * sp.ch_mult[i]
            for _ in range(sp.num_res_blocks):
                h = self._resblock(f"down_blocks.{d}", [h], H, W, out_ch)
                in_ch = out_ch
                if sp.has_attn(i):
                    h = self._attn(f"down_attn.{d}", h)
                hs.append(h)
                d += 1
            hs.append(h)
            if i != L - 1:
                h = self._conv(f"downsample.{i}", [h], H, W, in_ch, f"downsample.{i}.Conv_0.w",
                               f"downsample.{i}.Conv_0.bias", pad=0, stride=2)
                H, W = h.H, h.W
        h = self._resblock("mid_block1", [h], H, W, in_ch)
        if sp.mid_attn:
            h = self._attn("mid_attn", h)
        h = self._resblock("mid_block2", [h], H, W, in_ch)
        u = 0
        for j, i in enumerate(reversed(range(L))):
            out_ch = nf * sp.ch_mult[i]
            for _ in range(sp.num_res_blocks + 1):
                skip = hs.pop()
                # ragged fix-up (ncsnpp.py:319-320): h is gathered to the skip's size by nearest mapping
                H, W = skip.H, skip.W
                h = self._resblock(f"up_blocks.{u}", [h, skip], H, W, out_ch)
                if sp.has_attn(i):
                    h = self._attn(f"up_attn.{u}", h)
                u += 1
            if i != 0:
                # Upsample (layerspp.py:122-124): nearest x2 folded into the operand gather
                h = self._conv(f"upsample.{j}", [h], 2 * h.H, 2 * h.W, out_ch, f"upsample.{j}.Conv_0.w",
                               f"upsample.{j}.Conv_0.bias")
        assert h.H == self.H and h.W == self.W, "network output size differs from its input size"
        op = D.Op()
        op.kind = D.RD_OP_OUT_HEAD
        o = op.u.outhead
        o.h, o.gamma, o.beta = h.ptr, w.ptr("out_norm.weight"), w.ptr("out_norm.bias")
        o.w, o.bias = w.ptr("out_conv.weight"), w.ptr("out_conv.bias")
        o.cfg_w, o.cfg_w_scalar, o.score = (self.cfg_w.data_ptr() if self.cfg else None), 0.0, self.score.data_ptr()
        o.B, o.B2, o.C, o.C_img, o.H, o.W = self.B, self.B2, h.C, sp.channels, self.H, self.W
        o.groups, o.cfg, o.eps = min(h.C // 4, 32), (1 if self.cfg else 0), 1e-6
        o.precision = self.prec
        if sp.scale_by_sigma:  # ncsnpp.py:350-351: h / sigma, folded into the output head
            o.sigma_table, o.step_ctr = self._sigma_source()
        self._add(op, "out_head")
        self.n_ops = lib().rd_plan_size(self.plan)

    def _sigma_source(self):
        """(sigma table pointer, step counter pointer or None) for scale_by_sigma; set by the subclasses."""
        raise NotImplementedError

    # ------------------------------------------------------------------ execution
    def run_plan(self):
        check(lib().rd_plan_run(self.plan, stream_ptr(self.device)), "rd_plan_run")

    def run_ops(self, first: int, count: int):
        check(lib().rd_plan_run_range(self.plan, first, count, stream_ptr(self.device)), "rd_plan_run_range")

    def activation(self, name: str) -> torch.Tensor:
        """NCHW fp32 copy of a named intermediate (layer-level parity tests)."""
        return self.tensors[name].t.permute(0, 3, 1, 2).float().contiguous()

    def __del__(self):
        try:
            if self._sampler is not None:
                lib().rd_sampler_destroy(self._sampler)
            if self.plan:
                lib().rd_plan_destroy(self.plan)
        except Exception:
            pass


class ForwardEngine(Engine):
    """Generic `model(x, sigma, labels)` evaluation: per-sample sigma rows, no CFG inside."""

    def __init__(self, spec, weights, B, H, W, device):
        super().__init__(spec, weights, B, H, W, cfg=False, device=device)
        self.sigma_rows = torch.ones((self.B2,), dtype=torch.float32, device=self.device)
        self.build()
        table = torch.zeros((self.B2, self.temb_dim), dtype=torch.float32, device=self.device)
        self.finalize(table, use_row_idx=True)

    def _sigma_source(self):
        return self.sigma_rows.data_ptr(), None  # one sigma per sample

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, sigma: torch.Tensor, labels: Optional[torch.Tensor]) -> torch.Tensor:
        self.x.copy_(x.reshape(self.x.shape))
        self.time_table.copy_(self.w.time_rows(sigma.reshape(-1)))
        self.sigma_rows.copy_(sigma.reshape(-1))
        if self.spec.conditional:
            if labels is None:
                raise TypeError("conditional NCSNpp needs class_labels (the reference fails the same way, "
                                "ncsnpp.py:262)")
            self.labels2.copy_(labels.reshape(self.labels2.shape))
        self.run_plan()
        return self.score.clone()


class SamplerEngine(Engine):
    """Native predictor-corrector sampler with classifier-free guidance (sampling.py:292-339)."""

    def __init__(self, spec, weights, B, H, W, device, sde, eps, snr, n_corrector_steps, cfg=True):
        super().__init__(spec, weights, B, H, W, cfg=cfg, device=device)
        if cfg:
            self.temb_rows = B + 1  # rows 0..B-1 conditional, row B shared by the whole unconditional half
        self.sde, self.eps, self.snr, self.n_corr = sde, eps, snr, n_corrector_steps
        self.N = sde.N
        t, sigma, g = sde.step_tables(eps)
        self.sigma_tab = sigma.to(self.device, torch.float32).contiguous()
        self.g_tab = g.to(self.device, torch.float32).contiguous()
        self.build()
        table = torch.zeros((self.N, self.temb_dim), dtype=torch.float32, device=self.device)
        self.finalize(table, use_row_idx=False)
        self.partial = torch.zeros((2 * ((B + 7) // 8) + 8,), dtype=torch.float32, device=self.device)
        self._tape = None
        self._desc = None
        self.stream = torch.cuda.Stream(device=self.device)

    def _sigma_source(self):
        return self.sigma_tab.data_ptr(), self.step_ctr.data_ptr()  # every sample shares sigma_i inside the sampler

    def refresh_tables(self):
        """Re-tabulate the batch-invariant time embedding after a weight change."""
        self.time_table.copy_(self.w.time_rows(self.sigma_tab))

    def _make_sampler(self, tape: Optional[torch.Tensor], seed: int):
        if self._sampler is not None:
            lib().rd_sampler_destroy(self._sampler)
            self._sampler = None
        d = D.SamplerDesc()
        d.forward, d.x, d.score, d.partial = self.plan, self.x.data_ptr(), self.score.data_ptr(), self.partial.data_ptr()
        d.g_table, d.step_ctr = self.g_tab.data_ptr(), self.step_ctr.data_ptr()
        d.noise_tape = tape.data_ptr() if tape is not None else None
        d.seed = seed
        dt = -1.0 / self.N
        d.snr, d.dt, d.sqrt_dt = float(self.snr), float(np.float32(dt)), float(np.float32(np.sqrt(-dt)))
        d.B, d.D, d.n_corrector_steps = self.B, self.x[0].numel(), self.n_corr
        s = C.c_void_p()
        check(lib().rd_sampler_create(C.byref(d), C.byref(s)), "rd_sampler_create")
        self._sampler, self._tape, self._desc = s, tape, d

    @torch.no_grad()
    def sample(self, x0: torch.Tensor, labels: Optional[torch.Tensor], weight, *, tape: Optional[torch.Tensor] = None,
               seed: int = 0, use_graph: bool = True, n_iter: Optional[int] = None, start_step: int = 0) -> torch.Tensor:
        """Run iterations start_step..N-2 (the last grid point is skipped, sampling.py:330) and return x.
        `start_step` / `n_iter` restrict the range (teacher-forced single-step parity checks)."""
        sp = self.spec
        self.x.copy_(x0.reshape(self.x.shape))
        if sp.conditional and self.cfg:
            self.labels2[: self.B].copy_(labels.reshape(self.B, -1))
            self.labels2[self.B:].zero_()  # unconditional half = label 0 (models/utils.py:123)
        if self.cfg:
            if weight is None:
                self.cfg_w.zero_()
            elif isinstance(weight, (float, int)):
                self.cfg_w.fill_(float(weight))
            else:
                self.cfg_w.copy_(weight.reshape(-1))
        self.step_ctr.fill_(start_step)
        if tape is not None:
            tape = tape.to(self.device, torch.float32).contiguous()
        iters = (self.N - 1 - start_step) if n_iter is None else n_iter
        # stream capture is not allowed on the legacy default stream: the loop runs on the engine's own stream
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self._sampler is None or (tape is not None) or (self._tape is not None) or self._desc.seed != seed:
                self._make_sampler(tape, seed)
                self.run_plan()  # one eager pass: first-use kernel attribute calls must not happen under capture
            check(lib().rd_sampler_run(self._sampler, iters, 1 if use_graph else 0, stream_ptr(self.device)),
                  "rd_sampler_run")
        cur.wait_stream(self.stream)
        return self.x.clone()

    def launches_per_iter(self) -> int:
        return lib().rd_sampler_launches_per_iter(self._sampler) if self._sampler is not None else 0
