"""Reference state_dict (SURVEY.md App. C key names) -> kernel weight layouts.

Conv / NIN weights become bf16 slabs in the un-swizzled K-major UMMA "B" layout the conv kernel
streams with bulk TMA copies:  [C_in/64][taps][8 (16-byte k-chunks)][C_out][8 channels].
Everything is written IN PLACE into buffers allocated once, so device pointers captured in a CUDA
graph stay valid when the caller swaps weights (EMA copy_to / restore around every sampler call,
reference Benchmark/gto_halo_benchmarking.py:230-239).
"""
from __future__ import annotations

from typing import Dict, List

import torch


MAX_N_PER_LAUNCH = 128  # output channels of one conv launch; wider layers are issued as channel slices (engine._conv)


def n_slices(c_out: int):
    """[(first channel, width)] of the launches that make up a layer with `c_out` output channels."""
    return [(n0, min(MAX_N_PER_LAUNCH, c_out - n0)) for n0 in range(0, c_out, MAX_N_PER_LAUNCH)]


def _planes(t: torch.Tensor, x3: bool) -> torch.Tensor:
    """[chunk, tap, 8, n, 8] fp32 -> bf16 operand image; fp32-class mode: [chunk, tap, 2 (hi|lo), 8, n, 8] with
    w ~= hi + lo (both bf16, round to nearest)."""
    hi = t.to(torch.bfloat16)
    if not x3:
        return hi.contiguous()
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo], dim=2).contiguous()


def pack_conv3x3(w: torch.Tensor, x3: bool = False) -> torch.Tensor:
    """[C_out, C_in, 3, 3] fp32 -> [C_in/64, 9, (2,) 8, C_out, 8] bf16 (tap = dy*3+dx)."""
    co, ci = w.shape[0], w.shape[1]
    assert ci % 64 == 0, "C_in must be a multiple of 64"
    t = w.permute(1, 2, 3, 0).reshape(ci // 64, 8, 8, 9, co)  # [chunk, kc, j, tap, n]
    return _planes(t.permute(0, 3, 1, 4, 2), x3)


def pack_1x1(W: torch.Tensor, x3: bool = False) -> torch.Tensor:
    """NIN weight [C_in, C_out] fp32 -> [C_in/64, 1, (2,) 8, C_out, 8] bf16."""
    ci, co = W.shape
    assert ci % 64 == 0
    t = W.reshape(ci // 64, 8, 8, 1, co)  # [chunk, kc, j, tap, n]
    return _planes(t.permute(0, 3, 1, 4, 2), x3)


def pad_rows(w: torch.Tensor, ld: int) -> torch.Tensor:
    """[rows, cols] fp32 -> [rows, ld] bf16 with zero padding (shared-memory image of the mma.sync B operand)."""
    out = torch.zeros((w.shape[0], ld), dtype=torch.bfloat16, device=w.device)
    out[:, : w.shape[1]] = w.to(torch.bfloat16)
    return out


class PackedWeights:
    """Device-resident packed parameters of one NCSNpp instance."""

    def __init__(self, device, x3: bool = False):
        self.device = device
        self.x3 = bool(x3)  # fp32-class plan: conv / NIN filters as bf16 hi + lo planes
        self.t: Dict[str, torch.Tensor] = {}
        self.dense_offsets: Dict[str, int] = {}
        self.n_dense_out = 0

    def _put(self, name: str, value: torch.Tensor):
        value = value.detach().to(self.device)
        if name in self.t:
            if self.t[name].shape != value.shape or self.t[name].dtype != value.dtype:
                raise RuntimeError(f"packed tensor {name} changed shape/dtype; rebuild the engine")
            self.t[name].copy_(value)
        else:
            self.t[name] = value.contiguous().clone()

    def ptr(self, name: str) -> int:
        return self.t[name].data_ptr()

    def _put_conv(self, name: str, w: torch.Tensor):
        """3x3 filter [C_out, C_in, 3, 3] -> one packed tensor per channel slice: `name.w` (first / only), `name.w.s1`, ..."""
        for i, (n0, n) in enumerate(n_slices(w.shape[0])):
            self._put(name + ".w" + (f".s{i}" if i else ""), pack_conv3x3(w[n0:n0 + n], self.x3))

    def _put_nin(self, name: str, W: torch.Tensor):
        """NIN weight [C_in, C_out], sliced the same way."""
        for i, (n0, n) in enumerate(n_slices(W.shape[1])):
            self._put(name + ".w" + (f".s{i}" if i else ""), pack_1x1(W[:, n0:n0 + n], self.x3))

    @torch.no_grad()
    def update(self, sd: Dict[str, torch.Tensor], res_blocks: List[str], attn_blocks: List[str]):
        f32 = lambda v: v.detach().to(self.device, torch.float32)  # noqa: E731
        for k in ("time_embed.W", "time_mlp.0.weight", "time_mlp.0.bias", "time_mlp.2.weight", "time_mlp.2.bias"):
            self._put(k, f32(sd[k]))
        if "label_emb.weight" in sd:
            self._put("label_emb.weight", f32(sd["label_emb.weight"]))
            self._put("label_emb.bias", f32(sd["label_emb.bias"]))
        self._put("input_conv.weight", f32(sd["input_conv.weight"]))
        self._put("input_conv.bias", f32(sd["input_conv.bias"]))
        self._put("out_conv.weight", f32(sd["out_conv.weight"]))
        self._put("out_conv.bias", f32(sd["out_conv.bias"]))
        self._put("out_norm.weight", f32(sd["out_norm.weight"]))
        self._put("out_norm.bias", f32(sd["out_norm.bias"]))
        dw, db, off = [], [], 0
        for p in res_blocks:
            for gn in ("GroupNorm_0", "GroupNorm_1"):
                self._put(f"{p}.{gn}.weight", f32(sd[f"{p}.{gn}.weight"]))
                self._put(f"{p}.{gn}.bias", f32(sd[f"{p}.{gn}.bias"]))
            for cv in ("Conv_0", "Conv_1"):
                self._put_conv(f"{p}.{cv}", f32(sd[f"{p}.{cv}.weight"]))
                self._put(f"{p}.{cv}.bias", f32(sd[f"{p}.{cv}.bias"]))
            if f"{p}.NIN_0.W" in sd:
                self._put_nin(f"{p}.NIN_0", f32(sd[f"{p}.NIN_0.W"]))
                self._put(f"{p}.NIN_0.bias", f32(sd[f"{p}.NIN_0.b"]))
                if not self.x3 and sd[f"{p}.NIN_0.W"].shape[1] <= MAX_N_PER_LAUNCH:
                    # fused shortcut (csrc/conv_gemm.cu): Conv_1's 3x3 slabs followed by NIN_0's 1x1 slabs, one bias
                    self._put(f"{p}.Conv_1.wf", torch.cat([pack_conv3x3(f32(sd[f"{p}.Conv_1.weight"])).reshape(-1),
                                                            pack_1x1(f32(sd[f"{p}.NIN_0.W"])).reshape(-1)]))
                    self._put(f"{p}.Conv_1.bias_f", f32(sd[f"{p}.Conv_1.bias"]) + f32(sd[f"{p}.NIN_0.b"]))
            if f"{p}.NIN_0.W" not in sd and not self.x3 and sd[f"{p}.Conv_1.weight"].shape[0] <= MAX_N_PER_LAUNCH:
                # identity shortcut as a fused 1x1 with the identity matrix (exact in bf16): the skip connection is added
                # by the tensor core instead of a residual read in the epilogue
                co = sd[f"{p}.Conv_1.weight"].shape[0]
                self._put(f"{p}.Conv_1.wf", torch.cat([pack_conv3x3(f32(sd[f"{p}.Conv_1.weight"])).reshape(-1),
                                                        pack_1x1(torch.eye(co, device=self.device)).reshape(-1)]))
                self._put(f"{p}.Conv_1.bias_f", f32(sd[f"{p}.Conv_1.bias"]))
            w = f32(sd[f"{p}.Dense_0.weight"])
            self.dense_offsets[p] = off
            off += w.shape[0]
            dw.append(w)
            db.append(f32(sd[f"{p}.Dense_0.bias"]))
        self.n_dense_out = off
        self._put("dense.weight", torch.cat(dw, 0))
        self._put("dense.bias", torch.cat(db, 0))
        for p in attn_blocks:
            self._put(f"{p}.GroupNorm_0.weight", f32(sd[f"{p}.GroupNorm_0.weight"]))
            self._put(f"{p}.GroupNorm_0.bias", f32(sd[f"{p}.GroupNorm_0.bias"]))
            qkv = torch.cat([f32(sd[f"{p}.NIN_{j}.W"]) for j in range(3)], dim=1)  # [C, 3C]
            self._put_nin(f"{p}.qkv", qkv)
            self._put(f"{p}.qkv.bias", torch.cat([f32(sd[f"{p}.NIN_{j}.b"]) for j in range(3)]))
            self._put_nin(f"{p}.proj", f32(sd[f"{p}.NIN_3.W"]))
            self._put(f"{p}.proj.bias", f32(sd[f"{p}.NIN_3.b"]))
            # fused attention kernel: weights transposed to [out][in] bf16 rows padded to 72 (conflict-free fragments)
            if qkv.shape[0] == 64:
                self._put(f"{p}.wqkv_t", pad_rows(qkv.t().contiguous(), 72))
                self._put(f"{p}.wproj_t", pad_rows(f32(sd[f"{p}.NIN_3.W"]).t().contiguous(), 72))
            # fused kernel: k / v carry no bias (csrc/attn_core.cu); the value bias reaches the output as b_v @ W_3
            self._put(f"{p}.proj.bias_fused", f32(sd[f"{p}.NIN_3.b"]) + f32(sd[f"{p}.NIN_2.b"]) @ f32(sd[f"{p}.NIN_3.W"]))
        for k in list(sd.keys()):
            if (k.startswith("downsample.") or k.startswith("upsample.")) and k.endswith(".Conv_0.weight"):
                p = k[: -len(".weight")]
                self._put_conv(p, f32(sd[k]))
                self._put(p + ".bias", f32(sd[p + ".bias"]))

    @torch.no_grad()
    def time_rows(self, sigma: torch.Tensor) -> torch.Tensor:
        """time_mlp(fourier(log sigma)) + label_emb.bias for each entry of `sigma` (ncsnpp.py:252-262,
        layerspp.py:26-28), evaluated with torch in fp32 -- batch-invariant inside the sampler, so it is
        tabulated once per (weights, schedule) instead of per step."""
        import numpy as np
        import torch.nn.functional as F
        t = self.t
        s = sigma.to(self.device, torch.float32)
        proj = torch.log(s)[:, None] * t["time_embed.W"][None, :] * 2 * np.pi
        emb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
        h = F.linear(emb, t["time_mlp.0.weight"], t["time_mlp.0.bias"])
        h = F.linear(F.silu(h), t["time_mlp.2.weight"], t["time_mlp.2.bias"])
        if "label_emb.bias" in t:
            h = h + t["label_emb.bias"][None, :]
        return h.contiguous()
