"""CPU: the C-ABI shared library loads and exports every symbol include/rdb200.h declares, and the
ctypes struct mirrors agree with the C compiler's layout.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from rdb200 import _lib, cdefs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rdb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"librdb200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "python binding table and header disagree"
    assert lib.rd_version() >= 100


def test_struct_layouts_match_the_c_compiler():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rdb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rd_conv_src), sizeof(rd_op_conv), sizeof(rd_op_attn),
         sizeof(rd_op_temb), sizeof(rd_op_inconv), sizeof(rd_op_outhead), sizeof(rd_op), sizeof(rd_sampler_desc));
  printf("%zu %zu %zu\n", sizeof(rd_op_attn_block), sizeof(rd_gto_halo_codec), offsetof(rd_gto_halo_codec, n_triplets));
  printf("%zu %zu %zu %zu %zu\n", offsetof(rd_op_conv, gn_gamma), offsetof(rd_op_conv, out), offsetof(rd_op, u),
         offsetof(rd_op_outhead, score), offsetof(rd_sampler_desc, seed));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(v) for v in out]
    mine = [C.sizeof(t) for t in (cdefs.ConvSrc, cdefs.OpConv, cdefs.OpAttn, cdefs.OpTemb, cdefs.OpInConv,
                                  cdefs.OpOutHead, cdefs.Op, cdefs.SamplerDesc)]
    mine += [C.sizeof(cdefs.OpAttnBlock), C.sizeof(cdefs.GtoHaloCodec), cdefs.GtoHaloCodec.n_triplets.offset]
    mine += [cdefs.OpConv.gn_gamma.offset, cdefs.OpConv.out.offset, cdefs.Op.u.offset, cdefs.OpOutHead.score.offset,
             cdefs.SamplerDesc.seed.offset]
    assert sizes == mine


def test_argument_validation_without_a_gpu():
    lib = _lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.rd_reflect_f32(None, None, 16, None) != 0
    assert b"null" in lib.rd_last_error()
    assert lib.rd_reflect_f32(None, None, 0, None) == 0  # empty input is a no-op
    assert lib.rd_score_hk_f32(None, None, None, 0.1, None, 0, 72, 20, 10, 0.01, None) == 0
    assert lib.rd_philox_normal_f32(None, 6, 1, 0, None) != 0
    p = C.c_void_p()
    assert lib.rd_plan_create(C.byref(p)) == 0 and lib.rd_plan_size(p) == 0
    op = cdefs.Op()
    op.kind = cdefs.RD_OP_CONV  # an all-zero conv op must be refused, not queued
    assert lib.rd_plan_add(p, C.byref(op)) != 0
    assert lib.rd_plan_destroy(p) == 0


def test_conv_tile_geometry():
    """Planner feedback for the GTO-Halo layer shapes: the library must find a tile geometry for each."""
    lib = _lib.lib()
    cases = [  # (C_in sources, H, W, ntaps, C_out, stride, pad)
        ([64], 8, 9, 9, 64, 1, 1), ([128, 64], 8, 9, 9, 64, 1, 1), ([64, 64], 8, 9, 9, 64, 1, 1),
        ([64], 8, 9, 1, 192, 1, 0), ([64], 8, 9, 9, 64, 2, 0), ([128], 4, 4, 9, 128, 1, 1),
        ([128, 128], 4, 4, 9, 128, 1, 1), ([128, 128], 2, 2, 9, 128, 1, 1), ([128], 8, 8, 9, 128, 1, 1),
        ([128, 128], 2, 2, 1, 128, 1, 0)]
    for cs, H, W, ntaps, co, stride, pad in cases:
        op = cdefs.OpConv()
        op.nsrc = len(cs)
        for i, c in enumerate(cs):
            op.src[i].ptr, op.src[i].C, op.src[i].Hs, op.src[i].Ws = 0x1000, c, H, W
        op.H_in, op.W_in, op.pad, op.stride, op.ntaps, op.C_out = H, W, pad, stride, ntaps, co
        op.H_out, op.W_out = ((H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1) if stride == 2 else (H, W)
        op.w, op.bias, op.out, op.B2 = 0x1000, 0x1000, 0x1000, 1000
        smem, grid, rows = C.c_int(), C.c_int(), C.c_int()
        rc = lib.rd_conv_launch_info(C.byref(op), C.byref(smem), C.byref(grid), C.byref(rows))
        assert rc == 0, lib.rd_last_error()
        assert 0 < smem.value <= 227 * 1024 and grid.value >= 1 and rows.value % 2 == 1
