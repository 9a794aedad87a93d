"""CPU: host-side logic of the drop-in -- registries, SDE tables, weight packing layouts, plan
topology bookkeeping, error behaviour without a GPU, and the 2-rank batch sharding (gloo)."""
import os
import sys

import numpy as np
import pytest
import torch

import cube
import sampling
import sde_lib
from models import utils as mutils
from oracle import rd_oracle as O
from rdb200 import pack
from rdb200.engine import NetSpec, spec_from_config
from helpers import load_golden, make_config, oracle_cfg, same_bits


def test_registries_and_errors():
    assert sampling.get_predictor("euler_maruyama") is sampling.ReflectedEulerMaruyamaPredictor
    assert sampling.get_corrector("langevin") is sampling.ReflectedLangevinCorrector
    assert sampling.get_corrector("none") is sampling.NoneCorrector
    assert set(("network", "mean", "none")) <= set(sampling._DENOISERS)
    with pytest.raises(ValueError):
        sampling.register_predictor(name="euler_maruyama")(type("X", (), {}))
    with pytest.raises(KeyError):
        sampling.get_predictor("nope")
    cfg = make_config()
    cfg.sampling.method = "bogus"
    with pytest.raises(ValueError):
        sampling.get_sampling_fn(cfg, sde_lib.RVESDE(0.01, 5, 10), (2, 1, 8, 9), 1e-5, "cpu")
    with pytest.raises(ValueError):
        mutils.register_model(name="ncsnpp")(type("Y", (), {}))
    assert mutils.get_model("ncsnpp").__name__ == "NCSNpp"


def test_sde_tables_match_reference_schedule():
    g = load_golden("schedule.npz")
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    t, sigma, gg = sde.step_tables(1e-5)
    assert same_bits(t, torch.from_numpy(g["t"])) and same_bits(sigma, torch.from_numpy(g["sigma"]))
    assert same_bits(gg, torch.from_numpy(g["g"]))
    # API surface of the reverse SDE
    r = sde.reverse(lambda x, tt: torch.ones_like(x), probability_flow=False)
    x = torch.zeros(3, 1, 2, 2)
    drift, diff = r.sde(x, torch.tensor([1.0, 0.5, 0.1]))
    assert drift.shape == x.shape and torch.allclose(drift[:, 0, 0, 0], -(diff ** 2))
    assert r.N == 1000 and r.T == 1
    f, G = sde.discretize(x, torch.tensor([1.0, 0.5, 0.0]))
    assert float(G[2]) == pytest.approx(float(sde.discrete_sigmas[0]))
    assert sde.prior_sampling((2, 3)).shape == (2, 3) and float(sde.prior_logp(x).abs().sum()) == 0


def test_no_cpu_fallback():
    x = torch.rand(4, 1, 8, 9)
    for fn in (cube.reflect, cube.inside, lambda t: cube.score_hk(t, t, 0.1)):
        with pytest.raises(RuntimeError):
            fn(x)
    model = mutils.create_model(make_config())
    with pytest.raises(RuntimeError):
        model(x, torch.ones(4), class_labels=torch.zeros(4, 1))


def test_module_matches_reference_state_dict_spec():
    model = mutils.create_model(make_config())
    sd = model.state_dict()
    shapes = O.state_dict_shapes(oracle_cfg())
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 6254913 - 64  # time_embed.W is frozen
    spec = spec_from_config(make_config())
    assert len(spec.res_blocks()) == 17 and spec.attn_blocks() == ["down_attn.0", "down_attn.1", "up_attn.6", "up_attn.7", "up_attn.8"]
    # degenerate init pattern of the reference
    assert float(sd["down_blocks.0.Conv_1.weight"].std()) < 1e-5 < float(sd["down_blocks.0.Conv_0.weight"].std())
    assert float(sd["out_conv.weight"].std()) < 1e-5 and float(sd["out_conv.bias"].abs().sum()) == 0
    # 9x9 shipped shape gates attention on image_size only
    assert spec_from_config(make_config(9, 9)).attn_blocks() == spec.attn_blocks()


def test_weight_packing_layout():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(128, 192, 3, 3, generator=g)
    p = pack.pack_conv3x3(w)
    assert p.shape == (3, 9, 8, 128, 8) and p.dtype == torch.bfloat16
    for (n, c, dy, dx) in [(0, 0, 0, 0), (5, 70, 1, 2), (127, 191, 2, 2), (64, 64, 0, 1)]:
        assert float(p[c // 64, dy * 3 + dx, (c % 64) // 8, n, c % 8]) == float(w[n, c, dy, dx].to(torch.bfloat16))
    W = torch.randn(256, 128, generator=g)
    q = pack.pack_1x1(W)
    assert q.shape == (4, 1, 8, 128, 8)
    for (k, n) in [(0, 0), (100, 17), (255, 127)]:
        assert float(q[k // 64, 0, (k % 64) // 8, n, k % 8]) == float(W[k, n].to(torch.bfloat16))


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from rdb200 import dist as rdd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rdd.shard_range(10, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1, 1).repeat(1, 1, 2, 2)
    full = rdd.all_gather_batch(local, 10)
    q.put((rank, lo, hi, full[:, 0, 0, 0].tolist(), rdd.philox_seed_for_rank(123, rank)))
    dist.destroy_process_group()


def test_two_rank_batch_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 5, 5, 10)
    assert res[0][3] == list(map(float, range(10))) == res[1][3]
    assert res[0][4] != res[1][4]


def test_conv_launch_geometry_invariants():
    """rd_conv_launch_info is pure host logic: for every layer shape of the GTO-Halo network the chosen geometry must
    fit the 227 KB shared-memory budget, fill at most one CTA per SM, and stage an odd number of rows (the un-swizzled
    K-major operand relies on it for conflict-free 16-byte row strides)."""
    import ctypes as C
    from rdb200 import _lib, cdefs as D
    lib = _lib.lib()
    shapes = [(64, 0, 64, 8, 9, 9), (128, 64, 64, 8, 9, 9), (64, 64, 64, 8, 9, 9), (64, 0, 128, 4, 4, 9), (128, 0, 128, 4, 4, 9),
              (128, 128, 128, 4, 4, 9), (128, 0, 128, 2, 2, 9), (128, 128, 128, 2, 2, 9), (128, 64, 64, 8, 9, 1),
              (128, 128, 128, 2, 2, 1), (64, 0, 192, 8, 9, 1)]
    for c0, c1, cout, H, W, taps in shapes:
        op = D.OpConv()
        op.nsrc = 2 if c1 else 1
        op.src[0].ptr, op.src[0].C, op.src[0].Hs, op.src[0].Ws = 0x1000, c0, H, W
        if c1:
            op.src[1].ptr, op.src[1].C, op.src[1].Hs, op.src[1].Ws = 0x1000, c1, H, W
        op.H_in, op.W_in, op.H_out, op.W_out = H, W, H, W
        op.pad, op.stride, op.ntaps, op.C_out = (1 if taps == 9 else 0), 1, taps, cout
        if taps == 9:
            op.gn_groups, op.gn_silu, op.gn_eps, op.gn_gamma, op.gn_beta = min((c0 + c1) // 4, 32), 1, 1e-6, 0x1000, 0x1000
        op.w, op.bias, op.out, op.out_scale, op.B2 = 0x1000, 0x1000, 0x1000, 1.0, 16384
        smem, grid, rows = C.c_int(), C.c_int(), C.c_int()
        assert lib.rd_conv_launch_info(C.byref(op), C.byref(smem), C.byref(grid), C.byref(rows)) == 0, lib.rd_last_error()
        assert 0 < smem.value <= 227 * 1024 - 2048
        assert 1 <= grid.value <= 148
        assert rows.value % 2 == 1 and rows.value >= 128
    # stride-2 (Downsample) launches: polyphase layout, four parity planes per operand stage.  The plain-gather transform
    # stages chunk i+1 before it reports chunk i, so these launches need two operand stages; a tile must hold at least one
    # sample's (Ho+1) x (Wo+1) positions
    for c0, cout, H, W in [(64, 64, 8, 9), (128, 128, 4, 4), (64, 64, 9, 9), (64, 64, 12, 10), (256, 128, 16, 16), (512, 128, 8, 8),
                           (512, 128, 4, 4), (64, 64, 6, 7)]:
        op = D.OpConv()
        op.nsrc = 1
        op.src[0].ptr, op.src[0].C, op.src[0].Hs, op.src[0].Ws = 0x1000, c0, H, W
        op.H_in, op.W_in, op.H_out, op.W_out = H, W, (H - 2) // 2 + 1, (W - 2) // 2 + 1
        op.pad, op.stride, op.ntaps, op.C_out = 0, 2, 9, cout
        op.w, op.bias, op.out, op.out_scale, op.B2 = 0x1000, 0x1000, 0x1000, 1.0, 16384
        geom = (C.c_int * 12)()
        assert lib.rd_conv_geometry(C.byref(op), geom) == 0, lib.rd_last_error()
        S, nt, R, n_groups, a_stages, w_res, w_stages, acc, xmode, smem, grid, tmem = list(geom)
        assert 0 < smem <= 227 * 1024 - 2048 and 1 <= grid <= 148 and R % 2 == 1
        assert a_stages >= 2 and xmode == 0
        assert S >= 1 and S * (op.H_out + 1) * (op.W_out + 1) <= nt * 128     # whole samples per group of accumulator tiles
        assert R >= 4 * nt * 128                                               # four parity planes of nt tiles each
        assert n_groups == (16384 + S - 1) // S and tmem in (32, 64, 128, 256, 512)
    bad = D.OpConv()
    bad.nsrc, bad.ntaps = 1, 5
    assert lib.rd_conv_launch_info(C.byref(bad), None, None, None) != 0 and b"ntaps" in lib.rd_last_error()


def test_engine_plan_accounting():
    """The algorithmic FLOP count the bench's roofline uses equals the closed form of the network (valid pixels only):
    the planner must not change it when ops are fused or split."""
    spec = spec_from_config(make_config(8, 8))
    # closed form for nf=64, ch_mult (1,2,2), 2 res blocks, 8x9 -> 4x4 -> 2x2 (SURVEY.md App. B layer table)
    def conv(hw, cin, cout, taps=9):
        return 2.0 * hw * cout * cin * taps
    total = 0.0
    P0, P1, P2 = 72, 16, 4
    total += 2 * (conv(P0, 64, 64) * 2)                                   # down level 0
    total += conv(P1, 64, 64)                                             # downsample.0 (4x4 out)
    total += conv(P1, 64, 128) + conv(P1, 128, 128) + conv(P1, 64, 128, 1) + 2 * conv(P1, 128, 128)   # down level 1
    total += conv(P2, 128, 128)                                           # downsample.1
    total += 2 * 2 * conv(P2, 128, 128) + 2 * 2 * conv(P2, 128, 128)      # down level 2 + mid
    total += 3 * (conv(P2, 256, 128) + conv(P2, 128, 128) + conv(P2, 256, 128, 1))   # up level 2
    total += conv(P1, 128, 128)                                           # upsample.0
    # the reference pushes the last block output of a level twice and never the downsample output
    # (ncsnpp.py:285-292), so all three up blocks of level 1 see 128 + 128 input channels
    total += 3 * (conv(P1, 256, 128) + conv(P1, 128, 128) + conv(P1, 256, 128, 1))
    total += conv(64, 128, 128)                                           # upsample.1 writes 8x8 (ncsnpp.py:319-320 fixes it up later)
    total += (conv(P0, 192, 64) + conv(P0, 64, 64) + conv(P0, 192, 64, 1)) + 2 * (conv(P0, 128, 64) + conv(P0, 64, 64) + conv(P0, 128, 64, 1))
    # the engine needs a GPU to build; the closed form is cross-checked against DESIGN.md's figure (SURVEY.md App. A:
    # 200.74 conv+linear+NIN minus attention NINs 11.8, temb linears 1.15, input/output convs 0.17) here and
    # against Engine.conv_flops_per_sample in tests/test_gpu_network.py
    assert abs(total / 1e6 - 187.63) < 0.01, total / 1e6
    assert spec.levels == 3 and spec.nf == 64


def test_bench_reference_arm_prints_one_json_line():
    """The driver parses bench.py's stdout: exactly one JSON line, whatever the libraries underneath print (the NCCL
    banner goes to stdout by default).  Run the CPU reference arm on a tiny sample and check the contract keys."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-batch", "8", "--ref-grid", "4", "--ref-passes", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    # kind "reference" = the unmodified reference from oracle/_ref (present wherever oracle/fetch_ref.py has run, i.e. the
    # build container and every snapshot shipped from it), "port" = the oracle restatement when it did not travel
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


# ------------------------------------------------------------------------------------------------ round-2 host logic (CPU)
def test_split_bf16_packing_and_channel_slices():
    """fp32-class plan: every packed filter is a bf16 hi plane + a bf16 lo plane with hi + lo == w to 2^-16 of |w| (two
    round-to-nearest bf16 steps); layers wider than 128 output channels are stored as 128-wide slices."""
    import torch
    from rdb200 import pack
    g = torch.Generator().manual_seed(0)
    w = torch.randn(192, 128, 3, 3, generator=g) * 0.05
    p1, p3 = pack.pack_conv3x3(w), pack.pack_conv3x3(w, x3=True)
    assert p1.shape == (2, 9, 8, 192, 8) and p3.shape == (2, 9, 2, 8, 192, 8) and p3.dtype == torch.bfloat16
    assert torch.equal(p3[:, :, 0], p1)
    rec = p3[:, :, 0].float() + p3[:, :, 1].float()
    # un-pack: [chunk, tap, kc, n, j] -> [n, chunk*64 + kc*8 + j, tap]
    back = rec.permute(3, 0, 2, 4, 1).reshape(192, 128, 9).reshape(192, 128, 3, 3)
    assert float((back - w).abs().max()) <= 2.0 ** -16 * float(w.abs().max())
    assert pack.n_slices(64) == [(0, 64)] and pack.n_slices(192) == [(0, 128), (128, 64)] and pack.n_slices(512)[-1] == (384, 128)
    W = torch.randn(256, 128, generator=g)
    assert pack.pack_1x1(W).shape == (4, 1, 8, 128, 8) and pack.pack_1x1(W, x3=True).shape == (4, 1, 2, 8, 128, 8)


def test_precision_selection_and_errors(monkeypatch):
    from rdb200 import engine
    cfg = __import__("helpers").make_config()
    assert engine.precision_from_config(cfg) == "bf16"
    cfg.model.rd_precision = "fp32"
    assert engine.precision_from_config(cfg) == "fp32" and engine.spec_from_config(cfg).precision == "fp32"
    monkeypatch.setenv("RDB200_PRECISION", "bf16")      # the environment overrides the config (unmodified consumer scripts)
    assert engine.precision_from_config(cfg) == "bf16"
    monkeypatch.setenv("RDB200_PRECISION", "fp16")
    with pytest.raises(ValueError):
        engine.precision_from_config(cfg)


def test_device_rk45_tables_match_scipy():
    """The device-side solver must be scipy's RK45: same Butcher tableau, error weights and control constants
    (scipy.integrate._ivp.rk)."""
    from scipy.integrate._ivp import rk
    from rdb200 import ode
    assert np.array_equal(ode.C_NODES, rk.RK45.C) and np.array_equal(ode.B_ROW, rk.RK45.B) and np.array_equal(ode.E_ROW, rk.RK45.E)
    for s, row in enumerate(ode.A_ROWS):
        assert np.array_equal(row, rk.RK45.A[s][:s]), s
    assert (ode.SAFETY, ode.MIN_FACTOR, ode.MAX_FACTOR) == (rk.SAFETY, rk.MIN_FACTOR, rk.MAX_FACTOR)
    assert ode.ERROR_ESTIMATOR_ORDER == rk.RK45.error_estimator_order and rk.RK45.n_stages == 6


def test_reference_copy_recipe_and_shared_bench_config():
    """oracle/fetch_ref.py keeps a byte-for-byte, sha-stamped copy of the reference's hot-path sources outside the history
    (when /root/reference exists, i.e. in the build container); both bench arms print the same `config` object."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle import fetch_ref
    if os.path.isdir(fetch_ref.SRC):
        assert fetch_ref.fetch() and fetch_ref.available() and fetch_ref.verify()
        with open(os.path.join(fetch_ref.DST, "cube.py"), "rb") as a, open(os.path.join(fetch_ref.SRC, "cube.py"), "rb") as b:
            assert a.read() == b.read()
        ign = open(os.path.join(root, ".gitignore")).read()
        assert "oracle/_ref/" in ign
    bench = importlib.import_module("bench")
    a, b = bench.static_config(8192, 1), bench.static_config(8192, 1)
    assert a == b and set(a) == {"workload", "batch_per_gpu", "global_batch", "step", "l2", "parallelism"}


def test_checksum_segment_table_layout():
    """rdb200.ops.checksum_segments: (address, first flat index, count) rows of at most 65536 elements covering every
    tensor exactly once, in order (the table csrc/weights.cu walks)."""
    import torch
    from rdb200 import ops
    ts = [torch.zeros(100000), torch.zeros(7), torch.zeros(65536)]
    segs, out = ops.checksum_segments(ts)
    rows = segs.tolist()
    assert [r[2] for r in rows] == [65536, 100000 - 65536, 7, 65536]
    assert [r[1] for r in rows] == [0, 65536, 100000, 100007]
    assert rows[1][0] == ts[0].data_ptr() + 4 * 65536 and rows[2][0] == ts[1].data_ptr() and out.dtype == torch.int64
