"""GPU parity: cube.reflect / score_hk / inside, the fused predictor / corrector updates, the CFG
combine and the Philox stream, each through the C ABI against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cube
from oracle import rd_oracle as O
from rdb200 import ops
from helpers import load_golden, rel_to_max, same_bits

DEV = "cuda"


def test_reflect_golden_bitwise():
    g = load_golden("reflect.npz")
    out = cube.reflect(torch.from_numpy(g["x"]).to(DEV))
    assert same_bits(out, torch.from_numpy(g["y"]))


def test_reflect_empty_ragged_and_inplace_semantics():
    assert cube.reflect(torch.empty(0, 1, 8, 9, device=DEV)).shape == (0, 1, 8, 9)
    for n in (1, 3, 5, 71, 1025):  # ragged tails around the 128-bit vector path
        x = (torch.rand(n, device=DEV) - 0.5) * 9
        keep = x.clone()
        assert same_bits(cube.reflect(x), O.reflect(x.cpu()))
        assert torch.equal(x, keep)  # the input is never modified (cube.py:47 copies)


def test_reflect_full_size_properties():
    # BASELINE config #2 shape: [2^20, 1, 8, 9]
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn((1 << 20, 1, 8, 9), device=DEV, generator=g) * 3 + 0.5
    r = cube.reflect(x)
    assert bool(cube.inside(r).all())
    assert torch.equal(cube.reflect(r), r)  # idempotent bitwise
    sl = slice(0, 4096)
    assert same_bits(r[sl], O.reflect(x[sl].cpu()))
    assert float((cube.reflect(-x[sl]) - r[sl]).abs().max()) <= 1e-6
    # linearity-free checksum: result only depends on x mod 2 up to rounding of the shift
    assert float((cube.reflect(x[sl] + 2) - r[sl]).abs().max()) <= 4e-6


def test_inside():
    x = torch.rand(33, 1, 8, 9, device=DEV)
    x[3, 0, 2, 2] = 1.0000001
    x[7, 0, 0, 0] = -1e-9
    x[9, 0, 1, 1] = float("nan")
    x[11, 0, 1, 1] = -0.0
    got = cube.inside(x).cpu()
    assert torch.equal(got, O.inside(x.cpu()))
    assert got.sum() == 30


def test_score_hk_golden():
    """north_star: within 1e-6 relative in fp32 -- except the band just above the branch cutoff where
    the reference's own fp32 evaluation is ill-conditioned: there the bar is the
    reference's own fp32-vs-fp64 error."""
    g = load_golden("score_hk.npz")
    for i in range(g["x"].shape[0]):
        x, x0, sg = (torch.from_numpy(g[k][i]).to(DEV) for k in ("x", "x_orig", "sigma"))
        r = cube.score_hk(x, x0, sg).cpu()
        ref32, ref64 = torch.from_numpy(g["ref32"][i]), torch.from_numpy(g["ref64"][i])
        scale = float(ref64.abs().max())
        floor = float((ref32.double() - ref64).abs().max())
        assert float((r - ref32).abs().max()) <= 1e-6 * scale + 2 * floor + 1e-30, i
        assert float((r.double() - ref64).abs().max()) <= 1e-6 * scale + 2 * floor + 1e-30, i
    # python-float sigma (cube.py:173-174); sigma = 0.25 sits in the ill-conditioned band -> floor-based bar
    x, x0 = torch.from_numpy(g["x"][-1]).to(DEV), torch.from_numpy(g["x_orig"][-1]).to(DEV)
    ref = torch.from_numpy(g["ref32_sigma_float_025"])
    r64 = O.score_hk(x.cpu().double(), x0.cpu().double(), 0.25)
    floor = float((ref.double() - r64).abs().max())
    got = cube.score_hk(x, x0, 0.25).cpu()
    assert float((got - ref).abs().max()) <= 1e-6 * float(r64.abs().max()) + 2 * floor


def test_score_hk_vs_oracle_shapes_and_args():
    g = torch.Generator().manual_seed(5)
    for shape in ((5, 1, 9, 9), (3, 7), (130, 1, 8, 9)):  # D % 4 != 0 takes the scalar path
        B = shape[0]
        mean = torch.rand(shape, generator=g)
        sg = torch.exp(torch.rand(B, generator=g) * 6.2 - 4.6)
        x = O.reflect(mean + sg.view((-1,) + (1,) * (len(shape) - 1)) * torch.randn(shape, generator=g))
        for kw in ({}, {"efs": 5}, {"refls": 0}, {"min_cutoff": 0.1}, {"efs": 0}):
            want = O.score_hk(x, mean, sg, **kw)
            got = cube.score_hk(x.to(DEV), mean.to(DEV), sg.to(DEV), **kw).cpu()
            w64 = O.score_hk(x.double(), mean.double(), sg.double(), **kw)
            floor = float((want.double() - w64).abs().max())
            assert float((got - want).abs().max()) <= 2e-6 * float(w64.abs().max()) + 2 * floor + 1e-30, (shape, kw)
    assert cube.score_hk(torch.empty(0, 1, 8, 9, device=DEV), torch.empty(0, 1, 8, 9, device=DEV), 0.1).numel() == 0


def test_pc_steps_golden():
    g, sch = load_golden("pc_steps.npz"), load_golden("schedule.npz")
    x, score, z = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "score", "z"))
    for idx in (0, 300, 700, 998):
        xp, xpm = ops.predictor_step(x, score, z, float(sch["g"][idx]), 1000)
        assert same_bits(xp, torch.from_numpy(g[f"pred_x_{idx}"])) and same_bits(xpm, torch.from_numpy(g[f"pred_mean_{idx}"]))
        gt = torch.full((x.shape[0],), float(sch["g"][idx]), device=DEV)
        xp2, _ = ops.predictor_step(x, score, z, gt, 1000)  # per-sample g path
        assert same_bits(xp2, xp)
    xc, xcm, stats = ops.corrector_step(x, score, z, 0.01)
    _, _, (gn, nn, step) = O.corrector_step(x.cpu(), score.cpu(), z.cpu(), 0.01)
    assert float(stats[0]) == pytest.approx(float(gn), rel=1e-6) and float(stats[1]) == pytest.approx(float(nn), rel=1e-6)
    assert float(stats[2]) == pytest.approx(float(step), rel=2e-6)
    # the batch-mean norms are summed in a different (fixed) order than torch: <= 1 ulp on the step size
    assert float((xc.cpu() - torch.from_numpy(g["corr_x_0"])).abs().max()) <= 5e-7
    assert float((xcm.cpu() - torch.from_numpy(g["corr_mean_0"])).abs().max()) <= 5e-7


def test_pc_steps_ragged_and_domain():
    g = torch.Generator().manual_seed(6)
    for shape in ((3, 1, 9, 9), (1, 1, 8, 9), (257, 1, 8, 9)):
        x = torch.rand(shape, generator=g); s = torch.randn(shape, generator=g) * 5; z = torch.randn(shape, generator=g)
        xp, xpm = ops.predictor_step(x.to(DEV), s.to(DEV), z.to(DEV), 17.63, 1000)
        op, opm = O.predictor_step(x, s, torch.full((shape[0],), 17.63), 1000, z)
        assert same_bits(xp, op) and same_bits(xpm, opm)
        xc, xcm, _ = ops.corrector_step(x.to(DEV), s.to(DEV), z.to(DEV), 0.16)
        oc, ocm, _ = O.corrector_step(x, s, z, 0.16)
        assert float((xc.cpu() - oc).abs().max()) <= 2e-6 and float((xcm.cpu() - ocm).abs().max()) <= 2e-6
        for t in (xp, xpm, xc, xcm):
            assert bool(cube.inside(t).all())


def test_cfg_combine():
    g = torch.Generator().manual_seed(7)
    s2 = torch.randn(10, 1, 8, 9, generator=g)
    for w in (None, 0, 1.5, torch.rand(5, generator=g) * 4):
        want = O.cfg_combine(s2, w)
        got = ops.cfg_combine(s2.to(DEV), w.to(DEV) if torch.is_tensor(w) else w).cpu()
        assert same_bits(got, want)


def test_philox_stream():
    z = ops.philox_normal((1 << 22,), 1234, 0, DEV)
    assert abs(float(z.mean())) < 3e-3 and abs(float(z.std()) - 1) < 3e-3
    assert abs(float((z ** 4).mean()) - 3) < 0.05 and float(z.abs().max()) > 4.5
    assert torch.equal(z, ops.philox_normal((1 << 22,), 1234, 0, DEV))          # reproducible
    assert not torch.equal(z, ops.philox_normal((1 << 22,), 1234, 1, DEV))      # draw index matters
    assert not torch.equal(z, ops.philox_normal((1 << 22,), 1235, 0, DEV))      # seed matters
    # the fused kernels consume exactly this stream: in-kernel noise == dumped tape
    x = torch.rand(64, 1, 8, 9, device=DEV); s = torch.randn(64, 1, 8, 9, device=DEV)
    tape = ops.philox_normal(x.shape, 99, 1, DEV)
    a, _ = ops.predictor_step(x, s, None, 2.0, 1000, seed=99)
    b, _ = ops.predictor_step(x, s, tape, 2.0, 1000)
    assert torch.equal(a, b)
    tape0 = ops.philox_normal(x.shape, 99, 0, DEV)
    c, _, _ = ops.corrector_step(x, s, None, 0.16, seed=99)
    d, _, _ = ops.corrector_step(x, s, tape0, 0.16)
    assert torch.equal(c, d)


def test_corrector_large_batch_paths():
    """Batches large enough for the all-lanes norm kernel (>= 151552 samples) and the bounded partial table: the batch
    means match torch, and in-kernel Philox noise still equals the dumped tape bit for bit."""
    g = torch.Generator(device=DEV).manual_seed(8)
    B = 1 << 18
    x = torch.rand((B, 1, 8, 9), device=DEV, generator=g)
    s = torch.randn((B, 1, 8, 9), device=DEV, generator=g) * 3
    z = torch.randn((B, 1, 8, 9), device=DEV, generator=g)
    xc, xcm, stats = ops.corrector_step(x, s, z, 0.16)
    gn = s.flatten(1).double().norm(dim=1).mean()
    nn = z.flatten(1).double().norm(dim=1).mean()
    assert float(stats[0]) == pytest.approx(float(gn), rel=2e-6) and float(stats[1]) == pytest.approx(float(nn), rel=2e-6)
    step = (0.16 * nn / gn) ** 2 * 2
    want = cube.reflect((x.double() + step * s.double() + torch.sqrt(step * 2) * z.double()).float())
    assert float((xc - want).abs().max()) <= 2e-5 and bool(cube.inside(xc).all()) and bool(cube.inside(xcm).all())
    tape0 = ops.philox_normal(x.shape, 99, 0, DEV)
    c, _, st_c = ops.corrector_step(x, s, None, 0.16, seed=99)
    d, _, st_d = ops.corrector_step(x, s, tape0, 0.16)
    assert torch.equal(c, d) and torch.equal(st_c, st_d)
