"""CPU: the oracle restatement (oracle/rd_oracle.py) against the golden vectors produced by the
unmodified reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import rd_oracle as O
from helpers import load_golden, oracle_cfg, rel_to_max, same_bits


def test_reflect_bitwise():
    g = load_golden("reflect.npz")
    assert same_bits(O.reflect(torch.from_numpy(g["x"])), torch.from_numpy(g["y"]))


def test_reflect_properties():
    x = (torch.rand(100000, generator=torch.Generator().manual_seed(1)) - 0.5) * 40
    r = O.reflect(x)
    assert bool(((r >= 0) & (r <= 1)).all())
    assert same_bits(O.reflect(r), r)  # idempotent, bitwise
    # even and 2-periodic only up to an ulp of the intermediate
    assert float((O.reflect(-x) - r).abs().max()) <= 1e-6
    assert float((O.reflect(x + 2) - r).abs().max()) <= 4e-6


def test_score_hk_against_reference():
    g = load_golden("score_hk.npz")
    for i in range(g["x"].shape[0]):
        x, x0, sg = (torch.from_numpy(g[k][i]) for k in ("x", "x_orig", "sigma"))
        r = O.score_hk(x, x0, sg)
        ref32, ref64 = torch.from_numpy(g["ref32"][i]), torch.from_numpy(g["ref64"][i])
        scale = float(ref64.abs().max())
        floor = float((ref32.double() - ref64).abs().max())
        assert float((r - ref32).abs().max()) <= 2e-6 * scale + 2 * floor + 1e-30


def test_score_hk_float_sigma():
    g = load_golden("score_hk.npz")
    x, x0 = torch.from_numpy(g["x"][-1]), torch.from_numpy(g["x_orig"][-1])
    ref = torch.from_numpy(g["ref32_sigma_float_025"])
    assert rel_to_max(O.score_hk(x, x0, 0.25), ref) <= 1e-5


def test_schedule_bitwise():
    g = load_golden("schedule.npz")
    s = O.VESchedule(0.01, 5.0, 1000, 1.0, 1e-5)
    t = s.timesteps()
    assert same_bits(t, torch.from_numpy(g["t"]))
    assert same_bits(s.sigma(t), torch.from_numpy(g["sigma"]))
    assert same_bits(s.diffusion(t), torch.from_numpy(g["g"]))


def test_pc_single_steps_bitwise():
    g = load_golden("pc_steps.npz")
    sch = load_golden("schedule.npz")
    x, score, z = (torch.from_numpy(g[k]) for k in ("x", "score", "z"))
    B = x.shape[0]
    for idx in (0, 300, 700, 998):
        gi = torch.full((B,), float(sch["g"][idx]))
        xp, xpm = O.predictor_step(x, score, gi, 1000, z)
        xc, xcm, _ = O.corrector_step(x, score, z, 0.01)
        assert same_bits(xp, torch.from_numpy(g[f"pred_x_{idx}"])) and same_bits(xpm, torch.from_numpy(g[f"pred_mean_{idx}"]))
        assert same_bits(xc, torch.from_numpy(g[f"corr_x_{idx}"])) and same_bits(xcm, torch.from_numpy(g[f"corr_mean_{idx}"]))


@pytest.mark.parametrize("tag,isz", [("8x9", 8), ("9x9", 9)])
def test_forward_against_reference(tag, isz):
    g = load_golden(f"forward_{tag}.npz")
    cfg = oracle_cfg(isz, isz)
    sd = O.synth_state_dict(cfg, seed=7)
    x, sigma, labels = (torch.from_numpy(g[k]) for k in ("x", "sigma", "labels"))
    taps = {}
    with torch.no_grad():
        y = O.ncsnpp_forward(x, sigma, labels, sd, cfg, taps=taps)
        s = O.guided_score(x, O.VESchedule().sigma(torch.from_numpy(g["t_cfg"])), labels, torch.from_numpy(g["w_cfg"]), sd, cfg)
    assert rel_to_max(y, torch.from_numpy(g["y"])) <= 2e-5
    assert rel_to_max(s, torch.from_numpy(g["score_cfg"])) <= 2e-5
    for k in g.files:
        if k.startswith("tap:"):
            assert rel_to_max(taps[k[4:]], torch.from_numpy(g[k])) <= 2e-5, k


def test_state_dict_spec_counts():
    shapes = O.state_dict_shapes(oracle_cfg(8, 8))
    assert len(shapes) == 261
    assert sum(int(np.prod(s)) for s in shapes.values()) == 6254913
    no_attn = O.state_dict_shapes(O.NetConfig(image_size=8, attn_resolutions=()))
    assert sum(int(np.prod(s)) for s in no_attn.values()) == 6171073


@pytest.mark.parametrize("tag,corrector", [("pc_N30", "langevin"), ("pred_only_N30", "none")])
def test_sampler_tape_replay(tag, corrector):
    g = load_golden(f"sampler_{tag}.npz")
    N, B = int(g["N"]), int(g["B"])
    cfg = oracle_cfg(8, 8)
    sd = O.synth_state_dict(cfg, seed=7)
    n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
    x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=int(g["tape_seed"]))
    labels, w = torch.from_numpy(g["labels"]), float(g["w"])
    with torch.no_grad():
        x = O.pc_sampler(lambda xx, sg: O.guided_score(xx, sg, labels, w, sd, cfg),
                         O.VESchedule(0.01, 5.0, N, 1.0, 1e-5), O.SamplerConfig(corrector=corrector), x0, noise)
    assert bool(O.inside(x).all())
    assert float((x - torch.from_numpy(g["x_final"])).abs().max()) <= 5e-4
    assert int(g["nfe"]) == N * 2  # the reference reports N*(n_steps+1) regardless of the corrector


def test_identities_the_kernels_rely_on_fp64():
    """Two exact reformulations used by the CUDA kernels, checked in fp64 on the oracle (no GPU involved):
    (1) csrc/attn_core.cu drops the key bias (softmax-invariant) and folds the value bias into the output projection
        bias (rdb200/pack.py `proj.bias_fused`);
    (2) csrc/elementwise.cu evaluates score_hk above the reference's branch cutoff with the image sum instead of the
        eigen-series (same function by Poisson summation; the 1e-12 denominator epsilon is rescaled by sqrt(4 pi t))."""
    import numpy as np
    from oracle import rd_oracle as O
    g = torch.Generator().manual_seed(12)
    # ---- (1) attention block with and without k / v biases
    B, C, H, W = 3, 64, 8, 9
    sd = {}
    for j in range(4):
        sd[f"a.NIN_{j}.W"] = torch.randn(C, C, generator=g, dtype=torch.float64) * 0.2
        sd[f"a.NIN_{j}.b"] = torch.randn(C, generator=g, dtype=torch.float64)
    sd["a.GroupNorm_0.weight"] = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    sd["a.GroupNorm_0.bias"] = torch.randn(C, generator=g, dtype=torch.float64)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    ref = O.attnblock(x, sd, "a")
    sd2 = dict(sd)
    sd2["a.NIN_1.b"] = torch.zeros(C, dtype=torch.float64)
    sd2["a.NIN_2.b"] = torch.zeros(C, dtype=torch.float64)
    sd2["a.NIN_3.b"] = sd["a.NIN_3.b"] + sd["a.NIN_2.b"] @ sd["a.NIN_3.W"]
    assert float((O.attnblock(x, sd2, "a") - ref).abs().max()) < 1e-12
    # ---- (2) eigen-series vs image sum of the reflected heat-kernel score, t = sigma^2/2 in (0.01, 0.214]
    n = 4096
    for sigma in (0.1415, 0.2, 0.3, 0.5, 0.65):
        t = sigma ** 2 / 2
        x0 = torch.rand(n, 1, generator=g, dtype=torch.float64)
        xx = O.reflect(x0 + sigma * torch.randn(n, 1, generator=g, dtype=torch.float64))
        sg = torch.full((n,), sigma, dtype=torch.float64)
        ef = O._score_hk_ef(xx, x0, sg ** 2 / 2, 40)
        im = O._score_hk_refl(xx, x0, sg ** 2 / 2, 10)
        scale = float(ef.abs().max())
        assert float((ef - im).abs().max()) <= 1e-9 * scale + 1e-9, sigma
        # the kernel's truncations at this t: Kc modes / Mc image pairs change nothing an fp32 result could see
        Kc = int(np.ceil(np.sqrt(17.5 / (np.pi ** 2 * t)))) + 1
        Mc = int(np.ceil(np.sqrt((70 * t + 1) / 4)))
        assert float((O._score_hk_ef(xx, x0, sg ** 2 / 2, Kc) - ef).abs().max()) <= 3e-7 * scale
        assert float((O._score_hk_refl(xx, x0, sg ** 2 / 2, Mc) - im).abs().max()) <= 3e-7 * scale


def test_round2_goldens_pin_the_oracle():
    """Round-2 fixtures (oracle/make_golden_r2.py, outputs of the unmodified reference): the codec restatements are
    bit-identical, and the oracle's forward reproduces the scale_by_sigma golden (the N=1000 / n_steps_each=2 / C5
    fixtures were bit-identical at generation time, tests/golden/REPORT_r2.txt; re-running them here would cost minutes
    of CPU, they are re-checked on the GPU box through the oracle in tests/test_gpu_round2.py)."""
    g = load_golden("codec.npz")
    dec = O.gto_halo_decode(g["latents"])
    assert dec.shape == g["physical"].shape == (64, 67)
    assert np.array_equal(dec, g["physical"]), float(np.abs(dec - g["physical"]).max())
    img, lab = O.gto_halo_encode(g["raw"], 9)
    assert np.array_equal(img, g["enc_img"]) and np.array_equal(lab, g["enc_label"])
    # padding entries are z-scored too: (0 - mean) / std
    assert abs(float(img[0, 0, 8, 8]) - (0.0 - 0.4652) / 0.1811) < 1e-6
    f = load_golden("forward_sbs.npz")
    cfg = O.NetConfig(image_size=8, attn_resolutions=(8,), scale_by_sigma=True)
    sd = O.synth_state_dict(cfg, seed=9, out_scale=0.003)
    with torch.no_grad():
        y = O.ncsnpp_forward(torch.from_numpy(f["x"]), torch.from_numpy(f["sigma"]), torch.from_numpy(f["labels"]), sd, cfg)
    assert float((y - torch.from_numpy(f["y"])).abs().max()) <= 2e-5 * float(np.abs(f["y"]).max())
    s = load_golden("sampler_pc_N40_ns2.npz")
    N, B = int(s["N"]), int(s["B"])
    cfg8 = O.NetConfig(image_size=8, attn_resolutions=(8,))
    sd8 = O.synth_state_dict(cfg8, seed=7)
    x0, noise = O.make_tape(B, (1, 8, 9), (N - 1) * 3, seed=int(s["tape_seed"]))
    labels = torch.from_numpy(s["labels"])
    with torch.no_grad():
        xo = O.pc_sampler(lambda xx, sg: O.guided_score(xx, sg, labels, float(s["w"]), sd8, cfg8), O.VESchedule(0.01, 5.0, N, 1.0, 1e-5),
                          O.SamplerConfig(n_steps_each=2), x0, noise)
    assert float((xo - torch.from_numpy(s["x_final"])).abs().max()) <= 5e-4
