"""Round-2 GPU parity: the fp32-class plan (split-bf16 tcgen05 operands, fp32 activations), the benchmark's own
1000-step grid, multi-step correctors, scale_by_sigma, BASELINE config C5 and in-place weight swaps -- all against
golden vectors generated from the UNMODIFIED reference (oracle/make_golden_r2.py, tests/golden/REPORT_r2.txt).

Tolerances are the measured errors of the shipped kernels times ~1.3 (profiles/r02_parity.txt), written next to the
north-star figures they are to be read against (1e-3 abs for the bf16 plan, 1e-5 for the fp32 mode).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cube
import sampling
import sde_lib
from models import utils as mutils
from oracle import rd_oracle as O
from helpers import load_golden, make_config, oracle_cfg, patched, rel_to_max

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def build(isz=8, attn=8, corrector="langevin", seed=7, precision=None, **kw):
    cfg = make_config(isz, attn, corrector, precision=precision, **{k: v for k, v in kw.items() if k in ("n_steps_each", "scale_by_sigma")})
    ocfg = oracle_cfg(isz, attn, **{k: v for k, v in kw.items() if k in ("scale_by_sigma",)})
    sd = O.synth_state_dict(ocfg, seed=seed, **{k: v for k, v in kw.items() if k in ("out_scale",)})
    model = mutils.create_model(cfg).to(DEV)
    model.load_state_dict(sd)
    return cfg, ocfg, sd, model.eval()


# ------------------------------------------------------------------------------------------------ fp32-class plan
@pytest.mark.parametrize("tag,isz", [("8x9", 8), ("9x9", 9)])
def test_fp32_mode_forward_vs_reference(tag, isz):
    """north_star: 1e-5 in fp32 mode.  The forward is compared rel-to-max with the reference's fp32 output; what the
    split-bf16 tensor-core product can give is 5e-6 .. 1e-5 per layer (tools/probe_split.cu)."""
    g = load_golden(f"forward_{tag}.npz")
    cfg, ocfg, sd, model = build(isz, isz, precision="fp32")
    assert model.rd_precision == "fp32"
    x, sigma, labels = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "sigma", "labels"))
    with torch.no_grad():
        y = model(x, sigma, class_labels=labels)
    err = rel_to_max(y.cpu(), torch.from_numpy(g["y"]))
    eng = list(model._rd_forward_engines.values())[0]
    taps = {k[4:]: rel_to_max(eng.activation(k[4:]).cpu(), torch.from_numpy(g[k])) for k in g.files
            if k.startswith("tap:") and k[4:] in eng.tensors}
    print(f"fp32-mode forward {tag}: rel-to-max err {err:.3e}; worst tap {max(taps, key=taps.get)} {max(taps.values()):.3e}")
    assert err <= 3e-5 and max(taps.values()) <= 3.5e-5      # measured 1.9e-5 / 2.5e-5 (profiles/r02b_pytest_round2.log)
    # guided score through the drop-in wrappers, per-sample guidance weights up to 4
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    t, w = torch.from_numpy(g["t_cfg"]).to(DEV), torch.from_numpy(g["w_cfg"]).to(DEV)
    with torch.no_grad():
        s = mutils.get_cf_score_fn(sde, model, labels, w)(x, t)
    e = rel_to_max(s.cpu(), torch.from_numpy(g["score_cfg"]))
    print(f"fp32-mode guided score {tag}: rel-to-max err {e:.3e}")
    assert e <= 8e-5                                          # measured 5.8e-5 (guidance weights up to 4 amplify by 1 + 2w)


@pytest.mark.parametrize("tag,corrector", [("pc_N200", "langevin"), ("pc_N1000", "langevin"), ("pred_only_N1000", "none")])
def test_fp32_mode_sampler_vs_reference(tag, corrector):
    g = load_golden(f"sampler_{tag}.npz")
    N, B, w = int(g["N"]), int(g["B"]), float(g["w"])
    cfg, ocfg, sd, model = build(8, 8, corrector, precision="fp32")
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
    x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=int(g["tape_seed"]))
    labels = torch.from_numpy(g["labels"]).to(DEV)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, DEV)
    with patched(torch, "rand", lambda *a, **k: x0.clone()):
        xg, nfe = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(DEV))
    assert bool(cube.inside(xg).all())
    d = (xg.cpu() - torch.from_numpy(g["x_final"])).abs()
    print(f"fp32-mode sampler {tag}: max {float(d.max()):.3e} mean {float(d.mean()):.3e}  (north_star fp32 mode: 1e-5)")
    # measured (profiles/r02b_pytest_round2.log) x 1.5; the reference's own sensitivity to a 1e-6 perturbation of x0 on this
    # kind of model is 4e-5 max / 7e-6 mean at N = 1000 with the corrector, 6e-6 / 1e-6 without
    bar = {"pc_N200": (1.0e-4, 3.2e-5), "pc_N1000": (2.8e-4, 7.7e-5), "pred_only_N1000": (3.0e-5, 5.9e-6)}[tag]
    assert float(d.max()) <= bar[0] and float(d.mean()) <= bar[1]


# ------------------------------------------------------------------------------------------------ bf16 plan, N = 1000
@pytest.mark.parametrize("tag,corrector", [("pc_N1000", "langevin"), ("pred_only_N1000", "none")])
def test_bf16_sampler_N1000_vs_reference(tag, corrector):
    """The headline configuration's grid (N = 1000), final samples against the reference's, next to PyTorch's own
    bf16-autocast run of the same model on the same tape."""
    g = load_golden(f"sampler_{tag}.npz")
    N, B, w = int(g["N"]), int(g["B"]), float(g["w"])
    cfg, ocfg, sd, model = build(8, 8, corrector)
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    sched, scfg = O.VESchedule(0.01, 5.0, N, 1.0, 1e-5), O.SamplerConfig(corrector=corrector)
    n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
    x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=int(g["tape_seed"]))
    labels = torch.from_numpy(g["labels"]).to(DEV)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, DEV)
    with patched(torch, "rand", lambda *a, **k: x0.clone()):
        xg, nfe = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(DEV))
    assert nfe == 2 * N and bool(cube.inside(xg).all())
    ref = torch.from_numpy(g["x_final"]).to(DEV)
    with torch.no_grad():
        def bf16_score(xx, sg):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return O.guided_score(xx, sg, labels, w, sdg, ocfg).float()
        xb = O.pc_sampler(bf16_score, sched, scfg, x0.to(DEV), noise.to(DEV))
    fl = (xb - ref).abs()
    d = (xg - ref).abs()
    print(f"bf16 sampler {tag}: max {float(d.max()):.3e} mean {float(d.mean()):.3e} | torch bf16-autocast floor max "
          f"{float(fl.max()):.3e} mean {float(fl.mean()):.3e}  (north_star bf16: 1e-3)")
    assert float(d.mean()) <= 1.1 * float(fl.mean()) + 1e-4     # measured 0.87-0.94 x the floor (profiles/r02_pytest_gpu.log)
    assert float(d.max()) <= 1.1 * float(fl.max()) + 1e-3


# ------------------------------------------------------------------------------------------------ multi-step corrector
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_multi_step_corrector_native(precision):
    """n_steps_each = 2 (sampling.py:221 loops n_steps Langevin moves per grid point) runs on the native engine."""
    g = load_golden("sampler_pc_N40_ns2.npz")
    N, B, w = int(g["N"]), int(g["B"]), float(g["w"])
    cfg, ocfg, sd, model = build(8, 8, "langevin", precision=precision, n_steps_each=2)
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    x0, noise = O.make_tape(B, (1, 8, 9), (N - 1) * 3, seed=int(g["tape_seed"]))
    labels = torch.from_numpy(g["labels"]).to(DEV)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, DEV)
    with patched(torch, "rand", lambda *a, **k: x0.clone()):
        xn, nfe = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(DEV))
        xl, _ = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(DEV), rd_native=False)
        xp, _ = fn(model, weight=w, class_labels=labels, rd_seed=3)
        xp2, _ = fn(model, weight=w, class_labels=labels, rd_seed=3)
    assert nfe == int(g["nfe"]) == 3 * N
    assert len(model._rd_sampler_engines) == 1 and list(model._rd_sampler_engines.values())[0].n_corr == 2
    assert bool(cube.inside(xn).all()) and bool(cube.inside(xp).all()) and torch.equal(xp, xp2)
    d = (xn.cpu() - torch.from_numpy(g["x_final"])).abs()
    dl = (xn - xl).abs()
    print(f"{precision} n_steps_each=2: native vs reference max {float(d.max()):.3e} mean {float(d.mean()):.3e}; "
          f"native vs generic loop max {float(dl.max()):.3e}")
    if precision == "fp32":
        assert float(d.max()) <= 3e-4 and float(d.mean()) <= 7.5e-5   # measured 2.0e-4 / 4.9e-5
    else:
        # the mean is the stable statistic; the max over 4 x 72 values of a chaotic 40-step trajectory moved between 3.6e-2, 5.2e-2
        # and 6.2e-2 across kernel revisions whose mean stayed at 1.0e-2
        assert float(d.mean()) <= 1.5e-2 and float(d.max()) <= 9e-2


# ------------------------------------------------------------------------------------------------ scale_by_sigma
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_scale_by_sigma(precision):
    """model.scale_by_sigma = True divides the network output by sigma (ncsnpp.py:350-351): generic forward, guided
    score, and the NATIVE sampler loop (round 1 silently ignored the flag there)."""
    g = load_golden("forward_sbs.npz")
    cfg, ocfg, sd, model = build(8, 8, precision=precision, seed=9, scale_by_sigma=True, out_scale=0.003)
    x, sigma, labels = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "sigma", "labels"))
    with torch.no_grad():
        y = model(x, sigma, class_labels=labels)
    ref = torch.from_numpy(g["y"])
    per_sample = ((y.cpu() - ref).abs().amax(dim=(1, 2, 3)) / ref.abs().amax(dim=(1, 2, 3)))
    print(f"{precision} scale_by_sigma forward: per-sample rel-to-max err {per_sample.tolist()}")
    assert float(per_sample.max()) <= (5e-6 if precision == "fp32" else 3e-3)   # measured 2.0e-6 / 1.7e-3
    gs = load_golden("sampler_pc_N30_sbs.npz")
    N, B, w = int(gs["N"]), int(gs["B"]), float(gs["w"])
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    x0, noise = O.make_tape(B, (1, 8, 9), (N - 1) * 2, seed=int(gs["tape_seed"]))
    lab = torch.from_numpy(gs["labels"]).to(DEV)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, DEV)
    with patched(torch, "rand", lambda *a, **k: x0.clone()):
        xn, _ = fn(model, weight=w, class_labels=lab, rd_tape=noise.to(DEV))
        xl, _ = fn(model, weight=w, class_labels=lab, rd_tape=noise.to(DEV), rd_native=False)
    assert len(model._rd_sampler_engines) == 1, "the native engine must take scale_by_sigma models"
    d = (xn.cpu() - torch.from_numpy(gs["x_final"])).abs()
    print(f"{precision} scale_by_sigma sampler N30: native vs reference max {float(d.max()):.3e} mean {float(d.mean()):.3e}; "
          f"native vs generic loop max {float((xn - xl).abs().max()):.3e}")
    assert bool(cube.inside(xn).all())
    assert float(d.mean()) <= (1.5e-6 if precision == "fp32" else 7e-4)          # measured 4.7e-7 / 3.4e-4
    assert float((xn - xl).abs().max()) <= (8e-6 if precision == "fp32" else 4e-3)  # measured 2.6e-6 / 1.5e-3


# ------------------------------------------------------------------------------------------------ BASELINE config C5
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_c5_forward_vs_reference(precision):
    """nf 256, ch_mult [1,2,2,2], 16x16 latents, attention at 16x16 (T = 256 tokens, C = 256): 139 M parameters."""
    g = load_golden("forward_c5.npz")
    cfg = make_config(16, 16, nf=256, ch_mult=(1, 2, 2, 2), W=16, precision=precision)
    ocfg = O.NetConfig(image_size=16, nf=256, ch_mult=(1, 2, 2, 2), attn_resolutions=(16,))
    sd = O.synth_state_dict(ocfg, seed=int(g["weights_seed"]))
    model = mutils.create_model(cfg).to(DEV)
    model.load_state_dict(sd)
    model.eval()
    del sd
    x, sigma, labels = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "sigma", "labels"))
    with torch.no_grad():
        y = model(x, sigma, class_labels=labels)
    err = rel_to_max(y.cpu(), torch.from_numpy(g["y"]))
    eng = list(model._rd_forward_engines.values())[0]
    taps = {k[4:]: rel_to_max(eng.activation(k[4:]).cpu(), torch.from_numpy(g[k])) for k in g.files
            if k.startswith("tap:") and k[4:] in eng.tensors}
    print(f"{precision} C5 forward: rel-to-max err {err:.3e}; taps " + ", ".join(f"{k} {v:.2e}" for k, v in taps.items()))
    assert err <= (3e-5 if precision == "fp32" else 1.4e-2)                 # measured 1.8e-5 / 9.3e-3
    assert max(taps.values()) <= (7e-5 if precision == "fp32" else 1.6e-2)  # measured 4.6e-5 / 1.1e-2


# ------------------------------------------------------------------------------------------------ weight swaps
def test_weight_swap_detection_is_content_based():
    """A norm-preserving in-place update (sign flip of one filter, permutation of two biases) must reach the kernels:
    round 1 compared per-tensor L2 norms and missed it.  Writes go through .data like ExponentialMovingAverage.copy_to."""
    cfg, ocfg, sd, model = build(8, 8)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(4, 1, 8, 9, generator=gen).to(DEV)
    sigma = torch.tensor([2.0, 0.5, 0.1, 0.02]).to(DEV)
    labels = torch.rand(4, 1, generator=gen).to(DEV)
    with torch.no_grad():
        y0 = model(x, sigma, class_labels=labels)
        assert model.rd_sync_weights() is False            # nothing changed: no repack
        p = dict(model.named_parameters())["up_blocks.8.Conv_1.weight"]
        p.data.mul_(-1.0)                                  # same L2 norm, different function
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2["up_blocks.8.Conv_1.weight"] = -sd2["up_blocks.8.Conv_1.weight"]
        y1 = model(x, sigma, class_labels=labels)
        ref1 = O.ncsnpp_forward(x, sigma, labels, {k: v.to(DEV) for k, v in sd2.items()}, ocfg)
        assert rel_to_max(y1, ref1) <= 3e-2 and rel_to_max(y1, y0) > 0.05
        p.data.mul_(-1.0)                                  # restore (EMA restore pattern)
        y2 = model(x, sigma, class_labels=labels)
        assert torch.equal(y2, y0)
        with model.rd_freeze_weights():                    # inside a frozen region the check is skipped by contract
            assert model.rd_sync_weights() is False


# ------------------------------------------------------------------------------------------------ stand-alone layers
@pytest.mark.parametrize("cin,cout,hw,precision", [(64, 64, (8, 9), "bf16"), (64, 128, (4, 4), "bf16"), (128, 128, (2, 2), "fp32"),
                                                     (64, 64, (9, 9), "fp32"), (256, 256, (4, 4), "bf16")])
def test_resblock_forward_standalone(cin, cout, hw, precision):
    """ResnetBlockDDPMpp.forward(x, temb) on its own (layerspp.py:198-214) through rd_resblock."""
    import torch.nn as nn
    from models import layerspp
    torch.manual_seed(5)
    blk = layerspp.ResnetBlockDDPMpp(nn.SiLU(), cin, cout, temb_dim=256, skip_rescale=True, init_scale=1.0).to(DEV).eval()
    blk.rd_precision = precision
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {"b." + k: v.detach() for k, v in blk.state_dict().items()}
    B = 5
    x = torch.randn(B, cin, *hw, device=DEV)
    temb = torch.randn(B, 256, device=DEV)
    y = blk(x, temb)
    with torch.no_grad():
        ref = O.resblock(x, temb, sd, "b", True)
    e = rel_to_max(y, ref)
    print(f"resblock {cin}->{cout} {hw} {precision}: rel-to-max err {e:.3e}")
    assert y.shape == ref.shape and e <= (3e-5 if precision == "fp32" else 1.5e-2)
    y2 = blk(x)   # temb=None: no Dense_0 term at all (layerspp.py:201: `if temb is not None`)
    sd0 = dict(sd)
    sd0["b.Dense_0.weight"], sd0["b.Dense_0.bias"] = torch.zeros_like(sd["b.Dense_0.weight"]), torch.zeros_like(sd["b.Dense_0.bias"])
    with torch.no_grad():
        ref2 = O.resblock(x, temb, sd0, "b", True)
    assert rel_to_max(y2, ref2) <= (3e-5 if precision == "fp32" else 1.5e-2)


def test_attnblock_forward_standalone():
    """AttnBlockpp.forward(x) on its own (layerspp.py:80-96) through rd_attn_block."""
    from models import layerspp
    torch.manual_seed(6)
    blk = layerspp.AttnBlockpp(64, skip_rescale=True, init_scale=1.0).to(DEV).eval()
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {"a." + k: v.detach() for k, v in blk.state_dict().items()}
    for hw in ((8, 9), (9, 9), (4, 4)):
        x = torch.randn(7, 64, *hw, device=DEV)
        y = blk(x)
        with torch.no_grad():
            ref = O.attnblock(x, sd, "a", True)
        e = rel_to_max(y, ref)
        print(f"attnblock {hw}: rel-to-max err {e:.3e}")
        assert e <= 1.5e-2
    with pytest.raises(NotImplementedError):
        layerspp.AttnBlockpp(128).to(DEV)(torch.randn(1, 128, 8, 9, device=DEV))


# ------------------------------------------------------------------------------------------------ other latent shapes
@pytest.mark.parametrize("H,W,attn", [(8, 8, 8), (12, 10, 12), (16, 16, 16), (6, 7, 0), (4, 4, 4)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_forward_other_latent_shapes(H, W, attn, precision):
    """Every shipped consumer of the reference builds a SQUARE sampling shape from image_size alone (run_train.py:124-127,
    Benchmark/gto_halo_benchmarking.py:201-206), so (B,1,8,8) is what an unmodified caller passes; other sizes exercise the
    tile-geometry chooser, the ragged nearest-neighbour fix-up (ncsnpp.py:319-320: 12x10 -> 6x5 -> 3x3 -> 6x6 -> 6x5), the
    fused shortcut at other pixel counts and the attention paths (fused block T <= 128, flash core T = 256)."""
    cfg = make_config(H, attn, W=W, precision=precision)
    ocfg = oracle_cfg(H, attn)
    sd = O.synth_state_dict(ocfg, seed=13)
    model = mutils.create_model(cfg).to(DEV)
    model.load_state_dict(sd)
    model.eval()
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    gen = torch.Generator().manual_seed(H * 100 + W)
    for B in (3, 37):
        x = torch.rand(B, 1, H, W, generator=gen).to(DEV)
        sigma = torch.exp(torch.rand(B, generator=gen) * 6.2 - 4.6).to(DEV)
        labels = torch.rand(B, 1, generator=gen).to(DEV)
        with torch.no_grad():
            y = model(x, sigma, class_labels=labels)
            ref = O.ncsnpp_forward(x, sigma, labels, sdg, ocfg)
        e = rel_to_max(y, ref)
        print(f"{precision} forward {H}x{W} attn@{attn} B={B}: rel-to-max err {e:.3e}")
        assert e <= (5e-5 if precision == "fp32" else 2e-2), (H, W, B, e)
    # a short native sampler run stays inside the cube and matches the oracle on an injected tape
    N, B = 12, 5
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    x0, noise = O.make_tape(B, (1, H, W), (N - 1) * 2, seed=3)
    lab = torch.rand(B, 1, generator=gen).to(DEV)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, H, W), 1e-5, DEV)
    with patched(torch, "rand", lambda *a, **k: x0.clone()):
        xs, _ = fn(model, weight=1.5, class_labels=lab, rd_tape=noise.to(DEV))
    with torch.no_grad():
        xo = O.pc_sampler(lambda xx, sg: O.guided_score(xx, sg, lab, 1.5, sdg, ocfg), O.VESchedule(0.01, 5.0, N, 1.0, 1e-5),
                          O.SamplerConfig(), x0.to(DEV), noise.to(DEV))
    d = (xs - xo).abs()
    print(f"{precision} sampler {H}x{W} N={N}: max {float(d.max()):.3e} mean {float(d.mean()):.3e}")
    assert bool(cube.inside(xs).all())
    assert float(d.mean()) <= (2e-4 if precision == "fp32" else 8e-2)   # (12 coarse steps: the bf16 band is wide, cf. pc_N30)


# ------------------------------------------------------------------------------------------------ the reference's own consumer
def test_reference_consumer_runs_unmodified_on_the_b200_package():
    """Drop-in at the level the reference is actually used: its benchmark driver (Benchmark/gto_halo_benchmarking.py,
    the UNMODIFIED file from oracle/_ref) imports `sampling`, `sde_lib`, `utils`, `losses`, `models.utils`, `models.ema` by
    bare name -- here they resolve to the B200 package -- and its `GTOHaloBenchmarker.generate_samples` runs the EMA swap,
    the sampler closure (square 9x9 shape built from image_size alone) and its own un-normalisation on our samples."""
    import os
    import sys
    import types
    from oracle import fetch_ref
    bench_dir = os.path.normpath(os.path.join(fetch_ref.DST, "..", "Benchmark"))
    if not (fetch_ref.available() and os.path.exists(os.path.join(bench_dir, "gto_halo_benchmarking.py"))):
        pytest.skip("oracle/_ref (copy of the unmodified reference) did not travel")
    for name in ("matplotlib", "matplotlib.pyplot", "omegaconf", "seaborn", "pydylan", "torchvision", "torchvision.transforms",
                 "torchvision.datasets", "PIL"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.OmegaConf = object
                sys.modules[name] = m
    sys.path.append(bench_dir)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import gto_halo_benchmarking as gb
    assert gb.sampling is sampling and gb.sde_lib is sde_lib and gb.mutils is mutils, "the consumer must bind the B200 modules"
    from models.ema import ExponentialMovingAverage
    from rdb200 import codec
    cfg, ocfg, sd, model = build(9, 9)
    ema = ExponentialMovingAverage(model.parameters(), decay=0.999)
    with torch.no_grad():
        for s_ in ema.shadow_params:
            s_.mul_(1.01)                      # the EMA weights differ from the live ones: copy_to must reach the kernels
    before = [p.detach().clone() for p in model.parameters()]
    N, bs = 20, 4
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    real_fn = sampling.get_sampling_fn(cfg, sde, (bs, 1, cfg.model.image_size, cfg.model.image_size), 1e-5, DEV)
    latents, tokens = [], []

    def recording_fn(score_model, **kw):
        tokens.append(score_model._rd_weights_token())   # parameters as the sampler sees them (EMA copied in)
        x, n = real_fn(score_model, **kw)
        latents.append(x.clone())
        return x, n

    b = object.__new__(gb.GTOHaloBenchmarker)
    b.config = types.SimpleNamespace(num_samples=2 * bs, batch_size=bs, guidance_weight=1.5)
    b.device, b.score_model, b.ema, b.sampling_fn = DEV, model, ema, recording_fn
    b.total_spherical_clips, b.total_spherical_elements = 0, 0
    live_token = model._rd_weights_token()
    with contextlib.redirect_stdout(io.StringIO()):
        phys, times = b.generate_samples()
    assert phys.shape == (2 * bs, 67) and np.isfinite(phys).all() and len(times) == 2
    assert all(t != live_token for t in tokens), "sampling must have run on the EMA weights"
    assert all(torch.equal(p, q) for p, q in zip(model.parameters(), before)), "ema.restore must put the live weights back"
    lat = torch.cat(latents, 0)
    assert bool(cube.inside(lat).all())
    ours = codec.gto_halo_decode(lat).cpu().numpy()
    d = np.abs(ours - phys)
    ang = np.zeros(67, bool)
    ang[4:64:3] = True
    ang[5:64:3] = True
    d[:, ang] = np.minimum(d[:, ang], np.abs(d[:, ang] - 2 * np.pi))
    assert float(d.max()) <= 2e-5 * max(1.0, float(np.abs(phys).max())), float(d.max())


# ------------------------------------------------------------------------------------------------ several GPUs, one process
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """The library keeps its launch state (shared-memory opt-ins, SM counts) per device ordinal and every stream call
    runs under the device its stream belongs to: a process that uses cuda:0 and then cuda:1 gets the same numbers."""
    cfg = make_config(8, 8)
    ocfg = oracle_cfg(8, 8)
    sd = O.synth_state_dict(ocfg, seed=7)
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(9, 1, 8, 9, generator=gen)
    sigma = torch.exp(torch.rand(9, generator=gen) * 6.2 - 4.6)
    labels = torch.rand(9, 1, generator=gen)
    outs = []
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        model = mutils.create_model(cfg).to(d)
        model.load_state_dict(sd)
        model.eval()
        with torch.no_grad():
            outs.append(model(x.to(d), sigma.to(d), class_labels=labels.to(d)).cpu())
            r = cube.reflect((x * 3 - 1).to(d))
            assert r.device == torch.device(d) and torch.equal(r.cpu(), O.reflect(x * 3 - 1))
        sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=10)
        fn = sampling.get_sampling_fn(cfg, sde, (9, 1, 8, 9), 1e-5, d)
        xs, _ = fn(model, weight=1.5, class_labels=labels.to(d), rd_seed=4)
        assert xs.device == torch.device(d) and bool(cube.inside(xs).all())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


# ------------------------------------------------------------------------------------------------ output head on mma.sync
_HEAD_SCRIPT = r"""
import sys, types, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
import sde_lib
from models import utils as mutils
from oracle import rd_oracle as O
from helpers import make_config, oracle_cfg
isz, scale_by_sigma = int(sys.argv[4]), bool(int(sys.argv[5]))
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
cfg = make_config(isz, isz, "langevin", scale_by_sigma=scale_by_sigma)
sd = O.synth_state_dict(oracle_cfg(isz, isz, scale_by_sigma=scale_by_sigma), seed=7, out_scale=1.0)
model = mutils.create_model(cfg).to("cuda"); model.load_state_dict(sd); model.eval()
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
B = 37   # not a multiple of the four samples per CTA
x = torch.rand(B, 1, isz, 9, device="cuda"); labels = torch.rand(B, 1, device="cuda")
t = torch.linspace(0.05, 1.0, B, device="cuda")
w = torch.linspace(0.0, 4.0, B, device="cuda")
with torch.no_grad():
    s = mutils.get_cf_score_fn(sde, model, labels, w)(x, t)
torch.save(s.cpu(), sys.argv[6])
"""


@pytest.mark.parametrize("isz,scale_by_sigma", [(8, False), (9, False), (8, True)])
def test_out_head_mma_matches_general_kernel(tmp_path, isz, scale_by_sigma):
    """The one-warp-per-sample output head (mma.sync, split-bf16 operands, csrc/small_ops.cu out_head_mma_kernel) against the
    general fp32 SIMT kernel on the same network activations (RD_OUTHEAD_MMA=0, read once per process -> two subprocesses):
    out_norm + SiLU + out_conv + guidance combine (ncsnpp.py:343-351, models/utils.py:124-138).  Both heads read the same
    bf16 activations, so they may differ only by the 2^-17 operand split and the summation order (a few 1e-6 per network
    pass), which the guidance combine (1 + w) a - w b amplifies by up to 1 + 2w = 9 here: bound 5e-5 of the largest score
    (measured 2.2e-5)."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    outs = []
    for flag in ("1", "0"):
        out = str(tmp_path / f"score_{flag}.pt")
        env = dict(os.environ, RD_OUTHEAD_MMA=flag)
        r = subprocess.run([sys.executable, "-c", _HEAD_SCRIPT, root, os.path.join(root, "optimized-diffusion-model_b200"), here,
                            str(isz), str(int(scale_by_sigma)), out], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(out))
    a, b = outs
    assert torch.isfinite(a).all() and a.abs().max() > 0
    err = float((a - b).abs().max() / b.abs().max())
    assert err <= 5e-5, f"mma output head deviates from the general kernel: {err:.3e} of max"


@pytest.mark.parametrize("isz", [8, 9])
def test_polyphase_stride2_matches_one_plane_layout(tmp_path, isz):
    """Downsample convolutions (layerspp.py:157-159) on the polyphase operand layout (four parity planes, accumulator rows =
    output positions) against the one-plane layout that computes every input position (RD_CONV_POLY=0, read once per
    process -> two subprocesses).  Both issue the same products in the same (chunk, tap, k-step) order into fp32 TMEM
    accumulators, so the guided score of the whole network must be bit-identical."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    outs = []
    for flag in ("1", "0"):
        out = str(tmp_path / f"score_poly_{flag}.pt")
        env = dict(os.environ, RD_CONV_POLY=flag)
        r = subprocess.run([sys.executable, "-c", _HEAD_SCRIPT, root, os.path.join(root, "optimized-diffusion-model_b200"), here,
                            str(isz), "0", out], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(out))
    a, b = outs
    assert torch.isfinite(a).all() and a.abs().max() > 0
    diff = float((a - b).abs().max())
    print(f"polyphase vs one-plane stride-2 layout at {isz}x9: max |diff| = {diff:.3e}")
    assert diff == 0.0
