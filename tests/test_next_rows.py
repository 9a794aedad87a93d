"""The "next" rows of SURVEY.md section 8f -- evaluation DSM loss, EMA, checkpoint ingestion and the
latent->physical codec.  CPU tests pin the oracle (and the torch-only host pieces) against the
reference-generated fixtures; GPU tests run the drop-in modules through the C ABI."""
import numpy as np
import pytest
import torch

import losses
import sde_lib
import utils as rd_utils
from models import utils as mutils
from models.ema import ExponentialMovingAverage
from oracle import rd_oracle as O
from helpers import load_golden, make_config, oracle_cfg


# ------------------------------------------------------------------------------------------ CPU
def test_oracle_dsm_loss_matches_reference_fixture():
    f = load_golden("dsm_loss.npz")
    ocfg = oracle_cfg(8, 8)
    sd = O.synth_state_dict(ocfg, seed=int(f["weights_seed"]))
    sched = O.VESchedule(0.01, 5.0, 1000, 1.0, 1e-5)
    batch, labels, u, z = (torch.from_numpy(f[k]) for k in ("batch", "labels", "u", "z"))
    for rm in (True, False):
        for lw in (True, False):
            with torch.no_grad():
                lo, per, pert, tgt = O.dsm_loss(lambda x, s: O.ncsnpp_forward(x, s, labels, sd, ocfg), sched, batch, u, z,
                                                reduce_mean=rm, likelihood_weighting=lw)
            ref = float(f["loss_rm%d_lw%d" % (rm, lw)])
            assert abs(float(lo) - ref) <= 2e-5 * abs(ref)
    assert float(pert.min()) >= 0.0 and float(pert.max()) <= 1.0


def test_ema_dropin_matches_reference_fixture():
    f = load_golden("ema.npz")
    params = [torch.nn.Parameter(torch.from_numpy(f["p0_a"]).clone()), torch.nn.Parameter(torch.from_numpy(f["p0_b"]).clone())]
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)  # skipped by the average, like the reference
    ema = ExponentialMovingAverage(params + [frozen], decay=0.999)
    assert len(ema.shadow_params) == 2
    for i in range(f["d_a"].shape[0]):
        with torch.no_grad():
            params[0].add_(torch.from_numpy(f["d_a"][i]))
            params[1].add_(torch.from_numpy(f["d_b"][i]))
        ema.update(params + [frozen])
    assert ema.num_updates == int(f["num_updates"])
    assert torch.allclose(ema.shadow_params[0], torch.from_numpy(f["shadow_a"]), rtol=0, atol=1e-6)
    assert torch.allclose(ema.shadow_params[1], torch.from_numpy(f["shadow_b"]), rtol=0, atol=1e-6)
    live = [p.detach().clone() for p in params]
    ema.store(params)
    ema.copy_to(params)
    assert torch.equal(params[0].data, ema.shadow_params[0])
    ema.restore(params)
    assert torch.equal(params[0].data, live[0]) and torch.equal(params[1].data, live[1])
    ema2 = ExponentialMovingAverage(params, decay=0.5)
    ema2.load_state_dict(ema.state_dict())
    assert ema2.decay == 0.999 and ema2.num_updates == ema.num_updates
    with pytest.raises(ValueError):
        ExponentialMovingAverage(params, decay=1.5)


def test_codec_oracle_hand_checked_sample():
    lat = np.zeros((2, 81), np.float32)
    lat[0, 0] = 0.5                                   # label -> halo energy 0.008 + 0.5 * 0.087
    lat[0, 1] = (0.25 - 0.4652) / 0.1811              # shooting time 0.25 * 40 = 10
    lat[0, 4:7] = (np.array([1.0, 0.5, 0.5]) - 0.4652) / 0.1811   # u = (1, 0, 0) -> alpha 0, beta 0, r 1
    lat[0, 7:10] = (np.array([0.5, 0.0, 0.5]) - 0.4652) / 0.1811  # u = (0, -1, 0) -> alpha 3pi/2
    lat[0, 64] = (0.5 - 0.4652) / 0.1811              # fuel mass 408 + 31
    out = O.gto_halo_decode(lat)
    assert out.shape == (2, 67)
    assert abs(out[0, 0] - 0.0515) < 1e-6 and abs(out[0, 1] - 10.0) < 1e-4
    assert np.allclose(out[0, 4:7], [0.0, 0.0, 1.0], atol=1e-5)
    assert np.allclose(out[0, 7:10], [1.5 * np.pi, 0.0, 1.0], atol=1e-5)
    assert abs(out[0, 64] - 439.0) < 1e-3
    assert np.all(out[:, 6:64:3] <= 1.0)


def test_oracle_ode_sampler_matches_reference_fixture():
    f = load_golden("ode_moll0.npz")
    ocfg = oracle_cfg(8, 8)
    sd = O.synth_state_dict(ocfg, seed=int(f["weights_seed"]))
    sched = O.VESchedule(0.01, 5.0, 1000, 1.0, 1e-5)
    x0, lab = torch.from_numpy(f["x0"]), torch.from_numpy(f["labels"])
    with torch.no_grad():
        x, nfe = O.ode_sampler(lambda xx, sg: O.guided_score(xx, sg, lab, float(f["w"]), sd, ocfg), sched, x0,
                               rtol=float(f["rtol"]), atol=float(f["atol"]), moll=float(f["moll"]))
    assert nfe == int(f["nfe"])
    assert float((x - torch.from_numpy(f["x_final"])).abs().max()) <= 1e-4


def test_training_entry_points_refuse():
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    with pytest.raises(NotImplementedError):
        losses.get_sde_loss_fn(sde, train=True)
    with pytest.raises(NotImplementedError):
        losses.get_step_fn(sde, train=True)
    with pytest.raises(ValueError):
        rd_utils.load_denoising_model("/nonexistent/ckpt.pth", None)


# ------------------------------------------------------------------------------------------ GPU
def _model(seed):
    cfg = make_config(8, 8)
    ocfg = oracle_cfg(8, 8)
    sd = O.synth_state_dict(ocfg, seed=seed)
    model = mutils.create_model(cfg).to("cuda")
    model.load_state_dict(sd)
    return cfg, ocfg, sd, model.eval()


@pytest.mark.gpu
def test_eval_loss_vs_reference_fixture():
    f = load_golden("dsm_loss.npz")
    _, ocfg, sd, model = _model(int(f["weights_seed"]))
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    batch, labels, u, z = (torch.from_numpy(f[k]).cuda() for k in ("batch", "labels", "u", "z"))
    # the kernels that do not involve the bf16 network are exact / 1e-6 against the reference's tensors
    from rdb200 import ops
    import cube
    t = u * (sde.T - 1e-5) + 1e-5
    mean, std = sde.marginal_prob(batch, t)
    pert = ops.perturb_reflect(mean, z, std)
    assert torch.equal(pert.cpu(), torch.from_numpy(f["perturbed"]))
    tgt = cube.score_hk(pert, mean, std)
    ref_t = torch.from_numpy(f["target"])
    assert float((tgt.cpu() - ref_t).abs().max()) <= 2e-5 * float(ref_t.abs().max())
    for rm in (True, False):
        for lw in (True, False):
            fn = losses.get_sde_loss_fn(sde, train=False, reduce_mean=rm, likelihood_weighting=lw)
            with torch.no_grad():
                got = float(fn(model, batch, class_labels=labels, rd_t=u, rd_z=z))
            ref = float(f["loss_rm%d_lw%d" % (rm, lw)])
            # bf16 network inside a squared error: same band as the guided-score check (test_gpu_network.py)
            assert abs(got - ref) <= 5e-2 * abs(ref), (rm, lw, got, ref)
    # reduction kernel alone against torch on identical inputs
    s = torch.randn(64, 72, device="cuda")
    g = torch.randn(64, 72, device="cuda")
    w = torch.rand(64, device="cuda")
    for rm in (True, False):
        want = (w[:, None] * (s - g) ** 2)
        want = want.mean(-1) if rm else 0.5 * want.sum(-1)
        assert torch.allclose(ops.dsm_reduce(s, g, w, rm), want, rtol=2e-6, atol=0)
    # unseeded call draws on the device and stays finite
    fn = losses.get_sde_loss_fn(sde, train=False)
    assert torch.isfinite(fn(model, batch, class_labels=labels))


@pytest.mark.gpu
def test_checkpoint_ingestion_and_ema_swap(tmp_path):
    """A checkpoint in the reference's layout (utils.py:77-86) restores into the B200 model; the EMA swap
    around a sampler call (gto_halo_benchmarking.py:230-239) changes the weights the kernels see and puts
    the live ones back."""
    import sampling
    cfg, ocfg, sd, model = _model(21)
    ema_sd = O.synth_state_dict(ocfg, seed=22)
    ema_sd["time_embed.W"] = sd["time_embed.W"].clone()  # frozen (requires_grad=False): not part of the average
    donor = mutils.create_model(cfg)
    donor.load_state_dict(ema_sd)
    ckpt = {"step": 1234, "model": {k: v.clone() for k, v in sd.items()}, "optimizer": {"state": {}, "param_groups": []},
            "ema": {"decay": 0.999, "num_updates": 1234, "shadow_params": [p.detach().clone() for p in donor.parameters() if p.requires_grad]},
            "scaler": None, "config": None}
    path = str(tmp_path / "checkpoint_1.pth")
    torch.save(ckpt, path)

    fresh = mutils.create_model(cfg).to("cuda").eval()
    state = dict(model=fresh, ema=ExponentialMovingAverage(fresh.parameters(), decay=0.5), step=0)
    state = rd_utils.restore_checkpoint(path, state, "cuda")
    assert state["step"] == 1234 and state["ema"].decay == 0.999
    assert rd_utils.restore_checkpoint(str(tmp_path / "missing" / "x.pth"), dict(state), "cuda")["step"] == 1234

    x = torch.rand(4, 1, 8, 9, device="cuda")
    sigma = torch.full((4,), 0.7, device="cuda")
    labels = torch.rand(4, 1, device="cuda")
    with torch.no_grad():
        y_live = fresh(x, sigma, class_labels=labels).clone()
        want_live = O.ncsnpp_forward(x.cpu(), sigma.cpu(), labels.cpu(), sd, ocfg)
        state["ema"].store(fresh.parameters())
        state["ema"].copy_to(fresh.parameters())
        y_ema = fresh(x, sigma, class_labels=labels).clone()
        want_ema = O.ncsnpp_forward(x.cpu(), sigma.cpu(), labels.cpu(), ema_sd, ocfg)
        state["ema"].restore(fresh.parameters())
        y_back = fresh(x, sigma, class_labels=labels).clone()
    tol = lambda want: 3e-2 * float(want.abs().max())  # noqa: E731
    assert float((y_live.cpu() - want_live).abs().max()) <= tol(want_live)
    assert float((y_ema.cpu() - want_ema).abs().max()) <= tol(want_ema)
    assert float((y_ema - y_live).abs().max()) > 10 * tol(want_live)  # the swap really changed the network
    assert torch.equal(y_back, y_live)

    # eval step function: EMA weights in, live weights back
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    step = losses.get_step_fn(sde, train=False, reduce_mean=True, likelihood_weighting=True)
    before = [p.detach().clone() for p in fresh.parameters()]
    loss = step(state, torch.rand(8, 1, 8, 9, device="cuda"), class_labels=torch.rand(8, 1, device="cuda"))
    assert torch.isfinite(loss)
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, fresh.parameters()))
    # save in the reference layout and read it back with the model-only loader
    out = str(tmp_path / "resaved.pth")
    rd_utils.save_checkpoint(out, state)
    again = rd_utils.load_denoising_model(out, mutils.create_model(cfg), device="cpu")
    assert all(torch.equal(a.cpu(), b) for a, b in zip(fresh.state_dict().values(), again.state_dict().values()))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["moll0", "moll200"])
def test_ode_sampler_vs_reference_fixture(tag):
    import sampling
    from rdb200 import ops
    f = load_golden(f"ode_{tag}.npz")
    cfg, ocfg, sd, model = _model(int(f["weights_seed"]))
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    x0, lab = torch.from_numpy(f["x0"]).cuda(), torch.from_numpy(f["labels"]).cuda()
    fn = sampling.get_ode_sampler(sde, tuple(x0.shape), rtol=float(f["rtol"]), atol=float(f["atol"]), eps=1e-5,
                                  moll=float(f["moll"]), device="cuda")
    x, nfe = fn(model, z=x0.clone(), weight=float(f["w"]), class_labels=lab)                      # device-side RK45
    xh, nfe_h = fn(model, z=x0.clone(), weight=float(f["w"]), class_labels=lab, rd_host_solver=True)  # scipy loop, same RHS
    ref = torch.from_numpy(f["x_final"])
    err = (x.cpu() - ref).abs()
    print("ode %s: nfe %d (reference %d, scipy loop over our RHS %d) max %.3e mean %.3e | device-vs-scipy max %.3e"
          % (tag, nfe, int(f["nfe"]), nfe_h, float(err.max()), float(err.mean()), float((x - xh).abs().max())))
    # same scheme, same control law, same right-hand side: the device solver and scipy take the same steps
    assert abs(nfe - nfe_h) <= max(6, 0.01 * nfe_h), (nfe, nfe_h)
    assert float((x - xh).abs().max()) <= 1e-6
    # (against the reference's own run the step SEQUENCE differs: with synthetic weights the mollified flow is sensitive
    #  to 1e-5 perturbations of the right-hand side -- the fp32-class plan lands at 1772 evaluations, bf16 at 1418, the
    #  reference at 1490 -- so the solver is pinned by the exact scipy match above and the samples by the floor below)
    assert abs(nfe - int(f["nfe"])) <= 0.25 * int(f["nfe"]), (nfe, int(f["nfe"]))
    # floor: the oracle's own ODE run with PyTorch bf16 autocast around the network, same weights and x0
    sdg = {k: v.cuda() for k, v in sd.items()}
    sched = O.VESchedule(0.01, 5.0, 1000, 1.0, 1e-5)

    def autocast_score(xx, sg):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.guided_score(xx.cuda(), sg.cuda(), lab, float(f["w"]), sdg, ocfg).float().cpu()

    with torch.no_grad():
        xa, nfa = O.ode_sampler(autocast_score, sched, x0.cpu(), rtol=float(f["rtol"]), atol=float(f["atol"]),
                                moll=float(f["moll"]))
    floor = (xa - ref).abs()
    print("   bf16-autocast floor: nfe %d max %.3e mean %.3e" % (nfa, float(floor.max()), float(floor.mean())))
    if tag == "moll200":  # the configured mode (configs/vis.yaml:17): must sit inside the bf16 band
        assert float(err.max()) <= max(1.5 * float(floor.max()), 2e-2)
        assert float(err.mean()) <= max(1.5 * float(floor.mean()), 5e-3)
    else:  # moll = 0 multiplies the drift by x itself: non-contracting flow, one trajectory, 122 coarse steps
        assert float(err.max()) <= 0.1 and float(err.mean()) <= max(3 * float(floor.mean()), 2e-2)
    assert nfe <= 4 * int(f["nfe"])
    # the drift kernel alone is exact against the reference expression
    xs, sc, g = torch.rand(8, 1, 8, 9, device="cuda"), torch.randn(8, 1, 8, 9, device="cuda"), torch.rand(8, device="cuda") * 17
    for moll in (200.0, 0.0):
        bump = ((-1 / (0.5 ** 2 - (0.5 - xs).pow(2)) + 4) / moll).exp() if moll > 0 else xs
        want = (torch.zeros_like(xs) - g[:, None, None, None] ** 2 * sc * 0.5) * bump
        got = ops.pf_drift(xs, sc, g, moll)
        assert torch.allclose(got, want, rtol=2e-6, atol=1e-30), float((got - want).abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(257, 1, 9, 9), (64, 1, 8, 9)])
def test_codec_vs_oracle(shape):
    from rdb200 import codec
    g = torch.Generator().manual_seed(5)
    lat = torch.rand(shape, generator=g) * 1.6 - 0.3  # includes |u| > 1 (clipped) and negative angles
    lat[0].zero_()
    lat[1, 0, 0, 4:7] = torch.tensor([0.5, 0.5, 0.5])
    lat[1, 0, 0, 4:7] = (lat[1, 0, 0, 4:7] - 0.4652) / 0.1811  # exactly zero control vector -> beta = 0 branch
    want = O.gto_halo_decode(lat.numpy())
    got = codec.gto_halo_decode(lat.cuda()).cpu().numpy()
    assert got.shape == want.shape == (shape[0], 67)
    # angles wrap at 2 pi: compare on the circle
    d = np.abs(got - want)
    ang = np.zeros(67, bool)
    ang[4:64:3] = True
    ang[5:64:3] = True
    d[:, ang] = np.minimum(d[:, ang], np.abs(d[:, ang] - 2 * np.pi))
    assert float(d.max()) <= 2e-5 * max(1.0, float(np.abs(want).max())), float(d.max())
    with pytest.raises(ValueError):
        codec.gto_halo_decode(torch.rand(4, 1, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        codec.gto_halo_decode(torch.rand(4, 81))


@pytest.mark.gpu
def test_codec_pinned_against_reference_outputs():
    """tests/golden/codec.npz holds outputs of the reference's OWN code: the tail of GTOHaloBenchmarker.generate_samples
    (Benchmark/gto_halo_benchmarking.py:255-363) and datasets.GTOHaloImageDataset (datasets.py:82-98), driven by
    oracle/make_golden_r2.py.  Decode: within 2e-5 (device asin / atan2 vs numpy's); encode: bit-exact."""
    from helpers import load_golden
    from rdb200 import codec
    g = load_golden("codec.npz")
    got = codec.gto_halo_decode(torch.from_numpy(g["latents"]).cuda()).cpu().numpy()
    want = g["physical"]
    d = np.abs(got - want)
    ang = np.zeros(67, bool)
    ang[4:64:3] = True
    ang[5:64:3] = True
    d[:, ang] = np.minimum(d[:, ang], np.abs(d[:, ang] - 2 * np.pi))
    assert got.shape == want.shape and float(d.max()) <= 2e-5 * max(1.0, float(np.abs(want).max())), float(d.max())
    lat, lab = codec.gto_halo_encode(torch.from_numpy(g["raw"]).cuda(), 9)
    assert torch.equal(lat.cpu(), torch.from_numpy(g["enc_img"])) and torch.equal(lab.cpu(), torch.from_numpy(g["enc_label"]))
    # encode -> decode round trip of in-range rows returns the physical-unit image of the row (size-independent property)
    raw = torch.rand(1000, 67, device="cuda")
    lat, lab = codec.gto_halo_encode(raw, 9)
    back = codec.gto_halo_decode(lat)
    want = torch.from_numpy(O.gto_halo_decode(O.gto_halo_encode(raw.cpu().numpy(), 9)[0])).cuda()
    assert float((back - want).abs().max()) <= 1e-3
    with pytest.raises(ValueError):
        codec.GtoHaloConstants(n_variables=66).c_struct()   # 59 control values: not whole triplets
    with pytest.raises(ValueError):
        codec.gto_halo_encode(torch.rand(4, 90, device="cuda"), 9)
