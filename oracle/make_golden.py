"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/Reflected-Diffusion, imported, never copied) on seeded inputs, and check the
oracle restatement (oracle/rd_oracle.py) against it while doing so.

Run in the build container only (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY -- see the header of oracle/rd_oracle.py.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/Reflected-Diffusion"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import rd_oracle as O  # noqa: E402


def ref_modules():
    sys.path.insert(0, REF)
    import cube  # noqa
    import sampling  # noqa
    import sde_lib  # noqa
    from models import ncsnpp  # noqa
    from models import utils as mutils  # noqa
    sys.path.pop(0)
    return cube, sampling, sde_lib, ncsnpp, mutils


def ref_config(cfg: O.NetConfig, scfg: O.SamplerConfig, W: int):
    m = types.SimpleNamespace(
        name="ncsnpp", channels=cfg.channels, image_size=cfg.image_size, image_width=W,
        num_classes=cfg.num_classes, cond_drop_prob=0.5, conditional=cfg.conditional, init_scale=0.0,
        ema_rate=0.999, nf=cfg.nf, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
        attn_resolutions=list(cfg.attn_resolutions), resamp_with_conv=True, embedding_type="fourier",
        fourier_scale=cfg.fourier_scale, skip_rescale=cfg.skip_rescale, nonlinearity="swish", fir=False,
        fir_kernel=[1, 3, 3, 1], dropout=0.2, scale_by_sigma=cfg.scale_by_sigma)
    s = types.SimpleNamespace(method="pc", predictor=scfg.predictor, corrector=scfg.corrector, denoiser="none",
                              snr=scfg.snr, n_steps_each=scfg.n_steps_each)
    return types.SimpleNamespace(model=m, sampling=s)


def build_ref_model(ncsnpp, cfg, scfg, W, sd):
    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints [DEBUG] lines
        model = ncsnpp.NCSNpp(ref_config(cfg, scfg, W))
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def maxabs(a, b):
    return float((a - b).abs().max())


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cube, sampling, sde_lib, ncsnpp, mutils = ref_modules()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    report = []

    # ---------------------------------------------------------------- reflect
    g = torch.Generator().manual_seed(100)
    edge = torch.tensor([0.0, -0.0, 1.0, 2.0, -1.0, -2.0, 3.0, 1e8, -1e8, -1e-9, 1e-9, 0.5, 1.5, -0.5, -1.5,
                         2.0 - 1e-7, 1.0 + 1e-7, -1e-45, 1e-45, 4000.25, -4000.25, float("inf"), float("-inf"),
                         float("nan"), 1.9999999, 0.99999994, 1.0000001, 16777216.0, 16777217.0, -3.9999998])
    x = torch.cat([edge, (torch.rand(4066, generator=g) - 0.5) * 12, torch.randn(4096, generator=g) * 4000])
    y = cube.reflect(x.clone())
    yo = O.reflect(x)
    assert torch.equal(torch.nan_to_num(y, nan=7.0), torch.nan_to_num(yo, nan=7.0)), "oracle.reflect != reference"
    assert torch.equal(torch.signbit(y), torch.signbit(yo))
    np.savez(os.path.join(out, "reflect.npz"), x=x.numpy(), y=y.numpy())
    report.append("reflect: oracle bitwise == reference on %d values" % x.numel())

    # ---------------------------------------------------------------- score_hk
    B, shape = 64, (1, 8, 9)
    sig_list = [0.01, 0.03, 0.1, 0.1414, 0.1415, 0.15, 0.2, 0.3, 0.5, 1.0, 2.0, 5.0]
    xs, x0s, sgs, refs, refs64 = [], [], [], [], []
    for si, sg in enumerate(sig_list):
        g = torch.Generator().manual_seed(200 + si)
        mean = torch.rand((B,) + shape, generator=g)
        xx = cube.reflect(mean + sg * torch.randn((B,) + shape, generator=g))
        sv = torch.full((B,), sg)
        r32 = cube.score_hk(xx, mean, sv)
        r64 = cube.score_hk(xx.double(), mean.double(), sv.double())
        ro = O.score_hk(xx, mean, sv)
        scale = float(r64.abs().max())
        err_o = maxabs(ro.double(), r64) / scale
        err_r = maxabs(r32.double(), r64) / scale
        assert maxabs(ro, r32) <= 2e-6 * scale + 2.0 * maxabs(r32.double(), r64), (sg, maxabs(ro, r32), scale)
        report.append("score_hk sigma=%g: ref32-vs-ref64 %.2e, oracle32-vs-ref64 %.2e (rel to max %.3g)"
                      % (sg, err_r, err_o, scale))
        xs.append(xx); x0s.append(mean); sgs.append(sv); refs.append(r32); refs64.append(r64)
    # mixed per-sample sigma (log-uniform) + python-float sigma path
    g = torch.Generator().manual_seed(299)
    mean = torch.rand((B,) + shape, generator=g)
    sv = torch.exp(torch.rand(B, generator=g) * (np.log(5.0) - np.log(0.01)) + np.log(0.01))
    xx = cube.reflect(mean + sv.view(-1, 1, 1, 1) * torch.randn((B,) + shape, generator=g))
    xs.append(xx); x0s.append(mean); sgs.append(sv)
    refs.append(cube.score_hk(xx, mean, sv)); refs64.append(cube.score_hk(xx.double(), mean.double(), sv.double()))
    r_float = cube.score_hk(xx, mean, 0.25)
    assert maxabs(O.score_hk(xx, mean, 0.25), r_float) <= 1e-5 * float(r_float.abs().max())
    np.savez(os.path.join(out, "score_hk.npz"), x=torch.stack(xs).numpy(), x_orig=torch.stack(x0s).numpy(),
             sigma=torch.stack(sgs).numpy(), ref32=torch.stack(refs).numpy(),
             ref64=torch.stack(refs64).numpy(), ref32_sigma_float_025=r_float.numpy())

    # ---------------------------------------------------------------- SDE schedule
    sched = O.VESchedule(0.01, 5.0, 1000, 1.0, 1e-5)
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
    ts = torch.linspace(sde.T, 1e-5, sde.N)
    _, gref = sde.sde(torch.zeros(1000, 1, 1, 1), ts)
    _, sref = sde.marginal_prob(torch.zeros(1000, 1, 1, 1), ts)
    assert torch.equal(gref, sched.diffusion(sched.timesteps())) and torch.equal(sref, sched.sigma(sched.timesteps()))
    np.savez(os.path.join(out, "schedule.npz"), t=ts.numpy(), sigma=sref.numpy(), g=gref.numpy())
    report.append("schedule: oracle bitwise == reference (N=1000)")

    # ---------------------------------------------------------------- single predictor / corrector steps
    g = torch.Generator().manual_seed(300)
    B = 16
    x = torch.rand((B,) + shape, generator=g)
    score = torch.randn((B,) + shape, generator=g) * 0.3
    z = torch.randn((B,) + shape, generator=g)
    steps = {}
    for idx in (0, 300, 700, 998):
        vec_t = torch.ones(B) * ts[idx]
        fixed = lambda xx, tt, s=score: s  # noqa: E731
        pred = sampling.ReflectedEulerMaruyamaPredictor(sde, fixed)
        with _patched(torch, "randn_like", lambda t, zz=z: zz.clone()):
            xp, xpm = pred.update_fn(x.clone(), vec_t)
        corr = sampling.ReflectedLangevinCorrector(sde, fixed, 0.01, 1)
        with _patched(torch, "randn_like", lambda t, zz=z: zz.clone()):
            xc, xcm = corr.update_fn(x.clone(), vec_t)
        op, opm = O.predictor_step(x, score, sched.diffusion(vec_t), 1000, z)
        oc, ocm, _ = O.corrector_step(x, score, z, 0.01)
        assert torch.equal(op, xp) and torch.equal(opm, xpm), "oracle predictor != reference"
        assert torch.equal(oc, xc) and torch.equal(ocm, xcm), "oracle corrector != reference"
        steps[f"pred_x_{idx}"] = xp.numpy(); steps[f"pred_mean_{idx}"] = xpm.numpy()
        steps[f"corr_x_{idx}"] = xc.numpy(); steps[f"corr_mean_{idx}"] = xcm.numpy()
    np.savez(os.path.join(out, "pc_steps.npz"), x=x.numpy(), score=score.numpy(), z=z.numpy(), **steps)
    report.append("predictor/corrector single steps: oracle bitwise == reference at i in (0,300,700,998)")

    # ---------------------------------------------------------------- network forward (8x9 attn@8, and 9x9 attn@9)
    for tag, cfg, W in (("8x9", O.NetConfig(image_size=8, attn_resolutions=(8,)), 9),
                        ("9x9", O.NetConfig(image_size=9, attn_resolutions=(9,)), 9)):
        scfg = O.SamplerConfig()
        sd = O.synth_state_dict(cfg, seed=7)
        model = build_ref_model(ncsnpp, cfg, scfg, W, sd)
        g = torch.Generator().manual_seed(400)
        B = 6
        x = torch.rand((B, 1, cfg.image_size, W), generator=g)
        sigma = torch.tensor([5.0, 1.3, 0.4, 0.1, 0.03, 0.01])
        labels = torch.rand((B, 1), generator=g)
        with torch.no_grad():
            yr = model(x, sigma, class_labels=labels)
            taps = {}
            yo = O.ncsnpp_forward(x, sigma, labels, sd, cfg, taps=taps)
            # guided score through the reference wrappers
            ts6 = torch.tensor([1.0, 0.8, 0.6, 0.4, 0.2, 1e-5])
            w = torch.tensor([0.0, 0.5, 1.0, 2.0, 4.0, 1.5])
            sr = mutils.get_cf_score_fn(sde, model, labels, w)(x, ts6)
            so = O.guided_score(x, sched.sigma(ts6), labels, w, sd, cfg)
        scale = float(yr.abs().max())
        assert maxabs(yr, yo) <= 2e-5 * scale, ("forward", tag, maxabs(yr, yo), scale)
        assert maxabs(sr, so) <= 2e-5 * float(sr.abs().max())
        keep = {k: taps[k].numpy() for k in ("temb", "input_conv", "down_blocks.0", "down_attn.0", "downsample.0",
                                              "down_blocks.2", "mid_block2", "up_blocks.0", "upsample.0",
                                              "up_blocks.5", "upsample.1", "up_blocks.6", "up_attn.8")}
        np.savez_compressed(os.path.join(out, f"forward_{tag}.npz"), x=x.numpy(), sigma=sigma.numpy(),
                            labels=labels.numpy(), y=yr.numpy(), t_cfg=ts6.numpy(), w_cfg=w.numpy(),
                            score_cfg=sr.numpy(), **{"tap:" + k: v for k, v in keep.items()})
        report.append("forward %s: |ref|max %.3g, oracle-vs-reference max abs %.2e; guided %.2e"
                      % (tag, scale, maxabs(yr, yo), maxabs(sr, so)))

    # ---------------------------------------------------------------- full pc_sampler with a noise tape
    cfg = O.NetConfig(image_size=8, attn_resolutions=(8,))
    sd = O.synth_state_dict(cfg, seed=7)
    for tag, corrector, N, B in (("pc_N30", "langevin", 30, 4), ("pred_only_N30", "none", 30, 4),
                                 ("pc_N200", "langevin", 200, 2)):
        scfg = O.SamplerConfig(corrector=corrector)
        model = build_ref_model(ncsnpp, cfg, scfg, 9, sd)
        sdeN = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
        schedN = O.VESchedule(0.01, 5.0, N, 1.0, 1e-5)
        n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
        x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=500)
        g = torch.Generator().manual_seed(501)
        labels = torch.rand((B, 1), generator=g)
        w = 1.5
        fn = sampling.get_sampling_fn(ref_config(cfg, scfg, 9), sdeN, (B, 1, 8, 9), 1e-5, "cpu")
        tape = [noise[i] for i in range(n_draws)]
        with _patched(torch, "rand", lambda *a, **k: x0.clone()), \
                _patched(torch, "randn_like", lambda t: tape.pop(0).clone()):
            xr, nfe = fn(model, weight=w, class_labels=labels)
        assert not tape, "tape not fully consumed"
        with torch.no_grad():
            xo = O.pc_sampler(lambda xx, sg: O.guided_score(xx, sg, labels, w, sd, cfg), schedN, scfg, x0, noise)
        assert bool(cube.inside(xr).all())
        d = maxabs(xr, xo)
        assert d <= 5e-4, (tag, d)
        np.savez(os.path.join(out, f"sampler_{tag}.npz"), x_final=xr.numpy(), labels=labels.numpy(),
                 w=np.float32(w), N=N, B=B, tape_seed=500, nfe=nfe)
        report.append("sampler %s: oracle-vs-reference max abs %.2e (nfe reported %d)" % (tag, d, nfe))

    with open(os.path.join(out, "REPORT.txt"), "w") as f:
        f.write("Generated by oracle/make_golden.py with torch %s on %d threads\n" % (torch.__version__, torch.get_num_threads()))
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


@contextlib.contextmanager
def _patched(mod, name, fn):
    old = getattr(mod, name)
    setattr(mod, name, fn)
    try:
        yield
    finally:
        setattr(mod, name, old)


if __name__ == "__main__":
    main()
