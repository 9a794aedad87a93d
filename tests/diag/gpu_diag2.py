"""Second diagnostic: sampler parity (free-running, teacher-forced) and a first timing."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import numpy as np, torch
from oracle import rd_oracle as O
import cube, sde_lib, sampling
from models import utils as mutils
dev = "cuda"
G = os.path.join(ROOT, "tests", "golden")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

def make_cfg(image_size, attn, corrector="langevin"):
    return types.SimpleNamespace(model=types.SimpleNamespace(name='ncsnpp', channels=1, image_size=image_size, image_width=9, num_classes=1,
        cond_drop_prob=0.5, conditional=True, init_scale=0.0, ema_rate=0.999, nf=64, ch_mult=[1, 2, 2], num_res_blocks=2, attn_resolutions=[attn],
        resamp_with_conv=True, embedding_type='fourier', fourier_scale=16, skip_rescale=True, nonlinearity='swish', fir=False, fir_kernel=[1, 3, 3, 1],
        dropout=0.2, scale_by_sigma=False),
        sampling=types.SimpleNamespace(method='pc', predictor='euler_maruyama', corrector=corrector, denoiser='none', snr=0.01, n_steps_each=1))

cfg = make_cfg(8, 8)
ocfg = O.NetConfig(image_size=8, attn_resolutions=(8,))
sd = O.synth_state_dict(ocfg, seed=7)
sdg = {k: v.to(dev) for k, v in sd.items()}
model = mutils.create_model(cfg).to(dev); model.load_state_dict(sd); model.eval()
for tag, corrector in (("pc_N30", "langevin"), ("pred_only_N30", "none"), ("pc_N200", "langevin")):
    g = np.load(os.path.join(G, f"sampler_{tag}.npz"))
    N, B = int(g["N"]), int(g["B"])
    cfg.sampling.corrector = corrector
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    sched = O.VESchedule(0.01, 5.0, N, 1.0, 1e-5)
    scfg = O.SamplerConfig(corrector=corrector)
    n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
    x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=int(g["tape_seed"]))
    labels = torch.from_numpy(g["labels"]).to(dev)
    w = float(g["w"])
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, dev)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: x0.clone()
    try:
        for graph in (False, True):
            xs, nfe = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(dev), rd_graph=graph)
            d = (xs.cpu() - torch.from_numpy(g["x_final"])).abs()
            print("sampler %s graph=%s: max abs diff vs reference %.3e mean %.3e inside=%s" % (tag, graph, float(d.max()), float(d.mean()), bool(cube.inside(xs).all())), flush=True)
        xs2, _ = fn(model, weight=w, class_labels=labels, rd_tape=noise.to(dev), rd_native=False)
        d = (xs2.cpu() - torch.from_numpy(g["x_final"])).abs()
        print("sampler %s generic loop: max abs diff %.3e mean %.3e ; native-vs-generic %.3e" % (tag, float(d.max()), float(d.mean()), float((xs2 - xs).abs().max())), flush=True)
    finally:
        torch.rand = real_rand
    # oracle on GPU: fp32 and torch-bf16-autocast free-running, plus teacher-forced per-step parity
    with torch.no_grad():
        trace = []
        xo = O.pc_sampler(lambda xx, sg: O.guided_score(xx, sg, labels, w, sdg, ocfg), sched, scfg, x0.to(dev), noise.to(dev), trace=trace)
        def bf16_score(xx, sg):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return O.guided_score(xx, sg, labels, w, sdg, ocfg).float()
        xb = O.pc_sampler(bf16_score, sched, scfg, x0.to(dev), noise.to(dev))
    ref = torch.from_numpy(g["x_final"]).to(dev)
    print("   oracle-GPU-fp32 vs reference-CPU: %.3e ; torch-bf16-autocast vs fp32: max %.3e mean %.3e" % (
        float((xo - ref).abs().max()), float((xb - xo).abs().max()), float((xb - xo).abs().mean())), flush=True)
    eng = list(model._rd_sampler_engines.values())[-1]
    worst = 0.0; worst_i = -1; errs = []
    for (i, x_c, x_p) in trace:
        x_in = x0.to(dev) if i == 0 else trace[i - 1][2]
        xs = eng.sample(x_in, labels, w, tape=noise.to(dev), seed=0, use_graph=False, n_iter=1, start_step=i)
        e = float((xs - x_p).abs().max()); errs.append(e)
        if e > worst: worst, worst_i = e, i
    print("   teacher-forced per-iteration max abs err: worst %.3e at i=%d ; median %.3e ; first 5 %s" % (worst, worst_i, float(np.median(errs)), ["%.1e" % v for v in errs[:5]]), flush=True)

# ---- first timing: B=8192, a few iterations
B = 8192
cfg = make_cfg(8, 8)
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
model = mutils.create_model(cfg).to(dev); model.eval()
labels = torch.rand(B, 1, device=dev)
eng = model.rd_sampler_engine(B, 8, 9, dev, sde, 1e-5, 0.01, 1, cfg=True)
x0 = torch.rand(B, 1, 8, 9, device=dev)
for iters in (3, 20):
    torch.cuda.synchronize(); t0 = time.time()
    xs = eng.sample(x0, labels, 1.5, seed=1, use_graph=True, n_iter=iters)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("B=%d: %d PC iterations in %.3fs -> %.2f ms/iter -> %.1f samples/s for 999 iters ; inside=%s" % (B, iters, dt, 1e3 * dt / iters, B / (dt / iters * 999), bool(cube.inside(xs).all())), flush=True)
# per-op timing of one forward
ev = [torch.cuda.Event(enable_timing=True) for _ in range(eng.n_ops + 1)]
eng.run_plan(); torch.cuda.synchronize()
ev[0].record()
for i in range(eng.n_ops):
    eng.run_ops(i, 1); ev[i + 1].record()
torch.cuda.synchronize()
ts = [(ev[i].elapsed_time(ev[i + 1]), eng.op_names[i]) for i in range(eng.n_ops)]
tot = sum(t for t, _ in ts)
print("one guided-score evaluation (2B=%d): %.3f ms over %d ops" % (2 * B, tot, eng.n_ops))
for t, n in sorted(ts, reverse=True)[:25]:
    print("   %-28s %.3f ms" % (n, t))
print("DIAG2 DONE")
