"""Diagnostic run on the GPU box: element-wise kernels and the network forward against the
committed golden fixtures / the oracle.  Prints error magnitudes; asserts nothing."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import numpy as np, torch
from oracle import rd_oracle as O
import cube, sde_lib, sampling
from models import utils as mutils
from rdb200 import ops

dev = "cuda"
G = os.path.join(ROOT, "tests", "golden")
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))

g = np.load(os.path.join(G, "reflect.npz"))
x = torch.from_numpy(g["x"]).to(dev); y = torch.from_numpy(g["y"])
out = cube.reflect(x).cpu()
same = torch.equal(torch.nan_to_num(out, nan=7.0), torch.nan_to_num(y, nan=7.0)) and torch.equal(torch.signbit(out), torch.signbit(y))
print("reflect bitwise:", same, flush=True)
if not same:
    bad = (torch.nan_to_num(out, nan=7.0) != torch.nan_to_num(y, nan=7.0)).nonzero().flatten()[:10]
    print("  bad idx", bad.tolist(), g["x"][bad.numpy()], out[bad], y[bad])

g = np.load(os.path.join(G, "score_hk.npz"))
for i in range(g["x"].shape[0]):
    xx = torch.from_numpy(g["x"][i]).to(dev); x0 = torch.from_numpy(g["x_orig"][i]).to(dev); sg = torch.from_numpy(g["sigma"][i]).to(dev)
    r = cube.score_hk(xx, x0, sg).cpu()
    r32 = torch.from_numpy(g["ref32"][i]); r64 = torch.from_numpy(g["ref64"][i])
    sc = float(r64.abs().max())
    print("score_hk case %d sigma0=%.4g: |max|=%.3g  ours-vs-ref32 %.2e  ours-vs-ref64 %.2e  ref32-vs-ref64 %.2e (rel to max)" % (
        i, float(sg[0]), sc, float((r - r32).abs().max()) / sc if sc > 0 else 0, float((r.double() - r64).abs().max()) / sc if sc > 0 else 0,
        float((r32.double() - r64).abs().max()) / sc if sc > 0 else 0), flush=True)

g = np.load(os.path.join(G, "pc_steps.npz"))
x = torch.from_numpy(g["x"]).to(dev); score = torch.from_numpy(g["score"]).to(dev); z = torch.from_numpy(g["z"]).to(dev)
sch = np.load(os.path.join(G, "schedule.npz"))
for idx in (0, 300, 700, 998):
    xp, xpm = ops.predictor_step(x, score, z, float(sch["g"][idx]), 1000)
    xc, xcm, st = ops.corrector_step(x, score, z, 0.01)
    print("step i=%d: pred x %.2e mean %.2e | corr x %.2e mean %.2e  stats %s" % (idx,
          float((xp.cpu() - torch.from_numpy(g[f"pred_x_{idx}"])).abs().max()), float((xpm.cpu() - torch.from_numpy(g[f"pred_mean_{idx}"])).abs().max()),
          float((xc.cpu() - torch.from_numpy(g[f"corr_x_{idx}"])).abs().max()), float((xcm.cpu() - torch.from_numpy(g[f"corr_mean_{idx}"])).abs().max()), st.tolist()), flush=True)

zz = ops.philox_normal((1 << 20,), 1234, 0, dev)
print("philox: mean %.4f std %.4f kurt %.4f max %.3f" % (float(zz.mean()), float(zz.std()), float((zz ** 4).mean()), float(zz.abs().max())), flush=True)

def make_cfg(image_size, attn):
    return types.SimpleNamespace(model=types.SimpleNamespace(name='ncsnpp', channels=1, image_size=image_size, image_width=9, num_classes=1,
        cond_drop_prob=0.5, conditional=True, init_scale=0.0, ema_rate=0.999, nf=64, ch_mult=[1, 2, 2], num_res_blocks=2, attn_resolutions=[attn],
        resamp_with_conv=True, embedding_type='fourier', fourier_scale=16, skip_rescale=True, nonlinearity='swish', fir=False, fir_kernel=[1, 3, 3, 1],
        dropout=0.2, scale_by_sigma=False),
        sampling=types.SimpleNamespace(method='pc', predictor='euler_maruyama', corrector='langevin', denoiser='none', snr=0.01, n_steps_each=1))

for tag, isz in (("8x9", 8), ("9x9", 9)):
    cfg = make_cfg(isz, isz)
    ocfg = O.NetConfig(image_size=isz, attn_resolutions=(isz,))
    sd = O.synth_state_dict(ocfg, seed=7)
    model = mutils.create_model(cfg).to(dev)
    model.load_state_dict(sd); model.eval()
    g = np.load(os.path.join(G, f"forward_{tag}.npz"))
    x = torch.from_numpy(g["x"]).to(dev); sigma = torch.from_numpy(g["sigma"]).to(dev); labels = torch.from_numpy(g["labels"]).to(dev)
    t0 = time.time()
    with torch.no_grad():
        y = model(x, sigma, class_labels=labels)
    torch.cuda.synchronize()
    print("forward %s ran in %.2fs; out vs golden: max abs %.3e (|ref|max %.3g) rel %.3e" % (tag, time.time() - t0,
          float((y.cpu() - torch.from_numpy(g["y"])).abs().max()), float(np.abs(g["y"]).max()), rel(y.cpu(), torch.from_numpy(g["y"]))), flush=True)
    eng = list(model._rd_forward_engines.values())[0]
    for k in g.files:
        if k.startswith("tap:") and k[4:] in eng.tensors:
            a = eng.activation(k[4:]).cpu(); b = torch.from_numpy(g[k])
            print("   tap %-16s rel-to-max err %.3e" % (k[4:], rel(a, b)), flush=True)
    # pytorch's own bf16 autocast noise floor on the same weights/inputs (oracle on GPU)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    with torch.no_grad():
        y32 = O.ncsnpp_forward(x, sigma, labels, sdg, ocfg)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = O.ncsnpp_forward(x, sigma, labels, sdg, ocfg)
    print("   oracle-on-GPU fp32 vs golden rel %.3e ; torch bf16-autocast vs fp32 rel %.3e" % (rel(y32.cpu(), torch.from_numpy(g["y"])), rel(y16.float(), y32)), flush=True)

# sampler tape parity
cfg = make_cfg(8, 8)
ocfg = O.NetConfig(image_size=8, attn_resolutions=(8,))
sd = O.synth_state_dict(ocfg, seed=7)
model = mutils.create_model(cfg).to(dev); model.load_state_dict(sd); model.eval()
for tag, corrector in (("pc_N30", "langevin"), ("pred_only_N30", "none"), ("pc_N200", "langevin")):
    g = np.load(os.path.join(G, f"sampler_{tag}.npz"))
    N, B = int(g["N"]), int(g["B"])
    cfg.sampling.corrector = corrector
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    n_draws = (N - 1) * (2 if corrector == "langevin" else 1)
    x0, noise = O.make_tape(B, (1, 8, 9), n_draws, seed=int(g["tape_seed"]))
    labels = torch.from_numpy(g["labels"]).to(dev)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 1, 8, 9), 1e-5, dev)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: x0.clone()
    try:
        for graph in (False, True):
            xs, nfe = fn(model, weight=float(g["w"]), class_labels=labels, rd_tape=noise.to(dev), rd_graph=graph)
            d = (xs.cpu() - torch.from_numpy(g["x_final"])).abs()
            print("sampler %s graph=%s: max abs diff vs reference %.3e mean %.3e inside=%s" % (tag, graph, float(d.max()), float(d.mean()), bool(cube.inside(xs).all())), flush=True)
        xs2, _ = fn(model, weight=float(g["w"]), class_labels=labels, rd_tape=noise.to(dev), rd_native=False)
        d = (xs2.cpu() - torch.from_numpy(g["x_final"])).abs()
        print("sampler %s generic loop: max abs diff %.3e mean %.3e" % (tag, float(d.max()), float(d.mean())), flush=True)
    finally:
        torch.rand = real_rand
print("DIAG DONE")
