"""Per-op timing of one guided-score evaluation at B=8192 (CUDA events around each op, eager)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import torch, bench, sde_lib
from models import utils as mutils
bench.use_config(os.environ.get("RD_PROF_CFG", "c3"))
H, W = bench.CONFIG["H"], bench.CONFIG["W"]
if os.environ.get("RD_PROF_HW"):   # e.g. RD_PROF_HW=9x9: the shape every shipped artefact of the reference uses
    H, W = (int(v) for v in os.environ["RD_PROF_HW"].split("x"))
    bench.CONFIG.update(H=H, W=W)
B = int(os.environ.get("RD_PROF_B", "2048" if bench.CONFIG["name"] == "c5" else "8192"))
dev = torch.device("cuda", 0)
model = mutils.create_model(bench.model_config()).to(dev).eval()
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
eng = model.rd_sampler_engine(B, H, W, dev, sde, 1e-5, 0.01, 1, cfg=True)
x0 = torch.rand(B, 1, H, W, device=dev); labels = torch.rand(B, 1, device=dev)
import time
for iters in (3, 30):
    torch.cuda.synchronize(); t0 = time.time()
    xs = eng.sample(x0, labels, 1.5, seed=1, use_graph=True, n_iter=iters)
    torch.cuda.synchronize(); dt = time.time() - t0
print("B=%d: %.2f ms/iter -> %.1f samples/s" % (B, 1e3 * dt / iters, B / (dt / iters * 999)))
reps = 5
per = [0.0] * eng.n_ops
for _ in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(eng.n_ops + 1)]
    ev[0].record()
    for i in range(eng.n_ops):
        eng.run_ops(i, 1); ev[i + 1].record()
    torch.cuda.synchronize()
    for i in range(eng.n_ops): per[i] += ev[i].elapsed_time(ev[i + 1]) / reps
tot = sum(per)
print("forward (2B=%d): %.3f ms over %d ops" % (2 * B, tot, eng.n_ops))
kinds = collections.Counter()
for t, n in zip(per, eng.op_names): kinds[eng.op_kinds[n]] += t
print("by kind:", {k: round(v, 3) for k, v in kinds.items()})
for t, n in zip(per, eng.op_names):
    i = eng.op_info.get(n)
    if i:
        mmas = i["n_groups"] * i["n_tiles"] * (i["cin"] // 16) * i["taps"]          # per launch, all CTAs
        cyc = t * 1e-3 * 1.965e9 * min(i["grid"], 148) / max(mmas, 1)                 # SM-cycles per MMA issued
        print("   %-26s %.3f ms  %6.1f TF/s  cin %4d n %3d %dx%d taps %d | S %2d nt %d groups %5d grid %3d a_st %d w_res %d w_st %d acc %d rc %2d smem %3dK | %5.1f cyc/MMA%s%s"
              % (n, t, i["flops"] * eng.B2 / (t * 1e-3) / 1e12, i["cin"], i["n"], i["hw"][0], i["hw"][1], i["taps"], i["S"], i["n_tiles"],
                 i["n_groups"], i["grid"], i["a_stages"], i["w_resident"], i["w_stages"], i["acc_bufs"], i["xmode"], i["smem"] // 1024, cyc,
                 " res" if i["res"] else "", " gn" if i["gn"] else ""))
    else:
        print("   %-26s %.3f ms" % (n, t))
