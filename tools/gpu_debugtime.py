import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import torch, bench, sde_lib
from models import utils as mutils
B = 8192
dev = torch.device("cuda", 0)
model = mutils.create_model(bench.model_config()).to(dev).eval()
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
eng = model.rd_sampler_engine(B, 8, 9, dev, sde, 1e-5, 0.01, 1, cfg=True)
eng.run_plan(); torch.cuda.synchronize()
reps = 5
per = [0.0] * eng.n_ops
for _ in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(eng.n_ops + 1)]
    ev[0].record()
    for i in range(eng.n_ops):
        eng.run_ops(i, 1); ev[i + 1].record()
    torch.cuda.synchronize()
    for i in range(eng.n_ops): per[i] += ev[i].elapsed_time(ev[i + 1]) / reps
sel = ["down_blocks.0.Conv_0", "down_blocks.0", "down_blocks.3.Conv_0", "mid_block1.Conv_0", "up_blocks.3.Conv_0", "upsample.1", "up_blocks.6.Conv_0", "up_blocks.7.Conv_0", "up_blocks.6.NIN_0"]
d = dict(zip(eng.op_names, per))
print("DEBUG=%s conv total %.3f | " % (os.environ.get("RD_CONV_DEBUG", "0"), sum(t for t, n in zip(per, eng.op_names) if eng.op_kinds[n] == "conv")) + " ".join("%s=%.3f" % (k.replace("blocks.", "b").replace("Conv_0", "c0"), d[k]) for k in sel))
