"""Small, deterministic target for ncu: two guided-score evaluations + one PC iteration at B=8192."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import torch
import bench
import sde_lib
from models import utils as mutils
B = int(os.environ.get("RD_PROF_B", "8192"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mutils.create_model(bench.model_config()).to(dev).eval()
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
eng = model.rd_sampler_engine(B, 8, 9, dev, sde, 1e-5, 0.01, 1, cfg=True)
x0 = torch.rand(B, 1, 8, 9, device=dev); labels = torch.rand(B, 1, device=dev)
xs = eng.sample(x0, labels, 1.5, seed=1, use_graph=False, n_iter=2)
torch.cuda.synchronize()
print("prof target done", float(xs.mean()))
