"""Diagnostic: guided-score error of the CUDA path vs fp32 oracle across noise levels, next to the
PyTorch bf16-autocast floor (same weights / inputs).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import rd_oracle as O
from models import utils as mutils
from helpers import make_config, oracle_cfg

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg, ocfg = make_config(8, 8), oracle_cfg(8, 8)
sd = O.synth_state_dict(ocfg, seed=11)
sdg = {k: v.cuda() for k, v in sd.items()}
model = mutils.create_model(cfg).cuda().eval(); model.load_state_dict(sd)
g = torch.Generator().manual_seed(3)
B = 64
lab = torch.rand(B, 1, generator=g).cuda()
for lo, hi in ((0.0, 1.0), (-0.5, 1.5)):
    x = (torch.rand(B, 1, 8, 9, generator=g) * (hi - lo) + lo).cuda()
    for s in (5.0, 2.0, 1.0, 0.5, 0.2, 0.1, 0.05, 0.02, 0.01):
        sg = torch.full((B,), s, device="cuda")
        with torch.no_grad():
            ref = O.guided_score(x, sg, lab, 1.5, sdg, ocfg)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ac = O.guided_score(x, sg, lab, 1.5, sdg, ocfg).float()
            ours = model.rd_guided_score(x, sg, lab, 1.5)
        m = float(ref.abs().max())
        print("x in [%.1f,%.1f] sigma %5.2f |ref|max %.3e  ours max %.3e mean %.3e | autocast max %.3e mean %.3e" % (
            lo, hi, s, m, float((ours - ref).abs().max()) / m, float((ours - ref).abs().mean()) / m,
            float((ac - ref).abs().max()) / m, float((ac - ref).abs().mean()) / m))
