import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optimized-diffusion-model_b200"))
import torch, bench, sde_lib, numpy as np
from models import utils as mutils
from rdb200._lib import lib
B = 8192
dev = torch.device("cuda", 0)
model = mutils.create_model(bench.model_config()).to(dev).eval()
sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=1000)
eng = model.rd_sampler_engine(B, 8, 9, dev, sde, 1e-5, 0.01, 1, cfg=True)
eng.run_plan(); torch.cuda.synchronize()
G = 28
for name in sys.argv[1:] or ["down_blocks.0.Conv_0"]:
    i = eng.op_names.index(name)
    buf = torch.zeros(3 * G * 8, dtype=torch.int64, device=dev)
    lib().rd_conv_set_trace(buf.data_ptr(), G)
    eng.run_ops(i, 1); torch.cuda.synchronize()
    lib().rd_conv_set_trace(None, 0)
    t = buf.cpu().numpy().reshape(3, G, 8)
    t0 = t[t > 0].min()
    r = np.where(t > 0, t - t0, -1)
    print("==== %s (cycles since first stamp; CTA 0)" % name)
    for li in range(min(G, 12)):
        print("g%02d MMA[top %6d accE %6d aFull %6d issued %6d] EPI[top %6d bar %6d accF %6d done %6d] XF[top %6d rec %6d bar %6d stat %6d aEmpty %6d norm %6d fence %6d arr %6d]" % (
            (li,) + tuple(r[0, li, :4]) + tuple(r[1, li, :4]) + tuple(r[2, li, :8])))
    last = r[r >= 0].max()
    print("total cycles", last)
    # compact steady-state summary (groups 4..11)
    gs = range(4, 12)
    def avg(f): return sum(f(li) for li in gs) / len(gs)
    print("SUMMARY %s debug=%s: period %.0f | MMA issue %.0f exec(accF - max(prev accF, aFull)) %.0f accE-wait %.0f | EPI accF->done %.0f wait-accF %.0f | XF load+rec %.0f stats %.0f aEmpty-wait %.0f norm+store %.0f" % (
        name, os.environ.get("RD_CONV_DEBUG", "0"),
        avg(lambda li: r[0, li + 1, 0] - r[0, li, 0]),
        avg(lambda li: r[0, li, 3] - r[0, li, 2]),
        avg(lambda li: r[1, li, 2] - max(r[1, li - 1, 2], r[0, li, 2])),
        avg(lambda li: r[0, li, 1] - r[0, li, 0]),
        avg(lambda li: r[1, li, 3] - r[1, li, 2]),
        avg(lambda li: r[1, li, 2] - r[1, li, 1]),
        avg(lambda li: r[2, li, 1] - r[2, li, 0]),
        avg(lambda li: r[2, li, 3] - r[2, li, 1]),
        avg(lambda li: r[2, li, 4] - r[2, li, 3]),
        avg(lambda li: r[2, li, 5] - r[2, li, 4])))
