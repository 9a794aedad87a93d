"""BASELINE config C2: cube.reflect / cube.score_hk / fused PC-update microbench on synthetic [2^20,1,8,9] latents
across the VESDE sigma range on one B200, through the C ABI with pre-allocated buffers.

Every kernel is timed with CUDA events on torch's current stream (the stream the C ABI launches on), 10 warm-up
launches then `--reps` timed ones; inputs are 302 MB each (> 126 MB L2), so there is no cross-iteration reuse.
GB/s = ALGORITHMIC bytes (reflect 8 B/element; score_hk 12 B/element + 4 B/sample; fused
predictor / corrector apply 12 B/element with in-kernel Philox noise; corrector norm pass 4 B/element) over the
mean launch time, against MEASURED_PEAKS.json's HBM copy bandwidth.  One JSON line per case.

  python tools/bench_c2.py [--log2B 20] [--reps 10] [--out gpurun_out/c2.jsonl]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimized-diffusion-model_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from rdb200._lib import check, lib, ptr, stream_ptr  # noqa: E402


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def timed(fn, reps):
    for _ in range(10):   # first launches on freshly allocated buffers run slow (seen on `reflect`, the first case)
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return sum(ts) / len(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2B", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the case name (for ncu captures)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    run(args)


def run(args, dev=None, quiet=False):
    """-> list of result dicts (one per case); `args` needs .log2B .reps .only .out"""
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    B, D = 1 << args.log2B, 72
    n = B * D
    peak, src = hbm_peak()
    L = lib()
    st = stream_ptr(dev)
    g = torch.Generator(device=dev).manual_seed(4)
    mean = torch.rand((B, 1, 8, 9), device=dev, generator=g)
    z = torch.randn((B, 1, 8, 9), device=dev, generator=g)
    out = torch.empty_like(mean)
    xin = torch.empty_like(mean)
    lines = []

    def emit(name, ms_avg, ms_min, nbytes, extra=None):
        gbs = nbytes / (ms_avg * 1e-3) / 1e9
        line = {"case": name, "B": B, "D": D, "ms_avg": ms_avg, "ms_min": ms_min, "algorithmic_bytes": nbytes,
                "GBps": gbs, "peak_GBps": peak, "peak_source": src, "frac": gbs / peak}
        if extra:
            line.update(extra)
        lines.append(line)
        if not quiet:
            print(json.dumps(line), flush=True)

    def want(name):
        return args.only is None or args.only in name

    # ---- cube.reflect on un-reflected mean + sigma*z (sigma = 1: ~40 % of the values leave [0,1])
    if want("reflect"):
        torch.add(mean, z, out=xin)
        a, m = timed(lambda: check(L.rd_reflect_f32(ptr(xin), ptr(out), n, st), "reflect"), args.reps)
        emit("reflect", a, m, 8 * n)

    # ---- cube.score_hk: fixed sigma sweep + per-sample log-uniform sigma over the VESDE range
    sig_cases = [("sigma=%g" % s, s) for s in (0.01, 0.03, 0.1, 0.1414, 0.1415, 0.2, 0.5, 1.0, 2.0, 5.0)]
    sig_cases.append(("sigma~logU[0.01,5]", None))
    for name, s in sig_cases:
        if not want("score_hk " + name):
            continue
        if s is None:
            sig = torch.exp(torch.rand(B, device=dev, generator=g) * (math.log(5.0) - math.log(0.01)) + math.log(0.01))
        else:
            sig = torch.full((B,), s, device=dev)
        torch.addcmul(mean, sig.view(-1, 1, 1, 1), z, out=xin)
        check(L.rd_reflect_f32(ptr(xin), ptr(xin), n, st), "reflect")   # x = reflect(mean + sigma z), in place
        ws = torch.empty((int(L.rd_score_hk_workspace_bytes(B)),), dtype=torch.uint8, device=dev)
        a, m = timed(lambda: check(L.rd_score_hk_ws_f32(ptr(xin), ptr(mean), ptr(sig), 0.0, ptr(out), B, D, 20, 10, 1e-2,
                                                        ptr(ws), ws.numel(), st), "score_hk"), args.reps)
        emit("score_hk " + name, a, m, 12 * n + 4 * B, {"finite": bool(torch.isfinite(out).all())})

    # ---- fused sampler updates (in-kernel Philox noise, both reflections fused)
    score = z
    gtab = torch.tensor([3.0], device=dev)
    if want("predictor"):
        a, m = timed(lambda: check(L.rd_pc_predictor_step(ptr(mean), ptr(score), None, ptr(gtab), -1e-3, math.sqrt(1e-3),
                                                          ptr(out), None, B, D, 7, 0, None, 0, 0, 0, st), "predictor"), args.reps)
        emit("predictor step (x,score in; x out; Philox)", a, m, 12 * n)
    partial = torch.empty(2 * ((B + 7) // 8), device=dev)
    nblk = C.c_int(0)
    if want("corrector"):
        a, m = timed(lambda: check(L.rd_pc_norms(ptr(score), None, ptr(partial), C.byref(nblk), B, D, 7, 0, None, 0, st),
                                   "norms"), args.reps)
        emit("corrector norms (grad in; Philox)", a, m, 4 * n)
        a, m = timed(lambda: check(L.rd_pc_corrector_apply(ptr(mean), ptr(score), None, ptr(partial), nblk.value, 0.01,
                                                           ptr(out), None, None, B, D, 7, 0, None, 0, st), "apply"), args.reps)
        emit("corrector apply (x,grad in; x out; Philox)", a, m, 12 * n)
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    return lines


if __name__ == "__main__":
    main()
