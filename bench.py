#!/usr/bin/env python
"""bench.py -- GTO-Halo samples/s of the 1000-step CFG reflected predictor-corrector sampler.

Workload (BASELINE.json configs[2] = "C3" per GPU; N GPUs x 8192 = configs[3] "C4" at N=8):
  NCSN++ 2D (8x9 latents, nf 64, ch_mult [1,2,2], 2 res blocks, attention at 8x9, conditional),
  random-init weights, RVESDE(0.01, 5, N=1000), eps 1e-5, Langevin corrector (snr 0.01, 1 step) +
  reflected Euler-Maruyama predictor, classifier-free guidance w=1.5, batch 8192 per GPU.

A "step" is ONE predictor-corrector iteration over the batch (2 guided-score evaluations = 4
network passes per sample + both fused updates) -- the unit the sampler repeats 999 times; all
iterations launch the same captured graph, so
    samples/s = n_gpus * B / (999 * ms_per_step + final all-gather).
The default --steps 999 times exactly one full sampler pass.  `e2e` times one complete call of the
public drop-in API (sampling.get_sampling_fn(...)(model, weight, class_labels)) with host labels,
the CPU-drawn prior copied H2D (as the reference does) and the samples copied back to the host.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "optimized-diffusion-model_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

_JSON_OUT = sys.stdout


def _isolate_stdout():
    """stdout carries exactly ONE JSON line per run: keep a private duplicate of the descriptor for that line and point
    fd 1 at stderr for everything else (NCCL prints its "NCCL version ..." banner on stdout)."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

SDE_N = 1000
ITERS_PER_PASS = SDE_N - 1          # sampling.py:330 skips the last grid point
FLOP_PER_SAMPLE_FORWARD = 207.374848e6   # 8x9, attention at 8 (torch FlopCounter, 2*MAC)
FLOP_PER_SAMPLE = FLOP_PER_SAMPLE_FORWARD * 2 * 2 * ITERS_PER_PASS  # CFG x2, corrector+predictor x2
METRIC = "GTO-Halo samples/sec (1000-step CFG PC sampler)"
UNIT = "samples/s"
# BASELINE.json configs[4] ("C5"): the scaled network; selected with --config c5 (the default line is C3)
C5 = {"H": 16, "W": 16, "nf": 256, "ch_mult": [1, 2, 2, 2], "attn": [16], "batch_per_gpu": 2048,
      "flop_per_sample_forward": 12608.866304e6,   # SURVEY.md App. A2 (torch FlopCounter on the reference model)
      "workload": ("C5: scaled NCSN++ 2D (16x16 reflected latents, nf256, ch_mult [1,2,2,2], 2 res blocks, attention at "
                   "16x16 = 256 tokens x 256 channels, CFG w=1.5) random-init, reflected PC sampler (Langevin snr 0.01 + "
                   "Euler-Maruyama), 1000 steps, batch 16384 over 8 GPUs = 2048 per GPU")}
CONFIG = {"name": "c3", "H": 8, "W": 9}

WORKLOAD = ("C3: GTO-Halo NCSN++ 2D (8x9, nf64, ch_mult [1,2,2], 2 res blocks, attn@8x9, CFG w=1.5) random-init, "
            "reflected PC sampler (Langevin snr 0.01 + Euler-Maruyama), 1000 steps, batch 8192 per GPU")


def use_config(name):
    """Switch the module-level workload description to BASELINE config `name` ("c3" default, "c5")."""
    global WORKLOAD, FLOP_PER_SAMPLE_FORWARD, FLOP_PER_SAMPLE
    if name == "c5":
        CONFIG.update(name="c5", H=C5["H"], W=C5["W"])
        WORKLOAD = C5["workload"]
        FLOP_PER_SAMPLE_FORWARD = C5["flop_per_sample_forward"]
        FLOP_PER_SAMPLE = FLOP_PER_SAMPLE_FORWARD * 2 * 2 * ITERS_PER_PASS


def model_config(corrector="langevin"):
    c5 = CONFIG["name"] == "c5"
    m = types.SimpleNamespace(
        name="ncsnpp", channels=1, image_size=CONFIG["H"], image_width=CONFIG["W"], num_classes=1, cond_drop_prob=0.5,
        conditional=True, init_scale=0.0, ema_rate=0.999, nf=C5["nf"] if c5 else 64, ch_mult=C5["ch_mult"] if c5 else [1, 2, 2],
        num_res_blocks=2, attn_resolutions=C5["attn"] if c5 else [CONFIG["H"]],
        resamp_with_conv=True, embedding_type="fourier", fourier_scale=16, skip_rescale=True, nonlinearity="swish",
        fir=False, fir_kernel=[1, 3, 3, 1], dropout=0.2, scale_by_sigma=False)
    s = types.SimpleNamespace(method="pc", predictor="euler_maruyama", corrector=corrector, denoiser="none", snr=0.01,
                              n_steps_each=1)
    return types.SimpleNamespace(model=m, sampling=s)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)),
                "bf16_burst": d.get("bf16_tflops", 1590.0), "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t_begin, self.continued = [], None, index, None, False

    def start(self):
        """Starts the poller (call it BEFORE the warm-up: nvidia-smi needs a few hundred ms to print its first row, longer
        than a 20-iteration timed region); rows are stamped with their arrival time."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def begin(self):
        """The timed region starts now: only rows that arrive from here on are reported."""
        self.t_begin = time.time()

    def rows_in_region(self):
        t0 = self.t_begin if self.t_begin is not None else 0.0
        return sum(1 for t, _ in self.rows if t >= t0)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.15)   # a row sampled under load just before the end is still in the pipe
        self.proc.terminate()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        rows = [r for t, r in self.rows if t0 <= t <= t_end + 0.15]
        window = "timed region" if not self.continued else "same load continued untimed until the first nvidia-smi row (region shorter than the polling period)"
        if not rows and self.rows:   # region shorter than the polling period: the row closest to it (the warm-up runs the same load)
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - t0))[1]]
            window = "closest sample to the timed region (warm-up, same load)"
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def static_config(B, world):
    """The workload description shared verbatim by both arms' JSON lines (run-dependent facts live outside `config`)."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
            "step": "one PC iteration (2 CFG score evaluations + fused corrector and predictor updates); "
                    "value = global_batch / (999*ms_per_step + all-gather)",
            "l2": "activations of one iteration (>5 GB at 2B=16384) exceed the 126 MB L2; no flush needed",
            "parallelism": f"batch-sharded x{world}, no collective in the loop"}


def _ref_available():
    from oracle import fetch_ref
    return fetch_ref.available() and fetch_ref.verify()


def stock_reference_pass(ref, B, N, device, threads=None, weights_seed=0):
    """ONE call of the unmodified reference's public API -- sampling.get_sampling_fn(config, sde, shape, eps, device)
    (model, weight=1.5, class_labels=...) -- with a forward-pre-hook on the model that only records timestamps
    (two network calls per PC iteration).  Returns (seconds for the whole call, timestamps, samples)."""
    if threads:
        torch.set_num_threads(threads)
    cfg = model_config()
    torch.manual_seed(weights_seed)
    model = ref.mutils.create_model(cfg).to(device).eval()   # untouched reference init, like our arm
    sde = ref.sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=N)
    fn = ref.sampling.get_sampling_fn(cfg, sde, (B, 1, CONFIG["H"], CONFIG["W"]), 1e-5, device)
    labels = torch.rand((B, 1), generator=torch.Generator().manual_seed(1)).to(device)
    stamps = []
    cuda = torch.device(device).type == "cuda"

    def hook(_m, _a):
        if cuda:
            torch.cuda.synchronize()
        stamps.append(time.perf_counter())
    h = model.register_forward_pre_hook(hook)
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, _ = fn(model, weight=1.5, class_labels=labels)
    if cuda:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    h.remove()
    return t1 - t0, stamps + [t1], x


def cpu_reference_c1(steps, warmup, passes=3, B=128, N=100):
    """BASELINE config C1 in full (batch 128, 100 steps, all host threads), `passes` times; the best pass is reported.
    ms_per_step = K consecutive PC iterations of that pass after W warm-up iterations (every iteration does the same
    work).  Falls back to the oracle port when oracle/_ref did not travel (kind says which)."""
    threads = os.cpu_count() or 1
    if _ref_available():
        from oracle import fetch_ref
        if steps + warmup > N - 1:
            N = steps + warmup + 1
        best = None
        with fetch_ref.imported() as ref:
            import contextlib, io
            for _ in range(passes):
                with contextlib.redirect_stdout(io.StringIO()):   # the stock constructor prints three [DEBUG] lines
                    total, st, x = stock_reference_pass(ref, B, N, "cpu", threads)
                ms = 1e3 * (st[2 * (warmup + steps)] - st[2 * warmup]) / steps
                if best is None or ms < best[0]:
                    best = (ms, total)
                assert bool(((x >= 0) & (x <= 1)).all())
        ms, total = best
        sample = (f"UNMODIFIED reference (oracle/_ref copy of Reflected-Diffusion, sha-verified), stock "
                  f"sampling.get_sampling_fn on the CPU, fp32, {threads} threads: BASELINE config C1 (batch {B}, "
                  f"{N} steps, CFG w=1.5; C1 in full is 128 / 100) best of {passes} passes = {total:.2f} s per pass; ms_per_step = {steps} consecutive PC "
                  f"iterations after {warmup} warm-up iterations of that pass; value extrapolates the per-iteration time to "
                  f"the 999 iterations of the 1000-step sampler; loadavg {os.getloadavg()[0]:.1f}")
        return {"ms": ms, "B": B, "kind": "reference", "cores": threads, "sample": sample, "c1_pass_seconds": total}
    # ---- fallback: the oracle port (bit-identical to the reference on the golden fixtures)
    from oracle import rd_oracle as O
    torch.set_num_threads(threads)
    cfg = O.NetConfig()
    sd = O.synth_state_dict(cfg, seed=0, degenerate=True)
    sched = O.VESchedule(0.01, 5.0, SDE_N, 1.0, 1e-5)
    scfg = O.SamplerConfig()
    g = torch.Generator().manual_seed(2)
    x = torch.rand((B, 1, 8, 9), generator=g)
    labels = torch.rand((B, 1), generator=g)
    ts = sched.timesteps()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            noise = torch.randn((2, B, 1, 8, 9), generator=g)
            t0 = time.perf_counter()
            vec_t = torch.ones(B) * ts[i]
            sig = sched.sigma(vec_t)
            grad = O.guided_score(x, sig, labels, 1.5, sd, cfg)
            x, _, _ = O.corrector_step(x, grad, noise[0], scfg.snr)
            score = O.guided_score(x, sig, labels, 1.5, sd, cfg)
            x, _ = O.predictor_step(x, score, sched.diffusion(vec_t), SDE_N, noise[1])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    sample = (f"oracle/_ref absent: ORACLE PORT of the reference sampler (fp32 PyTorch CPU), batch {B}, {steps} PC iterations "
              f"after {warmup} warm-up, extrapolated to 999 iterations; loadavg {os.getloadavg()[0]:.1f}")
    return {"ms": ms, "B": B, "kind": "port", "cores": threads, "sample": sample, "c1_pass_seconds": None}


def run_reference(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference_c1(args.steps, args.warmup, passes=args.ref_passes, B=args.ref_batch, N=args.ref_grid)
    value = r["B"] / (ITERS_PER_PASS * r["ms"] / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": static_config(args.batch, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                             "batch": r["B"], "c1_pass_seconds": r["c1_pass_seconds"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def eager_gpu_baseline(dev, B, iters=20):
    """The stock reference sampler on the SAME GPU in eager fp32 (TF32 off): what a user of sampling.py gets on this card
    today.  One warm-up call (cuDNN autotune, allocator) then one timed call of `iters` PC iterations (an RVESDE with
    N = iters + 1 grid points runs exactly `iters` iterations; the work per iteration does not depend on the grid)."""
    if not _ref_available():
        return {"unavailable": "oracle/_ref did not travel"}
    from oracle import fetch_ref
    import contextlib, io
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with fetch_ref.imported() as ref, contextlib.redirect_stdout(io.StringIO()):
            stock_reference_pass(ref, B, 4, dev)
            total, st, x = stock_reference_pass(ref, B, iters + 1, dev)
        ms = 1e3 * (st[-1] - st[0]) / iters
        inside = bool(((x >= 0) & (x <= 1)).all())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    return {"value": B / (ITERS_PER_PASS * ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": B, "iterations": iters,
            "kind": "reference", "all_samples_inside_cube": inside,
            "what": "unmodified reference sampling.get_sampling_fn on this GPU, eager PyTorch fp32 (cudnn/matmul TF32 off), "
                    "same model / SDE / guidance as our arm; per-iteration time extrapolated to 999 iterations"}


def _conv_source_sha():
    import hashlib
    h = hashlib.sha256()
    for f in ("conv_gemm.cu", "rd_ptx.cuh"):
        with open(os.path.join(PKG, "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def conv_traffic():
    """`roofline.traffic`: DRAM bytes per conv launch from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
    capture of this workload (tools/run_ncu_traffic.sh -> tools/summarise_ncu.py writes profiles/*_conv_traffic.json and
    stamps it with the sha of the kernel sources it profiled).  A capture of OTHER sources is refused: traffic = null."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_conv_traffic.json")))
    sha = _conv_source_sha()
    for path in reversed(files):
        try:
            with open(path) as f:
                js = json.load(f)
        except Exception:
            continue
        if js.get("kernel_source_sha") == sha:
            return js.get("dram_bytes_per_launch"), "%s (ncu dram__bytes_read+write per conv launch, B=8192, kernel sources %s)" % (
                os.path.relpath(path, ROOT), sha)
    return None, "no ncu capture of the current kernel sources (%s) under profiles/ -- stale captures are not reported" % sha


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=ITERS_PER_PASS)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8192, help="samples per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c5"],
                    help="BASELINE config: c3 = GTO-Halo 8x9 (the metric's configuration, default), c5 = scaled nf256 16x16")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"], help="network plan precision (default bf16)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total batch split evenly over the GPUs (BASELINE config C4 = 65536)")
    ap.add_argument("--weight", type=float, default=1.5, help="classifier-free guidance weight (C3 sweep)")
    ap.add_argument("--per-sample-weight", action="store_true", help="w_b = weight * U[0,1] per sample (run_train.py:173)")
    ap.add_argument("--ref-batch", type=int, default=128, help="reference arm: batch (C1 = 128)")
    ap.add_argument("--ref-grid", type=int, default=100, help="reference arm: sampler grid points (C1 = 100)")
    ap.add_argument("--ref-passes", type=int, default=3, help="reference arm: passes, best one reported")
    ap.add_argument("--no-eager-gpu", action="store_true", help="skip the stock-reference-on-this-GPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="skip the reflect / score_hk / fused-update HBM microbench")
    args = ap.parse_args()
    use_config(args.config)
    if args.config == "c5" and args.batch == 8192:
        args.batch = C5["batch_per_gpu"]
    _isolate_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    import cube
    import sampling
    import sde_lib
    from models import utils as mutils
    from rdb200 import dist as rdd

    B = args.batch
    scaling = "weak"
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of GPUs")
        B, scaling = args.global_batch // world, "strong"
    cfg = model_config()
    torch.manual_seed(0)
    model = mutils.create_model(cfg).to(dev).eval()   # untouched reference init: timing is weight-independent
    sde = sde_lib.RVESDE(sigma_min=0.01, sigma_max=5.0, N=SDE_N)
    labels_host = torch.rand((B, 1), generator=torch.Generator().manual_seed(1 + rank)).pin_memory()
    labels = labels_host.to(dev, non_blocking=True)
    H, W = CONFIG["H"], CONFIG["W"]
    if args.precision:
        model.rd_set_precision(args.precision)
    x0 = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(2 + rank)).to(dev)
    eng = model.rd_sampler_engine(B, H, W, dev, sde, 1e-5, 0.01, 1, cfg=True)
    seed = rdd.philox_seed_for_rank(3, rank)

    weight = args.weight
    if args.per_sample_weight:
        weight = (args.weight * torch.rand((B,), generator=torch.Generator().manual_seed(5 + rank))).to(dev)

    def run_iters(n, start=0):
        return eng.sample(x0, labels, weight, seed=seed, use_graph=True, n_iter=n, start_step=start)

    # ---- warm-up (also builds and instantiates the CUDA graph)
    clocks = ClockSampler(local_rank)
    clocks.start()   # polling from before the warm-up; only rows from the timed region are reported
    run_iters(args.warmup)
    torch.cuda.synchronize(dev)
    # ---- timed region: exactly K iterations, CUDA events on the engine's stream, barrier + sync on both sides
    K = args.steps
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = 0
    e0.record(torch.cuda.current_stream(dev))
    while done < K:  # the step counter indexes the sigma grid, so long runs wrap in chunks of one full pass
        n = min(K - done, ITERS_PER_PASS)
        xs = run_iters(n)
        done += n
    e1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    # a timed region shorter than nvidia-smi's polling period has no clock row yet: keep the SAME load running (untimed)
    # until one arrives, so that the reported clocks are clocks under this load
    t_wait = time.time()
    while clocks.proc is not None and clocks.rows_in_region() == 0 and time.time() - t_wait < 2.0:
        run_iters(min(5, ITERS_PER_PASS))
        torch.cuda.synchronize(dev)
        clocks.continued = True
    clk = clocks.stop()
    inside = bool(cube.inside(xs).all())
    # final all-gather (once per sampler pass)
    ag_ms = 0.0
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rdd.all_gather_batch(xs, B * world)
        torch.cuda.synchronize(dev)
        a0.record(); full = rdd.all_gather_batch(xs, B * world); a1.record()
        torch.cuda.synchronize(dev)
        ag_ms = a0.elapsed_time(a1)
        assert full.shape[0] == B * world
        t = torch.tensor([ms_total, ag_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ag_ms = float(t[0]), float(t[1])
    launches_per_iter = eng.launches_per_iter()
    ms_per_step = ms_total / K
    pass_s = (ITERS_PER_PASS * ms_per_step + ag_ms) / 1e3
    value = world * B / pass_s

    # ---- roofline of the dominant kernel (tcgen05 conv/NIN implicit GEMM): per-op CUDA events, eager replay
    pk = peaks()
    roof = None
    if rank == 0:
        eng.run_plan()
        torch.cuda.synchronize(dev)
        reps = 3
        per_op = [0.0] * eng.n_ops
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(eng.n_ops + 1)]
            ev[0].record()
            for i in range(eng.n_ops):
                eng.run_ops(i, 1)
                ev[i + 1].record()
            torch.cuda.synchronize(dev)
            for i in range(eng.n_ops):
                per_op[i] += ev[i].elapsed_time(ev[i + 1]) / reps
        conv_ms = sum(t for t, n in zip(per_op, eng.op_names) if eng.op_kinds[n] == "conv")
        fwd_ms = sum(per_op)
        conv_flop = eng.conv_flops_per_sample * eng.B2
        achieved = conv_flop / (conv_ms / 1e3) / 1e12
        n_conv = sum(1 for n in eng.op_names if eng.op_kinds[n] == "conv")
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (all %d launches of one guided-score evaluation)" % n_conv,
                "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
                "traffic": conv_traffic()[0], "traffic_source": conv_traffic()[1],
                "peak_source": pk["source"] + " (sustained)",
                "avg_launch_ms": conv_ms / n_conv, "share_of_forward": conv_ms / fwd_ms,
                "algorithmic_flop_per_launch_avg": conv_flop / n_conv,
                "sampler_frac": value / world * FLOP_PER_SAMPLE / (pk["bf16_tflops"] * 1e12),
                "forward_ms_by_kind": {k: round(sum(t for t, n in zip(per_op, eng.op_names) if eng.op_kinds[n] == k), 4)
                                       for k in sorted(set(eng.op_kinds.values()))}}
        if model.rd_precision == "fp32":
            # fp32 activations put the network below the ridge point: the bound to read this plan against is
            # HBM, 1.32 MB of fp32 activation traffic per sample-forward, not the bf16 tensor peak
            roof["fp32_plan_hbm_bound_samples_per_s"] = pk["hbm_gbs"] * 1e9 / (1.32e6 * 2 * 2 * ITERS_PER_PASS) if args.config == "c3" else None
            roof["fp32_plan_tensor_work"] = "3 MMAs per k-step (lo*hi + hi*lo + hi*hi): achieved counts algorithmic flops once"

    # ---- e2e through the public drop-in API with host buffers
    e2e = None
    if not args.no_e2e:
        fn = sampling.get_sampling_fn(cfg, sde, (B, 1, H, W), 1e-5, dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        lab = labels_host.to(dev, non_blocking=True)                  # H2D from pinned memory
        samples, nfe = fn(model, weight=weight, class_labels=lab, rd_seed=seed)   # prior drawn on the CPU and copied H2D inside
        if world > 1:
            samples = rdd.all_gather_batch(samples, B * world)
        host = samples.cpu()                                           # D2H of the result
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        assert bool(((host >= 0) & (host <= 1)).all())
        e2e = {"value": world * B / dt, "unit": UNIT, "h2d_bytes_per_step": int(B * 4 + B * H * W * 4),
               "d2h_bytes_per_step": int(host.numel() * 4 // (world if world > 1 else 1)), "seconds_per_pass": dt,
               "call": "sampling.get_sampling_fn(config, sde, shape, eps, device)(model, weight=1.5, class_labels=...)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the HBM-bound kernels of the path (BASELINE config C2 shape, [2^20,1,8,9]): achieved GB/s of algorithmic
    # bytes against the measured copy bandwidth, timed live with CUDA events (tools/bench_c2.py)
    hbm = None
    if not args.no_c2 and args.config == "c3":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_c2
            torch.cuda.empty_cache()   # fresh, contiguous allocations for the four 302 MB buffers (blocks recycled from the
            #                            sampler passes above cost `reflect` a quarter of its bandwidth)
            warm = torch.rand((1 << 24,), device=dev)
            for _ in range(50):   # the CPU-side bookkeeping above leaves the GPU idle; bring clocks back up first
                warm.mul_(1.0001)
            torch.cuda.synchronize(dev)
            del warm
            rows = bench_c2.run(types.SimpleNamespace(log2B=20, reps=10, only=None, out=None), dev=dev, quiet=True)
            hbm = [{"kernel": r["case"], "ms": round(r["ms_avg"], 4), "GBps": round(r["GBps"], 1), "frac": round(r["frac"], 4)}
                   for r in rows if r["case"] in ("reflect", "score_hk sigma=0.1", "score_hk sigma~logU[0.01,5]",
                                                  "predictor step (x,score in; x out; Philox)",
                                                  "corrector norms (grad in; Philox)", "corrector apply (x,grad in; x out; Philox)")]
        except Exception as e:  # the headline line must not depend on the side measurement
            hbm = [{"error": repr(e)}]

    cpu = None
    if not args.no_cpu_baseline and world == 1 and args.config == "c3":
        r = cpu_reference_c1(20, 5, passes=1)   # bounded sample: one C1 pass (batch 128, 100 steps) on the host cores
        cpu = {"value": r["B"] / (ITERS_PER_PASS * r["ms"] / 1e3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "batch": r["B"], "c1_pass_seconds": r["c1_pass_seconds"]}
    eager = None
    if not args.no_eager_gpu and world == 1:
        try:
            torch.cuda.empty_cache()
            eager = eager_gpu_baseline(dev, B)
        except Exception as e:   # a side measurement must not take the headline line down
            eager = {"error": repr(e)}

    config = static_config(B, world)
    if args.weight != 1.5 or args.per_sample_weight:
        config["guidance"] = ("per-sample w = %g*U[0,1]" % args.weight) if args.per_sample_weight else ("w = %g" % args.weight)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if model.rd_precision == "bf16" else "f32 (split-bf16 tensor-core operands)", "data": "synthetic",
            "config": config,
            "run": {"noise": "in-kernel Philox", "cuda_graph": True, "all_gather_ms": ag_ms, "all_samples_inside_cube": inside},
            "clocks": clk, "roofline": roof, "hbm_kernels": {"shape": "[2^20,1,8,9] fp32 (C2)", "peak_GBps": pk["hbm_gbs"],
                                                              "peak_source": pk["source"], "rows": hbm},
            "cpu_baseline": cpu, "eager_gpu_baseline": eager, "e2e": e2e, "gpu_launches": launches_per_iter * K}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
