#!/usr/bin/env python
"""bench.py — expected_Z_var query-point evaluations per second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (config.workload): BASELINE.json configs[1] — 1-D BQ, 64 observations, expected_Z_var over
a 10^6-point grid (SURVEY.md §8(d) generator).  A *step* is one pass of the hot path over the
grid: the scoring kernel (esm for every point, expected variance Zm^2 + Zv - esm in its epilogue,
per-CTA minima) and one small kernel that reduces them to the deterministic (min, first index)
choose_next needs and — at N > 1 — exchanges the ranks' pairs over peer-mapped memory (NCCL
all-gather if symmetric memory is unavailable); the result lands in page-locked host memory.  At
N > 1 every rank scores 10^6 points of an N x 10^6 grid (weak scaling, block-cyclic shards).  The
per-hyper-set setup kernel (Gram, Cholesky, Z_mean, Z_var — amortised over the grid)
is timed separately (setup_ms).

Timing: W warm-up steps, then K steps enqueued back to back between a barrier + synchronize on both
sides, each bracketed by CUDA events on the launching stream; cold L2 through eight rotating 24 MB
input/output sets (192 MB > 126 MB of L2); max over ranks.  `value` is device-resident throughput
(inputs already in HBM); `e2e` is the same pass through the public host API (BQ.expected_Z_var) with
page-locked host buffers, the host-to-device and device-to-host traffic inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NS = 64
NA = 10 ** 6
METRIC = "expected_Z_var candidate evals/sec"
UNIT = "evals/s"


def w_flop(ns, nc):
    """Algorithmic FP64 flops per evaluation."""
    n = ns + nc
    return n * n + ns * ns + 2 * (3 * n + 2 * ns) + 40


def w_exp(ns, nc):
    return ns + nc + ns + 4


def fp64_peak_tflops():
    """FP64 roofline denominator.  MEASURED_PEAKS.json carries no FP64 figure, so the peak is this
    repo's own microbenchmark (bench_micro/fp64_peaks.cu, DMMA m8n8k4 on all SMs), committed as
    profiles/fp64_peaks_r01.json."""
    p = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        with open(p) as fh:
            d = json.load(fh)
        return max(v for k, v in d.items() if k.startswith("dmma") or k.startswith("dfma")), "measured:profiles/fp64_peaks_r01.json"
    except Exception:
        return 37.2, "nominal:148 SM x 64 DFMA/clk x 1.965 GHz"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the scoring kernel on the bench workload, from the
    committed `ncu --set full` capture (profiles/ncu_score_traffic.json, written by profiles/summarize.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_score_traffic.json")) as fh:
            d = json.load(fh)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler(object):
    """`nvidia-smi -lms 200` in its own process for the duration of the timed regions (the profiling
    recipe's clocks line); parsed when stopped.  A separate process, so it never holds this process's GIL."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = "/tmp/bq_b200_clocks_%d_%d.csv" % (os.getpid(), index)
        self.fh = open(self.path, "w")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def start(self):
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.fh.close()
        rows = [[c.strip() for c in line.split(",")] for line in open(self.path) if line.strip()]
        os.unlink(self.path)

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 3 + i and r[3 + i].lower() == "active"})
        # "under load": samples at or above half the maximum clock (idle samples between phases are dropped)
        load = [v for v in sm if mx and v >= 0.5 * max(mx)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def build_problem():
    """C2 inputs.  The candidate draw uses the host RNG exactly like BQ.init (bq.py:967-991)."""
    from bayesian_quadrature_b200 import synthetic
    from bayesian_quadrature_b200.bq import BQ
    from bayesian_quadrature_b200.gp import GaussianKernel
    bq = synthetic.make_bq(BQ, GaussianKernel, NS)
    return bq


# ------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    lo, hi, na_total = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import warnings
    warnings.simplefilter("ignore")
    from oracle import build_ref
    from bayesian_quadrature_b200 import synthetic
    bqmod, gp = build_ref.import_reference()
    bq = synthetic.make_bq(bqmod.BQ, gp.GaussianKernel, NS)
    x_a = synthetic.query_grid(NS, na_total)[lo:hi]
    t0 = time.perf_counter()
    ev = bq.expected_Z_var(x_a)
    return time.perf_counter() - t0, float(ev.sum())


def reference_pass(cores, pts_per_core):
    """One bounded sample of the C2 workload on the reference's own CPU implementation: `cores`
    forked processes each score a contiguous slice of the 10^6 grid (the reference itself is single
    threaded; this is the most favourable multi-core use of it)."""
    import multiprocessing as mp
    starts = np.linspace(0, NA - pts_per_core, cores).astype(np.int64)
    jobs = [(int(s), int(s) + pts_per_core, NA) for s in starts]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    inner = max(r[0] for r in res)      # slowest worker's scoring time (excludes import/setup)
    return cores * pts_per_core / inner, inner, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build_ref
    kind = "reference" if build_ref.built() else "port"
    cores = os.cpu_count() or 1
    pts = 1500
    if kind == "port":
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.py where /root/reference exists")
    for _ in range(min(args.warmup, 1)):
        reference_pass(cores, 200)
    vals, times = [], []
    for _ in range(args.steps):
        v, inner, _ = reference_pass(cores, pts)
        vals.append(v)
        times.append(inner)
    v = float(np.mean(vals))
    sample = "%d forked processes x %d contiguous points of the 10^6 grid (ns=64), OPENBLAS_NUM_THREADS=1" % (cores, pts)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 1-D BQ, ns=64, expected_Z_var over a 10^6-point grid (bounded sample per step)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from bayesian_quadrature_b200 import synthetic, _lib
    from bayesian_quadrature_b200 import dist as bqdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    bq = build_problem()
    bq.device = local
    nc = bq.nc
    na_total = NA * world
    grid = synthetic.query_grid(NS, na_total)
    # block-cyclic shards (blocks of CYC points dealt round-robin): every rank sees the same mix of near- and far-field
    # points; with contiguous shards the ranks holding the observed region would do several times the work of the others
    CYC = 10000
    shard = bqdist.cyclic_shard(grid, world, rank, CYC) if os.environ.get("BQB_SHARD", "cyclic") == "cyclic" else grid[rank * NA:(rank + 1) * NA]
    cyc = CYC if os.environ.get("BQB_SHARD", "cyclic") == "cyclic" else 0

    # ---- setup (amortised over the grid, timed separately): wall time of a full device-model rebuild
    # (buffer allocation + uploads + setup kernel + header read-back)
    bq._device_model()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bq._invalidate_device()
    model = bq._device_model()
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    batch = model.batch

    x_d = torch.from_numpy(shard).to(dev)
    esm = torch.empty(1, NA, dtype=torch.float64, device=dev)
    em = torch.empty(1, NA, dtype=torch.float64, device=dev)
    st = torch.empty(1, NA, dtype=torch.int32, device=dev)
    evv = torch.empty(NA, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2
    pair = torch.empty(2, dtype=torch.float64, device=dev)
    pairs = torch.empty(world, 2, dtype=torch.float64, device=dev) if world > 1 else None

    # cross-rank exchange inside the reduction kernel (peer-mapped symmetric memory over NVLink); NCCL all-gather otherwise
    exch = None if os.environ.get("BQB_EXCHANGE", "p2p") == "nccl" else bqdist.PairExchange.create(dev)

    def step():
        # esm (bq.py:379-402), expected variance (bq.py:374-377) and the (min, first global index) over all shards:
        # the scoring kernel with its fused epilogue + one tiny reduce-and-exchange launch; result in page-locked memory
        if exch is not None:
            return exch.step(batch, x_d, esm[0], evv, rank * NA, cyclic_block=cyc)
        batch.choose_step_device(x_d, esm, evv, pair, offset=0 if cyc else rank * NA)
        if cyc:                                                    # local -> global index of a block-cyclic shard
            i = pair[1]
            pair[1] = (torch.floor(i / cyc) * world + rank) * cyc + torch.remainder(i, cyc)
        if world > 1:
            dist.all_gather_into_tensor(pairs.view(-1), pair)      # the path's only collective: W pairs of 16 B
            return bqdist.combine_argmin(pairs.cpu().numpy())
        p = pair.cpu().numpy()                                     # the step's device->host read (16 B)
        return float(p[0]), int(p[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- kernel-only timing of the dominant kernel (CUDA events on the launching stream)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    for _ in range(max(args.steps, 5)):
        flush.zero_()
        k0.record()
        batch.score_device(x_d, esm, em, st)
        k1.record()
        torch.cuda.synchronize()
        kern_ms.append(k0.elapsed_time(k1))
    kern_ms_avg = float(np.mean(kern_ms))
    # executed work of one launch (device counter of DMMA instructions), and the same launch with the band skipping
    # switched off (cut-off = inf: the dense border-update algorithm SURVEY §8(d)'s W_flop counts)
    batch.work_counter(True)
    batch.score_device(x_d, esm, em, st)
    torch.cuda.synchronize()
    dmma_per_launch = batch.work_counter(False)
    batch.set_cutoff(float("inf"))
    dense_ms = []
    for _ in range(max(args.steps, 5)):
        flush.zero_()
        k0.record()
        batch.score_device(x_d, esm, em, st)
        k1.record()
        torch.cuda.synchronize()
        dense_ms.append(k0.elapsed_time(k1))
    dense_ms_avg = float(np.mean(dense_ms[1:]))
    batch.set_cutoff(72.0)

    # ---- timed region: exactly K steps, L2 flushed between steps (outside the per-step events)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = batch.launch_count
    step_ms = []
    l2_note = "flushed between timed steps (256 MiB memset)"
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    result = None
    if exch is not None:
        # the K steps are enqueued back to back (L2 flush between them, outside the per-step events); the ranks stay in
        # step on the device through the exchange kernel, so no host wake-up jitter is billed to a step
        # Cold L2 without a flush kernel in the loop (its run-to-run jitter would leak into the other ranks' steps through the
        # exchange): NSETS rotating sets of input / output vectors, 24 MB each, 192 MB > 126 MB of L2 in total.
        NSETS = 8
        sets = [(x_d.clone(), torch.empty_like(esm[0]), torch.empty_like(evv)) for _ in range(NSETS)]
        flush.zero_()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k, (s0, s1) in enumerate(evs):
            xs_k, esm_k, ev_k = sets[k % NSETS]
            s0.record()
            exch.step_async(batch, xs_k, esm_k, ev_k, rank * NA, cyclic_block=cyc)
            s1.record()
        result = exch.result()
        step_ms = [s0.elapsed_time(s1) for s0, s1 in evs]
        l2_note = "%d rotating input/output sets (%d MiB > 126 MB L2), no flush inside the timed loop" % (NSETS, NSETS * 24)
    else:
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record()
            result = step()
            e1.record()
            torch.cuda.synchronize()
            step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = batch.launch_count - launches0
    total_ms = float(np.sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = na_total / (ms_per_step * 1e-3)

    # ---- end to end through the public host API: numpy in (pinned), numpy out
    x_pin = torch.from_numpy(shard).pin_memory()
    x_host = x_pin.numpy()
    keep = [bq.expected_Z_var(x_host) for _ in range(3)]     # steady state: the caller still holds the previous result
    del keep
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev_host = bq.expected_Z_var(x_host)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    barrier()
    t = torch.tensor([float(np.sum(e2e_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = na_total / (float(t.item()) / args.steps * 1e-3)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = fp64_peak_tflops()
        wf = w_flop(NS, nc)
        achieved = wf * NA / (kern_ms_avg * 1e-3) * 1e-12
        executed = dmma_per_launch * 512 / (kern_ms_avg * 1e-3) * 1e-12
        dense = wf * NA / (dense_ms_avg * 1e-3) * 1e-12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 1-D BQ, ns=%d nc=%d, expected_Z_var over a 10^6-point grid per GPU (%d points total)"
                                   % (NS, nc, na_total),
                       "l2": l2_note, "parallelism": "x_a sharded, %d rank(s), %s" % (world, "block-cyclic shards of %d points" % CYC if cyc else "contiguous shards"),
                       "exchange": "p2p stores from the reduction kernel (symmetric memory)" if exch is not None else "nccl all-gather"},
            "roofline": {"bound": "tensor", "pipe": "FP64 DMMA (mma.m8n8k4.f64; shares the FP64 datapath with DFMA)", "kernel": "bq_score_kernel<KS=16,NT=2,WARPS=8,MINB=2,STREAM=0,TABN=2048,ALIGN=1>", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic(), "traffic_unit": "bytes per launch (ncu dram read + write)",
                         "peak_source": peak_src,
                         "flop_per_eval": wf, "exp_per_eval": w_exp(NS, nc), "kernel_ms": kern_ms_avg,
                         "note": "achieved = SURVEY 8(d) algorithmic flops (dense border update) / kernel time; the kernel skips "
                                 "cross-kernel blocks below e^-72 of each point's leading element (band skipping, DESIGN.md 4.1), "
                                 "so it executes fewer flops than that count: see executed_* and dense_*",
                         "executed_dmma_per_launch": dmma_per_launch, "executed_tflops": executed, "executed_frac": executed / peak,
                         "dense_kernel_ms": dense_ms_avg, "dense_tflops": dense, "dense_frac": dense / peak,
                         "hbm_bytes_per_eval": 28, "hbm_gbs": 28 * NA / (kern_ms_avg * 1e-3) * 1e-9},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * NA, "d2h_bytes_per_step": bq._last_d2h_bytes},
            "gpu_launches": int(launches), "setup_ms": setup_ms,
            "clocks": clocks, "argmin": {"min": result[0], "index": result[1]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import build_ref
            if build_ref.built():
                cores = os.cpu_count() or 1
                pts = args.cpu_points
                v, inner, wall = reference_pass(cores, pts)
                v1, inner1, wall1 = reference_pass(1, 12000)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": "%d forked processes x %d contiguous points of the 10^6 grid, %.1f s (%.1f s scoring)"
                                                  % (cores, pts, wall, inner),
                                        "single_process_value": v1,
                                        "single_process_sample": "1 process, OPENBLAS_NUM_THREADS=1, 12000 points, %.1f s" % wall1}
            else:
                line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-points", type=int, default=40000,
                    help="points per host core of the cpu_baseline sample (default: ~10 s of CPU work per core)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
