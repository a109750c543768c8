#!/usr/bin/env python
"""Device-resident timings of the five BASELINE.json configs on ONE B200 (bench.py carries the
driver's headline line for configs[1]; this script records the others for DESIGN.md / profiles/).

    python bench_configs.py [c1 c2 c3 c4 c5] > profiles/configs_rNN.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bayesian_quadrature_b200 import BQ, GaussianKernel, _lib, synthetic, util  # noqa: E402

PEAK = 37.156


def w_flop(ns, nc):
    n = ns + nc
    return n * n + ns * ns + 2 * (3 * n + 2 * ns) + 40


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts)), float(np.mean(ts))


def single(name, ns, na):
    dev = torch.device("cuda", 0)
    bq = synthetic.make_bq(BQ, GaussianKernel, ns)
    batch = bq._device_model().batch
    x_d = torch.from_numpy(synthetic.query_grid(ns, na)).to(dev)
    esm = torch.empty(1, na, dtype=torch.float64, device=dev)
    ev = torch.empty(na, dtype=torch.float64, device=dev)
    pair = torch.empty(2, dtype=torch.float64, device=dev)

    def step():
        batch.score_device(x_d, esm)
        batch.expected_var_device(0, esm, ev)
        batch.argmin_pair_device(ev, 0, pair)
    best, mean = timed(step)
    kbest, _ = timed(lambda: batch.score_device(x_d, esm))
    wf = w_flop(ns, bq.nc)
    p = pair.cpu().numpy()
    return {"config": name, "ns": ns, "nc": int(bq.nc), "na": na, "step_ms": best, "score_kernel_ms": kbest,
            "evals_per_s": na / (best * 1e-3), "tflops": wf * na / (kbest * 1e-3) * 1e-12,
            "frac_fp64_peak": wf * na / (kbest * 1e-3) * 1e-12 / PEAK, "argmin_index": int(p[1])}


def c4(n_hyper=1024, na=10 ** 5, ns=64):
    dev = torch.device("cuda", 0)
    bq = synthetic.make_bq(BQ, GaussianKernel, ns)
    hyp4 = synthetic.hyper_sets(n_hyper)
    hyp = np.zeros((n_hyper, 6))
    hyp[:, 0], hyp[:, 1], hyp[:, 3], hyp[:, 4] = hyp4.T
    opt = synthetic.options(ns)
    prior = np.tile([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]], (n_hyper, 1))
    batch = _lib.Batch(n_hyper, ns)
    args = (np.full(n_hyper, ns), np.full(n_hyper, bq.nc), np.tile(bq.x_s, (n_hyper, 1)), np.tile(bq.l_s, (n_hyper, 1)),
            np.tile(bq.x_c, (n_hyper, 1)), hyp, prior)
    t0 = time.perf_counter()
    info = batch.setup(*args, check_max=True)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    assert (info["status"] == 0).all()
    x_d = torch.from_numpy(synthetic.query_grid(ns, na)).to(dev)
    esm = torch.empty(n_hyper, na, dtype=torch.float64, device=dev)
    loss = torch.empty(na, dtype=torch.float64, device=dev)
    pair = torch.empty(2, dtype=torch.float64, device=dev)

    def step():
        batch.score_device(x_d, esm)
        batch.mean_neg_device(esm, loss)                 # marginal loss, sample order (bq.py:662)
        batch.argmin_pair_device(loss, 0, pair)
    best, mean = timed(step, reps=3, warm=1)
    kbest, _ = timed(lambda: batch.score_device(x_d, esm), reps=3, warm=1)
    wf = w_flop(ns, bq.nc)
    n = n_hyper * na
    p = pair.cpu().numpy()
    batch.close()
    return {"config": "C4", "ns": ns, "nc": int(bq.nc), "n_hyper": n_hyper, "na": na, "setup_ms": setup_ms, "step_ms": best,
            "score_kernel_ms": kbest, "evals_per_s": n / (best * 1e-3), "tflops": wf * n / (kbest * 1e-3) * 1e-12,
            "frac_fp64_peak": wf * n / (kbest * 1e-3) * 1e-12 / PEAK, "argmin_index": int(p[1])}


def c5(n_prob=16384, ns=128, na=4096):
    dev = torch.device("cuda", 0)
    opt = synthetic.options(ns)
    x_s0, _ = synthetic.observations(ns)
    x_s, l_s, x_c = np.tile(x_s0, (n_prob, 1)), np.empty((n_prob, ns)), np.zeros((n_prob, 16))
    nc = np.zeros(n_prob, dtype=np.int32)
    t0 = time.perf_counter()
    for p in range(n_prob):
        l_s[p] = synthetic.likelihood(ns, synthetic.problem_shift(p))(x_s0)
        rs = np.random.RandomState(synthetic.SEED + p)
        xc = rs.uniform(x_s0.min() - 2.0, x_s0.max() + 2.0, opt["n_candidate"])     # bq.py:974-978, per-problem stream
        util.filter_candidates(xc, x_s0, opt["candidate_thresh"])
        xc = np.sort(xc[~np.isnan(xc)])
        nc[p] = xc.size
        x_c[p, :xc.size] = xc
    host_ms = (time.perf_counter() - t0) * 1e3
    batch = _lib.Batch(n_prob, ns)
    hyp = np.tile(list(synthetic.PARAMS_TL) + list(synthetic.PARAMS_L), (n_prob, 1))
    prior = np.tile([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]], (n_prob, 1))
    t0 = time.perf_counter()
    info = batch.setup(np.full(n_prob, ns), nc, x_s, l_s, x_c, hyp, prior)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    assert (info["status"] == 0).all(), np.bincount(info["status"])
    x_d = torch.from_numpy(synthetic.query_grid(ns, na)).to(dev)
    esm = torch.empty(n_prob, na, dtype=torch.float64, device=dev)
    mins = torch.empty(n_prob, dtype=torch.float64, device=dev)
    idxs = torch.empty(n_prob, dtype=torch.int64, device=dev)
    neg = torch.empty_like(esm)

    def step():
        batch.score_device(x_d, esm)
        torch.neg(esm, out=neg)                          # loss = -esm (bq.py:660); plumbing, not a hot kernel
        batch.argmin_rows_device(neg, mins, idxs)        # per-problem deterministic choose_next
    best, mean = timed(step, reps=3, warm=1)
    kbest, _ = timed(lambda: batch.score_device(x_d, esm), reps=3, warm=1)
    n = n_prob * na
    wf = float(np.mean([w_flop(ns, c) for c in nc]))
    batch.close()
    return {"config": "C5 (one active-sampling round)", "ns": ns, "nc_mean": float(nc.mean()), "n_problems": n_prob, "na": na,
            "host_candidate_draw_ms": host_ms, "setup_ms": setup_ms, "step_ms": best, "score_kernel_ms": kbest,
            "evals_per_s": n / (best * 1e-3), "tflops": wf * n / (kbest * 1e-3) * 1e-12,
            "frac_fp64_peak": wf * n / (kbest * 1e-3) * 1e-12 / PEAK, "chosen_index_hist_head": np.bincount(idxs.cpu().numpy())[:4].tolist()}


def scattered(ns=256, na=10 ** 6):
    """Query points in arbitrary order through the host API (BQ.expected_Z_var): with and without the device pre-sort."""
    from bayesian_quadrature_b200 import BQ, GaussianKernel
    bq = synthetic.make_bq(BQ, GaussianKernel, ns)
    grid = synthetic.query_grid(ns, na)
    x = np.random.RandomState(0).permutation(grid)
    out = {"config": "scattered query points, host API", "ns": ns, "na": na}
    model = bq._device_model()
    for label, mode, pts in (("sorted_ms", 1, grid), ("shuffled_presort_ms", 1, x), ("shuffled_no_presort_ms", 0, x)):
        model.batch.set_presort(mode)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            bq.expected_Z_var(pts)
            ts.append((time.perf_counter() - t0) * 1e3)
        out[label] = min(ts)
    model.batch.set_presort(1)
    return out


def c5_rounds(n_prob=16384, ns=128, na=4096, rounds=20, device_resident=True):
    """C5 as BASELINE.json states it: 16384 independent problems, 128 observations each, 20 rounds of
    score -> deterministic argmin -> add_observation -> re-init, all problems on this GPU.  device_resident=True keeps
    the observations, candidate generators and candidate filter on the GPU (the likelihood is still evaluated by the
    caller on the host: it is the user's black box); False drives every round from host arrays."""
    from bayesian_quadrature_b200 import BatchBQ
    opt = synthetic.options(ns)
    x0, _ = synthetic.observations(ns)
    shifts = np.array([synthetic.problem_shift(p) for p in range(n_prob)])
    sp = synthetic.span(ns)

    def lik(x, sh):     # vectorised synthetic.likelihood over problems: x [P], sh [P, 3]
        npdf = lambda x, m, s: np.exp(-0.5 * ((x - m) / s) ** 2) / (np.sqrt(2 * np.pi) * s)
        return (0.5 * npdf(x, (-0.3 + sh[:, 0]) * sp, 0.16 * sp) + 0.3 * npdf(x, (0.4 + sh[:, 1]) * sp, 0.10 * sp)
                + 0.2 * npdf(x, (0.1 + sh[:, 2]) * sp, 0.3 * sp))
    l0 = np.stack([lik(np.full(n_prob, x), shifts) for x in x0], axis=1)
    if device_resident:          # kernels of this capacity class are loaded by a 32-problem batch outside the timed loop (as bench.py does)
        warm = BatchBQ(np.tile(x0, (32, 1)), l0[:32], synthetic.PARAMS_TL, synthetic.PARAMS_L, opt["n_candidate"],
                       opt["candidate_thresh"], opt["x_mean"], opt["x_var"], seed=synthetic.SEED, ns_reserve=rounds, device_resident=True)
        gw = torch.from_numpy(synthetic.query_grid(ns, na)).cuda()
        for _ in range(2):
            _, xw = warm.choose_next(gw, on_device=True)
            warm.add_observations(xw, torch.from_numpy(lik(xw.cpu().numpy(), shifts[:32])).cuda())
        warm.close()
    t0 = time.perf_counter()
    bb = BatchBQ(np.tile(x0, (n_prob, 1)), l0, synthetic.PARAMS_TL, synthetic.PARAMS_L, opt["n_candidate"],
                 opt["candidate_thresh"], opt["x_mean"], opt["x_var"], seed=synthetic.SEED, ns_reserve=rounds,
                 device_resident=device_resident)
    init_s = time.perf_counter() - t0
    grid = synthetic.query_grid(ns, na)
    if device_resident:
        grid = torch.from_numpy(grid).cuda()
    score_ms, host_ms = [], []
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for r in range(rounds):
        t0 = time.perf_counter()
        idx, x_next = bb.choose_next(grid, on_device=device_resident)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if device_resident:      # only the chosen points go to the host likelihood and its values come back
            l_next = torch.from_numpy(lik(x_next.cpu().numpy(), shifts)).cuda()
            bb.add_observations(x_next, l_next)
        else:
            bb.add_observations(x_next, lik(x_next, shifts))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        score_ms.append((t1 - t0) * 1e3)
        host_ms.append((t2 - t1) * 1e3)
    total_s = time.perf_counter() - t_all
    evals = n_prob * na * rounds
    bb.sync_host()
    out = {"config": "C5 (20 rounds, %s)" % ("device-resident" if device_resident else "host-driven"), "n_problems": n_prob, "ns_start": ns, "ns_end_min": int(bb.ns.min()),
           "ns_end_max": int(bb.ns.max()), "na": na, "rounds": rounds, "init_s": init_s, "total_s": total_s,
           "score_ms_per_round_first": score_ms[0], "score_ms_per_round_last": score_ms[-1],
           "update_ms_per_round_mean": float(np.mean(host_ms)), "evals_per_s_whole_loop": evals / total_s,
           "evals_per_s_scoring_only": evals / (sum(score_ms) * 1e-3), "Z_mean_first_problem": float(bb.Z_mean()[0])}
    bb.close()
    return out


if __name__ == "__main__":
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5", "c5r"]
    runs = {"c1": lambda: single("C1", 8, 200), "c2": lambda: single("C2", 64, 10 ** 6), "c3": lambda: single("C3", 256, 10 ** 7),
            "c4": c4, "c5": c5, "c5r": c5_rounds, "c5rh": lambda: c5_rounds(device_resident=False),
            "scat": scattered, "scat64": lambda: scattered(64)}
    for k in which:
        print(json.dumps(runs[k]()), flush=True)
