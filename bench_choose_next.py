#!/usr/bin/env python
"""Wall time of BQ.choose_next end to end (sample_hypers -> marginal loss over the samples -> argmin; reference
bayesian_quadrature/bq.py:659-681) on this package's CUDA path, next to the unmodified reference (oracle/_ref) where it
finishes in seconds.  One JSON line per case:

    python bench_choose_next.py > profiles/choose_next_r02.jsonl

Cases: (a) BASELINE.md section 2's row -- the reference's test fixture (9 observations), 200 query points, 20 samples;
(b) ns = 64, 1500 points, 6 samples (the choose_c2 fixture); (c) C4-class: ns = 64, 10^5 points, 1024 samples -- the reference
is not run at that size (extrapolated from its per-point time)."""
import json
import os
import sys
import time
import warnings

import numpy as np
import scipy.stats

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")


def fixture(BQ, GaussianKernel):
    np.random.seed(8728)
    x = np.linspace(-5, 5, 9)
    bq = BQ(x, scipy.stats.norm.pdf(x, 0, 1), n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5, kernel=GaussianKernel,
            optim_method="L-BFGS-B")
    bq.init(params_tl=(15, 2, 0.), params_l=(0.2, 1.3, 0.))
    return bq


def timed(make, x_a, n, reps):
    out = []
    for _ in range(reps):
        bq = make()
        t0 = time.perf_counter()
        x = bq.choose_next(x_a, n=n, params=["h", "w"])
        out.append((time.perf_counter() - t0, float(x)))
    return out


def phases(make, x_a, n):
    """Where the product's time goes: the sampler, the batched marginal loss, the rest."""
    import torch
    bq = make()
    bq.choose_next(x_a, n=2, params=["h", "w"])             # warm the device model / allocator
    bq = make()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    htl, hl = bq.sample_hypers(["h", "w"], n=n, nburn=1)
    t1 = time.perf_counter()
    loss, batch = bq.marginal_loss(x_a, htl, hl, ["h", "w"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    batch.close()
    return {"sample_hypers_s": t1 - t0, "marginal_loss_s": t2 - t1}


def main():
    import logging
    logging.disable(logging.CRITICAL)
    from bayesian_quadrature_b200 import BQ, GaussianKernel, synthetic
    from oracle import build_ref
    have_ref = build_ref.built()
    if have_ref:
        bqmod, gp = build_ref.import_reference()
    cases = [
        ("fixture: 9 observations, 200 points, 20 samples (BASELINE.md section 2)", lambda C, K: fixture(C, K), np.linspace(-10, 10, 200), 20),
        ("ns = 64, 1500 points, 6 samples (seed 8738)", lambda C, K: synthetic.make_bq(C, K, 64, seed=8738), synthetic.query_grid(64, 1500), 6),
    ]
    for name, make, x_a, n in cases:
        mk = lambda: make(BQ, GaussianKernel)
        timed(mk, x_a, min(n, 4), 1)                          # warm-up (library load, allocator)
        mine = timed(mk, x_a, n, 3)
        line = {"case": name, "na": int(x_a.size), "n_samples": n, "cuda_s": min(t for t, _ in mine), "cuda_chosen": mine[0][1]}
        line.update(phases(mk, x_a, n))
        if have_ref:
            ref = timed(lambda: make(bqmod.BQ, gp.GaussianKernel), x_a, n, 1)
            line.update(reference_s=ref[0][0], reference_chosen=ref[0][1], same_point=bool(ref[0][1] == mine[0][1]),
                        speedup=ref[0][0] / line["cuda_s"])
        print(json.dumps(line), flush=True)

    # (c) C4 class: the marginal loss and argmin over 1024 GIVEN hyper-parameter sets (synthetic.hyper_sets: the sampler's window
    # takes most ns = 64 chains through numerically singular K_tl -- tests/golden/make_golden.py section 8 -- where a 1024-sample chain
    # does not terminate in reasonable time in the reference either), and the sampler's cost as log-density evaluations per second
    import torch
    bq = synthetic.make_bq(BQ, GaussianKernel, 64)
    hyp = synthetic.hyper_sets(1024)
    x_a = synthetic.query_grid(64, 10 ** 5)
    for _ in range(2):
        t0 = time.perf_counter()
        loss, batch = bq.marginal_loss(x_a, hyp[:, :2], hyp[:, 2:], ["h", "w"])
        mn, idx = batch.argmin_device(loss)
        torch.cuda.synchronize()
        t_loss = time.perf_counter() - t0
        batch.close()
    line = {"case": "C4 class: ns = 64, 10^5 points, marginal loss + argmin over 1024 given hyper-parameter sets", "na": 10 ** 5, "n_samples": 1024,
            "cuda_s": t_loss, "argmin_index": int(idx)}
    f = bq._make_llh_params(["h", "w"])
    rs = np.random.RandomState(0)
    props = np.stack([rs.uniform(13.5, 15.5, 400), rs.uniform(1.6, 2.2, 400), rs.uniform(0.15, 0.6, 400), rs.uniform(1.0, 1.5, 400)], axis=1)
    f(props[0])
    t0 = time.perf_counter()
    vals = [f(p) for p in props]
    line["log_density_evals_per_s"] = 400 / (time.perf_counter() - t0)
    bq.log_lh_batch(props[:, :2], props[:, 2:], ["h", "w"])     # (loads the many-instance variant of the setup kernel)
    t0 = time.perf_counter()
    got = bq.log_lh_batch(props[:, :2], props[:, 2:], ["h", "w"])
    line["log_density_batched_evals_per_s"] = 400 / (time.perf_counter() - t0)
    line["batched_equals_sequential"] = bool(np.allclose(got, vals, rtol=1e-12, atol=0))
    if have_ref:
        rb = synthetic.make_bq(bqmod.BQ, gp.GaussianKernel, 64)
        fr = rb._make_llh_params(["h", "w"])
        fr(props[0])
        t0 = time.perf_counter()
        rv = [fr(p) for p in props[:100]]
        line["reference_log_density_evals_per_s"] = 100 / (time.perf_counter() - t0)
        line["log_density_max_rel_diff_vs_reference"] = float(np.max(np.abs(np.array(vals[:100]) - np.array(rv)) / np.abs(rv)))
        t0 = time.perf_counter()
        rb.expected_squared_mean(x_a[:300])
        per_pt = (time.perf_counter() - t0) / 300
        line.update(reference_s_extrapolated=per_pt * x_a.size * 1025, reference_per_point_s=per_pt,
                    note="reference: 1025 scoring passes over x_a (bq.py:626-652), extrapolated from 300 points")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
