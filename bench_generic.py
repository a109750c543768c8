#!/usr/bin/env python
"""Timing of the generic device path (periodic kernel, trapezoid integrals) next to the reference's own
`use_approx` path on the host: `python bench_generic.py > profiles/generic_r02.jsonl` on a B200 box.  The reference leg
needs oracle/_ref (built where /root/reference exists; it travels with the snapshot)."""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.simplefilter("ignore")


def vmpdf(x, mu, kappa):
    from scipy.special import iv
    return np.exp(-np.log(2 * np.pi * iv(0, kappa)) + kappa * np.cos(x - mu))


def make(BQ, kernel, nobs, ptl, pl):
    np.random.seed(8728)
    x = np.linspace(-np.pi, np.pi, nobs + 1)[:-1]
    bq = BQ(x, vmpdf(x, 0.1, 1.1), n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5, kernel=kernel,
            optim_method="L-BFGS-B")
    bq.init(params_tl=ptl, params_l=pl)
    return bq


def main():
    import torch
    from bayesian_quadrature_b200 import BQ, PeriodicKernel
    from oracle import build_ref
    ref = build_ref.import_reference() if build_ref.built() else None
    for nobs, ptl, pl in ((5, (3, 1.2, 1, 0.), (0.3, 0.8, 1, 0.)), (24, (3, 0.42, 1, 0.), (0.3, 0.31, 1, 0.))):
        bq = make(BQ, PeriodicKernel, nobs, ptl, pl)
        out = {"case": "periodic kernel, %d observations, %d candidates, 1000-point trapezoid grid" % (bq.ns, bq.nc)}
        for na in (10 ** 4, 10 ** 5):
            x_a = np.linspace(-np.pi, np.pi, na)
            bq.expected_squared_mean(x_a)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            esm = bq.expected_squared_mean(x_a)
            out["cuda_s_na_%d" % na] = time.perf_counter() - t0
        out["cuda_evals_per_s"] = 10 ** 5 / out["cuda_s_na_100000"]
        if ref is not None:
            import logging
            logging.disable(logging.CRITICAL)
            bqmod, gp = ref
            rbq = make(bqmod.BQ, gp.PeriodicKernel, nobs, ptl, pl)
            sub = np.linspace(-np.pi, np.pi, 10 ** 5)[:: 10 ** 5 // 200][:200]
            t0 = time.perf_counter()
            r_esm = rbq.expected_squared_mean(sub)
            out["reference_s_200_points"] = time.perf_counter() - t0
            out["reference_evals_per_s"] = 200 / out["reference_s_200_points"]
            out["speedup"] = out["cuda_evals_per_s"] / out["reference_evals_per_s"]
            d_esm = bq.expected_squared_mean(sub)
            out["max_rel_diff_vs_reference"] = float(np.max(np.abs(d_esm - r_esm) / np.abs(r_esm)))
        print(json.dumps(out), flush=True)
    # more than 256 observations: Gaussian kernel, closed forms, capacity class 512 (generic kernel, dense algorithm)
    from bayesian_quadrature_b200 import GaussianKernel, synthetic
    for ns in (300, 512):
        bq = synthetic.make_bq(BQ, GaussianKernel, ns)
        x_a = synthetic.query_grid(ns, 10 ** 5)
        bq.expected_squared_mean(x_a[:1000])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bq.expected_squared_mean(x_a)
        dt = time.perf_counter() - t0
        print(json.dumps({"case": "Gaussian kernel, %d observations (capacity class 512, generic kernel), 10^5 points" % ns,
                          "cuda_s": dt, "cuda_evals_per_s": 10 ** 5 / dt}), flush=True)


if __name__ == "__main__":
    main()
