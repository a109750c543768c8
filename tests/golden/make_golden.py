#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref: the
reference's own Cython + bq.py, with the numpy stand-in for the un-vendored `gp`).

Run here (where /root/reference exists):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files only.

Every fixture stores the inputs a C-ABI call needs (x_s, l_s, x_c, hypers, prior, thresh,
x_a) and the reference's outputs (esm, em, expected_Z_var, Z_mean, Z_var, l_c, status).
"""
import os
import sys
import warnings

import numpy as np
import scipy.stats

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref                        # noqa: E402
from bayesian_quadrature_b200 import synthetic      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")

#: printed in /root/reference/docs/ipynb/visual-tests.ipynb (cell outputs at the cited lines)
NOTEBOOK_GOLDENS = {
    "sum_int_K": 0.033989188741,            # :346
    "sum_int_K1_K2": 5.55090592396,         # :407
    "sum_int_int_K1_K2_K1": 0.0230979040904,  # :468
    "sum_int_int_K1_K2": 0.576959869272,    # :529
    "int_int_K": 0.00342641751296,          # :586
    "Z_mean": 0.119771005796,               # :640
    "Z_var": 5.98039315292e-07,             # :694
}


def status_of(bq, x_a):
    """Classify points the way bq.py:456-459 / :481-490 branch (1 shortcut)."""
    st = np.zeros(x_a.size, dtype=np.int32)
    for i, x in enumerate(x_a):
        if np.isclose(x, bq.x_s, atol=1e-4).any():
            st[i] = 1
    return st


def record(bq, x_a, extra=None):
    esm_em = bq.expected_squared_mean_and_mean(x_a)
    Zm, Zv = bq.Z_mean(), bq.Z_var()
    d = dict(
        x_s=bq.x_s, l_s=bq.l_s, x_c=bq.x_c, l_c=bq.l_c,
        params_tl=np.array(bq.gp_log_l.params), params_l=np.array(bq.gp_l.params),
        x_mean=float(bq.options["x_mean"][0]), x_var=float(bq.options["x_cov"][0, 0]),
        candidate_thresh=bq.options["candidate_thresh"],
        x_a=x_a, esm=esm_em[:, 0], em=esm_em[:, 1], Z_mean=Zm, Z_var=Zv,
        expected_Z_var=Zm ** 2 + Zv - esm_em[:, 0], shortcut=status_of(bq, x_a),
        alpha_l=bq.gp_l.inv_Kxx_y, cond_K_l=np.linalg.cond(bq.gp_l.Kxx),
        cond_K_tl=np.linalg.cond(bq.gp_log_l.Kxx),
        log_lh=bq.gp_log_l.log_lh + bq.gp_l.log_lh)
    if extra:
        d.update(extra)
    return d


def edge_points(bq):
    """Branch coverage: on/near observations (isclose boundary), on candidates, at the strict
    `< candidate_thresh` boundary, between points, far field."""
    t = bq.options["candidate_thresh"]
    pts = []
    for xs in bq.x_s[:: max(1, bq.ns // 6)]:
        b = 1e-4 + 1e-5 * abs(xs)
        pts += [xs, xs + 0.5 * b, xs - 0.999 * b, xs + 1.001 * b, xs - 1.5 * b, xs + 10 * b]
    for xc in bq.x_c:
        pts += [xc, xc + 1e-9, xc - 1e-3, xc + 0.5 * t, xc - t, xc + t, np.nextafter(xc + t, -np.inf),
                np.nextafter(xc - t, np.inf), xc + 1.01 * t, xc - 0.99 * t]
    lo, hi = bq.x_sc.min(), bq.x_sc.max()
    w = hi - lo
    pts += [lo - 0.3 * w, hi + 0.3 * w, lo - 2 * w, hi + 3 * w, lo - 10 * w, 1e6, -1e6]
    return np.array(pts, dtype=np.float64)


def main():
    ok = build_ref.build()
    if not ok:
        raise SystemExit("oracle/_ref not available: run where /root/reference exists")
    bqmod, gp = build_ref.import_reference()
    BQ = bqmod.BQ
    from bayesian_quadrature import gauss_c

    # ---- 1. the reference's own test fixture (tests/util.py:46-59) + the notebook goldens
    np.random.seed(8728)
    x = np.linspace(-5, 5, 9)
    y = scipy.stats.norm.pdf(x, 0, 1)
    bq = BQ(x, y, n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5,
            kernel=gp.GaussianKernel, optim_method="L-BFGS-B")
    bq.init(params_tl=(15, 2, 0.), params_l=(0.2, 1.3, 0.))
    x_l = np.array(bq.gp_l.x[None], order="F"); h_l = bq.gp_l.K.h; w_l = np.array([bq.gp_l.K.w])
    x_tl = np.array(bq.gp_log_l.x[None], order="F"); h_tl = bq.gp_log_l.K.h; w_tl = np.array([bq.gp_log_l.K.w])
    xm, xC = bq.options["x_mean"], bq.options["x_cov"]
    got = {}
    c = np.empty(bq.nsc); gauss_c.int_K(c, x_l, h_l, w_l, xm, xC); got["sum_int_K"] = c.sum()
    c = np.empty((bq.nsc, bq.ns), order="F"); gauss_c.int_K1_K2(c, x_l, x_tl, h_l, w_l, h_tl, w_tl, xm, xC)
    got["sum_int_K1_K2"] = c.sum()
    c = np.empty((bq.nsc, bq.nsc), order="F"); gauss_c.int_int_K1_K2_K1(c, x_l, h_l, w_l, h_tl, w_tl, xm, xC)
    got["sum_int_int_K1_K2_K1"] = c.sum()
    c = np.empty(bq.ns); gauss_c.int_int_K1_K2(c, x_tl, h_l, w_l, h_tl, w_tl, xm, xC)
    got["sum_int_int_K1_K2"] = c.sum()
    got["int_int_K"] = gauss_c.int_int_K(1, h_l, w_l, xm, xC)
    got["Z_mean"] = bq._exact_Z_mean()
    got["Z_var"] = bq._exact_Z_var()
    for k, v in NOTEBOOK_GOLDENS.items():
        tol = 5e-9 if k == "Z_var" else 5e-12   # Z_var is a 6-digit cancellation
        assert abs(got[k] - v) <= tol * abs(v), (k, got[k], v)
    x_a = np.concatenate([np.linspace(-10, 10, 201), edge_points(bq)])
    np.savez(os.path.join(OUT, "fixture.npz"), **record(bq, x_a, extra={
        "notebook_" + k: v for k, v in NOTEBOOK_GOLDENS.items()}))
    print("fixture: nc=%d na=%d cond_l=%.3g cond_tl=%.3g" % (bq.nc, x_a.size, np.linalg.cond(bq.gp_l.Kxx),
                                                            np.linalg.cond(bq.gp_log_l.Kxx)))

    # ---- 2. BASELINE configs on bounded samples of their query grids
    rs = np.random.RandomState(1)
    for name, ns, na_full, nsub, nwin in [("c1", 8, 200, 200, 0), ("c2", 64, 10 ** 6, 400, 201),
                                          ("c5", 128, 4096, 256, 0), ("c3", 256, 10 ** 7, 120, 81)]:
        bq = synthetic.make_bq(BQ, gp.GaussianKernel, ns)
        grid = synthetic.query_grid(ns, na_full)
        if nsub >= na_full:
            idx = np.arange(na_full)
        else:
            idx = np.unique(np.concatenate([np.linspace(0, na_full - 1, nsub).astype(np.int64),
                                            rs.randint(0, na_full, size=nsub // 4)]))
        if nwin:
            # dense window of consecutive grid points around the coarse argmax of esm, to pin
            # the argmin identity of choose_next where the gap is ~1e-9 relative
            coarse = bq.expected_squared_mean(grid[idx])
            centre = idx[int(np.argmax(coarse))]
            lo = max(0, centre - nwin // 2)
            idx = np.unique(np.concatenate([idx, np.arange(lo, min(na_full, lo + nwin))]))
        x_a = np.concatenate([grid[idx], edge_points(bq)])
        np.savez(os.path.join(OUT, name + ".npz"), **record(bq, x_a, extra={"grid_idx": idx, "na_full": na_full}))
        print("%s: ns=%d nc=%d na=%d cond_l=%.3g cond_tl=%.3g" % (name, ns, bq.nc, x_a.size,
              np.linalg.cond(bq.gp_l.Kxx), np.linalg.cond(bq.gp_log_l.Kxx)))

    # ---- 3. C4: hyper-parameter sets pushed through _set_gp_log_l_params / _set_gp_l_params
    bq = synthetic.make_bq(BQ, gp.GaussianKernel, 64)
    hyp = synthetic.hyper_sets(8)
    grid = synthetic.query_grid(64, 301)
    esm = np.empty((hyp.shape[0], grid.size)); em = np.empty_like(esm)
    Zm = np.empty(hyp.shape[0]); Zv = np.empty(hyp.shape[0]); l_c = np.empty((hyp.shape[0], bq.nc))
    for i, (h_tl, w_tl, h_l, w_l) in enumerate(hyp):
        bq._set_gp_log_l_params({"h": h_tl, "w": w_tl})
        bq._set_gp_l_params({"h": h_l, "w": w_l})
        r = bq.expected_squared_mean_and_mean(grid)
        esm[i], em[i] = r[:, 0], r[:, 1]
        Zm[i], Zv[i], l_c[i] = bq.Z_mean(), bq.Z_var(), bq.l_c
    loss = (-esm).mean(axis=0)                      # bq.py:660-662
    np.savez(os.path.join(OUT, "c4.npz"), x_s=bq.x_s, l_s=bq.l_s, x_c=bq.x_c, hypers=hyp, x_a=grid,
             x_mean=float(bq.options["x_mean"][0]), x_var=float(bq.options["x_cov"][0, 0]),
             candidate_thresh=bq.options["candidate_thresh"], esm=esm, em=em, Z_mean=Zm, Z_var=Zv, l_c=l_c,
             loss=loss, argmin=int(np.argmin(loss)))
    print("c4: %d hyper sets x %d points, argmin=%d" % (hyp.shape[0], grid.size, int(np.argmin(loss))))


if __name__ == "__main__":
    main()
