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


def with_truth(bq, x_a, rec, prior):
    """Adds the multi-precision truth columns (oracle/truth.py) and the reference's own error against them."""
    from oracle import truth
    T = truth.Truth(bq.x_s, bq.l_s, bq.x_c, bq.gp_log_l.params, bq.gp_l.params, prior[0], prior[1], prior[2])
    te, tm, sc = T.esm_and_em(x_a)
    assert (sc == rec["shortcut"]).all()
    rec.update(truth_esm=te, truth_em=tm, truth_Z_mean=float(T.Z_mean()),
               truth_l_c=np.array([float(v) for v in T.l_c]))
    fin = np.isfinite(te) & (te != 0)
    rel = np.abs(rec["esm"][fin] - te[fin]) / np.abs(te[fin])
    rec["ref_err_esm"] = float(rel.max())
    return rec


def fixture_bq(BQ, gp, params_tl, params_l):
    np.random.seed(8728)
    x = np.linspace(-5, 5, 9)
    y = scipy.stats.norm.pdf(x, 0, 1)
    bq = BQ(x, y, n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5,
            kernel=gp.GaussianKernel, optim_method="L-BFGS-B")
    bq.init(params_tl=params_tl, params_l=params_l)
    return bq


def synth_bq(BQ, gp, ns, params_tl, params_l, seed=8728):
    x_s, l_s = synthetic.observations(ns)
    opt = synthetic.options(ns)
    opt["kernel"] = gp.GaussianKernel
    np.random.seed(seed)
    bq = BQ(x_s, l_s, **opt)
    bq.init(params_tl=params_tl, params_l=params_l)
    return bq, opt


def main_edge():
    """Round-2 fixtures: the branches and regimes the round-1 fixtures never reach."""
    import logging
    logging.disable(logging.CRITICAL)               # the reference logs every infinite esm
    bqmod, gp = build_ref.import_reference()
    BQ = bqmod.BQ

    # ---- 4. int_exp_norm overflow guards (gauss_c.pyx:87-91 -> bq_c.pyx:472-483): prior variance of log l
    #      h_tl^2 / (sqrt(2 pi) w_tl) = 404 (h_tl = 45: only exp(2 tm + 2 tC) overflows, esm = +inf, em finite) and 1616
    #      (h_tl = 90: exp(tm + tC / 2) overflows too, esm = em = +inf), far from the data
    for name, h_tl in (("edge_inf45", 45.0), ("edge_inf90", 90.0)):
        bq = fixture_bq(BQ, gp, (h_tl, 2, 0.), (0.2, 1.3, 0.))
        x_a = np.concatenate([np.linspace(-25, 25, 101), edge_points(bq)])
        rec = with_truth(bq, x_a, record(bq, x_a), (0.0, 10.0, 0.5))
        assert (np.isinf(rec["esm"]) == np.isinf(rec["truth_esm"])).all() and (np.isinf(rec["em"]) == np.isinf(rec["truth_em"])).all()
        np.savez(os.path.join(OUT, name + ".npz"), **rec)
        print("%s: esm inf at %d, em inf at %d of %d points; reference-vs-truth max rel err %.3g" % (
            name, np.isinf(rec["esm"]).sum(), np.isinf(rec["em"]).sum(), x_a.size, rec["ref_err_esm"]))

    # ---- 5. observation noise on both GPs (the bordered matrix of bq.py:465 ignores s_l, alpha_l does not)
    bq = fixture_bq(BQ, gp, (15, 2, 0.3), (0.2, 1.3, 0.05))
    x_a = np.concatenate([np.linspace(-10, 10, 101), edge_points(bq)])
    rec = record(bq, x_a)
    np.savez(os.path.join(OUT, "edge_noise.npz"), **rec)
    print("edge_noise: Z_mean %.12g Z_var %.6g" % (rec["Z_mean"], rec["Z_var"]))

    # ---- 6. conditioning sweep at ns = 64 with multi-precision truth.  The reference's own float64 error
    #      grows like cond * eps; the tests hold both implementations against the truth, not against each other.
    rs = np.random.RandomState(5)
    for tag, w_tl, w_l in (("a", 2.0, 1.3), ("b", 2.7, 1.66), ("c", 3.5, 2.0), ("d", 2.0, 2.5)):
        bq, opt = synth_bq(BQ, gp, 64, (15.0, w_tl, 0.0), (0.2, w_l, 0.0))
        x_a = np.sort(np.concatenate([synthetic.query_grid(64, 41), rs.uniform(bq.x_s.min() - 3, bq.x_s.max() + 3, 40), bq.x_c,
                                      bq.x_c + 0.3, bq.x_s[::16] + 1e-5]))
        rec = with_truth(bq, x_a, record(bq, x_a), (0.0, opt["x_var"], 0.5))
        np.savez(os.path.join(OUT, "illcond_%s.npz" % tag), **rec)
        print("illcond_%s: w_tl %.2f w_l %.2f nc %d cond_tl %.3g cond_l %.3g reference-vs-truth max rel err %.3g" % (
            tag, w_tl, w_l, bq.nc, rec["cond_K_tl"], rec["cond_K_l"], rec["ref_err_esm"]))

    # ---- 7. the not-positive-definite fallback of bq.py:481-490.  With the Gaussian kernel it cannot return: the jitter on the
    #      new point (1e-4 of the diagonal, bq.py:473-476) keeps the LAST pivot positive unless the leading nsc x nsc block
    #      already fails, and then the fallback's own Z_mean() (bq.py:488 -> gp_l.inv_Kxx_y) raises LinAlgError.  The scan
    #      records what the reference does as K_l degenerates: 0 = scores returned (number of fallback points stored),
    #      1 = LinAlgError.
    w_ls = np.array([2.5, 3.0, 3.04, 3.4, 4.0, 5.0])
    outcome, nfall, conds = [], [], []
    for w_l in w_ls:
        try:
            bq, opt = synth_bq(BQ, gp, 64, (15.0, 2.0, 0.0), (0.2, float(w_l), 0.0))
            x_a = np.linspace(bq.x_s.min() - 4, bq.x_s.max() + 4, 200)
            r = bq.expected_squared_mean_and_mean(x_a)
            Zm = bq.Z_mean()
            fb = (r[:, 0] == Zm ** 2) & (status_of(bq, x_a) == 0)
            outcome.append(0); nfall.append(int(fb.sum())); conds.append(np.linalg.cond(bq.gp_l.Kxx))
        except np.linalg.LinAlgError:
            outcome.append(1); nfall.append(-1); conds.append(np.inf)
    np.savez(os.path.join(OUT, "notpd_scan.npz"), w_l=w_ls, outcome=np.array(outcome), n_fallback=np.array(nfall),
             cond_K_l=np.array(conds), ns=64, params_tl=np.array([15.0, 2.0, 0.0]), h_l=0.2)
    print("notpd_scan: w_l %s -> outcome %s fallbacks %s" % (w_ls, outcome, nfall))

    # ---- 8. sampler and choose_next under a fixed numpy seed (bq.py:565-598, :659-681; util_c.pyx:25-148).  Two runs from the
    #      same seed: one calls choose_next, the other sample_hypers at the same RNG position (marginalize's shape probe,
    #      bq.py:626-633, consumes no random numbers), so the fixture holds the sampled sets AND the point chosen from them.
    #      max_cond: the worst condition number among the proposals the chain evaluated.  The slice sampler's window
    #      (2 * nparam, bq.py:567) takes it into regions where K_tl is numerically singular and the log likelihood is
    #      rounding noise; there an accept / reject decision of the reference is not a property of the model and no
    #      other implementation can be asked to repeat it.
    def traced(bq):
        log = []
        orig = bq._make_llh_params

        def patched(params):
            f = orig(params)

            def g(x):
                v = f(x)
                if np.isfinite(v):
                    log.append(max(np.linalg.cond(bq.gp_log_l.Kxx), np.linalg.cond(bq.gp_l.Kxx)))
                return v
            return g
        bq._make_llh_params = patched
        return log

    for name, make, x_a, n in (
            ("choose_fixture", lambda: fixture_bq(BQ, gp, (15, 2, 0.), (0.2, 1.3, 0.)), np.linspace(-10, 10, 200), 20),
            # seed 8738: the first seed >= 8728 whose ns = 64 chain keeps every evaluated proposal below cond 1e9 (with 8728
            # itself the chain evaluates proposals of cond 2.9e15, 27 of the 32 seeds 8728..8759 go beyond 1e13)
            ("choose_c2", lambda: synth_bq(BQ, gp, 64, synthetic.PARAMS_TL, synthetic.PARAMS_L, seed=8738)[0],
             synthetic.query_grid(64, 1500), 6)):
        bq = make()
        chosen = bq.choose_next(x_a, n=n, params=["h", "w"])
        bq = make()
        log = traced(bq)
        h_tl, h_l = bq.sample_hypers(["h", "w"], n=n, nburn=1)
        # the marginal loss and tie set those samples give (bq.py:660-665), for diagnosis when the chosen point differs
        bq = make()
        esm = np.empty((n, x_a.size))
        for i in range(n):
            bq._set_gp_log_l_params(dict(zip(["h", "w"], h_tl[i])))
            bq._set_gp_l_params(dict(zip(["h", "w"], h_l[i])))
            esm[i] = bq.expected_squared_mean(x_a)
        loss = (-esm).mean(axis=0)
        close = np.nonzero(np.isclose(loss, loss.min()))[0]
        np.savez(os.path.join(OUT, name + ".npz"), x_s=bq.x_s, l_s=bq.l_s, x_c=bq.x_c, x_a=x_a, n=n, chosen=chosen,
                 hypers_tl=h_tl, hypers_l=h_l, loss=loss, tie_set=close, ns=bq.ns, max_cond=max(log), n_eval=len(log))
        print("%s: chosen %.6f, %d samples, tie set of %d points, argmin %d, max cond over %d evaluated proposals %.3g" % (
            name, chosen, n, close.size, int(np.argmin(loss)), len(log), max(log)))

    # ---- 9. sample_hypers from several seeds on the ns = 64 workload (a proposal whose K_l is not positive
    #      definite must not poison the following evaluations)
    seeds = np.array([1, 2, 3, 4, 5, 6, 8738, 8743])
    out_tl, out_l, conds = [], [], []
    for sd in seeds:
        bq, _ = synth_bq(BQ, gp, 64, synthetic.PARAMS_TL, synthetic.PARAMS_L, seed=int(sd))
        log = traced(bq)
        a, b = bq.sample_hypers(["h", "w"], n=4, nburn=2)
        out_tl.append(a); out_l.append(b); conds.append(max(log))
    np.savez(os.path.join(OUT, "sample_hypers_c2.npz"), seeds=seeds, hypers_tl=np.array(out_tl), hypers_l=np.array(out_l),
             max_cond=np.array(conds))
    print("sample_hypers_c2: %d seeds x 4 samples, max cond per seed %s" % (seeds.size, np.array(conds)))



def main_approx():
    """Round-2 fixtures of the trapezoid path (`use_approx`, bq.py:251-252, :310-311, :498-510; bq_c.pyx:216-261, :358-422,
    :538-598): the Gaussian fixture with use_approx switched on (the reference's own test does that,
    tests/test_bq_object.py:179), and gp.PeriodicKernel problems (tests/util.py:76-91: von Mises likelihood and prior,
    wrapped domain)."""
    import logging
    import scipy.special
    logging.disable(logging.CRITICAL)
    bqmod, gp = build_ref.import_reference()
    BQ = bqmod.BQ

    def rec_approx(bq, x_a, kind):
        r = record(bq, x_a)
        r.update(kind=kind, xo=np.array(bq._approx_x), p_xo=np.array(bq._approx_px), use_approx=1,
                 wrapped=int(bool(bq.options["wrapped"])))
        return r

    # ---- 9. Gaussian kernel, integrals by the trapezoid rule
    bq = fixture_bq(BQ, gp, (15, 2, 0.), (0.2, 1.3, 0.))
    bq.options["use_approx"] = True
    x_a = np.concatenate([np.linspace(-10, 10, 101), edge_points(bq)])
    rec = rec_approx(bq, x_a, 0)
    np.savez(os.path.join(OUT, "approx_gauss.npz"), **rec)
    print("approx_gauss: Z_mean %.12g Z_var %.6g (exact: %.12g %.6g), nc=%d" % (rec["Z_mean"], rec["Z_var"], bq._exact_Z_mean(),
                                                                              bq._exact_Z_var(), bq.nc))

    # ---- 10. periodic kernel: (a) the reference's own test problem (tests/util.py:76-91), (b) narrower kernels and fewer
    #      observations, so that candidates survive the filter and most points take the regular branch
    def vmpdf(x, mu, kappa):
        return np.exp(-np.log(2 * np.pi * scipy.special.iv(0, kappa)) + kappa * np.cos(x - mu))
    for name, nobs, ptl, pl in (("periodic_a", 8, (5, 2 * np.pi, 1, 0.), (0.2, np.pi / 2., 1, 0.)),
                                ("periodic_b", 5, (3, 1.2, 1, 0.), (0.3, 0.8, 1, 0.))):
        np.random.seed(8728)
        x = np.linspace(-np.pi, np.pi, nobs + 1)[:-1]
        y = vmpdf(x, 0.1, 1.1)
        bq = BQ(x, y, n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5, kernel=gp.PeriodicKernel,
                optim_method="L-BFGS-B")
        bq.init(params_tl=ptl, params_l=pl)
        x_a = np.concatenate([np.linspace(-np.pi, np.pi, 61), edge_points(bq)[:-5]])
        rec = rec_approx(bq, x_a, 1)
        np.savez(os.path.join(OUT, name + ".npz"), **rec)
        moved = (~np.isclose(rec["esm"], rec["Z_mean"] ** 2, rtol=1e-12)).sum()
        print("%s: ns=%d nc=%d Z_mean %.12g Z_var %.6g cond_l %.3g cond_tl %.3g; %d of %d points off the shortcut / fallback value" % (
            name, bq.ns, bq.nc, rec["Z_mean"], rec["Z_var"], rec["cond_K_l"], rec["cond_K_tl"], moved, x_a.size))


if __name__ == "__main__":
    if "--approx-only" in sys.argv:
        main_approx()
    else:
        if "--edge-only" not in sys.argv:
            main()
        main_edge()
        main_approx()
