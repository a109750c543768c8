"""Synthetic workloads of BASELINE.json's configs.

The generator scales the reference's own test fixture
(/root/reference/bayesian_quadrature/tests/util.py:12-59: spacing 1.25, hypers
(15, 2, 0) / (0.2, 1.3, 0), options n_candidate=10, candidate_thresh=0.5, seed 8728) to
``ns`` observations of a three-component Gaussian-mixture likelihood.
"""
import numpy as np

SEED = 8728                      # tests/util.py:22-23
PARAMS_TL = (15.0, 2.0, 0.0)     # tests/util.py:43
PARAMS_L = (0.2, 1.3, 0.0)

#: hyper-parameter ranges of config C4 (the well-conditioned part of what sample_hypers visits)
HYPER_RANGES = {"h_tl": (13.5, 15.5), "w_tl": (1.6, 2.2), "h_l": (0.15, 0.6), "w_l": (1.0, 1.5)}


def span(ns):
    return 5.0 * (ns - 1) / 8.0


def _npdf(x, m, s):
    return np.exp(-0.5 * ((x - m) / s) ** 2) / (np.sqrt(2 * np.pi) * s)


def likelihood(ns, shift=(0.0, 0.0, 0.0)):
    """The mixture likelihood l(x) for an ``ns``-observation problem; ``shift`` moves the three
    component centres (in units of span) — used to make C5's independent problems differ."""
    sp = span(ns)

    def l(x):
        x = np.asarray(x, dtype=np.float64)
        return (0.5 * _npdf(x, (-0.3 + shift[0]) * sp, 0.16 * sp)
                + 0.3 * _npdf(x, (0.4 + shift[1]) * sp, 0.10 * sp)
                + 0.2 * _npdf(x, (0.1 + shift[2]) * sp, 0.3 * sp))
    return l


def observations(ns, shift=(0.0, 0.0, 0.0)):
    x_s = 1.25 * (np.arange(ns, dtype=np.float64) - (ns - 1) / 2.0)
    return x_s, likelihood(ns, shift)(x_s)


def options(ns):
    """BQ options without the 'kernel' entry (the caller adds its GaussianKernel class)."""
    S = (ns - 1) / 8.0
    return {"n_candidate": 10, "candidate_thresh": 0.5, "x_mean": 0.0, "x_var": 10.0 * S * S,
            "optim_method": "L-BFGS-B"}


def query_grid(ns, na):
    sp = span(ns)
    return np.linspace(-2 * sp, 2 * sp, na)


def make_bq(BQ, GaussianKernel, ns, shift=(0.0, 0.0, 0.0), seed=SEED):
    """Build and initialise a BQ object (reference's or this package's) for the ns-observation
    workload: ``np.random.seed(seed)`` immediately before the constructor, as §8(d) states."""
    x_s, l_s = observations(ns, shift)
    opt = options(ns)
    opt["kernel"] = GaussianKernel
    np.random.seed(seed)
    bq = BQ(x_s, l_s, **opt)
    bq.init(params_tl=PARAMS_TL, params_l=PARAMS_L)
    return bq


def hyper_sets(n, seed=SEED):
    """C4: ``n`` hyper-parameter sets (h_tl, w_tl, h_l, w_l) drawn uniformly from HYPER_RANGES."""
    rs = np.random.RandomState(seed)
    cols = [rs.uniform(*HYPER_RANGES[k], size=n) for k in ("h_tl", "w_tl", "h_l", "w_l")]
    return np.stack(cols, axis=1)


def problem_shift(p, seed=SEED):
    """C5: per-problem jitter of the mixture centres."""
    rs = np.random.RandomState(seed + int(p))
    return tuple(rs.uniform(-0.05, 0.05, size=3))
