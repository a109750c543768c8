"""Host-side helpers of the BQ interface that are outside the CUDA hot path: candidate filtering,
the slice sampler, the optimiser wrapper and two plot helpers (reference:
bayesian_quadrature/util.py, util_c.pyx, and bq_c.filter_candidates)."""
import ctypes
import logging

import numpy as np

logger = logging.getLogger("bayesian_quadrature.util")
DTYPE = np.dtype("float64")
MIN = np.log(np.exp2(np.float64(np.finfo(np.float64).minexp + 4)))
RAND_MAX = 2147483647


def filter_candidates(x_c, x_s, thresh):
    """In-place candidate filter with the semantics of bq_c.filter_candidates (bq_c.pyx:601-650):
    repeatedly merge candidates closer than `thresh` into their midpoint (the second one becomes
    NaN), then drop candidates closer than `thresh` to an observation."""
    nc = x_c.shape[0]
    merged = True
    while merged:
        merged = False
        for i in range(nc):
            if np.isnan(x_c[i]):
                continue
            for j in range(i + 1, nc):
                if np.isnan(x_c[j]):
                    continue
                if abs(x_c[i] - x_c[j]) < thresh:
                    x_c[i] = (x_c[i] + x_c[j]) / 2.0
                    x_c[j] = np.nan
                    merged = True
    if x_s.size:
        for i in range(nc):
            if not np.isnan(x_c[i]) and (np.abs(x_c[i] - x_s) < thresh).any():
                x_c[i] = np.nan


class _LibcRand(object):
    """libc srand/rand, so that a run seeded through numpy consumes the same two random streams as
    the reference sampler (util_c.pyx:21-22, :33)."""

    def __init__(self):
        self._libc = ctypes.CDLL(None)
        self._libc.rand.restype = ctypes.c_int
        self._libc.srand.argtypes = [ctypes.c_uint]

    def seed(self, s):
        self._libc.srand(int(s))

    def uniform(self, lo, hi):
        return (self._libc.rand() / float(RAND_MAX)) * (hi - lo) + lo


def slice_sample(logpdf, niter, w, xval, nburn=1, freq=1):
    """Multivariate slice sampler along random directions with stepping-out and shrinkage
    (util.py:45-75 wrapper + util_c.pyx:25-148).  Returns samples[nburn:][::freq]."""
    xval = np.asarray(xval, dtype=DTYPE)
    d = xval.size
    samples = np.empty((niter, d))
    samples[0] = xval
    rng = _LibcRand()
    rng.seed(np.random.randint(0, RAND_MAX))
    i = 0
    while i < niter - 1:
        cur = samples[i]
        height = logpdf(cur)
        if height == -np.inf:
            raise RuntimeError("zero probability encountered")
        logy = np.log(rng.uniform(0, np.exp(height)))
        direction = np.random.rand(d) - 0.5
        direction /= np.linalg.norm(direction)
        left, right = -w, w
        j = 0
        while logpdf(cur + left * direction) >= logy:     # step out, at most 101 times per side
            left -= w
            j += 1
            if j > 100:
                break
        j = 0
        while logpdf(cur + right * direction) >= logy:
            right += w
            j += 1
            if j > 100:
                break
        while True:
            if (right - left) < 1e-9:             # window collapsed: retry this iteration
                break
            loc = rng.uniform(left, right)
            samples[i + 1] = cur + loc * direction
            if logpdf(samples[i + 1]) > logy:
                i += 1
                break
            if loc < 0:
                left = loc
            else:
                right = loc
    return samples[nburn:][::freq]


def find_good_parameters(logpdf, x0, method, ntry=10):
    """Maximise `logpdf` with scipy (util.py:151-169): up to `ntry` restarts, returns None when no
    restart reaches a log-density above MIN."""
    import scipy.optimize as optim
    for i in range(ntry):
        logger.debug("Attempt #%d with %s", i + 1, method)
        res = optim.minimize(fun=lambda x: -logpdf(x), x0=x0, method=method)
        p = logpdf(res["x"])
        if p > MIN:
            return res["x"]
        if logpdf(x0) < p:
            x0 = res["x"]
    return None


def set_scientific(ax, low, high, axis=None):
    import matplotlib.pyplot as plt
    fmt = plt.ScalarFormatter()
    fmt.set_scientific(True)
    fmt.set_powerlimits((low, high))
    if axis is None or axis == "x":
        ax.get_xaxis().set_major_formatter(fmt)
    if axis is None or axis == "y":
        ax.get_yaxis().set_major_formatter(fmt)


def vlines(ax, x, **kwargs):
    ymin, ymax = ax.get_ylim()
    ax.vlines(x, ymin, ymax, **kwargs)


def hlines(ax, y, **kwargs):
    xmin, xmax = ax.get_xlim()
    ax.hlines(y, xmin, xmax, **kwargs)
