"""Helper of tests/test_gpu_setup_generations.py: runs the setup kernel on a fixed list of ragged instances and saves
every model block (header, vectors, fragment-ordered operands) plus the header read-back to an .npz.  Run once with
BQB_SETUP_V1=1 (first-generation kernel, global scratch) and once without (shared-memory kernel)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_quadrature_b200 import _lib, synthetic  # noqa: E402


def cases():
    """(ns, nc, s_tl, s_l, check_max, shuffled) per instance group; groups share a capacity class"""
    out = []
    for cap, sizes in ((16, (1, 2, 7, 8, 9, 15, 16)), (64, (17, 31, 40, 63, 64)), (128, (65, 100, 127, 128)),
                       (160, (129, 148, 160)), (256, (161, 200, 255, 256))):
        for k, ns in enumerate(sizes):
            out.append((cap, ns, (k * 3) % 7, 0.0, 0.0, bool(k & 1), bool(k & 2)))
        out.append((cap, sizes[-2], 3, 0.0, 0.05, True, False))       # noisy gp_l (SURVEY A.2 asymmetry)
        out.append((cap, sizes[-1], 16 if cap >= 64 else 5, 1e-3, 0.0, False, True))   # many candidates, noisy gp_log_l
    return out


def build(case, rs):
    cap, ns, nc, s_tl, s_l, check_max, shuffled = case
    x_s = 1.25 * (np.arange(ns, dtype=np.float64) - (ns - 1) / 2.0)
    l_s = synthetic.likelihood(max(ns, 9))(x_s)
    if shuffled:
        p = rs.permutation(ns)
        x_s, l_s = x_s[p], l_s[p]
    # candidates as bq.py:967-991 leaves them: at least candidate_thresh = 0.5 from every observation and from each other
    # (mid-points of random gaps of the 1.25-spaced observations, and points beyond both ends)
    xs = np.sort(x_s)
    spots = np.concatenate([xs[:-1] + 0.625, xs[:1] - 0.9 - 1.1 * np.arange(8), xs[-1:] + 0.9 + 1.1 * np.arange(8)])
    x_c = np.sort(rs.choice(spots, size=nc, replace=False)) if nc else np.zeros(0)
    opt = synthetic.options(max(ns, 9))
    hyp = np.array([15.0, 2.0, s_tl, 0.2, 1.3, s_l])
    prior = np.array([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]])
    return x_s, l_s, x_c, hyp, prior


def main(path):
    rs = np.random.RandomState(1234)
    res = {}
    by_cap = {}
    for c in cases():
        by_cap.setdefault((c[0], c[5]), []).append(c)
    for (cap, check_max), group in sorted(by_cap.items()):
        B = len(group)
        b = _lib.Batch(B, cap)
        ns = np.array([g[1] for g in group], dtype=np.int32)
        nc = np.array([g[2] for g in group], dtype=np.int32)
        X, L, XC = np.zeros((B, cap)), np.ones((B, cap)), np.zeros((B, 16))
        H, P = np.zeros((B, 6)), np.zeros((B, 3))
        for i, g in enumerate(group):
            x_s, l_s, x_c, hyp, prior = build(g, rs)
            X[i, :g[1]], L[i, :g[1]], XC[i, :g[2]], H[i], P[i] = x_s, l_s, x_c, hyp, prior
        info = b.setup(ns, nc, X, L, XC, H, P, check_max=check_max)
        key = "cap%d_cm%d" % (cap, int(check_max))
        res[key + "_models"] = np.stack([b.read_model(i) for i in range(B)])
        res[key + "_ns"], res[key + "_nc"] = ns, nc
        for k, v in info.items():
            res[key + "_" + k] = v
        b.close()
    np.savez(path, **res)


if __name__ == "__main__":
    main(sys.argv[1])
