import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

#: north_star tolerance: results must match the reference within rtol 1e-9 / atol 1e-12 (float64)
RTOL = 1e-9
ATOL = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def assert_close(got, want, what="", rtol=RTOL, atol=ATOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    err = np.abs(got - want)
    bad = ~both_inf & ~(err <= atol + rtol * np.abs(want))
    if bad.any():
        i = int(np.argmax(np.where(bad, err / (atol + rtol * np.abs(want)), 0)))
        raise AssertionError("%s: %d/%d outside rtol=%g atol=%g; worst at %d: got %.17g want %.17g" % (
            what, int(bad.sum()), bad.size, rtol, atol, i, got.flat[i], want.flat[i]))


def truth_err(got, truth):
    fin = np.isfinite(truth) & (truth != 0)
    assert (np.isinf(got) == np.isinf(truth)).all()
    return np.abs(got[fin] - truth[fin]) / np.abs(truth[fin])


def truth_bound(g):
    """Relative error allowed against the multi-precision truth: the parity tolerance, twice what the reference itself
    shows on the fixture, or a tenth of the a-priori forward-error scale cond * eps of a backward-stable solve -- the
    errors of two stable float64 algorithms on one ill-conditioned problem are unrelated samples from that scale (the C
    restatement is 12x the reference on fixture d and 2x better on b), so the reference's sample alone is no bound."""
    cond = max(float(g["cond_K_tl"]), float(g["cond_K_l"]))
    return max(RTOL, 2 * float(g["ref_err_esm"]), 0.1 * cond * np.finfo(np.float64).eps)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
