"""Host-side logic that needs no GPU: candidate filtering, option handling, constructor
validation, the GP host objects and the slice sampler (reference tests: test_bq_object.py:59-84,
test_util.py)."""
import numpy as np
import pytest

from bayesian_quadrature_b200 import util
from bayesian_quadrature_b200.bq import BQ
from bayesian_quadrature_b200.gp import GP, GaussianKernel, PeriodicKernel

OPTIONS = dict(n_candidate=10, x_mean=0.0, x_var=10.0, candidate_thresh=0.5, kernel=GaussianKernel,
               optim_method="L-BFGS-B")


def test_constructor_validation():
    # reference test_bad_init (test_bq_object.py:59-72)
    x = np.linspace(-5, 5, 9)
    y = np.exp(-0.5 * x ** 2)
    with pytest.raises(ValueError):
        BQ(x[:, None], y, **OPTIONS)
    with pytest.raises(ValueError):
        BQ(x, y[:, None], **OPTIONS)
    with pytest.raises(ValueError):
        BQ(x[:-1], y, **OPTIONS)
    with pytest.raises(ValueError):
        BQ(x, -y, **OPTIONS)
    with pytest.raises(TypeError):
        BQ(x, y, n_candidate=10)             # all six options are mandatory (bq.py:94)
    bq = BQ(x, y, **OPTIONS)
    assert not bq.initialized and bq.gp_l is None and bq.x_c is None
    assert (bq.tl_s == np.log(y)).all() and bq.ns == 9
    assert bq.options["use_approx"] is False and bq.options["wrapped"] is False
    assert bq.options["x_mean"].shape == (1,) and bq.options["x_cov"].shape == (1, 1)


def test_non_gaussian_kernel_options_and_prior():
    """gp.PeriodicKernel selects the trapezoid path on a wrapped domain (bq.py:124-125); the prior on the approximation
    grid is the reference's von Mises density, normalised with libc's j0 as bq_c.pyx:31-60 does (tests/golden/periodic_b.npz
    holds the grid and density the unmodified reference produced).  Without a GPU the device pass of init fails loudly."""
    from conftest import load_golden
    g = load_golden("periodic_b")
    opt = dict(OPTIONS, kernel=PeriodicKernel)
    x = np.linspace(-3, 3, 8)
    bq = BQ(x, np.ones_like(x), **opt)
    assert bq.options["use_approx"] and bq.options["wrapped"]
    np.testing.assert_allclose(bq._make_approx_px(g["xo"]), g["p_xo"], rtol=1e-13)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            bq.init((5, 6.0, 1, 0), (0.2, 1.5, 1, 0))


def test_filter_candidates_matches_reference_fixture_draw():
    # the reference fixture (tests/util.py:46-59, seed 8728) ends up with these two candidates
    np.random.seed(8728)
    x_s = np.linspace(-5, 5, 9)
    xc = np.random.uniform(x_s.min() - 2, x_s.max() + 2, 10)
    util.filter_candidates(xc, x_s, 0.5)
    x_c = np.sort(xc[~np.isnan(xc)])
    np.testing.assert_allclose(x_c, [-3.18507065, 4.31828185], atol=1e-8)
    # spacing invariants (reference test_choose_candidates, test_bq_object.py:75-84)
    assert (np.abs(x_c[:, None] - x_s[None]) >= 0.5).all()
    assert (np.diff(x_c) >= 0.5).all()


def test_filter_candidates_merges_close_pairs():
    xc = np.array([0.0, 0.2, 3.0, 10.0, 10.3, 10.5])
    util.filter_candidates(xc, np.array([3.1]), 0.5)
    kept = xc[~np.isnan(xc)]
    assert (np.diff(np.sort(kept)) >= 0.5).all()
    assert not np.isclose(kept, 3.0).any()


def test_gp_host_object():
    x = np.linspace(-5, 5, 9)
    y = np.sin(x)
    gp = GP(GaussianKernel(2.0, 1.5), x, y, s=0.1)
    K = gp.Kxx
    assert K is gp.Kxx                                            # memoised (reference test_bq_c.py:49)
    c = 4.0 / (np.sqrt(2 * np.pi) * 1.5)
    np.testing.assert_allclose(np.diag(K), c + 0.01)
    np.testing.assert_allclose(K[0, 1], c * np.exp(-0.5 * 1.25 ** 2 / 1.5 ** 2))
    np.testing.assert_allclose(np.dot(K, gp.inv_Kxx_y), y, atol=1e-9)
    np.testing.assert_allclose(gp.mean(x), y - 0.01 * gp.inv_Kxx_y, atol=1e-9)
    sign, logdet = np.linalg.slogdet(K)
    np.testing.assert_allclose(gp.log_lh, -0.5 * y @ np.linalg.solve(K, y) - 0.5 * logdet - 4.5 * np.log(2 * np.pi))
    gp.set_param("w", 2.0)
    assert gp._memoized == {} and gp.K.w == 2.0 and K is not gp.Kxx
    with pytest.raises(ValueError):
        gp.set_param("h", -1.0)
    np.testing.assert_allclose(gp.params, [2.0, 2.0, 0.1])
    g2 = gp.copy()
    g2.set_param("h", 3.0)
    assert gp.K.h == 2.0


def test_slice_sampler_recovers_a_gaussian():
    # reference test_util.py:20-34 (histogram test), here on the first two moments
    np.random.seed(8728)
    logpdf = lambda x: float(-0.5 * np.sum((x - 1.0) ** 2 / 0.25) - 0.5 * np.log(2 * np.pi * 0.25))
    s = util.slice_sample(logpdf, 4000, 1.0, np.array([0.0]), nburn=100, freq=1)
    assert s.shape == (3900, 1)
    assert abs(s.mean() - 1.0) < 0.05 and abs(s.std() - 0.5) < 0.05


def test_slice_sampler_zero_probability():
    with pytest.raises(RuntimeError):
        util.slice_sample(lambda x: -np.inf, 5, 1.0, np.array([0.0]))


def test_synthetic_workload_matches_survey_generator():
    from bayesian_quadrature_b200 import synthetic
    x_s, l_s = synthetic.observations(64)
    assert x_s.shape == (64,) and np.isclose(np.diff(x_s), 1.25).all() and abs(x_s.mean()) < 1e-12
    assert (l_s > 0).all()
    opt = synthetic.options(64)
    assert opt["x_var"] == 10.0 * (63 / 8.0) ** 2 and opt["n_candidate"] == 10
    g = synthetic.query_grid(64, 1001)
    assert g[0] == -2 * synthetic.span(64) and g[-1] == 2 * synthetic.span(64)
    h = synthetic.hyper_sets(1024)
    assert h.shape == (1024, 4) and (h[:, 1] >= 1.6).all() and (h[:, 1] <= 2.2).all()


def test_batched_candidate_filter_equals_the_scalar_one():
    from bayesian_quadrature_b200.batch import filter_candidates_batch
    rs = np.random.RandomState(0)
    P = 300
    ns = rs.randint(5, 10, P)
    x_s = np.zeros((P, 9))
    for p in range(P):
        x_s[p, :ns[p]] = np.sort(rs.uniform(-5, 5, ns[p]))
    xc = rs.uniform(-7, 7, (P, 10))
    ref = xc.copy()
    for p in range(P):
        util.filter_candidates(ref[p], x_s[p, :ns[p]], 0.5)
    filter_candidates_batch(xc, x_s, ns, 0.5)
    assert np.array_equal(np.isnan(xc), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(xc), np.nan_to_num(ref))


def test_looks_sorted_heuristic():
    """The sampled sortedness test that decides whether query points are sorted on the device first: grids pass,
    shuffled / reversed / NaN-carrying vectors do not (only speed depends on it, never results)."""
    from bayesian_quadrature_b200.bq import _looks_sorted
    rs = np.random.RandomState(0)
    grid = np.linspace(-50, 50, 100001)
    assert _looks_sorted(grid) and _looks_sorted(np.zeros(10000)) and _looks_sorted(np.array([1.0]))
    assert not _looks_sorted(grid[::-1]) and not _looks_sorted(rs.permutation(grid))
    bad = grid.copy()
    bad[0] = np.nan
    assert not _looks_sorted(bad)
    assert not _looks_sorted(np.concatenate([grid[50000:], grid[:50000]]))      # two sorted halves in the wrong order
