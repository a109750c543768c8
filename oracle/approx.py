"""TEST INFRASTRUCTURE -- numpy restatement of the reference's trapezoid (`use_approx`) path and of the periodic kernel.

Only tests/ may import this module (the product path never does; tests/test_capi_symbols.py enforces it).  Parity
pinned: tests/test_oracle_golden.py checks it against fixtures produced by the unmodified reference
(tests/golden/approx_gauss.npz, periodic_a.npz, periodic_b.npz; generator tests/golden/make_golden.py::main_approx).

What it restates, per function:
  * kernel()            gp.GaussianKernel / gp.PeriodicKernel of the un-vendored gaussian_processes==1.0.5
                        (h^2 N(x1 | x2, w^2) and h^2 exp(-2 sin^2((x1 - x2) / 2p) / w^2))
  * ApproxModel.__init__  bq.py:132-171, :967-991 (given candidates), :933-965
  * Z_mean / Z_var      bq.py:256-266 / :315-327 with bq_c.approx_Z_mean (bq_c.pyx:216-261) and approx_Z_var (:358-422)
  * esm_and_em          bq.py:447-527 with bq_c.improve_covariance_conditioning (bq_c.pyx:127-140),
                        approx_expected_squared_mean_and_mean (:538-598) and _esm_and_em (:425-490)
Pure numpy, one bordered Cholesky per query point as the reference does: small cases only.
"""
import numpy as np

EPS = np.finfo(np.float64).eps
MAX_EXPONENT = np.log(2.0 ** 1020)          # gauss_c.pyx:16


def kernel(kind, h, w, p, x1, x2):
    d = np.subtract.outer(np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64))
    if kind == 0:
        return h ** 2 / (np.sqrt(2 * np.pi) * w) * np.exp(-0.5 * d ** 2 / w ** 2)
    return h ** 2 * np.exp(-2.0 * np.sin(d / (2.0 * p)) ** 2 / w ** 2)


def trapz_rows(K, xo, p_xo):
    """int_K[i] = sum_j diff[j] (K[i, j] p[j] + K[i, j+1] p[j+1]) / 2   (bq_c.pyx:585-593; :247-259 for one row)"""
    diff = np.diff(xo)
    Kp = K * p_xo[None, :]
    return (diff[None, :] * (Kp[:, :-1] + Kp[:, 1:]) / 2.0).sum(axis=1)


def int_exp_norm(c, m, S):
    """gauss_c.pyx:65-92"""
    a = c * m + 0.5 * c ** 2 * S
    return np.inf if a > MAX_EXPONENT else np.exp(a)


class ApproxModel(object):
    def __init__(self, x_s, l_s, x_c, params_tl, params_l, x_mean, x_var, thresh, kind, xo, p_xo):
        """params_*: (h, w, s) for the Gaussian kernel, (h, w, p, s) for the periodic one."""
        self.kind = int(kind)
        self.x_s, self.l_s, self.x_c = (np.asarray(v, dtype=np.float64) for v in (x_s, l_s, x_c))
        self.ns, self.nc = self.x_s.size, self.x_c.size
        ptl, pl = list(params_tl), list(params_l)
        self.h_tl, self.w_tl, self.s_tl = ptl[0], ptl[1], ptl[-1]
        self.h_l, self.w_l, self.s_l = pl[0], pl[1], pl[-1]
        self.p_tl = ptl[2] if self.kind else 1.0
        self.p_l = pl[2] if self.kind else 1.0
        self.thresh = float(thresh)
        self.xo, self.p_xo = np.asarray(xo, dtype=np.float64), np.asarray(p_xo, dtype=np.float64)
        # gp_log_l
        self.K_tl = self.k_tl(self.x_s, self.x_s) + self.s_tl ** 2 * np.eye(self.ns)
        L = np.linalg.cholesky(self.K_tl)
        Li = np.linalg.inv(L)
        self.inv_K_tl = Li.T.dot(Li)
        self.a_tl = self.inv_K_tl.dot(np.log(self.l_s))
        # candidates' values and gp_l (bq.py:985, :144-165)
        self.l_c = np.exp(self.k_tl(self.x_c, self.x_s).dot(self.a_tl)) if self.nc else np.zeros(0)
        self.x_sc = np.concatenate([self.x_s, self.x_c])
        self.l_sc = np.concatenate([self.l_s, self.l_c])
        self.nsc = self.ns + self.nc
        K_l = self.k_l(self.x_sc, self.x_sc) + self.s_l ** 2 * np.eye(self.nsc)
        L = np.linalg.cholesky(K_l)
        Li = np.linalg.inv(L)
        self.alpha_l = Li.T.dot(Li).dot(self.l_sc)

    def k_tl(self, x1, x2):
        return kernel(self.kind, self.h_tl, self.w_tl, self.p_tl, x1, x2)

    def k_l(self, x1, x2):
        return kernel(self.kind, self.h_l, self.w_l, self.p_l, x1, x2)

    def l_mean(self, x):
        return self.k_l(x, self.x_sc).dot(self.alpha_l)

    def Z_mean(self):
        m = self.l_mean(self.xo)
        return float(trapz_rows(m[None, :], self.xo, self.p_xo)[0])

    def Z_var(self):
        m_l = self.l_mean(self.xo)
        Kxox = self.k_tl(self.xo, self.x_s)
        C_tl = self.k_tl(self.xo, self.xo) - Kxox.dot(self.inv_K_tl).dot(Kxox.T)
        buf = trapz_rows(C_tl * m_l[None, :], self.xo, self.p_xo)              # inner integral, bq_c.pyx:404-410
        return float(trapz_rows((buf * m_l)[None, :], self.xo, self.p_xo)[0])  # outer integral, :413-418

    def esm_and_em(self, x_a):
        x_a = np.asarray(x_a, dtype=np.float64)
        esm, em, st = np.empty(x_a.size), np.empty(x_a.size), np.zeros(x_a.size, dtype=np.int32)
        Zm = self.Z_mean()
        for t, xa in enumerate(x_a):
            if np.isclose(xa, self.x_s, atol=1e-4).any():                     # bq.py:456-459
                esm[t], em[t], st[t] = Zm ** 2, Zm, 1
                continue
            x_sca = np.concatenate([self.x_sc, [xa]])
            K = self.k_l(x_sca, x_sca)
            close = np.abs(self.x_c - xa) < self.thresh                        # bq.py:470
            if close.any():
                idx = np.nonzero(close)[0] + self.ns
                K[idx, idx] += max(EPS, K.max()) * 1e-4                        # bq_c.pyx:136-140
            K[self.nsc, self.nsc] += max(EPS, K.max()) * 1e-4
            try:
                L = np.linalg.cholesky(K)
            except np.linalg.LinAlgError:                                      # bq.py:481-490
                esm[t], em[t], st[t] = Zm ** 2, Zm, 2
                continue
            k_t = self.k_tl([xa], self.x_s)[0]
            tm_a = float(k_t.dot(self.a_tl))                                   # bq.py:493
            tC_a = float(self.k_tl([xa], [xa])[0, 0] - k_t.dot(self.inv_K_tl).dot(k_t))     # bq.py:496
            int_K = trapz_rows(self.k_l(x_sca, self.xo), self.xo, self.p_xo)   # bq_c.pyx:585-593
            A = np.linalg.solve(L.T, np.linalg.solve(L, int_K))                # bq_c.pyx:467
            A_a, A_sc_l = A[-1], float(A[:-1].dot(self.l_sc))
            e1 = int_exp_norm(1, tm_a, tC_a)
            if np.isinf(e1):
                esm[t] = em[t] = np.inf
                continue
            em[t] = A_sc_l + A_a * e1
            e2 = int_exp_norm(2, tm_a, tC_a)
            esm[t] = np.inf if np.isinf(e2) else A_sc_l ** 2 + 2 * A_sc_l * A_a * e1 + A_a ** 2 * e2
        return esm, em, st
