"""TEST INFRASTRUCTURE — multi-precision "truth" for the expected-squared-mean path.

Evaluates, with mpmath at 80 significant digits, the exact mathematical value of what the
reference computes in float64 for one query point (bq.py:447-527 + bq_c.pyx:425-535 +
gauss_c.pyx:65-164 + the `gp` arithmetic of bq.py:465,493,496), as a function of the *double
precision inputs* (x_s, l_s, x_c, hyper-parameters, prior, x_a).  The branch decisions the
reference takes in float64 (np.isclose shortcut bq.py:456, the strict `< candidate_thresh`
pattern bq.py:470, the jitter magnitudes bq_c.pyx:136) are taken in float64 here as well, so the
truth is the exact value of the SAME formula, not of a different model.

Used only by tests/golden/make_golden.py to attach a truth column to the ill-conditioned
fixtures (where cond(K) >~ 1e9 the reference's own float64 result is off by more
than the 1e-9 parity tolerance, so both implementations are judged against this instead of
against each other).  Never imported by the product.
"""
import numpy as np
from mpmath import mp, mpf

mp.dps = 80
EPS = float(np.finfo(np.float64).eps)


def _chol(A):
    """Lower Cholesky of a list-of-lists mpf matrix."""
    n = len(A)
    L = [[mpf(0)] * n for _ in range(n)]
    for i in range(n):
        Li = L[i]
        for j in range(i + 1):
            Lj = L[j]
            s = A[i][j]
            for k in range(j):
                s -= Li[k] * Lj[k]
            if i == j:
                if s <= 0:
                    raise ArithmeticError("truth: matrix not positive definite at pivot %d" % i)
                Li[j] = mp.sqrt(s)
            else:
                Li[j] = s / Lj[j]
    return L


def _fwd(L, b):
    n = len(b)
    y = [mpf(0)] * n
    for i in range(n):
        s = b[i]
        Li = L[i]
        for k in range(i):
            s -= Li[k] * y[k]
        y[i] = s / Li[i]
    return y


def _bwd(L, y):
    n = len(y)
    x = [mpf(0)] * n
    for i in range(n - 1, -1, -1):
        s = y[i]
        for k in range(i + 1, n):
            s -= L[k][i] * x[k]
        x[i] = s / L[i][i]
    return x


def _dot(a, b):
    return mp.fsum(x * y for x, y in zip(a, b))


class Truth(object):
    def __init__(self, x_s, l_s, x_c, params_tl, params_l, x_mean, x_var, thresh):
        self.x_s64 = np.asarray(x_s, dtype=np.float64)
        self.x_c64 = np.asarray(x_c, dtype=np.float64)
        self.thresh = float(thresh)
        self.x_s = [mpf(float(v)) for v in x_s]
        self.x_c = [mpf(float(v)) for v in x_c]
        self.ns, self.nc = len(self.x_s), len(self.x_c)
        h_tl, w_tl, s_tl = (mpf(float(v)) for v in params_tl)
        h_l, w_l, s_l = (mpf(float(v)) for v in params_l)
        self.mu, self.sig2 = mpf(float(x_mean)), mpf(float(x_var))
        self.h_l, self.w_l, self.w_tl = h_l, w_l, w_tl
        self.c_tl = h_tl ** 2 / (mp.sqrt(2 * mp.pi) * w_tl)
        self.c_l = h_l ** 2 / (mp.sqrt(2 * mp.pi) * w_l)
        # GP over log l (bq.py:73, :147): K_tl + s^2 I, alpha_tl, l_c = exp(mean(x_c)) (bq.py:985)
        tl_s = [mp.log(mpf(float(v))) for v in l_s]
        Ktl = [[self.k_tl(a, b) + (s_tl ** 2 if i == j else 0) for j, b in enumerate(self.x_s)] for i, a in enumerate(self.x_s)]
        self.L_tl = _chol(Ktl)
        self.a_tl = _bwd(self.L_tl, _fwd(self.L_tl, tl_s))
        self.l_c = [mp.exp(_dot([self.k_tl(c, b) for b in self.x_s], self.a_tl)) for c in self.x_c]
        self.x_sc = self.x_s + self.x_c
        self.l_sc = [mpf(float(v)) for v in l_s] + self.l_c
        self.b_sc = [self.int_K(x) for x in self.x_sc]
        # float64 jitter magnitudes exactly as bq_c.pyx:136 evaluates them (np.max of the float64 Gram matrix: its diagonal)
        kmax = float(np.float64(float(params_l[0])) ** 2 / (np.sqrt(2 * np.pi) * np.float64(float(params_l[1]))))
        self.kmax64 = kmax
        self.j1 = max(EPS, kmax) * 1e-4
        self._patterns = {}
        self.s_l = s_l

    def k_tl(self, a, b):
        return self.c_tl * mp.exp(-(a - b) ** 2 / (2 * self.w_tl ** 2))

    def k_l(self, a, b):
        return self.c_l * mp.exp(-(a - b) ** 2 / (2 * self.w_l ** 2))

    def int_K(self, x):
        """gauss_c.pyx:95-164 for d = 1: h^2 N(x | mu, w^2 + sigma^2)."""
        v = self.w_l ** 2 + self.sig2
        return self.h_l ** 2 * mp.exp(-(x - self.mu) ** 2 / (2 * v)) / mp.sqrt(2 * mp.pi * v)

    def _pattern(self, P):
        if P not in self._patterns:
            n = self.ns + self.nc
            K = [[self.k_l(a, b) for b in self.x_sc] for a in self.x_sc]
            for j in P:
                K[self.ns + j][self.ns + j] += mpf(self.j1)
            L = _chol(K)
            alpha = _bwd(L, _fwd(L, self.l_sc))
            gamma = _bwd(L, _fwd(L, self.b_sc))
            self._patterns[P] = (L, alpha, gamma)
        return self._patterns[P]

    def Z_mean(self):
        """bq_c.pyx:157-213 with alpha_l = (K_l + s_l^2 I)^-1 l_sc (gp.inv_Kxx_y)."""
        if self.s_l == 0:
            _, alpha, _ = self._pattern(())
        else:
            K = [[self.k_l(a, b) + (self.s_l ** 2 if i == j else 0) for j, b in enumerate(self.x_sc)] for i, a in enumerate(self.x_sc)]
            L = _chol(K)
            alpha = _bwd(L, _fwd(L, self.l_sc))
        return _dot(self.b_sc, alpha)

    def esm_and_em(self, x_a):
        """(esm, em, shortcut) lists; infinities follow the float64 guards of gauss_c.pyx:87-91 (exponent > log 2^1020)."""
        MAXE = mpf(float(np.log(np.exp2(np.float64(1020)))))
        Zm = None
        out_esm, out_em, out_sc = [], [], []
        for xa64 in np.asarray(x_a, dtype=np.float64):
            if np.isclose(xa64, self.x_s64, atol=1e-4).any():                    # bq.py:456-459
                if Zm is None:
                    Zm = self.Z_mean()
                out_esm.append(Zm ** 2); out_em.append(Zm); out_sc.append(1)
                continue
            xa = mpf(float(xa64))
            P = tuple(int(j) for j in np.nonzero(np.abs(self.x_c64 - xa64) < self.thresh)[0])   # bq.py:470
            L, alpha, gamma = self._pattern(P)
            # jitter on the new point: np.max re-evaluated after the first pass (bq_c.pyx:136, bq.py:473-476)
            j2 = max(EPS, self.kmax64 + (self.j1 if P else 0.0)) * 1e-4
            k_a = [self.k_l(xa, b) for b in self.x_sc]
            v = _fwd(L, k_a)
            s = self.c_l + mpf(j2) - _dot(v, v)
            b_a = self.int_K(xa)
            A_a = (b_a - _dot(k_a, gamma)) / s                                    # last entry of K^-1 int_K, bq_c.pyx:467-469
            A_scl = _dot(self.b_sc, alpha) - A_a * _dot(k_a, alpha)               # dot(A[:-1], l_sc), bq_c.pyx:470
            k_t = [self.k_tl(xa, b) for b in self.x_s]
            tm = _dot(k_t, self.a_tl)                                             # gp_log_l.mean, bq.py:493
            vt = _fwd(self.L_tl, k_t)
            tC = self.c_tl - _dot(vt, vt)                                         # gp_log_l.cov, bq.py:496
            a1, a2 = tm + tC / 2, 2 * tm + 2 * tC
            if a1 > MAXE:                                                         # bq_c.pyx:472-475
                out_esm.append(mp.inf); out_em.append(mp.inf); out_sc.append(0)
                continue
            e1 = mp.exp(a1)
            em = A_scl + A_a * e1
            if a2 > MAXE:                                                         # bq_c.pyx:477-483
                out_esm.append(mp.inf); out_em.append(em); out_sc.append(0)
                continue
            e2 = mp.exp(a2)
            out_esm.append(A_scl ** 2 + 2 * A_scl * A_a * e1 + A_a ** 2 * e2)     # bq_c.pyx:485
            out_em.append(em); out_sc.append(0)
        f = lambda v: np.array([float(x) for x in v])
        return f(out_esm), f(out_em), np.array(out_sc, dtype=np.int32)
