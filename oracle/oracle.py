"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/bq_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs import this module.  The product package never does (it fails loudly without its CUDA
library instead of falling back to anything here).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "bq_oracle.c")
LIB = os.path.join(HERE, "libbq_oracle.so")

ST_OK, ST_SHORTCUT, ST_NOTPD, ST_ESM_INF, ST_EM_INF, ST_ESM_BAD, ST_EM_BAD, ST_XA_BAD = 0, 1, 2, 4, 8, 16, 32, 64

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -ffp-contract=off: no FMA contraction, so the arithmetic is the reference's plain
        # IEEE double sequence
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.orc_model_new.restype = ctypes.c_void_p
        L.orc_model_new.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _dp,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.orc_model_free.argtypes = [ctypes.c_void_p]
        L.orc_model_error.argtypes = [ctypes.c_void_p]
        for f in ("orc_model_Z_mean", "orc_model_Z_var", "orc_model_log_lh"):
            getattr(L, f).restype = ctypes.c_double
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.orc_model_l_c.argtypes = [ctypes.c_void_p, _dp]
        L.orc_model_alpha_l.argtypes = [ctypes.c_void_p, _dp]
        L.orc_esm_and_em.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp, _ip]
        L.orc_expected_Z_var.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.orc_model_gp_log_l_mean_cov.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp]
        L.orc_int_K.argtypes = [ctypes.c_int, _dp] + [ctypes.c_double] * 4 + [_dp]
        L.orc_int_K1_K2.argtypes = [ctypes.c_int, _dp, ctypes.c_int, _dp] + [ctypes.c_double] * 6 + [_dp]
        L.orc_int_int_K1_K2_K1.argtypes = [ctypes.c_int, _dp] + [ctypes.c_double] * 6 + [_dp, _dp]
        L.orc_int_int_K1_K2.argtypes = [ctypes.c_int, _dp] + [ctypes.c_double] * 6 + [_dp]
        L.orc_int_int_K.restype = ctypes.c_double
        L.orc_int_int_K.argtypes = [ctypes.c_double] * 4
        L.orc_int_exp_norm.restype = ctypes.c_double
        L.orc_int_exp_norm.argtypes = [ctypes.c_double] * 3
        L.orc_cho_factor.argtypes = [ctypes.c_int, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class OracleError(Exception):
    pass


class OracleModel(object):
    """One BQ problem under one hyper-parameter set, as the reference would hold it after
    ``BQ(x, l, **options).init(params_tl, params_l)`` with the given candidate locations
    ``x_c`` (bq.py:132-171, :967-991), or after ``_set_gp_log_l_params`` /
    ``_set_gp_l_params`` (bq.py:933-965) when ``check_max`` is set."""

    def __init__(self, x_s, l_s, x_c, params_tl, params_l, x_mean, x_var, candidate_thresh, check_max=False):
        L = lib()
        self.x_s, self.l_s, self.x_c = _d(x_s), _d(l_s), _d(x_c)
        ptl, pl = _d(params_tl), _d(params_l)
        assert ptl.size == 3 and pl.size == 3
        xc = self.x_c if self.x_c.size else np.zeros(1)
        self._h = L.orc_model_new(self.x_s.size, _p(self.x_s), _p(self.l_s), self.x_c.size, _p(xc),
                                  _p(ptl), _p(pl), float(x_mean), float(x_var), float(candidate_thresh),
                                  int(check_max))
        err = L.orc_model_error(self._h)
        if err:
            self.close()
            raise np.linalg.LinAlgError({1: "gp_log_l Kxx is not positive definite",
                                         2: "gp_l Kxx is not positive definite",
                                         3: "GP mean is too large"}[err])
        self.ns, self.nc = self.x_s.size, self.x_c.size

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_model_free(self._h)
            self._h = None

    __del__ = close

    def Z_mean(self):
        return lib().orc_model_Z_mean(self._h)

    def Z_var(self):
        return lib().orc_model_Z_var(self._h)

    def log_lh(self):
        return lib().orc_model_log_lh(self._h)

    @property
    def l_c(self):
        out = np.empty(max(self.nc, 1))
        lib().orc_model_l_c(self._h, _p(out))
        return out[:self.nc]

    @property
    def alpha_l(self):
        out = np.empty(self.ns + self.nc)
        lib().orc_model_alpha_l(self._h, _p(out))
        return out

    def esm_and_em(self, x_a):
        x_a = _d(x_a)
        esm, em = np.empty(x_a.size), np.empty(x_a.size)
        st = np.empty(x_a.size, dtype=np.int32)
        lib().orc_esm_and_em(self._h, x_a.size, _p(x_a), _p(esm), _p(em), st.ctypes.data_as(_ip))
        return esm, em, st

    def expected_Z_var(self, x_a):
        esm, _, _ = self.esm_and_em(x_a)
        return self.Z_mean() ** 2 + self.Z_var() - esm

    def gp_log_l_mean_cov(self, x):
        x = _d(x)
        m, c = np.empty(x.size), np.empty(x.size)
        lib().orc_model_gp_log_l_mean_cov(self._h, x.size, _p(x), _p(m), _p(c))
        return m, c


def int_K(x, h, w, mu, cov):
    x = _d(x); out = np.empty(x.size)
    lib().orc_int_K(x.size, _p(x), h, w, mu, cov, _p(out))
    return out


def int_K1_K2(x1, x2, h1, w1, h2, w2, mu, cov):
    x1, x2 = _d(x1), _d(x2); out = np.empty((x1.size, x2.size), order="F")
    lib().orc_int_K1_K2(x1.size, _p(x1), x2.size, _p(x2), h1, w1, h2, w2, mu, cov, _p(out))
    return out


def int_int_K1_K2_K1(x, h1, w1, h2, w2, mu, cov):
    x = _d(x); out = np.empty((x.size, x.size), order="F"); work = np.empty(2 * x.size)
    lib().orc_int_int_K1_K2_K1(x.size, _p(x), h1, w1, h2, w2, mu, cov, _p(out), _p(work))
    return out


def int_int_K1_K2(x, h1, w1, h2, w2, mu, cov):
    x = _d(x); out = np.empty(x.size)
    lib().orc_int_int_K1_K2(x.size, _p(x), h1, w1, h2, w2, mu, cov, _p(out))
    return out


def int_int_K(h, w, mu, cov):
    return lib().orc_int_int_K(h, w, mu, cov)
