"""TEST INFRASTRUCTURE — numpy stand-in for the un-vendored third-party package
``gaussian_processes==1.0.5`` (import name ``gp``; /root/reference/requirements.txt:2,
imported at bayesian_quadrature/bq.py:7).

The package source is absent from /root/reference and there is no network, so its
published behaviour is restated here.  What pins it: the reference's closed forms only
equal the integral of the kernel if K(x, x') = h^2 N(x | x', w^2)
(bayesian_quadrature/gauss_c.pyx:110 and tests/test_gauss_c.py:84-104), and the seven
12-digit goldens printed in docs/ipynb/visual-tests.ipynb (checked in
tests/test_oracle_golden.py).  The floating-point *route* (explicit inverse through
inv(Lxx)) follows gp 1.0.5's memoised ``inv_Lxx``/``inv_Kxx`` properties.

build_ref.py copies this file to oracle/_ref/gp.py so that the unmodified reference
``bq.py`` can ``from gp import GP, GaussianKernel, PeriodicKernel``.
Nothing in the product package imports it.
"""
import numpy as np

DTYPE = np.float64


class _Kernel(object):
    _names = ()

    @property
    def params(self):
        return np.array([getattr(self, n) for n in self._names], dtype=DTYPE)

    @params.setter
    def params(self, val):
        for n, v in zip(self._names, val):
            setattr(self, n, v)

    def _check(self, name, v):
        v = float(v)
        if not v > 0:
            raise ValueError("invalid value for %s: %s" % (name, v))
        return v


class GaussianKernel(_Kernel):
    """K(x1, x2) = h^2 / (sqrt(2 pi) w) * exp(-(x1 - x2)^2 / (2 w^2))."""
    _names = ("h", "w")

    def __init__(self, h, w):
        self.h = h
        self.w = w

    h = property(lambda s: s._h, lambda s, v: setattr(s, "_h", s._check("h", v)))
    w = property(lambda s: s._w, lambda s, v: setattr(s, "_w", s._check("w", v)))

    def copy(self):
        return GaussianKernel(self.h, self.w)

    def __call__(self, x1, x2):
        x1 = np.asarray(x1, dtype=DTYPE)
        x2 = np.asarray(x2, dtype=DTYPE)
        d = x1[:, None] - x2[None, :]
        c = self.h ** 2 / (np.sqrt(2 * np.pi) * self.w)
        return c * np.exp(-0.5 * d ** 2 / self.w ** 2)

    K = __call__


class PeriodicKernel(_Kernel):
    """h^2 exp(-2 sin^2((x1 - x2) / (2 p)) / w^2) — out of the hot path; present so the
    reference module imports."""
    _names = ("h", "w", "p")

    def __init__(self, h, w, p):
        self.h = h
        self.w = w
        self.p = p

    h = property(lambda s: s._h, lambda s, v: setattr(s, "_h", s._check("h", v)))
    w = property(lambda s: s._w, lambda s, v: setattr(s, "_w", s._check("w", v)))
    p = property(lambda s: s._p, lambda s, v: setattr(s, "_p", s._check("p", v)))

    def copy(self):
        return PeriodicKernel(self.h, self.w, self.p)

    def __call__(self, x1, x2):
        x1 = np.asarray(x1, dtype=DTYPE)
        x2 = np.asarray(x2, dtype=DTYPE)
        d = x1[:, None] - x2[None, :]
        return self.h ** 2 * np.exp(-2.0 * np.sin(d / (2.0 * self.p)) ** 2 / self.w ** 2)

    K = __call__


def _memo(f):
    name = f.__name__

    def g(self):
        try:
            return self._memoized[name]
        except KeyError:
            v = self._memoized[name] = f(self)
            return v
    g.__name__ = name
    return property(g)


class GP(object):
    def __init__(self, K, x, y, s=0):
        self._memoized = {}
        self.K = K
        self._x = np.array(x, dtype=DTYPE, copy=True)
        self._y = np.array(y, dtype=DTYPE, copy=True)
        self._s = None
        self.s = s

    # -- data / parameters: every setter invalidates the caches ------------------
    @property
    def x(self):
        return self._x

    @x.setter
    def x(self, val):
        self._memoized = {}
        self._x = np.array(val, dtype=DTYPE, copy=True)

    @property
    def y(self):
        return self._y

    @y.setter
    def y(self, val):
        self._memoized = {}
        self._y = np.array(val, dtype=DTYPE, copy=True)

    @property
    def s(self):
        return self._s

    @s.setter
    def s(self, val):
        val = float(val)
        if not val >= 0:
            raise ValueError("invalid value for s: %s" % val)
        self._memoized = {}
        self._s = val

    @property
    def params(self):
        return np.array(list(self.K.params) + [self._s], dtype=DTYPE)

    @params.setter
    def params(self, val):
        self._memoized = {}
        self.K.params = val[:-1]
        self.s = val[-1]

    def get_param(self, name):
        if name == "s":
            return self._s
        return getattr(self.K, name)

    def set_param(self, name, val):
        if name == "s":
            self.s = val
        else:
            setattr(self.K, name, val)   # raises ValueError when invalid
            self._memoized = {}

    def copy(self, deep=True):
        new = GP(self.K.copy(), self._x, self._y, s=self._s)
        if hasattr(self, "jitter"):
            new.jitter = np.array(self.jitter, copy=True)
        return new

    # -- memoised linear algebra ---------------------------------------------------
    @_memo
    def Kxx(self):
        K = self.K(self._x, self._x)
        K += self._s ** 2 * np.eye(self._x.size)
        if np.isnan(K).any():
            raise ArithmeticError("Kxx contains invalid values")
        return K

    @_memo
    def Lxx(self):
        return np.linalg.cholesky(self.Kxx)

    @_memo
    def inv_Lxx(self):
        return np.linalg.inv(self.Lxx)

    @_memo
    def inv_Kxx(self):
        Li = self.inv_Lxx
        return np.dot(Li.T, Li)

    @_memo
    def inv_Kxx_y(self):
        return np.dot(self.inv_Kxx, self._y)

    @_memo
    def log_lh(self):
        y = self._y
        n = y.size
        try:
            L = self.Lxx
        except np.linalg.LinAlgError:
            return -np.inf
        data_fit = -0.5 * np.dot(y, self.inv_Kxx_y)
        complexity = -np.sum(np.log(np.diag(L)))
        return data_fit + complexity - 0.5 * n * np.log(2 * np.pi)

    # -- prediction ------------------------------------------------------------------
    def Kxoxo(self, xo):
        return self.K(xo, xo)

    def Kxxo(self, xo):
        return self.K(self._x, xo)

    def Kxox(self, xo):
        return self.K(xo, self._x)

    def mean(self, xo):
        return np.dot(self.Kxox(xo), self.inv_Kxx_y)

    def cov(self, xo):
        Kxox = self.Kxox(xo)
        return self.Kxoxo(xo) - np.dot(Kxox, np.dot(self.inv_Kxx, Kxox.T))

    def plot(self, *args, **kwargs):
        pass
