"""Host-side Gaussian-process objects that the ``BQ`` interface exposes (``bq.gp_log_l``,
``bq.gp_l``; reference: bayesian_quadrature/bq.py:147-165 builds them from the third-party
``gp`` package, ``gaussian_processes==1.0.5``, which is not vendored in the reference repo).

These objects carry the *state* (kernel parameters, x, y, jitter) and serve the parts of the
API that are outside the hot path (``l_mean`` / ``l_var`` bq.py:177-231, ``log_lh`` for the
hyper-parameter sampler bq.py:546, plotting, pickling).  The hot path never calls them: scoring
goes through the CUDA library with the parameters read from here.

Interface (the surface BQ and its tests use): ``GaussianKernel(h, w)``,
``PeriodicKernel(h, w, p)``, ``GP(K, x, y, s=0)`` with ``K, x, y, s, params, get_param, set_param,
Kxx, Lxx, inv_Kxx_y, log_lh, Kxoxo, Kxxo, Kxox, mean, cov, copy, plot`` and the ``_memoized`` cache
that every setter clears.
"""
import numpy as np
import scipy.linalg

DTYPE = np.float64
SQRT_2PI = np.sqrt(2 * np.pi)


class Kernel(object):
    names = ()

    def __init__(self, *values):
        if len(values) != len(self.names):
            raise ValueError("%s takes %d parameters" % (type(self).__name__, len(self.names)))
        for name, v in zip(self.names, values):
            setattr(self, name, v)

    def __setattr__(self, name, value):
        if name in self.names:
            value = float(value)
            if not (value > 0):
                raise ValueError("invalid value for %s: %s" % (name, value))
        object.__setattr__(self, name, value)

    @property
    def params(self):
        return np.array([getattr(self, name) for name in self.names], dtype=DTYPE)

    @params.setter
    def params(self, values):
        for name, v in zip(self.names, values):
            setattr(self, name, v)

    def copy(self):
        return type(self)(*self.params)

    def __call__(self, x1, x2):
        diff = np.subtract.outer(np.asarray(x1, dtype=DTYPE), np.asarray(x2, dtype=DTYPE))
        return self.of_difference(diff)

    def __getstate__(self):
        return {name: getattr(self, name) for name in self.names}

    def __setstate__(self, state):
        for name, v in state.items():
            setattr(self, name, v)


class GaussianKernel(Kernel):
    r"""K(x, x') = h^2 N(x | x', w^2) — the *normalised* Gaussian that the reference's closed
    forms integrate (gauss_c.pyx:110)."""
    names = ("h", "w")

    def of_difference(self, diff):
        return (self.h ** 2 / (SQRT_2PI * self.w)) * np.exp(-0.5 * diff ** 2 / self.w ** 2)


class PeriodicKernel(Kernel):
    r"""K(x, x') = h^2 exp(-2 sin^2((x - x') / 2p) / w^2).  BQ with this kernel takes the
    reference's trapezoid `approx_*` path, which is outside the CUDA hot path."""
    names = ("h", "w", "p")

    def of_difference(self, diff):
        return self.h ** 2 * np.exp(-2.0 * np.sin(diff / (2.0 * self.p)) ** 2 / self.w ** 2)


def memoized(f):
    def getter(self):
        cache = self._memoized
        if f.__name__ not in cache:
            cache[f.__name__] = f(self)
        return cache[f.__name__]
    getter.__name__ = f.__name__
    getter.__doc__ = f.__doc__
    return property(getter)


class GP(object):
    def __init__(self, K, x, y, s=0):
        self._memoized = {}
        self.K = K
        self._x = np.array(x, dtype=DTYPE)
        self._y = np.array(y, dtype=DTYPE)
        if self._x.shape != self._y.shape or self._x.ndim != 1:
            raise ValueError("x and y must be 1-D arrays of the same shape")
        self._s = 0.0
        self.s = s

    # ---- state; every setter drops the cache -------------------------------------------
    x = property(lambda self: self._x)
    y = property(lambda self: self._y)
    s = property(lambda self: self._s)

    @x.setter
    def x(self, value):
        self._memoized = {}
        self._x = np.array(value, dtype=DTYPE)

    @y.setter
    def y(self, value):
        self._memoized = {}
        self._y = np.array(value, dtype=DTYPE)

    @s.setter
    def s(self, value):
        value = float(value)
        if not (value >= 0):
            raise ValueError("invalid value for s: %s" % value)
        self._memoized = {}
        self._s = value

    @property
    def params(self):
        return np.append(self.K.params, self._s)

    @params.setter
    def params(self, values):
        self.K.params = values[:-1]
        self.s = values[-1]

    def get_param(self, name):
        return self._s if name == "s" else getattr(self.K, name)

    def set_param(self, name, value):
        if name == "s":
            self.s = value
        else:
            if name not in self.K.names:
                raise AttributeError("unknown parameter %r" % name)
            setattr(self.K, name, value)
            self._memoized = {}

    def copy(self, deep=True):
        new = GP(self.K.copy(), self._x, self._y, s=self._s)
        if hasattr(self, "jitter"):
            new.jitter = np.array(self.jitter)
        return new

    # ---- cached linear algebra -----------------------------------------------------------
    @memoized
    def Kxx(self):
        K = self.K(self._x, self._x)
        K[np.diag_indices_from(K)] += self._s ** 2
        if not np.isfinite(K).all():
            raise ArithmeticError("Kxx contains invalid values")
        return K

    @memoized
    def Lxx(self):
        return np.linalg.cholesky(self.Kxx)          # LinAlgError when not positive definite

    @memoized
    def inv_Kxx_y(self):
        return scipy.linalg.cho_solve((self.Lxx, True), self._y)

    @memoized
    def log_lh(self):
        try:
            L = self.Lxx
        except np.linalg.LinAlgError:
            return -np.inf
        n = self._y.size
        return (-0.5 * np.dot(self._y, self.inv_Kxx_y) - np.log(np.diag(L)).sum() - 0.5 * n * np.log(2 * np.pi))

    # ---- prediction --------------------------------------------------------------------------
    def Kxoxo(self, xo):
        return self.K(xo, xo)

    def Kxxo(self, xo):
        return self.K(self._x, xo)

    def Kxox(self, xo):
        return self.K(xo, self._x)

    def mean(self, xo):
        return np.dot(self.Kxox(xo), self.inv_Kxx_y)

    def cov(self, xo):
        V = scipy.linalg.solve_triangular(self.Lxx, self.Kxxo(xo), lower=True)
        return self.Kxoxo(xo) - np.dot(V.T, V)

    def plot(self, ax, xlim=None, color="k", markercolor="r"):
        x, y = self._x, self._y
        if xlim is None:
            xlim = (x.min(), x.max())
        X = np.linspace(xlim[0], xlim[1], 1000)
        mean = self.mean(X)
        std = np.sqrt(np.clip(np.diag(self.cov(X)), 0, None))
        ax.fill_between(X, mean - std, mean + std, color=color, alpha=0.2)
        ax.plot(X, mean, lw=2, color=color)
        ax.plot(x, y, "o", ms=5, color=markercolor)
        ax.set_xlim(*xlim)
