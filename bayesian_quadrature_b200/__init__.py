"""bayesian_quadrature_b200 — B200-native expected-variance active sampling for Bayesian
quadrature: a drop-in ``BQ`` (reference: jhamrick/bayesian-quadrature v0.2.0,
bayesian_quadrature/__init__.py:8-9 exports ``BQ``) whose scoring path runs in hand-written
sm_100a CUDA kernels behind the C-ABI of include/bq_b200.h."""
import logging

logger = logging.getLogger("bayesian_quadrature")

from .bq import BQ                                   # noqa: E402
from .gp import GP, GaussianKernel, PeriodicKernel   # noqa: E402
from .batch import BatchBQ                           # noqa: E402

__all__ = ["BQ", "BatchBQ", "GP", "GaussianKernel", "PeriodicKernel"]
