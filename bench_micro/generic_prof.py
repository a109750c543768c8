import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import BQ, GaussianKernel, synthetic
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bq = synthetic.make_bq(BQ, GaussianKernel, ns)
x_a = synthetic.query_grid(ns, 10 ** 5)
for _ in range(3):
    t0 = time.perf_counter(); bq.expected_squared_mean(x_a); print("ns=%d 1e5 points: %.2f ms" % (ns, (time.perf_counter() - t0) * 1e3))
