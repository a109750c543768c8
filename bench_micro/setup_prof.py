import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import _lib, synthetic, util
n_prob, ns = 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 128
opt = synthetic.options(ns)
x_s0, _ = synthetic.observations(ns)
x_s, l_s, x_c = np.tile(x_s0, (n_prob, 1)), np.empty((n_prob, ns)), np.zeros((n_prob, 16))
nc = np.zeros(n_prob, dtype=np.int32)
for p in range(n_prob):
    l_s[p] = synthetic.likelihood(ns, synthetic.problem_shift(p))(x_s0)
    rs = np.random.RandomState(synthetic.SEED + p)
    xc = rs.uniform(x_s0.min() - 2.0, x_s0.max() + 2.0, opt["n_candidate"])
    util.filter_candidates(xc, x_s0, opt["candidate_thresh"])
    xc = np.sort(xc[~np.isnan(xc)]); nc[p] = xc.size; x_c[p, :xc.size] = xc
batch = _lib.Batch(n_prob, ns)
hyp = np.tile(list(synthetic.PARAMS_TL) + list(synthetic.PARAMS_L), (n_prob, 1))
prior = np.tile([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]], (n_prob, 1))
for _ in range(3):
    t0 = time.perf_counter(); info = batch.setup(np.full(n_prob, ns), nc, x_s, l_s, x_c, hyp, prior); torch.cuda.synchronize()
    print("setup %d x ns=%d: %.2f ms" % (n_prob, ns, (time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); batch.setup_device(); print("setup_device: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
