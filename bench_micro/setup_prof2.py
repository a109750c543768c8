import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import _lib, synthetic
def run(B, ns, nc_=2, cm=True):
    opt = synthetic.options(ns)
    x_s0, l0 = synthetic.observations(ns)
    x_s, l_s = np.tile(x_s0, (B, 1)), np.tile(l0, (B, 1))
    x_c = np.zeros((B, 16)); x_c[:, 0] = x_s0[3] + 0.6; x_c[:, 1] = x_s0[-5] + 0.6
    for j in range(2, nc_): x_c[:, j] = x_s0[-1] + 0.7 * j
    hyp = np.tile(list(synthetic.PARAMS_TL) + list(synthetic.PARAMS_L), (B, 1))
    prior = np.tile([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]], (B, 1))
    b = _lib.Batch(B, ns)
    info = b.setup(np.full(B, ns, dtype=np.int32), np.full(B, nc_, dtype=np.int32), x_s, l_s, x_c, hyp, prior, check_max=cm)
    torch.cuda.synchronize()
    b.close()
run(1, 64, 4); run(1, 128, 2); run(1, 148, 4); run(1, 256, 1); run(300, 148, 4)
