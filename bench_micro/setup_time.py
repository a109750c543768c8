import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import _lib, synthetic
def run(B, ns, nc_=2):
    opt = synthetic.options(ns)
    x_s0, l0 = synthetic.observations(ns)
    x_s, l_s = np.tile(x_s0, (B, 1)), np.tile(l0, (B, 1))
    x_c = np.zeros((B, 16)); x_c[:, 0] = x_s0[3] + 0.6; x_c[:, 1] = x_s0[-5] + 0.6
    hyp = np.tile(list(synthetic.PARAMS_TL) + list(synthetic.PARAMS_L), (B, 1))
    prior = np.tile([opt["x_mean"], opt["x_var"], opt["candidate_thresh"]], (B, 1))
    b = _lib.Batch(B, ns)
    ns_a, nc_a = np.full(B, ns, dtype=np.int32), np.full(B, nc_, dtype=np.int32)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = b.setup(ns_a, nc_a, x_s, l_s, x_c, hyp, prior)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert (info["status"] == 0).all()
    print("setup B=%d ns=%d: best %.2f ms (%.1f us per instance), Zm %.6g" % (B, ns, min(ts), min(ts) / B * 1e3, info["Z_mean"][0]))
    b.close()
run(1, 64); run(1, 128); run(1, 256); run(1024, 64); run(16384, 128)
