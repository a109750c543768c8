import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import BatchBQ, synthetic
n_prob, ns, na, rounds = 16384, 128, 4096, 20
opt = synthetic.options(ns)
x0, _ = synthetic.observations(ns)
shifts = np.array([synthetic.problem_shift(p) for p in range(n_prob)])
sp = synthetic.span(ns)
def lik(x, sh):
    npdf = lambda x, m, s: np.exp(-0.5 * ((x - m) / s) ** 2) / (np.sqrt(2 * np.pi) * s)
    return (0.5 * npdf(x, (-0.3 + sh[:, 0]) * sp, 0.16 * sp) + 0.3 * npdf(x, (0.4 + sh[:, 1]) * sp, 0.10 * sp) + 0.2 * npdf(x, (0.1 + sh[:, 2]) * sp, 0.3 * sp))
l0 = np.stack([lik(np.full(n_prob, x), shifts) for x in x0], axis=1)
bb = BatchBQ(np.tile(x0, (n_prob, 1)), l0, synthetic.PARAMS_TL, synthetic.PARAMS_L, opt["n_candidate"], opt["candidate_thresh"], opt["x_mean"], opt["x_var"], seed=synthetic.SEED, ns_reserve=rounds, device_resident=True)
grid = torch.from_numpy(synthetic.query_grid(ns, na)).cuda()
rows = []
for r in range(rounds):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, x_next = bb.choose_next(grid, on_device=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    l_next = torch.from_numpy(lik(x_next.cpu().numpy(), shifts)).cuda(); t2 = time.perf_counter()
    bb.add_observations(x_next, l_next); torch.cuda.synchronize(); t3 = time.perf_counter()
    st = bb.batch.get_staged()
    rows.append((r, int(st["ns"].max()), int(st["nc"].max()), int((st["ns"] + st["nc"]).max()), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
for row in rows: print("round %2d ns_max %d nc_max %d n_max %d score %.1f ms lik %.1f ms update %.1f ms" % row)
