import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import _lib, synthetic
def load_golden(name):
    return dict(np.load('tests/golden/%s.npz' % name))
def run(name, ns, na, reps=8):
    g = load_golden(name)
    nc = g["x_c"].size
    b = _lib.Batch(1, ns)
    hyp = np.concatenate([g["params_tl"], g["params_l"]])
    prior = np.array([float(g["x_mean"]), float(g["x_var"]), float(g["candidate_thresh"])])
    b.setup([ns], [nc], g["x_s"][None], g["l_s"][None], g["x_c"][None], hyp[None], prior[None])
    dev = torch.device("cuda", 0)
    grid = synthetic.query_grid(ns, na)
    x_d = torch.from_numpy(grid).to(dev)
    esm = torch.empty(1, na, dtype=torch.float64, device=dev)
    em = torch.empty_like(esm); st = torch.empty(1, na, dtype=torch.int32, device=dev)
    for _ in range(3): b.score_device(x_d, esm, em, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); b.score_device(x_d, esm, em, st); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ev = torch.empty(na, dtype=torch.float64, device=dev); pair = torch.empty(2, dtype=torch.float64, device=dev)
    for _ in range(3): b.choose_step_device(x_d, esm[0], ev, pair)
    torch.cuda.synchronize(); tf = []
    for _ in range(reps):
        e0.record(); b.choose_step_device(x_d, esm[0], ev, pair); e1.record(); torch.cuda.synchronize(); tf.append(e0.elapsed_time(e1))
    tu = []
    for _ in range(reps):
        e0.record(); b.score_device(x_d, esm); b.expected_var_device(0, esm, ev); b.argmin_pair_device(ev, 0, pair); e1.record(); torch.cuda.synchronize(); tu.append(e0.elapsed_time(e1))
    print("   fused step %.4f ms, unfused step %.4f ms" % (min(tf), min(tu)))
    # parity on the fixture subset
    k = g["grid_idx"].size
    e = esm[0].cpu().numpy()
    if na == int(g["na_full"]):
        err = np.abs(e[g["grid_idx"]] - g["esm"][:k]) / np.abs(g["esm"][:k])
        perr = " max rel err vs reference %.2e" % err.max()
    else:
        perr = ""
    n = ns + nc
    wf = n * n + ns * ns + 2 * (3 * n + 2 * ns) + 40
    t = min(ts)
    print("%s cfg=%s ns=%d na=%d: best %.4f ms  avg %.4f ms  -> %.2f Gevals/s, %.2f TF/s (%.1f%% of 37.16)%s" % (
        name, os.environ.get("BQB_SCORE_CFG", "0"), ns, na, t, np.mean(ts), na / t / 1e6, wf * na / t / 1e9, 100 * wf * na / t / 1e9 / 37.156, perr))
    b.close()
which = sys.argv[1:] or ["c2"]
if "c2" in which: run("c2", 64, 10**6)
if "c5" in which: run("c5", 128, 10**6)
if "c3" in which: run("c3", 256, 10**6)
if "c1" in which: run("c1", 8, 10**6)
