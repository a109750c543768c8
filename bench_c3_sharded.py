"""BASELINE configs[2]: 1-D BQ with 256 observations, 10^7 query points sharded across the ranks (block-cyclic), argmin of
the expected variance exchanged by the reduction kernel.  STRONG scaling: the 10^7 points are fixed.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench_c3_sharded.py"""
import json, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.getcwd())
from bayesian_quadrature_b200 import BQ, GaussianKernel, synthetic
from bayesian_quadrature_b200 import dist as bqdist
local = int(os.environ.get("LOCAL_RANK", 0)); W = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if W > 1 else 0
ns, na, blk, steps = 256, 10 ** 7, 10 ** 4, 10
bq = synthetic.make_bq(BQ, GaussianKernel, ns)
batch = bq._device_model().batch
x = synthetic.query_grid(ns, na)
x_d = torch.from_numpy(bqdist.cyclic_shard(x, W, rank, blk)).to(dev)
esm, ev = torch.empty_like(x_d), torch.empty_like(x_d)
ex = bqdist.PairExchange.create(dev)
assert ex is not None
for _ in range(3):
    got = ex.step(batch, x_d, esm, ev, 0, cyclic_block=blk)
if W > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ex.step_async(batch, x_d, esm, ev, 0, cyclic_block=blk)
e1.record()
got = ex.result()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if W > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "C3 sharded (strong scaling)", "ns": ns, "na_total": na, "n_gpus": W, "ms_per_step": float(t.item()),
                      "evals_per_s": na / (float(t.item()) * 1e-3), "argmin_index": got[1], "min": got[0],
                      "shards": "block-cyclic, %d points per block" % blk, "exchange": "fused reduce + p2p exchange kernel"}))
if W > 1: dist.destroy_process_group()
