"""Run under torchrun on >= 2 GPUs (tests/test_gpu_dist.py launches it): the sharded scoring calls must return
what one GPU returns — bit-identical scores and argmin for points-sharding, 1e-12-close loss for samples-sharding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesian_quadrature_b200 import BQ, GaussianKernel, synthetic  # noqa: E402
from bayesian_quadrature_b200 import dist as bqdist                  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, W = dist.get_rank(), dist.get_world_size()
    bq = synthetic.make_bq(BQ, GaussianKernel, 64)           # same seed on every rank: identical candidates
    assert bq.device == local
    x_a = synthetic.query_grid(64, 100003)
    ev = bqdist.expected_Z_var_sharded(bq, x_a)
    ref = bq.expected_Z_var(x_a)                             # the whole vector on this rank's GPU
    assert np.array_equal(ev, ref), "sharded expected_Z_var differs from the single-GPU result"
    hyp = synthetic.hyper_sets(6)
    htl, hl = hyp[:, :2], hyp[:, 2:]
    loss, batch = bq.marginal_loss(x_a, htl, hl, ["h", "w"])
    mn1, idx1 = batch.argmin_device(loss)
    batch.close()
    xp, idx_p, mn_p = bqdist.choose_next_sharded(bq, x_a, htl, hl, ["h", "w"], shard="points")
    assert idx_p == idx1 and mn_p == mn1 and xp == x_a[idx1], (idx_p, idx1, mn_p, mn1)
    xs, idx_s, mn_s = bqdist.choose_next_sharded(bq, x_a, htl, hl, ["h", "w"], shard="samples")
    assert idx_s == idx1 and abs(mn_s - mn1) <= 1e-12 * abs(mn1), (idx_s, idx1, mn_s, mn1)
    # fused reduce + exchange (bqb_choose_step_exchange over peer-mapped symmetric memory) == the single-GPU argmin of
    # expected_Z_var over the whole vector, for several consecutive steps (slot parity, step tags)
    ex = bqdist.PairExchange.create(torch.device("cuda", local))
    assert ex is not None, "symmetric memory unavailable on this box"
    model = bq._device_model()
    lo, hi = bqdist.shard_bounds(x_a.size, W, rank)
    x_d = torch.from_numpy(x_a[lo:hi]).cuda()
    esm_d = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    ev_d = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    want = (float(ref.min()), int(np.argmin(ref)))
    for _ in range(5):
        got = ex.step(model.batch, x_d, esm_d, ev_d, lo)
        assert got == want, (got, want)
    assert np.array_equal(ev_d.cpu().numpy(), ref[lo:hi])
    # the same with block-cyclic shards (global index computed by the exchange kernel)
    blk, n_c = 1000, 100000
    want_c = (float(ref[:n_c].min()), int(np.argmin(ref[:n_c])))
    x_c = torch.from_numpy(bqdist.cyclic_shard(x_a[:n_c], W, rank, blk)).cuda()
    esm_c, ev_c = torch.empty_like(x_c), torch.empty_like(x_c)
    for _ in range(3):
        got = ex.step(model.batch, x_c, esm_c, ev_c, 0, cyclic_block=blk)
        assert got == want_c, (got, want_c)
    if rank == 0:
        print("DIST_GPU_CHECK_OK world=%d argmin=%d exchange=p2p" % (W, idx1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
