"""Multi-GPU plumbing of the scoring path: one process per GPU, query points
(or hyper-parameter samples, or independent problems) sharded across ranks with NO data-path
collective; the only exchanges are

* an all-gather of per-point scores when the caller wants the whole vector
  (``expected_Z_var`` over a sharded grid, or the reference's random tie-break of choose_next), and
* an all-gather of one ``(min, first global index)`` pair per rank for the deterministic argmin
  of ``choose_next`` (NCCL has no MINLOC; W pairs are reduced locally on every rank).

Works with any torch.distributed backend: ``nccl`` on the GPUs, ``gloo`` in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_bounds(n, world_size, rank):
    """Contiguous, balanced shard [lo, hi) of n items: the first n % W ranks get one extra item."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_argmin(pairs):
    """pairs: [W, 2] array of (local min, global index of its first occurrence); NaN mins never win.
    Returns (min, index) with ties resolved to the smallest global index — np.argmin semantics."""
    pairs = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
    vals = np.where(np.isnan(pairs[:, 0]), np.inf, pairs[:, 0])
    order = np.lexsort((pairs[:, 1], vals))
    return float(pairs[order[0], 0]), int(pairs[order[0], 1])


def all_argmin(local_min, local_idx, offset, device=None):
    """Deterministic global (min, argmin) from each rank's local result over its shard."""
    W, _ = world()
    mine = torch.tensor([float(local_min), float(local_idx + offset)], dtype=torch.float64, device=device)
    if W == 1:
        return float(mine[0]), int(mine[1])
    out = torch.empty(W * 2, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, mine)
    return combine_argmin(out.cpu().numpy())


def all_gather_scores(local, n_total):
    """Concatenate the ranks' contiguous shards (shard_bounds order) into the full score vector."""
    W, _ = world()
    if W == 1:
        return local
    sizes = [shard_bounds(n_total, W, r) for r in range(W)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(W * pad, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return torch.cat([out[r * pad: r * pad + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def all_reduce_loss(partial_sum, n_samples_total):
    """C4 sharded by hyper-parameter sample: every rank holds the SUM of -esm over its samples;
    the marginal loss is the all-reduced sum divided by the total number of samples."""
    W, _ = world()
    if W > 1:
        dist.all_reduce(partial_sum, op=dist.ReduceOp.SUM)
    return partial_sum / n_samples_total
