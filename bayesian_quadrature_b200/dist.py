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


def cyclic_shard(x, world_size, rank, block):
    """Block-cyclic shard of a vector whose length is a multiple of world_size * block: blocks rank, rank + W, rank + 2W, ...
    Unlike contiguous shards, every rank gets the same mix of cheap (far from every observation) and expensive query
    points, which matters since band skipping makes their cost differ."""
    x = np.asarray(x)
    if x.size % (world_size * block):
        raise ValueError("vector length must be a multiple of world_size * block")
    return np.ascontiguousarray(x.reshape(-1, world_size, block)[:, rank, :]).ravel()


#: block of the interleaved shards of the sharded BQ calls below
SHARD_BLOCK = 4096


def interleaved_indices(n, world_size, rank, block=SHARD_BLOCK):
    """Global indices (ascending) of rank's shard of a vector of ANY length n under block-interleaved sharding: blocks of
    `block` consecutive points are dealt round-robin (block j to rank j % W; the last block may be short).  Like
    cyclic_shard, every rank gets the same mix of cheap and expensive query points -- contiguous shards leave the ranks that
    hold the observed region with several times the work of the others under band skipping -- but no divisibility is
    required.  Local index i is global index ((i // block) * W + rank) * block + i % block."""
    n, W, block = int(n), int(world_size), int(block)
    nblk = (n + block - 1) // block
    mine = np.arange(rank, nblk, W, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < n]


def interleaved_global(i, world_size, rank, block=SHARD_BLOCK):
    """Global index of local index i of rank's interleaved shard."""
    return ((int(i) // block) * int(world_size) + int(rank)) * block + int(i) % block


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


class PairExchange(object):
    """Cross-rank exchange of the (min, first global index) pair of a sharded choose_next step, done by the reduction
    kernel itself (``bqb_choose_step_exchange``): every rank's exchange buffer is peer-mapped symmetric memory, the
    kernel stores its pair into all of them over NVLink and reduces what the others stored into its own.  The result
    lands in page-locked host memory, so a step needs no NCCL collective and no device-to-host copy call.

    With one rank the buffer is ordinary device memory.  ``PairExchange.create`` returns None when symmetric memory is
    not available (the caller then uses the NCCL all-gather of ``all_argmin`` / ``combine_argmin``)."""

    def __init__(self, device, buf, ptrs, keep=None):
        import ctypes
        self.world, self.rank = world()
        self.buf, self._keep = buf, keep
        self.ptrs = (ctypes.c_void_p * self.world)(*[int(q) for q in ptrs])
        self.out = torch.zeros(4, dtype=torch.float64).pin_memory()
        self._out_np = self.out.numpy()
        self.seq = 0

    @classmethod
    def create(cls, device):
        W, _ = world()
        import os
        if W == 1 or os.environ.get("BQB_EXCHANGE") == "self":      # "self": diagnostic -- every rank exchanges with itself only
            buf = torch.zeros(2 * 4, dtype=torch.float64, device=device)
            ex = cls(device, buf, [buf.data_ptr()])
            ex.world, ex.rank = 1, 0
            return ex
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(2 * W * 4, dtype=torch.float64, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            ptrs = list(hdl.buffer_ptrs)
            torch.cuda.synchronize(device)
            dist.barrier()
            return cls(device, buf, ptrs, keep=hdl)
        except Exception as e:                        # no peer access / API not present: NCCL path
            import warnings
            warnings.warn("PairExchange: symmetric memory unavailable (%s); using the NCCL all-gather" % (e,))
            return None

    def step_async(self, batch, x_d, esm, ev, offset, inst=0, cyclic_block=0):
        """Enqueue scoring + fused reduce/exchange on the current stream and return at once (the ranks stay in step on the
        device through the exchange itself); `result()` waits and reads the last step's pair."""
        self.seq += 1
        batch.choose_step_exchange(x_d, esm, ev, offset, self.ptrs, self.world, self.rank, self.seq, self.out.data_ptr(), inst=inst,
                                   cyclic_block=cyclic_block)

    def step(self, batch, x_d, esm, ev, offset, inst=0, cyclic_block=0):
        """One step, synchronously: (min, global index).  ``cyclic_block`` > 0: this rank holds blocks rank, rank + W, ... of
        ``cyclic_block`` points (``cyclic_shard``) instead of one contiguous shard starting at ``offset``."""
        self.step_async(batch, x_d, esm, ev, offset, inst=inst, cyclic_block=cyclic_block)
        return self.result()

    def result(self):
        torch.cuda.current_stream().synchronize()
        o = self._out_np
        if o[3] != self.seq or o[2] != 0:
            raise RuntimeError("PairExchange: a rank did not deliver its pair for step %d" % self.seq)
        return float(o[0]), int(o[1])


def all_gather_scores(local, n_total, block=0):
    """The full score vector from the ranks' shards: contiguous shards in shard_bounds order (block = 0), or
    block-interleaved shards (interleaved_indices) scattered back to their global positions."""
    W, _ = world()
    if W == 1:
        return local
    if block:
        index = [interleaved_indices(n_total, W, r, block) for r in range(W)]
        sizes = [ix.size for ix in index]
    else:
        sizes = [hi - lo for lo, hi in (shard_bounds(n_total, W, r) for r in range(W))]
    pad = max(max(sizes), 1)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(W * pad, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    if not block:
        return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(W)])
    full = torch.empty(n_total, dtype=local.dtype, device=local.device)
    for r in range(W):
        full[torch.from_numpy(index[r]).to(local.device)] = out[r * pad: r * pad + sizes[r]]
    return full


def raise_together(err, device=None):
    """Collective error check: every rank calls it with its own exception (or None) BEFORE entering the data
    collectives; if any rank failed, all ranks raise (the failing ones their own exception), so that a setup failure
    or a bad-score flag on one shard cannot leave the other ranks waiting in an all-gather / all-reduce."""
    W, rank = world()
    if W > 1:
        flag = torch.tensor([1.0 if err is not None else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if flag.item() and err is None:
            err = RuntimeError("sharded call aborted: another rank failed (see its exception)")
    if err is not None:
        raise err


def all_reduce_loss(partial_sum, n_samples_total):
    """C4 sharded by hyper-parameter sample: every rank holds the SUM of -esm over its samples;
    the marginal loss is the all-reduced sum divided by the total number of samples."""
    W, _ = world()
    if W > 1:
        dist.all_reduce(partial_sum, op=dist.ReduceOp.SUM)
    return partial_sum / n_samples_total


# --------------------------------------------------------------------------------------------------
# Sharded forms of the BQ scoring calls (one process per GPU; every rank holds the same BQ object)

def expected_Z_var_sharded(bq, x_a):
    """``bq.expected_Z_var(x_a)`` with the query points sharded across ranks (C2 / C3): each rank scores its
    block-interleaved shard (interleaved_indices: balanced work under band skipping) on its GPU, the shards are
    all-gathered and scattered back, every rank returns the full vector."""
    W, rank = world()
    x_a = np.ascontiguousarray(x_a, dtype=np.float64)
    mine = interleaved_indices(x_a.shape[0], W, rank)
    m = mine.size
    dev = torch.device("cuda", bq.device)
    model = bq._device_model()
    x_d = torch.from_numpy(x_a[mine]).to(dev)
    esm = torch.empty(max(m, 1), dtype=torch.float64, device=dev)
    ev = torch.empty_like(esm)
    pair = torch.empty(2, dtype=torch.float64, device=dev)
    if m:
        model.batch.choose_step_device(x_d, esm[:m], ev[:m], pair, offset=0)
    return all_gather_scores(ev[:m], x_a.shape[0], block=SHARD_BLOCK).cpu().numpy()


def choose_next_sharded(bq, x_a, hypers_tl, hypers_l, params, shard="points"):
    """Deterministic ``choose_next`` (first minimiser of the marginal loss, bq.py:660-663) over the hyper-parameter
    samples ``hypers_tl`` / ``hypers_l`` (as returned by ``bq.sample_hypers``; identical on every rank).

    shard="points"  (C2/C3/C4-by-points): every rank scores its block-interleaved shard of ``x_a`` under ALL samples; the mean over
                    samples is taken in sample order on the device (bit-identical to one GPU); the ranks exchange one
                    (min, first global index) pair each.
    shard="samples" (C4-by-samples): every rank scores ALL points under its contiguous subset of samples; the partial
                    sums are all-reduced (summation order differs from one GPU by ~1e-16 relative), then argmin.
    Returns (x_next, global index, minimal loss)."""
    W, rank = world()
    x_a = np.ascontiguousarray(x_a, dtype=np.float64)
    n = len(hypers_tl)
    dev = torch.device("cuda", bq.device)
    if shard not in ("points", "samples"):
        raise ValueError("shard must be 'points' or 'samples'")
    err = None
    if shard == "points":
        mn, idx = float("inf"), 0
        try:
            mine = interleaved_indices(x_a.shape[0], W, rank)
            if mine.size:
                loss, batch = bq.marginal_loss(x_a[mine], hypers_tl, hypers_l, params)
                try:
                    mn, idx = batch.argmin_device(loss)     # first local minimiser = smallest global index among this rank's
                finally:
                    batch.close()
                idx = interleaved_global(idx, W, rank)
        except Exception as e:                              # noqa: BLE001 -- re-raised on every rank below
            err = e
        raise_together(err, device=dev)
        mn, idx = all_argmin(mn, idx, 0, device=dev)
    else:
        lo, hi = shard_bounds(n, W, rank)
        total = torch.zeros(x_a.shape[0], dtype=torch.float64, device=dev)
        try:
            if hi > lo:
                part, batch = bq.marginal_loss(x_a, hypers_tl[lo:hi], hypers_l[lo:hi], params, reduce="sum")
                batch.close()
                total += part
        except Exception as e:                              # noqa: BLE001
            err = e
        raise_together(err, device=dev)
        loss = all_reduce_loss(total, n)
        mn, idx = bq._device_model().batch.argmin_device(loss)      # the object's own resident batch: any rank, even one without samples
    return x_a[idx], idx, mn
