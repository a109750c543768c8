"""world_size-2 gloo tests (CPU) of the multi-rank host logic: sharding, score all-gather, the
(min, first index) exchange of choose_next and the C4 loss all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesian_quadrature_b200 import dist as bqdist


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 10 ** 6, 10 ** 7 + 3):
        for W in (1, 2, 3, 8):
            b = [bqdist.shard_bounds(n, W, r) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_combine_argmin_ties_and_nan():
    assert bqdist.combine_argmin([[1.0, 5], [0.5, 9], [0.5, 7]]) == (0.5, 7)
    assert bqdist.combine_argmin([[np.nan, 0], [2.0, 3]]) == (2.0, 3)


def test_cyclic_shard_and_its_index_map():
    """Block-cyclic shards partition the vector, and the global index the exchange kernel computes for a local index
    (include/bq_b200.h, bqb_choose_step_exchange) points back at the same element."""
    x = np.arange(24000, dtype=np.float64)
    for W, blk in ((1, 1000), (2, 1000), (3, 2000), (8, 500)):
        parts = [bqdist.cyclic_shard(x, W, r, blk) for r in range(W)]
        assert sorted(np.concatenate(parts).tolist()) == x.tolist()
        for r, part in enumerate(parts):
            i = np.arange(part.size)
            g = ((i // blk) * W + r) * blk + i % blk
            assert np.array_equal(x[g], part)
            assert (np.diff(g) > 0).all()              # local order = global order: "first minimiser" survives sharding
    with pytest.raises(ValueError):
        bqdist.cyclic_shard(x, 7, 0, 1000)


def test_interleaved_shards_partition_any_length():
    """Block-interleaved shards (the sharded BQ calls): a partition of 0 .. n-1 for any n, ascending within a rank
    (so the first local minimiser is the rank's smallest global one), balanced to within one block, and
    interleaved_global is the index map."""
    for n in (0, 1, 4095, 4096, 4097, 100003):
        for W in (1, 2, 3, 8):
            for blk in (4096, 1000):
                parts = [bqdist.interleaved_indices(n, W, r, blk) for r in range(W)]
                allidx = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
                assert np.array_equal(np.sort(allidx), np.arange(n))
                sizes = [p.size for p in parts]
                assert max(sizes) - min(sizes) <= blk
                for r, p in enumerate(parts):
                    assert (np.diff(p) > 0).all()
                    for i in (0, p.size // 2, p.size - 1):
                        if 0 <= i < p.size:
                            assert bqdist.interleaved_global(i, W, r, blk) == p[i]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        full = rs.standard_normal(n)
        full[[n // 3, n // 3 + 4]] = full.min() - 1.0           # a tie that spans positions, first one must win
        lo, hi = bqdist.shard_bounds(n, world, rank)
        local = torch.from_numpy(full[lo:hi].copy())
        gathered = bqdist.all_gather_scores(local, n)
        ok_gather = bool((gathered.numpy() == full).all())
        for blk in (4, 4096):                                   # block-interleaved shards scattered back
            mine = bqdist.interleaved_indices(n, world, rank, blk)
            g2 = bqdist.all_gather_scores(torch.from_numpy(full[mine].copy()), n, block=blk)
            ok_gather = ok_gather and bool((g2.numpy() == full).all())
            li = int(np.argmin(full[mine])) if mine.size else 0
            mn2, idx2 = bqdist.all_argmin(float(full[mine].min()) if mine.size else float("inf"),
                                          bqdist.interleaved_global(li, world, rank, blk) if mine.size else 0, 0)
            ok_gather = ok_gather and idx2 == int(np.argmin(full)) and mn2 == full.min()
        mn, idx = bqdist.all_argmin(float(local.min()), int(local.argmin()), lo)
        ok_argmin = (idx == int(np.argmin(full))) and (mn == full.min())
        part = torch.from_numpy(full[lo:hi].copy()).sum().reshape(1)
        loss = bqdist.all_reduce_loss(part, n)
        ok_loss = abs(float(loss) - full.mean()) < 1e-12
        q.put((rank, ok_gather, ok_argmin, ok_loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 1001])
def test_two_rank_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, a, l in res:
        assert g and a and l, (rank, g, a, l)


def _raise_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        bqdist.raise_together(None)                       # nobody failed: returns on every rank
        out.append("ok")
        try:
            bqdist.raise_together(ValueError("rank 1 failed") if rank == 1 else None)
            out.append("no raise")
        except ValueError as e:
            out.append("own:" + str(e))
        except RuntimeError as e:
            out.append("peer:" + str(e)[:30])
        # the collectives that follow still line up: nobody is left waiting
        t = torch.ones(1)
        dist.all_reduce(t)
        out.append(float(t))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_a_failure_on_one_rank_raises_on_all():
    """a rank whose marginal_loss raised must not leave the others in the all-gather / all-reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_raise_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == "ok" and res[0][1].startswith("peer:sharded call aborted") and res[0][2] == 2.0
    assert res[1] == ["ok", "own:rank 1 failed", 2.0]
