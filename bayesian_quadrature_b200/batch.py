"""``BatchBQ`` — P independent 1-D BQ problems advanced in lock-step on one GPU (BASELINE config C5:
a batch of problems doing rounds of expected-variance active sampling, problems sharded across
GPUs with no collective).

Per problem the semantics are those of the reference's ``BQ`` object under fixed hyper-parameters:
``init`` draws and filters candidates (bq.py:967-991), a round scores a query grid with
``expected_squared_mean`` (bq.py:379-402), picks the first minimiser of ``-esm`` (bq.py:660-663,
deterministic mode), and ``add_observation`` merges or appends the new point and re-initialises
(bq.py:683-701).  Differences forced by batching, both documented in SURVEY appendix A.4: every
problem has its own ``RandomState(seed + p)`` candidate stream (the reference uses one global
RNG), and device arrays are padded to the batch's largest ``ns`` / ``nc``.

All per-problem bookkeeping is numpy-vectorised over the batch; scoring, the per-problem argmin
and the model setup run on the device through the C-ABI.
"""
import numpy as np

from . import _lib


def filter_candidates_batch(x_c, x_s, ns, thresh):
    """Vectorised ``bq_c.filter_candidates`` (bq_c.pyx:601-650) over a batch: x_c [P, m] is modified
    in place (NaN = removed), x_s [P, cap] holds ns[p] valid observations per row."""
    P, m = x_c.shape
    xt = np.ascontiguousarray(x_c.T)         # [m, P]: contiguous per-candidate vectors
    changed = True
    with np.errstate(invalid="ignore"):
        while changed:                       # the reference repeats the pair sweep while anything merged
            changed = False
            for i in range(m):
                for j in range(i + 1, m):
                    hit = np.abs(xt[i] - xt[j]) < thresh          # NaN (removed) compares False
                    if hit.any():
                        xt[i] = np.where(hit, (xt[i] + xt[j]) / 2.0, xt[i])
                        xt[j] = np.where(hit, np.nan, xt[j])
                        changed = True
    x_c[:] = xt.T
    # drop candidates closer than thresh to an observation: the nearest observation is one of the two
    # neighbours in sorted order, found with a vectorised binary search (same |x_c - x_s| < thresh test)
    cap = x_s.shape[1]
    valid = np.arange(cap)[None, :] < ns[:, None]
    xs = np.sort(np.where(valid, x_s, np.inf), axis=1)
    xs = np.concatenate([np.full((P, 1), -np.inf), xs, np.full((P, 1), np.inf)], axis=1)    # sentinels at both ends
    rows = np.arange(P)
    steps = int(np.ceil(np.log2(cap + 2))) + 1
    for i in range(m):
        q = x_c[:, i]
        lo = np.zeros(P, dtype=np.int64)                 # invariant: xs[lo] < q <= xs[hi] (NaN q: comparisons False)
        hi = np.full(P, cap + 1, dtype=np.int64)
        for _ in range(steps):
            mid = (lo + hi) >> 1
            go = xs[rows, mid] < q
            lo = np.where(go, mid, lo)
            hi = np.where(go, hi, mid)
        with np.errstate(invalid="ignore"):
            close = (np.abs(q - xs[rows, lo]) < thresh) | (np.abs(q - xs[rows, hi]) < thresh)
        x_c[close, i] = np.nan


class BatchBQ(object):
    def __init__(self, x_s, l_s, params_tl, params_l, n_candidate, candidate_thresh, x_mean, x_var, seed=0,
                 device=0, ns_reserve=0, device_resident=False):
        """x_s, l_s: [P, ns0] initial observations; params_*: (h, w, s) shared by all problems or [P, 3];
        x_mean / x_var scalars or [P].  ``ns_reserve`` extra observation slots are pre-allocated.

        ``device_resident=True`` keeps the observations, the candidate generators (one MT19937 stream per problem,
        bit-identical to ``np.random.RandomState(seed + p)``) and the candidate filter on the GPU: a round then moves
        only the chosen points and their likelihood values across PCIe.  Both modes produce the
        same candidates, choices and Z estimates (tests/test_gpu_bq_api.py)."""
        x_s, l_s = np.asarray(x_s, dtype=np.float64), np.asarray(l_s, dtype=np.float64)
        if x_s.ndim != 2 or x_s.shape != l_s.shape:
            raise ValueError("x_s and l_s must be [P, ns] arrays of the same shape")
        if (l_s <= 0).any():
            raise ValueError("l_s contains zero or negative values")
        self.P, ns0 = x_s.shape
        self.cap = ns0 + int(ns_reserve)
        self.x_s = np.zeros((self.P, self.cap)); self.x_s[:, :ns0] = x_s
        self.l_s = np.ones((self.P, self.cap)); self.l_s[:, :ns0] = l_s
        self.ns = np.full(self.P, ns0, dtype=np.int32)
        bc = lambda v, k: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.P, k)))
        self.hyp = np.concatenate([bc(params_tl, 3), bc(params_l, 3)], axis=1)
        self.prior = np.stack([bc(x_mean, 1)[:, 0], bc(x_var, 1)[:, 0], np.full(self.P, float(candidate_thresh))], axis=1)
        self.n_candidate, self.thresh = int(n_candidate), float(candidate_thresh)
        if self.n_candidate > _lib.NC_MAX:
            raise NotImplementedError("n_candidate > %d" % _lib.NC_MAX)
        self.device_resident = bool(device_resident)
        self.seed = int(seed)
        self.rngs = None if self.device_resident else [np.random.RandomState(seed + p) for p in range(self.P)]
        self._ns_max = ns0                                 # upper bound of ns over the problems (device mode)
        # capacity class chosen for ns0 + ns_reserve from the start (both modes, so that they run the same kernels): a
        # batch that is known to grow does not migrate to the next class after the first appended observation
        self._class_hint = min(self.cap, 512)
        self.device = int(device)
        self.batch = None
        self.x_c = np.zeros((self.P, _lib.NC_MAX))
        self.nc = np.zeros(self.P, dtype=np.int32)
        self.init()

    def close(self):
        if self.batch is not None:
            self.batch.close()
            self.batch = None

    # ---------------------------------------------------------------- bq.py:132-171 / :967-991 per problem
    def _check_info(self, info):
        bad = np.nonzero(info["status"])[0]
        if bad.size:
            raise np.linalg.LinAlgError("problem %d failed device setup with status %d" % (bad[0], info["status"][bad[0]]))
        self.info = info
        return info

    def _init_device(self):
        """Device-resident (re-)initialisation: candidates are drawn, filtered and sorted on the GPU."""
        if self.batch is None:                             # first call: upload the observations, seed the generators
            cap_class = _lib.load().bqb_ns_capacity(max(int(self._ns_max), self._class_hint))
            if cap_class < 0:
                raise NotImplementedError("more than 512 observations per problem")
            self.batch = _lib.Batch(self.P, cap_class, device=self.device)
            self._cap_class = cap_class
            stride = min(self.cap, cap_class)
            self.batch.stage(self.ns, self.x_s[:, :stride], self.l_s[:, :stride], self.hyp, self.prior)
            self.batch.seed_candidates((self.seed + np.arange(self.P)) & 0xffffffff)
        self.batch.draw_candidates(self.n_candidate)
        info = self._check_info(self.batch.setup_device())
        self._host_stale = True
        return info

    def sync_host(self):
        """Device mode: refresh the host mirrors x_s, l_s, ns, x_c, nc from the GPU."""
        if self.device_resident and getattr(self, "_host_stale", False):
            st = self.batch.get_staged()
            cap = st["x_s"].shape[1]
            if cap > self.cap:
                self.x_s = np.concatenate([self.x_s, np.zeros((self.P, cap - self.cap))], axis=1)
                self.l_s = np.concatenate([self.l_s, np.ones((self.P, cap - self.cap))], axis=1)
                self.cap = cap
            self.x_s[:, :cap], self.l_s[:, :cap] = st["x_s"], st["l_s"]
            self.ns, self.nc, self.x_c = st["ns"], st["nc"], st["x_c"]
            self._host_stale = False

    def _grow_device(self):
        """Move the device batch to the next capacity class (observations and generator states are carried over)."""
        st = self.batch.get_staged()
        mt, pos = self.batch.rng_get()
        old_cap = self.batch.capacity
        cap_class = _lib.load().bqb_ns_capacity(old_cap + 1)
        if cap_class < 0:
            raise NotImplementedError("more than 512 observations per problem")
        self.batch.close()
        self.batch = _lib.Batch(self.P, cap_class, device=self.device)
        self._cap_class = cap_class
        self.batch.stage(st["ns"], st["x_s"], st["l_s"], self.hyp, self.prior)
        self.batch.rng_set(mt, pos)

    def init(self):
        if self.device_resident:
            return self._init_device()
        w_tl = self.hyp[:, 1]
        idx = np.arange(self.cap)[None, :] < self.ns[:, None]
        lo = np.where(idx, self.x_s, np.inf).min(axis=1) - w_tl
        hi = np.where(idx, self.x_s, -np.inf).max(axis=1) + w_tl
        xc = np.stack([self.rngs[p].uniform(lo[p], hi[p], self.n_candidate) for p in range(self.P)])
        filter_candidates_batch(xc, self.x_s, self.ns, self.thresh)
        xc.sort(axis=1)                                   # NaNs sort last
        self.nc = (~np.isnan(xc)).sum(axis=1).astype(np.int32)
        self.x_c[:] = 0.0
        self.x_c[:, :self.n_candidate] = np.nan_to_num(xc, nan=0.0)
        cap_class = _lib.load().bqb_ns_capacity(max(int(self.ns.max()), self._class_hint))
        if cap_class < 0:
            raise NotImplementedError("more than 512 observations per problem")
        if self.batch is None or self._cap_class != cap_class:       # crossed a kernel capacity class: new device batch
            self.close()
            self.batch = _lib.Batch(self.P, cap_class, device=self.device)
            self._cap_class = cap_class
        stride = min(self.cap, cap_class)
        info = self.batch.setup(self.ns, self.nc, self.x_s[:, :stride], self.l_s[:, :stride], self.x_c, self.hyp, self.prior)
        return self._check_info(info)

    def Z_mean(self):
        return self.info["Z_mean"]

    def Z_var(self):
        return self.info["Z_var"]

    # ---------------------------------------------------------------- one active-sampling round
    def choose_next(self, x_a, on_device=False):
        """Deterministic choose_next of every problem over a shared grid ``x_a`` [na] or per-problem grids
        [P, na]: index and location of the first maximiser of expected_squared_mean.  ``x_a`` may be a float64
        CUDA tensor (it is then not uploaded again); with ``on_device=True`` the result is a pair of CUDA tensors."""
        import torch
        dev = torch.device("cuda", self.device)
        if isinstance(x_a, torch.Tensor):
            x_d = x_a.to(dev, torch.float64).contiguous()
            x_a = None
        else:
            x_a = np.ascontiguousarray(x_a, dtype=np.float64)
            x_d = torch.from_numpy(x_a).to(dev)
        na = x_d.shape[-1]
        if getattr(self, "_esm", None) is None or self._esm.shape != (self.P, na):
            self._esm = torch.empty(self.P, na, dtype=torch.float64, device=dev)
            self._mins = torch.empty(self.P, dtype=torch.float64, device=dev)
            self._idxs = torch.empty(self.P, dtype=torch.int64, device=dev)
            self._flags = torch.zeros(self.P, dtype=torch.int32, device=dev)
        self.batch.score_device(x_d, self._esm, None, None, self._flags)
        self._esm.neg_()                                  # loss = -esm (bq.py:660)
        self.batch.argmin_rows_device(self._esm, self._mins, self._idxs)
        bad = int((self._flags & (_lib.ST_ESM_BAD | _lib.ST_EM_BAD | _lib.ST_XA_BAD)).ne(0).any().item())
        if bad:
            fl = self._flags.cpu().numpy()
            raise RuntimeError("invalid expected squared mean in problem %d" % int(np.argmax(fl & 112 != 0)))
        if on_device:
            x_next = x_d[self._idxs] if x_d.dim() == 1 else x_d.gather(1, self._idxs[:, None])[:, 0]
            return self._idxs, x_next
        idx = self._idxs.cpu().numpy()
        if x_a is None:
            x_a = x_d.cpu().numpy()
        x_next = x_a[idx] if x_a.ndim == 1 else x_a[np.arange(self.P), idx]
        return idx, x_next

    def add_observations(self, x_new, l_new):
        """Vectorised ``add_observation`` (bq.py:683-701): average into the nearest observation when it is
        closer than candidate_thresh, append otherwise; then re-initialise every problem.  In device-resident mode
        ``x_new`` / ``l_new`` may be CUDA tensors and nothing but they crosses PCIe."""
        if self.device_resident:
            import torch
            dev = torch.device("cuda", self.device)
            as_dev = lambda v: (v.to(dev, torch.float64).contiguous() if isinstance(v, torch.Tensor)
                                else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev))
            x_d, l_d = as_dev(x_new), as_dev(l_new)
            if x_d.shape != (self.P,) or l_d.shape != (self.P,):
                raise ValueError("x_new and l_new must have one entry per problem")
            if not bool((l_d > 0).all().item()):
                raise ValueError("l_s contains zero or negative values")
            if self._ns_max + 1 > self.batch.capacity:
                self._grow_device()
            self.batch.add_observations(x_d, l_d)
            self._ns_max += 1
            return self._init_device()
        x_new, l_new = np.asarray(x_new, dtype=np.float64), np.asarray(l_new, dtype=np.float64)
        valid = np.arange(self.cap)[None, :] < self.ns[:, None]
        d = np.abs(x_new[:, None] - self.x_s)
        d[~valid] = np.inf
        c = d.argmin(axis=1)
        merge = d[np.arange(self.P), c] < self.thresh
        rows = np.nonzero(merge)[0]
        self.x_s[rows, c[rows]] = (self.x_s[rows, c[rows]] + x_new[rows]) / 2.0
        self.l_s[rows, c[rows]] = (self.l_s[rows, c[rows]] + l_new[rows]) / 2.0
        rows = np.nonzero(~merge)[0]
        if rows.size and self.ns[rows].max() >= self.cap:
            grow = max(16, self.cap // 4)
            self.x_s = np.concatenate([self.x_s, np.zeros((self.P, grow))], axis=1)
            self.l_s = np.concatenate([self.l_s, np.ones((self.P, grow))], axis=1)
            self.cap += grow
        self.x_s[rows, self.ns[rows]] = x_new[rows]
        self.l_s[rows, self.ns[rows]] = l_new[rows]
        self.ns[rows] += 1
        return self.init()
