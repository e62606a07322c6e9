"""Multi-GPU Gibbs sweep: k-d-tree blocks sharded over ranks, one process per GPU (torch.distributed).

Replaces, for the sweep, what Spark does in the reference:
  - one task per partition (GibbsUpdates.scala:137)            -> blocks placed on ranks by LPT on R_b * E_b
  - the shuffle `.partitionBy(partitioner)` (GU:144)           -> one all-to-all of the clusters whose new block is
                                                                  owned by another rank (NCCL over NVLink; gloo in
                                                                  the CPU tests of this host logic)
  - accumulators (SummaryAccumulators.scala:54-63)             -> all-reduce of A*F + A + 3 int64 words + 1 double
  - broadcast of theta (State.scala:84)                        -> nothing: every rank draws the same theta from the
                                                                  same counter-based stream

The compute engine is duck-typed (`begin/pack/unpack/end/...`), so the exchange logic below is exercised on CPU
with the gloo backend by tests/test_distributed.py; on GPUs it drives dblink_b200.GibbsEngine.
"""
import ctypes as C

import numpy as np

from . import _lib
from .engine import GibbsEngine, KDTreePartitioner, SAMPLERS, _check, _p


def lpt_assign(costs, world):
    """Longest-processing-time placement of blocks on ranks (the reference ships the same heuristic in
    partitioning/LPTScheduler.scala:57-76).  Deterministic: ties by block id, then lowest rank."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda b: (-costs[b], b))
    load = np.zeros(world)
    owner = np.zeros(len(costs), np.int32)
    for b in order:
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += costs[b]
    return owner


def exchange(dist, send_ent, ent_counts, send_rec, rec_counts, ent_words, device, torch):
    """All-to-all of the packed messages.  send_* are int32 tensors concatenated by destination rank;
    *_counts are per-destination MESSAGE counts.  Returns (recv_ent, n_ent_msgs, recv_rec, n_rec_msgs)."""
    world = dist.get_world_size()
    mine = torch.tensor(np.stack([ent_counts, rec_counts], axis=1).reshape(-1), dtype=torch.int64, device=device)
    theirs = torch.empty_like(mine)
    dist.all_to_all_single(theirs, mine)  # counts first
    theirs_h = theirs.cpu().numpy().reshape(world, 2)
    in_ent, in_rec = theirs_h[:, 0], theirs_h[:, 1]
    recv_ent = torch.empty(int(in_ent.sum()) * ent_words, dtype=torch.int32, device=device)
    recv_rec = torch.empty(int(in_rec.sum()) * 3, dtype=torch.int32, device=device)
    dist.all_to_all_single(recv_ent, send_ent, output_split_sizes=[int(c) * ent_words for c in in_ent],
                           input_split_sizes=[int(c) * ent_words for c in ent_counts])
    dist.all_to_all_single(recv_rec, send_rec, output_split_sizes=[int(c) * 3 for c in in_rec],
                           input_split_sizes=[int(c) * 3 for c in rec_counts])
    return recv_ent, int(in_ent.sum()), recv_rec, int(in_rec.sum())


def allreduce_summary(dist, counts, loglik, device, torch):
    """One all-reduce for the integer summary words and the log-likelihood: the counts travel as float64 (exact
    below 2^53; they are bounded by records x attributes)."""
    buf = np.empty(len(counts) + 1, np.float64)
    buf[:-1] = counts
    buf[-1] = loglik
    t = torch.from_numpy(buf).to(device)
    dist.all_reduce(t)
    out = t.cpu().numpy()
    return np.rint(out[:-1]).astype(np.int64), float(out[-1])


class ShardedGibbs:
    """Same surface as GibbsEngine (init_state / sweep / summary / download_state), block-sharded over the ranks of
    the default process group."""

    def __init__(self, indexes, alpha, beta, seed=0, num_files=1, levels=0, split_attrs=()):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.levels, self.split_attrs = levels, list(split_attrs)
        self.eng = GibbsEngine(indexes, alpha, beta, None, seed, num_files, rank=self.rank, world_size=self.world)
        self.A, self.F = self.eng.A, self.eng.F
        self.owner = None
        self._ms = 0.0
        self._stage = {}  # device staging buffers of upload_state
        self.trace = {}  # host wall-clock ms per phase of the sharded sweep, accumulated (rank-local)

    # ---- state ---------------------------------------------------------------------------------------
    def init_state(self, x, file_ids=None, population_size=0):
        """Every rank builds the same replicated initial state (State.deterministic), fits the same k-d tree on it,
        then keeps only the blocks LPT assigns to it."""
        e = self.eng
        e.init_state(x, file_ids, population_size)
        part = KDTreePartitioner(self.levels, self.split_attrs).fit(e.download_state()["y"])
        e.set_partitioner(part)
        self.partitioner = part
        link, blk = e.links()
        P = e.num_partitions
        ent = np.bincount(blk, minlength=P).astype(np.float64)
        rec = np.bincount(blk[link], minlength=P).astype(np.float64)
        self.set_owners(lpt_assign(ent * rec, self.world))

    def _gather_upload(self, name, a, dtype):
        """Full device copy of the host array `a` (identical on every rank): each rank copies 1/world of it over
        PCIe, one in-place all-gather over NVLink completes it."""
        torch, dist, W = self.torch, self.dist, self.world
        flat = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
        n = flat.size
        chunk = (n + W - 1) // W
        key = (name, n)
        full = self._stage.get(key)
        if full is None:
            full = self._stage[key] = torch.empty(W * chunk, dtype=torch.from_numpy(flat[:0].copy()).dtype,
                                                  device=self.device)
        lo, hi = min(n, self.rank * chunk), min(n, (self.rank + 1) * chunk)
        if hi > lo:
            full[lo:hi].copy_(torch.from_numpy(flat[lo:hi]), non_blocking=True)
        dist.all_gather_into_tensor(full, full[self.rank * chunk:(self.rank + 1) * chunk])
        return full

    def upload_state(self, x, file_ids, z, link, y, theta, iteration=0):
        """Every rank is handed the same full state (State.read) and keeps the blocks it owns.  The host-to-device
        traffic is shared: a rank copies its 1/world slice of each array and the slices are all-gathered."""
        if self.owner is None:
            raise RuntimeError("upload_state needs the partitioner and block placement of init_state first")
        x = np.asarray(x)
        y = np.asarray(y)
        if x.ndim != 2 or y.ndim != 2 or x.shape[1] != self.A or y.shape[1] != self.A:
            raise ValueError("state arrays do not match the model")
        dx = self._gather_upload("x", x, np.int32)
        df = self._gather_upload("file", file_ids, np.int32)
        dz = self._gather_upload("z", z, np.uint8)
        dl = self._gather_upload("link", link, np.int32)
        dy = self._gather_upload("y", y, np.int32)
        self.torch.cuda.current_stream().synchronize()  # the engine copies from these buffers on its own stream
        self.eng.upload_state_device(x.shape[0], y.shape[0], dx.data_ptr(), df.data_ptr(), dz.data_ptr(),
                                     dl.data_ptr(), dy.data_ptr(), theta, iteration)
        self.set_owners(self.owner)

    def set_owners(self, owner):
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        self.owner = owner
        _check(_lib.load().dbl_set_block_owners(self.eng._h, _p(owner, _lib.i32p)), "set_block_owners", self.eng._h)
        self._sync_summary()

    def _sync_summary(self):
        L = _lib.load()
        n = L.dbl_summary_words(self.eng._h)
        counts = np.zeros(n, np.int64)
        ll = C.c_double(0.0)
        _check(L.dbl_partial_summary(self.eng._h, _p(counts, _lib.i64p), C.byref(ll)), "partial_summary", self.eng._h)
        g, gll = allreduce_summary(self.dist, counts, ll.value, self.device, self.torch)
        g = np.ascontiguousarray(g, dtype=np.int64)
        _check(L.dbl_set_global_summary(self.eng._h, _p(g, _lib.i64p), gll), "set_global_summary", self.eng._h)

    # ---- transition ----------------------------------------------------------------------------------
    def sweep(self, sampler="PCG-I", n=1):
        torch, dist, L, h = self.torch, self.dist, _lib.load(), self.eng._h
        s = SAMPLERS[sampler] if isinstance(sampler, str) else int(sampler)
        W, ew = self.world, self.A + 1
        total_ms = 0.0
        import time
        tr = self.trace
        for _ in range(n):
            ec = np.zeros(W, np.int64)
            rc = np.zeros(W, np.int64)
            t0 = time.perf_counter()
            _check(L.dbl_sweep_begin(h, s, _p(ec, _lib.i64p), _p(rc, _lib.i64p)), "sweep_begin", h)
            t1 = time.perf_counter()
            send_ent = torch.empty(int(ec.sum()) * ew, dtype=torch.int32, device=self.device)
            send_rec = torch.empty(int(rc.sum()) * 3, dtype=torch.int32, device=self.device)
            _check(L.dbl_exchange_pack(h, send_ent.data_ptr() if send_ent.numel() else None,
                                       send_rec.data_ptr() if send_rec.numel() else None), "exchange_pack", h)
            t2 = time.perf_counter()
            recv_ent, ne, recv_rec, nr = exchange(dist, send_ent, ec, send_rec, rc, ew, self.device, torch)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            _check(L.dbl_exchange_unpack(h, recv_ent.data_ptr() if ne else None, ne,
                                         recv_rec.data_ptr() if nr else None, nr), "exchange_unpack", h)
            _check(L.dbl_sweep_end(h), "sweep_end", h)
            t4 = time.perf_counter()
            total_ms += self.eng.last_sweep_ms()
            self._sync_summary()
            t5 = time.perf_counter()
            for k, v in (("begin", t1 - t0), ("pack", t2 - t1), ("exchange", t3 - t2), ("unpack_end", t4 - t3),
                         ("summary", t5 - t4), ("sweeps", 1e-3)):
                tr[k] = tr.get(k, 0.0) + v * 1e3
            self.last_exchange = (int(ec.sum()), int(rc.sum()))
        self._ms = total_ms

    def last_sweep_ms(self):
        return self._ms

    # ---- read-out ------------------------------------------------------------------------------------
    def summary(self):
        return self.eng.summary()

    def kernel_launches(self):
        return self.eng.kernel_launches()

    def link_kernel_ms(self):
        return self.eng.link_kernel_ms()

    @property
    def iteration(self):
        return self.eng.iteration

    def owned_masks(self):
        e = self.eng
        em = np.zeros(e.num_entities, np.uint8)
        rm = np.zeros(e.num_records, np.uint8)
        _check(_lib.load().dbl_owned_masks(e._h, _p(em, _lib.u8p), _p(rm, _lib.u8p)), "owned_masks", e._h)
        return em.astype(bool), rm.astype(bool)

    def download_state(self, out=None):
        """The full state on every rank: each rank exports the rows it owns into device buffers (zeros elsewhere),
        one all-reduce (NCCL) per array sums them, one device-to-host copy brings them back.  `out` may hold
        preallocated (e.g. pinned) host arrays under the keys z, link, y, block; they are filled in place."""
        torch, dist, e = self.torch, self.dist, self.eng
        R, E, A = e.num_records, e.num_entities, self.A
        y = torch.empty(E * A, dtype=torch.int32, device=self.device)
        blk = torch.empty(E, dtype=torch.int32, device=self.device)
        link = torch.empty(R, dtype=torch.int32, device=self.device)
        z = torch.empty(R * A, dtype=torch.uint8, device=self.device)
        _check(_lib.load().dbl_export_owned_dev(e._h, y.data_ptr(), blk.data_ptr(), link.data_ptr(), z.data_ptr()),
               "export_owned", e._h)
        for t in (y, blk, link, z):
            dist.all_reduce(t)  # exactly one rank owns each row
        theta = np.zeros((A, self.F))
        _check(_lib.load().dbl_summary(e._h, None, None, None, _p(theta, _lib.f64p)), "summary", e._h)
        res = {"theta": theta}
        for k, dev, shape in (("y", y, (E, A)), ("block", blk, (E,)), ("link", link, (R,)), ("z", z, (R, A))):
            host = out.get(k) if out else None
            if host is None:
                res[k] = dev.cpu().numpy().reshape(shape)
            else:
                torch.from_numpy(host.reshape(-1)).copy_(dev, non_blocking=True)
                res[k] = host
        torch.cuda.current_stream().synchronize()
        if out and out.get("theta") is not None:
            out["theta"][...] = theta
            res["theta"] = out["theta"]
        return res
