"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch on the box, gloo in the CPU tests) for the three real exchange steps of
the path:

  * Krylov vector all-gather  -- rank r owns a contiguous row block of the
    projected H (global column ids) and the matching slice of every vector;
  * PT2 dedup                 -- the candidate space is partitioned by key hash, every
    unique candidate is summed on exactly one rank (pt2_select_sharded); the explicit
    exchange (exchange_by_owner: owner partition + all_to_all of (determinant, partial
    coupling) pairs) is kept as a utility for workflows whose sources cannot be replicated;
  * global top-k merge        -- all-gather of the per-rank top-k, identical
    deterministic merge on every rank.

The projected-H build itself needs no communication (the basis index and the
integral tables are replicated).  Everything here is tensor plumbing on top of
callables, so the exchange logic runs unchanged on CPU tensors under gloo.
The reference has no distributed code at all.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def row_block(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """contiguous equal row blocks (the last ranks may get one row less / be short)."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def allgather_vector(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """concatenate the per-rank slices (row_block layout) into the full length-n vector."""
    rank, ws = world()
    if ws == 1:
        return local
    per = -(-n // ws)
    if local.shape[0] != per:
        pad = torch.zeros(per - local.shape[0], dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad])
    if local.is_complex():
        out = torch.empty(per * ws, 2, dtype=torch.float64, device=local.device)
        dist.all_gather_into_tensor(out, torch.view_as_real(local.contiguous()), group=group)
        return torch.view_as_complex(out)[:n]
    out = torch.empty(per * ws, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n]


class ShardedOperator:
    """Row-block sharded operator: matvec(x_full) -> y_full (all-gathered).
    `local_matvec(x_full) -> y_local` is the rank's CSR block product (fgk_spmv_*)."""

    def __init__(self, n: int, local_matvec: Callable, local_diagonal: Optional[torch.Tensor] = None):
        self.n = n
        self.rank, self.world = world()
        self.row_begin, self.row_end = row_block(n, self.rank, self.world)
        self._mv = local_matvec
        self._diag_local = local_diagonal
        self._diag = None

    def matvec_local(self, x_full):
        return self._mv(x_full)

    def matvec(self, x_full):
        return allgather_vector(self._mv(x_full), self.n)

    def diagonal(self):
        if self._diag is None:
            self._diag = allgather_vector(self._diag_local, self.n)
        return self._diag


def owner_of(dets: torch.Tensor, world_size: int) -> torch.Tensor:
    """owner rank of every packed determinant: a multiplicative hash of the two words
    (int64 arithmetic wraps, which is what we want)."""
    a, b = dets[:, 0], dets[:, 1]
    h = (a * -7046029254386353131) ^ (b * -4417276706812531889)     # odd 64-bit constants
    h = h ^ ((h >> 31) & 0x1FFFFFFFF)
    h = h * -7723592293110705685
    return ((h >> 33) & 0x3FFFFFFF) % world_size


def exchange_by_owner(dets: torch.Tensor, vals: torch.Tensor, group=None):
    """route (determinant, value) pairs to their owner ranks; returns what this rank
    received (duplicates across senders still to be reduced by the caller)."""
    rank, ws = world()
    if ws == 1:
        return dets, vals
    if dets.is_cuda:
        # count + scatter kernels (fgk_partition_by_owner) instead of a 64-bit sort
        from . import _native as nat
        L, dev_i, st = nat.lib(), nat.device_index(dets.device), nat.stream_ptr(dets.device)
        dets, vals = dets.contiguous(), vals.to(torch.float64).contiguous()
        m = dets.shape[0]
        send_counts = torch.zeros(ws, dtype=torch.int64, device=dets.device)
        nat.check(L.fgk_partition_by_owner(nat.ptr(dets), None, m, ws, nat.ptr(send_counts), None, None,
                                           0, dev_i, st))
        cursors = torch.zeros(ws, dtype=torch.int64, device=dets.device)
        torch.cumsum(send_counts[:-1], 0, out=cursors[1:])
        sd, sv = torch.empty_like(dets), torch.empty_like(vals)
        nat.check(L.fgk_partition_by_owner(nat.ptr(dets), nat.ptr(vals), m, ws, nat.ptr(cursors),
                                           nat.ptr(sd), nat.ptr(sv), 1, dev_i, st))
        dets, vals = sd, sv
    else:
        own = owner_of(dets, ws)
        order = torch.argsort(own, stable=True)
        dets, vals, own = dets[order].contiguous(), vals[order].contiguous(), own[order]
        send_counts = torch.bincount(own, minlength=ws).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    total = int(sum(rc))
    rdets = torch.empty(total, 2, dtype=dets.dtype, device=dets.device)
    rvals = torch.empty(total, dtype=vals.dtype, device=vals.device)
    dist.all_to_all_single(rdets, dets, output_split_sizes=rc, input_split_sizes=sc, group=group)
    dist.all_to_all_single(rvals, vals, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return rdets, rvals


def merge_topk(dets: torch.Tensor, score: torch.Tensor, k: int, n_orb: int, group=None):
    """global top-k from per-rank candidates: all-gather every rank's local top-k
    (padded to k) and run the same deterministic selection everywhere."""
    from .expansion import select_top_k
    rank, ws = world()
    ld, ls = select_top_k(dets, score, k, n_orb)
    if ws == 1:
        return ld, ls
    m = ld.shape[0]
    # one collective: rows {alpha, beta, score bits}; padding rows carry -inf
    pk = torch.zeros(k, 3, dtype=torch.int64, device=dets.device)
    ps = torch.full((k,), float("-inf"), dtype=torch.float64, device=dets.device)
    ps[:m] = ls.to(torch.float64)
    pk[:m, :2] = ld
    pk[:, 2] = ps.view(torch.int64)
    gk = torch.empty(ws * k, 3, dtype=torch.int64, device=dets.device)
    dist.all_gather_into_tensor(gk, pk, group=group)
    gs = gk[:, 2].contiguous().view(torch.float64)
    live = gs > float("-inf")
    return select_top_k(gk[live][:, :2].contiguous(), gs[live], k, n_orb)


def allreduce_scalar(x: float, op="sum", device="cpu") -> float:
    rank, ws = world()
    if ws == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return float(t[0])


# ---- engine-facing helpers (CUDA) ----------------------------------------------------------
def build_sharded_h(ham, dets, mode, index=None, sort_rows=False, operator=False):
    """each rank builds the rows of its block; returns (ProjectedH block, ShardedOperator).
    operator=False: CSR rows (reference-comparable form); operator=True: the best storage for
    repeated H.v (ham.projected_operator: packed SELL-32 built directly when possible)."""
    from .hamiltonian import BasisIndex
    rank, ws = world()
    idx = index if index is not None else BasisIndex(dets)
    lo, hi = row_block(dets.shape[0], rank, ws)
    if operator:
        P = ham.projected_operator(dets, mode, row_begin=lo, row_end=hi, index=idx, packed=True, min_rows=1024)
        if not getattr(P, "sell_only", False):
            P.optimize_for_matvec(min_rows=0)
    else:
        P = ham.projected_csr(dets, mode, row_begin=lo, row_end=hi, index=idx, packed=True,
                              sort_rows=sort_rows)
    op = ShardedOperator(dets.shape[0], P.matvec, P.diagonal())
    return P, op


def pt2_select_sharded(ham, index, coeffs, energy, k, mode=None, workspace=None, coeff_cut=1e-8,
                       max_passes=4096):
    """Stage-3 selection on N GPUs by partitioning the CANDIDATE space: rank r owns the
    candidates whose key hash falls into its buckets; every rank walks all significant sources
    but accumulates only what it owns (the bucket test sits right after the key hash, before
    any table traffic), so dedup is local and nothing but the final k x 24 B top-k lists
    crosses NVLink.  The sweep is bound by random DRAM sectors of the accumulator table
    (0.5 ns per owned candidate) while a skipped candidate costs ~0.01 ns of integer work --
    measured alternative, sources sharded + all-to-all of partial sums to the owner + merge:
    2.67 s on 2 GPUs against 2.48 s on one (configs[4] shape, 4.4e9 candidates), because the
    owner-side merge repeats the random-access upsert for every exchanged pair.
    Returns (selected dets, scores, stats); identical on all ranks.  With one rank this is
    expansion.pt2_select."""
    from . import _native as nat
    from .expansion import (_raw_connections_per_det, default_pt2_workspace, planned_passes, pt2_select,
                            select_top_k)
    mode = nat.PT2_SUM if mode is None else mode
    rank, ws = world()
    if ws == 1:
        sel, sc, st = pt2_select(ham, index, coeffs, energy, k, workspace=workspace, mode=mode,
                                 coeff_cut=coeff_cut, max_passes=max_passes)
        st = dict(st)
        st["raw_candidates_total"] = st["raw_candidates"]
        st["unique_total"] = st["unique_candidates"]
        return sel, sc, st
    dev = ham.device
    c32 = coeffs.to(dev).to(torch.float32)
    src = torch.nonzero(c32.abs() > coeff_cut).squeeze(1)          # ALL significant sources
    cj = c32[src].double()
    n_src = int(src.numel())
    empty = (torch.empty(0, 2, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.float64, device=dev))
    if n_src == 0:
        return empty + (dict(n_sources=0, raw_candidates=0, passes=1, raw_candidates_total=0, unique_total=0),)
    wa = workspace if workspace is not None else default_pt2_workspace(ham, -(-n_src // ws))
    # every rank must run the same number of bucket passes: size them for the smallest workspace
    # (agreed once per workspace object: two small collectives less on every later sweep)
    shared = getattr(wa, "_shared_sizes", None)
    if shared is None or shared[0] != ws:
        t = torch.tensor([-float(wa.capacity), -float(wa.queue_pairs)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shared = wa._shared_sizes = (ws, int(-t[0]), int(-t[1]))
    _, cap, qp = shared
    raw_ub = n_src * _raw_connections_per_det(ham)
    local_passes = planned_passes(raw_ub, cap, ws)
    if qp:
        local_passes = max(local_passes, -(-raw_ub // (qp * ws)))
    while True:
        keep_d, keep_s, raw, uniq, over = [], [], 0, 0, False
        n_pass = ws * local_passes
        for p in range(local_passes):
            wa.reset()
            wa.accumulate(ham, index, src, cj, mode, n_pass, rank + ws * p)
            ns, nr, ov = wa.count()
            if ov:                      # this rank's pass overflowed: finish the loop cheaply, redo below
                over = True
                continue
            raw += nr
            d, sc, live = wa.select_head(ham if mode == nat.PT2_SUM else None, ns, energy, k)
            uniq += live
            sd, ss = select_top_k(d, sc, k, ham.n_orbitals)
            keep_d.append(sd.clone())
            keep_s.append(ss.clone())
            del d, sc
        # ONE collective per sweep: overflow flag (max) and the counters (sum) travel together
        t = torch.tensor([1.0 if over else 0.0, float(raw), float(uniq)], dtype=torch.float64, device=dev)
        tl = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(tl, t)
        tot = torch.stack(tl).cpu()
        if float(tot[:, 0].max()) == 0.0:
            break
        local_passes *= 2
        if ws * local_passes > max_passes:
            raise RuntimeError(f"PT2 candidate set does not fit the workspaces in {max_passes} passes")
    sel, sc = merge_topk(torch.cat(keep_d), torch.cat(keep_s), k, ham.n_orbitals)
    st = dict(n_sources=n_src, raw_candidates=raw, passes=local_passes, unique_local=uniq)
    st["raw_candidates_total"] = int(tot[:, 1].sum())
    st["unique_total"] = int(tot[:, 2].sum())
    return sel, sc, st


# ---- H.v with the all-gather fused into the kernel (peer memory over NVLink) --------------------
class _DevArray:
    """view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


class FusedShardedOperator:
    """Row-block sharded H whose product, all-gather and barrier are ONE kernel launch:
    fgk_peer_step stores each y_r straight into every rank's next-vector buffer (peer-mapped
    memory over NVLink) while it streams the matrix, and its last CTA closes the step with a flag
    barrier.  Buffers ping-pong.  Real FP64 and complex128 vectors; FP64 SELL-32 storage or, when
    the operator has it, the packed exact-float32 SELL-32 storage (8 B/nnz).

    gather_local / matvec_local serve ROW-SHARDED Krylov vectors (solvers.lowest_eigenpairs with
    a sharded operator): the rank's slice of x is distributed by peer stores (fgk_peer_gather) and
    only the local rows of y are produced -- nothing full-length is computed twice.
    matvec_host uploads only this rank's slice of a host vector."""

    ERR_POLL = 64          # steps between looks at the barrier's error flag
    AR_SLOT = 4096         # doubles per rank in one all-reduce (fgk_peer_allreduce_sum)

    def __init__(self, P, group=None, storage="auto"):
        import ctypes as C
        from . import _native as nat
        if storage not in ("auto", "sell", "packed"):
            raise ValueError("storage must be 'auto', 'sell' or 'packed'")
        if storage == "packed":
            P.to_sell_packed()
        if storage == "sell" or (storage == "auto" and getattr(P, "_sellf", None) is None):
            P.to_sell()
        self.packed = storage == "packed" or (storage == "auto" and getattr(P, "_sellf", None) is not None)
        self.P = P
        self.n = P.n
        self.rank, self.world = world()
        self.dev = nat.device_index(P.device)
        self.device = P.device
        self.row_begin, self.row_end = P.row_begin, P.row_end
        if P.n_rows < 1:
            raise ValueError("FusedShardedOperator: every rank needs at least one row")
        L = nat.lib()
        nbytes = 16 * self.n                    # room for a complex128 vector
        own, handles = [], []
        sizes = [nbytes, nbytes, 8 * 64, 8 * 2 * self.world * self.AR_SLOT]
        for b in range(4):                     # two vector buffers, the flag array, the all-reduce scratch
            ptr, h = C.c_void_p(), C.create_string_buffer(64)
            nat.check(L.fgk_peer_alloc(sizes[b], self.dev, C.byref(ptr), h))
            own.append(ptr.value)
            handles.append(h.raw)
        self._own = own
        allh = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(allh, handles, group=group)
        else:
            allh = [handles]
        self._opened = []
        ptrs = [[0] * self.world for _ in range(4)]
        for p in range(self.world):
            for b in range(4):
                if p == self.rank:
                    ptrs[b][p] = own[b]
                else:
                    q = C.c_void_p()
                    nat.check(L.fgk_peer_open(allh[p][b], self.dev, C.byref(q)))
                    ptrs[b][p] = q.value
                    self._opened.append(q.value)
        VP = C.c_void_p * self.world
        self._bufs = [VP(*ptrs[0]), VP(*ptrs[1])]
        self._flags = VP(*ptrs[2])
        self._scratch = VP(*ptrs[3])
        self._ar_calls = 0
        self._views = [torch.as_tensor(_DevArray(own[b], self.n), device=P.device) for b in range(2)]
        self._zviews = [torch.as_tensor(_DevArray(own[b], self.n, "<c16"), device=P.device) for b in range(2)]
        self._err = torch.zeros(1, dtype=torch.int64, device=P.device)
        self._done = torch.zeros(1, dtype=torch.int32, device=P.device)
        self._epoch = 0
        self._cur = 0
        self._cplx = False
        self._diag = None
        self._failed = False
        self._host_io = None
        if self.world > 1:
            dist.barrier(group=group)          # every rank has mapped every buffer

    def close(self):
        from . import _native as nat
        L = nat.lib()
        torch.cuda.synchronize()
        if self.world > 1 and not self._failed:     # after a lost peer the others may never arrive
            dist.barrier()
        for q in getattr(self, "_opened", []):
            L.fgk_peer_close(q, self.dev)
        self._opened = []
        self._views, self._zviews = [], []
        for q in getattr(self, "_own", []):
            L.fgk_peer_free(q, self.dev)
        self._own = []

    def current(self):
        """the full current vector on this rank (a view: valid until the step after next)."""
        return (self._zviews if self._cplx else self._views)[self._cur]

    def load(self, x):
        self._cplx = x.is_complex()
        self.current().copy_(x)

    def _sync_args(self):
        from . import _native as nat
        self._epoch += 1
        if self._epoch % self.ERR_POLL == 0:
            self.check()
        return (self._flags, self.rank, self.world, self._epoch, nat.ptr(self._done), nat.ptr(self._err, torch.int64),
                self.dev, nat.stream_ptr(self.device))

    def step(self):
        """cur <- H cur on every rank (one launch); returns the new current vector (view)."""
        import ctypes as C
        from . import _native as nat
        L = nat.lib()
        nxt = 1 - self._cur
        flags = (nat.PEER_COMPLEX if self._cplx else 0) | (nat.PEER_PACKED_F32 if self.packed else 0)
        if self.packed:
            sp, pk, dg = self.P._sellf
            a_cols, a_vals, a_diag = nat.ptr(pk), None, nat.ptr(dg, torch.float64)
        else:
            sp, sc, sv = self.P._sell
            a_cols, a_vals, a_diag = nat.ptr(sc, torch.int32), nat.ptr(sv, torch.float64), None
        nat.check(L.fgk_peer_step(
            self.P.n_rows, nat.ptr(sp, torch.int64), a_cols, a_vals, a_diag, C.c_void_p(self._own[self._cur]),
            self._bufs[nxt], flags, self.row_begin, *self._sync_args()))
        self._cur = nxt
        return self.current()

    def gather_local(self, x_local):
        """distribute this rank's slice (rows row_begin:row_end) of a row-sharded vector into every
        rank's current buffer; afterwards current() is the full vector everywhere."""
        import ctypes as C
        from . import _native as nat
        self._cplx = x_local.is_complex()
        x_local = x_local.contiguous()
        if x_local.shape[0] != self.P.n_rows:
            raise ValueError("gather_local: expected this rank's row block")
        w = 16 if self._cplx else 8
        nat.check(nat.lib().fgk_peer_gather(
            C.c_void_p(x_local.data_ptr()), w * self.P.n_rows, self._bufs[self._cur], w * self.row_begin,
            *self._sync_args()))
        return self.current()

    def matvec_local(self, x_local, out=None):
        """y[rows of this rank] = (H x)[rows] for a row-sharded x: peer gather of the input + the
        plain local product (no output broadcast)."""
        xf = self.gather_local(x_local)
        y = self.P.matvec(xf, out=out, fmt="packed" if self.packed else "sell")
        # the next gather overwrites the buffer xf lives in: ping-pong so that a peer that is one
        # call ahead cannot touch the vector this rank's product is still reading
        self._cur = 1 - self._cur
        return y

    def allreduce_sum_(self, t):
        """in-place sum over the ranks of a small contiguous float64 tensor (the dot products of a
        row-sharded iteration): one launch over peer memory, identical bits on every rank"""
        import ctypes as C
        from . import _native as nat
        if self.world == 1:
            return t
        if t.dtype != torch.float64 or not t.is_contiguous() or t.numel() > self.AR_SLOT:
            dist.all_reduce(t)
            return t
        area = self._ar_calls & 1
        self._ar_calls += 1
        nat.check(nat.lib().fgk_peer_allreduce_sum(
            C.c_void_p(t.data_ptr()), t.numel(), 1, C.c_void_p(t.data_ptr()), self._scratch, self.AR_SLOT, area,
            *self._sync_args()))
        return t

    def reduce_partials(self, partial, rows, n):
        """sum of `rows` partial rows of n doubles (fgk_davidson_step's per-CTA rows, contiguous at the
        start of `partial`) over the rows and over the ranks, in ONE launch -> new (n,) tensor"""
        import ctypes as C
        from . import _native as nat
        out = torch.empty(n, dtype=torch.float64, device=partial.device)
        if self.world == 1 or n > self.AR_SLOT:
            out.copy_(partial.view(-1)[: rows * n].view(rows, n).sum(dim=0))
            return self.allreduce_sum_(out)
        area = self._ar_calls & 1
        self._ar_calls += 1
        nat.check(nat.lib().fgk_peer_allreduce_sum(
            C.c_void_p(partial.data_ptr()), n, rows, C.c_void_p(out.data_ptr()), self._scratch, self.AR_SLOT, area,
            *self._sync_args()))
        return out

    def check(self):
        e = int(self._err.item())
        if e:
            self._failed = True
            raise RuntimeError(f"fgk_peer barrier: a peer never arrived at epoch {e}")

    def matvec(self, x_full):
        self.load(x_full)
        return self.step().clone()

    def matvec_host(self, x_host, out=None):
        """host-buffer product on N GPUs: this rank uploads ONLY its slice of x (pinned host
        memory), the slices are exchanged over NVLink (fgk_peer_gather), one fused step runs, and
        the rank's rows of y come back in a pinned host tensor."""
        if not torch.is_tensor(x_host):
            x_host = torch.from_numpy(x_host)
        key = x_host.dtype
        if self._host_io is None or self._host_io[0] != key:
            ydt = torch.complex128 if x_host.is_complex() else torch.float64
            self._host_io = (key, torch.empty(self.P.n_rows, dtype=ydt, device=self.device),
                             torch.empty(self.P.n_rows, dtype=ydt).pin_memory())
        _, xl, yh = self._host_io
        dst = out if out is not None else yh
        xs = x_host[self.row_begin:self.row_end]
        if not (xs.is_contiguous() and dst.is_contiguous() and not dst.is_cuda):
            raise ValueError("matvec_host: contiguous host tensors expected")
        import ctypes as C
        from . import _native as nat
        self._cplx = x_host.is_complex()
        flags = (nat.PEER_COMPLEX if self._cplx else 0) | (nat.PEER_PACKED_F32 if self.packed else 0)
        if self.packed:
            sp, pk, dg = self.P._sellf
            a_cols, a_vals, a_diag = nat.ptr(pk), None, nat.ptr(dg, torch.float64)
        else:
            sp, sc, sv = self.P._sell
            a_cols, a_vals, a_diag = nat.ptr(sc, torch.int32), nat.ptr(sv, torch.float64), None
        nxt = 1 - self._cur
        self._epoch += 2                          # the gather and the step
        if (self._epoch // 2) % self.ERR_POLL == 0:
            self.check()
        # ONE library call: H2D of the slice, peer gather, fused step, D2H of the rows, synchronise
        nat.check(nat.lib().fgk_peer_matvec_host(
            self.P.n_rows, nat.ptr(sp, torch.int64), a_cols, a_vals, a_diag, C.c_void_p(xs.data_ptr()),
            C.c_void_p(xl.data_ptr()), self._bufs[self._cur], self._bufs[nxt], C.c_void_p(dst.data_ptr()), flags,
            self.row_begin, self._flags, self.rank, self.world, self._epoch - 1, nat.ptr(self._done),
            nat.ptr(self._err, torch.int64), self.dev, nat.stream_ptr(self.device)))
        self._cur = nxt
        return dst

    def diagonal(self):
        if self._diag is None:
            self._diag = allgather_vector(self.P.diagonal(), self.n)
        return self._diag
