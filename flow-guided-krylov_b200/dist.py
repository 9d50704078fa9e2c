"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch on the box, gloo in the CPU tests) for the three real exchange steps of
the path:

  * Krylov vector all-gather  -- rank r owns a contiguous row block of the
    projected H (global column ids) and the matching slice of every vector;
  * PT2 dedup exchange        -- candidates are routed to an owner rank by key hash
    (all_to_all of (determinant, partial coupling) pairs), so every unique
    candidate is summed on exactly one rank;
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
    pd = torch.zeros(k, 2, dtype=torch.int64, device=dets.device)
    ps = torch.full((k,), float("-inf"), dtype=torch.float64, device=dets.device)
    pd[:m], ps[:m] = ld, ls.to(torch.float64)
    gd = torch.empty(ws * k, 2, dtype=torch.int64, device=dets.device)
    gs = torch.empty(ws * k, dtype=torch.float64, device=dets.device)
    dist.all_gather_into_tensor(gd, pd, group=group)
    dist.all_gather_into_tensor(gs, ps, group=group)
    live = gs > float("-inf")
    return select_top_k(gd[live], gs[live], k, n_orb)


def allreduce_scalar(x: float, op="sum", device="cpu") -> float:
    rank, ws = world()
    if ws == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return float(t[0])


# ---- engine-facing helpers (CUDA) ----------------------------------------------------------
def build_sharded_h(ham, dets, mode, index=None, sort_rows=True):
    """each rank builds CSR rows of its block; returns (ProjectedH block, ShardedOperator)."""
    from .hamiltonian import BasisIndex
    rank, ws = world()
    idx = index if index is not None else BasisIndex(dets)
    lo, hi = row_block(dets.shape[0], rank, ws)
    P = ham.projected_csr(dets, mode, row_begin=lo, row_end=hi, index=idx, packed=True,
                          sort_rows=sort_rows)
    op = ShardedOperator(dets.shape[0], P.matvec, P.diagonal())
    return P, op


def pt2_select_sharded(ham, index, coeffs, energy, k, mode=None, workspace=None):
    """Stage-3 selection with the significant sources dealt round-robin to the ranks and
    the candidates owned by key hash.  Returns (selected dets, importances, stats); identical on all ranks."""
    from . import _native as nat
    from .expansion import Pt2Workspace, pt2_candidates
    mode = nat.PT2_SUM if mode is None else mode
    rank, ws = world()
    cand, cpl, _, imp, st = pt2_candidates(ham, index, coeffs, energy, mode=mode,
                                           src_shard=(rank, ws), workspace=workspace)
    if ws > 1:
        rd, rv = exchange_by_owner(cand, cpl)
        wsp = Pt2Workspace(max(1024, int(rd.shape[0]) + 16), ham.device)
        wsp.merge(rd, rv, mode)
        ns, _, ov = wsp.count()
        if ov:
            raise RuntimeError("pt2_select_sharded: merge workspace overflow")
        cand, cpl, _, imp = wsp.export(ham, ns, energy)
    score = imp if mode == nat.PT2_SUM else cpl
    sel, simp = merge_topk(cand, score, k, ham.n_orbitals)
    st = dict(st)
    st["raw_candidates_total"] = int(allreduce_scalar(float(st["raw_candidates"]), "sum", ham.device))
    st["unique_local"] = int(cand.shape[0])
    st["unique_total"] = int(allreduce_scalar(float(cand.shape[0]), "sum", ham.device))
    return sel, simp, st
