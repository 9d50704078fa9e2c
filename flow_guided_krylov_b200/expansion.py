"""Stage-3 residual / PT2 basis expansion over the sm_100a engine.

Host-side mirror of reference src/krylov/residual_expansion.py:
  ResidualExpansionConfig (:27-57), SelectedCIExpander (:305-554),
  ResidualBasedExpander (:60-257) -- same constructor and method signatures, same
  stats dictionaries.  The Python loops over get_connections + dict lookups are
  replaced by fgk_pt2_* (enumerate -> filter -> hash-accumulate -> diagonal ->
  importance) and the dense n x n diagonalisation by the device CSR + solvers.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from .hamiltonian import BasisIndex, sort_unique_dets
from .solvers import lowest_eigenpairs
from . import solvers as _solvers


@dataclass
class ResidualExpansionConfig:
    """residual_expansion.py:27-57 (field for field)."""
    max_configs_per_iter: int = 100
    residual_threshold: float = 1e-4
    max_iterations: int = 10
    energy_convergence: float = 1e-6
    min_energy_improvement_mha: float = 0.05
    stagnation_patience: int = 2
    max_basis_size: int = 4096
    use_importance_sampling: bool = True
    n_importance_samples: int = 10000


class Pt2Workspace:
    """Device hash map candidate -> FP64 accumulator (fgk_pt2_*)."""

    L2_REGION_BYTES = 24 << 20      # table region + its pool share that should stay L2-resident

    def __init__(self, capacity, device, queue_pairs=None):
        """capacity: distinct candidates the map can hold.  queue_pairs: size of the radix
        partition in front of the hash (pairs per accumulate call); None = no partition,
        'auto' = 2 x capacity when the accumulator is much larger than L2."""
        self.capacity = int(capacity)
        self.device = device
        slots = 1
        while slots < 2 * self.capacity:
            slots *= 2
        slots = max(slots, 1024)
        # caller-allocated (torch caching allocator): table, pool of 32-byte {key, sum} entries, counters
        self._table = torch.empty(slots, dtype=torch.int64, device=device)
        self._pool = torch.empty(self.capacity, 4, dtype=torch.int64, device=device)
        self._counters = torch.zeros(4, dtype=torch.int64, device=device)
        h = C.c_void_p()
        nat.check(nat.lib().fgk_pt2_create(
            self.capacity, slots, nat.ptr(self._table), nat.ptr(self._pool),
            nat.ptr(self._counters), nat.device_index(device), C.byref(h)))
        self._h = h
        self.queue_pairs = 0
        foot = 8 * slots + 32 * self.capacity
        if queue_pairs == "auto":
            queue_pairs = 2 * self.capacity if foot > 4 * self.L2_REGION_BYTES else None
        if queue_pairs:
            region_bits = 0
            while (foot >> region_bits) > self.L2_REGION_BYTES and (slots >> (region_bits + 1)) >= 4096:
                region_bits += 1
            queue_bits = max(region_bits, 10)
            nq = 1 << queue_bits
            stride = int(1.25 * int(queue_pairs) / nq) + 256
            self._qdets = torch.empty(nq * stride, 2, dtype=torch.int64, device=device)
            self._qvals = torch.empty(nq * stride, dtype=torch.float64, device=device)
            self._qcur = torch.zeros(nq, dtype=torch.int64, device=device)
            nat.check(nat.lib().fgk_pt2_set_partition(
                self._h, region_bits, queue_bits, stride, nat.ptr(self._qdets), nat.ptr(self._qvals),
                nat.ptr(self._qcur)))
            self.queue_pairs = int(queue_pairs)
            self.partition = dict(region_bits=region_bits, queue_bits=queue_bits, stride=stride)
        self.reset()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                nat.lib().fgk_pt2_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self):
        nat.check(nat.lib().fgk_pt2_reset(self._h, nat.stream_ptr(self.device)))

    def accumulate(self, ham, index, src_idx, coeff, mode=nat.PT2_SUM, n_pass=1, pass_id=0):
        ham._require_particle_numbers(index, "PT2 accumulate")
        src_idx = src_idx.to(torch.int64).contiguous()
        coeff = coeff.to(torch.float64).contiguous()
        nat.check(nat.lib().fgk_pt2_accumulate(
            ham._h, index._h, self._h, nat.ptr(src_idx, torch.int64), nat.ptr(coeff, torch.float64),
            src_idx.shape[0], mode, n_pass, pass_id, nat.stream_ptr(self.device)))

    def merge(self, dets, vals, mode=nat.PT2_SUM):
        dets = dets.contiguous()
        vals = vals.to(torch.float64).contiguous()
        nat.check(nat.lib().fgk_pt2_merge(self._h, nat.ptr(dets, torch.int64),
                                          nat.ptr(vals, torch.float64), dets.shape[0], mode,
                                          nat.stream_ptr(self.device)))

    def count(self):
        """-> (slots used, raw candidates tested, overflowed)  [synchronises]"""
        ns, nr, ov = C.c_int64(0), C.c_int64(0), C.c_int(0)
        rc = nat.lib().fgk_pt2_count(self._h, nat.stream_ptr(self.device), C.byref(ns), C.byref(nr),
                                     C.byref(ov))
        if rc not in (nat.OK, nat.ERR_CAPACITY):
            nat.check(rc)
        return ns.value, nr.value, bool(ov.value)

    def export(self, ham, n_slots, energy=0.0, want_coupling=True, want_diag=True):
        """-> (dets, coupling, diag, importance) of the live candidates (unordered);
        coupling / diag are None when not requested (saves 8 B per candidate each)."""
        dev = self.device
        dets = torch.empty(n_slots, 2, dtype=torch.int64, device=dev)
        cpl = torch.empty(n_slots, dtype=torch.float64, device=dev) if want_coupling else None
        dg = torch.empty(n_slots, dtype=torch.float64, device=dev) if want_diag else None
        imp = torch.empty(n_slots, dtype=torch.float64, device=dev)
        live = C.c_int64(0)
        nat.check(nat.lib().fgk_pt2_export(
            ham._h if ham is not None else None, self._h, n_slots, float(energy),
            nat.ptr(dets, torch.int64), nat.ptr(cpl), nat.ptr(dg),
            nat.ptr(imp, torch.float64), C.byref(live), nat.stream_ptr(dev)))
        m = live.value
        return (dets[:m], cpl[:m] if cpl is not None else None,
                dg[:m] if dg is not None else None, imp[:m])


def _ws_select_head(self, ham, n_slots, energy, k):
    """Candidates that can be among the k best of this sweep, scored in place
    (fgk_pt2_score + fgk_pt2_gather): -> (dets, scores, n_live).  ham=None scores by
    |coupling| (MAXABS sweeps), else by the PT2 importance.  The head holds every candidate
    down to one binary exponent below the k-th score, so select_top_k on it equals
    select_top_k on the full list."""
    dev = self.device
    if getattr(self, "_hist", None) is None:
        self._hist = torch.empty(2048, dtype=torch.int32, device=dev)
        self._thr = torch.empty(2, dtype=torch.int64, device=dev)
    live, keep = C.c_int64(0), C.c_int64(0)
    nat.check(nat.lib().fgk_pt2_score(
        ham._h if ham is not None else None, self._h, n_slots, float(energy), int(k),
        nat.ptr(self._hist), nat.ptr(self._thr), C.byref(live), C.byref(keep), nat.stream_ptr(dev)))
    m = keep.value
    dets = torch.empty(m, 2, dtype=torch.int64, device=dev)
    score = torch.empty(m, dtype=torch.float64, device=dev)
    wrote = C.c_int64(0)
    if m:
        nat.check(nat.lib().fgk_pt2_gather(self._h, n_slots, nat.ptr(self._thr), nat.ptr(dets, torch.int64),
                                           nat.ptr(score, torch.float64), m, C.byref(wrote),
                                           nat.stream_ptr(dev)))
    return dets[:wrote.value], score[:wrote.value], live.value


Pt2Workspace.select_head = _ws_select_head


def _key_sort_order(dets, n_orb):
    """argsort of packed determinants by ascending unsigned (alpha, beta)."""
    a, b = dets[:, 0], dets[:, 1]
    if n_orb == 64:
        flip = torch.tensor(-2 ** 63, dtype=torch.int64, device=dets.device)
        a, b = a ^ flip, b ^ flip
    o = torch.argsort(b, stable=True)
    return o[torch.argsort(a[o], stable=True)]


def select_top_k(dets, score, k, n_orb, tie_eps=1e-9):
    """Deterministic top-k: score descending, ties by ascending key (the reference's
    torch.topk leaves ties unspecified, residual_expansion.py:552).  Scores within a
    relative `tie_eps` of the k-th score count as tied with it: FP64 atomics accumulate in
    arbitrary order, so exactly degenerate candidates (e.g. spin-flipped partners) differ in
    the last bits from run to run -- the cut must not depend on that noise."""
    n = dets.shape[0]
    k = min(int(k), n)
    if k == 0:
        return dets[:0], score[:0]
    kth = torch.topk(score, k).values[-1]
    band = tie_eps * kth.abs()
    sure = torch.nonzero(score > kth + band).squeeze(1)
    tie = torch.nonzero((score >= kth - band) & (score <= kth + band)).squeeze(1)
    need = k - sure.numel()
    if tie.numel() > need:
        tie = tie[_key_sort_order(dets[tie], n_orb)[:need]]
    pick = torch.cat([sure, tie])
    # final order: score desc, key asc
    ko = _key_sort_order(dets[pick], n_orb)
    pick = pick[ko]
    pick = pick[torch.argsort(score[pick], descending=True, stable=True)]
    return dets[pick], score[pick]


def pt2_candidates(ham, index, coeffs, energy, workspace=None, mode=nat.PT2_SUM,
                   coeff_cut=1e-8, max_passes=4096, src_shard=None):
    """Phase 1+2 of _find_important_configs (residual_expansion.py:481-548) for the
    sources of `index` (src_shard=(rank, world): only every world-th significant source,
    the multi-GPU split -- the significant sources, not the rows, are what must balance).
    Returns (cand_dets, coupling, diag, importance, stats).

    A pass needs one pool entry per DISTINCT candidate (the table slot is won before a pool
    entry is claimed); an undersized workspace just takes more bucket passes.  Coupling sums are
    exact 128-bit fixed-point accumulations: bit-identical from run to run, for any number of
    passes and any number of owner ranks."""
    dev = ham.device
    n = len(index)
    c32 = coeffs.to(dev).to(torch.float32)                         # :481
    src = torch.nonzero(c32.abs() > coeff_cut).squeeze(1)          # :489-490
    if src_shard is not None:
        src = src[src_shard[0]::src_shard[1]]
    cj = c32[src].double()
    stats = dict(n_sources=int(src.numel()), raw_candidates=0, passes=1)
    empty = (torch.empty(0, 2, dtype=torch.int64, device=dev),) + tuple(
        torch.empty(0, dtype=torch.float64, device=dev) for _ in range(3))
    if src.numel() == 0:
        return empty + (stats,)
    ws = workspace
    if ws is None:
        ws = default_pt2_workspace(ham, int(src.numel()))
    n_pass = 1
    if ws.queue_pairs:
        n_pass = max(1, -(-int(src.numel()) * _raw_connections_per_det(ham) // ws.queue_pairs))
    while True:
        outs, raw, ok = [], 0, True
        for p in range(n_pass):
            ws.reset()
            ws.accumulate(ham, index, src, cj, mode, n_pass, p)
            ns, nr, ov = ws.count()
            if ov:
                ok = False
                break
            raw += nr
            outs.append(ws.export(ham, ns, energy))
        if ok:
            break
        n_pass *= 2
        if n_pass > max_passes:
            raise RuntimeError("PT2 candidate set does not fit the workspace even in "
                               f"{max_passes} passes (capacity {ws.capacity})")
    stats.update(raw_candidates=raw, passes=n_pass)
    if len(outs) == 1:
        return outs[0] + (stats,)
    return tuple(torch.cat([o[i] for o in outs]) for i in range(4)) + (stats,)


DISTINCT_PER_RAW = 0.75     # planning figure: distinct candidates per raw connection (measured 0.37 on
                            # the configs[3] sweep, 0.60 at configs[4] shape); a wrong guess costs a repeated sweep


def default_pt2_capacity(ham, n_sources, partition=False):
    """distinct-candidate capacity for a sweep over n_sources determinants: every raw connection
    up to 2^28 of them (small and medium sweeps never need a second pass), DISTINCT_PER_RAW of
    them above; bounded by the free HBM (a sweep that does not fit is split into bucket passes by
    the callers).  Per unit of capacity: 16 B table + 32 B pool + 24 B head buffers, plus 2 x 30 B
    of partition queue."""
    n_conn = _raw_connections_per_det(ham)
    raw = n_sources * n_conn
    want = raw if raw <= (1 << 28) else max(1 << 28, int(DISTINCT_PER_RAW * raw))
    free = nat.device_info(ham.device)["free_bytes"]
    per = 72 + (60 if partition else 0) + 8
    return int(min(max(4096, 1.05 * want), 0.8 * free / per, 2 ** 32 - 8))


def planned_passes(raw_upper_bound, capacity, shares=1):
    """bucket passes (per rank) so that the expected distinct candidates of a pass fit the workspace"""
    return max(1, -(-int(DISTINCT_PER_RAW * raw_upper_bound) // (int(capacity) * shares)))


def default_pt2_workspace(ham, n_sources, partition=False):
    """partition=True puts the radix partition (fgk_pt2_set_partition) in front of the hash.
    Measured on B200 (profiles/README.md): enumerate + append runs at 1.8e10 candidates/s, but the
    queue-ordered fold is still bound by compulsory misses into the (sparse, over-provisioned)
    tag table, so the two-phase sweep is no faster than the direct one (18 vs 13.4 ms on the
    bench sweep); it stays available, off by default."""
    cap = default_pt2_capacity(ham, n_sources, partition)
    return Pt2Workspace(cap, ham.device, queue_pairs="auto" if partition else None)


def pt2_select(ham, index, coeffs, energy, k, workspace=None, mode=nat.PT2_SUM, coeff_cut=1e-8,
               max_passes=4096, src_shard=None):
    """Streaming form of the selection: enumerate -> filter -> accumulate -> importance ->
    top-k, keeping only k candidates per bucket pass, so candidate sets far beyond HBM
    (config 5: ~3e10 raw connections) are processed exactly in n_pass sweeps.
    Returns (selected dets, importances (or max |c.H|), stats)."""
    dev = ham.device
    c32 = coeffs.to(dev).to(torch.float32)
    src = torch.nonzero(c32.abs() > coeff_cut).squeeze(1)
    if src_shard is not None:
        src = src[src_shard[0]::src_shard[1]]
    cj = c32[src].double()
    stats = dict(n_sources=int(src.numel()), raw_candidates=0, passes=1, unique_candidates=0)
    if src.numel() == 0:
        return (torch.empty(0, 2, dtype=torch.int64, device=dev),
                torch.empty(0, dtype=torch.float64, device=dev), stats)
    ws = workspace if workspace is not None else default_pt2_workspace(ham, int(src.numel()))
    # first guess from the planning ratio; with a partition queue every raw candidate of a pass
    # must also fit the queue
    raw_ub = int(src.numel()) * _raw_connections_per_det(ham)
    n_pass = planned_passes(raw_ub, ws.capacity)
    if ws.queue_pairs:
        n_pass = max(n_pass, -(-raw_ub // ws.queue_pairs))
    while True:
        keep_d, keep_s, raw, uniq, ok = [], [], 0, 0, True
        for p in range(n_pass):
            ws.reset()
            ws.accumulate(ham, index, src, cj, mode, n_pass, p)
            ns, nr, ov = ws.count()
            if ov:
                ok = False
                break
            raw += nr
            d, sc, live = ws.select_head(ham if mode == nat.PT2_SUM else None, ns, energy, k)
            uniq += live
            sd, ss = select_top_k(d, sc, k, ham.n_orbitals)
            keep_d.append(sd.clone())
            keep_s.append(ss.clone())
            del d, sc
        if ok:
            break
        n_pass *= 2
        if n_pass > max_passes:
            raise RuntimeError(f"PT2 candidate set does not fit the workspace in {max_passes} passes")
    stats.update(raw_candidates=raw, passes=n_pass, unique_candidates=uniq)
    if n_pass == 1:
        return keep_d[0], keep_s[0], stats
    sd, ss = select_top_k(torch.cat(keep_d), torch.cat(keep_s), k, ham.n_orbitals)
    return sd, ss, stats


def _raw_connections_per_det(ham):
    n, na, nb = ham.n_orbitals, ham.n_alpha, ham.n_beta
    va, vb = n - na, n - nb
    c2 = lambda m: m * (m - 1) // 2
    return na * va + nb * vb + c2(na) * c2(va) + c2(nb) * c2(vb) + na * nb * va * vb


class SelectedCIExpander:
    """residual_expansion.py:305-554, same signatures."""

    def __init__(self, hamiltonian, config: ResidualExpansionConfig = None):
        self.hamiltonian = hamiltonian
        self.config = config or ResidualExpansionConfig()
        self.device = getattr(hamiltonian, "device", "cpu")
        self.last_stats = {}

    # Under torchrun (torch.distributed initialised, world > 1) the two heavy steps shard
    # themselves: rows of the projected H by rank + fused H.v (dist.FusedShardedOperator) for
    # bases above `sharded_min_rows`, and the PT2 sweep by candidate ownership
    # (dist.pt2_select_sharded).  Every rank ends up with identical results.
    sharded_min_rows = 65536

    @staticmethod
    def _world():
        import torch.distributed as tdist
        return tdist.get_world_size() if tdist.is_available() and tdist.is_initialized() else 1

    # :408-443 -- float64, symmetrised; dense eigh for small n, Davidson above
    def _diagonalize_packed(self, dets, warm=None, trusted=False):
        """-> (E0, v0, index).  The last result is kept: the pipeline's loop calls expand_basis
        with the basis the previous call returned, whose eigenpair was just computed (the
        reference recomputes it, residual_expansion.py:356).  warm = (old dets, old eigenvector):
        start vector for the iterative solver, embedded into the new basis.  trusted: the
        determinants are a checked basis plus excitations of it (particle numbers conserved)."""
        H = self.hamiltonian
        n = dets.shape[0]
        c = getattr(self, "_last_diag", None)
        if c is not None and c[0].shape == dets.shape and bool(torch.equal(c[0], dets)):
            return c[1], c[2], c[3]
        index = BasisIndex(dets)
        if trusted:
            index._particles_ok = (H.n_alpha, H.n_beta)
        v0 = None
        if warm is not None and n > _solvers.DENSE_EIG_MAX:
            pos = index.lookup(warm[0]).long()
            v0 = torch.zeros(n, dtype=torch.float64, device=dets.device)
            v0[pos[pos >= 0]] = warm[1][pos >= 0]
        if self._world() > 1 and n >= self.sharded_min_rows:
            from . import dist as fdist
            Pb, _ = fdist.build_sharded_h(H, dets, nat.H_SYM, index=index, operator=True)   # packed rows, built directly
            op = fdist.FusedShardedOperator(Pb)
            try:
                w, v = lowest_eigenpairs(op, k=1, sharded=op, v0=v0)   # row-sharded vectors, peer gather per product
                op.check()
            finally:
                op.close()
        else:
            P = H.projected_operator(dets, nat.H_SYM, index=index, packed=True)   # CSR below 16,384 rows, packed SELL-32 above
            w, v = lowest_eigenpairs(P, k=1, v0=v0)
        out = (float(w[0]), v[:, 0], index)
        self._last_diag = (dets,) + out
        return out

    def _diagonalize(self, basis: torch.Tensor) -> Tuple[float, np.ndarray]:
        E, v, _ = self._diagonalize_packed(self.hamiltonian.pack(basis))
        return E, v.cpu().numpy()

    # :451-554
    def _find_important_packed(self, dets, index, energy, v):
        H = self.hamiltonian
        # the workspace is kept between rounds (allocation + first-touch of a fresh table cost more
        # than a small sweep); every sharded rank sizes it for its share of the sources
        world = self._world()
        need = default_pt2_capacity(H, -(-int(dets.shape[0]) // world))
        ws = getattr(self, "_ws", None)
        if ws is None or ws.capacity < need or str(ws.device) != str(H.device):
            ws = self._ws = Pt2Workspace(int(need * 1.5) if need < (1 << 24) else need, H.device)
        if world > 1:
            from . import dist as fdist
            sel, imp, st = fdist.pt2_select_sharded(H, index, v, energy, self.config.max_configs_per_iter, workspace=ws)
        else:
            sel, imp, st = pt2_select(H, index, v, energy, self.config.max_configs_per_iter, workspace=ws)
        if ws.capacity > (1 << 24):             # large sweeps: give the memory back for the next H build
            self._ws = None
        self.last_stats = st
        return sel, imp

    def _find_important_configs(self, basis, energy, eigenvector):
        H = self.hamiltonian
        dets = H.pack(basis)
        v = torch.as_tensor(np.asarray(eigenvector), dtype=torch.float64, device=self.device)
        sel, imp = self._find_important_packed(dets, BasisIndex(dets), energy, v)
        if sel.shape[0] == 0:
            return (torch.empty(0, basis.shape[1], device=self.device),
                    torch.empty(0, device=self.device))
        return H.unpack(sel, basis.dtype), imp.to(H.output_dtype)

    # :334-406
    def expand_basis(self, current_basis: torch.Tensor) -> Tuple[torch.Tensor, Dict]:
        H = self.hamiltonian
        current_basis = current_basis.to(self.device)
        dets = H.pack(current_basis)
        energy, v, index = self._diagonalize_packed(dets)
        sel, _ = self._find_important_packed(dets, index, energy, v)
        if sel.shape[0] == 0:                                                   # :363-364
            return current_basis, {'configs_added': 0, 'energy': energy, 'initial_energy': energy,
                                   'final_energy': energy}
        expanded = sort_unique_dets(torch.cat([dets, sel], dim=0), H.n_orbitals)   # :367-368
        keep_old = getattr(self, "_last_diag", None)
        new_energy, _, _ = self._diagonalize_packed(expanded, warm=(dets, v), trusted=True)   # :371
        energy_improvement = energy - new_energy
        if energy_improvement < -1e-8:                                           # :378-393
            self._last_diag = keep_old          # the caller keeps the old basis
            return current_basis, {
                'initial_size': len(current_basis), 'final_size': len(current_basis),
                'configs_added': 0, 'initial_energy': energy, 'final_energy': energy,
                'energy_improvement': 0.0, 'energy_improvement_mha': 0.0,
                'variational_violation': True, 'rejected_energy': new_energy,
                'rejected_increase_mha': -energy_improvement * 1000}
        stats = {                                                                # :395-404
            'initial_size': len(current_basis), 'final_size': int(expanded.shape[0]),
            'configs_added': int(sel.shape[0]), 'initial_energy': energy,
            'final_energy': new_energy, 'energy_improvement': energy_improvement,
            'energy_improvement_mha': energy_improvement * 1000, 'variational_violation': False}
        return H.unpack(expanded, current_basis.dtype), stats


class ResidualBasedExpander:
    """residual_expansion.py:60-257, same signatures (max |c_j <x|H|j>| instead of PT2)."""

    def __init__(self, hamiltonian, config: ResidualExpansionConfig = None):
        self.hamiltonian = hamiltonian
        self.config = config or ResidualExpansionConfig()
        self.device = getattr(hamiltonian, "device", "cpu")

    # :162-172 -- np.linalg.eigh on the RAW (unsymmetrised) matrix reads the lower triangle
    def _diagonalize_packed(self, dets):
        H = self.hamiltonian
        P = H.projected_csr(dets, nat.H_RAW, packed=True, sort_rows=False)
        D = P.to_dense()
        L = torch.tril(D)
        S = L + torch.tril(D, -1).T
        w, v = torch.linalg.eigh(S)
        return float(w[0]), v[:, 0], P._index

    def _diagonalize(self, basis):
        E, v, _ = self._diagonalize_packed(self.hamiltonian.pack(basis))
        return E, v.cpu().numpy()

    # :174-253
    def _find_residual_packed(self, dets, index, v):
        H = self.hamiltonian
        cfg = self.config
        cand, res, _, _, _ = pt2_candidates(H, index, v, 0.0, mode=nat.PT2_MAXABS, coeff_cut=1e-10)
        if cand.shape[0] == 0:
            return cand, res
        res = res.float().double()            # the reference stores residuals in float32 (:228)
        m = res > cfg.residual_threshold      # :239
        cand, res = cand[m], res[m]
        return select_top_k(cand, res, cfg.max_configs_per_iter, H.n_orbitals)

    def expand_basis(self, current_basis, energy: Optional[float] = None,
                     eigenvector: Optional[np.ndarray] = None):
        H = self.hamiltonian
        cfg = self.config
        current_basis = current_basis.to(self.device)
        n_current = len(current_basis)
        dets = H.pack(current_basis)
        if energy is None or eigenvector is None:
            energy, v, index = self._diagonalize_packed(dets)
        else:
            v = torch.as_tensor(np.asarray(eigenvector), dtype=torch.float64, device=self.device)
            index = BasisIndex(dets)
        history = {'energies': [energy], 'basis_sizes': [n_current], 'configs_added': []}
        cur_e, cur_v = energy, v
        energy_change = None
        iteration = -1
        for iteration in range(cfg.max_iterations):                             # :117-148
            if dets.shape[0] >= cfg.max_basis_size:
                break
            sel, _ = self._find_residual_packed(dets, index, cur_v)
            if sel.shape[0] == 0:
                break
            dets = sort_unique_dets(torch.cat([dets, sel], dim=0), H.n_orbitals)
            new_e, new_v, index = self._diagonalize_packed(dets)
            history['energies'].append(new_e)
            history['basis_sizes'].append(int(dets.shape[0]))
            history['configs_added'].append(int(sel.shape[0]))
            energy_change = abs(new_e - cur_e)
            if energy_change < cfg.energy_convergence:
                break
            cur_e, cur_v = new_e, new_v
        stats = {
            'initial_basis_size': n_current, 'final_basis_size': int(dets.shape[0]),
            'configs_added_total': int(dets.shape[0]) - n_current, 'iterations': iteration + 1,
            'converged': (energy_change < cfg.energy_convergence) if energy_change is not None else True,
            'final_energy': cur_e, 'history': history}
        return H.unpack(dets, current_basis.dtype), stats
