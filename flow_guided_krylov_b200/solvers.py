"""Krylov drivers over the device SpMV (K6/K10).

* lowest_eigenpairs -- replaces np.linalg.eigh / scipy eigsh(which='SA') in
  residual_expansion.py:408-443, skqd.py:754-796, molecular.py:929-937.
  Small problems: dense torch.linalg.eigh on the device (cuSOLVER) -- the only
  dense step of the path.  Large problems: block Davidson with the diagonal
  preconditioner, H.v by fgk_spmv_f64.
* expm_multiply -- replaces scipy.sparse.linalg.expm_multiply(-i dt H, psi)
  (skqd.py:291-293): scaled, shifted Taylor series in complex128 with the
  Al-Mohy--Higham truncation test (tol 2^-53), H.v by fgk_spmv_z.
Vectors stay on the device; only scalars (norms, small projected matrices) are
looked at on the host.
"""
import math

import torch

DENSE_EIG_MAX = 3072
FUSED_DAVIDSON_MIN_ROWS = 16384     # above: fused iteration kernels (fgk_davidson_step) on one GPU too
MODEL_SPACE = 512                   # determinants (lowest diagonal) whose exact block seeds the fused Davidson


def _model_space_start(op, diag_full, nb):
    """Start vectors from the exact eigenvectors of the projected H inside the MODEL_SPACE
    determinants of lowest diagonal energy (a 512 x 512 dense problem, built by the row builder on
    that sub-basis: ~6 ms), embedded into the full space.  Unit vectors on the lowest diagonals --
    the plain start -- needed 78 products on a 14,400-determinant test problem, this start 59-63
    (CPU experiment with the same iteration; the preconditioner is unchanged).  None if the operator
    does not know its Hamiltonian / basis."""
    P = getattr(op, "P", None)
    ham, idx = getattr(P, "_ham", None), getattr(P, "_index", None)
    n = diag_full.shape[0]
    if ham is None or idx is None or n < 8 * MODEL_SPACE:
        return None, None
    sel = torch.argsort(diag_full)[:MODEL_SPACE]
    sub = idx.dets[sel].contiguous()
    D = ham.projected_csr(sub, P.mode & 3, packed=True).to_dense()
    _, U = torch.linalg.eigh(0.5 * (D + D.T))
    return sel, U[:, :nb].contiguous()



class _LocalOp:
    """one-GPU adapter with the contract of dist.FusedShardedOperator (row-sharded drivers)"""

    def __init__(self, P):
        self.P, self.n, self.row_begin, self.row_end, self.world, self.rank = P, P.n, 0, P.n, 1, 0

    def diagonal(self):
        return self.P.diagonal()

    def matvec_local(self, x, out=None):
        return self.P.matvec(x, out=out)

    def allreduce_sum_(self, t):
        return t

    def check(self):
        pass


def _full_matvec(P, x):
    if P.row_begin != 0 or P.row_end != P.n:
        raise ValueError("solver needs the full row range (use dist.ShardedH for row blocks)")
    return P.matvec(x)


class _Phases:
    """optional wall-clock split of the Davidson loop (synchronises; diagnostics only)."""

    def __init__(self, sink, dev):
        self.sink, self.dev, self.t = sink, dev, None

    def mark(self, name):
        if self.sink is None:
            return
        import time
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        if self.t is not None and name:
            self.sink[name] = self.sink.get(name, 0.0) + now - self.t
        self.t = now


def lowest_eigenpairs(P, k=1, tol=1e-11, max_iter=2000, max_space=None, matvec=None,
                      diagonal=None, dense_max=None, v0=None, phases=None, sharded=None):
    """k lowest eigenpairs of the symmetric operator P (a ProjectedH built with
    H_SYM, or any object with .n plus `matvec`/`diagonal` callables).
    Returns (w (k,) float64 tensor ascending, V (n,k)).

    Small n: dense eigh on the device.  Otherwise block Davidson with the diagonal
    preconditioner and THICK restart (the lowest Ritz vectors are kept, H V is rotated
    along, no extra products); the small projected matrix lives on the host (numpy eigh),
    so an iteration costs the H.v products plus a handful of n x m device GEMVs and one
    scalar read-back.

    sharded = a dist.FusedShardedOperator: the ROW-SHARDED form of the same iteration (see
    _davidson_sharded) -- every rank keeps only its row block of the basis vectors."""
    import numpy as np
    if dense_max is None:
        dense_max = DENSE_EIG_MAX                  # read at call time (tunable)
    if sharded is not None:
        fused = sharded.diagonal().is_cuda and (max_space is None or max_space <= 64)
        drv = _davidson_fused if fused else _davidson_sharded
        return drv(sharded, min(k, P.n), tol, max_iter, max_space, v0, phases)
    n = P.n
    mv = matvec if matvec is not None else (lambda x: _full_matvec(P, x))
    k = min(k, n)
    if n <= dense_max and matvec is None:
        D = P.to_dense()
        w, v = torch.linalg.eigh(0.5 * (D + D.T))
        return w[:k].clone(), v[:, :k].clone()
    diag = diagonal if diagonal is not None else P.diagonal()
    dev = diag.device
    if (matvec is None and dev.type == "cuda" and hasattr(P, "optimize_for_matvec") and n >= FUSED_DAVIDSON_MIN_ROWS
            and (max_space is None or max_space <= 64)):
        P.optimize_for_matvec()
        return _davidson_fused(_LocalOp(P), k, tol, max_iter, max_space, v0, phases)
    ph = _Phases(phases, dev)
    ph.mark(None)
    if matvec is None and hasattr(P, "optimize_for_matvec"):
        P.optimize_for_matvec()                    # many products ahead: SELL-32 / packed storage
    nb = min(max(2 * k, k + 2), n)                 # initial block: lowest diagonal entries
    m_max = min(n, max_space if max_space is not None else max(12 * k, 36))
    m_max = max(m_max, nb + k)
    keep = min(max(2 * k + 2, m_max // 3), m_max - k)
    # basis vectors are the ROWS of V / W (each one contiguous): the n x m products below are
    # then plain column-major GEMVs/GEMMs with leading dimension n, and H.v reads V[j] in place
    V = torch.zeros(m_max, n, dtype=torch.float64, device=dev)
    W = torch.zeros(m_max, n, dtype=torch.float64, device=dev)
    start = torch.argsort(diag)[:nb]
    V0 = torch.zeros(n, nb, dtype=torch.float64, device=dev)
    V0[start, torch.arange(nb, device=dev)] = 1.0
    # A seeded random admixture: unit vectors alone can be orthogonal to a whole
    # symmetry sector (e.g. triplets), which residual norms cannot detect.
    gen = torch.Generator(device=dev).manual_seed(20240229)
    V0 += 1e-2 * torch.randn(n, nb, dtype=torch.float64, generator=gen, device=dev)
    if v0 is not None:
        V0[:, 0] = v0.to(dev, torch.float64)
    V0, _ = torch.linalg.qr(V0)
    m = nb
    V[:m] = V0.T
    del V0
    for i in range(m):
        W[i] = mv(V[i])
    T = np.zeros((m_max, m_max))
    T[:m, :m] = (V[:m] @ W[:m].T).cpu().numpy()
    w_out = X = None
    ph.mark("setup")
    for _ in range(max_iter):
        Tm = 0.5 * (T[:m, :m] + T[:m, :m].T)
        th, s = np.linalg.eigh(Tm)
        s_dev = torch.from_numpy(np.ascontiguousarray(s.T)).to(dev)     # rows = Ritz vectors
        thk = torch.from_numpy(th[:k].copy()).to(dev)
        X = s_dev[:k] @ V[:m]                                           # (k, n)
        R = s_dev[:k] @ W[:m] - thk[:, None] * X
        rn = torch.linalg.norm(R, dim=1).cpu().numpy()
        ph.mark("ritz_residual")
        w_out = thk
        scale = max(1.0, float(np.abs(th[:k]).max()))
        if rn.max() < tol * scale:
            break
        todo = [i for i in range(k) if rn[i] >= tol * scale]
        if m + len(todo) > m_max:                   # thick restart: rotate V and H V together
            q = keep
            V[:q] = s_dev[:q] @ V[:m]
            W[:q] = s_dev[:q] @ W[:m]
            T[:] = 0.0
            T[np.arange(q), np.arange(q)] = th[:q]
            m = q
        added = 0
        norms = []
        for i in todo:
            den = thk[i] - diag
            den = torch.where(den.abs() < 1e-8, torch.full_like(den, -1e-8), den)
            t = R[i] / den
            for _ in range(2):                      # CGS2 against everything kept so far
                t = t - (V[:m + added] @ t) @ V[:m + added]
            nt = torch.linalg.norm(t)
            norms.append(nt)
            # normalised on the device; a vanished correction (nt ~ 0) is detected from the norm
            # that comes back with the projection below, and dropped then
            V[m + added] = t * torch.where(nt > 1e-10, 1.0 / nt, torch.zeros_like(nt))   # zero row if vanished
            added += 1
        ph.mark("correction_orth")
        for j in range(m, m + added):
            W[j] = mv(V[j])
        ph.mark("matvec")
        blk_dev = V[:m + added] @ W[m:m + added].T                  # new columns of T
        back = torch.cat([blk_dev.reshape(-1), torch.stack(norms)]).cpu().numpy()   # ONE read-back
        blk = back[:blk_dev.numel()].reshape(blk_dev.shape)
        ok = back[blk_dev.numel():] > 1e-10
        if not ok.all():                            # drop vanished corrections (rare: breakdown)
            good = [j for j in range(added) if ok[j]]
            if not good:
                break
            sel = torch.tensor([m + j for j in good], device=dev)
            V[m:m + len(good)] = V[sel]
            W[m:m + len(good)] = W[sel]
            rows = list(range(m)) + [m + j for j in good]
            blk = blk[np.ix_(rows, good)]
            added = len(good)
        T[:m + added, m:m + added] = blk
        T[m:m + added, :m + added] = blk.T
        m += added
        ph.mark("project")
    return w_out.clone(), X.T.contiguous()


def _davidson_fused(op, k, tol, max_iter, max_space, v0, phases):
    """The row-sharded block Davidson of _davidson_sharded with its vector algebra in four fused
    kernels per correction (fgk_davidson_step: Ritz residual + correction + first projection;
    orthogonalisation pass + second projection; normalised new vector; projection of H v) instead
    of ~45 small tensor kernels, and the dot products summed by one peer-memory all-reduce launch
    each.  Same iteration (start vectors, restart rule, lagged convergence test)."""
    import ctypes as C
    import numpy as np
    from . import _native as nat
    from .dist import allgather_vector
    L = nat.lib()
    n, lo, hi = op.n, op.row_begin, op.row_end
    nl = hi - lo
    diag_full = op.diagonal()
    diag = diag_full[lo:hi].contiguous()
    dev = diag.device
    dev_i = nat.device_index(dev)
    ph = _Phases(phases, dev)
    ph.mark(None)
    allsum = op.allreduce_sum_

    nb = min(max(2 * k, k + 2), n)
    m_max = min(n, max_space if max_space is not None else max(12 * k, 36))
    m_max = min(max(m_max, nb + k), 64)
    keep = min(max(2 * k + 2, m_max // 3), m_max - k)
    V = torch.zeros(m_max, nl, dtype=torch.float64, device=dev)
    W = torch.zeros(m_max, nl, dtype=torch.float64, device=dev)
    V0 = torch.zeros(n, nb, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev).manual_seed(20240229)       # same stream on every rank
    sel, U = _model_space_start(op, diag_full, nb)
    if sel is not None:                                  # exact eigenvectors of the model-space block
        V0[sel] = U
        V0 += 1e-3 * torch.randn(n, nb, dtype=torch.float64, generator=gen, device=dev)
    else:
        start = torch.argsort(diag_full)[:nb]
        V0[start, torch.arange(nb, device=dev)] = 1.0
        V0 += 1e-2 * torch.randn(n, nb, dtype=torch.float64, generator=gen, device=dev)
    if v0 is not None:
        V0[:, 0] = v0.to(dev, torch.float64)
    V0 = V0[lo:hi].contiguous()
    for _ in range(2):                                   # CholQR2 on the sharded block
        G = allsum((V0.T @ V0).contiguous())
        R = torch.linalg.cholesky(G, upper=True)
        V0 = torch.linalg.solve_triangular(R, V0, upper=True, left=False)
    m = nb
    V[:m] = V0.T
    del V0

    def apply_h(j0, j1):
        """W[j] = H V[j] for j0 <= j < j1.  Two real vectors ride in ONE complex product (x1 + i x2):
        the packed complex H.v costs 12 % more than the real one (2.93 vs 2.61 ms on configs[3]), so a
        pair costs 0.56 of two separate products."""
        j = j0
        while j + 1 < j1:
            y = op.matvec_local(torch.complex(V[j], V[j + 1]))
            W[j].copy_(y.real)
            W[j + 1].copy_(y.imag)
            j += 2
        if j < j1:
            op.matvec_local(V[j], out=W[j])

    apply_h(0, m)
    T = np.zeros((m_max, m_max))
    T[:m, :m] = allsum((V[:m] @ W[:m].T).contiguous()).cpu().numpy()

    props = torch.cuda.get_device_properties(dev)
    G_ = max(1, min(6 * props.multi_processor_count, -(-nl // 128)))      # CTAs of 4 warps, one 32-row tile per warp step
    partial = torch.empty(G_, m_max + 1, dtype=torch.float64, device=dev)
    tvec = torch.empty(nl, dtype=torch.float64, device=dev)
    st = nat.stream_ptr(dev)
    HostCoef = C.c_double * 64

    def step(mode, mm, host_coef=None, dev_coef=None, tt=None, theta=0.0, w=None, out=None, nrm=None):
        nat.check(L.fgk_davidson_step(
            mode, nl, nl, mm, nat.ptr(V), nat.ptr(W), host_coef, nat.ptr(dev_coef), nat.ptr(tt), float(theta),
            nat.ptr(diag), nat.ptr(tvec), nat.ptr(w), nat.ptr(out), nat.ptr(partial), G_, nat.ptr(nrm), dev_i, st))

    fused_reduce = getattr(op, "reduce_partials", None)

    def reduced(mm):
        """block partials -> one vector (rows added in order) -> sum over the ranks"""
        if fused_reduce is not None:
            return fused_reduce(partial, G_, mm + 1)
        return allsum(partial.view(-1)[: G_ * (mm + 1)].view(G_, mm + 1).sum(dim=0))

    w_out = X = None
    s_last = None
    ph.mark("setup")
    rn_prev = np.full(k, np.inf)
    for _ in range(max_iter):
        Tm = 0.5 * (T[:m, :m] + T[:m, :m].T)
        th, s = np.linalg.eigh(Tm)
        s_last, m_last = s, m
        w_out = th[:k].copy()
        scale = max(1.0, float(np.abs(th[:k]).max()))
        todo = [i for i in range(k) if rn_prev[i] >= tol * scale]
        if m + len(todo) > m_max:                   # thick restart: rotate V and H V together
            q = keep
            s_dev = torch.from_numpy(np.ascontiguousarray(s.T)).to(dev)
            V[:q] = s_dev[:q] @ V[:m]
            W[:q] = s_dev[:q] @ W[:m]
            T[:] = 0.0
            T[np.arange(q), np.arange(q)] = th[:q]
            s = np.eye(q)                           # the kept Ritz vectors are the new basis
            s_last, m_last = s, q
            m = q
        added = 0
        norms, rrs = [], []
        for i in todo:
            mm = m + added
            coef = np.zeros(64)
            coef[:m] = s[:m, i]                     # Ritz vector i in the (rotated) basis; new rows get 0
            step(0, mm, host_coef=HostCoef(*coef), theta=th[i])
            v1 = reduced(mm)
            rrs.append(v1[mm:mm + 1])
            step(1, mm, dev_coef=v1)
            v2 = reduced(mm)
            nrm = torch.empty(1, dtype=torch.float64, device=dev)
            step(2, mm, dev_coef=v2, tt=v2[mm:], out=V[mm], nrm=nrm)
            norms.append(nrm)
            added += 1
        ph.mark("correction_orth")
        apply_h(m, m + added)
        ph.mark("matvec")
        cols = []
        for j in range(m, m + added):
            step(3, m + added, w=W[j])
            cols.append(reduced(m + added)[: m + added])
        back = torch.cat(cols + norms + rrs).cpu().numpy()       # ONE read-back per iteration
        nbk = (m + added) * added
        blk = back[:nbk].reshape(added, m + added).T
        ok = back[nbk:nbk + added] > 1e-10
        rn_new = np.sqrt(np.maximum(back[nbk + added:], 0.0))
        for q_, i in enumerate(todo):
            rn_prev[i] = rn_new[q_]
        if rn_prev.max() < tol * scale:
            break                                   # the Ritz pairs of this iteration are converged
        if not ok.all():                            # drop vanished corrections (rare: breakdown)
            good = [j for j in range(added) if ok[j]]
            if not good:
                break
            sel = torch.tensor([m + j for j in good], device=dev)
            V[m:m + len(good)] = V[sel]
            W[m:m + len(good)] = W[sel]
            rows = list(range(m)) + [m + j for j in good]
            blk = blk[np.ix_(rows, good)]
            added = len(good)
        T[:m + added, m:m + added] = blk
        T[m:m + added, :m + added] = blk.T
        m += added
        ph.mark("project")
    op.check()
    s_dev = torch.from_numpy(np.ascontiguousarray(s_last[:m_last, :k].T)).to(dev)
    X = s_dev @ V[:m_last]
    Xf = torch.stack([allgather_vector(X[i].contiguous(), n) for i in range(X.shape[0])], dim=1) if op.world > 1 \
        else X.T.contiguous()
    return torch.from_numpy(w_out).to(dev), Xf.contiguous()


def _davidson_sharded(op, k, tol, max_iter, max_space, v0, phases):
    """Block Davidson with ROW-SHARDED vectors over a dist.FusedShardedOperator: rank r keeps rows
    [row_begin, row_end) of every basis vector V_j and of H V_j; a product gathers the new vector
    over peer memory (fgk_peer_gather) and multiplies the local row block; dot products are summed
    with one small all-reduce each.  The replicated form repeated all full-length algebra on
    every rank (0.30 parallel efficiency at 8 GPUs); here only O(m) numbers per
    iteration are replicated.  Same iteration as lowest_eigenpairs (same start vectors, same
    restart rule), so the Ritz values agree with it to rounding.  Returns (w, X) with X full-length
    (all-gathered) on every rank."""
    import numpy as np
    import torch.distributed as tdist
    from .dist import allgather_vector
    n, lo, hi = op.n, op.row_begin, op.row_end
    nl = hi - lo
    diag_full = op.diagonal()
    diag = diag_full[lo:hi]
    dev = diag.device
    ph = _Phases(phases, dev)
    ph.mark(None)
    multi = op.world > 1

    peer_sum = getattr(op, "allreduce_sum_", None)      # dist.FusedShardedOperator: one launch over peer memory

    def allsum(t):
        if multi:
            if peer_sum is not None:
                return peer_sum(t.contiguous())
            tdist.all_reduce(t)
        return t

    nb = min(max(2 * k, k + 2), n)
    m_max = min(n, max_space if max_space is not None else max(12 * k, 36))
    m_max = max(m_max, nb + k)
    keep = min(max(2 * k + 2, m_max // 3), m_max - k)
    V = torch.zeros(m_max, nl, dtype=torch.float64, device=dev)
    W = torch.zeros(m_max, nl, dtype=torch.float64, device=dev)
    start = torch.argsort(diag_full)[:nb]
    V0 = torch.zeros(n, nb, dtype=torch.float64, device=dev)
    V0[start, torch.arange(nb, device=dev)] = 1.0
    gen = torch.Generator(device=dev).manual_seed(20240229)       # same stream on every rank
    V0 += 1e-2 * torch.randn(n, nb, dtype=torch.float64, generator=gen, device=dev)
    if v0 is not None:
        V0[:, 0] = v0.to(dev, torch.float64)
    V0 = V0[lo:hi].contiguous()
    for _ in range(2):                                   # CholQR2 on the sharded block
        G = allsum(V0.T @ V0)
        R = torch.linalg.cholesky(G, upper=True)
        V0 = torch.linalg.solve_triangular(R, V0, upper=True, left=False)
    m = nb
    V[:m] = V0.T
    del V0
    for i in range(m):
        op.matvec_local(V[i], out=W[i])
    T = np.zeros((m_max, m_max))
    T[:m, :m] = allsum(V[:m] @ W[:m].T).cpu().numpy()
    w_out = X = None
    ph.mark("setup")
    # One host synchronisation and three small all-reduces per iteration: the residual norms of
    # the Ritz vectors travel with the first orthogonalisation pass and are LOOKED AT one
    # iteration late (together with the projection block), so a converged solve spends one extra
    # product instead of a read-back per iteration.
    rn_prev = np.full(k, np.inf)
    for _ in range(max_iter):
        Tm = 0.5 * (T[:m, :m] + T[:m, :m].T)
        th, s = np.linalg.eigh(Tm)
        s_dev = torch.from_numpy(np.ascontiguousarray(s.T)).to(dev)
        thk = torch.from_numpy(th[:k].copy()).to(dev)
        X = s_dev[:k] @ V[:m]
        R = s_dev[:k] @ W[:m] - thk[:, None] * X
        rr_local = (R * R).sum(dim=1)
        ph.mark("ritz_residual")
        w_out = thk
        scale = max(1.0, float(np.abs(th[:k]).max()))
        todo = [i for i in range(k) if rn_prev[i] >= tol * scale]
        if m + len(todo) > m_max:
            q = keep
            V[:q] = s_dev[:q] @ V[:m]
            W[:q] = s_dev[:q] @ W[:m]
            T[:] = 0.0
            T[np.arange(q), np.arange(q)] = th[:q]
            m = q
        added = 0
        norms = []
        rr_glob = None
        for i in todo:
            den = thk[i] - diag
            den = torch.where(den.abs() < 1e-8, torch.full_like(den, -1e-8), den)
            t = R[i] / den
            # pass 1 (+ all residual norms on the first correction), pass 2 (+ t.t: the norm after the
            # pass follows from Pythagoras, V is orthonormal)
            buf = allsum(torch.cat([V[:m + added] @ t, rr_local if rr_glob is None else rr_local[:0]]))
            if rr_glob is None:
                rr_glob = buf[m + added:]
            t = t - buf[:m + added] @ V[:m + added]
            buf = allsum(torch.cat([V[:m + added] @ t, (t * t).sum().reshape(1)]))
            c2 = buf[:m + added]
            t = t - c2 @ V[:m + added]
            nt = torch.sqrt(torch.clamp(buf[m + added] - (c2 * c2).sum(), min=0.0))
            norms.append(nt)
            V[m + added] = t * torch.where(nt > 1e-10, 1.0 / nt, torch.zeros_like(nt))
            added += 1
        ph.mark("correction_orth")
        for j in range(m, m + added):
            op.matvec_local(V[j], out=W[j])
        ph.mark("matvec")
        blk_dev = allsum(V[:m + added] @ W[m:m + added].T)
        back = torch.cat([blk_dev.reshape(-1), torch.stack(norms), rr_glob]).cpu().numpy()   # ONE read-back
        nb_ = blk_dev.numel()
        blk = back[:nb_].reshape(blk_dev.shape)
        ok = back[nb_:nb_ + added] > 1e-10
        rn_prev = np.sqrt(np.maximum(back[nb_ + added:], 0.0))
        if rn_prev.max() < tol * scale:
            break                                   # (w_out, X) of this iteration's Ritz step are converged
        if not ok.all():                            # drop vanished corrections (rare: breakdown)
            good = [j for j in range(added) if ok[j]]
            if not good:
                break
            sel = torch.tensor([m + j for j in good], device=dev)
            V[m:m + len(good)] = V[sel]
            W[m:m + len(good)] = W[sel]
            rows = list(range(m)) + [m + j for j in good]
            blk = blk[np.ix_(rows, good)]
            added = len(good)
        T[:m + added, m:m + added] = blk
        T[m:m + added, :m + added] = blk.T
        m += added
        ph.mark("project")
    op.check()
    Xf = torch.stack([allgather_vector(X[i].contiguous(), n) for i in range(X.shape[0])], dim=1)
    return w_out.clone(), Xf.contiguous()


# Al-Mohy & Higham 2011, table 3.1 (tol = 2^-53): theta_m for selected m
_THETA = {1: 2.29e-16, 2: 2.58e-8, 3: 1.39e-5, 4: 3.40e-4, 5: 2.40e-3, 6: 9.07e-3, 7: 2.38e-2,
          8: 5.00e-2, 9: 8.96e-2, 10: 1.44e-1, 11: 2.14e-1, 12: 3.00e-1, 13: 4.00e-1, 14: 5.14e-1,
          15: 6.41e-1, 16: 7.81e-1, 17: 9.31e-1, 18: 1.09, 19: 1.26, 20: 1.44, 21: 1.62, 22: 1.82,
          23: 2.01, 24: 2.22, 25: 2.43, 26: 2.64, 27: 2.86, 28: 3.08, 29: 3.31, 30: 3.54, 35: 4.7,
          40: 6.0, 45: 7.2, 50: 8.5, 55: 9.9}


def one_norm(P):
    """exact ||H||_1 = max column abs sum of a CSR operator (device scatter-add)."""
    cs = torch.zeros(P.n, dtype=torch.float64, device=P.vals.device)
    cs.index_add_(0, P.cols.long(), P.vals.abs())
    return cs


def complex_abs2(v):
    """|v_i|^2 of a complex (or real) tensor with plain real kernels"""
    if v.is_complex():
        return torch.view_as_real(v).pow(2).sum(dim=-1)
    return v * v


def spectral_radius_estimate(matvec, n, mu, device, iters=12, seed=20240301):
    """|lambda|_max of (H - mu I) by power iteration on a seeded random complex vector (the same
    on every rank).  Converges from below; expm_multiply adds a margin and checks the series a
    posteriori.  Costs `iters` products, once per operator."""
    gen = torch.Generator(device=device).manual_seed(seed)
    v = torch.randn(n, 2, dtype=torch.float64, generator=gen, device=device)
    v = torch.view_as_complex(v).contiguous()
    v = v / complex_abs2(v).sum().sqrt()
    est = None
    for _ in range(iters):
        w = matvec(v) - mu * v
        est = complex_abs2(w).sum().sqrt()
        v = w / est
    return float(est)


def _taylor_parameters(a, m_cap=55):
    """(m*, s) minimising m * ceil(a / theta_m) (Al-Mohy & Higham 2011, eq. 3.11 with alpha = a).
    m_cap bounds the degree: with alpha = the spectral radius the scaled argument theta_m is
    really reached, and the largest Taylor term is ~ e^theta (m <= 30: theta <= 3.54, a factor 34)."""
    if a == 0.0:
        return 1, 1
    best = None
    for m, th in _THETA.items():
        if m > m_cap:
            continue
        sm = max(1, math.ceil(a / th))
        if best is None or m * sm < best[0]:
            best = (m * sm, m, sm)
    return best[1], best[2]


def expm_multiply(P, psi, t, matvec=None, mu=None, norm1=None, allreduce_max=None, rho=None):
    """exp(t * H) psi for a real CSR operator H and complex scalar t (t = -i dt in
    SKQD, skqd.py:291).  Shift by mu = trace(H)/n, scale by s, Taylor degree m*
    chosen to minimise m * ceil(|t| alpha / theta_m); early exit when two successive terms fall
    under 2^-53 relative to the running sum.

    alpha = ||H - mu||_1 (norm1, exact column sums) by default.  With rho = an estimate of the
    spectral radius of H - mu (spectral_radius_estimate), alpha = 1.25 rho: for the nearly
    normal operators of this path the Taylor terms are governed by the spectral radius, which
    is several times smaller than the 1-norm (2,221 entries per row on configs[3]), so far fewer
    products are needed.  The choice is checked a posteriori: rounding errors of the series are
    bounded by u * sum_j ||term_j||; if that sum exceeds 1e3 ||result||, or the last of the m*
    terms is still above 1e-9 ||result||, the step is redone with the 1-norm parameters.

    On CUDA vectors one fused kernel per term (fgk_taylor_update_z) forms B <- c (H B - mu B),
    F <- F + B and both infinity norms."""
    mv = matvec if matvec is not None else (lambda x: _full_matvec(P, x))
    n = P.n
    psi = psi.to(torch.complex128)
    if mu is None:
        mu = float(P.diagonal().sum()) / n
    if norm1 is None and rho is None:
        d = P.diagonal()
        cs = one_norm(P) - d.abs() + (d - mu).abs()       # ||H - mu I||_1, exact
        norm1 = float(cs.max())
    tol = 2.0 ** -53
    fused = psi.is_cuda
    if fused:
        from . import _native as nat
        L = nat.lib()
        dev_i, st = nat.device_index(psi.device), nat.stream_ptr(psi.device)
        norms = torch.empty(2, dtype=torch.float64, device=psi.device)

    def inf_norm(v):
        # |z|^2 from the real view: torch's complex abs is a jiterator kernel (NVRTC compile on first
        # use in every process -- 70 ms inside a timed exp step on a box without a kernel cache)
        x = float(complex_abs2(v).max().sqrt()) if v.numel() else 0.0
        return allreduce_max(x) if allreduce_max is not None else x

    def run(alpha, guard):
        m_star, s = _taylor_parameters(abs(t) * alpha, 30 if guard else 55)
        eta = complex(math.e) ** (t * mu / s)
        F = psi.clone().contiguous()
        B = psi.clone().contiguous()
        for _ in range(s):
            c1 = inf_norm(B)
            total = c1
            nf = c1
            c2 = 0.0
            converged = False
            for j in range(1, m_star + 1):
                c = t / (s * j)
                if fused:
                    y = mv(B).contiguous()
                    nat.check(L.fgk_taylor_update_z(n, nat.ptr(y), nat.ptr(B), nat.ptr(F), float(mu),
                                                    float(c.real), float(c.imag), nat.ptr(norms, torch.float64),
                                                    dev_i, st))
                    c2, nf = norms.tolist()                 # the one read-back of this term
                else:
                    B = (mv(B) - mu * B) * c
                    c2 = inf_norm(B)
                    F = F + B
                    nf = inf_norm(F)
                total += c2
                if c1 + c2 <= tol * nf:
                    converged = True
                    break
                c1 = c2
            if guard and (total > 1e3 * nf or (not converged and c2 > 1e-9 * nf)):
                return None        # cancellation, or the series is not done after m* terms: the estimate was too small
            F = F * eta
            B = F.clone()
        return F

    if rho is not None:
        out = run(1.25 * rho, guard=norm1 is not None)
        if out is not None:
            return out
    return run(norm1, guard=False)
