"""ctypes front-end of the CPU oracle (oracle/fgk_oracle.c) + numpy restatements
of the reference's host-side numerics.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under flow_guided_krylov_b200/
does: the product path fails loudly when its CUDA library is missing.

Parity pin: checked against tests/golden/*.npz, which were produced by running
the reference itself (tests/golden/make_golden.py).

All `file:line` citations are under /root/reference/src.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """gcc the C restatement into oracle/liboracle.so (git-ignored)."""
    src = os.path.join(_HERE, "fgk_oracle.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    subprocess.check_call(
        ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, dbl, ci = C.c_void_p, C.c_long, C.c_double, C.c_int
        L.orc_ham_create.restype = vp
        L.orc_ham_create.argtypes = [vp, vp, ci, dbl]
        L.orc_ham_destroy.argtypes = [vp]
        L.orc_num_threads.restype = ci
        L.orc_set_num_threads.argtypes = [ci]
        L.orc_sign_single.restype = ci
        L.orc_sign_single.argtypes = [vp, ci, ci]
        L.orc_sign_double.restype = ci
        L.orc_sign_double.argtypes = [vp, ci, ci, ci, ci]
        L.orc_diag.argtypes = [vp, vp, i64, vp]
        L.orc_connections.restype = i64
        L.orc_connections.argtypes = [vp, vp, vp, vp, i64]
        L.orc_connections_count.argtypes = [vp, vp, i64, vp]
        L.orc_connections_fill.argtypes = [vp, vp, i64, vp, vp, vp, vp]
        L.orc_offdiag_coo.restype = i64
        L.orc_offdiag_coo.argtypes = [vp, vp, i64, vp, vp, vp, i64]
        L.orc_offdiag_coo_kets.restype = i64
        L.orc_offdiag_coo_kets.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp, i64]
        L.orc_pt2_candidates.restype = i64
        L.orc_pt2_candidates.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp]
        L.orc_csr_matvec_f64.argtypes = [i64, vp, vp, vp, vp, vp]
        L.orc_csr_matvec_z.argtypes = [i64, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(cfgs):
    return np.ascontiguousarray(np.asarray(cfgs), dtype=np.uint8)


def sort_unique(cfgs):
    """torch.unique(dim=0) (residual_expansion.py:368, skqd.py:942): rows sorted
    lexicographically, i.e. ascending site-0-MSB key."""
    return np.unique(_u8(cfgs), axis=0)


def keys_of(cfgs):
    """python-int keys, site 0 = MSB (molecular.py:498-500)."""
    out = []
    for row in _u8(cfgs):
        k = 0
        for b in row:
            k = (k << 1) | int(b)
        out.append(k)
    return out


def top_k_protocol(cand, imp, nsel):
    """indices of the nsel selected candidates: everything clearly above the nsel-th importance,
    then the members of its relative-1e-9 tie band by ascending key; ordered by importance
    descending, key ascending (cand: (m, S) 0/1 rows, site 0 most significant)."""
    kth = np.sort(imp)[::-1][nsel - 1]
    band = 1e-9 * abs(kth)
    sure = np.nonzero(imp > kth + band)[0]
    tie = np.nonzero((imp >= kth - band) & (imp <= kth + band))[0]
    need = nsel - len(sure)
    if len(tie) > need:
        tie = tie[np.lexsort(tuple(cand[tie][:, ::-1].T))[:need]]
    pick = np.concatenate([sure, tie])
    pick = pick[np.lexsort(tuple(cand[pick][:, ::-1].T))]          # key ascending ...
    return pick[np.argsort(-imp[pick], kind="stable")]             # ... inside importance descending


class OracleHam:
    """MolecularHamiltonian (molecular.py:35-117) over float32 tables."""

    def __init__(self, h1, g, n_alpha, n_beta, e_nuc=0.0):
        self.h1 = np.ascontiguousarray(h1, dtype=np.float32)     # molecular.py:68
        self.g = np.ascontiguousarray(g, dtype=np.float32)       # molecular.py:69
        self.n_orb = int(self.h1.shape[0])
        self.S = 2 * self.n_orb
        self.n_alpha, self.n_beta, self.e_nuc = int(n_alpha), int(n_beta), float(e_nuc)
        self._h = lib().orc_ham_create(_p(self.h1), _p(self.g), self.n_orb, self.e_nuc)

    def __del__(self):
        try:
            lib().orc_ham_destroy(self._h)
        except Exception:
            pass

    # molecular.py:778-792
    def hf_state(self):
        c = np.zeros(self.S, np.uint8)
        c[: self.n_alpha] = 1
        c[self.n_orb: self.n_orb + self.n_beta] = 1
        return c

    def sign_single(self, cfg, p, q):
        cfg = _u8(cfg)
        return lib().orc_sign_single(_p(cfg), p, q)

    def sign_double(self, cfg, p, r, q, s):
        cfg = _u8(cfg)
        return lib().orc_sign_double(_p(cfg), p, r, q, s)

    # molecular.py:133-184 (FP64 restatement on the float32 tables)
    def diag(self, cfgs):
        cfgs = _u8(cfgs).reshape(-1, self.S)
        out = np.empty(len(cfgs), np.float64)
        lib().orc_diag(self._h, _p(cfgs), len(cfgs), _p(out))
        return out

    # molecular.py:194-327
    def connections(self, cfg):
        cfg = _u8(cfg)
        n = lib().orc_connections(self._h, _p(cfg), None, None, 0)
        oc = np.empty((n, self.S), np.uint8)
        oe = np.empty(n, np.float32)
        lib().orc_connections(self._h, _p(cfg), _p(oc), _p(oe), n)
        return oc, oe

    # molecular.py:329-377 / 518-578 (deterministic order: source ascending)
    def connections_batch(self, cfgs):
        cfgs = _u8(cfgs).reshape(-1, self.S)
        n = len(cfgs)
        cnt = np.zeros(n, np.int64)
        lib().orc_connections_count(self._h, _p(cfgs), n, _p(cnt))
        offs = np.zeros(n + 1, np.int64)
        np.cumsum(cnt, out=offs[1:])
        tot = int(offs[-1])
        oc = np.empty((tot, self.S), np.uint8)
        oe = np.empty(tot, np.float32)
        src = np.empty(tot, np.int64)
        lib().orc_connections_fill(self._h, _p(cfgs), n, _p(offs), _p(oc), _p(oe), _p(src))
        return oc, oe, src, offs

    # molecular.py:580-638
    def offdiag_coo(self, basis):
        basis = _u8(basis).reshape(-1, self.S)
        n = len(basis)
        tot = lib().orc_offdiag_coo(self._h, _p(basis), n, None, None, None, 0)
        r = np.empty(tot, np.int64)
        c = np.empty(tot, np.int64)
        v = np.empty(tot, np.float32)
        lib().orc_offdiag_coo(self._h, _p(basis), n, _p(r), _p(c), _p(v), tot)
        return r, c, v

    def offdiag_coo_kets(self, basis, kets):
        """the same loop for a subset of kets; cols = position in `kets`."""
        basis = _u8(basis).reshape(-1, self.S)
        kets = np.ascontiguousarray(kets, np.int64)
        n = len(basis)
        tot = lib().orc_offdiag_coo_kets(self._h, _p(basis), n, _p(kets), len(kets), None, None, None, 0)
        r = np.empty(tot, np.int64)
        c = np.empty(tot, np.int64)
        v = np.empty(tot, np.float32)
        lib().orc_offdiag_coo_kets(self._h, _p(basis), n, _p(kets), len(kets), _p(r), _p(c), _p(v), tot)
        return r, c, v

    # molecular.py:471-516 matrix_elements_fast (dense; FP64 diagonal, exact
    # float32 off-diagonals; later writes win like H[i, j] = elements[k])
    def dense_H(self, basis):
        basis = _u8(basis).reshape(-1, self.S)
        n = len(basis)
        H = np.zeros((n, n), np.float64)
        H[np.arange(n), np.arange(n)] = self.diag(basis)
        r, c, v = self.offdiag_coo(basis)
        H[r, c] = v.astype(np.float64)
        return H

    # skqd.py:135-177
    def fci_basis(self):
        from itertools import combinations
        rows = []
        for a in combinations(range(self.n_orb), self.n_alpha):
            for b in combinations(range(self.n_orb), self.n_beta):
                c = np.zeros(self.S, np.uint8)
                c[list(a)] = 1
                c[[i + self.n_orb for i in b]] = 1
                rows.append(c)
        return np.stack(rows)

    # skqd.py:374-419 : raw directed CSR over a basis, entries (row=i, col=j)
    def raw_csr(self, basis):
        import scipy.sparse as sp
        basis = _u8(basis).reshape(-1, self.S)
        n = len(basis)
        r, c, v = self.offdiag_coo(basis)
        rows = np.concatenate([np.arange(n), r])
        cols = np.concatenate([np.arange(n), c])
        data = np.concatenate([self.diag(basis), v.astype(np.float64)])
        M = sp.csr_matrix((data, (rows, cols)), shape=(n, n))
        M.sort_indices()
        return M

    # residual_expansion.py:408-443 (SelectedCIExpander._diagonalize)
    def diagonalize(self, basis):
        H = self.dense_H(basis)
        H = 0.5 * (H + H.T)                                       # :425
        n = len(H)
        if n > 500:                                               # :428-440
            from scipy.sparse import csr_matrix
            from scipy.sparse.linalg import eigsh
            w, v = eigsh(csr_matrix(H), k=1, which="SA", tol=1e-12, maxiter=1000)
            return float(w[0]), v[:, 0]
        w, v = np.linalg.eigh(H)                                  # :442
        return float(w[0]), v[:, 0]

    # residual_expansion.py:451-554
    def pt2_candidates(self, basis, eigenvector):
        basis = _u8(basis).reshape(-1, self.S)
        n = len(basis)
        c32 = np.asarray(eigenvector, np.float64).astype(np.float32)   # :481
        mag = np.abs(c32)
        order = np.argsort(-mag, kind="stable")                        # :486
        order = order[mag[order] > np.float32(1e-8)].astype(np.int64)  # :489-490
        cap = 1
        raw = np.zeros(1, np.int64)
        nu = lib().orc_pt2_candidates(self._h, _p(basis), n, _p(order), len(order), _p(c32),
                                      None, None, None, 0, _p(raw))
        cap = max(int(nu), 1)
        oc = np.empty((cap, self.S), np.uint8)
        c32o = np.empty(cap, np.float32)
        c64o = np.empty(cap, np.float64)
        nu = lib().orc_pt2_candidates(self._h, _p(basis), n, _p(order), len(order), _p(c32),
                                      _p(oc), _p(c32o), _p(c64o), cap, _p(raw))
        return oc[:nu], c32o[:nu], c64o[:nu], int(raw[0])

    def find_important_configs(self, basis, energy, eigenvector, k, precision="f64"):
        """-> (selected cfgs, importances, all candidates, all importances).
        precision="f32" mirrors the reference's float32 chain (:515-548);
        "f64" is the 1e-9 value oracle.  Ties: importance desc, key asc."""
        cand, c32, c64, raw = self.pt2_candidates(basis, eigenvector)
        if len(cand) == 0:
            return cand, np.zeros(0), cand, np.zeros(0), raw
        ex = self.diag(cand)                                           # :539
        if precision == "f32":
            coup = c32.astype(np.float32)
            den = np.abs(np.float32(energy) - ex.astype(np.float32)) + np.float32(1e-10)
            imp = (coup * coup) / den
        else:
            imp = (c64 * c64) / (np.abs(float(energy) - ex) + 1e-10)   # :547-548
        nsel = min(int(k), len(cand))                                  # :551
        if nsel == 0:
            return cand[:0], imp[:0], cand, imp, raw
        # The reference's torch.topk leaves ties unspecified (:552).  Deterministic protocol,
        # the same as the product's select_top_k: candidates within a relative 1e-9 of the k-th
        # importance count as tied with it (symmetry-equivalent determinants of a real molecule
        # are exactly degenerate; their FP64 sums differ in the last bits only through the
        # summation order) and the tie is broken by ascending key; result ordered by importance
        # descending, key ascending.
        pick = top_k_protocol(cand, imp, nsel)
        return cand[pick], imp[pick], cand, imp, raw

    # residual_expansion.py:334-406
    def expand_basis(self, basis, k):
        basis = _u8(basis).reshape(-1, self.S)
        E, v = self.diagonalize(basis)
        sel, imp, _, _, _ = self.find_important_configs(basis, E, v, k)
        if len(sel) == 0:
            return basis, dict(configs_added=0, energy=E, initial_energy=E, final_energy=E)
        exp = sort_unique(np.concatenate([basis, sel]))                # :367-368
        E2, _ = self.diagonalize(exp)
        if E - E2 < -1e-8:                                             # :376-393
            return basis, dict(configs_added=0, initial_energy=E, final_energy=E,
                               variational_violation=True, rejected_energy=E2)
        return exp, dict(initial_size=len(basis), final_size=len(exp), configs_added=len(sel),
                         initial_energy=E, final_energy=E2, energy_improvement=E - E2,
                         variational_violation=False)

    # skqd.py:683-807
    def ground_state_energy(self, basis, return_eigenvector=False, regularization=1e-8,
                            reference_compat=True):
        H = self.dense_H(basis)
        H = 0.5 * (H + H.T)                                            # :725
        n = len(H)
        if regularization > 0:
            H = H + regularization * np.eye(n)                         # :738-739
        if np.linalg.cond(H) > 1e12:                                   # :742-750 -> _svd_ground_state, :809-843
            U, s, Vh = np.linalg.svd(H, hermitian=True)
            thr = 1e-10 * s.max()
            H_reg = U @ np.diag(np.where(s > thr, s, thr)) @ Vh
            w, v = np.linalg.eigh(H_reg)
            return float(w[0]), (v[:, 0] if return_eigenvector else None)
        w, v = np.linalg.eigh(H)
        if n < 100 or return_eigenvector or not reference_compat:      # :754-758, :790-793
            return float(w[0]), v[:, 0]
        # :784-796 : eigsh(k=min(2,n-1), which='SA', return_eigenvectors=False)[0];
        # scipy returns the k values in ascending order of ... see SURVEY F5:
        # result[0] is the LARGER of the two lowest eigenvalues.
        kk = min(2, n - 1)
        return float(w[kk - 1]), None


def csr_matvec(indptr, indices, data, x):
    """scipy csr_matvec restatement (real H; x real or complex)."""
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int32)
    data = np.ascontiguousarray(data, np.float64)
    n = len(indptr) - 1
    if np.iscomplexobj(x):
        x = np.ascontiguousarray(x, np.complex128)
        y = np.empty(n, np.complex128)
        lib().orc_csr_matvec_z(n, _p(indptr), _p(indices), _p(data), _p(x), _p(y))
    else:
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(n, np.float64)
        lib().orc_csr_matvec_f64(n, _p(indptr), _p(indices), _p(data), _p(x), _p(y))
    return y


def expm_multiply_taylor(indptr, indices, data, psi, dt):
    """exp(-i dt H) psi by plain scaled Taylor series in complex128 (an
    independent check of scipy.sparse.linalg.expm_multiply, skqd.py:291-293)."""
    psi = np.asarray(psi, np.complex128)
    n = len(psi)
    colsum = np.zeros(n)
    np.add.at(colsum, indices, np.abs(data))
    nrm = dt * colsum.max()
    s = max(1, int(np.ceil(nrm)))
    out = psi.copy()
    for _ in range(s):
        term = out.copy()
        acc = out.copy()
        for j in range(1, 200):
            term = (-1j * dt / (s * j)) * csr_matvec(indptr, indices, data, term)
            acc = acc + term
            if np.abs(term).max() <= 1e-18 * max(np.abs(acc).max(), 1e-300):
                break
        out = acc
    return out
