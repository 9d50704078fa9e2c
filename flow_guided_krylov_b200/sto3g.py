"""Integral front-end without PySCF.

The reference gets its integrals from PySCF (`compute_molecular_integrals`,
src/hamiltonians/molecular.py:945-1003: `gto.Mole` in STO-3G, `scf.RHF`,
`mo_coeff.T @ hcore @ mo_coeff`, `ao2mo` restored to the 4-index (pq|rs) tensor).
PySCF is not available offline, so BASELINE configs[0..2] (LiH / BeH2 / N2 in STO-3G)
would otherwise run on molecule-shaped synthetic integrals only.  This module
restates that recipe in numpy for first-row atoms:

  * STO-3G (Hehre, Stewart, Pople 1969): the universal 3-Gaussian fits of the 1s
    and 2sp Slater functions, scaled by the standard molecular exponents zeta^2;
  * one- and two-electron integrals over contracted Cartesian s / p Gaussians by the
    McMurchie-Davidson scheme (Hermite expansion coefficients + Boys function);
  * closed-shell RHF with DIIS;
  * AO -> MO transformation to the `MolecularIntegrals` fields the reference fills.

It is host-side set-up code (a few seconds per molecule), not part of the GPU hot path;
the factories below mirror the reference's names and default geometries
(molecular.py:1006-1139).  Checked in tests/test_sto3g.py against textbook STO-3G
energies (H2 RHF / FCI, HeH+) and against the number the reference publishes for its own
LiH Hamiltonian (SKQD_VALIDATION_REPORT.md:87).
"""
from math import exp, pi, sqrt
from typing import List, Sequence, Tuple

import numpy as np

ANGSTROM_TO_BOHR = 1.0 / 0.52917721092          # PySCF's (CODATA 2010 / 2014) Bohr radius in Angstrom

Z = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9}
# standard molecular Slater exponents of STO-3G: (1s,) or (1s, 2sp)
ZETA = {"H": (1.24,), "He": (1.69,), "Li": (2.69, 0.80), "Be": (3.68, 1.15), "B": (4.68, 1.50),
        "C": (5.67, 1.72), "N": (6.67, 1.95), "O": (7.66, 2.25), "F": (8.65, 2.55)}
# least-squares fits of Slater functions with zeta = 1 by three Gaussians
_A1S = (2.227660584, 0.405771156, 0.1098175104)
_D1S = (0.154328967, 0.535328142, 0.444634542)
_A2SP = (0.994203, 0.231031, 0.0751386)
_D2S = (-0.0999672292, 0.399512826, 0.700115469)
_D2P = (0.155916275, 0.607683719, 0.391957393)
# exponents as tabulated (and as PySCF ships them) for the elements of the reference's molecules;
# other first-row atoms use fit x zeta^2, which agrees with the table to ~1e-7 relative
_TABLE = {
    "H": ((3.42525091, 0.62391373, 0.16885540),),
    "Li": ((16.1195750, 2.9362007, 0.7946505), (0.6362897, 0.1478601, 0.0480887)),
    "Be": ((30.1678710, 5.4951153, 1.4871927), (1.3148331, 0.3055389, 0.0993707)),
    "C": ((71.6168370, 13.0450960, 3.5305122), (2.9412494, 0.6834831, 0.2222899)),
    "N": ((99.1061690, 18.0523120, 4.8856602), (3.7804559, 0.8784966, 0.2857144)),
    "O": ((130.7093200, 23.8088610, 6.4436083), (5.0331513, 1.1695961, 0.3803890)),
}
_D1S_TAB = (0.15432897, 0.53532814, 0.44463454)
_D2S_TAB = (-0.09996723, 0.39951283, 0.70011547)
_D2P_TAB = (0.15591627, 0.60768372, 0.39195739)


def shell_exponents(sym):
    """((1s exponents), (2sp exponents))-tuple of an element."""
    if sym in _TABLE:
        return _TABLE[sym]
    zeta = ZETA[sym]
    out = (tuple(a * zeta[0] ** 2 for a in _A1S),)
    if len(zeta) > 1:
        out += (tuple(a * zeta[1] ** 2 for a in _A2SP),)
    return out


class _CGF:
    """contracted Cartesian Gaussian: centre, angular momentum (lx,ly,lz), exponents, coefficients
    (coefficients include the primitive norms and the contraction norm)."""

    def __init__(self, centre, lmn, exps, coefs):
        self.A = np.asarray(centre, float)
        self.lmn = tuple(lmn)
        self.exps = np.asarray(exps, float)
        l = sum(lmn)
        # primitive normalisation for s (l=0) and p (l=1)
        norm = (2.0 * self.exps / pi) ** 0.75 * (4.0 * self.exps) ** (l / 2.0)
        self.coefs = np.asarray(coefs, float) * norm
        # normalise the contraction
        s = 0.0
        for a, ca in zip(self.exps, self.coefs):
            for b, cb in zip(self.exps, self.coefs):
                s += ca * cb * _overlap_prim(a, self.lmn, self.A, b, self.lmn, self.A)
        self.coefs = self.coefs / sqrt(s)


def build_basis(geometry: Sequence[Tuple[str, Sequence[float]]]) -> List[_CGF]:
    """STO-3G functions in the order 1s, 2s, 2px, 2py, 2pz per atom; geometry in Angstrom."""
    out = []
    for sym, xyz in geometry:
        R = np.asarray(xyz, float) * ANGSTROM_TO_BOHR
        shells = shell_exponents(sym)
        out.append(_CGF(R, (0, 0, 0), shells[0], _D1S_TAB))
        if len(shells) > 1:
            out.append(_CGF(R, (0, 0, 0), shells[1], _D2S_TAB))
            for lmn in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                out.append(_CGF(R, lmn, shells[1], _D2P_TAB))
    return out


# ---- McMurchie-Davidson machinery ---------------------------------------------------------------
def _E(i, j, t, Qx, a, b):
    """Hermite expansion coefficient E_t^{ij} of the product of two 1D Gaussians."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return exp(-q * Qx * Qx)
    if j == 0:
        return (1.0 / (2.0 * p)) * _E(i - 1, j, t - 1, Qx, a, b) - (q * Qx / a) * _E(i - 1, j, t, Qx, a, b) + \
            (t + 1) * _E(i - 1, j, t + 1, Qx, a, b)
    return (1.0 / (2.0 * p)) * _E(i, j - 1, t - 1, Qx, a, b) + (q * Qx / b) * _E(i, j - 1, t, Qx, a, b) + \
        (t + 1) * _E(i, j - 1, t + 1, Qx, a, b)


def _overlap_prim(a, lmn1, A, b, lmn2, B):
    p = a + b
    s = 1.0
    for k in range(3):
        s *= _E(lmn1[k], lmn2[k], 0, A[k] - B[k], a, b)
    return s * (pi / p) ** 1.5


def _kinetic_prim(a, lmn1, A, b, lmn2, B):
    l2, m2, n2 = lmn2
    t0 = b * (2 * (l2 + m2 + n2) + 3) * _overlap_prim(a, lmn1, A, b, lmn2, B)
    t1 = -2.0 * b * b * (_overlap_prim(a, lmn1, A, b, (l2 + 2, m2, n2), B) +
                         _overlap_prim(a, lmn1, A, b, (l2, m2 + 2, n2), B) +
                         _overlap_prim(a, lmn1, A, b, (l2, m2, n2 + 2), B))
    t2 = -0.5 * (l2 * (l2 - 1) * _overlap_prim(a, lmn1, A, b, (l2 - 2, m2, n2), B) +
                 m2 * (m2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2 - 2, n2), B) +
                 n2 * (n2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2, n2 - 2), B))
    return t0 + t1 + t2


def _boys(n, x):
    """F_n(x) = int_0^1 t^(2n) exp(-x t^2) dt."""
    if x < 1e-8:
        return 1.0 / (2 * n + 1) - x / (2 * n + 3)
    from scipy.special import gamma, gammainc
    return 0.5 * x ** (-(n + 0.5)) * gamma(n + 0.5) * gammainc(n + 0.5, x)


def _R(t, u, v, n, p, PC, RPC2, cache):
    """Hermite Coulomb integral R^n_{tuv}."""
    key = (t, u, v, n)
    if key in cache:
        return cache[key]
    if t < 0 or u < 0 or v < 0:
        return 0.0
    if t == u == v == 0:
        val = (-2.0 * p) ** n * _boys(n, p * RPC2)
    elif t == u == 0:
        val = (v - 1) * _R(t, u, v - 2, n + 1, p, PC, RPC2, cache) + PC[2] * _R(t, u, v - 1, n + 1, p, PC, RPC2, cache)
    elif t == 0:
        val = (u - 1) * _R(t, u - 2, v, n + 1, p, PC, RPC2, cache) + PC[1] * _R(t, u - 1, v, n + 1, p, PC, RPC2, cache)
    else:
        val = (t - 1) * _R(t - 2, u, v, n + 1, p, PC, RPC2, cache) + PC[0] * _R(t - 1, u, v, n + 1, p, PC, RPC2, cache)
    cache[key] = val
    return val


def _nuclear_prim(a, lmn1, A, b, lmn2, B, C):
    p = a + b
    P = (a * A + b * B) / p
    PC = P - C
    RPC2 = float(PC @ PC)
    cache = {}
    val = 0.0
    for t in range(lmn1[0] + lmn2[0] + 1):
        Et = _E(lmn1[0], lmn2[0], t, A[0] - B[0], a, b)
        for u in range(lmn1[1] + lmn2[1] + 1):
            Eu = _E(lmn1[1], lmn2[1], u, A[1] - B[1], a, b)
            for v in range(lmn1[2] + lmn2[2] + 1):
                Ev = _E(lmn1[2], lmn2[2], v, A[2] - B[2], a, b)
                val += Et * Eu * Ev * _R(t, u, v, 0, p, PC, RPC2, cache)
    return 2.0 * pi / p * val


def _hermite_pair(f1, f2):
    """all primitive pairs of two contracted functions: (p, P, coefficient, {(t,u,v): E_t E_u E_v})."""
    out = []
    for a, ca in zip(f1.exps, f1.coefs):
        for b, cb in zip(f2.exps, f2.coefs):
            p = a + b
            P = (a * f1.A + b * f2.A) / p
            E = {}
            for t in range(f1.lmn[0] + f2.lmn[0] + 1):
                Et = _E(f1.lmn[0], f2.lmn[0], t, f1.A[0] - f2.A[0], a, b)
                for u in range(f1.lmn[1] + f2.lmn[1] + 1):
                    Eu = _E(f1.lmn[1], f2.lmn[1], u, f1.A[1] - f2.A[1], a, b)
                    for v in range(f1.lmn[2] + f2.lmn[2] + 1):
                        Ev = _E(f1.lmn[2], f2.lmn[2], v, f1.A[2] - f2.A[2], a, b)
                        e = Et * Eu * Ev
                        if e != 0.0:
                            E[(t, u, v)] = e
            out.append((p, P, ca * cb, E))
    return out


def _eri_contracted(pair_ab, pair_cd):
    val = 0.0
    for p, P, cab, Eab in pair_ab:
        for q, Q, ccd, Ecd in pair_cd:
            alpha = p * q / (p + q)
            PQ = P - Q
            RPQ2 = float(PQ @ PQ)
            cache = {}
            s = 0.0
            for (t, u, v), e1 in Eab.items():
                for (tt, uu, vv), e2 in Ecd.items():
                    s += e1 * e2 * (-1.0) ** (tt + uu + vv) * _R(t + tt, u + uu, v + vv, 0, alpha, PQ, RPQ2, cache)
            val += cab * ccd * s * 2.0 * pi ** 2.5 / (p * q * sqrt(p + q))
    return val


def ao_integrals(geometry):
    """-> S, T, V, eri (chemist order (ab|cd)), E_nuc, n_electrons for a neutral molecule."""
    basis = build_basis(geometry)
    n = len(basis)
    nuclei = [(Z[s], np.asarray(x, float) * ANGSTROM_TO_BOHR) for s, x in geometry]
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            fi, fj = basis[i], basis[j]
            s = t = v = 0.0
            for a, ca in zip(fi.exps, fi.coefs):
                for b, cb in zip(fj.exps, fj.coefs):
                    c = ca * cb
                    s += c * _overlap_prim(a, fi.lmn, fi.A, b, fj.lmn, fj.A)
                    t += c * _kinetic_prim(a, fi.lmn, fi.A, b, fj.lmn, fj.A)
                    for zc, C in nuclei:
                        v -= zc * c * _nuclear_prim(a, fi.lmn, fi.A, b, fj.lmn, fj.A, C)
            S[i, j] = S[j, i] = s
            T[i, j] = T[j, i] = t
            V[i, j] = V[j, i] = v
    pairs = {(i, j): _hermite_pair(basis[i], basis[j]) for i in range(n) for j in range(i + 1)}
    eri = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(n):
                for l in range(k + 1):
                    if k * (k + 1) // 2 + l > ij:
                        continue
                    val = _eri_contracted(pairs[(i, j)], pairs[(k, l)])
                    for a, b, c, d in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k),
                                       (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                        eri[a, b, c, d] = val
    e_nuc = 0.0
    for a in range(len(nuclei)):
        for b in range(a):
            e_nuc += nuclei[a][0] * nuclei[b][0] / np.linalg.norm(nuclei[a][1] - nuclei[b][1])
    return S, T, V, eri, e_nuc, sum(z for z, _ in nuclei)


def rhf(S, hcore, eri, n_electrons, e_nuc, conv=1e-12, max_iter=200, guess="gwh"):
    """closed-shell RHF with DIIS; -> (E_total, C, orbital energies).  guess: 'core' (bare
    nuclei) or 'gwh' (generalised Wolfsberg-Helmholz, F_ij = 1.75 S_ij (h_ii + h_jj) / 2)."""
    n_occ = n_electrons // 2
    s_val, s_vec = np.linalg.eigh(S)
    X = s_vec @ np.diag(s_val ** -0.5) @ s_vec.T
    F0 = hcore
    if guess == "gwh":
        dh = np.diag(hcore)
        F0 = 1.75 * S * (dh[:, None] + dh[None, :]) / 2.0
        F0[np.diag_indices_from(F0)] = dh
    eps, Cp = np.linalg.eigh(X.T @ F0 @ X)
    C = X @ Cp
    D = 2.0 * C[:, :n_occ] @ C[:, :n_occ].T
    fock_list, err_list = [], []
    e_old = 0.0
    for _ in range(max_iter):
        J = np.einsum("pqrs,rs->pq", eri, D)
        K = np.einsum("prqs,rs->pq", eri, D)
        F = hcore + J - 0.5 * K
        e_el = 0.5 * np.sum(D * (hcore + F))
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        fock_list.append(F)
        err_list.append(err)
        if len(fock_list) > 8:
            fock_list.pop(0)
            err_list.pop(0)
        if len(fock_list) > 1:
            m = len(fock_list)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for a in range(m):
                for b in range(m):
                    B[a, b] = np.sum(err_list[a] * err_list[b])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                w = np.linalg.solve(B, rhs)[:m]
                F = sum(wi * Fi for wi, Fi in zip(w, fock_list))
            except np.linalg.LinAlgError:
                pass
        eps, Cp = np.linalg.eigh(X.T @ F @ X)
        C = X @ Cp
        D_new = 2.0 * C[:, :n_occ] @ C[:, :n_occ].T
        done = abs(e_el - e_old) < conv and np.abs(err).max() < 1e-9
        D, e_old = D_new, e_el
        if done:
            break
    return e_el + e_nuc, C, eps


def compute_molecular_integrals(geometry, basis: str = "sto-3g", charge: int = 0, spin: int = 0):
    """Drop-in for the reference's compute_molecular_integrals (molecular.py:945-1003) for
    closed-shell first-row molecules in STO-3G.  Returns flow_guided_krylov_b200.MolecularIntegrals
    (h1e, h2e in the RHF MO basis, chemist order)."""
    from .hamiltonian import MolecularIntegrals
    if basis.lower().replace("-", "") != "sto3g":
        raise ValueError("only STO-3G is built in (use PySCF for other basis sets)")
    if spin != 0:
        raise ValueError("closed-shell RHF only (spin = 0)")
    geometry = [(s, tuple(float(v) for v in x)) for s, x in geometry]
    S, T, V, eri, e_nuc, n_el = ao_integrals(geometry)
    n_el -= charge
    hcore = T + V
    # the SCF equations have several solutions; start from two different guesses and keep the
    # lower one (from the bare-nuclei guess alone N2 lands on a saddle 0.7 Ha above the ground state)
    e_hf, C = None, None
    for guess in ("gwh", "core"):
        e, Cg, _ = rhf(S, hcore, eri, n_el, e_nuc, guess=guess)
        if e_hf is None or e < e_hf - 1e-9:
            e_hf, C = e, Cg
    h1 = C.T @ hcore @ C
    g = np.einsum("pi,qj,pqrs,rk,sl->ijkl", C, C, eri, C, C, optimize=True)
    out = MolecularIntegrals(h1e=h1, h2e=g, nuclear_repulsion=float(e_nuc), n_electrons=int(n_el),
                             n_orbitals=len(S), n_alpha=n_el // 2, n_beta=n_el // 2)
    out.hf_energy = float(e_hf)
    return out


# geometries of the reference's factories (molecular.py:1006-1139)
def h2_geometry(bond_length=0.74):
    return [("H", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, bond_length))]


def lih_geometry(bond_length=1.6):
    return [("Li", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, bond_length))]


def h2o_geometry(oh_length=0.96, angle=104.5):
    a = np.radians(angle)
    return [("O", (0.0, 0.0, 0.0)), ("H", (oh_length, 0.0, 0.0)),
            ("H", (oh_length * np.cos(a), oh_length * np.sin(a), 0.0))]


def beh2_geometry(bond_length=1.33):
    return [("Be", (0.0, 0.0, 0.0)), ("H", (0.0, 0.0, bond_length)), ("H", (0.0, 0.0, -bond_length))]


def nh3_geometry(nh_length=1.01, hnh_angle=107.8):
    ang = np.radians(hnh_angle)
    h = nh_length * np.cos(np.arcsin(np.sin(ang / 2) / np.sin(np.radians(60))))
    r = np.sqrt(nh_length ** 2 - h ** 2)
    return [("N", (0.0, 0.0, h)), ("H", (r, 0.0, 0.0)),
            ("H", (r * np.cos(np.radians(120)), r * np.sin(np.radians(120)), 0.0)),
            ("H", (r * np.cos(np.radians(240)), r * np.sin(np.radians(240)), 0.0))]


def n2_geometry(bond_length=1.10):
    return [("N", (0.0, 0.0, 0.0)), ("N", (0.0, 0.0, bond_length))]


def ch4_geometry(ch_length=1.09):
    a = ch_length / np.sqrt(3)
    return [("C", (0.0, 0.0, 0.0)), ("H", (a, a, a)), ("H", (a, -a, -a)), ("H", (-a, a, -a)),
            ("H", (-a, -a, a))]


def _make(geometry, device):
    from .hamiltonian import MolecularHamiltonian
    return MolecularHamiltonian(compute_molecular_integrals(geometry, basis="sto-3g"), device=device)


# same names, parameters and defaults as the reference's factories (device: CUDA only here)
def create_h2_hamiltonian(bond_length: float = 0.74, device: str = "cuda"):
    return _make(h2_geometry(bond_length), device)


def create_lih_hamiltonian(bond_length: float = 1.6, device: str = "cuda"):
    return _make(lih_geometry(bond_length), device)


def create_h2o_hamiltonian(oh_length: float = 0.96, angle: float = 104.5, device: str = "cuda"):
    return _make(h2o_geometry(oh_length, angle), device)


def create_beh2_hamiltonian(bond_length: float = 1.33, device: str = "cuda"):
    return _make(beh2_geometry(bond_length), device)


def create_nh3_hamiltonian(nh_length: float = 1.01, hnh_angle: float = 107.8, device: str = "cuda"):
    return _make(nh3_geometry(nh_length, hnh_angle), device)


def create_n2_hamiltonian(bond_length: float = 1.10, device: str = "cuda"):
    return _make(n2_geometry(bond_length), device)


def create_ch4_hamiltonian(ch_length: float = 1.09, device: str = "cuda"):
    return _make(ch4_geometry(ch_length), device)
