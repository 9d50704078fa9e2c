"""The kernels' own index arithmetic (fgk_core.cuh, compiled for the host)
against the oracle and the golden vectors -- runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from conftest import load_golden
from helpers import hostcheck, pack_np, unpack_np

CASES = ["lih", "beh2", "n2", "ragged", "sparse", "edge_full_alpha", "edge_no_beta", "wide",
         "lih_sto3g", "beh2_sto3g", "n2_sto3g"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make(g):
    n_orb, na, nb = (int(x) for x in g["shape"])
    h1 = np.ascontiguousarray(g["h1"], np.float64)
    gg = np.ascontiguousarray(g["g"], np.float64)
    e_nuc = float(g.get("e_nuc", 0.0))
    hc = hostcheck().hc_ham_create(_p(h1), _p(gg), n_orb, na, nb, e_nuc)
    return hc, orc.OracleHam(g["h1"], g["g"], na, nb, e_nuc), n_orb


@pytest.mark.parametrize("name", CASES)
def test_pack_roundtrip_and_key_order(name):
    g = load_golden("ham_" + name)
    n_orb = int(g["shape"][0])
    dets = g["dets"]
    pk = pack_np(dets, n_orb)
    assert np.array_equal(unpack_np(pk, n_orb), dets)
    # (alpha, beta) lexicographic order == torch.unique(dim=0) row order
    order_words = np.lexsort((pk[:, 1], pk[:, 0]))
    order_rows = np.lexsort(tuple(dets[:, ::-1].T))
    assert np.array_equal(pk[order_words], pk[order_rows])


@pytest.mark.parametrize("name", CASES)
def test_ket_enumeration_matches_reference(name):
    g = load_golden("ham_" + name)
    hc, _, n_orb = make(g)
    offs = g["conn_offsets"]
    pk = pack_np(g["dets"], n_orb)
    for j in range(len(pk)):
        cap = int(offs[j + 1] - offs[j]) + 8
        od = np.zeros((cap, 2), np.uint64)
        oe = np.zeros(cap, np.float32)
        m = hostcheck().hc_connections(hc, int(pk[j, 0]), int(pk[j, 1]), _p(od), _p(oe), cap)
        assert m == offs[j + 1] - offs[j]
        assert np.array_equal(unpack_np(od[:m], n_orb), g["conn_cfgs"][offs[j]:offs[j + 1]])
        assert np.array_equal(oe[:m].view(np.uint32),
                              g["conn_elems"][offs[j]:offs[j + 1]].view(np.uint32))
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
def test_diag_matches_oracle_1e9(name):
    g = load_golden("ham_" + name)
    hc, H, n_orb = make(g)
    pk = pack_np(g["dets"], n_orb)
    out = np.zeros(len(pk))
    hostcheck().hc_diag(hc, _p(pk), len(pk), _p(out))
    assert np.abs(out - H.diag(g["dets"])).max() < 1e-9          # tolerance: 1e-9 Ha
    out2 = np.zeros(len(pk))
    hostcheck().hc_diag_loops(hc, _p(pk), len(pk), _p(out2))     # pair-loop form == nibble-table form
    assert np.abs(out - out2).max() < 1e-11
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_bra_rows_match_oracle_dense(name, mode):
    g = load_golden("ham_" + name)
    hc, H, n_orb = make(g)
    basis = g["basis"]
    pk = pack_np(basis, n_orb)
    D = H.dense_H(basis)
    if mode == 1:
        D = 0.5 * (D + D.T)
    n = len(basis)
    got = np.zeros_like(D)
    pattern = np.zeros(D.shape, bool)
    for i in range(n):
        cap = n + 4
        oc = np.zeros(cap, np.int32)
        ov = np.zeros(cap, np.float64)
        m = hostcheck().hc_bra_row(hc, _p(pk), n, i, mode, _p(oc), _p(ov), cap)
        assert m <= cap
        assert len(set(oc[:m].tolist())) == m
        got[i, oc[:m]] = ov[:m]
        pattern[i, oc[:m]] = True
    off = ~np.eye(n, dtype=bool)
    assert np.array_equal(got[off], D[off])                       # bit-exact off-diagonals
    assert np.abs(np.diag(got) - np.diag(D)).max() < 1e-9
    if mode == 0:
        # raw directed pattern == the reference's hits (explicit entries only where it writes)
        ref_pat = np.zeros(D.shape, bool)
        ref_pat[g["coo_rows"], g["coo_cols"]] = True
        assert np.array_equal(pattern & off, ref_pat & off)
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("scan", [0, 1])
def test_structured_bra_rows_equal_flat_enumeration(name, mode, scan):
    """the string-set driven row builder (k_projh2 strategy) produces exactly the entries of
    the flat reference-order enumeration"""
    g = load_golden("ham_" + name)
    hc, H, n_orb = make(g)
    basis = np.concatenate([g["basis"], g["dets"]])
    basis = np.unique(basis, axis=0)
    pk = pack_np(basis, n_orb)
    n = len(basis)
    for i in range(0, n, max(1, n // 25)):
        cap = n + 4
        c1, v1 = np.zeros(cap, np.int32), np.zeros(cap)
        c2, v2 = np.zeros(cap, np.int32), np.zeros(cap)
        m1 = hostcheck().hc_bra_row(hc, _p(pk), n, i, mode, _p(c1), _p(v1), cap)
        m2 = hostcheck().hc_bra_row2(hc, _p(pk), n, i, mode, scan, _p(c2), _p(v2), cap)
        assert m1 == m2
        o1, o2 = np.argsort(c1[:m1], kind="stable"), np.argsort(c2[:m2], kind="stable")
        assert np.array_equal(c1[:m1][o1], c2[:m2][o2])
        assert np.array_equal(v1[:m1][o1], v2[:m2][o2])
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_rank_based_bra_rows_equal_flat_enumeration(name, mode):
    """the rank-based row builder (k_projh3 strategy: sorted string lists, pair map, separable
    alpha-beta offsets and parities) produces exactly the entries of the flat enumeration"""
    g = load_golden("ham_" + name)
    hc, H, n_orb = make(g)
    basis = np.unique(np.concatenate([g["basis"], g["dets"]]), axis=0)
    pk = pack_np(basis, n_orb)
    n = len(basis)
    for i in range(0, n, max(1, n // 25)):
        cap = n + 4
        c1, v1 = np.zeros(cap, np.int32), np.zeros(cap)
        c3, v3 = np.zeros(cap, np.int32), np.zeros(cap)
        m1 = hostcheck().hc_bra_row(hc, _p(pk), n, i, mode, _p(c1), _p(v1), cap)
        m3 = hostcheck().hc_bra_row3(hc, _p(pk), n, i, mode, _p(c3), _p(v3), cap)
        assert m1 == m3
        o1, o3 = np.argsort(c1[:m1], kind="stable"), np.argsort(c3[:m3], kind="stable")
        assert np.array_equal(c1[:m1][o1], c3[:m3][o3])
        assert np.array_equal(v1[:m1][o1].view(np.uint64), v3[:m3][o3].view(np.uint64))
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_string_driven_bra_rows_equal_flat_enumeration(name, mode):
    """the string-driven row builder (k_projh4 strategy, fgk_lists.cuh: replacement lists with
    signed values, rows assembled from list entries) produces exactly the entries of the flat
    enumeration -- raw, symmetrised, and symmetrised with exact zeros dropped"""
    g = load_golden("ham_" + name)
    hc, H, n_orb = make(g)
    basis = np.unique(np.concatenate([g["basis"], g["dets"]]), axis=0)
    pk = pack_np(basis, n_orb)
    n = len(basis)
    for i in range(0, n, max(1, n // 25)):
        cap = n + 4
        c1, v1 = np.zeros(cap, np.int32), np.zeros(cap)
        c4, v4 = np.zeros(cap, np.int32), np.zeros(cap)
        m1 = hostcheck().hc_bra_row(hc, _p(pk), n, i, mode & 1, _p(c1), _p(v1), cap)
        m4 = hostcheck().hc_bra_row4(hc, _p(pk), n, i, mode, _p(c4), _p(v4), cap)
        if mode & 2:        # FGK_H_DROP_ZEROS: the flat walk keeps explicit zeros (F3 cancellations), drop them here
            keep = np.ones(m1, bool)
            keep[1:] = v1[1:m1] != 0.0
            c1[:keep.sum()], v1[:keep.sum()] = c1[:m1][keep], v1[:m1][keep]
            m1 = int(keep.sum())
        assert m1 == m4
        o1, o4 = np.argsort(c1[:m1], kind="stable"), np.argsort(c4[:m4], kind="stable")
        assert np.array_equal(c1[:m1][o1], c4[:m4][o4])
        assert np.array_equal(v1[:m1][o1].view(np.uint64), v4[:m4][o4].view(np.uint64))
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
def test_pt2_walk_by_singles_lists_equals_reference_connections(name):
    """the PT2 walk of k_pt2_accumulate2 (singles lists, separable alpha-beta elements) emits
    exactly the reference's connections with exactly its float32 elements, in another order"""
    g = load_golden("ham_" + name)
    hc, _, n_orb = make(g)
    offs = g["conn_offsets"]
    pk = pack_np(g["dets"], n_orb)
    for j in range(len(pk)):
        cap = int(offs[j + 1] - offs[j]) + 8
        od, oe = np.zeros((cap, 2), np.uint64), np.zeros(cap, np.float32)
        m = hostcheck().hc_pt2_walk2(hc, int(pk[j, 0]), int(pk[j, 1]), _p(od), _p(oe), cap)
        assert m == offs[j + 1] - offs[j]
        ref = pack_np(g["conn_cfgs"][offs[j]:offs[j + 1]], n_orb)
        ref_el = g["conn_elems"][offs[j]:offs[j + 1]].view(np.uint32)
        o_got = np.lexsort((od[:m, 1], od[:m, 0]))
        o_ref = np.lexsort((ref[:, 1], ref[:, 0]))
        assert np.array_equal(od[:m][o_got], ref[o_ref])
        assert np.array_equal(oe[:m].view(np.uint32)[o_got], ref_el[o_ref])
    hostcheck().hc_ham_destroy(hc)


def test_64_orbitals_bit63_edges():
    """n_orb = 64 uses every bit of both words (shift-by-64 hazards, sign bit)."""
    from helpers import synth_integrals
    n_orb, na, nb = 64, 2, 1
    rng = np.random.default_rng(64)
    h1 = rng.standard_normal((n_orb, n_orb)); h1 = 0.5 * (h1 + h1.T)
    g = np.zeros((n_orb,) * 4)
    idx = rng.integers(0, n_orb, size=(40000, 4))
    vals = rng.standard_normal(40000) * 0.1
    for perm in ((0, 1, 2, 3), (1, 0, 2, 3), (0, 1, 3, 2), (1, 0, 3, 2), (2, 3, 0, 1), (3, 2, 0, 1), (2, 3, 1, 0), (3, 2, 1, 0)):
        g[idx[:, perm[0]], idx[:, perm[1]], idx[:, perm[2]], idx[:, perm[3]]] = vals
    hc = hostcheck().hc_ham_create(_p(np.ascontiguousarray(h1)), _p(g), n_orb, na, nb, 0.25)
    H = orc.OracleHam(h1.astype(np.float32), g.astype(np.float32), na, nb, 0.25)
    dets = np.zeros((4, 2 * n_orb), np.uint8)
    dets[0, [0, 1]] = 1; dets[0, n_orb + 0] = 1            # lowest orbitals (highest bits)
    dets[1, [62, 63]] = 1; dets[1, n_orb + 63] = 1         # highest orbitals (bit 0)
    dets[2, [0, 63]] = 1; dets[2, n_orb + 31] = 1
    dets[3, [17, 40]] = 1; dets[3, n_orb + 5] = 1
    pk = pack_np(dets, n_orb)
    assert np.array_equal(unpack_np(pk, n_orb), dets)
    out = np.zeros(4)
    hostcheck().hc_diag(hc, _p(pk), 4, _p(out))
    assert np.abs(out - H.diag(dets)).max() < 1e-9
    for j in range(4):
        oc, oe = H.connections(dets[j])
        cap = len(oc) + 8
        od, el = np.zeros((cap, 2), np.uint64), np.zeros(cap, np.float32)
        m = hostcheck().hc_connections(hc, int(pk[j, 0]), int(pk[j, 1]), _p(od), _p(el), cap)
        assert m == len(oc)
        assert np.array_equal(unpack_np(od[:m], n_orb), oc)
        assert np.array_equal(el[:m].view(np.uint32), oe.view(np.uint32))
    hostcheck().hc_ham_destroy(hc)


@pytest.mark.parametrize("name", CASES)
def test_split_value_and_fast_parity_equal_generic_elements(name):
    g = load_golden("ham_" + name)
    hc, _, n_orb = make(g)
    dets = np.concatenate([g["dets"], g["basis"]])
    pk = pack_np(dets, n_orb)
    assert hostcheck().hc_check_split(hc, _p(pk), len(pk)) == 0
    hostcheck().hc_ham_destroy(hc)


def _random_case(seed):
    rng = np.random.default_rng(seed)
    n_orb = int(rng.integers(2, 10))
    na = int(rng.integers(0, n_orb + 1))
    nb = int(rng.integers(0, n_orb + 1))
    h1 = rng.standard_normal((n_orb, n_orb)); h1 = 0.5 * (h1 + h1.T)
    g = rng.standard_normal((n_orb,) * 4) * 0.2
    if seed % 2:                                   # 8-fold symmetric half of the time
        g = g + g.transpose(1, 0, 2, 3); g = g + g.transpose(0, 1, 3, 2); g = g + g.transpose(2, 3, 0, 1)
    if seed % 3 == 0:                              # zeros so that the 1e-12 filters act
        h1[rng.random(h1.shape) < 0.4] = 0.0
        g[rng.random(g.shape) < 0.5] = 0.0
    return n_orb, na, nb, h1, g, rng


@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_against_oracle(seed):
    """random orbital counts / fillings (incl. empty and full spin blocks), NON-symmetric
    integrals and sparsified tables: enumeration, diagonal and both row builders vs the oracle"""
    from helpers import random_dets
    n_orb, na, nb, h1, g, rng = _random_case(seed)
    hc = hostcheck().hc_ham_create(_p(np.ascontiguousarray(h1)), _p(np.ascontiguousarray(g)), n_orb, na, nb, 0.3)
    H = orc.OracleHam(h1.astype(np.float32), g.astype(np.float32), na, nb, 0.3)
    dets = np.unique(random_dets(n_orb, na, nb, 12, rng), axis=0)
    pk = pack_np(dets, n_orb)
    out = np.zeros(len(pk))
    hostcheck().hc_diag(hc, _p(pk), len(pk), _p(out))
    assert np.abs(out - H.diag(dets)).max() < 1e-9
    for j in range(len(pk)):
        oc, oe = H.connections(dets[j])
        cap = len(oc) + 4
        od, el = np.zeros((cap, 2), np.uint64), np.zeros(cap, np.float32)
        m = hostcheck().hc_connections(hc, int(pk[j, 0]), int(pk[j, 1]), _p(od), _p(el), cap)
        assert m == len(oc)
        assert np.array_equal(unpack_np(od[:m], n_orb), oc)
        assert np.array_equal(el[:m].view(np.uint32), oe.view(np.uint32))
    assert hostcheck().hc_check_split(hc, _p(pk), len(pk)) == 0
    n = len(dets)
    D = H.dense_H(dets)
    for mode, ref in ((0, D), (1, 0.5 * (D + D.T))):
        for builder in (0, 1, 2, 3):
            got = np.zeros_like(D)
            for i in range(n):
                cap = n + 4
                c, v = np.zeros(cap, np.int32), np.zeros(cap)
                if builder == 0:
                    m = hostcheck().hc_bra_row(hc, _p(pk), n, i, mode, _p(c), _p(v), cap)
                elif builder == 3:
                    m = hostcheck().hc_bra_row3(hc, _p(pk), n, i, mode, _p(c), _p(v), cap)
                else:
                    m = hostcheck().hc_bra_row2(hc, _p(pk), n, i, mode, builder - 1, _p(c), _p(v), cap)
                got[i, c[:m]] = v[:m]
            off = ~np.eye(n, dtype=bool)
            assert np.array_equal(got[off], ref[off])
    hostcheck().hc_ham_destroy(hc)


def test_pt2_fixed_point_accumulator_is_exact_and_order_independent():
    """fgk_pt2.cu accumulates couplings as 128-bit fixed point (fx_from_double / fx_atomic_add /
    fx_to_double): any order gives the same bits, and the result is the correctly rounded sum of
    the addends (each rounded to 2^-70 first)."""
    import math
    from fractions import Fraction
    L = hostcheck()
    rng = np.random.default_rng(7)
    for trial in range(20):
        n = int(rng.integers(1, 4000))
        scale = 10.0 ** rng.integers(-14, 6)
        v = (rng.standard_normal(n) * scale * 10.0 ** rng.integers(-6, 1, n)).astype(np.float64)
        if trial % 4 == 0:
            v[: n // 2] = -v[n // 2: 2 * (n // 2)][::-1]        # heavy cancellation
        outs = []
        for perm in (np.arange(n), np.arange(n)[::-1].copy(), rng.permutation(n)):
            out = np.zeros(1)
            perm = np.ascontiguousarray(perm, dtype=np.int64)
            assert L.hc_fx_sum(v.ctypes.data, perm.ctypes.data, n, out.ctypes.data) == 0
            outs.append(out[0])
        assert outs[0].tobytes() == outs[1].tobytes() == outs[2].tobytes()
        # exact reference: every addend rounded to a multiple of 2^-70 (half away from zero), exact rational sum
        q = Fraction(1, 2 ** 70)
        tot = Fraction(0)
        for x in v:
            fx = Fraction(float(x)) / q
            r = math.floor(abs(fx) + Fraction(1, 2))
            tot += (r if fx >= 0 else -r) * q
        assert outs[0] == float(tot)                           # float(Fraction) rounds correctly
        assert abs(outs[0] - math.fsum(v)) <= 1e-20 * n + 4e-16 * abs(math.fsum(v))
    # range guard: |v| >= 2^30, inf, nan are refused
    for bad in (2.0 ** 30, -1e12, np.inf, np.nan):
        out = np.zeros(1)
        one = np.zeros(1, np.int64)
        assert L.hc_fx_sum(np.array([bad]).ctypes.data, one.ctypes.data, 1, out.ctypes.data) == -1
    out = np.zeros(1)
    one = np.zeros(1, np.int64)
    assert L.hc_fx_sum(np.array([2.0 ** 30 - 1.0]).ctypes.data, one.ctypes.data, 1, out.ctypes.data) == 0
    assert out[0] == 2.0 ** 30 - 1.0
