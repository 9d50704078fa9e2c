"""The CPU oracle (oracle/) pinned against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import oracle as orc
from conftest import load_golden

HAM_CASES = ["lih", "beh2", "n2", "ragged", "sparse", "edge_full_alpha", "edge_no_beta", "wide",
             "lih_sto3g", "beh2_sto3g", "n2_sto3g"]        # *_sto3g: real molecules (sto3g.py integrals)


def ham_of(g):
    n_orb, na, nb = (int(x) for x in g["shape"])
    return orc.OracleHam(g["h1"], g["g"], na, nb, float(g.get("e_nuc", 0.0)))


@pytest.mark.parametrize("name", HAM_CASES)
def test_connections_emission_order_and_values(name):
    g = load_golden("ham_" + name)
    H = ham_of(g)
    offs = g["conn_offsets"]
    for j, det in enumerate(g["dets"]):
        c, e = H.connections(det)
        ref_c = g["conn_cfgs"][offs[j]:offs[j + 1]]
        ref_e = g["conn_elems"][offs[j]:offs[j + 1]]
        assert c.shape == ref_c.shape
        assert np.array_equal(c, ref_c)                 # same dets, same order
        assert np.array_equal(e.view(np.uint32), ref_e.view(np.uint32))   # bit-exact float32


@pytest.mark.parametrize("name", HAM_CASES)
def test_batch_matches_single(name):
    g = load_golden("ham_" + name)
    H = ham_of(g)
    c, e, src, offs = H.connections_batch(g["dets"])
    assert np.array_equal(offs, g["conn_offsets"])
    assert np.array_equal(c, g["conn_cfgs"])
    assert np.array_equal(e.view(np.uint32), g["conn_elems"].view(np.uint32))
    assert np.array_equal(src, np.repeat(np.arange(len(g["dets"])), np.diff(offs)))


@pytest.mark.parametrize("name", HAM_CASES)
def test_diag_fp64_within_float32_envelope(name):
    g = load_golden("ham_" + name)
    H = ham_of(g)
    d = H.diag(g["dets"])
    # the reference diagonal is a float32 einsum: agreement is float32-limited
    scale = max(1.0, float(np.abs(d).max()))
    assert np.abs(d - g["diag32"].astype(np.float64)).max() < 2e-5 * scale


@pytest.mark.parametrize("name", HAM_CASES)
def test_projected_h_offdiag_bit_exact(name):
    g = load_golden("ham_" + name)
    H = ham_of(g)
    r, c, v = H.offdiag_coo(g["basis"])
    assert np.array_equal(r, g["coo_rows"])
    assert np.array_equal(c, g["coo_cols"])
    assert np.array_equal(v.view(np.uint32), g["coo_vals"].astype(np.float32).view(np.uint32))
    D = H.dense_H(g["basis"])
    ref = g["H_dense32"].astype(np.float64)
    off = ~np.eye(len(D), dtype=bool)
    assert np.array_equal(D[off], ref[off])             # pattern + values, bit-exact
    assert np.abs(np.diag(D) - np.diag(ref)).max() < 2e-5 * max(1.0, np.abs(ref).max())


def test_sign_double_antisymmetry_quirk_F3():
    """molecular.py:391-423 is not antisymmetry-consistent: some same-spin
    doubles have H[i,j] = -H[j,i].  The oracle must reproduce it."""
    g = load_golden("skqd_lih")
    H = ham_of(g)
    D = H.dense_H(g["subspace"])
    off = D - np.diag(np.diag(D))
    anti = np.abs(off + off.T) < 1e-30
    sym = np.abs(off - off.T) < 1e-30
    nz = np.abs(off) > 0
    assert (anti & nz).sum() > 0 and (sym & nz).sum() > 0
    assert ((anti | sym) | ~nz).all()


@pytest.mark.parametrize("name", ["lih", "beh2", "lih_sto3g"])
def test_fci_energy(name):
    g = load_golden("fci_" + name)
    H = ham_of(g)
    E, _ = H.diagonalize(H.fci_basis())
    assert abs(E - float(g["fci"])) < 5e-6          # float32 diagonal envelope


@pytest.mark.parametrize("name", ["lih", "beh2", "beh2_wide", "lih_sto3g", "beh2_sto3g", "n2_sto3g"])
def test_selected_ci_rounds(name):
    g = load_golden("sci_" + name)
    H = ham_of(g)
    k = int(g["k"])
    basis = g["basis0"]
    for rd in range(int(g["rounds"])):
        # selection on the REFERENCE's own eigenpair: candidates + importances
        sel, imp, cand, imp_all, _ = H.find_important_configs(
            basis, float(g[f"r{rd}_E"]), g[f"r{rd}_v"], k, precision="f32")
        ref_sel, ref_imp = g[f"r{rd}_sel"], g[f"r{rd}_imp"]
        assert len(sel) == len(ref_sel)
        # same selected SET unless a near-tie sits at the cut (epsilon band)
        got = {bytes(r) for r in sel}
        want = {bytes(r) for r in ref_sel}
        if got != want:
            cut = float(ref_imp.min())
            for r in got ^ want:
                i = [bytes(x) for x in cand].index(r)
                assert abs(imp_all[i] - cut) <= 1e-5 * cut
        srt = np.sort(ref_imp.astype(np.float64))[::-1]
        assert np.allclose(np.sort(imp.astype(np.float64))[::-1], srt, rtol=2e-4, atol=1e-12)
        # full round through the oracle's own numerics (FP64, deterministic tie protocol)
        new_basis, st = H.expand_basis(basis, k)
        ref_basis = g[f"r{rd}_basis"]
        if np.array_equal(new_basis, ref_basis):
            # float32 diagonals in the reference: the envelope scales with |E| (N2: ~108 Ha)
            assert abs(st["final_energy"] - float(g[f"r{rd}_final_energy"])) < 5e-6 * max(1.0, abs(st["final_energy"]) / 8.0)
        else:
            # real molecules: symmetry-equivalent determinants have exactly degenerate importances;
            # when such a group straddles the cut the reference's float32 topk picks by rounding
            # noise.  Only members of that group may differ, and as many are taken.
            assert len(new_basis) == len(ref_basis)
            cut = float(ref_imp.min())
            keys = [bytes(x) for x in cand]
            for r in {bytes(x) for x in new_basis} ^ {bytes(x) for x in ref_basis}:
                assert abs(imp_all[keys.index(r)] - cut) <= 1e-5 * cut
            assert abs(st["final_energy"] - float(g[f"r{rd}_final_energy"])) < 1e-3
        basis = ref_basis                      # next round starts from the reference's basis


def check_hashed_csr(M, g):
    """large subspaces (make_golden.py hashed=True): indptr + SHA-256 of the column ids and of the
    float32 off-diagonal values + the reference's float32 diagonal"""
    import hashlib
    assert np.array_equal(M.indptr, g["H_indptr"]) and M.nnz == int(g["H_nnz"])
    assert hashlib.sha256(M.indices.astype(np.int32).tobytes()).hexdigest() == str(g["H_indices_sha256"])
    isdiag = M.indices == np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    assert hashlib.sha256(M.data[~isdiag].astype(np.float32).tobytes()).hexdigest() == str(g["H_offdiag_f32_sha256"])
    assert np.array_equal(M.data[~isdiag].astype(np.float32).astype(np.float64), M.data[~isdiag])
    d = M.data[isdiag]
    assert np.abs(d - g["H_diag32"]).max() < 2e-5 * max(1.0, np.abs(d).max())


def test_skqd_subspace_csr_and_time_evolution():
    for name in ("lih", "h5", "beh2_sto3g", "n2_sto3g"):
        g = load_golden("skqd_" + name)
        H = ham_of(g)
        sub = H.fci_basis()
        assert np.array_equal(sub, g["subspace"])
        M = H.raw_csr(sub)
        if "H_indices" not in g:            # hashed fixture: the oracle's matrix carries the evolution
            M.sort_indices()
            check_hashed_csr(M, g)
            psi = np.zeros(len(sub), np.complex128)
            psi[int(g["hf_index"])] = 1.0
            for step in range(3):
                psi = orc.expm_multiply_taylor(M.indptr.astype(np.int64), M.indices, M.data, psi, 0.1)
                assert np.abs(psi - g["psi_steps"][step]).max() < 2e-5     # float32 diagonal in the reference
            continue
        assert np.array_equal(M.indptr, g["H_indptr"])
        assert np.array_equal(M.indices, g["H_indices"])
        ref = g["H_data"]
        isdiag = M.indices == np.repeat(np.arange(len(sub)), np.diff(M.indptr))
        assert np.array_equal(M.data[~isdiag], ref[~isdiag])
        assert np.abs(M.data[isdiag] - ref[isdiag]).max() < 2e-5 * max(1.0, np.abs(ref).max())
        # time evolution on the REFERENCE's matrix: Taylor vs scipy expm_multiply
        psi = np.zeros(len(sub), np.complex128)
        psi[int(g["hf_index"])] = 1.0
        for step in range(3):
            psi = orc.expm_multiply_taylor(g["H_indptr"], g["H_indices"], ref, psi, 0.1)
            assert np.abs(psi - g["psi_steps"][step]).max() < 1e-12


@pytest.mark.parametrize("name", ["lih", "beh2_sto3g", "n2_sto3g"])
def test_ground_state_energy_quirk_F5(name):
    g = load_golden("skqd_" + name)
    H = ham_of(g)
    for tag in ("big", "small"):
        b = g[f"gse_{tag}_basis"]
        e_vec, v = H.ground_state_energy(b, True)
        e_no, _ = H.ground_state_energy(b, False)
        env = 5e-6 * max(1.0, abs(e_vec) / 8.0)          # float32 diagonals: relative envelope
        assert abs(e_vec - float(g[f"gse_{tag}_E_vec"])) < env
        assert abs(e_no - float(g[f"gse_{tag}_E_novec"])) < env
        # the reference's eigenvector is a ground vector of the oracle's matrix (Rayleigh quotient;
        # an overlap test would fail on the symmetry-degenerate N2 ground level)
        D = H.dense_H(b)
        S = 0.5 * (D + D.T)
        vr = g[f"gse_{tag}_v"].astype(np.float64)
        assert abs(vr @ S @ vr / (vr @ vr) + 1e-8 - e_vec) < env
        assert abs(v @ S @ v + 1e-8 - e_vec) < 1e-9
    assert float(g["gse_big_E_novec"]) > float(g["gse_big_E_vec"]) + 1e-3   # lambda_1, not lambda_0


@pytest.mark.parametrize("name", ["lih", "beh2_sto3g", "n2_sto3g"])
def test_skqd_energies_on_reference_samples(name):
    g = load_golden("skqd_" + name)
    H = ham_of(g)
    nf = g["nf_basis"]
    e_nf, _ = H.ground_state_energy(nf, False)
    env = 5e-6 * max(1.0, abs(e_nf) / 8.0)
    assert abs(e_nf - float(g["energy_nf_only"])) < env
    for k in range(1, int(g["kdim"])):
        kb = g[f"krylov_basis_{k}"]
        comb = orc.sort_unique(np.concatenate([nf, kb]))
        assert len(kb) == int(g["basis_sizes_krylov"][k - 1])
        assert len(comb) == int(g["basis_sizes_combined"][k - 1])
        e_k, _ = H.ground_state_energy(kb, False)
        e_c, _ = H.ground_state_energy(comb, False)
        assert abs(e_k - float(g["energies_krylov"][k - 1])) < env
        assert abs(e_c - float(g["energies_combined"][k - 1])) < env


def test_spmv_oracle_vs_numpy():
    g = load_golden("skqd_lih")
    rng = np.random.default_rng(0)
    n = len(g["H_indptr"]) - 1
    import scipy.sparse as sp
    M = sp.csr_matrix((g["H_data"], g["H_indices"], g["H_indptr"]), shape=(n, n))
    x = rng.standard_normal(n)
    assert np.allclose(orc.csr_matvec(g["H_indptr"], g["H_indices"], g["H_data"], x), M @ x,
                       rtol=0, atol=1e-12)
    z = x + 1j * rng.standard_normal(n)
    assert np.allclose(orc.csr_matvec(g["H_indptr"], g["H_indices"], g["H_data"], z), M @ z,
                       rtol=0, atol=1e-12)


def test_ground_state_energy_svd_fallback_branch():
    """oracle restatement of skqd.py:742-750 / :809-843: an ill-conditioned projected H (one
    eigenvalue shifted to ~0, regularization 0) goes through the SVD-regularised matrix; the ground
    energy is unchanged, the near-null mode is clamped to 1e-10 * s_max"""
    g = load_golden("skqd_lih")
    n_orb, na, nb = (int(x) for x in g["shape"])
    basis = g["gse_small_basis"]
    D = orc.OracleHam(g["h1"], g["g"], na, nb, 0.0).dense_H(basis)
    lam = np.linalg.eigvalsh(0.5 * (D + D.T))
    e_nuc = -float(lam[len(lam) // 2])
    H = orc.OracleHam(g["h1"], g["g"], na, nb, e_nuc)
    Ds = H.dense_H(basis)
    S = 0.5 * (Ds + Ds.T)
    assert np.linalg.cond(S) > 1e12
    e, v = H.ground_state_energy(basis, True, regularization=0.0)
    assert abs(e - (lam[0] + e_nuc)) < 1e-9
    assert abs(v @ S @ v - e) < 1e-9
    e2, none = H.ground_state_energy(basis, False, regularization=0.0)
    assert none is None and abs(e2 - e) < 1e-12
    # with the default regularisation the matrix is well conditioned again: plain branch
    e3, _ = H.ground_state_energy(basis, True)
    assert abs(e3 - (e + 1e-8)) < 1e-9
