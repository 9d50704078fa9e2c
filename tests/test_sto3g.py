"""Integral front-end without PySCF (flow_guided_krylov_b200.sto3g): textbook STO-3G numbers, and the
one energy the reference publishes for its own molecular Hamiltonian.  CPU tier -- the GPU
counterpart (same energies through the engine) is test_gpu_parity.py::test_real_molecules_sto3g."""
import numpy as np
import pytest

from flow_guided_krylov_b200 import sto3g
from oracle import oracle as orc

# reference SKQD_VALIDATION_REPORT.md:87 -- "Exact (FCI)" of the reference's LiH Hamiltonian
# (create_lih_hamiltonian(): STO-3G, 1.6 Angstrom; its element rules, SURVEY F2/F3, float32)
LIH_PUBLISHED_FCI = -7.96379759


def test_tabulated_exponents_agree_with_scaled_slater_fits():
    for sym, shells in sto3g._TABLE.items():
        zeta = sto3g.ZETA[sym]
        for k, exps in enumerate(shells):
            fit = np.array(sto3g._A1S if k == 0 else sto3g._A2SP) * zeta[k] ** 2
            assert np.abs(fit / np.array(exps) - 1.0).max() < 5e-6, (sym, k)
    assert sto3g.shell_exponents("F")[1][0] == pytest.approx(0.994203 * 2.55 ** 2)


def test_h2_rhf_and_fci_textbook_values():
    I = sto3g.compute_molecular_integrals(sto3g.h2_geometry(0.74))
    assert (I.n_orbitals, I.n_electrons, I.n_alpha, I.n_beta) == (2, 2, 1, 1)
    assert abs(I.hf_energy - (-1.11675930739643)) < 1e-10          # PySCF RHF/STO-3G at 0.74 A
    h, g = I.h1e, I.h2e
    ci = np.array([[2 * h[0, 0] + g[0, 0, 0, 0], g[0, 1, 0, 1]], [g[0, 1, 0, 1], 2 * h[1, 1] + g[1, 1, 1, 1]]])
    assert abs(np.linalg.eigvalsh(ci)[0] + I.nuclear_repulsion - (-1.1372838344885)) < 1e-10


@pytest.mark.parametrize("name,geom,n_orb,n_el,e_hf", [
    ("lih", sto3g.lih_geometry(), 6, 4, -7.8618648),
    ("beh2", sto3g.beh2_geometry(), 7, 6, -15.5600984),
    ("n2", sto3g.n2_geometry(), 10, 14, -107.4965005),
])
def test_molecule_integrals_are_consistent(name, geom, n_orb, n_el, e_hf):
    S, T, V, eri, e_nuc, nel = sto3g.ao_integrals(geom)
    assert nel == n_el and len(S) == n_orb
    assert np.abs(np.diag(S) - 1.0).max() < 1e-12                  # normalised contractions
    I = sto3g.compute_molecular_integrals(geom)
    g = I.h2e
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)):        # 8-fold symmetry of real orbitals
        assert np.abs(g - g.transpose(perm)).max() < 1e-12
    assert abs(I.hf_energy - e_hf) < 5e-7                          # lowest SCF solution (N2: not the saddle)
    # the MO integrals reproduce the SCF energy: E = 2 sum_i h_ii + sum_ij (2 J_ij - K_ij) + E_nuc
    o = n_el // 2
    e = 2 * np.trace(I.h1e[:o, :o]) + e_nuc
    e += sum(2 * g[i, i, j, j] - g[i, j, j, i] for i in range(o) for j in range(o))
    assert abs(e - I.hf_energy) < 1e-9
    # Brillouin: occupied-virtual Fock elements vanish at convergence
    F = I.h1e + np.einsum("pqii->pq", g[:, :, :o, :o]) * 2 - np.einsum("piiq->pq", g[:, :o, :o, :])
    assert np.abs(F[:o, o:]).max() < 1e-7


def test_lih_reproduces_the_reference_published_energy():
    """The reference's element rules (restated by the oracle) on OUR integrals give the number its
    validation report prints for LiH -- an end-to-end pin that does not go through fixtures
    generated in this repository.  The oracle evaluates in FP64 on the float32-rounded integrals;
    the published value comes from the reference's float32 diagonals, hence the float32 envelope
    (measured gap 5.7e-7 Ha; the reference itself on our integrals: 3e-8, next test)."""
    I = sto3g.compute_molecular_integrals(sto3g.lih_geometry(1.6))
    O = orc.OracleHam(I.h1e.astype(np.float32), I.h2e.astype(np.float32), I.n_alpha, I.n_beta,
                      I.nuclear_repulsion)
    basis = O.fci_basis()
    assert len(basis) == 225
    e, _ = O.diagonalize(basis)
    assert abs(e - LIH_PUBLISHED_FCI) < 2e-6


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/src"),
                    reason="needs the reference checkout (build container only)")
def test_unmodified_reference_on_our_integrals_prints_its_published_energy():
    import sys
    sys.path.insert(0, "/root/reference/src")
    try:
        from hamiltonians.molecular import MolecularHamiltonian as RefH, MolecularIntegrals as RefI
    finally:
        sys.path.remove("/root/reference/src")
    I = sto3g.compute_molecular_integrals(sto3g.lih_geometry(1.6))
    R = RefH(RefI(I.h1e, I.h2e, I.nuclear_repulsion, I.n_electrons, I.n_orbitals, I.n_alpha, I.n_beta),
             device="cpu")
    assert abs(R.fci_energy() - LIH_PUBLISHED_FCI) < 1e-7
