#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE itself (CPU, this container).

    python tests/golden/make_golden.py               # synthetic integrals -> tests/golden/*.npz
    python tests/golden/make_golden.py --molecules   # real LiH / BeH2 / N2 STO-3G integrals
                                                     # (flow_guided_krylov_b200.sto3g) -> *_sto3g.npz
    python tests/golden/make_golden.py --molecules2 [sci_n2] [skqd_beh2] [skqd_n2]
                                                     # Stage 3 on N2, Stage 4 on BeH2 / N2 (minutes)

The reference (/root/reference, pure Python) is imported, never copied.  It
cannot travel to the GPU box, so its outputs on seeded synthetic integrals
(SURVEY.md Appendix D generator) are committed here as small .npz fixtures.
They pin (a) the oracle restatement (tests/test_oracle_golden.py, CPU) and (b)
the CUDA path (tests/test_gpu_parity.py, -m gpu).

Reference functions exercised (file:line under /root/reference/src):
  hamiltonians/molecular.py:133-184   diagonal_elements_batch
  hamiltonians/molecular.py:194-327   get_connections
  hamiltonians/molecular.py:471-516   matrix_elements_fast
  hamiltonians/molecular.py:580-638   get_sparse_matrix_elements
  hamiltonians/molecular.py:872-942   fci_energy
  krylov/residual_expansion.py:334-554 SelectedCIExpander
  krylov/residual_expansion.py:60-257  ResidualBasedExpander
  krylov/skqd.py:135-177,374-419      subspace + subspace CSR
  krylov/skqd.py:275-296              expm_multiply time evolution
  krylov/skqd.py:683-807              compute_ground_state_energy
  krylov/skqd.py:946-1059             FlowGuidedSKQD.run_with_nf
"""
import contextlib
import io
import os
import sys
from itertools import combinations

import numpy as np
import torch

REF = os.environ.get("FGK_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
from hamiltonians.molecular import MolecularHamiltonian, MolecularIntegrals  # noqa: E402
from krylov.residual_expansion import (  # noqa: E402
    ResidualBasedExpander, ResidualExpansionConfig, SelectedCIExpander)
from krylov.skqd import FlowGuidedSKQD, SKQDConfig  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synth_arrays(n_orb, seed=0, h1_scale=1.0, h2_scale=0.1, sparsify=0.0):
    """SURVEY.md Appendix D generator; `sparsify` zeroes a symmetric fraction of
    entries so the reference's 1e-12 filters (molecular.py:106,266,288,308) bite."""
    rng = np.random.default_rng(seed)
    h1 = rng.standard_normal((n_orb, n_orb)) * h1_scale
    h1 = 0.5 * (h1 + h1.T)
    g = rng.standard_normal((n_orb,) * 4) * h2_scale
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    if sparsify > 0:
        m1 = rng.random((n_orb, n_orb)) < sparsify
        m1 = m1 | m1.T
        np.fill_diagonal(m1, False)
        h1 = np.where(m1, 0.0, h1)
        m2 = rng.random((n_orb,) * 4) < sparsify
        m2 = m2 | m2.transpose(1, 0, 2, 3)
        m2 = m2 | m2.transpose(0, 1, 3, 2)
        m2 = m2 | m2.transpose(2, 3, 0, 1)
        g = np.where(m2, 0.0, g)
    # the reference keeps float32 tables only (molecular.py:68-69); the fixtures
    # store exactly those, so round-trip through float32 here
    return h1.astype(np.float32).astype(np.float64), g.astype(np.float32).astype(np.float64)


MOLECULES = {}      # name -> (h1, g, e_nuc): real STO-3G integrals, float32-rounded


def molecule_arrays(molecule):
    """Real STO-3G / RHF integrals of the reference's own molecules (its factories' default
    geometries, molecular.py:1019-1115) from the repository's PySCF-free front-end."""
    if molecule not in MOLECULES:
        sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
        from flow_guided_krylov_b200 import sto3g
        geom = {"lih": sto3g.lih_geometry, "beh2": sto3g.beh2_geometry, "n2": sto3g.n2_geometry}[molecule]()
        I = sto3g.compute_molecular_integrals(geom)
        MOLECULES[molecule] = (I.h1e.astype(np.float32).astype(np.float64),
                               I.h2e.astype(np.float32).astype(np.float64), float(I.nuclear_repulsion))
    return MOLECULES[molecule]


def make_h(n_orb, na, nb, seed=0, e_nuc=0.0, molecule=None, **kw):
    if molecule is not None:
        h1, g, e_nuc = molecule_arrays(molecule)
        assert h1.shape == (n_orb, n_orb)
    else:
        h1, g = synth_arrays(n_orb, seed, **kw)
    integ = MolecularIntegrals(h1, g, e_nuc, na + nb, n_orb, na, nb)
    return MolecularHamiltonian(integ, device="cpu"), h1, g


def fci_basis(n_orb, na, nb):
    """combinations order, alpha-major (skqd.py:155-169, molecular.py:894-905)."""
    rows = []
    for a in combinations(range(n_orb), na):
        for b in combinations(range(n_orb), nb):
            c = np.zeros(2 * n_orb, dtype=np.int64)
            c[list(a)] = 1
            c[[i + n_orb for i in b]] = 1
            rows.append(c)
    return np.stack(rows)


def random_dets(n_orb, na, nb, n, rng):
    out = np.zeros((n, 2 * n_orb), dtype=np.int64)
    for i in range(n):
        out[i, rng.choice(n_orb, na, replace=False)] = 1
        out[i, n_orb + rng.choice(n_orb, nb, replace=False)] = 1
    return out


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def connections_block(H, dets):
    offs, cfgs, els = [0], [], []
    for d in dets:
        c, e = H.get_connections(torch.from_numpy(d))
        offs.append(offs[-1] + len(c))
        if len(c):
            cfgs.append(c.numpy().astype(np.uint8))
            els.append(e.numpy().astype(np.float32))
    S = dets.shape[1]
    cfgs = np.concatenate(cfgs) if cfgs else np.zeros((0, S), np.uint8)
    els = np.concatenate(els) if els else np.zeros((0,), np.float32)
    return np.array(offs, np.int64), cfgs, els


def gen_hamiltonian_case(name, n_orb, na, nb, seed, n_rand, n_basis, e_nuc=0.0, **kw):
    H, h1, g = make_h(n_orb, na, nb, seed, e_nuc=e_nuc, **kw)
    e_nuc = float(H.nuclear_repulsion) if hasattr(H, "nuclear_repulsion") else e_nuc
    rng = np.random.default_rng(1000 + seed)
    hf = H.get_hf_state().numpy()
    dets = np.concatenate([hf[None], random_dets(n_orb, na, nb, n_rand, rng)])
    offs, ccfg, cel = connections_block(H, dets)
    diag32 = H.diagonal_elements_batch(torch.from_numpy(dets)).numpy()
    # projected H over a basis without duplicates (torch.unique => sorted)
    basis = torch.unique(torch.from_numpy(
        np.concatenate([hf[None], random_dets(n_orb, na, nb, n_basis, rng)])), dim=0)
    Hd = H.matrix_elements_fast(basis).numpy()
    r, c, v = H.get_sparse_matrix_elements(basis)
    np.savez_compressed(
        os.path.join(OUT, f"ham_{name}.npz"),
        shape=np.array([n_orb, na, nb]), e_nuc=np.array(e_nuc),
        h1=h1.astype(np.float32), g=g.astype(np.float32), dets=dets.astype(np.uint8), conn_offsets=offs, conn_cfgs=ccfg,
        conn_elems=cel, diag32=diag32, basis=basis.numpy().astype(np.uint8),
        H_dense32=Hd, coo_rows=r.numpy(), coo_cols=c.numpy(), coo_vals=v.numpy())
    print(f"ham_{name}: {len(dets)} dets, {offs[-1]} connections, basis {len(basis)}")
    return H


def gen_expander_case(name, n_orb, na, nb, seed, k, rounds, start="hf", molecule=None):
    H, h1, g = make_h(n_orb, na, nb, seed, molecule=molecule)
    hf = H.get_hf_state()
    if start == "hf":
        b = torch.stack([hf])
    else:
        rng = np.random.default_rng(77)
        b = torch.unique(torch.from_numpy(
            np.concatenate([hf.numpy()[None], random_dets(n_orb, na, nb, start, rng)])), dim=0)
    out = dict(shape=np.array([n_orb, na, nb]), h1=h1.astype(np.float32), g=g.astype(np.float32), k=np.array(k),
               basis0=b.numpy().astype(np.uint8), e_nuc=np.array(float(getattr(H, "nuclear_repulsion", 0.0))))
    ex = SelectedCIExpander(H, ResidualExpansionConfig(max_configs_per_iter=k))
    for rd in range(rounds):
        # the inner selection, on the reference's own eigenpair
        E, v = ex._diagonalize(b)
        sel, imp = ex._find_important_configs(b, E, v)
        out[f"r{rd}_E"] = np.array(E)
        out[f"r{rd}_v"] = np.asarray(v, dtype=np.float64)
        out[f"r{rd}_sel"] = sel.numpy().astype(np.uint8)
        out[f"r{rd}_imp"] = imp.numpy().astype(np.float32)
        b, st = ex.expand_basis(b)
        out[f"r{rd}_basis"] = b.numpy().astype(np.uint8)
        out[f"r{rd}_final_energy"] = np.array(st["final_energy"])
        out[f"r{rd}_configs_added"] = np.array(st["configs_added"])
        print(f"  sci_{name} round {rd}: size {len(b)} E {st['final_energy']:.9f}")
    out["rounds"] = np.array(rounds)
    np.savez_compressed(os.path.join(OUT, f"sci_{name}.npz"), **out)


def gen_residual_case(name, n_orb, na, nb, seed, k, iters):
    H, h1, g = make_h(n_orb, na, nb, seed)
    hf = H.get_hf_state()
    ex = ResidualBasedExpander(H, ResidualExpansionConfig(
        max_configs_per_iter=k, max_iterations=iters, residual_threshold=1e-4))
    b, st = ex.expand_basis(torch.stack([hf]))
    np.savez_compressed(
        os.path.join(OUT, f"res_{name}.npz"), shape=np.array([n_orb, na, nb]), h1=h1.astype(np.float32), g=g.astype(np.float32),
        k=np.array(k), iters=np.array(iters), basis=b.numpy().astype(np.uint8),
        energies=np.array(st["history"]["energies"]),
        sizes=np.array(st["history"]["basis_sizes"]),
        final_energy=np.array(st["final_energy"]))
    print(f"res_{name}: size {len(b)} E {st['final_energy']:.9f}")


def gen_skqd_case(name, n_orb, na, nb, seed, h2_scale, kdim, shots, n_nf, molecule=None, hashed=False):
    """hashed=True (large subspaces): the subspace CSR is stored as indptr + SHA-256 digests of
    the column ids and of the float32 off-diagonal values (the arrays themselves would be tens
    of MB), the diagonal as the reference's float32 values."""
    if molecule is not None:
        H, h1, g = make_h(n_orb, na, nb, seed, molecule=molecule)
    else:
        H, h1, g = make_h(n_orb, na, nb, seed, h2_scale=h2_scale)
    rng = np.random.default_rng(5)
    hf = H.get_hf_state().numpy()
    nf = torch.unique(torch.from_numpy(
        np.concatenate([hf[None], random_dets(n_orb, na, nb, n_nf, rng)])), dim=0)
    cfg = SKQDConfig(max_krylov_dim=kdim, time_step=0.1, shots_per_krylov=shots,
                     use_gpu=False)
    sk = quiet(FlowGuidedSKQD, H, nf, cfg)
    sub = sk._subspace_basis.numpy().astype(np.uint8)
    Hs = quiet(sk._build_subspace_hamiltonian)
    Hs.sort_indices()
    sk._sparse_H = Hs
    # three exact time steps from the HF unit vector in the subspace
    from scipy.sparse.linalg import expm_multiply
    psi = np.zeros(len(sub), dtype=np.complex128)
    hf_idx = int(np.where((sub == hf.astype(np.uint8)).all(1))[0][0])
    psi[hf_idx] = 1.0
    psis = []
    for _ in range(3):
        psi = expm_multiply(-1j * sk.time_step * Hs, psi)
        psis.append(psi.copy())
    torch.manual_seed(0)
    res = quiet(sk.run_with_nf, progress=False)
    # cumulative Krylov bases exactly as the reference formed them
    out = dict(shape=np.array([n_orb, na, nb]), h1=h1.astype(np.float32), g=g.astype(np.float32), kdim=np.array(kdim),
               shots=np.array(shots), nf_basis=nf.numpy().astype(np.uint8),
               subspace=sub, H_indptr=Hs.indptr.astype(np.int64),
               H_imag_max=np.array(np.abs(Hs.data.imag).max()),
               e_nuc=np.array(float(getattr(H, "nuclear_repulsion", 0.0))),
               psi_steps=np.stack(psis), hf_index=np.array(hf_idx),
               energy_nf_only=np.array(res["energy_nf_only"]),
               energies_krylov=np.array(res["energies_krylov"]),
               energies_combined=np.array(res["energies_combined"]),
               basis_sizes_krylov=np.array(res["basis_sizes_krylov"]),
               basis_sizes_combined=np.array(res["basis_sizes_combined"]),
               best_stable_energy=np.array(res["best_stable_energy"]))
    if hashed:
        import hashlib
        rows = np.repeat(np.arange(len(sub)), np.diff(Hs.indptr))
        offd = Hs.indices != rows
        out["H_indices_sha256"] = np.array(hashlib.sha256(Hs.indices.astype(np.int32).tobytes()).hexdigest())
        out["H_offdiag_f32_sha256"] = np.array(
            hashlib.sha256(Hs.data.real[offd].astype(np.float32).tobytes()).hexdigest())
        out["H_diag32"] = Hs.data.real[~offd].astype(np.float32)
        out["H_nnz"] = np.array(Hs.nnz)
    else:
        out["H_indices"] = Hs.indices.astype(np.int32)
        out["H_data"] = Hs.data.real.astype(np.float64)
    for k in range(kdim):
        out[f"krylov_basis_{k}"] = sk.get_basis_states(k).numpy().astype(np.uint8)
    # ground-state solver quirks (F5): both return modes on a >=100 and a <100 basis
    big = torch.from_numpy(sub[:150].astype(np.int64))
    small = torch.from_numpy(sub[:60].astype(np.int64))
    for tag, bs in (("big", big), ("small", small)):
        e_t, v_t = quiet(sk.compute_ground_state_energy, bs, True, 1e-8)
        e_f, _ = quiet(sk.compute_ground_state_energy, bs, False, 1e-8)
        out[f"gse_{tag}_basis"] = bs.numpy().astype(np.uint8)
        out[f"gse_{tag}_E_vec"] = np.array(e_t)
        out[f"gse_{tag}_E_novec"] = np.array(e_f)
        out[f"gse_{tag}_v"] = v_t.numpy()
    np.savez_compressed(os.path.join(OUT, f"skqd_{name}.npz"), **out)
    print(f"skqd_{name}: subspace {len(sub)} nnz {Hs.nnz} E_nf {res['energy_nf_only']:.9f} "
          f"best {res['best_stable_energy']:.9f}")


def gen_fci(name, n_orb, na, nb, seed, molecule=None):
    H, h1, g = make_h(n_orb, na, nb, seed, molecule=molecule)
    e = quiet(H.fci_energy)
    np.savez_compressed(os.path.join(OUT, f"fci_{name}.npz"),
                        shape=np.array([n_orb, na, nb]), h1=h1.astype(np.float32), g=g.astype(np.float32), fci=np.array(e),
                        e_nuc=np.array(float(getattr(H, "nuclear_repulsion", 0.0))))
    print(f"fci_{name}: {e:.10f}")


def main_molecules():
    """fixtures on REAL STO-3G integrals (symmetry zeros, numerical-noise entries around the
    1e-12 filters, non-zero nuclear repulsion); kept separate so that the synthetic fixtures
    of the first generation stay byte-identical"""
    torch.set_num_threads(1)
    gen_hamiltonian_case("lih_sto3g", 6, 2, 2, 0, n_rand=24, n_basis=60, molecule="lih")
    gen_hamiltonian_case("beh2_sto3g", 7, 3, 3, 0, n_rand=16, n_basis=80, molecule="beh2")
    gen_hamiltonian_case("n2_sto3g", 10, 7, 7, 0, n_rand=6, n_basis=120, molecule="n2")
    gen_fci("lih_sto3g", 6, 2, 2, 0, molecule="lih")
    gen_expander_case("lih_sto3g", 6, 2, 2, 0, k=10, rounds=3, molecule="lih")
    gen_expander_case("beh2_sto3g", 7, 3, 3, 0, k=20, rounds=3, molecule="beh2")


def main_molecules2():
    """round-2 fixtures: Stage 3 on N2 and Stage 4 (SKQD) on BeH2 / N2, real STO-3G integrals
    (residual_expansion.py:334-406, skqd.py:946-1059)"""
    torch.set_num_threads(1)
    which = [a for a in sys.argv[1:] if not a.startswith("--")]
    if not which or "sci_n2" in which:
        gen_expander_case("n2_sto3g", 10, 7, 7, 0, k=100, rounds=2, start=250, molecule="n2")
    if not which or "skqd_beh2" in which:
        gen_skqd_case("beh2_sto3g", 7, 3, 3, 0, h2_scale=None, kdim=4, shots=5000, n_nf=100, molecule="beh2")
    if not which or "skqd_n2" in which:
        gen_skqd_case("n2_sto3g", 10, 7, 7, 0, h2_scale=None, kdim=3, shots=3000, n_nf=150, molecule="n2",
                      hashed=True)


def main():
    torch.set_num_threads(1)
    # SURVEY Appendix B shapes + ragged / sparse / edge shapes
    gen_hamiltonian_case("lih", 6, 2, 2, 0, n_rand=24, n_basis=60)
    gen_hamiltonian_case("beh2", 7, 3, 3, 0, n_rand=16, n_basis=80)
    gen_hamiltonian_case("n2", 10, 7, 7, 0, n_rand=6, n_basis=120)
    gen_hamiltonian_case("ragged", 5, 2, 3, 3, n_rand=20, n_basis=40, e_nuc=1.25)
    gen_hamiltonian_case("sparse", 6, 3, 2, 4, n_rand=20, n_basis=50, sparsify=0.35)
    gen_hamiltonian_case("edge_full_alpha", 4, 4, 1, 5, n_rand=3, n_basis=3)
    gen_hamiltonian_case("edge_no_beta", 4, 2, 0, 6, n_rand=5, n_basis=5)
    gen_hamiltonian_case("wide", 33, 3, 2, 7, n_rand=2, n_basis=30, sparsify=0.25)   # > 64 sites (F6/Q8)
    gen_fci("lih", 6, 2, 2, 0)
    gen_fci("beh2", 7, 3, 3, 0)
    gen_expander_case("lih", 6, 2, 2, 0, k=10, rounds=3)
    gen_expander_case("beh2", 7, 3, 3, 0, k=10, rounds=3)
    gen_expander_case("beh2_wide", 7, 3, 3, 1, k=40, rounds=2, start=30)
    gen_residual_case("lih", 6, 2, 2, 0, k=10, iters=3)
    gen_skqd_case("lih", 6, 2, 2, 0, h2_scale=0.1, kdim=4, shots=2000, n_nf=40)
    gen_skqd_case("h5", 5, 2, 2, 2, h2_scale=0.05, kdim=3, shots=1000, n_nf=25)


if __name__ == "__main__":
    if "--molecules2" in sys.argv:
        main_molecules2()
    elif "--molecules" in sys.argv:
        main_molecules()
    else:
        main()
