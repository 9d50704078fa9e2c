"""Time the REFERENCE itself (pure Python, CPU) on BASELINE.json configs[0..2] shapes.
Runs only where /root/reference exists (the build container); results are committed to
profiles/ref_cpu_timings.json and quoted next to the GPU numbers of tools/bench_configs.py.
PySCF is not available offline, so the molecules are molecule-SHAPED synthetic integrals
(SURVEY Appendix D generator + HF-like shift), same seeds as tools/bench_configs.py."""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hamiltonians.molecular import MolecularHamiltonian, MolecularIntegrals
from krylov.residual_expansion import SelectedCIExpander, ResidualExpansionConfig
from krylov.skqd import FlowGuidedSKQD, SKQDConfig
from bench import synth_integrals
from flow_guided_krylov_b200 import sto3g

REAL = "--real" in sys.argv      # real STO-3G integrals from the PySCF-free front-end (sto3g.py)
GEOM = {"lih": sto3g.lih_geometry, "beh2": sto3g.beh2_geometry, "n2": sto3g.n2_geometry}
SHAPES = {"lih": (6, 2, 2), "beh2": (7, 3, 3), "n2": (10, 7, 7)}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    torch.set_num_threads(os.cpu_count())
    out = {"host_cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
           "integrals": "STO-3G, RHF MOs (flow_guided_krylov_b200.sto3g)" if REAL else "synthetic, molecule-shaped",
           "configs": {}}
    for name, (n, na, nb) in SHAPES.items():
        if REAL:
            I = sto3g.compute_molecular_integrals(GEOM[name]())
            h1, g, e_nuc = I.h1e, I.h2e, I.nuclear_repulsion
        else:
            (h1, g), e_nuc = synth_integrals(n, seed=0), 0.0
        H = MolecularHamiltonian(MolecularIntegrals(h1, g, e_nuc, na + nb, n, na, nb), device="cpu")
        r = {}
        # Stage 3: three selected-CI rounds from the HF determinant
        k = {"lih": 150, "beh2": 200, "n2": 300}[name]
        ex = SelectedCIExpander(H, ResidualExpansionConfig(max_configs_per_iter=k))
        b = torch.stack([H.get_hf_state()])
        t0 = time.perf_counter()
        energies = []
        for _ in range(3):
            b, st = quiet(ex.expand_basis, b)
            energies.append(st["final_energy"])
        r["expand_basis_3_rounds_s"] = time.perf_counter() - t0
        r["expand_basis_sizes_final"] = int(len(b))
        r["expand_basis_energies"] = energies
        # Stage 4: subspace build + 2 time steps + projected solves
        t0 = time.perf_counter()
        sk = quiet(FlowGuidedSKQD, H, b, SKQDConfig(max_krylov_dim=3, shots_per_krylov=2000, use_gpu=False))
        r["skqd_subspace_setup_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        Hs = quiet(sk._build_subspace_hamiltonian)
        r["subspace_H_build_s"] = time.perf_counter() - t0
        r["subspace_H_nnz"] = int(Hs.nnz)
        sk._sparse_H = Hs
        torch.manual_seed(0)
        t0 = time.perf_counter()
        res = quiet(sk.run_with_nf, progress=False)
        r["run_with_nf_kdim3_s"] = time.perf_counter() - t0
        r["best_stable_energy"] = res["best_stable_energy"]
        print(name, json.dumps(r), flush=True)
        out["configs"][name] = r
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                     "ref_cpu_timings_sto3g.json" if REAL else "ref_cpu_timings.json")
    json.dump(out, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
