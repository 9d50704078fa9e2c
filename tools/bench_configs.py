"""BASELINE.json configs[0..2] (LiH / BeH2 / N2 shapes) through the drop-in classes on one
B200: Stage-3 selected-CI rounds and Stage-4 SKQD, with the same inputs tools/ref_timings.py
gives the Python reference, plus parity of every energy against the FP64 oracle.
PySCF is not available offline: molecule-SHAPED synthetic integrals (same seeds)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_guided_krylov_b200 as fgk  # noqa: E402
from bench import synth_integrals  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from flow_guided_krylov_b200 import sto3g  # noqa: E402

REAL = "--real" in sys.argv      # real STO-3G integrals from the PySCF-free front-end (sto3g.py)
GEOM = {"lih": sto3g.lih_geometry, "beh2": sto3g.beh2_geometry, "n2": sto3g.n2_geometry}
SHAPES = {"lih": (6, 2, 2), "beh2": (7, 3, 3), "n2": (10, 7, 7)}


def sync():
    torch.cuda.synchronize()


def main():
    ref = {}
    p = os.path.join(ROOT, "profiles", "ref_cpu_timings_sto3g.json" if REAL else "ref_cpu_timings.json")
    if os.path.exists(p):
        ref = json.load(open(p))["configs"]
    for name, (n, na, nb) in SHAPES.items():
        if REAL:
            I = sto3g.compute_molecular_integrals(GEOM[name]())
            h1, g, e_nuc = I.h1e, I.h2e, I.nuclear_repulsion
        else:
            (h1, g), e_nuc = synth_integrals(n, seed=0), 0.0
        H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, e_nuc, na + nb, n, na, nb), "cuda:0")
        O = orc.OracleHam(h1.astype(np.float32), g.astype(np.float32), na, nb, e_nuc)
        k = {"lih": 150, "beh2": 200, "n2": 300}[name]
        ex = fgk.SelectedCIExpander(H, fgk.ResidualExpansionConfig(max_configs_per_iter=k))
        r = {}
        for rep in range(2):                      # rep 0 warms the kernels up
            b = H.get_hf_state().unsqueeze(0)
            sync(); t0 = time.perf_counter()
            energies = []
            for _ in range(3):
                b, st = ex.expand_basis(b)
                energies.append(st["final_energy"])
            sync(); r["expand_basis_3_rounds_s"] = time.perf_counter() - t0
        r["expand_basis_sizes_final"] = int(b.shape[0])
        r["expand_basis_energies"] = energies
        # parity of the three rounds against the FP64 oracle (same selections, 1e-9 Ha)
        ob = O.hf_state()[None]
        oe = []
        for _ in range(3):
            ob, ost = O.expand_basis(ob, k)
            oe.append(ost["final_energy"])
        r["oracle_energy_max_abs_diff"] = float(np.abs(np.array(oe) - np.array(energies)).max())
        r["oracle_basis_equal"] = bool(np.array_equal(ob, b.cpu().numpy().astype(np.uint8)))
        # Stage 4
        cfg = fgk.SKQDConfig(max_krylov_dim=3, shots_per_krylov=2000)
        for rep in range(2):
            sync(); t0 = time.perf_counter()
            sk = fgk.FlowGuidedSKQD(H, b, cfg)
            sync(); r["skqd_subspace_setup_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            P = sk._build_subspace_hamiltonian()
            sync(); r["subspace_H_build_s"] = time.perf_counter() - t0
            r["subspace_H_nnz"] = P.nnz
            torch.manual_seed(0)
            t0 = time.perf_counter()
            res = sk.run_with_nf(progress=False)
            sync(); r["run_with_nf_kdim3_s"] = time.perf_counter() - t0
        r["best_stable_energy"] = res["best_stable_energy"]
        # the energies of the sampled bases against the oracle on the same bases
        comb = sk.get_combined_basis(2).cpu().numpy().astype(np.uint8)
        e_o, _ = O.ground_state_energy(comb, False)
        r["oracle_combined_energy_abs_diff"] = abs(e_o - res["energies_combined"][-1])
        if name in ref:
            rr = ref[name]
            r["reference_cpu"] = {k2: rr[k2] for k2 in ("expand_basis_3_rounds_s", "subspace_H_build_s",
                                                        "run_with_nf_kdim3_s", "subspace_H_nnz")}
            r["speedup_expand_basis"] = rr["expand_basis_3_rounds_s"] / r["expand_basis_3_rounds_s"]
            r["speedup_subspace_build"] = rr["subspace_H_build_s"] / r["subspace_H_build_s"]
            r["speedup_run_with_nf"] = rr["run_with_nf_kdim3_s"] / r["run_with_nf_kdim3_s"]
            r["nnz_equal_reference"] = rr["subspace_H_nnz"] == r["subspace_H_nnz"]
            # Stage-3 energies vs the raw float32 reference: float32 envelope
            r["reference_energy_max_abs_diff"] = float(
                np.abs(np.array(rr["expand_basis_energies"]) - np.array(energies)).max())
        print(json.dumps({"config": name, "shape": [n, na, nb],
                          "integrals": "STO-3G" if REAL else "synthetic", **r}), flush=True)


if __name__ == "__main__":
    main()
