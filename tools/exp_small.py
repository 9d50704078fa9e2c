"""Experiments on the small configurations (BASELINE configs[0..2], real STO-3G integrals):
(1) lowest eigenpair of the projected H: dense torch.linalg.eigh (cuSOLVER) against block Davidson
    over the engine's H.v, cold and warm-started, for growing selected-CI bases -- where does the
    iterative solver win (solvers.DENSE_EIG_MAX)?
(2) Davidson on configs[3] as a function of the subspace size (max_space).
(3) three selected-CI rounds per molecule with a per-phase split."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import sto3g, solvers
from bench import synth_integrals, cas_window_basis

dev = "cuda:0"


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


if "eig" in sys.argv or len(sys.argv) == 1:
    I = sto3g.compute_molecular_integrals(sto3g.n2_geometry())
    H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(I.h1e, I.h2e, I.nuclear_repulsion, 14, 10, 7, 7), dev)
    ex = fgk.SelectedCIExpander(H, fgk.ResidualExpansionConfig(max_configs_per_iter=300))
    b = H.get_hf_state().unsqueeze(0)
    prev = None
    for rd in range(9):
        b, st = ex.expand_basis(b)
        dets = H.pack(b)
        n = dets.shape[0]
        P = H.projected_csr(dets, fgk.H_SYM, packed=True)
        t_dense, (wd, vd) = timed(lambda: solvers.lowest_eigenpairs(P, k=1, dense_max=10 ** 9))
        t_cold, (wc, _) = timed(lambda: solvers.lowest_eigenpairs(P, k=1, dense_max=0))
        v0 = None
        if prev is not None:
            pos = fgk.BasisIndex(dets).lookup(prev[0]).long()
            v0 = torch.zeros(n, dtype=torch.float64, device=dev)
            v0[pos] = prev[1]
        t_warm, (ww, _) = timed(lambda: solvers.lowest_eigenpairs(P, k=1, dense_max=0, v0=v0)) if v0 is not None else (None, (wc, None))
        print(json.dumps({"exp": "eig", "n": n, "dense_ms": 1e3 * t_dense, "davidson_cold_ms": 1e3 * t_cold,
                          "davidson_warm_ms": None if t_warm is None else 1e3 * t_warm,
                          "diff_cold": abs(float(wd[0] - wc[0])), "diff_warm": abs(float(wd[0] - ww[0]))}), flush=True)
        prev = (dets, vd[:, 0].clone())

if "dav" in sys.argv or len(sys.argv) == 1:
    h1, g = synth_integrals(32, 0)
    H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
    dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
    P = H.projected_packed(dets, fgk.H_SYM, packed=True)
    for ms in (36, 48, 64, 96):
        calls = [0]

        def mv(v):
            calls[0] += 1
            return P.matvec(v)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        w, v = solvers.lowest_eigenpairs(P, k=1, tol=1e-9, matvec=mv, diagonal=P.diagonal(), dense_max=0, max_space=ms)
        torch.cuda.synchronize()
        print(json.dumps({"exp": "davidson_configs3", "max_space": ms, "seconds": time.perf_counter() - t0,
                          "matvecs": calls[0], "e0": float(w[0])}), flush=True)
    del P

if "rounds" in sys.argv or len(sys.argv) == 1:
    for name, geo, (n, na, nb), k in (("lih", sto3g.lih_geometry, (6, 2, 2), 150), ("beh2", sto3g.beh2_geometry, (7, 3, 3), 200),
                                      ("n2", sto3g.n2_geometry, (10, 7, 7), 300)):
        I = sto3g.compute_molecular_integrals(geo())
        H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(I.h1e, I.h2e, I.nuclear_repulsion, na + nb, n, na, nb), dev)
        for dm in (3072, 64, 0):
            solvers.DENSE_EIG_MAX = dm
            for rep in range(2):
                ex = fgk.SelectedCIExpander(H, fgk.ResidualExpansionConfig(max_configs_per_iter=k))
                b = H.get_hf_state().unsqueeze(0)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                es = []
                for _ in range(3):
                    b, st = ex.expand_basis(b)
                    es.append(st["final_energy"])
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(json.dumps({"exp": "rounds", "config": name, "dense_eig_max": dm, "three_rounds_ms": 1e3 * dt,
                              "final_size": int(b.shape[0]), "energies": es}), flush=True)
        solvers.DENSE_EIG_MAX = 3072
