"""The direct packed build on configs[3] (for ncu -k regex:k_projh4|k_lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from bench import synth_integrals, cas_window_basis
dev = "cuda:0"
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
idx = fgk.BasisIndex(dets)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    P = H.projected_packed(dets, fgk.H_SYM, index=idx, packed=True, profile=True)
    print("ok", P.nnz, P.build_profile, flush=True)
    del P
