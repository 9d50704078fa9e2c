"""A few fused Davidson iterations on the configs[3] packed operator (for ncu -k regex:k_dav)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import solvers
from bench import synth_integrals, cas_window_basis
dev = "cuda:0"
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
P = H.projected_packed(dets, fgk.H_SYM, packed=True)
it = int(sys.argv[1]) if len(sys.argv) > 1 else 30
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
w, v = solvers._davidson_fused(solvers._LocalOp(P), 1, 1e-9, it, None, None, None)
torch.cuda.synchronize()
print("ok", float(w[0]), time.perf_counter() - t0)
