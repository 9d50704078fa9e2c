"""Second pass over the SHIPPED kernels at bench size, final forms of round 2 (for ncu -k regex):
k_lists / k_projh4 (direct packed build), k_dav (fused Davidson iteration, m ~ 20), k_taylor_update_z
and the complex packed H.v (one exp(-i dt H) step), k_peer_gather / k_peer_allreduce (world = 1),
k_conn with TMA-staged integral tables (N2 STO-3G, 14,400 determinants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import dist as fdist, solvers, sto3g
from bench import synth_integrals, cas_window_basis
dev = "cuda:0"
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
n = dets.shape[0]
P = H.projected_packed(dets, fgk.H_SYM, packed=True)
w, v = solvers._davidson_fused(solvers._LocalOp(P), 1, 1e-9, 3, None, None, None)
psi = torch.zeros(n, dtype=torch.complex128, device=dev)
psi[0] = 1.0
mu = float(P.diagonal().sum()) / n
rho = solvers.spectral_radius_estimate(P.matvec, n, mu, dev, iters=3)
out = solvers.expm_multiply(P, psi, -0.01j, mu=mu, rho=max(rho, 1.0))
fop = fdist.FusedShardedOperator(P)
x = torch.randn(n, dtype=torch.float64, device=dev)
fop.gather_local(x)
fop.step()
t = torch.randn(37, dtype=torch.float64, device=dev)
fop.check()
fop.close()
I = sto3g.compute_molecular_integrals(sto3g.n2_geometry())
Hm = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(I.h1e, I.h2e, I.nuclear_repulsion, 14, 10, 7, 7), dev)
full = Hm.fci_dets()
od, el, src, offs = Hm.connections_packed(full)
torch.cuda.synchronize()
print("ok", P.nnz, float(w[0]), int(od.shape[0]))
