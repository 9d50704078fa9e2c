"""One short pass over the SHIPPED kernels at bench size, for ncu (-k regex filters pick the
kernels): index + projected-H build (k_projh3 count / fill at 1,002,001 rows), packed SELL-32
H.v (k_spmv_sell_f32), FP64 SELL H.v (k_spmv_sell), the fused multi-GPU step kernel with itself as
the only peer (k_spmv_sell_bcast, world = 1), a PT2 selection sweep (k_pt2_accumulate2,
k_pt2_score, k_pt2_gather).  Small mode (--small) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import dist as fdist
from flow_guided_krylov_b200.expansion import pt2_select, default_pt2_workspace
from bench import synth_integrals, cas_window_basis

small = "--small" in sys.argv
dev = "cuda:0"
n_orb, na, nfz, nact, nel = (16, 4, 1, 8, 3) if small else (32, 8, 4, 14, 4)
h1, g = synth_integrals(n_orb, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 2 * na, n_orb, na, na), dev)
dets = torch.from_numpy(cas_window_basis(n_orb, nfz, nact, nel).view(np.int64)).to(dev)
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
P = H.projected_csr(dets, fgk.H_SYM, index=idx, packed=True)
P.to_sell()
x = torch.randn(n, dtype=torch.float64, device=dev)
y0 = P.matvec(x, fmt="sell")
P.to_sell_packed()
y1 = P.matvec(x, fmt="packed")
z = torch.complex(x, x.flip(0))
P.matvec(z, fmt="packed")
sellf = P._sellf
P._sellf = None
fop = fdist.FusedShardedOperator(P)
y2 = fop.matvec(x)
fop.check()
fop.close()
P._sellf = sellf
assert float((y0 - y1).abs().max()) < 1e-9 and float((y0 - y2).abs().max()) < 1e-9
ns = min(256 if small else 2048, n)
coeff = torch.zeros(n, dtype=torch.float64, device=dev)
coeff[:ns] = torch.exp(-torch.arange(ns, dtype=torch.float64, device=dev) / (0.25 * ns))
coeff /= torch.linalg.norm(coeff)
ws = default_pt2_workspace(H, ns)
sel, imp, st = pt2_select(H, idx, coeff, -30.0, 500, workspace=ws)
ws_small = fgk.Pt2Workspace(max(4096, st["unique_candidates"] // 5), dev)      # multi-pass + contention
sel2, imp2, st2 = pt2_select(H, idx, coeff, -30.0, 500, workspace=ws_small)
assert torch.equal(sel, sel2) and torch.equal(imp, imp2), "pass-count invariance (exact accumulation)"
torch.cuda.synchronize()
print(f"ok n={n} nnz={P.nnz} pt2 raw={st['raw_candidates']} unique={st['unique_candidates']} passes2={st2['passes']}")
