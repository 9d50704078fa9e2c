"""Experiment: PT2 sweep at config-5 shape (48 orbitals, 12+12 electrons, CAS window basis of
C(11,4)^2 = 108,900 determinants, 270,648 connections per source).  Not part of the product."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from bench import synth_integrals, cas_window_basis

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cap = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
shape = int(sys.argv[3]) if len(sys.argv) > 3 else 5
if shape == 5:
    h1, g = synth_integrals(48, 0)
    H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 24, 48, 12, 12), "cuda:0")
    dets = torch.from_numpy(cas_window_basis(48, 8, 11, 4).view(np.int64)).cuda()
else:
    h1, g = synth_integrals(32, 0)
    H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), "cuda:0")
    dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).cuda()
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
coeff = torch.zeros(n, dtype=torch.float64, device="cuda")
ns = min(n_src, n)
perm = (torch.randperm(n, generator=torch.Generator().manual_seed(0))[:ns].cuda() if shape == 5
        else torch.arange(ns, device="cuda"))
coeff[perm] = torch.exp(-torch.arange(ns, dtype=torch.float64, device="cuda") / (0.25 * ns))
coeff /= torch.linalg.norm(coeff)
from flow_guided_krylov_b200.expansion import default_pt2_workspace
ns = min(n_src, n)
part = os.environ.get("FGK_PT2_PARTITION", "0") == "1"
ws = fgk.Pt2Workspace(cap, "cuda:0") if cap else default_pt2_workspace(H, ns, partition=part)
print("partition:", getattr(ws, "partition", None), "capacity", ws.capacity, flush=True)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    sel, imp, st = fgk.pt2_select(H, idx, coeff, -60.0, 500, workspace=ws)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"n={n} sources={st['n_sources']} raw={st['raw_candidates']:.4g} unique={st['unique_candidates']:.4g} "
          f"passes={st['passes']} time={dt:.3f}s  {st['raw_candidates']/dt:.3g} cand/s  top imp {float(imp[0]):.3e} "
          f"mem={torch.cuda.max_memory_allocated()/1e9:.1f}GB bps={os.environ.get('FGK_PT2_BPS', 'default')}", flush=True)
