"""Experiment: PT2 selection sweep (bench shape: configs[3] basis, 2,048 sources, 1.08e8 raw
candidates) as a function of the workspace capacity -- smaller workspaces mean more bucket
passes, each with an accumulator that fits further into the 126 MB L2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200.expansion import pt2_select, Pt2Workspace
from bench import synth_integrals, cas_window_basis

dev = "cuda:0"
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
ns = 2048
coeff = torch.zeros(n, dtype=torch.float64, device=dev)
coeff[:ns] = torch.exp(-torch.arange(ns, dtype=torch.float64, device=dev) / (0.25 * ns))
coeff /= torch.linalg.norm(coeff)
ref = None
for cap in [int(x) for x in (sys.argv[1:] or ["114000000", "57000000", "45000000", "12000000", "6000000", "3000000", "1500000", "800000"])]:
    ws = Pt2Workspace(cap, dev)
    sel, imp, st = pt2_select(H, idx, coeff, -30.0, 500, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sel, imp, st = pt2_select(H, idx, coeff, -30.0, 500, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if ref is None:
        ref = (sel.clone(), imp.clone())
    same = bool(torch.equal(sel, ref[0])) and bool(torch.equal(imp, ref[1]))
    print(f"capacity={cap} table={8 * ws._table.numel() / 2**20:.0f}MiB pool={32 * cap / 2**20:.0f}MiB passes={st['passes']} "
          f"raw={st['raw_candidates']:.4g} unique={st['unique_candidates']:.4g} ms={ms:.2f} "
          f"cand/s={st['raw_candidates'] / ms * 1e3:.3g} identical_to_first={same}", flush=True)
    del ws
