"""Experiment (python or torchrun): Stage-3 PT2 selection at BASELINE configs[4] shape
(48 orbitals, 12+12 electrons, 108,900-determinant CAS basis, 270,648 connections/source)
on N GPUs: sources dealt round-robin, dedup exchange by owner, global top-k merge."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import dist as fd
from bench import synth_integrals, cas_window_basis

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
rank = dist.get_rank() if world > 1 else 0
n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
h1, g = synth_integrals(48, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 24, 48, 12, 12), dev)
dets = torch.from_numpy(cas_window_basis(48, 8, 11, 4).view(np.int64)).to(dev)
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
ns = min(n_src, n)
coeff = torch.zeros(n, dtype=torch.float64, device=dev)
perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))[:ns].to(dev)
coeff[perm] = torch.exp(-torch.arange(ns, dtype=torch.float64, device=dev) / (0.25 * ns))
coeff /= torch.linalg.norm(coeff)
for rep in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.time()
    sel, sc, st = fd.pt2_select_sharded(H, idx, coeff, -60.0, 500)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.time() - t0
    if rank == 0:
        print(f"world={world} sources={ns} raw={st['raw_candidates_total']:.4g} unique={st['unique_total']:.4g} "
              f"passes={st['passes']} time={dt:.3f}s {st['raw_candidates_total']/dt:.3g} cand/s "
              f"sel0={sel[0].tolist()} imp0={float(sc[0]):.6e} mem={torch.cuda.max_memory_allocated()/1e9:.0f}GB", flush=True)
if world > 1:
    dist.destroy_process_group()
