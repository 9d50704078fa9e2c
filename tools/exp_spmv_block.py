"""Experiment: the SELL H.v kernel on a 1/N row block of configs[3] on ONE GPU (what a rank of an
N-GPU run executes, without the exchange): time per launch against full-size / N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from bench import synth_integrals, cas_window_basis
dev = "cuda:0"
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, 14, 4).view(np.int64)).to(dev)
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
x = torch.randn(n, dtype=torch.float64, device=dev)
for N in (1, 2, 4, 8):
    per = -(-n // N)
    P = H.projected_csr(dets, fgk.H_SYM, row_begin=0, row_end=per, index=idx, packed=True).to_sell(keep_csr=False)
    y = torch.empty(per, dtype=torch.float64, device=dev)
    for _ in range(5):
        P.matvec(x, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        P.matvec(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    # the one-launch step with this GPU as its only peer: per-CTA fence + counter + flag barrier, no NVLink
    from flow_guided_krylov_b200 import dist as fd
    fop = fd.FusedShardedOperator(P, storage="sell")
    fop.load(x)
    for _ in range(5):
        fop.step()
    e0.record()
    for _ in range(100):
        fop.step()
    e1.record(); torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / 100
    fop.close()
    print(f"unroll={os.environ.get('FGK_SPMV_UNROLL', '4')} N={N} rows={per} kernel_ms={ms:.4f} ms*N={ms * N:.4f} "
          f"peer_step_world1_ms={ms_step:.4f}", flush=True)
    del P, y, fop
