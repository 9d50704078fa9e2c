"""Experiment: SpMV variants on the config-4 matrix (CSR-vector vs SELL-32, sorted vs
enumeration-order rows).  Not part of the product; prints one line per variant."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flow_guided_krylov_b200 as fgk
from bench import synth_integrals, cas_window_basis

n_active = int(sys.argv[1]) if len(sys.argv) > 1 else 14
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), "cuda:0")
dets = torch.from_numpy(cas_window_basis(32, 4, n_active, 4).view(np.int64)).cuda()
n = dets.shape[0]
idx = fgk.BasisIndex(dets)
x = torch.randn(n, dtype=torch.float64, device="cuda")
z = torch.complex(x, torch.randn_like(x))

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for sort in (False, True):
    t0 = time.time()
    P = H.projected_csr(dets, fgk.H_SYM, index=idx, packed=True, sort_rows=False)
    torch.cuda.synchronize(); t1 = time.time()
    if sort:
        P.sort_rows(); torch.cuda.synchronize()
    t2 = time.time()
    b = P.bytes_per_matvec()
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: P.matvec(x, out=y))
    print(f"sorted={sort} build {t1-t0:.2f}s sort {t2-t1:.2f}s  CSR-vector  {ms:.3f} ms  {b/ms/1e6:.0f} GB/s", flush=True)
    yref = y.clone()
    t3 = time.time(); P.to_sell(); torch.cuda.synchronize(); t4 = time.time()
    ms = timeit(lambda: P.matvec(x, out=y))
    print(f"sorted={sort} to_sell {t4-t3:.2f}s            SELL-32     {ms:.3f} ms  {b/ms/1e6:.0f} GB/s  maxdiff {float((y-yref).abs().max()):.2e}", flush=True)
    yz = torch.empty(n, dtype=torch.complex128, device="cuda")
    bz = P.bytes_per_matvec(True)
    ms = timeit(lambda: P.matvec(z, out=yz))
    print(f"sorted={sort}                         SELL-32 cplx {ms:.3f} ms  {bz/ms/1e6:.0f} GB/s", flush=True)
    ms = timeit(lambda: P.matvec(z, out=yz, fmt='csr'))
    print(f"sorted={sort}                         CSR cplx     {ms:.3f} ms  {bz/ms/1e6:.0f} GB/s", flush=True)
    del P, y, yz
    torch.cuda.empty_cache()
