"""Experiment: irregular basis (SURVEY 8d "irregular variant").  Excitation-level-truncated CI
inside the config-4 CAS window: all determinants at most L excitations away from the HF
determinant (L = 5: 412,501 determinants).  Row lengths of the projected H vary by a factor
of several between excitation levels, so this exercises SELL-32 padding and the builder's
load balance; the CAS product basis of the headline has identical rows.  Not part of the product."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import flow_guided_krylov_b200 as fgk
from bench import synth_integrals, cas_window_strings

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n_orb, na, nb = 32, 8, 8
h1, g = synth_integrals(n_orb, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, n_orb, na, nb), "cuda:0")
s = cas_window_strings(n_orb, 4, 14, 4)
hf = s.max()                       # orbitals 0..7 occupied = highest bits
lvl = np.array([bin(int(x) ^ int(hf)).count("1") // 2 for x in s])
ia, ib = np.meshgrid(np.arange(len(s)), np.arange(len(s)), indexing="ij")
keep = (lvl[ia] + lvl[ib]) <= L
dnp = np.stack([s[ia[keep]], s[ib[keep]]], axis=1)
dets = torch.from_numpy(dnp.view(np.int64)).cuda()
n = dets.shape[0]
out = {"basis": f"excitation level <= {L} inside CAS(8e,14o), 32 orbitals", "n_dets": n}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = fgk.BasisIndex(dets)
    torch.cuda.synchronize(); t_index = time.perf_counter() - t0
    P = H.projected_csr(dets, fgk.H_SYM, packed=True, index=idx, profile=True)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
out["index_s"] = t_index
lens = (P.row_ptr[1:] - P.row_ptr[:-1]).double()
out.update(nnz=P.nnz, dense_pairs=idx.info()["dense_pairs"], build_s=t_build, build_kernels=P.build_profile,
           row_len_min=float(lens.min()), row_len_mean=float(lens.mean()), row_len_max=float(lens.max()))
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")


def timed(fmt, reps=30):
    for _ in range(3):
        P.matvec(x, out=y, fmt=fmt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        P.matvec(x, out=y, fmt=fmt)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_csr = timed("csr")
yc = y.clone()
P.to_sell()
ms_sell = timed("sell")
assert float((y - yc).abs().max()) < 1e-9
padded = int(P._sell[1].numel())
P.to_sell_packed()
ms_pack = timed("packed")
assert float((y - yc).abs().max()) < 1e-9
alg = 12.0 * P.nnz + 20.0 * n
out.update(csr_ms=ms_csr, sell_ms=ms_sell, packed_ms=ms_pack, sell_padding=padded / P.nnz - 1.0,
           sell_algorithmic_GBs=alg / ms_sell / 1e6, packed_stored_GBs=(8.0 * P.nnz + 28.0 * n) / ms_pack / 1e6,
           sell_nnz_per_s=P.nnz / ms_sell * 1e3)
print(json.dumps(out))
