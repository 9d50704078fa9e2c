"""Experiment (torchrun): cost breakdown of the fused multi-GPU H.v step."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import flow_guided_krylov_b200 as fgk
from flow_guided_krylov_b200 import dist as fd, _native as nat
from bench import synth_integrals, cas_window_basis

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
rank, world = dist.get_rank(), dist.get_world_size()
n_active = int(sys.argv[1]) if len(sys.argv) > 1 else 14
h1, g = synth_integrals(32, 0)
H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 16, 32, 8, 8), dev)
dets = torch.from_numpy(cas_window_basis(32, 4, n_active, 4).view(np.int64)).to(dev)
n = dets.shape[0]
P, op = fd.build_sharded_h(H, dets, fgk.H_SYM)
P.to_sell()
fop = fd.FusedShardedOperator(P)
x = torch.randn(n, dtype=torch.float64, device=dev) * 1e-3
fop.load(x)
L = nat.lib(); st = nat.stream_ptr(dev)
sp, sc, sv = P._sell
y_local = torch.empty(P.n_rows, dtype=torch.float64, device=dev)

def timeit(fn, reps=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

def k_plain(): P.matvec(x, out=y_local)
def k_bcast_only():
    nat.check(L.fgk_spmv_sell_f64_allgather(P.n_rows, nat.ptr(sp), nat.ptr(sc), nat.ptr(sv),
              C.c_void_p(fop._own[0]), fop._bufs[1], world, P.row_begin, fop.dev, st))
def k_bcast_self_only():
    VP = C.c_void_p * 1
    nat.check(L.fgk_spmv_sell_f64_allgather(P.n_rows, nat.ptr(sp), nat.ptr(sc), nat.ptr(sv),
              C.c_void_p(fop._own[0]), VP(fop._own[1]), 1, P.row_begin, fop.dev, st))
def barrier_only():
    fop._epoch += 1
    nat.check(L.fgk_peer_barrier(fop._flags, rank, world, fop._epoch, nat.ptr(fop._err), fop.dev, st))
def nccl_step():
    P.matvec(x, out=y_local); fd.allgather_vector(y_local, n)
def nccl_only(): fd.allgather_vector(y_local, n)
res = {}
for name, fn in [("plain_kernel", k_plain), ("bcast_kernel_self_only", k_bcast_self_only), ("bcast_kernel", k_bcast_only),
                 ("barrier_only", barrier_only), ("fused_step", fop.step), ("nccl_allgather_only", nccl_only), ("nccl_step", nccl_step)]:
    res[name] = timeit(fn)
fop.check()
if rank == 0:
    print(f"world={world} n={n} " + " ".join(f"{k}={v:.4f}ms" for k, v in res.items()), flush=True)
fop.close()
dist.destroy_process_group()
