"""N > 1 exchange logic on CPU tensors, world_size 2, gloo backend (runs without a
GPU): Krylov-vector all-gather, PT2 dedup exchange by owner, global top-k merge,
and a Davidson solve through the row-block sharded operator."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flow_guided_krylov_b200 import dist as fd
        from flow_guided_krylov_b200.expansion import select_top_k
        from flow_guided_krylov_b200.solvers import lowest_eigenpairs, expm_multiply

        # --- row blocks + vector all-gather (n not divisible by world) ---
        n = 101
        lo, hi = fd.row_block(n, rank, world)
        full = torch.arange(n, dtype=torch.float64) * 0.5
        got = fd.allgather_vector(full[lo:hi].clone(), n)
        assert torch.equal(got, full)
        z = torch.complex(full, -full)
        assert torch.equal(fd.allgather_vector(z[lo:hi].clone(), n), z)

        # --- sharded operator: local dense row block, gathered result ---
        g = torch.Generator().manual_seed(7)
        A = torch.randn(n, n, dtype=torch.float64, generator=g)
        A = 0.5 * (A + A.T) + torch.diag(torch.linspace(-5, 5, n, dtype=torch.float64))
        op = fd.ShardedOperator(n, lambda x: A[lo:hi].to(x.dtype) @ x, torch.diagonal(A)[lo:hi].clone())
        x = torch.randn(n, dtype=torch.float64, generator=g)
        assert torch.allclose(op.matvec(x), A @ x, atol=1e-12)
        assert torch.equal(op.diagonal(), torch.diagonal(A))
        w, v = lowest_eigenpairs(op, k=2, matvec=op.matvec, diagonal=op.diagonal(), dense_max=0)
        wr = torch.linalg.eigvalsh(A)[:2]
        assert torch.allclose(w, wr, atol=1e-9), (w, wr)

        # --- ROW-SHARDED Davidson (solvers._davidson_sharded): the operator contract of
        # dist.FusedShardedOperator (matvec_local on the rank's slice, diagonal, check) served by
        # gloo all-gathers; must agree with the replicated iteration and with dense eigh ---
        class MockFused:
            def __init__(self):
                self.n, self.row_begin, self.row_end, self.world, self.rank = n, lo, hi, world, rank
                self.calls = 0

            def diagonal(self):
                return torch.diagonal(A).clone()

            def matvec_local(self, x_local, out=None):
                self.calls += 1
                assert x_local.shape[0] == hi - lo
                y = A[lo:hi] @ fd.allgather_vector(x_local.clone(), n)
                if out is not None:
                    out.copy_(y)
                    return out
                return y

            def check(self):
                pass

        mop = MockFused()
        ws_, vs_ = lowest_eigenpairs(mop, k=2, sharded=mop)
        assert torch.allclose(ws_, wr, atol=1e-9), (ws_, wr)
        assert vs_.shape == (n, 2)
        res = A @ vs_[:, 0] - ws_[0] * vs_[:, 0]
        assert float(torch.linalg.norm(res)) < 1e-8
        assert abs(float(torch.linalg.norm(vs_[:, 0])) - 1.0) < 1e-9
        assert torch.allclose(ws_, w, atol=1e-9)

        class Fake:  # expm through the sharded operator
            n = 101
        psi = torch.zeros(n, dtype=torch.complex128)
        psi[3] = 1.0
        cs = A.abs().sum(0)
        out = expm_multiply(Fake, psi, -0.05j, matvec=op.matvec, mu=float(torch.diagonal(A).mean()),
                            norm1=float(cs.max()) + 5.0)
        ref = torch.matrix_exp(-0.05j * A.to(torch.complex128)) @ psi
        assert torch.allclose(out, ref, atol=1e-11)

        # --- PT2 dedup exchange: every key lands on its owner, nothing lost ---
        rng = np.random.default_rng(100 + rank)
        keys = torch.from_numpy(rng.integers(0, 50, size=(400, 2)).astype(np.int64))
        vals = torch.from_numpy(rng.standard_normal(400))
        rd, rv = fd.exchange_by_owner(keys, vals)
        assert torch.all(fd.owner_of(rd, world) == rank)
        tot_sent = torch.tensor([float(vals.sum()), 400.0], dtype=torch.float64)
        tot_recv = torch.tensor([float(rv.sum()), float(rd.shape[0])], dtype=torch.float64)
        dist.all_reduce(tot_sent)
        dist.all_reduce(tot_recv)
        assert abs(float(tot_sent[0] - tot_recv[0])) < 1e-9 and tot_sent[1] == tot_recv[1]
        # reduce duplicates locally, then the global top-k must equal the single-process answer
        uniq, inv = torch.unique(rd, dim=0, return_inverse=True)
        summed = torch.zeros(uniq.shape[0], dtype=torch.float64).index_add_(0, inv, rv)
        sel, sc = fd.merge_topk(uniq, summed.abs(), 7, 20)
        gathered = [None] * world
        dist.all_gather_object(gathered, (keys.numpy(), vals.numpy()))
        ak = torch.from_numpy(np.concatenate([a for a, _ in gathered]))
        av = torch.from_numpy(np.concatenate([b for _, b in gathered]))
        u2, i2 = torch.unique(ak, dim=0, return_inverse=True)
        s2 = torch.zeros(u2.shape[0], dtype=torch.float64).index_add_(0, i2, av)
        rsel, rsc = select_top_k(u2, s2.abs(), 7, 20)
        assert torch.equal(sel, rsel)
        assert torch.allclose(sc, rsc, atol=1e-12)
        assert abs(fd.allreduce_scalar(float(rank + 1), "sum") - 3.0) < 1e-12
        assert fd.allreduce_scalar(float(rank), "max") == 1.0
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    port = 29000 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok0")) and os.path.exists(os.path.join(tmp, "ok1"))
