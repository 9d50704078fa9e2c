"""torchrun worker for tests/test_gpu_multi.py: N-rank results must equal 1-rank results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    rank, world = dist.get_rank(), dist.get_world_size()
    import flow_guided_krylov_b200 as fgk
    from flow_guided_krylov_b200 import dist as fd
    from flow_guided_krylov_b200.solvers import lowest_eigenpairs, expm_multiply
    from bench import synth_integrals, cas_window_basis

    n_orb = 20
    h1, g = synth_integrals(n_orb, 2)
    H = fgk.MolecularHamiltonian(fgk.MolecularIntegrals(h1, g, 0.0, 10, n_orb, 5, 5), dev)
    dets = torch.from_numpy(cas_window_basis(n_orb, 2, 9, 3).view(np.int64)).to(dev)   # C(9,3)^2 = 7056
    n = dets.shape[0]
    idx = fgk.BasisIndex(dets)

    # projected H: row blocks == rows of the full build; sharded H.v == full H.v
    Pfull = H.projected_csr(dets, fgk.H_SYM, index=idx, packed=True, sort_rows=True)
    Pblk, op = fd.build_sharded_h(H, dets, fgk.H_SYM, index=idx, sort_rows=True)
    lo, hi = fd.row_block(n, rank, world)
    assert torch.equal(Pblk.row_ptr, Pfull.row_ptr[lo:hi + 1] - Pfull.row_ptr[lo])
    assert torch.equal(Pblk.cols, Pfull.cols[Pfull.row_ptr[lo]:Pfull.row_ptr[hi]])
    assert torch.equal(Pblk.vals, Pfull.vals[Pfull.row_ptr[lo]:Pfull.row_ptr[hi]])
    gen = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(n, dtype=torch.float64, generator=gen).to(dev)
    assert float((op.matvec(x) - Pfull.matvec(x)).abs().max()) < 1e-12
    Pblk.to_sell()
    assert float((op.matvec(x) - Pfull.matvec(x)).abs().max()) < 1e-11
    z = torch.complex(x, torch.flip(x, [0]))
    assert float((op.matvec(z) - Pfull.matvec(z)).abs().max()) < 1e-11

    # one-launch step (product + peer broadcast + barrier): same vectors, several ping-pong steps,
    # FP64 and packed storage, real and complex vectors
    yref = Pfull.matvec(x)
    zref = Pfull.matvec(z)
    wf = None
    for storage in ("sell", "packed"):
        fop = fd.FusedShardedOperator(Pblk, storage=storage)
        assert fop.packed == (storage == "packed")
        assert float((fop.matvec(x) - yref).abs().max()) < 1e-11
        assert float((fop.matvec(z) - zref).abs().max()) < 1e-11
        fop.load(x)
        cur = x
        for _ in range(5):
            cur = Pfull.matvec(cur)
            cur = cur / torch.linalg.norm(cur)
            got = fop.step()
            got /= torch.linalg.norm(got)          # in-place on the view: every rank scales its own copy
            assert float((got - cur).abs().max()) < 1e-11
        fop.load(z)
        cz = z
        for _ in range(3):
            cz = Pfull.matvec(cz)
            cz = cz / torch.linalg.norm(cz)
            got = fop.step()
            got /= torch.linalg.norm(got)
            assert got.is_complex() and float((got - cz).abs().max()) < 1e-11
        # row-sharded input: peer gather + local rows only
        yl = fop.matvec_local(x[lo:hi].contiguous())
        assert float((yl - yref[lo:hi]).abs().max()) < 1e-11
        zl = fop.matvec_local(z[lo:hi].contiguous())
        assert float((zl - zref[lo:hi]).abs().max()) < 1e-11
        # host vectors: every rank uploads its slice only
        yh = fop.matvec_host(x.cpu().pin_memory())
        assert float((yh.to(dev) - yref[lo:hi]).abs().max()) < 1e-11
        yh = fop.matvec_host(x.cpu().pin_memory())          # twice: buffers ping-pong
        assert float((yh.to(dev) - yref[lo:hi]).abs().max()) < 1e-11
        # small all-reduce over peer memory: identical bits on every rank, equals NCCL's sum
        for rep in range(5):                                  # consecutive calls alternate the scratch area
            part = torch.randn(37 + rep, dtype=torch.float64, generator=torch.Generator().manual_seed(100 * rank + rep)).to(dev)
            ref_sum = part.clone()
            dist.all_reduce(ref_sum)
            got_sum = fop.allreduce_sum_(part.clone())
            assert float((got_sum - ref_sum).abs().max()) < 1e-13
            gathered = [torch.empty_like(got_sum) for _ in range(world)]
            dist.all_gather(gathered, got_sum)
            assert all(torch.equal(gathered[0], gq) for gq in gathered)
        # replicated Davidson through matvec, row-sharded Davidson through matvec_local
        wf, _ = lowest_eigenpairs(fop, k=1, matvec=fop.matvec, diagonal=fop.diagonal(), dense_max=0)
        ws, vs = lowest_eigenpairs(fop, k=2, sharded=fop)
        fop.check()
        assert abs(float(wf[0]) - float(ws[0])) < 1e-9
        res = Pfull.matvec(vs[:, 0].contiguous()) - ws[0] * vs[:, 0]
        assert float(torch.linalg.norm(res)) < 1e-8
        fop.close()

    # Davidson and Taylor expm through the sharded operator
    w1, v1 = lowest_eigenpairs(Pfull, k=2, dense_max=0)
    w2, v2 = lowest_eigenpairs(op, k=2, matvec=op.matvec, diagonal=op.diagonal(), dense_max=0)
    assert float((w1 - w2).abs().max()) < 1e-9, (w1, w2)
    assert abs(float(wf[0]) - float(w1[0])) < 1e-9
    assert float((ws - w1).abs().max()) < 1e-9
    psi = torch.zeros(n, dtype=torch.complex128, device=dev)
    psi[0] = 1.0
    e1 = expm_multiply(Pfull, psi, -0.1j)
    from flow_guided_krylov_b200.solvers import one_norm
    d = Pfull.diagonal()
    mu = float(d.sum()) / n
    nrm = float((one_norm(Pfull) - d.abs() + (d - mu).abs()).max())
    e2 = expm_multiply(op, psi, -0.1j, matvec=op.matvec, mu=mu, norm1=nrm)
    assert float((e1 - e2).abs().max()) < 1e-12

    # PT2: N-rank selection == 1-rank selection (also with a tiny workspace -> multi-pass)
    coeff = torch.zeros(n, dtype=torch.float64, device=dev)
    pick = torch.randperm(n, generator=gen)[:300].to(dev)
    coeff[pick] = torch.randn(300, dtype=torch.float64, generator=gen).to(dev)
    E = float(w1[0])
    s_ref, i_ref, st_ref = fgk.pt2_select(H, idx, coeff, E, 64)
    for cap in (None, 20000):
        wsp = fgk.Pt2Workspace(cap, dev) if cap else None
        s_sh, i_sh, st = fd.pt2_select_sharded(H, idx, coeff, E, 64, workspace=wsp)
        assert st["raw_candidates_total"] == st_ref["raw_candidates"], (st, st_ref)
        assert st["unique_total"] == st_ref["unique_candidates"]
        assert torch.equal(s_sh, s_ref)
        assert torch.allclose(i_sh, i_ref, rtol=1e-10, atol=1e-18)
        if cap:
            assert st["passes"] > 1
    s_mx, r_mx, _ = fd.pt2_select_sharded(H, idx, coeff, 0.0, 40, mode=fgk.PT2_MAXABS)
    s_m1, r_m1, _ = fgk.pt2_select(H, idx, coeff, 0.0, 40, mode=fgk.PT2_MAXABS)
    assert torch.equal(s_mx, s_m1) and torch.equal(r_mx, r_m1)
    # Stage 3 end to end under torchrun: sharded PT2 (and, above the row threshold, sharded H +
    # fused H.v Davidson) must reproduce the single-process expander
    ex = fgk.SelectedCIExpander(H, fgk.ResidualExpansionConfig(max_configs_per_iter=150))
    b0 = H.unpack(dets[:600].contiguous())
    ex.sharded_min_rows = 10 ** 9
    b1, st1 = ex.expand_basis(b0)                       # sharded PT2 only
    ex.sharded_min_rows = 64
    b2, st2 = ex.expand_basis(b0)                       # + sharded H / fused Davidson
    wsz = dist.get_world_size()
    fgk.SelectedCIExpander._world = staticmethod(lambda: 1)
    b3, st3 = fgk.SelectedCIExpander(H, fgk.ResidualExpansionConfig(max_configs_per_iter=150)).expand_basis(b0)
    assert torch.equal(b1, b3) and torch.equal(b2, b3)
    assert abs(st1["final_energy"] - st3["final_energy"]) < 1e-9
    assert abs(st2["final_energy"] - st3["final_energy"]) < 1e-9
    assert st3["configs_added"] == 150
    # Stage 4 on a growing determinant set (adaptive SKQD): rows of H_S sharded, complex one-launch
    # step for exp(-i dt H_S), growth by the sharded MAXABS selection -- same sets, same amplitudes
    # and same energies as the single-process run
    cfg = fgk.SKQDConfig(max_krylov_dim=3, shots_per_krylov=4000, subspace_mode="adaptive",
                         max_subspace_size=4096 * wsz + 12000, expand_sources=48,
                         expand_new_per_round=4096 * wsz + 4000)     # the set must pass 4096 rows per rank
    nf = H.unpack(dets[torch.randperm(n, generator=torch.Generator().manual_seed(4))[:300].to(dev)].contiguous())
    fgk.SampleBasedKrylovDiagonalization._world = staticmethod(lambda: wsz)
    torch.manual_seed(5)
    sk_n = fgk.FlowGuidedSKQD(H, nf, cfg)
    res_n = sk_n.run_with_nf(progress=False)
    assert sk_n._subspace_op is not None            # the sharded branch ran
    fgk.SampleBasedKrylovDiagonalization._world = staticmethod(lambda: 1)
    torch.manual_seed(5)
    sk_1 = fgk.FlowGuidedSKQD(H, nf, cfg)
    res_1 = sk_1.run_with_nf(progress=False)
    assert sk_n.subspace_history == sk_1.subspace_history and sk_n.subspace_history[-1] >= 4096 * wsz
    assert torch.equal(sk_n._subspace_dets, sk_1._subspace_dets)
    for a, b in zip(sk_n.krylov_states, sk_1.krylov_states):
        assert float((a - b).abs().max()) < 1e-11
    for a, b in zip(sk_n.krylov_sample_dets, sk_1.krylov_sample_dets):
        assert torch.equal(a, b)
    assert np.abs(np.array(res_n["energies_combined"]) - np.array(res_1["energies_combined"])).max() < 1e-9
    sk_n._subspace_op.close()
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} n={n} nnz={Pfull.nnz} raw={st_ref['raw_candidates']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print("WORKER_FAILED rank", os.environ.get("RANK"), flush=True)
        traceback.print_exc(file=sys.stdout)
        sys.stdout.flush()
        raise
