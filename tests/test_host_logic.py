"""CPU tier: the C-ABI library loads and exports every symbol the header declares
(no compute without a GPU), the product path refuses to run without CUDA, and the
host-side Krylov / selection logic (pure torch) is correct on CPU tensors."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from helpers import pack_np

HEADER = os.path.join(ROOT, "include", "fgk_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    from flow_guided_krylov_b200.build import build_library
    return build_library()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fgk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fgk_b200.h but not exported"
    L.fgk_version.restype = C.c_int
    assert L.fgk_version() >= 100


def test_binding_table_matches_header(built_lib):
    from flow_guided_krylov_b200 import _native as nat
    assert sorted(nat._SIGNATURES) == declared_symbols()
    nat.lib()


def test_no_cpu_fallback():
    import flow_guided_krylov_b200 as f
    integ = f.MolecularIntegrals(np.eye(2), np.zeros((2,) * 4), 0.0, 2, 2, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.MolecularHamiltonian(integ, device="cpu")
    if not torch.cuda.is_available():
        from flow_guided_krylov_b200 import _native as nat
        with pytest.raises(RuntimeError):
            nat.ptr(torch.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flow_guided_krylov_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "orc_"):
                    assert needle not in txt, f"{fn} reaches into the oracle ({needle})"


class DenseOp:
    """stand-in for ProjectedH on CPU tensors (host-logic tests only)."""

    def __init__(self, A):
        self.A = A
        self.n = A.shape[0]
        self.row_begin, self.row_end = 0, self.n
        idx = A.nonzero()
        self.cols = idx[:, 1].to(torch.int32)
        self.vals = A[idx[:, 0], idx[:, 1]]

    def matvec(self, x):
        return (self.A.to(x.dtype)) @ x

    def diagonal(self):
        return torch.diagonal(self.A).clone()

    def to_dense(self):
        return self.A


def test_davidson_matches_dense_eigh():
    from flow_guided_krylov_b200.solvers import lowest_eigenpairs
    g = load_golden("skqd_lih")
    import scipy.sparse as sp
    n = len(g["H_indptr"]) - 1
    M = sp.csr_matrix((g["H_data"], g["H_indices"], g["H_indptr"]), shape=(n, n)).toarray()
    S = torch.from_numpy(0.5 * (M + M.T))
    w_ref = np.linalg.eigvalsh(S.numpy())
    op = DenseOp(S)
    w, v = lowest_eigenpairs(op, k=2, dense_max=0)
    assert np.abs(w.numpy() - w_ref[:2]).max() < 1e-9
    assert float(torch.linalg.norm(S @ v[:, 0] - w[0] * v[:, 0])) < 1e-8
    w1, _ = lowest_eigenpairs(op, k=1, dense_max=0, max_space=12)    # exercises thick restart
    assert abs(float(w1[0]) - w_ref[0]) < 1e-9
    wd, _ = lowest_eigenpairs(op, k=3)                               # dense branch
    assert np.abs(wd.numpy() - w_ref[:3]).max() < 1e-10


def test_expm_multiply_matches_scipy():
    from scipy.sparse.linalg import expm_multiply as sp_expm
    import scipy.sparse as sp
    from flow_guided_krylov_b200.solvers import expm_multiply
    g = load_golden("skqd_lih")
    n = len(g["H_indptr"]) - 1
    M = sp.csr_matrix((g["H_data"], g["H_indices"], g["H_indptr"]), shape=(n, n))
    op = DenseOp(torch.from_numpy(M.toarray()))
    psi = np.zeros(n, np.complex128)
    psi[int(g["hf_index"])] = 1.0
    ours = torch.from_numpy(psi)
    for step in range(3):
        ours = expm_multiply(op, ours, -0.1j)
        assert np.abs(ours.numpy() - g["psi_steps"][step]).max() < 1e-12
    big = expm_multiply(op, torch.from_numpy(psi), -2.5j)            # several scaling steps
    assert np.abs(big.numpy() - sp_expm(-2.5j * M, psi)).max() < 1e-11


def test_expm_multiply_spectral_radius_scaling_and_guard():
    """Taylor scaling by the estimated spectral radius: same result as the 1-norm scaling with
    fewer products on the reference's (non-Hermitian, F3) subspace matrix; an estimate that is far
    too small trips the a-posteriori check and falls back to the 1-norm parameters"""
    import scipy.sparse as sp
    from scipy.sparse.linalg import expm_multiply as sp_expm
    from flow_guided_krylov_b200.solvers import (_taylor_parameters, complex_abs2, expm_multiply,
                                                 spectral_radius_estimate)
    g = load_golden("skqd_lih")
    n = len(g["H_indptr"]) - 1
    M = sp.csr_matrix((g["H_data"], g["H_indices"], g["H_indptr"]), shape=(n, n))
    A = torch.from_numpy(M.toarray())
    op = DenseOp(A)
    calls = [0]

    def mv(x):
        calls[0] += 1
        return A.to(x.dtype) @ x
    psi = torch.zeros(n, dtype=torch.complex128)
    psi[int(g["hf_index"])] = 1.0
    mu = float(torch.diagonal(A).mean())
    d = torch.diagonal(A)
    norm1 = float((A.abs().sum(0) - d.abs() + (d - mu).abs()).max())
    rho = spectral_radius_estimate(mv, n, mu, "cpu", iters=30)
    lam = np.abs(np.linalg.eigvals(M.toarray() - mu * np.eye(n))).max()
    assert 0.7 * lam <= rho <= 1.05 * lam and rho < norm1
    for t in (-0.1j, -2.5j):
        ref = sp_expm(t * M, psi.numpy())
        calls[0] = 0
        a = expm_multiply(op, psi, t, matvec=mv, mu=mu, norm1=norm1)
        n_a = calls[0]
        calls[0] = 0
        b = expm_multiply(op, psi, t, matvec=mv, mu=mu, norm1=norm1, rho=rho)
        n_b = calls[0]
        assert np.abs(a.numpy() - ref).max() < 1e-11 and np.abs(b.numpy() - ref).max() < 1e-11
        assert n_b <= n_a
    # a far too small estimate (symmetrised matrix: a unitary, well-conditioned evolution): the series is
    # not finished after m* terms, the a-posteriori check redoes the step with the 1-norm parameters
    Ssym = 0.5 * (M + M.T)
    As = torch.from_numpy(Ssym.toarray())

    def mvs(x):
        calls[0] += 1
        return As.to(x.dtype) @ x
    ds = torch.diagonal(As)
    norm1s = float((As.abs().sum(0) - ds.abs() + (ds - mu).abs()).max())
    ref_s = sp_expm(-3.0j * Ssym, psi.numpy())
    calls[0] = 0
    good = expm_multiply(DenseOp(As), psi, -3.0j, matvec=mvs, mu=mu, norm1=norm1s)
    n_good = calls[0]
    calls[0] = 0
    c = expm_multiply(DenseOp(As), psi, -3.0j, matvec=mvs, mu=mu, norm1=norm1s, rho=0.02 * rho)
    assert np.abs(c.numpy() - ref_s).max() < 1e-10 and np.abs(good.numpy() - ref_s).max() < 1e-10
    assert calls[0] > n_good                                  # the aborted attempt + the redone step
    assert abs(float(torch.linalg.norm(c)) - 1.0) < 1e-10
    # parameter table: cost grows with the argument, degree capped in spectral-radius mode
    costs = [m * s_ for m, s_ in (_taylor_parameters(a_) for a_ in (0.01, 0.1, 1.0, 10.0, 100.0))]
    assert costs == sorted(costs) and _taylor_parameters(0.0) == (1, 1)
    assert all(_taylor_parameters(a_, 30)[0] <= 30 for a_ in (0.5, 5.0, 50.0, 500.0))
    z = torch.tensor([3 + 4j, 1j], dtype=torch.complex128)
    assert torch.equal(complex_abs2(z), torch.tensor([25.0, 1.0], dtype=torch.float64))


def test_pt2_pass_planning():
    from flow_guided_krylov_b200.expansion import DISTINCT_PER_RAW, planned_passes
    assert planned_passes(1000, 10 ** 9) == 1
    assert planned_passes(4_434_296_832, 1_700_000_000) == 2            # configs[4] shape on one GPU
    assert planned_passes(4_434_296_832, 440_000_000, 8) == 1           # ... on eight
    assert planned_passes(10 ** 9, 10 ** 6) == int(np.ceil(DISTINCT_PER_RAW * 1e9 / 1e6))


def test_select_top_k_ties_and_order():
    from flow_guided_krylov_b200.expansion import select_top_k
    dets = torch.tensor([[5, 1], [2, 9], [2, 3], [7, 0], [1, 1], [2, 4]], dtype=torch.int64)
    score = torch.tensor([1.0, 3.0, 3.0, 0.5, 3.0, 2.0], dtype=torch.float64)
    d, s = select_top_k(dets, score, 2, 10)
    assert s.tolist() == [3.0, 3.0]
    assert d.tolist() == [[1, 1], [2, 3]]                 # ties resolved by ascending key
    d, s = select_top_k(dets, score, 4, 10)
    assert d.tolist() == [[1, 1], [2, 3], [2, 9], [2, 4]]
    d, s = select_top_k(dets, score, 100, 10)
    assert d.shape[0] == 6 and s.tolist() == sorted(score.tolist(), reverse=True)
    d, s = select_top_k(dets[:0], score[:0], 3, 10)
    assert d.shape[0] == 0


def test_sort_unique_dets_is_reference_order():
    from flow_guided_krylov_b200.hamiltonian import sort_unique_dets
    g = load_golden("ham_wide")          # 66 sites: beyond the reference's int64 key
    n_orb = int(g["shape"][0])
    cfg = np.concatenate([g["basis"], g["basis"][::3], g["dets"]])
    want = torch.unique(torch.from_numpy(cfg.astype(np.int64)), dim=0).numpy().astype(np.uint8)
    got = sort_unique_dets(torch.from_numpy(pack_np(cfg, n_orb).view(np.int64)), n_orb)
    assert np.array_equal(got.numpy().view(np.uint64), pack_np(want, n_orb))
    # n_orb == 64 uses the sign-flip path
    d = torch.tensor([[-1, 3], [1, 2], [-(2 ** 63), 0], [1, 2], [0, -1]], dtype=torch.int64)
    out = sort_unique_dets(d, 64).numpy().view(np.uint64)
    keys = [(int(a) << 64) | int(b) for a, b in out]
    assert keys == sorted(set(keys)) and len(keys) == 4


def test_bench_stdout_carries_only_the_json_line(tmp_path):
    """bench.py's stdout guard: C-level writes to fd 1 (NCCL's version banner) and Python prints
    inside the run go to stderr; only the emitted JSON line reaches stdout."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "t.py"
    script.write_text(
        "import sys, ctypes, json\n"
        f"sys.path.insert(0, {root!r})\n"
        "from bench import CleanStdout\n"
        "libc = ctypes.CDLL(None)\n"
        "with CleanStdout() as out:\n"
        "    libc.puts(b'NCCL version 2.28.9+cuda12.9'); libc.fflush(None)\n"
        "    print('chatter')\n"
        "    out.emit(json.dumps({'metric': 'm', 'value': 1.5}))\n")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout) == {"metric": "m", "value": 1.5}
    assert "NCCL version" in r.stderr and "chatter" in r.stderr


@pytest.mark.parametrize("seed", range(6))
def test_select_top_k_equals_the_oracle_tie_protocol(seed):
    """The product's select_top_k (torch, runs on CPU tensors too) and the oracle's top_k_protocol
    pick the same candidates when groups of exactly / nearly degenerate scores straddle the cut --
    the situation real molecules create (symmetry-equivalent determinants), where the two sides see
    last-bit different sums."""
    import torch
    from flow_guided_krylov_b200.expansion import select_top_k
    from helpers import pack_np, random_dets
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    n_orb = [6, 20, 33, 48, 64, 10][seed]
    cfg = np.unique(random_dets(n_orb, 3, 2, 400, rng), axis=0)
    m = len(cfg)
    base = rng.random(m // 4 + 1) ** 4                       # few distinct levels -> big tie groups
    score = base[rng.integers(0, len(base), m)]
    noise_a = score * (1.0 + 1e-15 * rng.standard_normal(m))   # what "the GPU" sees
    noise_b = score * (1.0 + 1e-15 * rng.standard_normal(m))   # what "the oracle" sees
    dets = torch.from_numpy(pack_np(cfg, n_orb).view(np.int64))
    for k in (1, 5, m // 3, m // 2, m - 1, m, m + 7):
        sd, ss = select_top_k(dets, torch.from_numpy(noise_a), k, n_orb)
        pick = orc.top_k_protocol(cfg, noise_b, min(k, m))
        mine = {bytes(r) for r in sd.numpy().view(np.uint64)}
        theirs = {bytes(r) for r in pack_np(cfg[pick], n_orb)}
        assert mine == theirs and len(mine) == min(k, m)
        assert np.all(np.diff(ss.numpy()) <= 1e-9 * ss.numpy()[:-1])


def test_particle_number_guard_of_indexed_bases():
    """projected_csr / projected_sell / PT2 accumulate size per-warp lists from the Hamiltonian's
    n_alpha / n_beta; a basis with other particle numbers must be refused up front (host check,
    runs on CPU tensors)."""
    import types
    import torch
    from flow_guided_krylov_b200.hamiltonian import MolecularHamiltonian, popcount64
    from helpers import pack_np, random_dets
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2 ** 64, size=5000, dtype=np.uint64)
    x[:4] = [0, 2 ** 64 - 1, 2 ** 63, 1]
    assert np.array_equal(popcount64(torch.from_numpy(x.view(np.int64))).numpy(),
                          [bin(int(v)).count("1") for v in x])
    ham = types.SimpleNamespace(n_alpha=3, n_beta=2)

    class Idx:
        def __init__(self, dets):
            self.dets = dets

        def __len__(self):
            return self.dets.shape[0]

    good = torch.from_numpy(pack_np(random_dets(20, 3, 2, 50, rng), 20).view(np.int64))
    idx = Idx(good)
    MolecularHamiltonian._require_particle_numbers(ham, idx, "test")
    assert idx._particles_ok == (3, 2)
    bad = good.clone()
    bad[17, 1] |= 1 << 19                       # one more beta electron (orbital 0 of 20)
    with pytest.raises(ValueError, match="determinant 17"):
        MolecularHamiltonian._require_particle_numbers(ham, Idx(bad), "test")
    MolecularHamiltonian._require_particle_numbers(ham, Idx(good[:0]), "test")      # empty basis: fine
