"""Test-side helpers (numpy only): packing, hostcheck loader, synthetic integrals."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flow_guided_krylov_b200", "csrc")


def pack_np(cfgs, n_orb):
    """(n, 2*n_orb) 0/1 -> (n, 2) uint64 {alpha, beta}; orbital p -> bit n_orb-1-p."""
    cfgs = np.asarray(cfgs).reshape(-1, 2 * n_orb).astype(np.uint64)
    w = (np.uint64(1) << np.arange(n_orb - 1, -1, -1, dtype=np.uint64))
    out = np.empty((len(cfgs), 2), np.uint64)
    out[:, 0] = (cfgs[:, :n_orb] * w).sum(1, dtype=np.uint64)
    out[:, 1] = (cfgs[:, n_orb:] * w).sum(1, dtype=np.uint64)
    return out


def unpack_np(dets, n_orb):
    dets = np.asarray(dets, np.uint64).reshape(-1, 2)
    sh = np.arange(n_orb - 1, -1, -1, dtype=np.uint64)
    a = ((dets[:, 0:1] >> sh) & np.uint64(1)).astype(np.uint8)
    b = ((dets[:, 1:2] >> sh) & np.uint64(1)).astype(np.uint8)
    return np.concatenate([a, b], axis=1)


def synth_integrals(n_orb, seed=0, h1_scale=1.0, h2_scale=0.1):
    """SURVEY.md Appendix D generator (same as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    h1 = rng.standard_normal((n_orb, n_orb)) * h1_scale
    h1 = 0.5 * (h1 + h1.T)
    g = rng.standard_normal((n_orb,) * 4) * h2_scale
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    return h1, g


def random_dets(n_orb, na, nb, n, rng):
    out = np.zeros((n, 2 * n_orb), dtype=np.uint8)
    for i in range(n):
        out[i, rng.choice(n_orb, na, replace=False)] = 1
        out[i, n_orb + rng.choice(n_orb, nb, replace=False)] = 1
    return out


_hc = None


def hostcheck():
    """g++ build of csrc/fgk_hostcheck.cpp: the kernels' own __host__ __device__
    arithmetic, runnable without a GPU (test infrastructure, not a fallback)."""
    global _hc
    if _hc is None:
        so = os.path.join(CSRC, "libfgk_hostcheck.so")
        srcs = [os.path.join(CSRC, f) for f in ("fgk_hostcheck.cpp", "fgk_core.cuh", "fgk_lists.cuh", "fgk_tables.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-x", "c++", "-std=c++17", "-fPIC", "-shared",
                                   "-o", so, srcs[0]])
        L = C.CDLL(so)
        vp, i64, ci, dbl, u64 = C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_uint64
        L.hc_ham_create.restype = vp
        L.hc_ham_create.argtypes = [vp, vp, ci, ci, ci, dbl]
        L.hc_ham_destroy.argtypes = [vp]
        L.hc_diag.argtypes = [vp, vp, i64, vp]
        L.hc_diag_loops.argtypes = [vp, vp, i64, vp]
        L.hc_connections.restype = i64
        L.hc_connections.argtypes = [vp, u64, u64, vp, vp, i64]
        L.hc_bra_row.restype = i64
        L.hc_bra_row.argtypes = [vp, vp, i64, i64, ci, vp, vp, i64]
        L.hc_bra_row2.restype = i64
        L.hc_bra_row2.argtypes = [vp, vp, i64, i64, ci, ci, vp, vp, i64]
        L.hc_bra_row3.restype = i64
        L.hc_bra_row3.argtypes = [vp, vp, i64, i64, ci, vp, vp, i64]
        L.hc_bra_row4.restype = i64
        L.hc_bra_row4.argtypes = [vp, vp, i64, i64, ci, vp, vp, i64]
        L.hc_pt2_walk2.restype = i64
        L.hc_pt2_walk2.argtypes = [vp, u64, u64, vp, vp, i64]
        L.hc_check_split.restype = i64
        L.hc_check_split.argtypes = [vp, vp, i64]
        L.hc_fx_sum.restype = ci
        L.hc_fx_sum.argtypes = [vp, vp, i64, vp]
        _hc = L
    return _hc
