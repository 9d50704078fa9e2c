"""ctypes binding of libfgk_b200.so (include/fgk_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a
RuntimeError carrying fgk_last_error() is raised.  torch is used only for
device memory and streams (tensors own every buffer the C ABI writes into).
"""
import ctypes as C
import os

import torch

from .build import LIB

_lib = None

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
H_RAW, H_SYM, H_DROP_ZEROS, H_FLAT_WALK, H_HASH_WALK = 0, 1, 2, 4, 8
PT2_SUM, PT2_MAXABS = 0, 1
PEER_COMPLEX, PEER_PACKED_F32 = 1, 2

vp, i64, ci, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double

_SIGNATURES = {
    "fgk_version": (ci, []),
    "fgk_last_error": (C.c_char_p, []),
    "fgk_device_info": (ci, [ci, C.POINTER(ci), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                             C.POINTER(C.c_size_t)]),
    "fgk_ham_create": (ci, [vp, vp, ci, ci, ci, dbl, ci, C.POINTER(vp)]),
    "fgk_ham_destroy": (ci, [vp]),
    "fgk_pack_i64": (ci, [vp, i64, ci, vp, ci, vp]),
    "fgk_unpack_i64": (ci, [vp, i64, ci, vp, ci, vp]),
    "fgk_diag": (ci, [vp, vp, i64, vp, vp]),
    "fgk_conn_count": (ci, [vp, vp, i64, vp, vp]),
    "fgk_conn_fill": (ci, [vp, vp, i64, vp, vp, vp, vp, vp]),
    "fgk_index_create": (ci, [vp, i64, ci, vp, C.POINTER(vp)]),
    "fgk_index_destroy": (ci, [vp]),
    "fgk_index_lookup": (ci, [vp, vp, i64, vp, vp]),
    "fgk_index_info": (ci, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "fgk_index_layout": (ci, [vp, C.POINTER(ci)]),
    "fgk_projh_count": (ci, [vp, vp, i64, i64, ci, vp, vp]),
    "fgk_projh_fill": (ci, [vp, vp, i64, i64, ci, vp, vp, vp, vp]),
    "fgk_projh_fill_sell": (ci, [vp, vp, i64, i64, ci, vp, vp, vp, vp]),
    "fgk_csr_sort_rows": (ci, [i64, vp, vp, vp, ci, vp]),
    "fgk_strlists_create": (ci, [vp, vp, vp, C.POINTER(vp)]),
    "fgk_strlists_destroy": (ci, [vp]),
    "fgk_strlists_info": (ci, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "fgk_projh_packed_bound": (ci, [vp, vp, vp, i64, i64, vp, vp]),
    "fgk_projh_packed_count": (ci, [vp, vp, vp, i64, i64, ci, i64, vp, vp]),
    "fgk_projh_packed_fill": (ci, [vp, vp, vp, i64, i64, ci, vp, vp, vp, vp, vp]),
    "fgk_spmv_f64": (ci, [i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_spmv_z": (ci, [i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_sell_fill": (ci, [i64, vp, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_spmv_sell_f64": (ci, [i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_spmv_sell_z": (ci, [i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_sell_pack_f32": (ci, [i64, i64, vp, vp, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_spmv_sell_f32_f64": (ci, [i64, i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_spmv_sell_f32_z": (ci, [i64, i64, vp, vp, vp, vp, vp, ci, vp]),
    "fgk_davidson_step": (ci, [ci, i64, i64, ci, vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, ci, vp, ci, vp]),
    "fgk_taylor_update_z": (ci, [i64, vp, vp, vp, dbl, dbl, dbl, vp, ci, vp]),
    "fgk_peer_alloc": (ci, [C.c_size_t, ci, C.POINTER(vp), C.c_char_p]),
    "fgk_peer_open": (ci, [C.c_char_p, ci, C.POINTER(vp)]),
    "fgk_peer_close": (ci, [vp, ci]),
    "fgk_peer_free": (ci, [vp, ci]),
    "fgk_peer_barrier": (ci, [C.POINTER(vp), ci, ci, C.c_uint64, vp, ci, vp]),
    "fgk_peer_step": (ci, [i64, vp, vp, vp, vp, vp, C.POINTER(vp), ci, i64, C.POINTER(vp), ci, ci, C.c_uint64,
                           vp, vp, ci, vp]),
    "fgk_peer_allreduce_sum": (ci, [vp, i64, i64, vp, C.POINTER(vp), i64, ci, C.POINTER(vp), ci, ci, C.c_uint64, vp, vp, ci, vp]),
    "fgk_peer_matvec_host": (ci, [i64, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, ci, i64, C.POINTER(vp),
                                  ci, ci, C.c_uint64, vp, vp, ci, vp]),
    "fgk_peer_gather": (ci, [vp, i64, C.POINTER(vp), i64, C.POINTER(vp), ci, ci, C.c_uint64, vp, vp, ci, vp]),
    "fgk_pt2_create": (ci, [i64, i64, vp, vp, vp, ci, C.POINTER(vp)]),
    "fgk_pt2_destroy": (ci, [vp]),
    "fgk_pt2_set_partition": (ci, [vp, ci, ci, i64, vp, vp, vp]),
    "fgk_pt2_reset": (ci, [vp, vp]),
    "fgk_pt2_score": (ci, [vp, vp, i64, dbl, i64, vp, vp, C.POINTER(i64), C.POINTER(i64), vp]),
    "fgk_pt2_gather": (ci, [vp, i64, vp, vp, vp, i64, C.POINTER(i64), vp]),
    "fgk_pt2_accumulate": (ci, [vp, vp, vp, vp, vp, i64, ci, ci, ci, vp]),
    "fgk_pt2_merge": (ci, [vp, vp, vp, i64, ci, vp]),
    "fgk_pt2_count": (ci, [vp, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(ci)]),
    "fgk_partition_by_owner": (ci, [vp, vp, i64, ci, vp, vp, vp, ci, ci, vp]),
    "fgk_pt2_export": (ci, [vp, vp, i64, dbl, vp, vp, vp, vp, C.POINTER(i64), vp]),
}


def lib():
    """Load libfgk_b200.so; raise (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(
                f"{LIB} is not built. Run `python -m flow_guided_krylov_b200.build` "
                "(needs nvcc). There is no CPU fallback for this engine.")
        L = C.CDLL(LIB)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().fgk_last_error().decode(errors="replace")
        raise RuntimeError(f"libfgk_b200 error {rc}: {msg}")


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("libfgk_b200 works on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("tensor passed to libfgk_b200 must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def device_index(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(
            f"flow_guided_krylov_b200 needs a CUDA device (got {device!r}); there is no CPU fallback")
    return d.index if d.index is not None else torch.cuda.current_device()


def device_info(device):
    sm, l2, fr, tot = ci(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    check(lib().fgk_device_info(device_index(device), C.byref(sm), C.byref(l2), C.byref(fr),
                                C.byref(tot)))
    return dict(sm_count=sm.value, l2_bytes=l2.value, free_bytes=fr.value, total_bytes=tot.value)
