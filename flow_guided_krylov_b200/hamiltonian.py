"""MolecularHamiltonian -- host-side mirror of the reference operator interface
(reference src/hamiltonians/molecular.py:35-942, src/hamiltonians/base.py:9-58)
over the sm_100a engine.  Same constructor, attribute and method names, same
argument meaning and the same empty-result conventions; every method body is a
thin call into libfgk_b200.so.  No CPU path exists.

Parity tiers: connected sets, emission order, off-diagonal values
and nonzero patterns are bit-exact with the reference; diagonals / energies are
FP64 on the reference's float32-rounded integrals (the reference's own float32
einsum has no defined summation order), within 1e-9 Ha of the FP64 oracle.
"""
import ctypes as C
from dataclasses import dataclass
from itertools import combinations
from math import comb
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat


@dataclass
class MolecularIntegrals:
    """Same fields and order as the reference container (molecular.py:22-32)."""
    h1e: np.ndarray
    h2e: np.ndarray
    nuclear_repulsion: float
    n_electrons: int
    n_orbitals: int
    n_alpha: int
    n_beta: int


class ProjectedH:
    """Rows [row_begin, row_end) of a projected Hamiltonian in device CSR
    (int64 row_ptr, int32 GLOBAL column ids, FP64 values)."""

    def __init__(self, n, row_ptr, cols, vals, device, row_begin=0, row_end=None, mode=nat.H_RAW,
                 sorted_rows=False):
        self.n = int(n)
        self.row_ptr, self.cols, self.vals = row_ptr, cols, vals
        self.device = device
        self.row_begin = int(row_begin)
        self.row_end = int(n if row_end is None else row_end)
        self.mode = mode
        self.sorted_rows = sorted_rows

    @property
    def n_rows(self):
        return self.row_end - self.row_begin

    @property
    def nnz(self):
        nz = getattr(self, "_nnz", None)
        if torch.is_tensor(nz):                 # left on the device by the builder: read once
            nz = self._nnz = int(nz.item())
        return int(nz or self.vals.numel())

    def _need_csr(self, what):
        if getattr(self, "sell_only", False):
            raise RuntimeError(f"{what} needs the CSR arrays; this operator was built SELL-only "
                               "(use projected_csr instead of projected_sell)")

    def sort_rows(self):
        self._need_csr("sort_rows")
        if not self.sorted_rows and self.nnz:
            nat.check(nat.lib().fgk_csr_sort_rows(
                self.n_rows, nat.ptr(self.row_ptr, torch.int64), nat.ptr(self.cols, torch.int32),
                nat.ptr(self.vals, torch.float64), nat.device_index(self.device),
                nat.stream_ptr(self.device)))
        self.sorted_rows = True
        return self

    def to_sell(self, keep_csr=True):
        """Build the SELL-32 copy used by matvec (fgk_sell_fill).  The CSR arrays stay
        the canonical, reference-comparable form unless keep_csr=False."""
        if getattr(self, "_sell", None) is not None:
            return self
        dev = self.cols.device
        n_slices = (self.n_rows + 31) // 32
        lens = torch.zeros(n_slices * 32, dtype=torch.int64, device=dev)
        lens[: self.n_rows] = self.row_ptr[1:] - self.row_ptr[:-1]
        width = lens.view(n_slices, 32).max(dim=1).values
        width = (width + 1) // 2 * 2
        slice_ptr = torch.zeros(n_slices + 1, dtype=torch.int64, device=dev)
        torch.cumsum(width * 32, 0, out=slice_ptr[1:])
        total = int(slice_ptr[-1].item()) if n_slices else 0
        sc = torch.empty(total, dtype=torch.int32, device=dev)
        sv = torch.empty(total, dtype=torch.float64, device=dev)
        if total:
            nat.check(nat.lib().fgk_sell_fill(
                self.n_rows, nat.ptr(self.row_ptr, torch.int64), nat.ptr(self.cols, torch.int32),
                nat.ptr(self.vals, torch.float64), nat.ptr(slice_ptr, torch.int64),
                nat.ptr(sc, torch.int32), nat.ptr(sv, torch.float64),
                nat.device_index(self.device), nat.stream_ptr(self.device)))
        self._sell = (slice_ptr, sc, sv)
        self._nnz = self.nnz
        if not keep_csr:
            self._diag_cache = self.diagonal().clone()
            self.cols = self.cols[:0]
            self.vals = self.vals[:0]
            self.sell_only = True
        return self

    def to_sell_packed(self, keep_csr=True):
        """SELL-32 with exact-float32 off-diagonal storage + FP64 diagonal (fgk_sell_pack_f32):
        8 B per nonzero instead of 12, FP64 arithmetic on the same numbers (bit-identical
        products).  Possible whenever every off-diagonal value is a float32 number -- always for
        H_RAW, and for H_SYM when <i|H|j> = <j|H|i> (symmetric integrals).  Raises if not."""
        if getattr(self, "_sellf", None) is not None:
            return self
        self._need_csr("to_sell_packed")
        dev = self.cols.device
        n_slices = (self.n_rows + 31) // 32
        lens = torch.zeros(n_slices * 32, dtype=torch.int64, device=dev)
        lens[: self.n_rows] = self.row_ptr[1:] - self.row_ptr[:-1]
        width = (lens.view(n_slices, 32).max(dim=1).values + 1) // 2          # pair-columns per slice
        slice_ptr = torch.zeros(n_slices + 1, dtype=torch.int64, device=dev)  # in 16-byte units
        torch.cumsum(width * 32, 0, out=slice_ptr[1:])
        total = int(slice_ptr[-1].item()) if n_slices else 0
        packed = torch.empty(max(total, 1), 4, dtype=torch.int32, device=dev)
        diag = torch.empty(self.n_rows, dtype=torch.float64, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.n_rows:
            nat.check(nat.lib().fgk_sell_pack_f32(
                self.n_rows, self.row_begin, nat.ptr(self.row_ptr, torch.int64),
                nat.ptr(self.cols, torch.int32), nat.ptr(self.vals, torch.float64),
                nat.ptr(slice_ptr, torch.int64), nat.ptr(packed), nat.ptr(diag, torch.float64),
                nat.ptr(flag, torch.int32), nat.device_index(self.device), nat.stream_ptr(self.device)))
        if int(flag.item()):
            raise RuntimeError("to_sell_packed: some off-diagonal values are not float32-exact "
                               "(non-symmetric integrals with H_SYM?); use to_sell()")
        self._sellf = (slice_ptr, packed, diag)
        self._nnz = self.nnz
        if not keep_csr:
            self._diag_cache = diag
            self.cols = self.cols[:0]
            self.vals = self.vals[:0]
            self.sell_only = True
        return self

    def optimize_for_matvec(self, min_rows=16384):
        """Pick the fastest storage for repeated H.v: packed SELL-32 when every off-diagonal is
        float32-exact, else SELL-32; operators below `min_rows` rows stay on CSR (launch bound)."""
        if self.n_rows < min_rows or getattr(self, "sell_only", False):
            return self
        try:
            return self.to_sell_packed()
        except RuntimeError:
            return self.to_sell()

    def matvec(self, x, out=None, fmt=None):
        """y = H[row_begin:row_end, :] @ x ; x real FP64 or complex128, length n.
        Uses the packed SELL-32 copy if to_sell_packed() was called, else the SELL-32 copy of
        to_sell(), else CSR (fmt = 'packed' | 'sell' | 'csr' forces one)."""
        if x.shape[0] != self.n:
            raise ValueError(f"matvec: x has {x.shape[0]} entries, H has {self.n} columns")
        dev = nat.device_index(self.device)
        sellf = getattr(self, "_sellf", None)
        if sellf is not None and fmt in (None, "packed"):
            cplx = x.is_complex()
            x = x.to(torch.complex128 if cplx else torch.float64).contiguous()
            y = out if out is not None else torch.empty(self.n_rows, dtype=x.dtype, device=x.device)
            fn = nat.lib().fgk_spmv_sell_f32_z if cplx else nat.lib().fgk_spmv_sell_f32_f64
            nat.check(fn(self.n_rows, self.row_begin, nat.ptr(sellf[0], torch.int64), nat.ptr(sellf[1]),
                         nat.ptr(sellf[2], torch.float64), C.c_void_p(x.data_ptr()),
                         C.c_void_p(y.data_ptr()), dev, nat.stream_ptr(self.device)))
            return y
        sell = getattr(self, "_sell", None)
        if sell is not None and fmt in (None, "sell", "packed"):
            cplx = x.is_complex()
            x = x.to(torch.complex128 if cplx else torch.float64).contiguous()
            y = out if out is not None else torch.empty(self.n_rows, dtype=x.dtype, device=x.device)
            fn = nat.lib().fgk_spmv_sell_z if cplx else nat.lib().fgk_spmv_sell_f64
            nat.check(fn(self.n_rows, nat.ptr(sell[0], torch.int64), nat.ptr(sell[1], torch.int32),
                         nat.ptr(sell[2], torch.float64), C.c_void_p(x.data_ptr()),
                         C.c_void_p(y.data_ptr()), dev, nat.stream_ptr(self.device)))
            return y
        if x.is_complex():
            if x.dtype != torch.complex128:
                x = x.to(torch.complex128)
            x = x.contiguous()
            y = out if out is not None else torch.empty(self.n_rows, dtype=torch.complex128, device=x.device)
            nat.check(nat.lib().fgk_spmv_z(
                self.n_rows, nat.ptr(self.row_ptr, torch.int64), nat.ptr(self.cols, torch.int32),
                nat.ptr(self.vals, torch.float64), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()),
                dev, nat.stream_ptr(self.device)))
            return y
        if x.dtype != torch.float64:
            x = x.to(torch.float64)
        x = x.contiguous()
        y = out if out is not None else torch.empty(self.n_rows, dtype=torch.float64, device=x.device)
        nat.check(nat.lib().fgk_spmv_f64(
            self.n_rows, nat.ptr(self.row_ptr, torch.int64), nat.ptr(self.cols, torch.int32),
            nat.ptr(self.vals, torch.float64), nat.ptr(x, torch.float64), nat.ptr(y, torch.float64),
            dev, nat.stream_ptr(self.device)))
        return y

    def matvec_host(self, x_host, out=None):
        """Host-buffer form of matvec: x_host (CPU tensor or numpy array, ideally pinned) is
        copied to the device, y = H x is computed, and y comes back in a CPU tensor (pinned,
        reused between calls unless `out` is given).  The call returns when y is complete."""
        if not torch.is_tensor(x_host):
            x_host = torch.from_numpy(np.ascontiguousarray(x_host))
        dev = self.row_ptr.device
        key = (x_host.dtype, x_host.shape[0])
        buf = getattr(self, "_host_io", None)
        if buf is None or buf[0] != key:
            ydt = torch.complex128 if x_host.is_complex() else torch.float64
            buf = (key, torch.empty(self.n, dtype=ydt, device=dev),
                   torch.empty(self.n_rows, dtype=ydt, device=dev),
                   torch.empty(self.n_rows, dtype=ydt).pin_memory())
            self._host_io = buf
        _, xd, yd, yh = buf
        xd.copy_(x_host, non_blocking=True)
        self.matvec(xd, out=yd)
        dst = out if out is not None else yh
        dst.copy_(yd, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return dst

    def diagonal(self):
        """the diagonal is the first entry of every row until sort_rows()."""
        if getattr(self, "_diag_cache", None) is not None:
            return self._diag_cache
        if self.sorted_rows:
            rows = torch.repeat_interleave(
                torch.arange(self.row_begin, self.row_end, device=self.cols.device),
                self.row_ptr[1:] - self.row_ptr[:-1])
            m = self.cols.long() == rows
            return self.vals[m]
        return self.vals[self.row_ptr[:-1]]

    def to_dense(self):
        self._need_csr("to_dense")
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.cols.device),
                                       self.row_ptr[1:] - self.row_ptr[:-1])
        D = torch.zeros(self.n_rows, self.n, dtype=torch.float64, device=self.cols.device)
        D[rows, self.cols.long()] = self.vals
        return D

    def to_scipy(self, dtype=np.float64):
        import scipy.sparse as sp
        self.sort_rows()
        M = sp.csr_matrix((self.vals.cpu().numpy().astype(dtype), self.cols.cpu().numpy(),
                           self.row_ptr.cpu().numpy()), shape=(self.n_rows, self.n))
        return M

    def packed_to_coo(self):
        """(rows, cols, vals) of the packed SELL-32 operator, diagonal entries included (local row
        ids; FP64 values).  Export / parity helper: plain tensor indexing, not a hot path."""
        sp, pk, dg = self._sellf
        dev = pk.device
        n = self.n_rows
        lens = getattr(self, "_row_len", None)
        r = torch.arange(n, device=dev)
        if lens is None:                        # to_sell_packed(): off-diagonal count = CSR length - 1
            lens = (self.row_ptr[1:] - self.row_ptr[:-1] - 1).to(torch.int64)
        lens = lens.to(torch.int64)
        rows = torch.repeat_interleave(r, lens)
        start = torch.cumsum(lens, 0) - lens
        k = torch.arange(rows.shape[0], device=dev) - start[rows]
        unit = sp[rows // 32] + (k // 2) * 32 + rows % 32
        half = k % 2
        flat = pk.view(-1)
        vals = flat[unit * 4 + half].view(torch.float32).double()
        cols = flat[unit * 4 + 2 + half].long()
        rows = torch.cat([r, rows])
        cols = torch.cat([r + self.row_begin, cols])
        vals = torch.cat([dg, vals])
        return rows, cols, vals

    def bytes_per_matvec(self, complex_x=False):
        """algorithmic HBM bytes of one product: 12 B/nnz + 20 (36) B/row."""
        return 12 * self.nnz + (36 if complex_x else 20) * self.n_rows


class BasisIndex:
    """Device hash index of a packed basis (K4)."""

    def __init__(self, dets):
        self.dets = dets.contiguous()          # kept alive: the table stores indices into it
        self.device = dets.device
        h = C.c_void_p()
        nat.check(nat.lib().fgk_index_create(
            nat.ptr(self.dets, torch.int64), self.dets.shape[0], nat.device_index(self.device),
            nat.stream_ptr(self.device), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            for h in getattr(self, "_lists", {}).values():
                nat.lib().fgk_strlists_destroy(h)
            self._lists = {}
            if getattr(self, "_h", None):
                nat.lib().fgk_index_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __len__(self):
        return self.dets.shape[0]

    def string_lists(self, ham):
        """single / double replacement lists of the basis' distinct strings for `ham`
        (fgk_strlists_create), built once per (index, Hamiltonian) pair"""
        cache = self.__dict__.setdefault("_lists", {})
        key = id(ham)
        if key not in cache:
            h = C.c_void_p()
            nat.check(nat.lib().fgk_strlists_create(ham._h, self._h, nat.stream_ptr(self.device), C.byref(h)))
            cache[key] = h
        return cache[key]

    def lookup(self, query):
        query = query.contiguous()
        out = torch.empty(query.shape[0], dtype=torch.int32, device=self.device)
        nat.check(nat.lib().fgk_index_lookup(self._h, nat.ptr(query, torch.int64), query.shape[0],
                                             nat.ptr(out, torch.int32), nat.stream_ptr(self.device)))
        return out

    def info(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        nat.check(nat.lib().fgk_index_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        dense = C.c_int(0)
        nat.check(nat.lib().fgk_index_layout(self._h, C.byref(dense)))
        return dict(n_dets=a.value, n_alpha_strings=b.value, n_beta_strings=c.value,
                    dense_pairs=bool(dense.value))


def popcount64(x: torch.Tensor) -> torch.Tensor:
    """number of set bits of every int64 element (SWAR; works on CPU and CUDA tensors)."""
    m1, m2, m4 = 0x5555555555555555, 0x3333333333333333, 0x0F0F0F0F0F0F0F0F
    x = x - ((x >> 1) & m1)
    x = (x & m2) + ((x >> 2) & m2)
    x = (x + (x >> 4)) & m4
    return ((x * 0x0101010101010101) >> 56) & 0xFF


def sort_unique_dets(dets, n_orb):
    """Sorted set of packed determinants: ascending (alpha, beta) as unsigned
    128-bit == the row order of torch.unique(configs, dim=0)
    (residual_expansion.py:368, skqd.py:942)."""
    if dets.shape[0] == 0:
        return dets
    if n_orb == 64:      # bit 63 in use: map unsigned order onto torch's signed order
        flip = torch.tensor(-2 ** 63, dtype=torch.int64, device=dets.device)
        return torch.unique(dets ^ flip, dim=0) ^ flip
    return torch.unique(dets, dim=0)


class MolecularHamiltonian:
    """Drop-in for reference `MolecularHamiltonian(integrals, device)`
    (molecular.py:57-61).  `device` must be a CUDA device."""

    def __init__(self, integrals, device: str = "cuda"):
        self.num_sites = 2 * integrals.n_orbitals          # base.py:21-24
        self.local_dim = 2
        self.hilbert_dim = 2 ** self.num_sites
        self._dev_index = nat.device_index(device)          # raises without CUDA
        self.device = f"cuda:{self._dev_index}"
        self.integrals = integrals
        self.nuclear_repulsion = float(integrals.nuclear_repulsion)
        self.n_orbitals = int(integrals.n_orbitals)
        self.n_electrons = int(integrals.n_electrons)
        self.n_alpha = int(integrals.n_alpha)
        self.n_beta = int(integrals.n_beta)
        h1 = np.ascontiguousarray(integrals.h1e, dtype=np.float64)
        g = np.ascontiguousarray(integrals.h2e, dtype=np.float64)
        if h1.shape != (self.n_orbitals,) * 2 or g.shape != (self.n_orbitals,) * 4:
            raise ValueError("integral shapes do not match n_orbitals")
        # float32 device copies under the reference's attribute names (molecular.py:68-69)
        self.h1e = torch.from_numpy(h1).float().to(self.device)
        self.h2e = torch.from_numpy(g).float().to(self.device)
        self.output_dtype = torch.float64      # set to torch.float32 for reference-typed outputs
        h = C.c_void_p()
        nat.check(nat.lib().fgk_ham_create(
            h1.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), self.n_orbitals,
            self.n_alpha, self.n_beta, self.nuclear_repulsion, self._dev_index, C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                nat.lib().fgk_ham_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- packing (K1) -------------------------------------------------------------
    def pack(self, configs: torch.Tensor) -> torch.Tensor:
        """(n, num_sites) 0/1 any dtype/device -> (n, 2) int64 words {alpha, beta} on device."""
        if configs.dim() == 1:
            configs = configs.unsqueeze(0)
        if configs.shape[1] != self.num_sites:
            raise ValueError(f"configs have {configs.shape[1]} sites, expected {self.num_sites}")
        c = configs.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(c.shape[0], 2, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().fgk_pack_i64(nat.ptr(c, torch.int64), c.shape[0], self.n_orbitals,
                                         nat.ptr(out, torch.int64), self._dev_index,
                                         nat.stream_ptr(self.device)))
        return out

    def unpack(self, dets: torch.Tensor, dtype=torch.int64) -> torch.Tensor:
        dets = dets.contiguous()
        out = torch.empty(dets.shape[0], self.num_sites, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().fgk_unpack_i64(nat.ptr(dets, torch.int64), dets.shape[0], self.n_orbitals,
                                           nat.ptr(out, torch.int64), self._dev_index,
                                           nat.stream_ptr(self.device)))
        return out if dtype == torch.int64 else out.to(dtype)

    # ---- diagonal (K2) ---------------------------------------------------------------
    def diag_packed(self, dets):
        dets = dets.contiguous()
        out = torch.empty(dets.shape[0], dtype=torch.float64, device=self.device)
        nat.check(nat.lib().fgk_diag(self._h, nat.ptr(dets, torch.int64), dets.shape[0],
                                     nat.ptr(out, torch.float64), nat.stream_ptr(self.device)))
        return out

    @torch.no_grad()
    def diagonal_elements_batch(self, configs: torch.Tensor) -> torch.Tensor:
        """molecular.py:133-184.  (batch, num_sites) -> (batch,)."""
        return self.diag_packed(self.pack(configs)).to(self.output_dtype)

    def diagonal_element(self, config: torch.Tensor) -> torch.Tensor:
        """molecular.py:186-192."""
        return self.diagonal_elements_batch(config.unsqueeze(0))[0]

    # ---- connections (K3) --------------------------------------------------------------
    def connections_packed(self, dets, want_dets=True, want_src=True):
        """-> (out_dets (N,2) int64, elems (N,) float32, src (N,) int64, offsets (n+1,) int64),
        connections of dets[j] at offsets[j]:offsets[j+1] in the reference's emission order."""
        dets = dets.contiguous()
        n = dets.shape[0]
        st = nat.stream_ptr(self.device)
        counts = torch.empty(n, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().fgk_conn_count(self._h, nat.ptr(dets, torch.int64), n,
                                           nat.ptr(counts, torch.int64), st))
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        total = int(offsets[-1].item()) if n else 0
        out_dets = torch.empty(total, 2, dtype=torch.int64, device=self.device) if want_dets else None
        elems = torch.empty(total, dtype=torch.float32, device=self.device)
        src = torch.empty(total, dtype=torch.int64, device=self.device) if want_src else None
        if total:
            nat.check(nat.lib().fgk_conn_fill(
                self._h, nat.ptr(dets, torch.int64), n, nat.ptr(offsets, torch.int64),
                nat.ptr(out_dets), nat.ptr(elems, torch.float32), nat.ptr(src), st))
        return out_dets, elems, src, offsets

    def get_connections(self, config: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """molecular.py:194-327: (connected (N, num_sites) in config's dtype, elements (N,) float32);
        empty results are shaped tensors (molecular.py:320-321)."""
        dets = self.pack(config)
        od, el, _, _ = self.connections_packed(dets, want_src=False)
        if od.shape[0] == 0:
            return torch.empty(0, self.num_sites, device=self.device), torch.empty(0, device=self.device)
        return self.unpack(od, config.dtype), el

    @torch.no_grad()
    def get_connections_batch(self, configs: torch.Tensor):
        """The hook utils/connection_cache.py:250-254 looks for (the reference never
        implements it): (all_connected, all_elements, config_indices)."""
        if configs.shape[0] == 0:
            return (torch.empty(0, self.num_sites, device=self.device),
                    torch.empty(0, device=self.device),
                    torch.empty(0, dtype=torch.long, device=self.device))
        od, el, src, _ = self.connections_packed(self.pack(configs))
        if od.shape[0] == 0:
            return (torch.empty(0, self.num_sites, device=self.device),
                    torch.empty(0, device=self.device),
                    torch.empty(0, dtype=torch.long, device=self.device))
        return self.unpack(od, configs.dtype), el, src

    def get_all_connections_with_indices(self, configs):
        """molecular.py:329-377."""
        return self.get_connections_batch(configs)

    def get_connections_parallel(self, configs, max_workers: int = 8):
        """molecular.py:518-578 (there: a thread pool with arbitrary completion order;
        here: one launch, source-ascending order)."""
        return self.get_connections_batch(configs)

    @torch.no_grad()
    def local_energies(self, configs: torch.Tensor, log_amplitude, max_connections=8_000_000,
                       nqs_chunk_size=16384) -> torch.Tensor:
        """Stage-1 local energies E_loc(x) = <x|H|x> + sum_x' elem(x -> x') psi(x') / psi(x), the
        quantity flows/physics_guided_training.py:335-457 assembles from
        diagonal_elements_batch + get_connections(_parallel) + scatter_add -- here with bounded
        memory: the batch is walked in groups whose connections (counted first with
        fgk_conn_count) fit `max_connections`, connections stay packed until the amplitude
        network needs them, and only `nqs_chunk_size` rows are unpacked at a time.
        log_amplitude: callable (m, num_sites) float32 -> (m,) log psi (real or complex)."""
        dets = self.pack(configs)
        n = dets.shape[0]
        out = self.diag_packed(dets).clone()
        if n == 0:
            return out
        st = nat.stream_ptr(self.device)
        counts = torch.empty(n, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().fgk_conn_count(self._h, nat.ptr(dets, torch.int64), n,
                                           nat.ptr(counts, torch.int64), st))
        csum = torch.cumsum(counts, 0).cpu()
        log_psi0 = log_amplitude(configs.to(self.device).float()).clone()
        off = torch.zeros(n, dtype=log_psi0.dtype if log_psi0.is_complex() else torch.float64,
                          device=self.device)
        lo = 0
        while lo < n:
            base = int(csum[lo - 1]) if lo else 0
            hi = int(torch.searchsorted(csum, base + max_connections, right=True))
            hi = max(hi, lo + 1)
            od, el, src, _ = self.connections_packed(dets[lo:hi].contiguous())
            m = od.shape[0]
            if m:
                lp = torch.empty(m, dtype=log_psi0.dtype, device=self.device)
                for a in range(0, m, nqs_chunk_size):
                    b = min(m, a + nqs_chunk_size)
                    lp[a:b] = log_amplitude(self.unpack(od[a:b]).float())
                ratio = torch.exp(lp - log_psi0[lo:hi][src])
                off[lo:hi].index_add_(0, src, (el.to(torch.float64) * ratio).to(off.dtype))
            lo = hi
        res = out + off
        return res.real if res.is_complex() else res

    def _require_particle_numbers(self, index: "BasisIndex", what: str):
        """The row builders and the PT2 walk size their per-warp excitation lists from this
        Hamiltonian's n_alpha / n_beta: every determinant of an indexed basis must carry exactly
        those particle numbers (the pipeline's bases always do).  get_connections,
        diagonal_elements_batch and matrix_elements accept any occupation, like the reference.
        Checked once per index."""
        key = (self.n_alpha, self.n_beta)
        if getattr(index, "_particles_ok", None) == key or len(index) == 0:
            return
        cnt = popcount64(index.dets)
        bad = (cnt[:, 0] != self.n_alpha) | (cnt[:, 1] != self.n_beta)
        if bool(bad.any()):
            i = int(torch.nonzero(bad)[0])
            raise ValueError(f"{what}: determinant {i} of the basis has {int(cnt[i, 0])}+{int(cnt[i, 1])} "
                             f"electrons, the Hamiltonian has {self.n_alpha}+{self.n_beta}")
        index._particles_ok = key

    # ---- projected Hamiltonian (K4 + K5) ---------------------------------------------------
    def projected_csr(self, basis, mode=nat.H_RAW, row_begin=0, row_end=None, sort_rows=False,
                      index: Optional[BasisIndex] = None, packed=False, profile=False) -> ProjectedH:
        """CSR rows [row_begin,row_end) of <i|H|j> over `basis` (configs, or packed
        words if packed=True).  mode: H_RAW | H_SYM [| H_DROP_ZEROS].  Rows come in the
        builder's (deterministic) emission order with the diagonal first; H.v does not
        care.  sort_rows=True / .sort_rows() / .to_scipy() order them by column."""
        dets = basis if packed else self.pack(basis)
        idx = index if index is not None else BasisIndex(dets)
        self._require_particle_numbers(idx, "projected_csr")
        n = len(idx)
        row_end = n if row_end is None else row_end
        rows = row_end - row_begin
        st = nat.stream_ptr(self.device)
        import time as _time
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if profile else None
        counts = torch.empty(rows, dtype=torch.int64, device=self.device)
        if profile:
            ev[0].record()
        nat.check(nat.lib().fgk_projh_count(self._h, idx._h, row_begin, row_end, mode,
                                            nat.ptr(counts, torch.int64), st))
        if profile:
            ev[1].record()
        row_ptr = torch.zeros(rows + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=row_ptr[1:])
        nnz = int(row_ptr[-1].item()) if rows else 0
        t0 = _time.perf_counter()
        cols = torch.empty(nnz, dtype=torch.int32, device=self.device)
        vals = torch.empty(nnz, dtype=torch.float64, device=self.device)
        t_alloc = _time.perf_counter() - t0
        if profile:
            ev[2].record()
        if nnz:
            nat.check(nat.lib().fgk_projh_fill(self._h, idx._h, row_begin, row_end, mode,
                                               nat.ptr(row_ptr, torch.int64),
                                               nat.ptr(cols, torch.int32),
                                               nat.ptr(vals, torch.float64), st))
        P = ProjectedH(n, row_ptr, cols, vals, self.device, row_begin, row_end, mode)
        P._index = idx
        P._ham = self
        if profile:
            ev[3].record()
            ev[3].synchronize()
            P.build_profile = {"count_ms": ev[0].elapsed_time(ev[1]), "fill_ms": ev[2].elapsed_time(ev[3]),
                               "alloc_ms": 1e3 * t_alloc}
        return P.sort_rows() if sort_rows else P

    def projected_sell(self, basis, mode=nat.H_SYM, row_begin=0, row_end=None,
                       index: Optional[BasisIndex] = None, packed=False) -> ProjectedH:
        """Rows [row_begin,row_end) of the projected H built STRAIGHT into SELL-32 storage
        (count -> slice widths -> fgk_projh_fill_sell): the operator for Krylov work, at half the
        peak memory of projected_csr(...).to_sell() -- but the lane-interleaved layout makes the
        fill a strided write (measured 2.5x slower than CSR fill + conversion on config 4), so use it
        when memory, not build time, is the limit.  matvec / diagonal / nnz work; the CSR-only
        views (to_dense, to_scipy, sort_rows) are not available on it."""
        dets = basis if packed else self.pack(basis)
        idx = index if index is not None else BasisIndex(dets)
        self._require_particle_numbers(idx, "projected_sell")
        n = len(idx)
        row_end = n if row_end is None else row_end
        rows = row_end - row_begin
        st = nat.stream_ptr(self.device)
        counts = torch.empty(rows, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().fgk_projh_count(self._h, idx._h, row_begin, row_end, mode,
                                            nat.ptr(counts, torch.int64), st))
        row_ptr = torch.zeros(rows + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=row_ptr[1:])
        n_slices = (rows + 31) // 32
        lens = torch.zeros(n_slices * 32, dtype=torch.int64, device=self.device)
        lens[:rows] = counts
        width = (lens.view(n_slices, 32).max(dim=1).values + 1) // 2 * 2
        slice_ptr = torch.zeros(n_slices + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(width * 32, 0, out=slice_ptr[1:])
        total = int(slice_ptr[-1].item()) if n_slices else 0
        sc = torch.zeros(total, dtype=torch.int32, device=self.device)
        sv = torch.zeros(total, dtype=torch.float64, device=self.device)
        if total:
            nat.check(nat.lib().fgk_projh_fill_sell(
                self._h, idx._h, row_begin, row_end, mode, nat.ptr(slice_ptr, torch.int64),
                nat.ptr(sc, torch.int32), nat.ptr(sv, torch.float64), st))
        empty_c = torch.empty(0, dtype=torch.int32, device=self.device)
        empty_v = torch.empty(0, dtype=torch.float64, device=self.device)
        P = ProjectedH(n, row_ptr, empty_c, empty_v, self.device, row_begin, row_end, mode)
        P._index = idx
        P._sell = (slice_ptr, sc, sv)
        P._nnz = int(row_ptr[-1].item()) if rows else 0
        # the diagonal is entry 0 of every row: slice base + 2 * lane
        r = torch.arange(rows, device=self.device)
        P._diag_cache = sv[slice_ptr[r // 32] + 2 * (r % 32)] if rows else empty_v
        P.sell_only = True
        return P

    PACKED_MAX_STRINGS = 200_000      # the replacement lists are found by an all-pairs string scan

    def projected_packed(self, basis, mode=nat.H_SYM, row_begin=0, row_end=None,
                         index: Optional[BasisIndex] = None, packed=False, profile=False) -> ProjectedH:
        """Rows [row_begin,row_end) of the projected H built STRAIGHT into the packed SELL-32
        operator (exact-float32 off-diagonals + FP64 diagonal, 8 B/nnz): no CSR arrays, no
        CSR -> SELL pass, no re-pack (fgk_strlists_create + fgk_projh_packed_*; string-driven,
        one warp per 32-row slice with lane = row).  The operator for Krylov work -- matvec /
        matvec_host / diagonal / nnz / packed_to_coo; the CSR views (to_dense, to_scipy, sort_rows)
        need projected_csr.  Row lengths: a PRODUCT basis (every alpha string paired with every beta
        string: CAS windows, full spaces) is filled straight into storage sized by the bound from the
        list lengths -- no count pass at all; only if filtered values leave more than 2 % padding
        (real molecules: symmetry zeros) it is refilled with the lengths the first fill measured.
        Other bases run the exact count pass.  Raises if a value is not float32-exact
        (non-symmetric integrals with H_SYM): use projected_csr(...).to_sell() then."""
        dets = basis if packed else self.pack(basis)
        idx = index if index is not None else BasisIndex(dets)
        self._require_particle_numbers(idx, "projected_packed")
        n = len(idx)
        row_end = n if row_end is None else row_end
        rows = row_end - row_begin
        dev, st, L = self.device, nat.stream_ptr(self.device), nat.lib()
        if n == 0 or rows <= 0:
            raise ValueError("projected_packed: empty row range")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if profile else None
        if profile:
            ev[0].record()
        lists = idx.string_lists(self)
        if profile:
            ev[1].record()
        info = idx.info()
        product = info["dense_pairs"] and n == info["n_alpha_strings"] * info["n_beta_strings"]
        cnt = torch.empty(rows, dtype=torch.int64, device=dev)
        if product:     # every string pair is a determinant: the bound can only be missed through filtered values
            nat.check(L.fgk_projh_packed_bound(self._h, idx._h, lists, row_begin, row_end, nat.ptr(cnt, torch.int64), st))
        else:
            nat.check(L.fgk_projh_packed_count(self._h, idx._h, lists, row_begin, row_end, mode, 1,
                                               nat.ptr(cnt, torch.int64), st))
        if profile:
            ev[2].record()
        n_slices = (rows + 31) // 32
        diag = self.diag_packed(dets[row_begin:row_end])
        count_pass = "none (list-length bound of a product basis)" if product else "exact"
        for attempt in range(2):
            lens = torch.zeros(n_slices * 32, dtype=torch.int64, device=dev)
            lens[:rows] = cnt
            width = (lens.view(n_slices, 32).max(dim=1).values + 1) // 2           # pair-columns per slice
            slice_ptr = torch.zeros(n_slices + 1, dtype=torch.int64, device=dev)   # in 16-byte units
            torch.cumsum(width * 32, 0, out=slice_ptr[1:])
            total = int(slice_ptr[-1].item())
            pk = torch.empty(max(total, 1), 4, dtype=torch.int32, device=dev)
            row_len = torch.empty(rows, dtype=torch.int32, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            if profile and attempt == 0:
                ev[3].record()
            nat.check(L.fgk_projh_packed_fill(self._h, idx._h, lists, row_begin, row_end, mode,
                                              nat.ptr(slice_ptr, torch.int64), nat.ptr(pk), nat.ptr(row_len, torch.int32),
                                              nat.ptr(flag, torch.int32), st))
            # one read-back: inexact / too-narrow flag, and how much of the bound-sized storage is padding
            wasteful = (row_len.sum() * 50 < cnt.sum() * 49).to(torch.int32).reshape(1)
            bad, waste = torch.cat([flag, wasteful]).tolist()
            if bad:
                raise RuntimeError("projected_packed: an off-diagonal value is not float32-exact (non-symmetric "
                                   "integrals with H_SYM?); use projected_csr(...).to_sell()")
            if not (product and waste and attempt == 0):
                break
            cnt = row_len.to(torch.int64)       # > 2 % padding (filtered values): refill with the exact lengths
            count_pass = "the first fill (bound-sized) served as the count pass"
            del pk
        if profile:
            ev[4].record()
        row_ptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
        torch.cumsum(row_len.to(torch.int64) + 1, 0, out=row_ptr[1:])
        empty_c = torch.empty(0, dtype=torch.int32, device=dev)
        empty_v = torch.empty(0, dtype=torch.float64, device=dev)
        P = ProjectedH(n, row_ptr, empty_c, empty_v, self.device, row_begin, row_end, mode)
        P._index = idx
        P._ham = self
        P._sellf = (slice_ptr, pk, diag)
        P._row_len = row_len
        P._diag_cache = diag
        P._nnz = row_ptr[-1]
        P.sell_only = True
        P.count_pass = count_pass
        if profile:
            ev[4].synchronize()
            P.build_profile = {"lists_ms": ev[0].elapsed_time(ev[1]), "bound_count_ms": ev[1].elapsed_time(ev[2]),
                               "alloc_diag_ms": ev[2].elapsed_time(ev[3]), "fill_ms": ev[3].elapsed_time(ev[4]),
                               "count_pass": P.count_pass}
        return P

    def projected_operator(self, basis, mode=nat.H_SYM, row_begin=0, row_end=None,
                           index: Optional[BasisIndex] = None, packed=False, min_rows=16384) -> ProjectedH:
        """The projected H in the best storage for repeated H.v (Krylov drivers): built straight
        into the packed SELL-32 form when the basis is large enough to be bandwidth-bound and its
        values are float32-exact; CSR (+ optimize_for_matvec) otherwise."""
        dets = basis if packed else self.pack(basis)
        idx = index if index is not None else BasisIndex(dets)
        n = len(idx)
        re_ = n if row_end is None else row_end
        if re_ - row_begin >= min_rows:
            info = idx.info()
            if max(info["n_alpha_strings"], info["n_beta_strings"]) <= self.PACKED_MAX_STRINGS:
                try:
                    return self.projected_packed(dets, mode, row_begin, row_end, index=idx, packed=True)
                except RuntimeError as e:
                    if "float32-exact" not in str(e):
                        raise
        P = self.projected_csr(dets, mode, row_begin, row_end, index=idx, packed=True, sort_rows=False)
        return P.optimize_for_matvec(min_rows=min_rows)

    @torch.no_grad()
    def matrix_elements_fast(self, configs: torch.Tensor) -> torch.Tensor:
        """molecular.py:471-516: dense (n, n), H[i, j] = <i|H|j> raw directed."""
        if configs.shape[0] == 0:
            return torch.zeros(0, 0, dtype=self.output_dtype, device=self.device)
        return self.projected_csr(configs, nat.H_RAW, sort_rows=False).to_dense().to(self.output_dtype)

    def matrix_elements(self, configs_bra: torch.Tensor, configs_ket: torch.Tensor) -> torch.Tensor:
        """molecular.py:640-685."""
        if configs_bra.shape == configs_ket.shape and bool(
                torch.all(configs_bra.to(self.device) == configs_ket.to(self.device))):
            return self.matrix_elements_fast(configs_bra)
        bra = self.pack(configs_bra)
        ket = self.pack(configs_ket)
        idx = BasisIndex(bra)
        H = torch.zeros(bra.shape[0], ket.shape[0], dtype=torch.float64, device=self.device)
        hit = idx.lookup(ket).long()                       # diagonal, :671-674
        kk = torch.nonzero(hit >= 0).squeeze(1)
        if kk.numel():
            H[hit[kk], kk] = self.diag_packed(ket[kk])
        od, el, src, _ = self.connections_packed(ket)      # off-diagonal, :677-683
        if od.shape[0]:
            i = idx.lookup(od).long()
            m = i >= 0
            H[i[m], src[m]] = el[m].double()
        return H.to(self.output_dtype)

    @torch.no_grad()
    def get_sparse_matrix_elements(self, configs: torch.Tensor):
        """molecular.py:580-638: COO (rows, cols, values) of the OFF-diagonal hits,
        j ascending, emission order inside j."""
        if configs.shape[0] == 0:
            return (torch.tensor([], dtype=torch.long, device=self.device),
                    torch.tensor([], dtype=torch.long, device=self.device),
                    torch.tensor([], dtype=torch.float32, device=self.device))
        dets = self.pack(configs)
        idx = BasisIndex(dets)
        od, el, src, _ = self.connections_packed(dets)
        i = idx.lookup(od).long() if od.shape[0] else torch.empty(0, dtype=torch.long, device=self.device)
        m = i >= 0
        return i[m], src[m], el[m]

    # ---- misc reference API -------------------------------------------------------------------
    def get_hf_state(self) -> torch.Tensor:
        """molecular.py:778-792."""
        config = torch.zeros(self.num_sites, dtype=torch.long, device=self.device)
        config[: self.n_alpha] = 1
        config[self.n_orbitals: self.n_orbitals + self.n_beta] = 1
        return config

    def _config_to_index(self, config: torch.Tensor) -> int:
        """base.py: big-endian integer of the configuration."""
        idx = 0
        for b in config.tolist():
            idx = (idx << 1) | int(b)
        return idx

    def fci_dets(self) -> torch.Tensor:
        """all C(n,na)*C(n,nb) determinants, packed, in the reference's order
        (itertools.combinations, alpha-major: skqd.py:155-169, molecular.py:894-905)."""
        n = self.n_orbitals

        def strings(k):
            out = np.zeros(comb(n, k), dtype=np.uint64)
            for i, occ in enumerate(combinations(range(n), k)):
                w = 0
                for p in occ:
                    w |= 1 << (n - 1 - p)
                out[i] = w
            return out

        a, b = strings(self.n_alpha), strings(self.n_beta)
        dets = np.empty((len(a), len(b), 2), dtype=np.uint64)
        dets[:, :, 0] = a[:, None]
        dets[:, :, 1] = b[None, :]
        return torch.from_numpy(dets.reshape(-1, 2).view(np.int64)).to(self.device)

    def fci_energy(self) -> float:
        """molecular.py:872-942: lowest eigenvalue of the symmetrised projected H over
        the full particle-conserving space."""
        from .solvers import lowest_eigenpairs
        dets = self.fci_dets()
        P = self.projected_operator(dets, nat.H_SYM, packed=True)
        w, _ = lowest_eigenpairs(P, k=1)
        return float(w[0])
