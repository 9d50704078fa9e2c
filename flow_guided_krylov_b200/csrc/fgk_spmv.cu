// fgk_spmv.cu -- K6: FP64 CSR sparse H.v (real and complex right-hand side).
//
// HBM-bound: 12 B per nonzero (8 B value + 4 B column) + 20 B per row
// (8 B y write, 8 B compulsory x read, 4 B row_ptr amortised)  [SURVEY 8(d)].
// One warp per row ("CSR-vector"): the row's value / column segments are
// streamed with 128-bit / 64-bit loads that bypass L1 (ld.global.nc
// L1::no_allocate), two nonzeros per lane per load, UNROLL loads in flight per
// lane; x is gathered through the read-only path and is L2-resident (8-16 MB at
// 1e6 rows against 126 MB of L2).  Replaces scipy's csr_matvec under eigsh
// (skqd.py:784, residual_expansion.py:435) and expm_multiply (skqd.py:291-293).
#include <stdlib.h>

#include "fgk_internal.cuh"

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ int2 ld_stream_s32x2(const int32_t* p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

template <bool CPLX>
struct Acc;
template <>
struct Acc<false> {
    double re;
    __device__ __forceinline__ Acc() : re(0.0) {}
    __device__ __forceinline__ void fma(double v, const double* x, int c) { re = ::fma(v, __ldg(x + c), re); }
    __device__ __forceinline__ void reduce() { re = warp_sum(re); }
    __device__ __forceinline__ void store(double* y, i64 r) const { y[r] = re; }
    static __device__ __forceinline__ double imag(const Acc&) { return 0.0; }
};
template <>
struct Acc<true> {
    double re, im;
    __device__ __forceinline__ Acc() : re(0.0), im(0.0) {}
    __device__ __forceinline__ void fma(double v, const double* x, int c)
    {
        double2 z = __ldg(reinterpret_cast<const double2*>(x) + c);
        re = ::fma(v, z.x, re);
        im = ::fma(v, z.y, im);
    }
    __device__ __forceinline__ void reduce() { re = warp_sum(re); im = warp_sum(im); }
    __device__ __forceinline__ void store(double* y, i64 r) const
    {
        reinterpret_cast<double2*>(y)[r] = make_double2(re, im);
    }
    static __device__ __forceinline__ double imag(const Acc& a) { return a.im; }
};

template <bool CPLX, int UNROLL>
__global__ void __launch_bounds__(FGK_BLOCK)
k_spmv_csr_vector(i64 n_rows, const i64* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                  const double* __restrict__ vals, const double* __restrict__ x,
                  double* __restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    for (i64 r = warp0; r < n_rows; r += nwarps) {
        const i64 s = __ldg(row_ptr + r), e = __ldg(row_ptr + r + 1);
        Acc<CPLX> acc;
        // peel to even indices so that value pairs are 16-byte and column pairs 8-byte aligned
        i64 s2 = (s + 1) & ~1ll;
        if (s2 > e) s2 = e;
        i64 e2 = e & ~1ll;
        if (e2 < s2) e2 = s2;
        if (lane == 0 && s < s2) acc.fma(__ldg(vals + s), x, __ldg(cols + s));
        if (lane == 1 && e2 < e) acc.fma(__ldg(vals + e2), x, __ldg(cols + e2));
        i64 k = s2 + 2 * lane;
        // main loop: UNROLL independent 24-byte loads per lane before the first use
        for (; k + 64 * (UNROLL - 1) < e2; k += 64 * UNROLL) {
            double2 v[UNROLL];
            int2 c[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                v[u] = ld_stream_f64x2(vals + k + 64 * u);
                c[u] = ld_stream_s32x2(cols + k + 64 * u);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                acc.fma(v[u].x, x, c[u].x);
                acc.fma(v[u].y, x, c[u].y);
            }
        }
        for (; k < e2; k += 64) {
            double2 v = ld_stream_f64x2(vals + k);
            int2 c = ld_stream_s32x2(cols + k);
            acc.fma(v.x, x, c.x);
            acc.fma(v.y, x, c.y);
        }
        acc.reduce();
        if (lane == 0) acc.store(y, r);
    }
}

template <bool CPLX>
static int launch_spmv(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols,
                       const double* vals, const double* x, double* y, int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!row_ptr || !x || !y || n_rows < 0) return fgk_fail(FGK_ERR_ARG, "fgk_spmv: bad argument");
    if (((uintptr_t)vals & 15) || ((uintptr_t)cols & 7))
        return fgk_fail(FGK_ERR_ARG, "fgk_spmv: vals must be 16-byte and cols 8-byte aligned");
    FGK_CUDA(cudaSetDevice(device));
    i64 need = (n_rows + FGK_WARPS_PER_BLOCK - 1) / FGK_WARPS_PER_BLOCK;
    i64 cap = (i64)fgk_sm_count(device) * 8;      // 8 resident CTAs of 8 warps = 64 warps / SM
    int grid = (int)(need < cap ? need : cap);
    k_spmv_csr_vector<CPLX, 4><<<grid, FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        n_rows, (const i64*)row_ptr, cols, vals, x, y);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_spmv_f64(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols,
                            const double* vals, const double* x, double* y, int device, void* stream)
{
    return launch_spmv<false>(n_rows, row_ptr, cols, vals, x, y, device, stream);
}

extern "C" int fgk_spmv_z(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols,
                          const double* vals, const double* x, double* y, int device, void* stream)
{
    return launch_spmv<true>(n_rows, row_ptr, cols, vals, x, y, device, stream);
}

// ======================================================================================
// SELL-32: sliced ELLPACK with slice height 32 (one lane per row) and two entries per
// lane per 128-bit load.  Entry k of row r = 32 s + l lives at
//     slice_ptr[s] + ((k >> 1) * 32 + l) * 2 + (k & 1)
// so a warp's value load is one fully coalesced, 16-byte aligned 512 B request and its
// column load a 256 B one, with no per-row peel, no warp reduction and a coalesced y
// store.  More importantly for this path, lane = row makes the 32 x-gathers of one
// instruction fall into the SAME region of x: neighbouring basis determinants connect
// to neighbouring columns (sorted rows), which the one-warp-per-row CSR kernel cannot
// exploit (its 32 lanes walk 32 different columns of a single row).
// One CTA of SELL_WARPS warps per slice; warp w takes pair-columns k2 = w, w+W, ... so
// that the CTA streams one contiguous region; partial sums meet in shared memory.
// The hardware CTA scheduler balances the ~n/32 slices over the SMs.
// ======================================================================================
static const int SELL_WARPS = 4;

template <bool CPLX, int UNROLL>
__global__ void __launch_bounds__(SELL_WARPS * 32)
k_spmv_sell(i64 n_rows, const i64* __restrict__ slice_ptr, const int32_t* __restrict__ cols,
            const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y)
{
    __shared__ double s_part[SELL_WARPS][32][CPLX ? 2 : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 s = blockIdx.x;
    const i64 base = __ldg(slice_ptr + s);
    const i64 npair = (__ldg(slice_ptr + s + 1) - base) >> 6;      // pair-columns in this slice
    const double* v0 = vals + base + 2 * lane;
    const int32_t* c0 = cols + base + 2 * lane;
    Acc<CPLX> acc;
    i64 k2 = w;
    for (; k2 + (i64)SELL_WARPS * (UNROLL - 1) < npair; k2 += (i64)SELL_WARPS * UNROLL) {
        double2 v[UNROLL];
        int2 c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            v[u] = ld_stream_f64x2(v0 + (k2 + (i64)SELL_WARPS * u) * 64);
            c[u] = ld_stream_s32x2(c0 + (k2 + (i64)SELL_WARPS * u) * 64);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            acc.fma(v[u].x, x, c[u].x);
            acc.fma(v[u].y, x, c[u].y);
        }
    }
    for (; k2 < npair; k2 += SELL_WARPS) {
        double2 v = ld_stream_f64x2(v0 + k2 * 64);
        int2 c = ld_stream_s32x2(c0 + k2 * 64);
        acc.fma(v.x, x, c.x);
        acc.fma(v.y, x, c.y);
    }
    s_part[w][lane][0] = acc.re;
    if (CPLX) s_part[w][lane][CPLX ? 1 : 0] = Acc<CPLX>::imag(acc);
    __syncthreads();
    if (w == 0) {
        const i64 r = s * 32 + lane;
        if (r < n_rows) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int q = 0; q < SELL_WARPS; q++) {
                re += s_part[q][lane][0];
                if (CPLX) im += s_part[q][lane][CPLX ? 1 : 0];
            }
            if (CPLX) reinterpret_cast<double2*>(y)[r] = make_double2(re, im);
            else y[r] = re;
        }
    }
}

// CSR -> SELL-32.  One CTA per slice; thread (pair-column k2, lane l) reads row l's
// entries 2*k2, 2*k2+1 (L1 absorbs the row-strided reads) and writes coalesced.
// Padding: value 0, column = 0.
__global__ void __launch_bounds__(256)
k_sell_fill(i64 n_rows, const i64* __restrict__ row_ptr, const int32_t* __restrict__ cols,
            const double* __restrict__ vals, const i64* __restrict__ slice_ptr,
            int32_t* __restrict__ sc, double* __restrict__ sv)
{
    const i64 s = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const i64 base = slice_ptr[s];
    const i64 npair = (slice_ptr[s + 1] - base) >> 6;
    const i64 r = s * 32 + lane;
    i64 rs = 0, len = 0;
    if (r < n_rows) { rs = row_ptr[r]; len = row_ptr[r + 1] - rs; }
    for (i64 k2 = w; k2 < npair; k2 += nw) {
        double2 v = make_double2(0.0, 0.0);
        int2 c = make_int2(0, 0);
        i64 k = 2 * k2;
        if (k < len) { v.x = vals[rs + k]; c.x = cols[rs + k]; }
        if (k + 1 < len) { v.y = vals[rs + k + 1]; c.y = cols[rs + k + 1]; }
        reinterpret_cast<double2*>(sv + base)[k2 * 32 + lane] = v;
        reinterpret_cast<int2*>(sc + base)[k2 * 32 + lane] = c;
    }
}

extern "C" int fgk_sell_fill(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols,
                             const double* vals, const int64_t* slice_ptr, int32_t* sell_cols,
                             double* sell_vals, int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!row_ptr || !slice_ptr || !sell_cols || !sell_vals || n_rows < 0)
        return fgk_fail(FGK_ERR_ARG, "fgk_sell_fill: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    i64 n_slices = (n_rows + 31) / 32;
    k_sell_fill<<<(unsigned)n_slices, 256, 0, (cudaStream_t)stream>>>(
        n_rows, (const i64*)row_ptr, cols, vals, (const i64*)slice_ptr, sell_cols, sell_vals);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

template <bool CPLX>
static int launch_spmv_sell(int64_t n_rows, const int64_t* slice_ptr, const int32_t* cols,
                            const double* vals, const double* x, double* y, int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!slice_ptr || !cols || !vals || !x || !y || n_rows < 0)
        return fgk_fail(FGK_ERR_ARG, "fgk_spmv_sell: bad argument");
    if (((uintptr_t)vals & 15) || ((uintptr_t)cols & 7))
        return fgk_fail(FGK_ERR_ARG, "fgk_spmv_sell: vals must be 16-byte and cols 8-byte aligned");
    FGK_CUDA(cudaSetDevice(device));
    i64 n_slices = (n_rows + 31) / 32;
    static const int unroll = getenv("FGK_SPMV_UNROLL") ? atoi(getenv("FGK_SPMV_UNROLL")) : 4;    // experiment knob
    if (unroll == 8)
        k_spmv_sell<CPLX, 8><<<(unsigned)n_slices, SELL_WARPS * 32, 0, (cudaStream_t)stream>>>(
            n_rows, (const i64*)slice_ptr, cols, vals, x, y);
    else if (unroll == 2)
        k_spmv_sell<CPLX, 2><<<(unsigned)n_slices, SELL_WARPS * 32, 0, (cudaStream_t)stream>>>(
            n_rows, (const i64*)slice_ptr, cols, vals, x, y);
    else
    k_spmv_sell<CPLX, 4><<<(unsigned)n_slices, SELL_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n_rows, (const i64*)slice_ptr, cols, vals, x, y);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_spmv_sell_f64(int64_t n_rows, const int64_t* slice_ptr, const int32_t* sell_cols,
                                 const double* sell_vals, const double* x, double* y, int device,
                                 void* stream)
{
    return launch_spmv_sell<false>(n_rows, slice_ptr, sell_cols, sell_vals, x, y, device, stream);
}

extern "C" int fgk_spmv_sell_z(int64_t n_rows, const int64_t* slice_ptr, const int32_t* sell_cols,
                               const double* sell_vals, const double* x, double* y, int device,
                               void* stream)
{
    return launch_spmv_sell<true>(n_rows, slice_ptr, sell_cols, sell_vals, x, y, device, stream);
}

// ======================================================================================
// SELL-32 with exact-float32 off-diagonal storage ("packed" flavour).
// Every off-diagonal element of the projected H is an exact float32 number (the reference
// keeps float32 integrals: +-h_pq, +-g_pqrs, +-fp32(g - g); the symmetrised flavour too
// whenever <i|H|j> = <j|H|i> or one of them is filtered), only the diagonal is a genuine
// FP64 sum.  Storing {float v0, float v1, int c0, int c1} per lane per pair-column makes one
// 16-byte load carry two nonzeros (8 B/nnz instead of 12) while the arithmetic stays FP64:
// (double)v is exact, products and sums are the same FP64 operations on the same numbers.
// The diagonal lives in its own FP64 array and is applied in the epilogue.
// fgk_sell_pack_f32 refuses (flag) if any off-diagonal value is not float32-exact.
// ======================================================================================
__device__ __forceinline__ uint4 ld_stream_u32x4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <bool CPLX, int UNROLL>
__global__ void __launch_bounds__(SELL_WARPS * 32)
k_spmv_sell_f32(i64 n_rows, i64 row_offset, const i64* __restrict__ slice_ptr,
                const uint4* __restrict__ packed, const double* __restrict__ diag,
                const double* __restrict__ x, double* __restrict__ y)
{
    __shared__ double s_part[SELL_WARPS][32][CPLX ? 2 : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 s = blockIdx.x;
    const i64 base = __ldg(slice_ptr + s);                 // in 16-byte units
    const i64 npair = (__ldg(slice_ptr + s + 1) - base) >> 5;
    const uint4* p0 = packed + base + lane;
    Acc<CPLX> acc;
    i64 k2 = w;
    for (; k2 + (i64)SELL_WARPS * (UNROLL - 1) < npair; k2 += (i64)SELL_WARPS * UNROLL) {
        uint4 q[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) q[u] = ld_stream_u32x4(p0 + (k2 + (i64)SELL_WARPS * u) * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            acc.fma((double)__uint_as_float(q[u].x), x, (int)q[u].z);
            acc.fma((double)__uint_as_float(q[u].y), x, (int)q[u].w);
        }
    }
    for (; k2 < npair; k2 += SELL_WARPS) {
        uint4 q = ld_stream_u32x4(p0 + k2 * 32);
        acc.fma((double)__uint_as_float(q.x), x, (int)q.z);
        acc.fma((double)__uint_as_float(q.y), x, (int)q.w);
    }
    s_part[w][lane][0] = acc.re;
    if (CPLX) s_part[w][lane][CPLX ? 1 : 0] = Acc<CPLX>::imag(acc);
    __syncthreads();
    if (w == 0) {
        const i64 r = s * 32 + lane;
        if (r < n_rows) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int qd = 0; qd < SELL_WARPS; qd++) {
                re += s_part[qd][lane][0];
                if (CPLX) im += s_part[qd][lane][CPLX ? 1 : 0];
            }
            const double dg = __ldg(diag + r);
            if (CPLX) {
                double2 xr = __ldg(reinterpret_cast<const double2*>(x) + row_offset + r);
                reinterpret_cast<double2*>(y)[r] = make_double2(fma(dg, xr.x, re), fma(dg, xr.y, im));
            } else {
                y[r] = fma(dg, __ldg(x + row_offset + r), re);
            }
        }
    }
}

// CSR rows -> packed SELL-32.  The diagonal entry (column == row_offset + r) goes to diag[r];
// slice_ptr (16-byte units) is sized by the caller from (row length - 1).  *inexact is set if an
// off-diagonal value does not survive the float32 round trip.
__global__ void __launch_bounds__(256)
k_sell_pack_f32(i64 n_rows, i64 row_offset, const i64* __restrict__ row_ptr,
                const int32_t* __restrict__ cols, const double* __restrict__ vals,
                const i64* __restrict__ slice_ptr, uint4* __restrict__ packed,
                double* __restrict__ diag, int* inexact)
{
    const i64 s = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const i64 base = slice_ptr[s];
    const i64 npair = (slice_ptr[s + 1] - base) >> 5;
    const i64 r = s * 32 + lane;
    i64 rs = 0, len = 0, dpos = -1;
    if (r < n_rows) {
        rs = row_ptr[r]; len = row_ptr[r + 1] - rs;
        // position of the diagonal inside the row (first entry for unsorted rows; search otherwise)
        for (i64 k = 0; k < len; k++)
            if (cols[rs + k] == (int32_t)(row_offset + r)) { dpos = k; break; }
        if (w == 0) diag[r] = dpos >= 0 ? vals[rs + dpos] : 0.0;
    }
    const i64 noff = dpos >= 0 ? len - 1 : len;           // off-diagonal entries of this row
    bool bad = false;
    for (i64 k2 = w; k2 < npair; k2 += nw) {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        for (int h = 0; h < 2; h++) {
            i64 k = 2 * k2 + h;                            // k-th off-diagonal entry
            if (k < noff) {
                i64 src = rs + (dpos >= 0 && k >= dpos ? k + 1 : k);
                double v = vals[src];
                float f = (float)v;
                if ((double)f != v) bad = true;
                if (h == 0) { q.x = __float_as_uint(f); q.z = (unsigned)cols[src]; }
                else { q.y = __float_as_uint(f); q.w = (unsigned)cols[src]; }
            }
        }
        packed[base + k2 * 32 + lane] = q;
    }
    if (bad) atomicExch(inexact, 1);
}

extern "C" int fgk_sell_pack_f32(int64_t n_rows, int64_t row_offset, const int64_t* row_ptr,
                                 const int32_t* cols, const double* vals, const int64_t* slice_ptr,
                                 void* packed, double* diag, int* inexact_flag, int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!row_ptr || !cols || !vals || !slice_ptr || !packed || !diag || !inexact_flag || n_rows < 0)
        return fgk_fail(FGK_ERR_ARG, "fgk_sell_pack_f32: bad argument");
    if ((uintptr_t)packed & 15) return fgk_fail(FGK_ERR_ARG, "fgk_sell_pack_f32: packed must be 16-byte aligned");
    FGK_CUDA(cudaSetDevice(device));
    i64 n_slices = (n_rows + 31) / 32;
    k_sell_pack_f32<<<(unsigned)n_slices, 256, 0, (cudaStream_t)stream>>>(
        n_rows, row_offset, (const i64*)row_ptr, cols, vals, (const i64*)slice_ptr, (uint4*)packed, diag,
        inexact_flag);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

template <bool CPLX>
static int launch_spmv_sell_f32(int64_t n_rows, int64_t row_offset, const int64_t* slice_ptr,
                                const void* packed, const double* diag, const double* x, double* y,
                                int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!slice_ptr || !packed || !diag || !x || !y || n_rows < 0)
        return fgk_fail(FGK_ERR_ARG, "fgk_spmv_sell_f32: bad argument");
    if ((uintptr_t)packed & 15) return fgk_fail(FGK_ERR_ARG, "fgk_spmv_sell_f32: packed must be 16-byte aligned");
    FGK_CUDA(cudaSetDevice(device));
    i64 n_slices = (n_rows + 31) / 32;
    k_spmv_sell_f32<CPLX, 4><<<(unsigned)n_slices, SELL_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n_rows, row_offset, (const i64*)slice_ptr, (const uint4*)packed, diag, x, y);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_spmv_sell_f32_f64(int64_t n_rows, int64_t row_offset, const int64_t* slice_ptr,
                                     const void* packed, const double* diag, const double* x, double* y,
                                     int device, void* stream)
{
    return launch_spmv_sell_f32<false>(n_rows, row_offset, slice_ptr, packed, diag, x, y, device, stream);
}

extern "C" int fgk_spmv_sell_f32_z(int64_t n_rows, int64_t row_offset, const int64_t* slice_ptr,
                                   const void* packed, const double* diag, const double* x, double* y,
                                   int device, void* stream)
{
    return launch_spmv_sell_f32<true>(n_rows, row_offset, slice_ptr, packed, diag, x, y, device, stream);
}

// ======================================================================================
// One Taylor term of exp(t (H - mu)) psi (solvers.expm_multiply, reference skqd.py:291-293):
//   B <- c (y - mu B)   with y = H B already computed,   F <- F + B,
// and the two infinity norms the truncation test needs, max |B_i| and max |F_i|, folded into
// norms[0..1] with atomicMax on the bit patterns of the (non-negative) moduli.  One pass over
// three complex vectors instead of six elementwise kernels and two reductions.
// ======================================================================================
__global__ void __launch_bounds__(256)
k_taylor_update_z(i64 n, const double2* __restrict__ y, double2* __restrict__ B, double2* __restrict__ F,
                  double mu, double c_re, double c_im, unsigned long long* norms)
{
    double mb = 0.0, mf = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double2 yi = y[i], bi = B[i];
        const double ur = yi.x - mu * bi.x, ui = yi.y - mu * bi.y;
        const double2 nb = make_double2(c_re * ur - c_im * ui, c_re * ui + c_im * ur);
        double2 f = F[i];
        f.x += nb.x;
        f.y += nb.y;
        B[i] = nb;
        F[i] = f;
        mb = fmax(mb, hypot(nb.x, nb.y));
        mf = fmax(mf, hypot(f.x, f.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(norms, (unsigned long long)__double_as_longlong(mb));
        atomicMax(norms + 1, (unsigned long long)__double_as_longlong(mf));
    }
}

extern "C" int fgk_taylor_update_z(int64_t n, const double* y, double* B, double* F, double mu, double c_re,
                                   double c_im, double* norms, int device, void* stream)
{
    if (n == 0) return FGK_OK;
    if (!y || !B || !F || !norms || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_taylor_update_z: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    FGK_CUDA(cudaMemsetAsync(norms, 0, 2 * sizeof(double), st));
    i64 need = (n + 255) / 256, cap = (i64)fgk_sm_count(device) * 8;
    k_taylor_update_z<<<(int)(need < cap ? need : cap), 256, 0, st>>>(
        n, (const double2*)y, (double2*)B, (double2*)F, mu, c_re, c_im, (unsigned long long*)norms);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// ======================================================================================
// Fused vector algebra of one block-Davidson iteration (solvers._davidson_fused): the basis
// vectors V_j and their images W_j = H V_j are the ROWS of two (m_max x n_local) arrays; all
// dot products against the basis are produced by the same pass that forms the vector.
//   MODE 0  ritz + correction : r = W^T s - theta V^T s,  t = r / (theta - diag)  (guarded),
//                               acc_j = V_j . t ,  extra = r . r
//   MODE 1  orthogonalise     : t -= V^T c ,            acc_j = V_j . t ,  extra = t . t
//   MODE 2  finish            : out = (t - V^T c) / sqrt(tt - c . c)   (the new basis vector;
//                               Pythagoras: V is orthonormal), norm written to *nrm_out
//   MODE 3  project           : acc_j = V_j . w   (j < m: new column of the projected matrix)
// Reductions: thread-private accumulators over a grid-stride range -> warp shuffle -> shared
// memory -> one partial row per CTA (partial[cta][0..m] , extra at [m]); the caller adds the
// rows in order (deterministic), then across ranks.  Replaces ~45 small torch kernels per
// iteration (0.8 ms of launch latency per iteration on 2 GPUs, profiles/r02h).
// ======================================================================================
struct DavCoef { double c[64]; };

// One lane = one row of a 32-row tile.  The m dot products of a tile are reduced by a transpose
// through shared memory: lane l writes its m products, lane j adds the 32 products of basis
// vector j (and j + 32).  Modes 1-3 keep the tile's V entries in registers (one read, all loads
// in flight: 34-57 us per pass at 2.7-4.5 TB/s, ncu r02k); the Ritz mode reads V and W and would
// need 187 registers that way (8 warps per SM, 1.06 TB/s), so it streams both and re-reads V
// from cache (92 us).
template <int MCAP, int MODE>
__global__ void __launch_bounds__(128)
k_dav(i64 nl, i64 ld, int m, const double* __restrict__ V, const double* __restrict__ W,
      const __grid_constant__ DavCoef host_coef, const double* __restrict__ dev_coef, const double* __restrict__ tt_dev,
      double theta, const double* __restrict__ diag, double* __restrict__ t, const double* __restrict__ w,
      double* __restrict__ out, double* __restrict__ partial, double* nrm_out)
{
    extern __shared__ double s_tile[];                 // [4 warps][MCAP][33]
    __shared__ double s_c[MCAP];
    __shared__ double s_acc[4][MCAP + 1];
    __shared__ double s_scale;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* const tile_s = s_tile + (size_t)wid * MCAP * 33;
    if (threadIdx.x < MCAP)
        s_c[threadIdx.x] = threadIdx.x < m ? (MODE == 0 ? host_coef.c[threadIdx.x] : (MODE == 3 ? 0.0 : dev_coef[threadIdx.x])) : 0.0;
    __syncthreads();
    if (MODE == 2) {
        if (threadIdx.x == 0) {
            double cc = 0.0;
            for (int j = 0; j < m; j++) cc += s_c[j] * s_c[j];
            const double n2 = *tt_dev - cc;
            const double nt = n2 > 0.0 ? sqrt(n2) : 0.0;
            s_scale = nt > 1e-10 ? 1.0 / nt : 0.0;          // a vanished correction becomes a zero row
            if (blockIdx.x == 0) *nrm_out = nt;
        }
        __syncthreads();
    }
    double extra = 0.0;
    double acc0 = 0.0, acc1 = 0.0;                     // dot products of basis vectors `lane` and `lane + 32`
    const i64 n_tiles = (nl + 31) >> 5;
    for (i64 tile = (i64)blockIdx.x * 4 + wid; tile < n_tiles; tile += (i64)gridDim.x * 4) {
        const i64 i = tile * 32 + lane;
        const bool live = i < nl;
        const i64 ii = live ? i : nl - 1;              // dead lanes read a valid row and contribute nothing
        double ti = 0.0;
        if (MODE == 0) {
            // Ritz mode: V and W are streamed (16 + 16 loads in flight), V is read again for the
            // products below (L1 / L2: the warp has just loaded these lines)
            double xs = 0.0, ws = 0.0;
#pragma unroll 16
            for (int j = 0; j < m; j++) {
                xs = fma(s_c[j], __ldg(V + j * ld + ii), xs);
                ws = fma(s_c[j], __ldg(W + j * ld + ii), ws);
            }
            if (live) {
                const double r = ws - theta * xs;
                double den = theta - diag[i];
                if (fabs(den) < 1e-8) den = -1e-8;
                ti = r / den;
                extra = fma(r, r, extra);
                t[i] = ti;
            }
#pragma unroll 16
            for (int j = 0; j < m; j++) tile_s[j * 33 + lane] = __ldg(V + j * ld + ii) * ti;
        } else {
            // the other modes keep the tile's V entries in registers: one read, all loads in flight
            double v[MCAP];
#pragma unroll
            for (int j = 0; j < MCAP; j++) v[j] = j < m ? __ldg(V + j * ld + ii) : 0.0;
            if (MODE == 1 || MODE == 2) {
                double pr = 0.0;
#pragma unroll
                for (int j = 0; j < MCAP; j++)
                    if (j < m) pr = fma(s_c[j], v[j], pr);
                if (live) {
                    ti = t[i] - pr;
                    if (MODE == 1) { extra = fma(ti, ti, extra); t[i] = ti; }
                    else out[i] = ti * s_scale;
                }
            } else if (live) {
                ti = w[i];
            }
            if (MODE != 2) {
#pragma unroll
                for (int j = 0; j < MCAP; j++)
                    if (j < m) tile_s[j * 33 + lane] = v[j] * ti;
            }
        }
        if (MODE != 2) {
            __syncwarp();
            if (lane < m) {
                double a = 0.0;
#pragma unroll
                for (int q = 0; q < 32; q++) a += tile_s[lane * 33 + q];
                acc0 += a;
            }
            if (MCAP > 32 && lane + 32 < m) {
                double a = 0.0;
#pragma unroll
                for (int q = 0; q < 32; q++) a += tile_s[(lane + 32) * 33 + q];
                acc1 += a;
            }
            __syncwarp();
        }
    }
    if (MODE == 2) return;
    if (lane < MCAP) s_acc[wid][lane] = acc0;
    if (MCAP > 32) s_acc[wid][lane + 32] = acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, o);
    if (lane == 0) s_acc[wid][MCAP] = extra;
    __syncthreads();
    if (threadIdx.x <= m) {
        const int j = threadIdx.x == m ? MCAP : threadIdx.x;
        partial[(i64)blockIdx.x * (m + 1) + threadIdx.x] = s_acc[0][j] + s_acc[1][j] + s_acc[2][j] + s_acc[3][j];
    }
}

template <int MCAP>
static void launch_dav(int mode, int grid, cudaStream_t st, i64 nl, i64 ld, int m, const double* V, const double* W,
                       const DavCoef& hc, const double* dc, const double* tt, double theta, const double* diag,
                       double* t, const double* w, double* out, double* partial, double* nrm)
{
    const size_t smem = sizeof(double) * 4 * MCAP * 33;
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        cudaFuncSetAttribute(k_dav<MCAP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_dav<MCAP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_dav<MCAP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_dav<MCAP, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    switch (mode) {
    case 0: k_dav<MCAP, 0><<<grid, 128, smem, st>>>(nl, ld, m, V, W, hc, dc, tt, theta, diag, t, w, out, partial, nrm); break;
    case 1: k_dav<MCAP, 1><<<grid, 128, smem, st>>>(nl, ld, m, V, W, hc, dc, tt, theta, diag, t, w, out, partial, nrm); break;
    case 2: k_dav<MCAP, 2><<<grid, 128, smem, st>>>(nl, ld, m, V, W, hc, dc, tt, theta, diag, t, w, out, partial, nrm); break;
    default: k_dav<MCAP, 3><<<grid, 128, smem, st>>>(nl, ld, m, V, W, hc, dc, tt, theta, diag, t, w, out, partial, nrm); break;
    }
}

extern "C" int fgk_davidson_step(int mode, int64_t n_local, int64_t ld, int m, const double* V, const double* W,
                                 const double* host_coef, const double* dev_coef, const double* tt_dev, double theta,
                                 const double* diag, double* t, const double* w, double* out, double* partial,
                                 int n_blocks, double* nrm_out, int device, void* stream)
{
    if (mode < 0 || mode > 3 || n_local < 1 || m < 1 || m > 64 || !V || n_blocks < 1 || ld < n_local)
        return fgk_fail(FGK_ERR_ARG, "fgk_davidson_step: bad argument");
    if ((mode == 0 && (!W || !host_coef || !diag || !t || !partial)) || (mode == 1 && (!dev_coef || !t || !partial)) ||
        (mode == 2 && (!dev_coef || !tt_dev || !t || !out || !nrm_out)) || (mode == 3 && (!w || !partial)))
        return fgk_fail(FGK_ERR_ARG, "fgk_davidson_step: missing buffer for this mode");
    FGK_CUDA(cudaSetDevice(device));
    DavCoef hc;
    for (int j = 0; j < 64; j++) hc.c[j] = (mode == 0 && j < m) ? host_coef[j] : 0.0;
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 16) launch_dav<16>(mode, n_blocks, st, n_local, ld, m, V, W, hc, dev_coef, tt_dev, theta, diag, t, w, out, partial, nrm_out);
    else if (m <= 32) launch_dav<32>(mode, n_blocks, st, n_local, ld, m, V, W, hc, dev_coef, tt_dev, theta, diag, t, w, out, partial, nrm_out);
    else launch_dav<64>(mode, n_blocks, st, n_local, ld, m, V, W, hc, dev_coef, tt_dev, theta, diag, t, w, out, partial, nrm_out);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}
