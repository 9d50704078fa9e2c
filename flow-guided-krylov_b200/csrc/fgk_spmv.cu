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
