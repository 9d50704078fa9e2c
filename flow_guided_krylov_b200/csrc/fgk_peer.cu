// fgk_peer.cu -- multi-GPU H.v with the all-gather fused into the product.
//
// One process per GPU (torchrun).  Every rank owns a row block of H (SELL-32) and a
// full-length, peer-mapped copy of the Krylov vector.  k_spmv_sell_bcast computes the
// rank's slice of y = H x and stores each y_r straight into the NEXT vector buffer of
// EVERY rank (its own and the peers', over NVLink through cudaIpc-mapped pointers), so
// the "all-gather" is 8-byte-per-row remote stores overlapped with the 12 B/nnz HBM
// stream instead of a separate collective.  A flag barrier over the same peer mapping
// (k_peer_barrier) closes the step.  Buffers ping-pong, so one barrier per step is enough:
// a rank can only start overwriting buffer A (as output of step k+1) after every rank
// has finished reading it (as input of step k).
#include <string.h>

#include "fgk_internal.cuh"

extern "C" int fgk_peer_alloc(size_t bytes, int device, void** dev_ptr, unsigned char* handle64)
{
    if (!dev_ptr || !handle64 || bytes == 0) return fgk_fail(FGK_ERR_ARG, "fgk_peer_alloc: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    FGK_CUDA(cudaMalloc(&p, bytes));
    FGK_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fgk_fail(FGK_ERR_CUDA, "cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return FGK_OK;
}

extern "C" int fgk_peer_open(const unsigned char* handle64, int device, void** dev_ptr)
{
    if (!dev_ptr || !handle64) return fgk_fail(FGK_ERR_ARG, "fgk_peer_open: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    FGK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return FGK_OK;
}

extern "C" int fgk_peer_close(void* dev_ptr, int device)
{
    if (!dev_ptr) return FGK_OK;
    FGK_CUDA(cudaSetDevice(device));
    FGK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return FGK_OK;
}

extern "C" int fgk_peer_free(void* dev_ptr, int device)
{
    if (!dev_ptr) return FGK_OK;
    FGK_CUDA(cudaSetDevice(device));
    FGK_CUDA(cudaFree(dev_ptr));
    return FGK_OK;
}

__device__ __forceinline__ double2 ldp_f64x2(const double* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ int2 ldp_s32x2(const int32_t* p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

static const int PEER_MAX = 16;
struct PeerPtrs { double* out[PEER_MAX]; };

// same product as k_spmv_sell<false,4> (fgk_spmv.cu); only the epilogue differs:
// y_r goes to out[p][row_offset + r] on every rank p.  x must not alias any out[p]
// (it is plain, non-coherent load traffic; the outputs are the OTHER ping-pong buffer).
template <int UNROLL>
__global__ void __launch_bounds__(128, 16)
k_spmv_sell_bcast(i64 n_rows, const i64* __restrict__ slice_ptr, const int32_t* __restrict__ cols,
                  const double* __restrict__ vals, const double* __restrict__ x,
                  const __grid_constant__ PeerPtrs P, int world, i64 row_offset)
{
    __shared__ double s_part[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 s = blockIdx.x;
    const i64 base = __ldg(slice_ptr + s);
    const i64 npair = (__ldg(slice_ptr + s + 1) - base) >> 6;
    const double* v0 = vals + base + 2 * lane;
    const int32_t* c0 = cols + base + 2 * lane;
    double acc = 0.0;
    i64 k2 = w;
    for (; k2 + 4ll * (UNROLL - 1) < npair; k2 += 4ll * UNROLL) {
        double2 v[UNROLL];
        int2 c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            v[u] = ldp_f64x2(v0 + (k2 + 4ll * u) * 64);
            c[u] = ldp_s32x2(c0 + (k2 + 4ll * u) * 64);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            acc = fma(v[u].x, __ldg(x + c[u].x), acc);
            acc = fma(v[u].y, __ldg(x + c[u].y), acc);
        }
    }
    for (; k2 < npair; k2 += 4) {
        double2 v = ldp_f64x2(v0 + k2 * 64);
        int2 c = ldp_s32x2(c0 + k2 * 64);
        acc = fma(v.x, __ldg(x + c.x), acc);
        acc = fma(v.y, __ldg(x + c.y), acc);
    }
    s_part[w][lane] = acc;
    __syncthreads();
    // 4 warps x 32 lanes: warp w serves peers w, w+4, ...: a coalesced 256-byte store per peer
    const i64 r = s * 32 + lane;
    if (r < n_rows) {
        const double y = s_part[0][lane] + s_part[1][lane] + s_part[2][lane] + s_part[3][lane];
        for (int p = w; p < world; p += 4) P.out[p][row_offset + r] = y;
    }
}

// flags[p] points at rank p's flag array (world entries, peer-mapped); entry [rank] of
// rank p's array is written by `rank`.  Arrive: store epoch into every rank's array at
// my index (system scope, after a system fence); wait: until my own array shows
// epoch everywhere.  A bounded spin turns a lost peer into an error instead of a hang.
struct PeerFlags { unsigned long long* f[PEER_MAX]; };

__global__ void k_peer_barrier(const __grid_constant__ PeerFlags F, int rank, int world,
                               unsigned long long epoch, unsigned long long* err)
{
    const int p = threadIdx.x;
    if (p < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(F.f[p] + rank), "l"(epoch) : "memory");
        unsigned long long seen = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(F.f[rank] + p) : "memory");
            if (++spins > 400000000ll) { *err = epoch; break; }
        } while (seen < epoch);
    }
}

extern "C" int fgk_spmv_sell_f64_allgather(int64_t n_rows, const int64_t* slice_ptr,
                                           const int32_t* sell_cols, const double* sell_vals,
                                           const double* x, double* const* peer_out_host, int world,
                                           int64_t row_offset, int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!slice_ptr || !sell_cols || !sell_vals || !x || !peer_out_host || world < 1 || world > PEER_MAX)
        return fgk_fail(FGK_ERR_ARG, "fgk_spmv_sell_f64_allgather: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    PeerPtrs P;
    for (int p = 0; p < PEER_MAX; p++) P.out[p] = p < world ? peer_out_host[p] : nullptr;
    i64 n_slices = (n_rows + 31) / 32;
    k_spmv_sell_bcast<4><<<(unsigned)n_slices, 128, 0, (cudaStream_t)stream>>>(
        n_rows, (const i64*)slice_ptr, sell_cols, sell_vals, x, P, world, row_offset);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_peer_barrier(uint64_t* const* peer_flags_host, int rank, int world, uint64_t epoch,
                                uint64_t* err_flag, int device, void* stream)
{
    if (!peer_flags_host || !err_flag || world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return fgk_fail(FGK_ERR_ARG, "fgk_peer_barrier: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    PeerFlags F;
    for (int p = 0; p < PEER_MAX; p++)
        F.f[p] = p < world ? (unsigned long long*)peer_flags_host[p] : nullptr;
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(F, rank, world, epoch, (unsigned long long*)err_flag);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}
