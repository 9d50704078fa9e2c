// fgk_peer.cu -- multi-GPU H.v with the all-gather fused into the product.
//
// One process per GPU (torchrun).  Every rank owns a row block of H (SELL-32) and a
// full-length, peer-mapped copy of the Krylov vector.  k_peer_step computes the rank's slice of
// y = H x and stores each y_r straight into the NEXT vector buffer of EVERY rank (its own and
// the peers', over NVLink through cudaIpc-mapped pointers), so the "all-gather" is 8-byte-per-row
// remote stores overlapped with the HBM stream instead of a separate collective, and its last
// CTA runs the flag barrier that closes the step.  Buffers ping-pong, so one barrier per step is
// enough: a rank can only start overwriting buffer A (as output of step k+1) after every rank
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


// ======================================================================================
// One-launch step: product + broadcast + barrier.
// k_peer_step is the same product (FP64 SELL-32 or packed exact-float32 SELL-32 storage, real or
// complex vector); every CTA stores its 32 rows into the NEXT vector of every rank, fences
// (system scope) and bumps a device counter; the CTA that finishes LAST performs the flag
// barrier that k_peer_barrier used to do in a second launch: arrive on every rank's flag array,
// then wait until every peer has arrived.  When the kernel completes, the next vector is
// complete on this rank and every peer has finished reading the current one.
// One CTA per GPU waits, and the ranks' kernels run on different GPUs: no co-residency
// assumption between launches.
// ======================================================================================
__device__ __forceinline__ uint4 ldp_u32x4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct PeerSync {
    unsigned long long* f[PEER_MAX];    // flag arrays (see k_peer_barrier)
    unsigned* done;                     // device counter of finished CTAs (self-resetting)
    unsigned long long* err;
    unsigned long long epoch;
    int rank, world;
};

// called by every thread of every CTA after its remote stores; one warp of the last CTA runs the barrier
__device__ __forceinline__ void peer_finish(const PeerSync& S)
{
    __shared__ bool s_last;
    __syncthreads();                              // the CTA's remote stores are ordered before thread 0's fence
    if (threadIdx.x == 0) {
        // ONE system-scope fence per CTA (cumulative over the barrier above): every thread fencing its
        // own stores cost 3 % of the step at N = 2 (128 MEMBAR.SC.SYS per CTA, each waiting for NVLink acks)
        __threadfence_system();
        const unsigned prev = atomicAdd(S.done, 1u);
        s_last = prev == gridDim.x * gridDim.y - 1;
        if (s_last) *S.done = 0;                  // ready for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();                       // every CTA's stores happened before its counter bump
    const int p = threadIdx.x;
    if (p < S.world) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(S.f[p] + S.rank), "l"(S.epoch) : "memory");
        unsigned long long seen = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(S.f[S.rank] + p) : "memory");
            if (++spins > 400000000ll) { *S.err = S.epoch; break; }
        } while (seen < S.epoch);
    }
}

template <bool CPLX, bool PACKED, int UNROLL>
__global__ void __launch_bounds__(128, CPLX ? 10 : 14)
k_peer_step(i64 n_rows, const i64* __restrict__ slice_ptr, const void* __restrict__ cols_or_packed,
            const double* __restrict__ vals, const double* __restrict__ diag, const double* __restrict__ x,
            const __grid_constant__ PeerPtrs P, const __grid_constant__ PeerSync S, i64 row_offset)
{
    __shared__ double s_part[4][32][CPLX ? 2 : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 s = blockIdx.x;
    const i64 base = __ldg(slice_ptr + s);
    double re = 0.0, im = 0.0;
    auto add = [&](double v, int c) {
        if (CPLX) {
            const double2 z = __ldg(reinterpret_cast<const double2*>(x) + c);
            re = fma(v, z.x, re);
            im = fma(v, z.y, im);
        } else {
            re = fma(v, __ldg(x + c), re);
        }
    };
    if (PACKED) {
        const i64 npair = (__ldg(slice_ptr + s + 1) - base) >> 5;          // slice_ptr in 16-byte units
        const uint4* p0 = reinterpret_cast<const uint4*>(cols_or_packed) + base + lane;
        i64 k2 = w;
        for (; k2 + 4ll * (UNROLL - 1) < npair; k2 += 4ll * UNROLL) {
            uint4 q[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) q[u] = ldp_u32x4(p0 + (k2 + 4ll * u) * 32);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                add((double)__uint_as_float(q[u].x), (int)q[u].z);
                add((double)__uint_as_float(q[u].y), (int)q[u].w);
            }
        }
        for (; k2 < npair; k2 += 4) {
            const uint4 q = ldp_u32x4(p0 + k2 * 32);
            add((double)__uint_as_float(q.x), (int)q.z);
            add((double)__uint_as_float(q.y), (int)q.w);
        }
    } else {
        const i64 npair = (__ldg(slice_ptr + s + 1) - base) >> 6;
        const double* v0 = vals + base + 2 * lane;
        const int32_t* c0 = reinterpret_cast<const int32_t*>(cols_or_packed) + base + 2 * lane;
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
            for (int u = 0; u < UNROLL; u++) { add(v[u].x, c[u].x); add(v[u].y, c[u].y); }
        }
        for (; k2 < npair; k2 += 4) {
            const double2 v = ldp_f64x2(v0 + k2 * 64);
            const int2 c = ldp_s32x2(c0 + k2 * 64);
            add(v.x, c.x);
            add(v.y, c.y);
        }
    }
    s_part[w][lane][0] = re;
    if (CPLX) s_part[w][lane][CPLX ? 1 : 0] = im;
    __syncthreads();
    const i64 r = s * 32 + lane;
    if (r < n_rows) {
        double yr = s_part[0][lane][0] + s_part[1][lane][0] + s_part[2][lane][0] + s_part[3][lane][0];
        double yi = 0.0;
        if (CPLX) yi = s_part[0][lane][CPLX ? 1 : 0] + s_part[1][lane][CPLX ? 1 : 0] +
                       s_part[2][lane][CPLX ? 1 : 0] + s_part[3][lane][CPLX ? 1 : 0];
        if (PACKED) {                             // the diagonal lives in its own FP64 array
            const double dg = __ldg(diag + r);
            if (CPLX) {
                const double2 xr = __ldg(reinterpret_cast<const double2*>(x) + row_offset + r);
                yr = fma(dg, xr.x, yr);
                yi = fma(dg, xr.y, yi);
            } else {
                yr = fma(dg, __ldg(x + row_offset + r), yr);
            }
        }
        // warp w serves peers w, w+4, ...: one coalesced 256-byte (512-byte) store per peer
        for (int p = w; p < S.world; p += 4) {
            if (CPLX) reinterpret_cast<double2*>(P.out[p])[row_offset + r] = make_double2(yr, yi);
            else P.out[p][row_offset + r] = yr;
        }
    }
    peer_finish(S);
}

// all-gather of a vector whose row block lives on this rank: src[0 .. n_local) (16-byte words)
// -> dst_p[dst_offset + i] on every rank p, then the same last-CTA barrier.  Used to distribute
// the INPUT of a product (row-sharded Krylov vectors; host vectors uploaded in N slices).
__global__ void __launch_bounds__(256)
k_peer_gather(const double* __restrict__ src, i64 n_words, const __grid_constant__ PeerPtrs P,
              const __grid_constant__ PeerSync S, i64 dst_offset_words)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (int p = 0; p < S.world; p++) {
        double* dst = P.out[p] + dst_offset_words;
        for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) dst[i] = __ldg(src + i);
    }
    peer_finish(S);
}

// Small all-reduce (sum) over peer memory for the handful of dot products of a row-sharded Krylov
// iteration: every rank stores its n partial sums into slot [rank] of every rank's scratch
// (one of two areas, alternating per call), one flag barrier, then every rank adds the `world`
// partials in rank order -- identical bits on every rank, one launch, no library round trip
// (an 8-rank NCCL all-reduce of a few doubles costs ~40 us, this ~6 us).
__global__ void __launch_bounds__(256)
k_peer_allreduce(const double* __restrict__ src, int n, int src_rows, double* __restrict__ dst,
                 const __grid_constant__ PeerPtrs P, const __grid_constant__ PeerSync S, i64 area_offset, i64 slot_stride)
{
    // src holds src_rows partial rows of n (one per CTA of the producing kernel): added first, in a
    // fixed order (8 interleaved row groups per column, then the groups in order)
    __shared__ double s_grp[8][128];
    for (int i0 = 0; i0 < n; i0 += 128) {
        const int ncol = n - i0 < 128 ? n - i0 : 128;
        for (int idx = threadIdx.x; idx < ncol * 8; idx += blockDim.x) {
            const int i = idx % ncol, gq = idx / ncol;
            double v = 0.0;
            for (int b = gq; b < src_rows; b += 8) v += src[(i64)b * n + i0 + i];
            s_grp[gq][i] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ncol; i += blockDim.x) {
            double v = 0.0;
#pragma unroll
            for (int gq = 0; gq < 8; gq++) v += s_grp[gq][i];
            for (int p = 0; p < S.world; p++) P.out[p][area_offset + (i64)S.rank * slot_stride + i0 + i] = v;
        }
        __syncthreads();
    }
    peer_finish(S);             // gridDim = 1: this CTA is the last one; arrives and waits for every peer
    __syncthreads();
    const double* mine = P.out[S.rank] + area_offset;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < S.world; r++) acc += __ldcg(mine + (i64)r * slot_stride + i);
        dst[i] = acc;
    }
}

static int fill_sync(PeerSync& S, uint64_t* const* peer_flags_host, int rank, int world, uint64_t epoch,
                     uint32_t* done_counter, uint64_t* err_flag)
{
    if (!peer_flags_host || !done_counter || !err_flag || world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return fgk_fail(FGK_ERR_ARG, "fgk_peer: bad synchronisation argument");
    for (int p = 0; p < PEER_MAX; p++) S.f[p] = p < world ? (unsigned long long*)peer_flags_host[p] : nullptr;
    S.done = done_counter;
    S.err = (unsigned long long*)err_flag;
    S.epoch = epoch;
    S.rank = rank;
    S.world = world;
    return FGK_OK;
}

extern "C" int fgk_peer_step(int64_t n_rows, const int64_t* slice_ptr, const void* cols_or_packed,
                             const double* vals, const double* diag, const double* x,
                             double* const* peer_out_host, int flags, int64_t row_offset,
                             uint64_t* const* peer_flags_host, int rank, int world, uint64_t epoch,
                             uint32_t* done_counter, uint64_t* err_flag, int device, void* stream)
{
    const bool cplx = (flags & FGK_PEER_COMPLEX) != 0, packed = (flags & FGK_PEER_PACKED_F32) != 0;
    if (n_rows <= 0) return fgk_fail(FGK_ERR_ARG, "fgk_peer_step: every rank needs at least one row");
    if (!slice_ptr || !cols_or_packed || !x || !peer_out_host || (packed ? !diag : !vals))
        return fgk_fail(FGK_ERR_ARG, "fgk_peer_step: bad argument");
    PeerSync S;
    int rc = fill_sync(S, peer_flags_host, rank, world, epoch, done_counter, err_flag);
    if (rc != FGK_OK) return rc;
    FGK_CUDA(cudaSetDevice(device));
    PeerPtrs P;
    for (int p = 0; p < PEER_MAX; p++) P.out[p] = p < world ? peer_out_host[p] : nullptr;
    const unsigned grid = (unsigned)((n_rows + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
#define FGK_STEP(C, K)                                                                               \
    k_peer_step<C, K, 4><<<grid, 128, 0, st>>>(n_rows, (const i64*)slice_ptr, cols_or_packed, vals, diag, x, P, S, row_offset)
    if (cplx && packed) FGK_STEP(true, true);
    else if (cplx) FGK_STEP(true, false);
    else if (packed) FGK_STEP(false, true);
    else FGK_STEP(false, false);
#undef FGK_STEP
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_peer_gather(const void* src_local, int64_t n_bytes, void* const* peer_dst_host,
                               int64_t dst_offset_bytes, uint64_t* const* peer_flags_host, int rank, int world,
                               uint64_t epoch, uint32_t* done_counter, uint64_t* err_flag, int device, void* stream)
{
    if (!src_local || !peer_dst_host || n_bytes < 0 || (n_bytes & 7) || (dst_offset_bytes & 7) ||
        ((uintptr_t)src_local & 7))
        return fgk_fail(FGK_ERR_ARG, "fgk_peer_gather: pointers, size and offset must be 8-byte multiples");
    PeerSync S;
    int rc = fill_sync(S, peer_flags_host, rank, world, epoch, done_counter, err_flag);
    if (rc != FGK_OK) return rc;
    FGK_CUDA(cudaSetDevice(device));
    PeerPtrs P;
    for (int p = 0; p < PEER_MAX; p++) P.out[p] = p < world ? (double*)peer_dst_host[p] : nullptr;
    const i64 words = n_bytes / 8;
    i64 need = (words + 255) / 256, cap = (i64)fgk_sm_count(device) * 4;
    if (need < 1) need = 1;
    k_peer_gather<<<(unsigned)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(
        (const double*)src_local, words, P, S, dst_offset_bytes / 8);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// Host-buffer product on N GPUs in ONE library call: this rank's slice of x comes from (pinned)
// host memory, the slices are exchanged over peer memory, one fused step runs, and the rank's rows
// of y go back to (pinned) host memory; returns when y_host is complete.  Five host-side calls
// (copy, gather, step, copy, synchronise) folded into one entry point: the e2e step at N = 8 is
// ~0.65 ms, of which the interpreter overhead of five separate calls was ~0.05 ms.
extern "C" int fgk_peer_matvec_host(int64_t n_rows, const int64_t* slice_ptr, const void* cols_or_packed,
                                    const double* vals, const double* diag, const void* x_host_slice,
                                    void* x_dev_slice, double* const* peer_cur_host, double* const* peer_next_host,
                                    void* y_host, int flags, int64_t row_offset, uint64_t* const* peer_flags_host,
                                    int rank, int world, uint64_t epoch, uint32_t* done_counter, uint64_t* err_flag,
                                    int device, void* stream)
{
    if (!x_host_slice || !x_dev_slice || !peer_cur_host || !peer_next_host || !y_host || rank < 0 || rank >= world)
        return fgk_fail(FGK_ERR_ARG, "fgk_peer_matvec_host: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t w = (flags & FGK_PEER_COMPLEX) ? 16 : 8;
    FGK_CUDA(cudaMemcpyAsync(x_dev_slice, x_host_slice, (size_t)(w * n_rows), cudaMemcpyHostToDevice, st));
    int rc = fgk_peer_gather(x_dev_slice, w * n_rows, (void* const*)peer_cur_host, w * row_offset, peer_flags_host, rank,
                             world, epoch, done_counter, err_flag, device, stream);
    if (rc != FGK_OK) return rc;
    rc = fgk_peer_step(n_rows, slice_ptr, cols_or_packed, vals, diag, peer_cur_host[rank], peer_next_host, flags,
                       row_offset, peer_flags_host, rank, world, epoch + 1, done_counter, err_flag, device, stream);
    if (rc != FGK_OK) return rc;
    const char* y_rows = reinterpret_cast<const char*>(peer_next_host[rank]) + w * row_offset;
    FGK_CUDA(cudaMemcpyAsync(y_host, y_rows, (size_t)(w * n_rows), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaStreamSynchronize(st));
    return FGK_OK;
}

extern "C" int fgk_peer_allreduce_sum(const double* src, int64_t n, int64_t src_rows, double* dst, double* const* peer_scratch_host,
                                      int64_t slot_stride, int area, uint64_t* const* peer_flags_host, int rank,
                                      int world, uint64_t epoch, uint32_t* done_counter, uint64_t* err_flag,
                                      int device, void* stream)
{
    if (!src || !dst || !peer_scratch_host || n < 0 || n > slot_stride || slot_stride < 1 || (area != 0 && area != 1) ||
        src_rows < 1 || src_rows > (1 << 20))
        return fgk_fail(FGK_ERR_ARG, "fgk_peer_allreduce_sum: bad argument");
    PeerSync S;
    int rc = fill_sync(S, peer_flags_host, rank, world, epoch, done_counter, err_flag);
    if (rc != FGK_OK) return rc;
    FGK_CUDA(cudaSetDevice(device));
    PeerPtrs P;
    for (int p = 0; p < PEER_MAX; p++) P.out[p] = p < world ? peer_scratch_host[p] : nullptr;
    k_peer_allreduce<<<1, 256, 0, (cudaStream_t)stream>>>(src, (int)n, (int)src_rows, dst, P, S,
                                                          (i64)area * world * slot_stride, slot_stride);
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
