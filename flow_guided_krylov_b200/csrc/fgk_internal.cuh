// fgk_internal.cuh -- handles, error plumbing and warp-level helpers shared by
// the .cu translation units of libfgk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fgk_b200.h"
#include "fgk_core.cuh"

typedef long long i64;

// ---- error plumbing -----------------------------------------------------------
std::string& fgk_err_slot();
int fgk_fail(int code, const char* fmt, ...);

#define FGK_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return fgk_fail(FGK_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,  \
                            cudaGetErrorString(e__));                                   \
    } while (0)

#define FGK_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return fgk_fail(FGK_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__,        \
                            __LINE__, cudaGetErrorString(e__));                         \
    } while (0)

// ---- handles --------------------------------------------------------------------
struct fgk_ham {
    int device;
    HamView v;          // device pointers
    float *h1, *g, *w;  // views into itab
    float* itab;        // [h1 | g | w] float32 in ONE contiguous buffer (16-byte granular segments)
    unsigned itab_bytes;  // its size when it fits shared memory next to the kernels' own (else 0):
                          // one TMA bulk copy stages all integral tables for the connection enumerator
    double* dtab;       // [h_pp (padded to even)] [nibble (J-K) rows] [nibble J rows]: one TMA bulk copy
    unsigned dtab_bytes;  // 0 when the nibble tables would not fit shared memory (n_orb > 56)
    double* jkab;       // plain jks | jab (n*n each), for the pair-loop fallback
};

struct IndexView {
    const fgk_det* dets;
    i64 n;
    const u64* table;   // (tag << 32 | index), empty = ~0
    u64 mask;
    const u64* aset;    // distinct alpha strings, empty = ~0
    u64 amask;
    const u64* bset;
    u64 bmask;
    // rank form of the basis (null for an empty basis): ra[i] / rb[i] = position of determinant
    // i's alpha / beta string in the sorted distinct-string lists; pair[ra * n_bstr + rb] = basis
    // index of that (alpha, beta) combination or -1 -- present only when the basis fills its
    // string product densely enough (see fgk_index_create), else null.
    const int32_t* ra;
    const int32_t* rb;
    const int32_t* pair;
    i64 n_bstr;
};

struct fgk_index {
    int device;
    cudaStream_t stream;            // creation stream: the buffers are freed in its order
    IndexView v;
    u64 *table, *aset, *bset;
    u64 *alist, *blist;             // the distinct alpha / beta strings, ascending
    i64 n_alpha_strings, n_beta_strings;
    int32_t *ra, *rb, *pair;        // string ranks per determinant; dense pair table (may be null)
};

struct Pt2View {
    u64* table;         // (tag << 32 | slot), empty = ~0
    u64 mask;           // table_slots - 1
    u64* pool;          // 4 words per slot: {alpha, beta, accumulator lo, hi} (128-bit fixed point; MAXABS: FP64 bits, 0)
    i64 capacity;
    unsigned long long* counters;   // [0] slots used, [1] raw candidates tested, [2] overflow
    // The table is split into 2^region_bits regions selected by the TOP hash bits; linear
    // probing stays inside a region.  With region_bits = 0 it is one flat table.
    int region_bits;
    u64 region_mask;    // (table_slots >> region_bits) - 1
    // Optional partition queues (radix partition before the hash): 2^queue_bits queues selected
    // by the top hash bits (queue_bits >= region_bits, so queue order is region order), each
    // `qstride` (determinant, value) pairs long; qcursors[q] = pairs appended to queue q.
    int queue_bits;
    i64 qstride;
    fgk_det* qdets;
    double* qvals;
    unsigned long long* qcursors;
};

struct fgk_pt2 {
    int device;
    Pt2View v;
    int mode;           // FGK_PT2_SUM / FGK_PT2_MAXABS of the sweep in progress, -1 after reset
    bool scored;        // accumulators already converted to {coupling, score} by fgk_pt2_score
};

static const u64 FGK_EMPTY = ~0ull;
static const int FGK_WARPS_PER_BLOCK = 8;
static const int FGK_BLOCK = FGK_WARPS_PER_BLOCK * 32;

int fgk_sm_count(int device);
// stream-ordered allocation from the library's own memory pool (fgk_index.cu); free with cudaFreeAsync
cudaError_t fgk_pool_alloc(void** p, size_t bytes, cudaStream_t st, int device);

// ---- device helpers ---------------------------------------------------------------
#if defined(__CUDACC__)

// ---- TMA (bulk async copy) staging of a contiguous table into shared memory -------------
// One thread arms an mbarrier with the byte count and issues cp.async.bulk
// (global -> shared, completes on the mbarrier); every thread then waits on parity 0.
// bytes: multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_stage_table(void* smem_dst, const void* gmem_src, unsigned bytes,
                                                unsigned long long* mbar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
    }
    // all threads: wait for phase 0 to complete
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FGK_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra FGK_TMA_DONE;\n"
        "bra FGK_TMA_WAIT;\n"
        "FGK_TMA_DONE:\n"
        "}\n" :: "r"(smem_u32(mbar)) : "memory");
}

struct LdgF { __device__ __forceinline__ float operator()(const float* p) const { return __ldg(p); } };
struct LdgD { __device__ __forceinline__ double operator()(const double* p) const { return __ldg(p); } };

// explicit shared-space load for tables staged in shared memory (a table pointer rebased onto
// the staging buffer is a generic address to the compiler: it emits LD instead of LDS)
struct LdsF {
    __device__ __forceinline__ float operator()(const float* p) const
    {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
        return v;
    }
};

struct LdsD {
    __device__ __forceinline__ double operator()(const double* p) const
    {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
        return v;
    }
};

struct WarpLists { uint8_t occ_a[64], virt_a[64], occ_b[64], virt_b[64]; };

// one lane per orbital builds the ascending occupied / virtual lists of d
__device__ __forceinline__ void warp_build_ctx(DetCtx& c, int n, fgk_det d, WarpLists& L, int lane)
{
    __syncwarp();
    for (int p = lane; p < n; p += 32) {
        u64 bit = orb_bit(n, p), below = below_mask(n, p);
        int ka = __popcll(d.a & below), kb = __popcll(d.b & below);
        if (d.a & bit) L.occ_a[ka] = (uint8_t)p; else L.virt_a[p - ka] = (uint8_t)p;
        if (d.b & bit) L.occ_b[kb] = (uint8_t)p; else L.virt_b[p - kb] = (uint8_t)p;
    }
    __syncwarp();
    c.n = n; c.d = d;
    c.noa = __popcll(d.a); c.nva = n - c.noa;
    c.nob = __popcll(d.b); c.nvb = n - c.nob;
    c.occ_a = L.occ_a; c.virt_a = L.virt_a; c.occ_b = L.occ_b; c.virt_b = L.virt_b;
    detctx_sizes(c);
}

// Walk every excitation of c.d in the reference's order, 32 index slots per step.
// fs(valid_a, valid_b, p, q)  -- singles step (alpha then beta for the pair)
// fd(valid, x)                -- doubles step
// Both are called by ALL lanes, convergently, so they may use warp collectives.
template <class FS, class FD>
__device__ __forceinline__ void warp_enumerate(const DetCtx& c, int lane, FS&& fs, FD&& fd)
{
    for (int t0 = 0; t0 < c.n_s; t0 += 32) {
        int t = t0 + lane, p = 0, q = 0;
        bool va = false, vb = false;
        if (t < c.n_s) decode_single(c, t, p, q, va, vb);
        fs(va, vb, p, q);
    }
#pragma unroll 1
    for (int st = 2; st <= 4; st++) {
        const int size = st == 2 ? c.n_aa : (st == 3 ? c.n_bb : c.n_ab);
        for (int t0 = 0; t0 < size; t0 += 32) {
            int t = t0 + lane;
            Excitation x;
            x.cls = st; x.h0 = x.h1 = x.e0 = x.e1 = 0;
            bool valid = t < size;
            if (valid) decode_double(c, st, t, x);
            fd(valid, x);
        }
    }
}

__device__ __forceinline__ bool set_has(const u64* set, u64 mask, u64 w)
{
    u64 slot = word_hash(w) & mask;
    while (true) {
        u64 e = __ldg(set + slot);
        if (e == w) return true;
        if (e == FGK_EMPTY) return false;
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ int index_find_h(const IndexView& I, fgk_det o, u64 h)
{
    u64 tag = h >> 32, slot = h & I.mask;
    while (true) {
        u64 e = __ldg(I.table + slot);
        if (e == FGK_EMPTY) return -1;
        if ((e >> 32) == tag) {
            unsigned idx = (unsigned)(e & 0xffffffffu);
            ulonglong2 k = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + idx);
            if (k.x == o.a && k.y == o.b) return (int)idx;
        }
        slot = (slot + 1) & I.mask;
    }
}

__device__ __forceinline__ int index_find(const IndexView& I, fgk_det o)
{
    return index_find_h(I, o, det_hash(o.a, o.b));
}

// cheap necessary conditions first (alpha / beta string sets are small and
// L1-resident), then the full-key probe.  cls tells which words changed.
__device__ __forceinline__ int index_find_filtered(const IndexView& I, fgk_det o, int cls)
{
    if (cls != 1 && cls != 3) { if (!set_has(I.aset, I.amask, o.a)) return -1; }
    if (cls != 0 && cls != 2) { if (!set_has(I.bset, I.bmask, o.b)) return -1; }
    return index_find(I, o);
}

__device__ __forceinline__ int index_find_filtered_h(const IndexView& I, fgk_det o, u64 h, int cls)
{
    if (cls != 1 && cls != 3) { if (!set_has(I.aset, I.amask, o.a)) return -1; }
    if (cls != 0 && cls != 2) { if (!set_has(I.bset, I.bmask, o.b)) return -1; }
    return index_find_h(I, o, h);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// <D|H|D> computed by a whole warp: lane p (and p + 32) takes orbital p's h_pp and nibble row
// sums, then one butterfly sum.  In the row builders a single lane used to walk all occupied
// orbitals (~200 dependent loads, 13 % of k_projh3's instructions, ncu r01r); every lane gets
// the result.  Called convergently by all 32 lanes.
__device__ __forceinline__ double warp_diag_element(const HamView& H, fgk_det d, int lane)
{
    LdgD ldd;
    if (!H.nib_jk) {                    // n_orb > 56: pair-loop form on one lane
        double e = 0.0;
        if (lane == 0) e = diag_element_loops(H, d, ldd);
        return __shfl_sync(0xffffffffu, e, 0);
    }
    const int n = H.n_orb, nc = H.nchunk;
    double part = 0.0;
    for (int p = lane; p < n; p += 32) {
        const u64 bit = orb_bit(n, p);
        if (d.a & bit)
            part += ldd(H.hdiag + p) + 0.5 * nib_rowsum(H.nib_jk + (size_t)p * nc * 16, nc, d.a, ldd) +
                    nib_rowsum(H.nib_jab + (size_t)p * nc * 16, nc, d.b, ldd);
        if (d.b & bit)
            part += ldd(H.hdiag + p) + 0.5 * nib_rowsum(H.nib_jk + (size_t)p * nc * 16, nc, d.b, ldd);
    }
    return H.e_nuc + warp_sum(part);
}

__device__ __forceinline__ i64 warp_sum_i64(i64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
