// fgk_index.cu -- K4: device index of a determinant basis.
//
// Three open-addressing tables (linear probing, power-of-two sizes):
//   table : 64-bit entries (tag << 32 | basis index); the 16-byte key is read from
//           the caller's basis array only when the 32-bit tag matches, so a miss
//           usually costs ONE 8-byte load (most probes of the projected-H build
//           and of the PT2 filter are misses);
//   aset / bset : the distinct alpha / beta strings of the basis.  A determinant
//           can only be in the basis if both its strings are; these tables are a
//           few KB for CAS-like bases and stay L1-resident, so they reject most
//           excitations before the full-key probe.
// Replaces the Python dict / set of molecular.py:501,512,
// residual_expansion.py:445-449,513 and skqd.py:171-175,405-407.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "fgk_internal.cuh"

// The index owns ~10 small device buffers and is rebuilt for every new basis.  Plain cudaMalloc /
// cudaFree calls get slow (tens of ms per index on some boxes) once a caching allocator holds
// most of the device, so the buffers come from a stream-ordered memory pool OWNED BY THIS
// LIBRARY (one per device, blocks kept between indices).  The device's default pool -- which
// belongs to the host application -- is not touched.
cudaError_t fgk_pool_alloc(void** p, size_t bytes, cudaStream_t st, int device)
{
    static cudaMemPool_t pools[64] = {nullptr};
    cudaMemPool_t& pool = pools[device & 63];
    if (!pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess) { pool = nullptr; return e; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 16, pool, st);
}

static u64 pow2_at_least(u64 x)
{
    u64 p = 1;
    while (p < x) p <<= 1;
    return p;
}

__global__ void __launch_bounds__(256)
k_index_insert(const fgk_det* __restrict__ dets, i64 n, u64* table, u64 mask)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
        u64 h = det_hash(d.x, d.y);
        u64 tag = h >> 32, slot = h & mask;
        u64 entry = (tag << 32) | (u64)(unsigned)i;
        while (true) {
            u64 prev = atomicCAS((unsigned long long*)&table[slot], FGK_EMPTY, entry);
            if (prev == FGK_EMPTY) break;
            if ((prev >> 32) == tag) {
                unsigned idx = (unsigned)(prev & 0xffffffffu);
                ulonglong2 k = __ldg(reinterpret_cast<const ulonglong2*>(dets) + idx);
                if (k.x == d.x && k.y == d.y) {
                    // duplicate determinant: the reference's dict keeps the LAST index
                    atomicMax((unsigned long long*)&table[slot], entry);
                    break;
                }
            }
            slot = (slot + 1) & mask;
        }
    }
}

// word `which` (0 = alpha, 1 = beta) of every determinant
__global__ void __launch_bounds__(256)
k_extract_words(const fgk_det* __restrict__ dets, i64 n, int which, u64* __restrict__ out)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
        out[i] = which ? d.y : d.x;
    }
}

// insert the (distinct) words of a list into a word set
__global__ void __launch_bounds__(256)
k_set_insert_list(const u64* __restrict__ list, i64 n, u64* set, u64 mask)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const u64 w = __ldg(list + i);
        u64 slot = word_hash(w) & mask;
        while (true) {
            u64 prev = atomicCAS((unsigned long long*)&set[slot], FGK_EMPTY, w);
            if (prev == FGK_EMPTY || prev == w) break;
            slot = (slot + 1) & mask;
        }
    }
}

// ra[i] / rb[i] = rank of determinant i's alpha / beta string in the ascending distinct lists
__global__ void __launch_bounds__(256)
k_string_ranks(const fgk_det* __restrict__ dets, i64 n, const u64* __restrict__ alist, i64 na,
               const u64* __restrict__ blist, i64 nb, int32_t* __restrict__ ra, int32_t* __restrict__ rb)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
        i64 lo = 0, hi = na;
        while (lo < hi) { i64 mid = (lo + hi) >> 1; if (__ldg(alist + mid) < d.x) lo = mid + 1; else hi = mid; }
        ra[i] = (int32_t)lo;
        lo = 0; hi = nb;
        while (lo < hi) { i64 mid = (lo + hi) >> 1; if (__ldg(blist + mid) < d.y) lo = mid + 1; else hi = mid; }
        rb[i] = (int32_t)lo;
    }
}

// duplicates resolve to the LAST index, like the hash table
__global__ void __launch_bounds__(256)
k_pair_fill(const int32_t* __restrict__ ra, const int32_t* __restrict__ rb, i64 n, i64 nb, int32_t* pair)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        atomicMax(pair + (i64)ra[i] * nb + rb[i], (int32_t)i);
}

__global__ void __launch_bounds__(256)
k_index_lookup(IndexView I, const fgk_det* __restrict__ q, i64 m, int32_t* __restrict__ out)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(q) + i);
        fgk_det o = {d.x, d.y};
        out[i] = index_find_filtered(I, o, 4);
    }
}

static int grid1d(i64 n, int device)
{
    i64 need = (n + 255) / 256;
    i64 cap = (i64)fgk_sm_count(device) * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static void index_release(fgk_index* I, cudaStream_t st)
{
    void* bufs[8] = {I->table, I->aset, I->bset, I->alist, I->blist, I->ra, I->rb, I->pair};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, st);
    delete I;
}

// on a CUDA error: give every buffer back (stream-ordered) and free the handle
#define FGK_CUDA_I(call)                                                                 \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            for (void* t__ : tmp) if (t__) cudaFreeAsync(t__, st);                       \
            index_release(I, st);                                                        \
            return fgk_fail(FGK_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
                            cudaGetErrorString(e__));                                    \
        }                                                                                \
    } while (0)

// The index is complete in STREAM ORDER on `stream` when the call returns (one host
// synchronisation inside, to size the string tables); use it on that stream, or order other
// streams after it.  The distinct alpha / beta strings are found by a device radix sort +
// unique (ascending lists: the projected-H builder's emission order is deterministic).
extern "C" int fgk_index_create(const uint64_t* dets, int64_t n, int device, void* stream,
                                fgk_index_t* out)
{
    if (!out || n < 0 || (n > 0 && !dets)) return fgk_fail(FGK_ERR_ARG, "fgk_index_create: bad argument");
    if (n >= (1ll << 31)) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_index_create: n >= 2^31");
    FGK_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    fgk_index* I = new fgk_index();
    I->device = device;
    I->stream = st;
    I->table = I->aset = I->bset = I->alist = I->blist = nullptr;
    I->ra = I->rb = I->pair = nullptr;
    I->n_alpha_strings = I->n_beta_strings = 0;
    void* tmp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // words, sorted, unique a, unique b, cub / counts
    const fgk_det* d = (const fgk_det*)dets;
    u64 tsize = pow2_at_least((u64)(n > 0 ? 2 * n : 1) < 64 ? 64 : (u64)2 * n);
    FGK_CUDA_I(fgk_pool_alloc((void**)&I->table, tsize * sizeof(u64), st, device));
    FGK_CUDA_I(cudaMemsetAsync(I->table, 0xFF, tsize * sizeof(u64), st));
    unsigned long long h_cnt[2] = {0, 0};
    if (n > 0) {
        k_index_insert<<<grid1d(n, device), 256, 0, st>>>(d, n, I->table, tsize - 1);
        FGK_CUDA_I(cudaGetLastError());
        // distinct strings per spin: extract -> radix sort -> unique (+ count)
        size_t cub_a = 0, cub_b = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, cub_a, (const u64*)nullptr, (u64*)nullptr, (int)n, 0, 64, st);
        cub::DeviceSelect::Unique(nullptr, cub_b, (const u64*)nullptr, (u64*)nullptr, (unsigned long long*)nullptr,
                                  (int)n, st);
        const size_t cub_bytes = ((cub_a > cub_b ? cub_a : cub_b) + 255) & ~(size_t)255;
        for (int t = 0; t < 4; t++) FGK_CUDA_I(fgk_pool_alloc(&tmp[t], (size_t)n * sizeof(u64), st, device));
        FGK_CUDA_I(fgk_pool_alloc(&tmp[4], cub_bytes + 2 * sizeof(unsigned long long), st, device));
        unsigned long long* d_cnt = (unsigned long long*)((char*)tmp[4] + cub_bytes);
        for (int which = 0; which < 2; which++) {
            size_t tb = cub_bytes;
            k_extract_words<<<grid1d(n, device), 256, 0, st>>>(d, n, which, (u64*)tmp[0]);
            FGK_CUDA_I(cudaGetLastError());
            FGK_CUDA_I(cub::DeviceRadixSort::SortKeys(tmp[4], tb, (const u64*)tmp[0], (u64*)tmp[1], (int)n, 0, 64, st));
            tb = cub_bytes;
            FGK_CUDA_I(cub::DeviceSelect::Unique(tmp[4], tb, (const u64*)tmp[1], (u64*)tmp[2 + which],
                                                 d_cnt + which, (int)n, st));
        }
        FGK_CUDA_I(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        FGK_CUDA_I(cudaStreamSynchronize(st));          // the one host sync: sizes of the string tables
    }
    I->n_alpha_strings = (i64)h_cnt[0];
    I->n_beta_strings = (i64)h_cnt[1];
    u64 asz = pow2_at_least(h_cnt[0] * 4 < 64 ? 64 : h_cnt[0] * 4);
    u64 bsz = pow2_at_least(h_cnt[1] * 4 < 64 ? 64 : h_cnt[1] * 4);
    FGK_CUDA_I(fgk_pool_alloc((void**)&I->aset, asz * sizeof(u64), st, device));
    FGK_CUDA_I(fgk_pool_alloc((void**)&I->bset, bsz * sizeof(u64), st, device));
    FGK_CUDA_I(cudaMemsetAsync(I->aset, 0xFF, asz * sizeof(u64), st));
    FGK_CUDA_I(cudaMemsetAsync(I->bset, 0xFF, bsz * sizeof(u64), st));
    // right-sized ascending lists of the distinct strings (scanned by the projected-H builder)
    FGK_CUDA_I(fgk_pool_alloc((void**)&I->alist, (h_cnt[0] ? h_cnt[0] : 1) * sizeof(u64), st, device));
    FGK_CUDA_I(fgk_pool_alloc((void**)&I->blist, (h_cnt[1] ? h_cnt[1] : 1) * sizeof(u64), st, device));
    if (n > 0) {
        FGK_CUDA_I(cudaMemcpyAsync(I->alist, tmp[2], h_cnt[0] * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        FGK_CUDA_I(cudaMemcpyAsync(I->blist, tmp[3], h_cnt[1] * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        k_set_insert_list<<<grid1d((i64)h_cnt[0], device), 256, 0, st>>>(I->alist, (i64)h_cnt[0], I->aset, asz - 1);
        FGK_CUDA_I(cudaGetLastError());
        k_set_insert_list<<<grid1d((i64)h_cnt[1], device), 256, 0, st>>>(I->blist, (i64)h_cnt[1], I->bset, bsz - 1);
        FGK_CUDA_I(cudaGetLastError());
        // rank form: string ranks per determinant and, when the basis covers at least 1/16 of its
        // alpha x beta string product (always for CAS-like / product bases), the dense pair table
        // the rank-based projected-H builder reads instead of probing the hash table
        FGK_CUDA_I(fgk_pool_alloc((void**)&I->ra, (size_t)n * sizeof(int32_t), st, device));
        FGK_CUDA_I(fgk_pool_alloc((void**)&I->rb, (size_t)n * sizeof(int32_t), st, device));
        k_string_ranks<<<grid1d(n, device), 256, 0, st>>>(d, n, I->alist, I->n_alpha_strings, I->blist,
                                                          I->n_beta_strings, I->ra, I->rb);
        FGK_CUDA_I(cudaGetLastError());
        const i64 prod = I->n_alpha_strings * I->n_beta_strings;
        const i64 lim = 16 * n > (1ll << 20) ? 16 * n : (1ll << 20);
        if (prod <= lim && prod < (1ll << 31)) {
            FGK_CUDA_I(fgk_pool_alloc((void**)&I->pair, (size_t)prod * sizeof(int32_t), st, device));
            FGK_CUDA_I(cudaMemsetAsync(I->pair, 0xFF, (size_t)prod * sizeof(int32_t), st));
            k_pair_fill<<<grid1d(n, device), 256, 0, st>>>(I->ra, I->rb, n, I->n_beta_strings, I->pair);
            FGK_CUDA_I(cudaGetLastError());
        }
        for (void*& t : tmp) { if (t) cudaFreeAsync(t, st); t = nullptr; }
    }
    I->v.ra = I->ra; I->v.rb = I->rb; I->v.pair = I->pair; I->v.n_bstr = I->n_beta_strings;
    I->v.dets = d; I->v.n = n;
    I->v.table = I->table; I->v.mask = tsize - 1;
    I->v.aset = I->aset; I->v.amask = asz - 1;
    I->v.bset = I->bset; I->v.bmask = bsz - 1;
    *out = I;
    return FGK_OK;
}
#undef FGK_CUDA_I

// The buffers go back to the library's pool in stream order on the stream the index was created
// on: work queued on that stream before the call still sees them (no device-wide
// synchronisation).  Work on OTHER streams must have been ordered before this call by the caller.
extern "C" int fgk_index_destroy(fgk_index_t idx)
{
    if (!idx) return FGK_OK;
    cudaSetDevice(idx->device);
    index_release(idx, idx->stream);
    return FGK_OK;
}

extern "C" int fgk_index_lookup(fgk_index_t idx, const uint64_t* query, int64_t m, int32_t* out_idx,
                                void* stream)
{
    if (!idx) return fgk_fail(FGK_ERR_ARG, "fgk_index_lookup: null handle");
    if (m == 0) return FGK_OK;
    if (!query || !out_idx || m < 0) return fgk_fail(FGK_ERR_ARG, "fgk_index_lookup: bad argument");
    FGK_CUDA(cudaSetDevice(idx->device));
    k_index_lookup<<<grid1d(m, idx->device), 256, 0, (cudaStream_t)stream>>>(
        idx->v, (const fgk_det*)query, m, out_idx);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_index_info(fgk_index_t idx, int64_t* n_dets, int64_t* n_alpha_strings,
                              int64_t* n_beta_strings)
{
    if (!idx) return fgk_fail(FGK_ERR_ARG, "fgk_index_info: null handle");
    if (n_dets) *n_dets = idx->v.n;
    if (n_alpha_strings) *n_alpha_strings = idx->n_alpha_strings;
    if (n_beta_strings) *n_beta_strings = idx->n_beta_strings;
    return FGK_OK;
}

extern "C" int fgk_index_layout(fgk_index_t idx, int* dense_pairs)
{
    if (!idx) return fgk_fail(FGK_ERR_ARG, "fgk_index_layout: null handle");
    if (dense_pairs) *dense_pairs = idx->pair ? 1 : 0;
    return FGK_OK;
}
