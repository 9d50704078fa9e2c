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
#include <algorithm>
#include <vector>

#include "fgk_internal.cuh"

// The index owns ~10 small device buffers.  Plain cudaMalloc / cudaFree calls get slow (tens of
// ms per index on some boxes) once a caching allocator holds most of the device, and an index is
// rebuilt for every new basis; the stream-ordered allocator with an unbounded release threshold
// keeps the blocks in the device's default pool instead.
static cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st, int device)
{
    static bool configured[64] = {false};
    if (!configured[device & 63]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[device & 63] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 16, st);
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

// insert word `which` (0 = alpha, 1 = beta) of every determinant into a word set
__global__ void __launch_bounds__(256)
k_set_insert(const fgk_det* __restrict__ dets, i64 n, int which, u64* set, u64 mask,
             unsigned long long* distinct)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
        u64 w = which ? d.y : d.x;
        u64 slot = word_hash(w) & mask;
        while (true) {
            u64 prev = atomicCAS((unsigned long long*)&set[slot], FGK_EMPTY, w);
            if (prev == FGK_EMPTY) { atomicAdd(distinct, 1ull); break; }
            if (prev == w) break;
            slot = (slot + 1) & mask;
        }
    }
}

// compact the occupied slots of a word set into a dense list (order irrelevant: the
// projected-H builder sorts its rows afterwards)
__global__ void __launch_bounds__(256)
k_set_compact(const u64* __restrict__ set, u64 size, u64* __restrict__ list, unsigned long long* counter)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (u64)gridDim.x * blockDim.x) {
        u64 w = set[i];
        if (w != FGK_EMPTY) list[atomicAdd(counter, 1ull)] = w;
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

extern "C" int fgk_index_create(const uint64_t* dets, int64_t n, int device, void* stream,
                                fgk_index_t* out)
{
    if (!out || n < 0 || (n > 0 && !dets)) return fgk_fail(FGK_ERR_ARG, "fgk_index_create: bad argument");
    if (n >= (1ll << 31)) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_index_create: n >= 2^31");
    FGK_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    fgk_index* I = new fgk_index();
    I->device = device;
    I->table = I->aset = I->bset = nullptr;
    const fgk_det* d = (const fgk_det*)dets;
    u64 tsize = pow2_at_least((u64)(n > 0 ? 2 * n : 1) < 64 ? 64 : (u64)2 * n);
    FGK_CUDA(pool_alloc((void**)&I->table, tsize * sizeof(u64), st, device));
    FGK_CUDA(cudaMemsetAsync(I->table, 0xFF, tsize * sizeof(u64), st));
    if (n > 0) {
        k_index_insert<<<grid1d(n, device), 256, 0, st>>>(d, n, I->table, tsize - 1);
        FGK_LAUNCH_CHECK();
    }
    // string sets: first pass into a worst-case table to count the distinct
    // strings, second pass into a right-sized (cache friendly) one
    unsigned long long* d_cnt = nullptr;
    u64* tmp = nullptr;
    FGK_CUDA(pool_alloc((void**)&d_cnt, 2 * sizeof(unsigned long long), st, device));
    FGK_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
    FGK_CUDA(pool_alloc((void**)&tmp, tsize * sizeof(u64), st, device));
    unsigned long long h_cnt[2] = {0, 0};
    for (int which = 0; which < 2 && n > 0; which++) {
        FGK_CUDA(cudaMemsetAsync(tmp, 0xFF, tsize * sizeof(u64), st));
        k_set_insert<<<grid1d(n, device), 256, 0, st>>>(d, n, which, tmp, tsize - 1, d_cnt + which);
        FGK_LAUNCH_CHECK();
    }
    FGK_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaStreamSynchronize(st));
    I->n_alpha_strings = (i64)h_cnt[0];
    I->n_beta_strings = (i64)h_cnt[1];
    u64 asz = pow2_at_least(h_cnt[0] * 4 < 64 ? 64 : h_cnt[0] * 4);
    u64 bsz = pow2_at_least(h_cnt[1] * 4 < 64 ? 64 : h_cnt[1] * 4);
    FGK_CUDA(pool_alloc((void**)&I->aset, asz * sizeof(u64), st, device));
    FGK_CUDA(pool_alloc((void**)&I->bset, bsz * sizeof(u64), st, device));
    FGK_CUDA(cudaMemsetAsync(I->aset, 0xFF, asz * sizeof(u64), st));
    FGK_CUDA(cudaMemsetAsync(I->bset, 0xFF, bsz * sizeof(u64), st));
    if (n > 0) {
        k_set_insert<<<grid1d(n, device), 256, 0, st>>>(d, n, 0, I->aset, asz - 1, d_cnt);
        FGK_LAUNCH_CHECK();
        k_set_insert<<<grid1d(n, device), 256, 0, st>>>(d, n, 1, I->bset, bsz - 1, d_cnt + 1);
        FGK_LAUNCH_CHECK();
    }
    // dense lists of the distinct strings (scanned by the projected-H builder)
    I->alist = I->blist = nullptr;
    FGK_CUDA(pool_alloc((void**)&I->alist, (h_cnt[0] ? h_cnt[0] : 1) * sizeof(u64), st, device));
    FGK_CUDA(pool_alloc((void**)&I->blist, (h_cnt[1] ? h_cnt[1] : 1) * sizeof(u64), st, device));
    FGK_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
    if (n > 0) {
        k_set_compact<<<grid1d((i64)asz, device), 256, 0, st>>>(I->aset, asz, I->alist, d_cnt);
        FGK_LAUNCH_CHECK();
        k_set_compact<<<grid1d((i64)bsz, device), 256, 0, st>>>(I->bset, bsz, I->blist, d_cnt + 1);
        FGK_LAUNCH_CHECK();
    }
    FGK_CUDA(cudaStreamSynchronize(st));
    // ascending order makes the builder's emission order deterministic (the atomic append
    // above is not); the lists are small next to the basis, a host sort is enough
    for (int which = 0; which < 2 && n > 0; which++) {
        u64* dl = which ? I->blist : I->alist;
        size_t m = (size_t)h_cnt[which];
        std::vector<u64> hl(m);
        FGK_CUDA(cudaMemcpy(hl.data(), dl, m * sizeof(u64), cudaMemcpyDeviceToHost));
        std::sort(hl.begin(), hl.end());
        FGK_CUDA(cudaMemcpy(dl, hl.data(), m * sizeof(u64), cudaMemcpyHostToDevice));
    }
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(d_cnt, st);
    // rank form: string ranks per determinant and, when the basis covers at least 1/16 of its
    // alpha x beta string product (always for CAS-like / product bases), the dense pair table
    // the rank-based projected-H builder reads instead of probing the hash table
    I->ra = I->rb = I->pair = nullptr;
    if (n > 0) {
        FGK_CUDA(pool_alloc((void**)&I->ra, (size_t)n * sizeof(int32_t), st, device));
        FGK_CUDA(pool_alloc((void**)&I->rb, (size_t)n * sizeof(int32_t), st, device));
        k_string_ranks<<<grid1d(n, device), 256, 0, st>>>(d, n, I->alist, I->n_alpha_strings, I->blist,
                                                          I->n_beta_strings, I->ra, I->rb);
        FGK_LAUNCH_CHECK();
        const i64 prod = I->n_alpha_strings * I->n_beta_strings;
        const i64 lim = 16 * n > (1ll << 20) ? 16 * n : (1ll << 20);
        if (prod <= lim && prod < (1ll << 31)) {
            FGK_CUDA(pool_alloc((void**)&I->pair, (size_t)prod * sizeof(int32_t), st, device));
            FGK_CUDA(cudaMemsetAsync(I->pair, 0xFF, (size_t)prod * sizeof(int32_t), st));
            k_pair_fill<<<grid1d(n, device), 256, 0, st>>>(I->ra, I->rb, n, I->n_beta_strings, I->pair);
            FGK_LAUNCH_CHECK();
        }
        FGK_CUDA(cudaStreamSynchronize(st));
    }
    I->v.ra = I->ra; I->v.rb = I->rb; I->v.pair = I->pair; I->v.n_bstr = I->n_beta_strings;
    I->v.dets = d; I->v.n = n;
    I->v.table = I->table; I->v.mask = tsize - 1;
    I->v.aset = I->aset; I->v.amask = asz - 1;
    I->v.bset = I->bset; I->v.bmask = bsz - 1;
    *out = I;
    return FGK_OK;
}

extern "C" int fgk_index_destroy(fgk_index_t idx)
{
    if (!idx) return FGK_OK;
    cudaSetDevice(idx->device);
    // same guarantee as cudaFree (no kernel on any stream still reads the tables), but the blocks
    // go back to the pool instead of the driver
    cudaDeviceSynchronize();
    void* bufs[8] = {idx->table, idx->aset, idx->bset, idx->alist, idx->blist, idx->ra, idx->rb, idx->pair};
    for (void* b : bufs)
        if (b) cudaFreeAsync(b, 0);
    delete idx;
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
