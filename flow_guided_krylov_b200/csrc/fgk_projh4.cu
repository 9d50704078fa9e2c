// fgk_projh4.cu -- projected H built STRAIGHT into the packed SELL-32 operator (8 B/nnz).
//
// String-driven builder (fgk_lists.cuh): per distinct alpha / beta string of the basis the
// single / double replacement lists are built once (k_lists_*), then ONE WARP PER 32-ROW SLICE
// assembles the rows with LANE = ROW -- the layout the SELL-32 H.v kernels read.  Lane l emits
// entry k of its row into 16-byte unit  slice_base + (k >> 1) * 32 + l , so when the 32 rows of
// a slice run in lockstep (CAS-like / product bases: neighbouring determinants share their alpha
// string and have equally long lists) a warp store is one contiguous 512-byte line: the fill
// is coalesced without a transpose, and neither a CSR copy nor a CSR -> SELL pass nor a
// float32 re-pack exists.  Replaces fgk_projh_count + fgk_projh_fill + fgk_sell_fill +
// fgk_sell_pack_f32 for Krylov work (71 GB -> 17.8 GB of operator storage on configs[3]).
// Same values, filters and flavours as k_projh3 (reference molecular.py:471-516, skqd.py:374-419).
#include <cub/device/device_scan.cuh>

#include "fgk_internal.cuh"
#include "fgk_lists.cuh"

struct StrListView {
    const i64* sptr;            // singles of string t: singles[sptr[t] .. sptr[t+1])
    const i64* dptr;
    const LEntry* singles;
    const LEntry* doubles;
};

struct fgk_strlists {
    int device;
    cudaStream_t stream;
    StrListView a, b;
    void* bufs[8];
    i64 n_single[2], n_double[2];
};

// ---- list construction: one warp per string, scan of the ascending distinct-string list ------
template <bool FILL>
__global__ void __launch_bounds__(256)
k_lists(HamView H, const u64* __restrict__ list, i64 n_str, i64* __restrict__ cnt_s, i64* __restrict__ cnt_d,
        const i64* __restrict__ sptr, const i64* __restrict__ dptr, LEntry* __restrict__ singles,
        LEntry* __restrict__ doubles)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const i64 warp0 = (i64)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (i64)gridDim.x * 8;
    LdgF ldf;
    for (i64 t = warp0; t < n_str; t += nwarps) {
        const u64 w = __ldg(list + t);
        i64 ns = 0, nd = 0;
        const i64 s0 = FILL ? sptr[t] : 0, d0 = FILL ? dptr[t] : 0;
        for (i64 u0 = 0; u0 < n_str; u0 += 32) {
            const i64 u = u0 + lane;
            const u64 w2 = u < n_str ? __ldg(list + u) : w;
            const int pc = __popcll(w ^ w2);
            const unsigned bs = __ballot_sync(0xffffffffu, pc == 2), bd = __ballot_sync(0xffffffffu, pc == 4);
            if (FILL) {
                if (pc == 2) singles[s0 + ns + __popc(bs & lt)] = single_entry(H, w, w2, (int)u, ldf);
                if (pc == 4) doubles[d0 + nd + __popc(bd & lt)] = double_entry(H, w, w2, (int)u, ldf);
            }
            ns += __popc(bs);
            nd += __popc(bd);
        }
        if (!FILL && lane == 0) { cnt_s[t] = ns; cnt_d[t] = nd; }
    }
}

static int lists_grid(i64 n_str, int device)
{
    i64 need = (n_str + 7) / 8, cap = (i64)fgk_sm_count(device) * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int fgk_strlists_destroy(fgk_strlists_t L)
{
    if (!L) return FGK_OK;
    cudaSetDevice(L->device);
    for (void* b : L->bufs)
        if (b) cudaFreeAsync(b, L->stream);
    delete L;
    return FGK_OK;
}

extern "C" int fgk_strlists_create(fgk_ham_t h, fgk_index_t idx, void* stream, fgk_strlists_t* out)
{
    if (!h || !idx || !out) return fgk_fail(FGK_ERR_ARG, "fgk_strlists_create: bad argument");
    if (h->device != idx->device) return fgk_fail(FGK_ERR_ARG, "fgk_strlists_create: device mismatch");
    if (h->v.n_orb > 64) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_strlists_create: n_orb > 64");
    FGK_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    fgk_strlists* L = new fgk_strlists();
    L->device = h->device;
    L->stream = st;
    for (void*& b : L->bufs) b = nullptr;
#define FGK_CUDA_L(call)                                                                 \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            if (scratch) cudaFreeAsync(scratch, st);                                     \
            fgk_strlists_destroy(L);                                                     \
            return fgk_fail(FGK_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
                            cudaGetErrorString(e__));                                    \
        }                                                                                \
    } while (0)
    void* scratch = nullptr;
    const i64 ns[2] = {idx->n_alpha_strings, idx->n_beta_strings};
    const u64* lists[2] = {idx->alist, idx->blist};
    size_t cub_bytes = 0;
    const i64 nmax = ns[0] > ns[1] ? ns[0] : ns[1];
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const i64*)nullptr, (i64*)nullptr, (int)(nmax + 1), st);
    cub_bytes = (cub_bytes + 255) & ~(size_t)255;
    // scratch: cub | counts singles (nmax+1) | counts doubles (nmax+1)
    FGK_CUDA_L(fgk_pool_alloc(&scratch, cub_bytes + 2 * (size_t)(nmax + 1) * sizeof(i64), st, h->device));
    i64* cnt_s = (i64*)((char*)scratch + cub_bytes);
    i64* cnt_d = cnt_s + (nmax + 1);
    i64 totals[2][2] = {{0, 0}, {0, 0}};
    i64* ptrs[2][2];
    for (int spin = 0; spin < 2; spin++) {
        const i64 n = ns[spin];
        FGK_CUDA_L(fgk_pool_alloc(&L->bufs[4 * spin + 0], (size_t)(n + 1) * sizeof(i64), st, h->device));
        FGK_CUDA_L(fgk_pool_alloc(&L->bufs[4 * spin + 1], (size_t)(n + 1) * sizeof(i64), st, h->device));
        ptrs[spin][0] = (i64*)L->bufs[4 * spin + 0];
        ptrs[spin][1] = (i64*)L->bufs[4 * spin + 1];
        FGK_CUDA_L(cudaMemsetAsync(cnt_s, 0, 2 * (size_t)(nmax + 1) * sizeof(i64), st));
        if (n > 0) {
            k_lists<false><<<lists_grid(n, h->device), 256, 0, st>>>(h->v, lists[spin], n, cnt_s, cnt_d, nullptr,
                                                                     nullptr, nullptr, nullptr);
            FGK_CUDA_L(cudaGetLastError());
        }
        size_t tb = cub_bytes;
        FGK_CUDA_L(cub::DeviceScan::ExclusiveSum(scratch, tb, cnt_s, ptrs[spin][0], (int)(n + 1), st));
        tb = cub_bytes;
        FGK_CUDA_L(cub::DeviceScan::ExclusiveSum(scratch, tb, cnt_d, ptrs[spin][1], (int)(n + 1), st));
        FGK_CUDA_L(cudaMemcpyAsync(&totals[spin][0], ptrs[spin][0] + n, sizeof(i64), cudaMemcpyDeviceToHost, st));
        FGK_CUDA_L(cudaMemcpyAsync(&totals[spin][1], ptrs[spin][1] + n, sizeof(i64), cudaMemcpyDeviceToHost, st));
        FGK_CUDA_L(cudaStreamSynchronize(st));      // sizes of this spin's lists (cnt_* are reused)
    }
    for (int spin = 0; spin < 2; spin++) {
        const i64 n = ns[spin];
        L->n_single[spin] = totals[spin][0];
        L->n_double[spin] = totals[spin][1];
        FGK_CUDA_L(fgk_pool_alloc(&L->bufs[4 * spin + 2], (size_t)(totals[spin][0] + 1) * sizeof(LEntry), st, h->device));
        FGK_CUDA_L(fgk_pool_alloc(&L->bufs[4 * spin + 3], (size_t)(totals[spin][1] + 1) * sizeof(LEntry), st, h->device));
        if (n > 0) {
            k_lists<true><<<lists_grid(n, h->device), 256, 0, st>>>(
                h->v, lists[spin], n, nullptr, nullptr, ptrs[spin][0], ptrs[spin][1], (LEntry*)L->bufs[4 * spin + 2],
                (LEntry*)L->bufs[4 * spin + 3]);
            FGK_CUDA_L(cudaGetLastError());
        }
        StrListView& V = spin ? L->b : L->a;
        V.sptr = ptrs[spin][0];
        V.dptr = ptrs[spin][1];
        V.singles = (const LEntry*)L->bufs[4 * spin + 2];
        V.doubles = (const LEntry*)L->bufs[4 * spin + 3];
    }
    cudaFreeAsync(scratch, st);
#undef FGK_CUDA_L
    *out = L;
    return FGK_OK;
}

extern "C" int fgk_strlists_info(fgk_strlists_t L, int64_t* n_single_a, int64_t* n_double_a, int64_t* n_single_b,
                                 int64_t* n_double_b)
{
    if (!L) return fgk_fail(FGK_ERR_ARG, "fgk_strlists_info: null handle");
    if (n_single_a) *n_single_a = L->n_single[0];
    if (n_double_a) *n_double_a = L->n_double[0];
    if (n_single_b) *n_single_b = L->n_single[1];
    if (n_double_b) *n_double_b = L->n_double[1];
    return FGK_OK;
}

// ---- row assembly: one warp per slice, lane = row -----------------------------------------------
// MODE 0: upper bound of the off-diagonal row length from the list lengths alone
// MODE 1: exact off-diagonal row length (same walk, no stores); slices s = 0, stride, 2 stride, ...
// MODE 2: fill the packed SELL-32 units + actual row lengths
template <int MODE, bool DENSE>
__global__ void __launch_bounds__(256, 4)
k_projh4(HamView H, IndexView I, StrListView LA, StrListView LB, const u64* __restrict__ alist,
         const u64* __restrict__ blist, i64 row_begin, i64 row_end, int mode, i64 slice_stride,
         i64* __restrict__ counts, const i64* __restrict__ slice_ptr, uint4* __restrict__ packed,
         int32_t* __restrict__ rowlen, int* inexact)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = (i64)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (i64)gridDim.x * 8;
    const i64 rows = row_end - row_begin, n_slices = (rows + 31) >> 5;
    const bool sym = (mode & FGK_H_SYM) != 0, drop0 = (mode & FGK_H_DROP_ZEROS) != 0;
    const int nbs = (int)I.n_bstr;
    bool bad = false;
    for (i64 s = warp0 * slice_stride; s < n_slices; s += nwarps * slice_stride) {
        const i64 rl = s * 32 + lane, i = row_begin + rl;
        const bool valid = i < row_end;
        int ia = 0, ib = 0;
        i64 sa0 = 0, sb0 = 0, da0 = 0, db0 = 0;
        int nsa = 0, nsb = 0, nda = 0, ndb = 0;
        if (valid) {
            ia = __ldg(I.ra + i);
            ib = __ldg(I.rb + i);
            sa0 = __ldg(LA.sptr + ia); nsa = (int)(__ldg(LA.sptr + ia + 1) - sa0);
            sb0 = __ldg(LB.sptr + ib); nsb = (int)(__ldg(LB.sptr + ib + 1) - sb0);
            da0 = __ldg(LA.dptr + ia); nda = (int)(__ldg(LA.dptr + ia + 1) - da0);
            db0 = __ldg(LB.dptr + ib); ndb = (int)(__ldg(LB.dptr + ib + 1) - db0);
        }
        if (MODE == 0) {
            if (valid) counts[rl] = (i64)nsa + nsb + nda + ndb + (i64)nsa * nsb;
            continue;
        }
        const i64 base = MODE == 2 ? __ldg(slice_ptr + s) : 0;
        const i64 width = MODE == 2 ? (__ldg(slice_ptr + s + 1) - base) >> 5 : 0;     // pair-columns
        int pos = 0;
        float pend_v = 0.f;
        int pend_c = 0;
        auto column = [&](int ra2, int rb2) -> int {
            if (DENSE) return __ldg(I.pair + (i64)ra2 * nbs + rb2);
            fgk_det o = {__ldg(alist + ra2), __ldg(blist + rb2)};
            return index_find(I, o);
        };
        auto emit = [&](int j, float f) {
            if (MODE == 2) {
                if (pos & 1) {
                    if ((pos >> 1) < width)
                        packed[base + (i64)(pos >> 1) * 32 + lane] =
                            make_uint4(__float_as_uint(pend_v), __float_as_uint(f), (unsigned)pend_c, (unsigned)j);
                } else {
                    pend_v = f;
                    pend_c = j;
                }
            }
            pos++;
        };
        auto value = [&](int j, float vij, float vji) {
            float f;
            double v;
            bool exact;
            if (j >= 0 && entry_value_f32(sym, drop0, vij, vji, f, v, exact)) {
                if (!exact) bad = true;
                emit(j, f);
            }
        };
        // singles and same-spin doubles: the other spin keeps its own string.  Four list entries per
        // round, their pair-table loads in flight together.
        auto run_list = [&](const LEntry* __restrict__ list, i64 base0, int len, bool alpha_side) {
            const int mlen = __reduce_max_sync(0xffffffffu, len);
            for (int k0 = 0; k0 < mlen; k0 += 4) {
                float vij[4], vji[4];
                int j[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    j[u] = -1;
                    vij[u] = vji[u] = 0.f;
                    if (k0 + u < len) {
                        const LEntry e = list[base0 + k0 + u];
                        vij[u] = e.vij;
                        vji[u] = e.vji;
                        j[u] = alpha_side ? column(e.rank, ib) : column(ia, e.rank);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) value(j[u], vij[u], vji[u]);
            }
        };
        run_list(LA.singles, sa0, nsa, true);
        run_list(LB.singles, sb0, nsb, false);
        run_list(LA.doubles, da0, nda, true);
        run_list(LB.doubles, db0, ndb, false);
        // alpha-beta doubles: beta single outside (differs per lane), alpha single inside (neighbouring
        // rows share their alpha string: the inner loads are warp-wide broadcasts)
        const int mb = __reduce_max_sync(0xffffffffu, nsb), ma = __reduce_max_sync(0xffffffffu, nsa);
        const int n2 = H.n_orb * H.n_orb;
        for (int kb = 0; kb < mb; kb++) {
            const bool vb = kb < nsb;
            LEntry eb;
            eb.rank = 0; eb.vij = eb.vji = 0.f; eb.info = 0u;
            if (vb) eb = LB.singles[sb0 + kb];
            const float* const g_b = H.g + ((eb.info >> 12) & 0xfffu);     // bra-side column of g
            const float* const g_k = H.g + (eb.info & 0xfffu);            // ket-side column
            // four alpha singles per round: their pair-table and integral loads are all issued before
            // the first result is used (ncu r02l: 45 % of the stall samples sat on the dependent
            // pair -> g load chain of the one-at-a-time loop); the integral loads do not depend on
            // the pair lookup, so they are unconditional (offsets of a zeroed entry are valid)
            for (int ka0 = 0; ka0 < ma; ka0 += 4) {
                unsigned info[4];
                int j[4];
                float rb[4], rk[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int ka = ka0 + u;
                    info[u] = 0u;
                    j[u] = -1;
                    if (vb && ka < nsa) {
                        const LEntry ea = LA.singles[sa0 + ka];
                        info[u] = ea.info;
                        j[u] = column(ea.rank, eb.rank);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    rb[u] = __ldg(g_b + (size_t)((info[u] >> 12) & 0xfffu) * n2);
                    rk[u] = sym ? __ldg(g_k + (size_t)(info[u] & 0xfffu) * n2) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (j[u] >= 0) {
                        const unsigned px = info[u] ^ eb.info;
                        const float vij = (((px >> 25) ^ 1u) & 1u) ? -rb[u] : rb[u];      // parity (bra side) = pb_a ^ pb_b ^ 1
                        const float vji = (((px >> 24) ^ 1u) & 1u) ? -rk[u] : rk[u];      // parity (ket side) = pk_a ^ pk_b ^ 1
                        value(j[u], vij, vji);
                    }
                }
            }
        }
        if (MODE == 1) {
            if (valid) counts[rl] = pos;
            continue;
        }
        // tail of the row: the half-filled unit, then zero padding up to the slice width
        i64 k2 = pos >> 1;
        if ((pos & 1) && k2 < width) {
            packed[base + k2 * 32 + lane] = make_uint4(__float_as_uint(pend_v), 0u, (unsigned)pend_c, 0u);
            k2++;
        }
        for (; k2 < width; k2++) packed[base + k2 * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) rowlen[rl] = pos;
        if (((pos + 1) >> 1) > width) bad = true;      // the caller's slice was too narrow
    }
    if (MODE == 2 && bad) atomicExch(inexact, 1);
}

template <int MODE>
static int launch_projh4(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t L, i64 row_begin, i64 row_end, int mode,
                         i64 stride, i64* counts, const i64* slice_ptr, void* packed, int32_t* rowlen, int* inexact,
                         cudaStream_t st)
{
    const i64 n_slices = (row_end - row_begin + 31) / 32;
    const i64 work = (n_slices + stride - 1) / stride;
    i64 need = (work + 7) / 8, cap = (i64)fgk_sm_count(h->device) * 8;
    if (need < 1) need = 1;
    const int grid = (int)(need < cap ? need : cap);
    if (idx->pair)
        k_projh4<MODE, true><<<grid, 256, 0, st>>>(h->v, idx->v, L->a, L->b, idx->alist, idx->blist, row_begin, row_end,
                                                   mode, stride, counts, slice_ptr, (uint4*)packed, rowlen, inexact);
    else
        k_projh4<MODE, false><<<grid, 256, 0, st>>>(h->v, idx->v, L->a, L->b, idx->alist, idx->blist, row_begin, row_end,
                                                    mode, stride, counts, slice_ptr, (uint4*)packed, rowlen, inexact);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

static int check_packed_args(const char* who, fgk_ham_t h, fgk_index_t idx, fgk_strlists_t L, i64 row_begin, i64 row_end)
{
    if (!h || !idx || !L) return fgk_fail(FGK_ERR_ARG, "%s: null handle", who);
    if (row_begin < 0 || row_end > idx->v.n || row_begin > row_end) return fgk_fail(FGK_ERR_ARG, "%s: bad row range", who);
    if (h->device != idx->device || h->device != L->device) return fgk_fail(FGK_ERR_ARG, "%s: device mismatch", who);
    if (!idx->ra || !idx->rb) return fgk_fail(FGK_ERR_ARG, "%s: the index has no rank form (empty basis)", who);
    return FGK_OK;
}

extern "C" int fgk_projh_packed_bound(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t L, int64_t row_begin,
                                      int64_t row_end, int64_t* counts, void* stream)
{
    if (row_begin == row_end) return FGK_OK;
    int rc = check_packed_args("fgk_projh_packed_bound", h, idx, L, row_begin, row_end);
    if (rc != FGK_OK) return rc;
    if (!counts) return fgk_fail(FGK_ERR_ARG, "fgk_projh_packed_bound: null counts");
    FGK_CUDA(cudaSetDevice(h->device));
    return launch_projh4<0>(h, idx, L, row_begin, row_end, 0, 1, (i64*)counts, nullptr, nullptr, nullptr, nullptr,
                            (cudaStream_t)stream);
}

extern "C" int fgk_projh_packed_count(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t L, int64_t row_begin,
                                      int64_t row_end, int mode, int64_t slice_stride, int64_t* counts, void* stream)
{
    if (row_begin == row_end) return FGK_OK;
    int rc = check_packed_args("fgk_projh_packed_count", h, idx, L, row_begin, row_end);
    if (rc != FGK_OK) return rc;
    if (!counts || slice_stride < 1) return fgk_fail(FGK_ERR_ARG, "fgk_projh_packed_count: bad argument");
    FGK_CUDA(cudaSetDevice(h->device));
    return launch_projh4<1>(h, idx, L, row_begin, row_end, mode, slice_stride, (i64*)counts, nullptr, nullptr, nullptr,
                            nullptr, (cudaStream_t)stream);
}

extern "C" int fgk_projh_packed_fill(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t L, int64_t row_begin,
                                     int64_t row_end, int mode, const int64_t* slice_ptr, void* packed,
                                     int32_t* row_len, int* flag, void* stream)
{
    if (row_begin == row_end) return FGK_OK;
    int rc = check_packed_args("fgk_projh_packed_fill", h, idx, L, row_begin, row_end);
    if (rc != FGK_OK) return rc;
    if (!slice_ptr || !packed || !row_len || !flag || ((uintptr_t)packed & 15))
        return fgk_fail(FGK_ERR_ARG, "fgk_projh_packed_fill: bad argument");
    FGK_CUDA(cudaSetDevice(h->device));
    return launch_projh4<2>(h, idx, L, row_begin, row_end, mode, 1, nullptr, (const i64*)slice_ptr, packed, row_len, flag,
                            (cudaStream_t)stream);
}
