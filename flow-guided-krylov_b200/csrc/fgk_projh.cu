// fgk_projh.cu -- K5: projected Hamiltonian over an indexed basis, as CSR rows.
//
// One warp per ROW determinant i ("bra mode"): it walks every excitation x of
// D_i, forms j = D_i + x, rejects it with the alpha/beta string sets, probes the
// full-key table and, on a hit, evaluates <i|H|j> exactly as the reference's
// get_connections(j) would report the connection j -> i (and <j|H|i> for the
// symmetrised flavour).  Count pass -> exclusive scan (caller) -> fill pass; the
// rank of an entry inside its row comes from warp ballots, so the layout is
// deterministic.  Replaces matrix_elements_fast (molecular.py:471-516),
// get_sparse_matrix_elements (:580-638), _build_subspace_hamiltonian
// (skqd.py:374-419) -- all of which loop over kets in Python and probe a dict.
#include "fgk_internal.cuh"

template <bool FILL>
__global__ void __launch_bounds__(FGK_BLOCK)
k_projh(HamView H, IndexView I, i64 row_begin, i64 row_end, int mode, i64* __restrict__ counts,
        const i64* __restrict__ row_ptr, int32_t* __restrict__ cols, double* __restrict__ vals)
{
    __shared__ WarpLists s_lists[FGK_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const bool sym = (mode & FGK_H_SYM) != 0, drop0 = (mode & FGK_H_DROP_ZEROS) != 0;
    LdgF ldf;
    LdgD ldd;
    for (i64 i = row_begin + warp0; i < row_end; i += nwarps) {
        ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + i);
        fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, H.n_orb, d, s_lists[wib], lane);
        i64 pos = FILL ? row_ptr[i - row_begin] : 0;
        // diagonal first (always stored, like skqd.py:394-397)
        if (FILL && lane == 0) {
            cols[pos] = (int32_t)i;
            vals[pos] = diag_element(H, d, ldd);
        }
        pos += 1;
        auto visit = [&](bool valid, const Excitation& x) {
            int j = -1;
            double v = 0.0;
            bool keep = false;
            if (valid) {
                fgk_det o = apply_excitation(d, c.n, x);
                j = index_find_filtered(I, o, x.cls);
                if (j >= 0) {
                    float vij = 0.f, vji = 0.f;
                    bool kij = bra_element(H, d, x, ldf, vij);
                    bool kji = sym ? ket_element(H, d, x, ldf, vji) : false;
                    keep = kij || kji;
                    v = sym ? 0.5 * ((double)(kij ? vij : 0.f) + (double)(kji ? vji : 0.f))
                            : (double)vij;
                    if (drop0 && v == 0.0) keep = false;
                }
            }
            unsigned b = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) {
                i64 o = pos + __popc(b & lt);
                cols[o] = j;
                vals[o] = v;
            }
            pos += __popc(b);
        };
        warp_enumerate(
            c, lane,
            [&](bool va, bool vb, int p, int q) {
                Excitation x;
                x.h0 = q; x.e0 = p; x.h1 = 0; x.e1 = 0;
                x.cls = 0;
                visit(va, x);
                x.cls = 1;
                visit(vb, x);
            },
            visit);
        if (!FILL && lane == 0) counts[i - row_begin] = pos;
    }
}

static int grid_rows(i64 rows, int device)
{
    i64 need = (rows + FGK_WARPS_PER_BLOCK - 1) / FGK_WARPS_PER_BLOCK;
    i64 cap = (i64)fgk_sm_count(device) * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int fgk_projh_count(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end,
                               int mode, int64_t* counts, void* stream)
{
    if (!h || !idx) return fgk_fail(FGK_ERR_ARG, "fgk_projh_count: null handle");
    if (row_begin < 0 || row_end > idx->v.n || row_begin > row_end)
        return fgk_fail(FGK_ERR_ARG, "fgk_projh_count: bad row range");
    if (row_begin == row_end) return FGK_OK;
    if (!counts) return fgk_fail(FGK_ERR_ARG, "fgk_projh_count: null counts");
    if (h->device != idx->device) return fgk_fail(FGK_ERR_ARG, "fgk_projh_count: device mismatch");
    FGK_CUDA(cudaSetDevice(h->device));
    k_projh<false><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        h->v, idx->v, row_begin, row_end, mode, (i64*)counts, nullptr, nullptr, nullptr);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_projh_fill(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end,
                              int mode, const int64_t* row_ptr, int32_t* cols, double* vals,
                              void* stream)
{
    if (!h || !idx) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill: null handle");
    if (row_begin < 0 || row_end > idx->v.n || row_begin > row_end)
        return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill: bad row range");
    if (row_begin == row_end) return FGK_OK;
    if (!row_ptr || !cols || !vals) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill: null pointer");
    if (h->device != idx->device) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill: device mismatch");
    FGK_CUDA(cudaSetDevice(h->device));
    k_projh<true><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        h->v, idx->v, row_begin, row_end, mode, nullptr, (const i64*)row_ptr, cols, vals);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// ---- row sort: one CTA per row, ascending-only bitonic network --------------------------
// (virtual +inf padding up to the next power of two; every compare-exchange moves
// the smaller column to the lower index, so padded slots never move).
// Rows up to SORT_SMEM_MAX entries are sorted in shared memory, longer rows in place.
static const int SORT_SMEM_MAX = 4096;

template <bool SMEM>
__device__ __forceinline__ void bitonic_row(int32_t* c, double* v, int len, int N)
{
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                int a, b;
                if (j == (k >> 1)) {           // flip step: mirror inside the block of size k
                    int blk = t / j, off = t - blk * j;
                    a = blk * k + off;
                    b = blk * k + (k - 1 - off);
                } else {                        // disperse steps
                    int blk = t / j, off = t - blk * j;
                    a = blk * 2 * j + off;
                    b = a + j;
                }
                if (b < len) {
                    int32_t ca = c[a], cb = c[b];
                    if (ca > cb) {
                        double va = v[a], vb = v[b];
                        c[a] = cb; c[b] = ca;
                        v[a] = vb; v[b] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
k_sort_rows(i64 n_rows, const i64* __restrict__ row_ptr, int32_t* __restrict__ cols,
            double* __restrict__ vals)
{
    extern __shared__ unsigned char s_raw[];
    double* sv = reinterpret_cast<double*>(s_raw);
    int32_t* sc = reinterpret_cast<int32_t*>(s_raw + sizeof(double) * SORT_SMEM_MAX);
    for (i64 r = blockIdx.x; r < n_rows; r += gridDim.x) {
        i64 s = row_ptr[r];
        int len = (int)(row_ptr[r + 1] - s);
        if (len < 2) continue;                 // uniform per block
        int N = 1;
        while (N < len) N <<= 1;
        if (len <= SORT_SMEM_MAX) {
            for (int t = threadIdx.x; t < len; t += blockDim.x) { sc[t] = cols[s + t]; sv[t] = vals[s + t]; }
            __syncthreads();
            bitonic_row<true>(sc, sv, len, N);
            for (int t = threadIdx.x; t < len; t += blockDim.x) { cols[s + t] = sc[t]; vals[s + t] = sv[t]; }
            __syncthreads();
        } else {
            __syncthreads();
            bitonic_row<false>(cols + s, vals + s, len, N);
        }
    }
}

extern "C" int fgk_csr_sort_rows(int64_t n_rows, const int64_t* row_ptr, int32_t* cols, double* vals,
                                 int device, void* stream)
{
    if (n_rows == 0) return FGK_OK;
    if (!row_ptr || !cols || !vals || n_rows < 0)
        return fgk_fail(FGK_ERR_ARG, "fgk_csr_sort_rows: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    size_t smem = (sizeof(double) + sizeof(int32_t)) * SORT_SMEM_MAX;   // 48 KB
    static bool attr_set[64] = {false};
    if (!attr_set[device & 63]) {
        FGK_CUDA(cudaFuncSetAttribute(k_sort_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[device & 63] = true;
    }
    i64 cap = (i64)fgk_sm_count(device) * 4;
    int grid = (int)(n_rows < cap ? n_rows : cap);
    k_sort_rows<<<grid, 256, smem, (cudaStream_t)stream>>>(n_rows, (const i64*)row_ptr, cols, vals);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}
