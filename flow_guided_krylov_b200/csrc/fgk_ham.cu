// fgk_ham.cu -- error plumbing, Hamiltonian handle, K1 pack/unpack, K2 diagonal,
// K3 connection enumeration in the reference's emission order.
#include <stdarg.h>
#include <vector>

#include "fgk_internal.cuh"
#include "fgk_tables.h"

// ---- errors ---------------------------------------------------------------------
std::string& fgk_err_slot()
{
    static thread_local std::string s;
    return s;
}

int fgk_fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    fgk_err_slot() = buf;
    return code;
}

int fgk_sm_count(int device)
{
    static int cache[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (!cache[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0)
            v = 148;
        cache[device] = v;
    }
    return cache[device];
}

extern "C" int fgk_version(void) { return 100; }
extern "C" const char* fgk_last_error(void) { return fgk_err_slot().c_str(); }

extern "C" int fgk_device_info(int device, int* sm_count, size_t* l2_bytes, size_t* free_bytes,
                               size_t* total_bytes)
{
    FGK_CUDA(cudaSetDevice(device));
    int v = 0;
    FGK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = v;
    FGK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    if (l2_bytes) *l2_bytes = (size_t)v;
    size_t f = 0, t = 0;
    FGK_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return FGK_OK;
}

// ---- Hamiltonian handle -------------------------------------------------------------
template <class T>
static int upload(const std::vector<T>& src, T** dst)
{
    FGK_CUDA(cudaMalloc((void**)dst, sizeof(T) * (src.size() ? src.size() : 1)));
    FGK_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return FGK_OK;
}

extern "C" int fgk_ham_create(const double* h1_host, const double* g_host, int n_orb, int n_alpha,
                              int n_beta, double e_nuc, int device, fgk_ham_t* out)
{
    if (!h1_host || !g_host || !out) return fgk_fail(FGK_ERR_ARG, "fgk_ham_create: null pointer");
    if (n_orb < 1 || n_orb > 64)
        return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_ham_create: n_orb=%d outside 1..64", n_orb);
    if (n_alpha < 0 || n_beta < 0 || n_alpha > n_orb || n_beta > n_orb)
        return fgk_fail(FGK_ERR_ARG, "fgk_ham_create: bad electron counts");
    FGK_CUDA(cudaSetDevice(device));
    HostTables T;
    build_host_tables(h1_host, g_host, n_orb, T);
    fgk_ham* H = new fgk_ham();
    H->device = device;
    // diagonal tables: h_pp + nibble row-sum tables in ONE contiguous, 16-byte-granular buffer
    // (a single TMA bulk copy stages it into shared memory); plain (J-K) / J matrices kept for
    // the pair-loop fallback when the nibble tables exceed shared memory (n_orb > 56)
    const size_t npad = ((size_t)n_orb + 1) & ~(size_t)1, n2 = (size_t)n_orb * n_orb;
    const size_t nn = (size_t)n_orb * T.nchunk * 16;
    std::vector<double> dt(npad + 2 * nn, 0.0), jk(2 * n2, 0.0);
    for (int p = 0; p < n_orb; p++) dt[p] = T.hdiag[p];
    for (size_t i = 0; i < nn; i++) { dt[npad + i] = T.nib_jk[i]; dt[npad + nn + i] = T.nib_jab[i]; }
    for (size_t i = 0; i < n2; i++) { jk[i] = T.jks[i]; jk[n2 + i] = T.jab[i]; }
    // integral tables h1 | g | w in one buffer, every segment padded to 16 bytes: a single TMA bulk
    // copy stages them in shared memory when they fit (n_orb <= 12: 166 KB; all of the reference's
    // molecules), see k_conn
    auto pad4 = [](size_t x) { return (x + 3) & ~(size_t)3; };
    const size_t o_g = pad4(T.h1.size()), o_w = o_g + pad4(T.g.size()), n_it = o_w + pad4(T.w.size());
    std::vector<float> it(n_it, 0.f);
    std::copy(T.h1.begin(), T.h1.end(), it.begin());
    std::copy(T.g.begin(), T.g.end(), it.begin() + o_g);
    std::copy(T.w.begin(), T.w.end(), it.begin() + o_w);
    int rc;
    if ((rc = upload(it, &H->itab)) || (rc = upload(dt, &H->dtab)) || (rc = upload(jk, &H->jkab))) {
        delete H;
        return rc;
    }
    H->h1 = H->itab;
    H->g = H->itab + o_g;
    H->w = H->itab + o_w;
    H->itab_bytes = n_it * sizeof(float) <= 176 * 1024 ? (unsigned)(n_it * sizeof(float)) : 0u;
    const size_t bytes = dt.size() * sizeof(double);
    const bool nib_ok = bytes <= 208 * 1024;
    H->dtab_bytes = nib_ok ? (unsigned)bytes : 0;
    H->v.n_orb = n_orb; H->v.n_alpha = n_alpha; H->v.n_beta = n_beta; H->v.e_nuc = e_nuc;
    H->v.h1 = H->h1; H->v.g = H->g; H->v.w = H->w;
    H->v.hdiag = H->dtab; H->v.jks = H->jkab; H->v.jab = H->jkab + n2;
    H->v.nchunk = T.nchunk;
    H->v.nib_jk = nib_ok ? H->dtab + npad : nullptr;
    H->v.nib_jab = nib_ok ? H->dtab + npad + nn : nullptr;
    *out = H;
    return FGK_OK;
}

extern "C" int fgk_ham_destroy(fgk_ham_t h)
{
    if (!h) return FGK_OK;
    cudaSetDevice(h->device);
    cudaFree(h->itab);
    cudaFree(h->dtab); cudaFree(h->jkab);
    delete h;
    return FGK_OK;
}

// ---- K1 pack / unpack ------------------------------------------------------------------
// one warp per determinant: lanes read 32 consecutive int64 sites (coalesced 256 B),
// turn them into word contributions and OR-reduce across the warp.
__global__ void __launch_bounds__(FGK_BLOCK)
k_pack_i64(const int64_t* __restrict__ cfg, i64 n, int n_orb, fgk_det* __restrict__ dets)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const int S = 2 * n_orb;
    for (i64 j = warp0; j < n; j += nwarps) {
        const int64_t* row = cfg + j * S;
        unsigned alo = 0, ahi = 0, blo = 0, bhi = 0;
        for (int s = lane; s < S; s += 32) {
            if (row[s] != 0) {
                u64 bit = s < n_orb ? orb_bit(n_orb, s) : orb_bit(n_orb, s - n_orb);
                if (s < n_orb) { alo |= (unsigned)bit; ahi |= (unsigned)(bit >> 32); }
                else { blo |= (unsigned)bit; bhi |= (unsigned)(bit >> 32); }
            }
        }
        alo = __reduce_or_sync(0xffffffffu, alo);
        ahi = __reduce_or_sync(0xffffffffu, ahi);
        blo = __reduce_or_sync(0xffffffffu, blo);
        bhi = __reduce_or_sync(0xffffffffu, bhi);
        if (lane == 0) {
            ulonglong2 o;
            o.x = ((u64)ahi << 32) | alo;
            o.y = ((u64)bhi << 32) | blo;
            reinterpret_cast<ulonglong2*>(dets)[j] = o;
        }
    }
}

// one thread per output site: coalesced 8-byte stores
__global__ void __launch_bounds__(256)
k_unpack_i64(const fgk_det* __restrict__ dets, i64 n, int n_orb, int64_t* __restrict__ cfg)
{
    const int S = 2 * n_orb;
    const i64 total = n * S;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (i64)gridDim.x * blockDim.x) {
        i64 j = t / S;
        int s = (int)(t - j * S);
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + j);
        u64 w = s < n_orb ? d.x : d.y;
        int p = s < n_orb ? s : s - n_orb;
        cfg[t] = (int64_t)((w >> (n_orb - 1 - p)) & 1ull);
    }
}

static int grid_for(i64 work_items, int items_per_block, int device, int blocks_per_sm)
{
    i64 need = (work_items + items_per_block - 1) / items_per_block;
    i64 cap = (i64)fgk_sm_count(device) * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int fgk_pack_i64(const int64_t* cfg, int64_t n, int n_orb, uint64_t* dets, int device,
                            void* stream)
{
    if (n_orb < 1 || n_orb > 64) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_pack_i64: n_orb=%d", n_orb);
    if (n == 0) return FGK_OK;
    if (!cfg || !dets || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_pack_i64: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    k_pack_i64<<<grid_for(n, FGK_WARPS_PER_BLOCK, device, 8), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        cfg, n, n_orb, (fgk_det*)dets);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_unpack_i64(const uint64_t* dets, int64_t n, int n_orb, int64_t* cfg, int device,
                              void* stream)
{
    if (n_orb < 1 || n_orb > 64) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_unpack_i64: n_orb=%d", n_orb);
    if (n == 0) return FGK_OK;
    if (!cfg || !dets || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_unpack_i64: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    k_unpack_i64<<<grid_for(n * 2 * n_orb, 256, device, 8), 256, 0, (cudaStream_t)stream>>>(
        (const fgk_det*)dets, n, n_orb, cfg);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// ---- K2 diagonal -----------------------------------------------------------------------
// one thread per determinant; h_pp and the nibble row-sum tables (66 KB at 32 orbitals, 148 KB
// at 48) are staged in shared memory by ONE TMA bulk copy per CTA; a row sum over an occupation
// word is then ceil(n/4) independent shared-memory reads.  Above 56 orbitals the tables do not
// fit and the pair-loop form reads the plain matrices through the read-only path.
static const int DIAG_BLOCK = 1024;

__global__ void __launch_bounds__(DIAG_BLOCK)
k_diag(HamView H, unsigned tab_bytes, const fgk_det* __restrict__ dets, i64 n, double* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    if (tab_bytes) {
        tma_stage_table(s_raw, H.hdiag, tab_bytes, &s_mbar);
        const double* s_tab = reinterpret_cast<const double*>(s_raw);
        const double* g0 = H.hdiag;
        H.nib_jk = s_tab + (H.nib_jk - g0);
        H.nib_jab = s_tab + (H.nib_jab - g0);
        H.hdiag = s_tab;
    }
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + j);
        fgk_det dd = {d.x, d.y};
        out[j] = tab_bytes ? diag_element(H, dd, LdsD())
                           : diag_element(H, dd, [](const double* p) { return __ldg(p); });
    }
}

extern "C" int fgk_diag(fgk_ham_t h, const uint64_t* dets, int64_t n, double* out, void* stream)
{
    if (!h) return fgk_fail(FGK_ERR_ARG, "fgk_diag: null handle");
    if (n == 0) return FGK_OK;
    if (!dets || !out || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_diag: bad argument");
    FGK_CUDA(cudaSetDevice(h->device));
    size_t smem = h->dtab_bytes;
    static bool attr_set[64] = {false};
    if (smem > 48 * 1024 && !attr_set[h->device & 63]) {
        FGK_CUDA(cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        attr_set[h->device & 63] = true;
    }
    k_diag<<<grid_for(n, DIAG_BLOCK, h->device, 2), DIAG_BLOCK, smem, (cudaStream_t)stream>>>(
        h->v, h->dtab_bytes, (const fgk_det*)dets, n, out);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// ---- K3 connections, reference emission order ----------------------------------------------
// one warp per source determinant.  FILL = false: count survivors of the
// reference's |val| > 1e-12 filters; FILL = true: write them at offsets[j] + rank,
// rank = position in the reference's emission order, obtained by ballot + popc.
// STAGED: the integral tables h1 | g | w (one contiguous buffer) are staged in shared memory by a
// single TMA bulk copy (cp.async.bulk + mbarrier, as k_diag does for the diagonal tables) and
// read with ld.shared; used when they fit (n_orb <= 12).  Larger tables (4.2 MB each at 32
// orbitals) stay L2-resident and are read through the read-only path.
template <bool STAGED> struct ConnLoader { typedef LdgF type; };
template <> struct ConnLoader<true> { typedef LdsF type; };

template <bool FILL, bool STAGED>
__global__ void __launch_bounds__(FGK_BLOCK)
k_conn(HamView H, unsigned tab_bytes, const fgk_det* __restrict__ dets, i64 n, i64* __restrict__ counts,
       const i64* __restrict__ offsets, fgk_det* __restrict__ out_dets,
       float* __restrict__ out_elems, i64* __restrict__ out_src)
{
    extern __shared__ __align__(128) unsigned char s_itab[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ WarpLists s_lists[FGK_WARPS_PER_BLOCK];
    if (STAGED) {
        tma_stage_table(s_itab, H.h1, tab_bytes, &s_mbar);
        const float* s_tab = reinterpret_cast<const float*>(s_itab);
        const float* g0 = H.h1;
        H.g = s_tab + (H.g - g0);
        H.w = s_tab + (H.w - g0);
        H.h1 = s_tab;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    typename ConnLoader<STAGED>::type ldf;
    for (i64 j = warp0; j < n; j += nwarps) {
        ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(dets) + j);
        fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, H.n_orb, d, s_lists[wib], lane);
        i64 pos = FILL ? offsets[j] : 0;
        auto put = [&](i64 o, const Excitation& x, float e) {
            if (out_dets) {
                fgk_det t = apply_excitation(d, c.n, x);
                reinterpret_cast<ulonglong2*>(out_dets)[o] = make_ulonglong2(t.a, t.b);
            }
            if (out_elems) out_elems[o] = e;
            if (out_src) out_src[o] = j;
        };
        warp_enumerate(
            c, lane,
            [&](bool va, bool vb, int p, int q) {
                Excitation x;
                x.h0 = q; x.e0 = p; x.h1 = 0; x.e1 = 0;
                float ea = 0.f, eb = 0.f;
                x.cls = 0;
                bool ka = va && ket_element_fast(H, d, x, ldf, ea);
                x.cls = 1;
                bool kb = vb && ket_element_fast(H, d, x, ldf, eb);
                unsigned ba = __ballot_sync(0xffffffffu, ka), bb = __ballot_sync(0xffffffffu, kb);
                if (FILL) {
                    i64 o = pos + __popc(ba & lt) + __popc(bb & lt);
                    if (ka) { x.cls = 0; put(o, x, ea); }
                    if (kb) { x.cls = 1; put(o + (ka ? 1 : 0), x, eb); }
                }
                pos += __popc(ba) + __popc(bb);
            },
            [&](bool valid, const Excitation& x) {
                float e = 0.f;
                bool k = valid && ket_element_fast(H, d, x, ldf, e);
                unsigned b = __ballot_sync(0xffffffffu, k);
                if (FILL && k) put(pos + __popc(b & lt), x, e);
                pos += __popc(b);
            });
        if (!FILL && lane == 0) counts[j] = pos;
    }
}

static int conn_ctas_per_sm(unsigned tab_bytes)
{
    int c = (int)((200u * 1024u) / (tab_bytes + 4096u));
    return c < 1 ? 1 : (c > 8 ? 8 : c);
}

template <bool FILL>
static int conn_smem_attr(unsigned bytes)
{
    static bool done = false;           // per instantiation; the limit is per function, not per device state
    if (!done && bytes > 48 * 1024) {
        FGK_CUDA(cudaFuncSetAttribute(k_conn<FILL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    return FGK_OK;
}

extern "C" int fgk_conn_count(fgk_ham_t h, const uint64_t* dets, int64_t n, int64_t* counts,
                              void* stream)
{
    if (!h) return fgk_fail(FGK_ERR_ARG, "fgk_conn_count: null handle");
    if (n == 0) return FGK_OK;
    if (!dets || !counts || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_conn_count: bad argument");
    FGK_CUDA(cudaSetDevice(h->device));
    if (h->itab_bytes) {
        int rc = conn_smem_attr<false>(h->itab_bytes);
        if (rc != FGK_OK) return rc;
        // the staging copy is paid per CTA: long-lived CTAs (grid-stride over the determinants), as many
        // per SM as the staged tables leave room for
        k_conn<false, true><<<grid_for(n, 4 * FGK_WARPS_PER_BLOCK, h->device, conn_ctas_per_sm(h->itab_bytes)), FGK_BLOCK, h->itab_bytes, (cudaStream_t)stream>>>(
            h->v, h->itab_bytes, (const fgk_det*)dets, n, (i64*)counts, nullptr, nullptr, nullptr, nullptr);
    } else
    k_conn<false, false><<<grid_for(n, FGK_WARPS_PER_BLOCK, h->device, 8), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        h->v, 0u, (const fgk_det*)dets, n, (i64*)counts, nullptr, nullptr, nullptr, nullptr);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_conn_fill(fgk_ham_t h, const uint64_t* dets, int64_t n, const int64_t* offsets,
                             uint64_t* out_dets, float* out_elems, int64_t* out_src, void* stream)
{
    if (!h) return fgk_fail(FGK_ERR_ARG, "fgk_conn_fill: null handle");
    if (n == 0) return FGK_OK;
    if (!dets || !offsets || n < 0) return fgk_fail(FGK_ERR_ARG, "fgk_conn_fill: bad argument");
    FGK_CUDA(cudaSetDevice(h->device));
    if (h->itab_bytes) {
        int rc = conn_smem_attr<true>(h->itab_bytes);
        if (rc != FGK_OK) return rc;
        k_conn<true, true><<<grid_for(n, 4 * FGK_WARPS_PER_BLOCK, h->device, conn_ctas_per_sm(h->itab_bytes)), FGK_BLOCK, h->itab_bytes, (cudaStream_t)stream>>>(
            h->v, h->itab_bytes, (const fgk_det*)dets, n, nullptr, (const i64*)offsets, (fgk_det*)out_dets, out_elems,
            (i64*)out_src);
    } else
    k_conn<true, false><<<grid_for(n, FGK_WARPS_PER_BLOCK, h->device, 8), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        h->v, 0u, (const fgk_det*)dets, n, nullptr, (const i64*)offsets, (fgk_det*)out_dets, out_elems,
        (i64*)out_src);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}
