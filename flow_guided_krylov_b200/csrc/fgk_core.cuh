// fgk_core.cuh -- bit-level determinant algebra shared by every kernel.
//
// Everything here is __host__ __device__ so that the exact same arithmetic the
// kernels run can be compiled by g++ into a CPU self-check (csrc/fgk_hostcheck.cpp,
// used by the "not gpu" tests) and compared with the oracle without a GPU.
//
// Conventions (DESIGN.md "Data layout"):
//   * a determinant is two 64-bit occupation words {alpha, beta};
//   * orbital p of a spin block lives in bit (n_orb-1-p), so that the pair
//     (alpha, beta) compared as a 128-bit unsigned number orders determinants
//     exactly like the reference's site-0-is-MSB integer key
//     (reference molecular.py:498-500, torch.unique order);
//   * "site" = orbital for alpha, n_orb + orbital for beta (molecular.py:43-45).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FGK_HD __host__ __device__ __forceinline__
#else
#define FGK_HD inline
#endif

typedef unsigned long long u64;

struct fgk_det { u64 a, b; };   // 16 B, loaded/stored as one 128-bit word

// ---- device-side view of a Hamiltonian handle -------------------------------
struct HamView {
    int n_orb, n_alpha, n_beta;
    double e_nuc;
    const float* h1;      // (n,n)  float32 h1[p*n+q]                molecular.py:68
    const float* g;       // (n,n,n,n) float32 (pq|rs) chemist order  molecular.py:69
    const float* w;       // w[x,y,z,u] = fp32(g[x,y,z,u] - g[x,u,z,y]) same-spin value (molecular.py:265,287)
    const double* hdiag;  // h_pp
    const double* jks;    // 0.5*(J_pq+J_qp) - 0.5*(K_pq+K_qp)   (molecular.py:163-182, p!=q)
    const double* jab;    // J_pq = g[p,p,q,q]                   (molecular.py:171)
    // nibble row-sum tables (may be null): nib_x[(p*nchunk + c)*16 + v] = sum of x[p][q] over the
    // orbitals q whose bits are set in nibble value v of 4-bit chunk c of an occupation word
    const double* nib_jk;
    const double* nib_jab;
    int nchunk;           // ceil(n_orb / 4)
};

FGK_HD int fgk_popc(u64 x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

FGK_HD int fgk_clz(u64 x)   // x != 0
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// bit of orbital p inside a spin word
FGK_HD u64 orb_bit(int n_orb, int p) { return 1ull << (n_orb - 1 - p); }

// mask of the orbitals with index < p (they sit in the bits above orbital p's bit)
FGK_HD u64 below_mask(int n_orb, int p) { return (~0ull << (n_orb - 1 - p)) << 1; }

// k-th (0-based, ascending orbital index) set orbital of word x; x has > k bits set
FGK_HD int nth_orbital(u64 x, int n_orb, int k)
{
    for (int i = 0; i < k; i++) x &= ~(1ull << (63 - fgk_clz(x)));
    return n_orb - 1 - (63 - fgk_clz(x));
}

// Jordan-Wigner sign of a+_p a_q inside one spin block (molecular.py:379-389):
// parity of the occupied orbitals strictly between p and q.  Orbitals of the
// other spin block never lie between two same-spin sites.
FGK_HD int sign1_parity(u64 word, int n_orb, int p, int q)
{
    int lo = p < q ? p : q, hi = p < q ? q : p;
    u64 between = below_mask(n_orb, hi) & ~below_mask(n_orb, lo) & ~orb_bit(n_orb, lo);
    return fgk_popc(word & between) & 1;
}

// The reference's double-excitation sign (molecular.py:391-423), evaluated on
// the KET {a,b} for a+_p a+_r a_s a_q given as SITE indices (0..2n-1).
// total = P(p) + P(r) + P(s) + P(q) + [p<s] + [r<s] + [p<q] + [r<q]
//         - [q<r] c_q - [q<s] c_q - [s<q] c_s,   P(x) = #occupied sites < x.
// Returns total mod 2 (1 => sign -1).  No assumption on the order of p,r,q,s.
FGK_HD int sign2_parity(u64 a, u64 b, int n_orb, int p, int r, int q, int s)
{
    u64 ma = 0, mb = 0;
    int nbeta = 0;
    const int st[4] = {p, r, s, q};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) {
        int x = st[i];
        if (x < n_orb) ma ^= below_mask(n_orb, x);
        else { mb ^= below_mask(n_orb, x - n_orb); nbeta++; }
    }
    int t = fgk_popc(a & ma) + fgk_popc(b & mb) + nbeta * fgk_popc(a);
    int cq = (q < n_orb) ? (int)((a >> (n_orb - 1 - q)) & 1) : (int)((b >> (2 * n_orb - 1 - q)) & 1);
    int cs = (s < n_orb) ? (int)((a >> (n_orb - 1 - s)) & 1) : (int)((b >> (2 * n_orb - 1 - s)) & 1);
    t += (p < s) + (r < s) + (p < q) + (r < q);
    t -= (q < r) ? cq : 0;
    t -= (q < s) ? cq : 0;
    t -= (s < q) ? cs : 0;
    return t & 1;
}

// index into an (n,n,n,n) table; n <= 64, so the index (< 2^24) fits 32-bit arithmetic
FGK_HD size_t idx4(int n, int x, int y, int z, int u)
{
    return (size_t)(unsigned)(((x * n + y) * n + z) * n + u);
}

// decode t in [0, m(m-1)/2) -> (k<l), row-major over k (the i<j / k<l loops of
// molecular.py:257-264).
FGK_HD void tri_decode(int m, int t, int& k, int& l)
{
    // offset(k) = k*(2m-k-1)/2
    float fm = (float)(2 * m - 1);
#if defined(__CUDA_ARCH__)
    int kk = (int)((fm - sqrtf(fm * fm - 8.0f * (float)t)) * 0.5f);
#else
    int kk = (int)((fm - __builtin_sqrtf(fm * fm - 8.0f * (float)t)) * 0.5f);
#endif
    if (kk < 0) kk = 0;
    if (kk > m - 2) kk = m - 2;
    while (kk > 0 && kk * (2 * m - kk - 1) / 2 > t) kk--;
    while ((kk + 1) * (2 * m - kk - 2) / 2 <= t) kk++;
    k = kk;
    l = t - kk * (2 * m - kk - 1) / 2 + kk + 1;
}

// 64-bit mix of a determinant (splitmix/murmur finaliser); upper 32 bits are the
// tag stored next to the index in the basis table, lower bits pick the slot.
FGK_HD u64 det_hash(u64 a, u64 b)
{
    u64 h = a * 0x9E3779B97F4A7C15ull;
    h ^= (b + 0xD6E8FEB86659FD93ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    h *= 0x9FB21C651E98DF25ull;
    h ^= h >> 32;
    return h;
}

FGK_HD u64 word_hash(u64 a)
{
    u64 h = (a + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    return h;
}

// ---- one excitation of a determinant D --------------------------------------
// cls: 0 = alpha single, 1 = beta single, 2 = alpha-alpha, 3 = beta-beta, 4 = alpha-beta
// (h0,h1) orbitals emptied in D, (e0,e1) orbitals filled (h1/e1 unused for
// singles; for cls 4, h0/e0 are alpha and h1/e1 beta orbitals).  For same-spin
// doubles h0<h1 and e0<e1.
struct Excitation { int cls, h0, h1, e0, e1; };

FGK_HD fgk_det apply_excitation(fgk_det d, int n, const Excitation& x)
{
    switch (x.cls) {
    case 0: d.a ^= orb_bit(n, x.h0) | orb_bit(n, x.e0); break;
    case 1: d.b ^= orb_bit(n, x.h0) | orb_bit(n, x.e0); break;
    case 2: d.a ^= orb_bit(n, x.h0) | orb_bit(n, x.h1) | orb_bit(n, x.e0) | orb_bit(n, x.e1); break;
    case 3: d.b ^= orb_bit(n, x.h0) | orb_bit(n, x.h1) | orb_bit(n, x.e0) | orb_bit(n, x.e1); break;
    default:
        d.a ^= orb_bit(n, x.h0) | orb_bit(n, x.e0);
        d.b ^= orb_bit(n, x.h1) | orb_bit(n, x.e1);
    }
    return d;
}

// KET element: the value the reference's get_connections(ket) attaches to the
// connection ket -> (ket with x applied), molecular.py:234-318.
// Returns false if the reference's |val| > 1e-12 filter drops it.
// `ldf` abstracts the table read (plain load on the host, __ldg on the device).
template <class Ld>
FGK_HD bool ket_element(const HamView& H, fgk_det ket, const Excitation& x, Ld ldf, float& out)
{
    const int n = H.n_orb;
    float val;
    int par;
    switch (x.cls) {
    case 0:   // a+_p a_q, p = e0, q = h0 : sign * h_pq          (:234-242)
    case 1: { //                                                  (:244-251)
        val = ldf(H.h1 + (size_t)x.e0 * n + x.h0);
        par = sign1_parity(x.cls == 0 ? ket.a : ket.b, n, x.e0, x.h0);
        break;
    }
    case 2: { // q=h0 < s=h1, p=e0 < r=e1 : g[p,q,r,s]-g[p,s,r,q] (:254-274)
        val = ldf(H.w + idx4(n, x.e0, x.h0, x.e1, x.h1));
        par = sign2_parity(ket.a, ket.b, n, x.e0, x.e1, x.h0, x.h1);
        break;
    }
    case 3: { //                                                  (:276-300)
        val = ldf(H.w + idx4(n, x.e0, x.h0, x.e1, x.h1));
        par = sign2_parity(ket.a, ket.b, n, x.e0 + n, x.e1 + n, x.h0 + n, x.h1 + n);
        break;
    }
    default: { // q=h0 (alpha), s=h1 (beta), p=e0, r=e1 : g[p,q,r,s] (:302-318)
        val = ldf(H.g + idx4(n, x.e0, x.h0, x.e1, x.h1));
        par = sign2_parity(ket.a, ket.b, n, x.e0, x.e1 + n, x.h0, x.h1 + n);
    }
    }
    float av = val < 0.f ? -val : val;
    if (!(av > 1e-12f)) return false;
    out = par ? -val : val;
    return true;
}

// BRA element: <D|H|j> with j = D + x, i.e. what get_connections(j) reports for
// the connection j -> D.  Seen from j the holes are x's particles and vice versa.
template <class Ld>
FGK_HD bool bra_element(const HamView& H, fgk_det bra, const Excitation& x, Ld ldf, float& out)
{
    fgk_det ket = apply_excitation(bra, H.n_orb, x);
    Excitation rx;
    rx.cls = x.cls; rx.h0 = x.e0; rx.h1 = x.e1; rx.e0 = x.h0; rx.e1 = x.h1;
    return ket_element(H, ket, rx, ldf, out);
}

// FP64 diagonal <D|H|D> on the float32-rounded tables (molecular.py:133-184 in
// bit form).  `ldd` abstracts the table read.  Generic pair-loop form.
template <class Ldd>
FGK_HD double diag_element_loops(const HamView& H, fgk_det d, Ldd ldd)
{
    const int n = H.n_orb;
    double e = H.e_nuc;
    u64 xa = d.a;
    while (xa) {
        int bp = 63 - fgk_clz(xa);
        xa &= ~(1ull << bp);
        int p = n - 1 - bp;
        e += ldd(H.hdiag + p);
        u64 ya = xa;                      // orbitals q > p of the same spin
        while (ya) {
            int bq = 63 - fgk_clz(ya);
            ya &= ~(1ull << bq);
            e += ldd(H.jks + (size_t)p * n + (n - 1 - bq));
        }
        u64 yb = d.b;                     // every beta orbital (p == q included, :171)
        while (yb) {
            int bq = 63 - fgk_clz(yb);
            yb &= ~(1ull << bq);
            e += ldd(H.jab + (size_t)p * n + (n - 1 - bq));
        }
    }
    u64 xb = d.b;
    while (xb) {
        int bp = 63 - fgk_clz(xb);
        xb &= ~(1ull << bp);
        int p = n - 1 - bp;
        e += ldd(H.hdiag + p);
        u64 yb = xb;
        while (yb) {
            int bq = 63 - fgk_clz(yb);
            yb &= ~(1ull << bq);
            e += ldd(H.jks + (size_t)p * n + (n - 1 - bq));
        }
    }
    return e;
}

// Same quantity from nibble row-sum tables: the (J-K) matrix is symmetric with a zero diagonal
// (J_pp = K_pp), so  sum_{p<q in S} jks[p][q] = 1/2 sum_{p in S} rowsum_p(S), and a row sum over
// an occupation word is nchunk table reads (one per 4-bit nibble) instead of a bit loop.
template <class Ldd>
FGK_HD double nib_rowsum(const double* T, int nchunk, u64 w, Ldd ldd)
{
    double s = 0.0;
    for (int c = 0; c < nchunk; c++) s += ldd(T + c * 16 + (int)((w >> (4 * c)) & 15ull));
    return s;
}

template <class Ldd>
FGK_HD double diag_element(const HamView& H, fgk_det d, Ldd ldd)
{
    if (!H.nib_jk) return diag_element_loops(H, d, ldd);
    const int n = H.n_orb, nc = H.nchunk;
    double e = H.e_nuc, same = 0.0;
    u64 xa = d.a;
    while (xa) {
        int bp = 63 - fgk_clz(xa);
        xa &= ~(1ull << bp);
        int p = n - 1 - bp;
        e += ldd(H.hdiag + p);
        same += nib_rowsum(H.nib_jk + (size_t)p * nc * 16, nc, d.a, ldd);
        e += nib_rowsum(H.nib_jab + (size_t)p * nc * 16, nc, d.b, ldd);
    }
    u64 xb = d.b;
    while (xb) {
        int bp = 63 - fgk_clz(xb);
        xb &= ~(1ull << bp);
        int p = n - 1 - bp;
        e += ldd(H.hdiag + p);
        same += nib_rowsum(H.nib_jk + (size_t)p * nc * 16, nc, d.b, ldd);
    }
    return e + 0.5 * same;
}

// ---- per-determinant enumeration context -------------------------------------
// Orbital lists in ascending orbital index (np.where order, molecular.py:220-223).
// On the device they live in shared memory, one set per warp.
struct DetCtx {
    int n;                 // n_orb
    fgk_det d;
    const uint8_t *occ_a, *virt_a, *occ_b, *virt_b;
    int noa, nva, nob, nvb;
    int n_s;               // singles index space  : n*n   (p = t / n, q = t % n)
    int n_aa, n_bb, n_ab;  // doubles index spaces : C(noa,2)C(nva,2), C(nob,2)C(nvb,2), noa*nob*nva*nvb
};

FGK_HD void detctx_sizes(DetCtx& c)
{
    c.n_s = c.n * c.n;
    c.n_aa = (c.noa * (c.noa - 1) / 2) * (c.nva * (c.nva - 1) / 2);
    c.n_bb = (c.nob * (c.nob - 1) / 2) * (c.nvb * (c.nvb - 1) / 2);
    c.n_ab = c.noa * c.nob * c.nva * c.nvb;
}

// host-side list construction (the kernels build the same lists with one lane per orbital)
inline void detctx_fill_host(DetCtx& c, int n, fgk_det d, uint8_t* buf /* 4*64 bytes */)
{
    uint8_t *oa = buf, *va = buf + 64, *ob = buf + 128, *vb = buf + 192;
    c.n = n; c.d = d; c.noa = c.nva = c.nob = c.nvb = 0;
    for (int p = 0; p < n; p++) {
        if (d.a & orb_bit(n, p)) oa[c.noa++] = (uint8_t)p; else va[c.nva++] = (uint8_t)p;
        if (d.b & orb_bit(n, p)) ob[c.nob++] = (uint8_t)p; else vb[c.nvb++] = (uint8_t)p;
    }
    c.occ_a = oa; c.virt_a = va; c.occ_b = ob; c.virt_b = vb;
    detctx_sizes(c);
}

// singles index t -> (p = t / n particle, q = t % n hole); valid_a / valid_b say
// whether the alpha / beta move exists in D (occupancy only; the |h_pq| filter is
// applied by ket_element/bra_element).  Reference order: p outer, q inner, alpha
// before beta for each pair (molecular.py:234-251).
FGK_HD void decode_single(const DetCtx& c, int t, int& p, int& q, bool& valid_a, bool& valid_b)
{
    p = t / c.n;
    q = t - p * c.n;
    u64 bp = orb_bit(c.n, p), bq = orb_bit(c.n, q);
    valid_a = (p != q) && (c.d.a & bq) && !(c.d.a & bp);
    valid_b = (p != q) && (c.d.b & bq) && !(c.d.b & bp);
}

// doubles: stage 2 = alpha-alpha, 3 = beta-beta, 4 = alpha-beta; t inside the
// stage's index space, reference loop order (molecular.py:257-264, 279-286, 303-306).
FGK_HD void decode_double(const DetCtx& c, int stage, int t, Excitation& x)
{
    x.cls = stage;
    if (stage == 4) {
        int l = t % c.nvb; t /= c.nvb;
        int k = t % c.nva; t /= c.nva;
        int j = t % c.nob; t /= c.nob;
        x.h0 = c.occ_a[t]; x.h1 = c.occ_b[j]; x.e0 = c.virt_a[k]; x.e1 = c.virt_b[l];
        return;
    }
    const uint8_t* occ = stage == 2 ? c.occ_a : c.occ_b;
    const uint8_t* virt = stage == 2 ? c.virt_a : c.virt_b;
    int no = stage == 2 ? c.noa : c.nob, nv = stage == 2 ? c.nva : c.nvb;
    int npp = nv * (nv - 1) / 2;
    int hp = t / npp, pp = t - hp * npp;
    int i, j, k, l;
    tri_decode(no, hp, i, j);
    tri_decode(nv, pp, k, l);
    x.h0 = occ[i]; x.h1 = occ[j]; x.e0 = virt[k]; x.e1 = virt[l];
}

// ---- excitation recovered from two strings of the same spin ------------------------
// x = from ^ to.  popc(x) == 2: single (hole = the bit set in `from`, particle = the bit
// set in `to`); popc(x) == 4: double, holes h0 < h1 and particles e0 < e1 in ORBITAL
// order (orbital p lives in bit n-1-p, so the higher bit is the lower orbital).
FGK_HD int fgk_ctz(u64 x)
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

FGK_HD void single_from_strings(u64 from, u64 to, int n, int& h, int& e)
{
    u64 x = from ^ to;
    h = n - 1 - (63 - fgk_clz(x & from));
    e = n - 1 - (63 - fgk_clz(x & to));
}

FGK_HD void double_from_strings(u64 from, u64 to, int n, int& h0, int& h1, int& e0, int& e1)
{
    u64 x = from ^ to, hs = x & from, es = x & to;
    h0 = n - 1 - (63 - fgk_clz(hs));
    h1 = n - 1 - fgk_ctz(hs);
    e0 = n - 1 - (63 - fgk_clz(es));
    e1 = n - 1 - fgk_ctz(es);
}

// ---- split element evaluation: table value and sign parity separately ------------------------
// (the count pass of the projected-H builder only needs |value| > 1e-12; the fill pass adds
// the sign).  The fast parities use c_q = c_s = 1 (the holes ARE occupied in the ket) and the
// fixed site order of each class; fgk_hostcheck verifies them against sign2_parity.
template <class Ld>
FGK_HD float exc_value_ket(const HamView& H, const Excitation& x, Ld ldf)
{
    const int n = H.n_orb;
    if (x.cls < 2) return ldf(H.h1 + (size_t)x.e0 * n + x.h0);
    return ldf((x.cls == 4 ? H.g : H.w) + idx4(n, x.e0, x.h0, x.e1, x.h1));
}

template <class Ld>
FGK_HD float exc_value_bra(const HamView& H, const Excitation& x, Ld ldf)
{
    const int n = H.n_orb;
    if (x.cls < 2) return ldf(H.h1 + (size_t)x.h0 * n + x.e0);
    return ldf((x.cls == 4 ? H.g : H.w) + idx4(n, x.h0, x.e0, x.h1, x.e1));
}

// orbitals o with min(x,y) <= o < max(x,y)
FGK_HD u64 span_mask(int n, int x, int y) { return below_mask(n, x) ^ below_mask(n, y); }

// parity of the reference sign for the connection ket -> ket + x (x's holes q=h0,s=h1 occupied
// and particles p=e0,r=e1 empty in `ket`)
FGK_HD int exc_parity_ket(fgk_det ket, int n, const Excitation& x)
{
    const int q = x.h0, s = x.h1, p = x.e0, r = x.e1;
    switch (x.cls) {
    case 0: return sign1_parity(ket.a, n, p, q);
    case 1: return sign1_parity(ket.b, n, p, q);
    case 2:
    case 3: {
        const u64 w = x.cls == 2 ? ket.a : ket.b;
        int t = fgk_popc(w & (span_mask(n, p, q) ^ span_mask(n, r, s)));
        t += (p < s) + (r < s) + (p < q) + (r < q) + (q < r) + 1;     // -[q<r] - 1  ==  +[q<r] + 1 (mod 2)
        return t & 1;
    }
    default: {
        int t = fgk_popc(ket.a & span_mask(n, p, q)) + fgk_popc(ket.b & span_mask(n, r, s));
        t += (r < s) + (p < q) + 1;
        return t & 1;
    }
    }
}

// ---- separable form of the alpha-beta double excitation -------------------------------------
// For a+_p a+_r a_s a_q with (q -> p) in the alpha and (s -> r) in the beta block, both the
// table offset and the reference's sign parity split into one factor per single excitation:
//   g[p,q,r,s]          = g[(p n + q) n^2 + (r n + s)]
//   parity (ket side)   = pk_alpha ^ pk_beta ^ 1,  pk = popc(word & span(e,h)) + [e < h]
//   parity (bra side)   = pb_alpha ^ pb_beta ^ 1,  pb = popc(word' & span(h,e)) + [h < e]
// (word' = word with the single applied; exc_parity_ket, class 4).  sk / sb are the single's
// own parities on the ket / bra side (class 0 / 1).  Used by k_projh3 and k_pt2_accumulate2,
// verified against the generic forms by fgk_hostcheck (hc_bra_row3, hc_pt2_walk2).
FGK_HD void single_factors(u64 w, u64 w2, int n, int hh, int ee, unsigned& pk, unsigned& pb,
                           unsigned& sk, unsigned& sb)
{
    pk = (unsigned)(fgk_popc(w & span_mask(n, ee, hh)) + (ee < hh)) & 1u;
    pb = (unsigned)(fgk_popc(w2 & span_mask(n, hh, ee)) + (hh < ee)) & 1u;
    sk = (unsigned)sign1_parity(w, n, ee, hh);
    sb = (unsigned)sign1_parity(w2, n, hh, ee);
}

// <D|H|D+x> as get_connections(D+x) reports it: roles reversed on the ket D+x
FGK_HD int exc_parity_bra(fgk_det bra, int n, const Excitation& x)
{
    fgk_det ket = apply_excitation(bra, n, x);
    Excitation rx;
    rx.cls = x.cls; rx.h0 = x.e0; rx.h1 = x.e1; rx.e0 = x.h0; rx.e1 = x.h1;
    return exc_parity_ket(ket, n, rx);
}

// ket_element / bra_element with the split, class-specialised evaluation (bit-identical
// results; verified against the generic forms by fgk_hostcheck's hc_check_split)
template <class Ld>
FGK_HD bool ket_element_fast(const HamView& H, fgk_det ket, const Excitation& x, Ld ldf, float& out)
{
    float v = exc_value_ket(H, x, ldf);
    float av = v < 0.f ? -v : v;
    if (!(av > 1e-12f)) return false;
    out = exc_parity_ket(ket, H.n_orb, x) ? -v : v;
    return true;
}

FGK_HD u64 fgk_double_bits(double v)
{
#if defined(__CUDA_ARCH__)
    return (u64)__double_as_longlong(v);
#else
    u64 b; __builtin_memcpy(&b, &v, 8); return b;
#endif
}

// v * 2^e for a normal result (exact): builds the power of two from its exponent field
FGK_HD double fgk_scale2(double v, int e)
{
    // split so that both factors are normal doubles (|e| <= 200 here)
    const u64 b1 = (u64)(1023 + e / 2) << 52, b2 = (u64)(1023 + (e - e / 2)) << 52;
#if defined(__CUDA_ARCH__)
    return v * __longlong_as_double((long long)b1) * __longlong_as_double((long long)b2);
#else
    double p1, p2; __builtin_memcpy(&p1, &b1, 8); __builtin_memcpy(&p2, &b2, 8);
    return v * p1 * p2;
#endif
}

// ---- exact accumulation --------------------------------------------------------------------
// A coupling sum  sum_j c_j <x|H|j>  is accumulated as a signed 128-bit FIXED-POINT integer
// (resolution 2^-70, |addend| < 2^30, 27 bits of headroom for the number of addends): integer
// addition is associative, so the sum does not depend on the order in which the warps arrive --
// the sweep is bit-reproducible from run to run and for any number of passes / owner ranks
// (FP64 atomicAdd in arrival order was not).  Each addend is rounded to 2^-70 once (8.5e-22,
// far below the 1e-12 parity tolerance of the couplings); the total is converted to FP64 with
// one correctly rounded step when the candidates are scored or exported.
static const int PT2_FX_SHIFT = 70;

FGK_HD bool fx_from_double(double v, u64& lo, u64& hi)
{
    const u64 bits = fgk_double_bits(v);
    const int ex = (int)((bits >> 52) & 0x7ff);
    lo = hi = 0;
    if (ex == 0) return true;                           // zero / subnormal
    const u64 mant = (bits & 0xfffffffffffffull) | (1ull << 52);
    const int sh = ex - 1075 + PT2_FX_SHIFT;            // v = mant * 2^(ex - 1075)
    if (sh > 47) return false;                          // |v| >= 2^30 (or inf / nan): out of range
    if (sh >= 0) {
        lo = mant << sh;
        hi = sh ? mant >> (64 - sh) : 0;
    } else if (sh > -54) {
        lo = (mant + (1ull << (-sh - 1))) >> (-sh);     // round half up on the magnitude
    }
    if (bits >> 63) {                                   // two's complement negate
        lo = ~lo + 1;
        hi = ~hi + (lo == 0 ? 1 : 0);
    }
    return true;
}

FGK_HD double fx_to_double(u64 lo, u64 hi)
{
    const bool neg = (hi >> 63) != 0;
    if (neg) { lo = ~lo + 1; hi = ~hi + (lo == 0 ? 1 : 0); }
    double r;
    if (hi == 0) {
        r = fgk_scale2((double)lo, -PT2_FX_SHIFT);
    } else {
        const int lz = fgk_clz(hi);
        u64 top = lz ? (hi << lz) | (lo >> (64 - lz)) : hi;
        const u64 rest = lz ? (lo << lz) : lo;
        if (rest) top |= 1ull;                          // sticky bit: one correct rounding
        r = fgk_scale2((double)top, 64 - lz - PT2_FX_SHIFT);
    }
    return neg ? -r : r;
}

