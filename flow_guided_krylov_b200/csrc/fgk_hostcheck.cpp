// fgk_hostcheck.cpp -- CPU self-check of the kernels' index arithmetic.
//
// NOT a fallback and never imported by the product path: it exists so that the
// "not gpu" test tier can run the exact __host__ __device__ functions the CUDA
// kernels use (fgk_core.cuh: decode, sign, element, diagonal) against the
// oracle on a machine without a GPU.  Built by tests with plain g++.
#include <map>
#include <utility>
#include <vector>
#include "fgk_core.cuh"
#include "fgk_lists.cuh"
#include "fgk_tables.h"

struct HcHam { HostTables T; HamView V; };

static inline float ldf_host(const float* p) { return *p; }
static inline double ldd_host(const double* p) { return *p; }

extern "C" {

void* hc_ham_create(const double* h1, const double* g, int n_orb, int n_alpha, int n_beta, double e_nuc)
{
    HcHam* H = new HcHam();
    build_host_tables(h1, g, n_orb, H->T);
    H->V.n_orb = n_orb; H->V.n_alpha = n_alpha; H->V.n_beta = n_beta; H->V.e_nuc = e_nuc;
    H->V.h1 = H->T.h1.data(); H->V.g = H->T.g.data(); H->V.w = H->T.w.data();
    H->V.hdiag = H->T.hdiag.data(); H->V.jks = H->T.jks.data(); H->V.jab = H->T.jab.data();
    H->V.nib_jk = H->T.nib_jk.data(); H->V.nib_jab = H->T.nib_jab.data(); H->V.nchunk = H->T.nchunk;
    return H;
}
void hc_ham_destroy(void* h) { delete (HcHam*)h; }

void hc_diag(void* h, const u64* dets, long n, double* out)
{
    HcHam* H = (HcHam*)h;
    for (long i = 0; i < n; i++) {
        fgk_det d = {dets[2 * i], dets[2 * i + 1]};
        out[i] = diag_element(H->V, d, ldd_host);
    }
}

// the generic pair-loop diagonal (cross-check of the nibble-table form hc_diag uses)
void hc_diag_loops(void* h, const u64* dets, long n, double* out)
{
    HcHam* H = (HcHam*)h;
    for (long i = 0; i < n; i++) {
        fgk_det d = {dets[2 * i], dets[2 * i + 1]};
        out[i] = diag_element_loops(H->V, d, ldd_host);
    }
}

// ket-mode enumeration in reference order; returns count, stores up to cap
long hc_connections(void* h, u64 a, u64 b, u64* out_dets, float* out_el, long cap)
{
    HcHam* H = (HcHam*)h;
    uint8_t buf[256];
    DetCtx c;
    fgk_det d = {a, b};
    detctx_fill_host(c, H->V.n_orb, d, buf);
    long m = 0;
    auto emit = [&](const Excitation& x) {
        float v;
        if (!ket_element(H->V, d, x, ldf_host, v)) return;
        if (m < cap) {
            fgk_det o = apply_excitation(d, c.n, x);
            out_dets[2 * m] = o.a; out_dets[2 * m + 1] = o.b; out_el[m] = v;
        }
        m++;
    };
    for (int t = 0; t < c.n_s; t++) {
        int p, q; bool va, vb;
        decode_single(c, t, p, q, va, vb);
        Excitation x; x.h1 = x.e1 = 0; x.h0 = q; x.e0 = p;
        if (va) { x.cls = 0; emit(x); }
        if (vb) { x.cls = 1; emit(x); }
    }
    const int sizes[3] = {c.n_aa, c.n_bb, c.n_ab};
    for (int st = 2; st <= 4; st++)
        for (int t = 0; t < sizes[st - 2]; t++) {
            Excitation x;
            decode_double(c, st, t, x);
            emit(x);
        }
    return m;
}

// bra-mode row i of the projected H over `basis` (n dets): every (col j, value)
// with j = basis index of D_i + x.  mode 0 = raw directed <i|H|j>;
// mode 1 = symmetrised 0.5*(<i|H|j> + <j|H|i>).  Diagonal included first.
long hc_bra_row(void* h, const u64* basis, long n, long i, int mode, int* out_cols,
                double* out_vals, long cap)
{
    HcHam* H = (HcHam*)h;
    std::map<std::pair<u64, u64>, long> index;
    for (long k = 0; k < n; k++) index[{basis[2 * k], basis[2 * k + 1]}] = k;   // last wins
    uint8_t buf[256];
    DetCtx c;
    fgk_det d = {basis[2 * i], basis[2 * i + 1]};
    detctx_fill_host(c, H->V.n_orb, d, buf);
    long m = 0;
    if (m < cap) { out_cols[m] = (int)i; out_vals[m] = diag_element(H->V, d, ldd_host); }
    m++;
    auto emit = [&](const Excitation& x) {
        fgk_det o = apply_excitation(d, c.n, x);
        auto it = index.find({o.a, o.b});
        if (it == index.end()) return;
        float vij = 0.f, vji = 0.f;
        bool kij = bra_element(H->V, d, x, ldf_host, vij);
        bool kji = (mode == 1) ? ket_element(H->V, d, x, ldf_host, vji) : false;
        if (!kij && !kji) return;
        double v = mode == 1 ? 0.5 * ((double)(kij ? vij : 0.f) + (double)(kji ? vji : 0.f))
                             : (double)vij;
        if (m < cap) { out_cols[m] = (int)it->second; out_vals[m] = v; }
        m++;
    };
    for (int t = 0; t < c.n_s; t++) {
        int p, q; bool va, vb;
        decode_single(c, t, p, q, va, vb);
        Excitation x; x.h1 = x.e1 = 0; x.h0 = q; x.e0 = p;
        if (va) { x.cls = 0; emit(x); }
        if (vb) { x.cls = 1; emit(x); }
    }
    const int sizes[3] = {c.n_aa, c.n_bb, c.n_ab};
    for (int st = 2; st <= 4; st++)
        for (int t = 0; t < sizes[st - 2]; t++) {
            Excitation x;
            decode_double(c, st, t, x);
            emit(x);
        }
    return m;
}

// Structured bra-mode row (the k_projh2 strategy): string-set filtered singles lists,
// doubles found either by enumeration + set test (scan = 0) or by scanning the list of
// distinct strings (scan = 1), alpha-beta doubles as the product of the two singles lists.
long hc_bra_row2(void* h, const u64* basis, long n, long i, int mode, int scan, int* out_cols,
                 double* out_vals, long cap)
{
    HcHam* H = (HcHam*)h;
    std::map<std::pair<u64, u64>, long> index;
    std::map<u64, int> aset, bset;
    for (long k = 0; k < n; k++) {
        index[{basis[2 * k], basis[2 * k + 1]}] = k;
        aset[basis[2 * k]] = 1; bset[basis[2 * k + 1]] = 1;
    }
    uint8_t buf[256];
    DetCtx c;
    fgk_det d = {basis[2 * i], basis[2 * i + 1]};
    const int nn = H->V.n_orb;
    detctx_fill_host(c, nn, d, buf);
    long m = 0;
    if (m < cap) { out_cols[m] = (int)i; out_vals[m] = diag_element(H->V, d, ldd_host); }
    m++;
    auto emit = [&](const Excitation& x) {
        fgk_det o = apply_excitation(d, c.n, x);
        auto it = index.find({o.a, o.b});
        if (it == index.end()) return;
        float vij = 0.f, vji = 0.f;
        bool kij = bra_element(H->V, d, x, ldf_host, vij);
        bool kji = (mode == 1) ? ket_element(H->V, d, x, ldf_host, vji) : false;
        if (!kij && !kji) return;
        double v = mode == 1 ? 0.5 * ((double)(kij ? vij : 0.f) + (double)(kji ? vji : 0.f))
                             : (double)vij;
        if (m < cap) { out_cols[m] = (int)it->second; out_vals[m] = v; }
        m++;
    };
    std::vector<std::pair<int, int>> L[2];
    for (int spin = 0; spin < 2; spin++) {
        u64 w = spin ? d.b : d.a;
        auto& set = spin ? bset : aset;
        const uint8_t* occ = spin ? c.occ_b : c.occ_a;
        const uint8_t* virt = spin ? c.virt_b : c.virt_a;
        int no = spin ? c.nob : c.noa, nv = spin ? c.nvb : c.nva;
        if (scan) {
            for (auto& kv : set) {
                u64 x = kv.first ^ w;
                int pc = fgk_popc(x);
                if (pc == 2) {
                    int hh, ee;
                    single_from_strings(w, kv.first, nn, hh, ee);
                    L[spin].push_back({hh, ee});
                } else if (pc == 4) {
                    Excitation e; e.cls = 2 + spin;
                    double_from_strings(w, kv.first, nn, e.h0, e.h1, e.e0, e.e1);
                    emit(e);
                }
            }
        } else {
            for (int t = 0; t < no * nv; t++) {
                int hh = occ[t / nv], ee = virt[t % nv];
                u64 w2 = w ^ orb_bit(nn, hh) ^ orb_bit(nn, ee);
                if (set.count(w2)) L[spin].push_back({hh, ee});
            }
            int size = spin ? c.n_bb : c.n_aa;
            for (int t = 0; t < size; t++) {
                Excitation e;
                decode_double(c, 2 + spin, t, e);
                fgk_det o = apply_excitation(d, nn, e);
                if (set.count(spin ? o.b : o.a)) emit(e);
            }
        }
    }
    for (int spin = 0; spin < 2; spin++)
        for (auto& he : L[spin]) {
            Excitation e; e.cls = spin; e.h0 = he.first; e.e0 = he.second; e.h1 = e.e1 = 0;
            emit(e);
        }
    for (auto& a : L[0])
        for (auto& b : L[1]) {
            Excitation e; e.cls = 4; e.h0 = a.first; e.e0 = a.second; e.h1 = b.first; e.e1 = b.second;
            emit(e);
        }
    return m;
}

// Rank-based bra-mode row (the k_projh3 strategy): sorted distinct-string lists are scanned,
// singles become entries {offk, offb, parity factors, rank}, the column comes from the
// (alpha rank, beta rank) pair map, alpha-beta values and signs from the separable form.
long hc_bra_row3(void* h, const u64* basis, long n, long i, int mode, int* out_cols, double* out_vals,
                 long cap)
{
    HcHam* H = (HcHam*)h;
    const HamView& V = H->V;
    const int nn = V.n_orb, n2 = nn * nn;
    std::map<u64, int> arank, brank;
    for (long k = 0; k < n; k++) { arank[basis[2 * k]] = 0; brank[basis[2 * k + 1]] = 0; }
    std::vector<u64> alist, blist;
    for (auto& kv : arank) { kv.second = (int)alist.size(); alist.push_back(kv.first); }
    for (auto& kv : brank) { kv.second = (int)blist.size(); blist.push_back(kv.first); }
    std::map<std::pair<int, int>, long> pair;
    for (long k = 0; k < n; k++) pair[{arank[basis[2 * k]], brank[basis[2 * k + 1]]}] = k;   // last index wins
    fgk_det d = {basis[2 * i], basis[2 * i + 1]};
    const int ia = arank[d.a], ib = brank[d.b];
    const bool sym = mode == 1;
    long m = 0;
    if (m < cap) { out_cols[m] = (int)i; out_vals[m] = diag_element(V, d, ldd_host); }
    m++;
    auto column = [&](int ra, int rb) -> long {
        auto it = pair.find({ra, rb});
        return it == pair.end() ? -1 : it->second;
    };
    auto emit = [&](long j, float rb_, float rk, unsigned parb, unsigned park) {
        if (j < 0) return;
        const bool kij = (rb_ < 0 ? -rb_ : rb_) > 1e-12f;
        const bool kji = sym && (rk < 0 ? -rk : rk) > 1e-12f;
        if (!kij && !kji) return;
        const float vij = kij ? (parb ? -rb_ : rb_) : 0.f;
        const float vji = kji ? (park ? -rk : rk) : 0.f;
        const double v = sym ? 0.5 * ((double)vij + (double)vji) : (double)vij;
        if (m < cap) { out_cols[m] = (int)j; out_vals[m] = v; }
        m++;
    };
    struct Ent { int offk, offb; unsigned pk, pb, sk, sb; int rank; };
    std::vector<Ent> L[2];
    for (int spin = 0; spin < 2; spin++) {
        const u64 w = spin ? d.b : d.a;
        const std::vector<u64>& list = spin ? blist : alist;
        for (int t = 0; t < (int)list.size(); t++) {
            const u64 w2 = list[t];
            const int pc = fgk_popc(w2 ^ w);
            if (pc == 2) {
                int hh, ee;
                single_from_strings(w, w2, nn, hh, ee);
                Ent e;
                single_factors(w, w2, nn, hh, ee, e.pk, e.pb, e.sk, e.sb);
                e.offk = ee * nn + hh; e.offb = hh * nn + ee; e.rank = t;
                L[spin].push_back(e);
            } else if (pc == 4) {
                const long j = spin ? column(ia, t) : column(t, ib);
                int h0, h1, e0, e1;
                double_from_strings(w, w2, nn, h0, h1, e0, e1);
                const float rb_ = V.w[((h0 * nn + e0) * nn + h1) * nn + e1];
                const float rk = sym ? V.w[((e0 * nn + h0) * nn + e1) * nn + h1] : 0.f;
                Excitation x, rx;
                x.cls = rx.cls = 2;
                x.h0 = h0; x.h1 = h1; x.e0 = e0; x.e1 = e1;
                rx.h0 = e0; rx.h1 = e1; rx.e0 = h0; rx.e1 = h1;
                const fgk_det kd = {w, 0}, ko = {w2, 0};
                emit(j, rb_, rk, (unsigned)exc_parity_ket(ko, nn, rx), (unsigned)exc_parity_ket(kd, nn, x));
            }
        }
    }
    for (int spin = 0; spin < 2; spin++)
        for (auto& e : L[spin])
            emit(spin ? column(ia, e.rank) : column(e.rank, ib), V.h1[e.offb], sym ? V.h1[e.offk] : 0.f, e.sb, e.sk);
    for (auto& a : L[0])
        for (auto& b : L[1])
            emit(column(a.rank, b.rank), V.g[a.offb * n2 + b.offb], sym ? V.g[a.offk * n2 + b.offk] : 0.f,
                 a.pb ^ b.pb ^ 1u, a.pk ^ b.pk ^ 1u);
    return m;
}

// The PT2 walk of k_pt2_accumulate2 (one bucket): singles lists over every (occupied, virtual)
// pair, alpha-beta candidates from the separable form, same-spin doubles by decode_double.
// Emits (candidate, element) in the kernel's chunk order; the set must equal hc_connections'.
long hc_pt2_walk2(void* h, u64 a, u64 b, u64* out_dets, float* out_el, long cap)
{
    HcHam* H = (HcHam*)h;
    const HamView& V = H->V;
    const int nn = V.n_orb, n2 = nn * nn;
    uint8_t buf[256];
    DetCtx c;
    fgk_det d = {a, b};
    detctx_fill_host(c, nn, d, buf);
    long m = 0;
    auto put = [&](fgk_det o, float el) {
        if (m < cap) { out_dets[2 * m] = o.a; out_dets[2 * m + 1] = o.b; out_el[m] = el; }
        m++;
    };
    struct Ent { int hh, ee; unsigned pk, s1; };
    std::vector<Ent> L[2];
    for (int spin = 0; spin < 2; spin++) {
        const u64 w = spin ? d.b : d.a;
        const uint8_t* occ = spin ? c.occ_b : c.occ_a;
        const uint8_t* virt = spin ? c.virt_b : c.virt_a;
        const int nv = spin ? c.nvb : c.nva, ns = (spin ? c.nob : c.noa) * nv;
        for (int t = 0; t < ns; t++) {
            Ent e;
            e.hh = occ[t / nv]; e.ee = virt[t % nv];
            unsigned pb, sb;
            single_factors(w, w ^ orb_bit(nn, e.hh) ^ orb_bit(nn, e.ee), nn, e.hh, e.ee, e.pk, pb, e.s1, sb);
            L[spin].push_back(e);
        }
    }
    for (int spin = 0; spin < 2; spin++)
        for (auto& e : L[spin]) {
            const float v = V.h1[e.ee * nn + e.hh];
            if (!((v < 0 ? -v : v) > 1e-12f)) continue;
            fgk_det o = d;
            const u64 flip = orb_bit(nn, e.hh) ^ orb_bit(nn, e.ee);
            if (spin) o.b ^= flip; else o.a ^= flip;
            put(o, e.s1 ? -v : v);
        }
    const int sizes[2] = {c.n_aa, c.n_bb};
    for (int st = 2; st <= 3; st++)
        for (int t = 0; t < sizes[st - 2]; t++) {
            Excitation x;
            decode_double(c, st, t, x);
            float el;
            if (ket_element_fast(V, d, x, ldf_host, el)) put(apply_excitation(d, nn, x), el);
        }
    for (auto& ea : L[0])
        for (auto& eb : L[1]) {
            const float v = V.g[(ea.ee * nn + ea.hh) * n2 + eb.ee * nn + eb.hh];
            if (!((v < 0 ? -v : v) > 1e-12f)) continue;
            fgk_det o = d;
            o.a ^= orb_bit(nn, ea.hh) ^ orb_bit(nn, ea.ee);
            o.b ^= orb_bit(nn, eb.hh) ^ orb_bit(nn, eb.ee);
            put(o, ((ea.pk ^ eb.pk) & 1u) ? v : -v);
        }
    return m;
}

// fast split evaluation (exc_value_* / exc_parity_*) against the generic element functions,
// over every excitation of the given determinants; returns the number of mismatches
long hc_check_split(void* h, const u64* dets, long n_dets)
{
    HcHam* H = (HcHam*)h;
    long bad = 0;
    for (long i = 0; i < n_dets; i++) {
        uint8_t buf[256];
        DetCtx c;
        fgk_det d = {dets[2 * i], dets[2 * i + 1]};
        detctx_fill_host(c, H->V.n_orb, d, buf);
        auto chk = [&](const Excitation& x) {
            float vk = 0.f, vb = 0.f;
            bool kk = ket_element(H->V, d, x, ldf_host, vk);
            bool kb = bra_element(H->V, d, x, ldf_host, vb);
            float rk = exc_value_ket(H->V, x, ldf_host), rb = exc_value_bra(H->V, x, ldf_host);
            bool fk = (rk < 0 ? -rk : rk) > 1e-12f, fb = (rb < 0 ? -rb : rb) > 1e-12f;
            if (fk != kk || fb != kb) { bad++; return; }
            if (kk && (exc_parity_ket(d, c.n, x) ? -rk : rk) != vk) bad++;
            if (kb && (exc_parity_bra(d, c.n, x) ? -rb : rb) != vb) bad++;
        };
        for (int t = 0; t < c.n_s; t++) {
            int p, q; bool va, vb;
            decode_single(c, t, p, q, va, vb);
            Excitation x; x.h1 = x.e1 = 0; x.h0 = q; x.e0 = p;
            if (va) { x.cls = 0; chk(x); }
            if (vb) { x.cls = 1; chk(x); }
        }
        const int sizes[3] = {c.n_aa, c.n_bb, c.n_ab};
        for (int st = 2; st <= 4; st++)
            for (int t = 0; t < sizes[st - 2]; t++) {
                Excitation x;
                decode_double(c, st, t, x);
                chk(x);
            }
    }
    return bad;
}

// String-driven row (the k_projh4 strategy, fgk_lists.cuh): replacement lists per distinct string
// with signed values and separable alpha-beta factors, rows assembled from list entries and the
// (alpha rank, beta rank) pair map in the kernel's order: diagonal, alpha singles, beta singles,
// alpha doubles, beta doubles, alpha-beta (beta single outside, alpha single inside).
long hc_bra_row4(void* h, const u64* basis, long n, long i, int mode, int* out_cols, double* out_vals, long cap)
{
    HcHam* H = (HcHam*)h;
    const HamView& V = H->V;
    std::map<u64, int> arank, brank;
    for (long k = 0; k < n; k++) { arank[basis[2 * k]] = 0; brank[basis[2 * k + 1]] = 0; }
    std::vector<u64> alist, blist;
    for (auto& kv : arank) { kv.second = (int)alist.size(); alist.push_back(kv.first); }
    for (auto& kv : brank) { kv.second = (int)blist.size(); blist.push_back(kv.first); }
    std::map<std::pair<int, int>, long> pair;
    for (long k = 0; k < n; k++) pair[{arank[basis[2 * k]], brank[basis[2 * k + 1]]}] = k;   // last index wins
    const fgk_det d = {basis[2 * i], basis[2 * i + 1]};
    const int ia = arank[d.a], ib = brank[d.b];
    const bool sym = (mode & 1) != 0, drop0 = (mode & 2) != 0;
    std::vector<LEntry> S[2], D[2];
    for (int spin = 0; spin < 2; spin++) {
        const u64 w = spin ? d.b : d.a;
        const std::vector<u64>& list = spin ? blist : alist;
        for (int t = 0; t < (int)list.size(); t++) {
            const int pc = fgk_popc(list[t] ^ w);
            if (pc == 2) S[spin].push_back(single_entry(V, w, list[t], t, ldf_host));
            else if (pc == 4) D[spin].push_back(double_entry(V, w, list[t], t, ldf_host));
        }
    }
    long m = 0;
    if (m < cap) { out_cols[m] = (int)i; out_vals[m] = diag_element(V, d, ldd_host); }
    m++;
    auto column = [&](int ra, int rb) -> long {
        auto it = pair.find({ra, rb});
        return it == pair.end() ? -1 : it->second;
    };
    auto one = [&](float vij, float vji, long j) {
        float f;
        double v;
        bool exact;
        if (j >= 0 && entry_value_f32(sym, drop0, vij, vji, f, v, exact)) {     // the kernel's decision function
            if (m < cap) { out_cols[m] = (int)j; out_vals[m] = exact ? (double)f : v; }
            m++;
        }
    };
    for (const LEntry& e : S[0]) one(e.vij, e.vji, column(e.rank, ib));
    for (const LEntry& e : S[1]) one(e.vij, e.vji, column(ia, e.rank));
    for (const LEntry& e : D[0]) one(e.vij, e.vji, column(e.rank, ib));
    for (const LEntry& e : D[1]) one(e.vij, e.vji, column(ia, e.rank));
    for (const LEntry& eb : S[1])
        for (const LEntry& ea : S[0]) {
            const long j = column(ea.rank, eb.rank);
            if (j < 0) continue;
            float vij, vji;
            ab_values(V, ea, eb, sym, ldf_host, vij, vji);
            one(vij, vji, j);
        }
    return m;
}

// the PT2 accumulator arithmetic (fgk_pt2.cu: fx_from_double -> 128-bit integer adds with the
// carry rule of fx_atomic_add -> fx_to_double): sum of vals[order[i]]; returns -1 if out of range
int hc_fx_sum(const double* vals, const long* order, long n, double* out)
{
    u64 acc_lo = 0, acc_hi = 0;
    for (long i = 0; i < n; i++) {
        u64 lo, hi;
        if (!fx_from_double(vals[order[i]], lo, hi)) return -1;
        const u64 old = acc_lo;
        acc_lo += lo;
        acc_hi += hi + ((old + lo) < old ? 1ull : 0ull);
    }
    *out = fx_to_double(acc_lo, acc_hi);
    return 0;
}

}  // extern "C"
