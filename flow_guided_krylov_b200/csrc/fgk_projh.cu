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
    for (i64 i = row_begin + warp0; i < row_end; i += nwarps) {
        ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + i);
        fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, H.n_orb, d, s_lists[wib], lane);
        i64 pos = FILL ? row_ptr[i - row_begin] : 0;
        // diagonal first (always stored, like skqd.py:394-397)
        if (FILL) {
            const double dg = warp_diag_element(H, d, lane);
            if (lane == 0) {
                cols[pos] = (int32_t)i;
                vals[pos] = dg;
            }
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

// ---- structured row builder -----------------------------------------------------------
// The flat walk above visits every excitation of the row determinant (52,704 at 32
// orbitals / 8+8 electrons) although only those whose alpha AND beta strings occur in the
// basis can hit (2,220 for the CAS basis).  This kernel turns the string sets into the
// loop structure itself:
//   1. per spin, the list of single excitations whose target string is in the basis
//      (found by enumeration + set probe, or -- when the basis has fewer distinct strings
//      than the row has double excitations -- by one scan over the distinct strings,
//      classifying each by popc(string ^ own): 2 = single, 4 = double);
//   2. same-spin doubles: from the same scan (or enumeration + set probe);
//   3. alpha-beta doubles: the product of the two singles lists -- the only candidates
//      whose both strings are in the basis.
// Only these survivors reach the full-key probe and the element evaluation.
struct ProjLists {
    const u64* alist; i64 na;      // distinct alpha strings of the basis
    const u64* blist; i64 nb;
    int scan_a, scan_b;            // 1: classify by scanning the list, 0: enumerate + probe the set
};

static const int SINGLE_LIST_CAP = 1024;   // n_occ * n_virt <= 32 * 32

template <bool FILL>
__global__ void __launch_bounds__(FGK_BLOCK)
k_projh2(HamView H, IndexView I, ProjLists PL, i64 row_begin, i64 row_end, int mode,
         i64* __restrict__ counts, const i64* __restrict__ row_ptr, const i64* __restrict__ slice_ptr,
         int32_t* __restrict__ cols, double* __restrict__ vals)
{
    __shared__ WarpLists s_lists[FGK_WARPS_PER_BLOCK];
    __shared__ unsigned short s_single[FGK_WARPS_PER_BLOCK][2][SINGLE_LIST_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const bool sym = (mode & FGK_H_SYM) != 0, drop0 = (mode & FGK_H_DROP_ZEROS) != 0;
    const int n = H.n_orb;
    LdgF ldf;
    for (i64 i = row_begin + warp0; i < row_end; i += nwarps) {
        ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + i);
        fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, n, d, s_lists[wib], lane);
        // pos = index of the next entry INSIDE the row; at(k) = where entry k of this row lives:
        // CSR (row_ptr) or directly SELL-32 (slice_ptr: lane-interleaved pairs, see fgk_spmv.cu)
        const i64 rl = i - row_begin;
        const i64 row_base = !FILL ? 0 : (slice_ptr ? __ldg(slice_ptr + (rl >> 5)) + 2 * (rl & 31)
                                                    : __ldg(row_ptr + rl));
        auto at = [&](i64 k) -> i64 {
            return slice_ptr ? row_base + (k >> 1) * 64 + (k & 1) : row_base + k;
        };
        i64 pos = 0;
        if (FILL) {
            const double dg = warp_diag_element(H, d, lane);
            if (lane == 0) {
                cols[at(0)] = (int32_t)i;
                vals[at(0)] = dg;
            }
        }
        pos += 1;
        // full-key probe + element of one candidate per lane; rank inside the row by ballot.
        // The count pass needs only |value| > 1e-12 (no signs) unless exact zeros are dropped.
        auto emit = [&](bool valid, const Excitation& x) {
            int j = -1;
            double v = 0.0;
            bool keep = false;
            if (valid) {
                fgk_det o = apply_excitation(d, n, x);
                j = index_find(I, o);
                if (j >= 0) {
                    const float rb = exc_value_bra(H, x, ldf);
                    const bool kij = fabsf(rb) > 1e-12f;
                    float rk = 0.f;
                    bool kji = false;
                    if (sym) { rk = exc_value_ket(H, x, ldf); kji = fabsf(rk) > 1e-12f; }
                    keep = kij || kji;
                    if (keep && (FILL || drop0)) {
                        Excitation rx;
                        rx.cls = x.cls; rx.h0 = x.e0; rx.h1 = x.e1; rx.e0 = x.h0; rx.e1 = x.h1;
                        const float vij = kij ? (exc_parity_ket(o, n, rx) ? -rb : rb) : 0.f;   // <i|H|j>, ket j = o
                        const float vji = kji ? (exc_parity_ket(d, n, x) ? -rk : rk) : 0.f;    // <j|H|i>, ket i = d
                        v = sym ? 0.5 * ((double)vij + (double)vji) : (double)vij;
                        if (drop0 && v == 0.0) keep = false;
                    }
                }
            }
            unsigned b = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) {
                const i64 o = at(pos + __popc(b & lt));
                cols[o] = j;
                vals[o] = v;
            }
            pos += __popc(b);
        };
        int n_single[2] = {0, 0};
#pragma unroll 1
        for (int spin = 0; spin < 2; spin++) {
            const u64 w = spin ? d.b : d.a;
            unsigned short* L = s_single[wib][spin];
            int cnt = 0;
            auto push = [&](bool ok, int hh, int ee) {
                unsigned b = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    int o = cnt + __popc(b & lt);
                    if (o < SINGLE_LIST_CAP) L[o] = (unsigned short)((hh << 8) | ee);
                }
                cnt += __popc(b);
            };
            const bool scan = spin ? PL.scan_b : PL.scan_a;
            if (scan) {
                const u64* list = spin ? PL.blist : PL.alist;
                const i64 nl = spin ? PL.nb : PL.na;
                for (i64 t0 = 0; t0 < nl; t0 += 32) {
                    i64 t = t0 + lane;
                    u64 w2 = t < nl ? __ldg(list + t) : w;
                    int pc = __popcll(w2 ^ w);
                    int hh = 0, ee = 0;
                    if (pc == 2) single_from_strings(w, w2, n, hh, ee);
                    push(pc == 2, hh, ee);
                    Excitation x;
                    x.cls = 2 + spin; x.h0 = x.h1 = x.e0 = x.e1 = 0;
                    if (pc == 4) double_from_strings(w, w2, n, x.h0, x.h1, x.e0, x.e1);
                    if (__any_sync(0xffffffffu, pc == 4)) emit(pc == 4, x);
                }
            } else {
                const u64* set = spin ? I.bset : I.aset;
                const u64 smask = spin ? I.bmask : I.amask;
                const uint8_t* occ = spin ? c.occ_b : c.occ_a;
                const uint8_t* virt = spin ? c.virt_b : c.virt_a;
                const int no = spin ? c.nob : c.noa, nv = spin ? c.nvb : c.nva;
                for (int t0 = 0; t0 < no * nv; t0 += 32) {
                    int t = t0 + lane;
                    bool ok = false;
                    int hh = 0, ee = 0;
                    if (t < no * nv) {
                        hh = occ[t / nv]; ee = virt[t % nv];
                        ok = set_has(set, smask, w ^ orb_bit(n, hh) ^ orb_bit(n, ee));
                    }
                    push(ok, hh, ee);
                }
                const int size = spin ? c.n_bb : c.n_aa;
                for (int t0 = 0; t0 < size; t0 += 32) {
                    int t = t0 + lane;
                    Excitation x;
                    x.cls = 2 + spin; x.h0 = x.h1 = x.e0 = x.e1 = 0;
                    bool ok = false;
                    if (t < size) {
                        decode_double(c, 2 + spin, t, x);
                        u64 w2 = w ^ orb_bit(n, x.h0) ^ orb_bit(n, x.h1) ^ orb_bit(n, x.e0) ^ orb_bit(n, x.e1);
                        ok = set_has(set, smask, w2);
                    }
                    if (__any_sync(0xffffffffu, ok)) emit(ok, x);
                }
            }
            n_single[spin] = cnt < SINGLE_LIST_CAP ? cnt : SINGLE_LIST_CAP;
        }
        __syncwarp();
        // singles
#pragma unroll 1
        for (int spin = 0; spin < 2; spin++) {
            const unsigned short* L = s_single[wib][spin];
            for (int t0 = 0; t0 < n_single[spin]; t0 += 32) {
                int t = t0 + lane;
                Excitation x;
                x.cls = spin; x.h0 = x.h1 = x.e0 = x.e1 = 0;
                bool ok = t < n_single[spin];
                if (ok) { unsigned short he = L[t]; x.h0 = he >> 8; x.e0 = he & 0xff; }
                emit(ok, x);
            }
        }
        // alpha-beta doubles: product of the two lists
        for (int ia = 0; ia < n_single[0]; ia++) {
            const unsigned short ha = s_single[wib][0][ia];
            for (int t0 = 0; t0 < n_single[1]; t0 += 32) {
                int t = t0 + lane;
                Excitation x;
                x.cls = 4; x.h0 = ha >> 8; x.e0 = ha & 0xff; x.h1 = x.e1 = 0;
                bool ok = t < n_single[1];
                if (ok) { unsigned short he = s_single[wib][1][t]; x.h1 = he >> 8; x.e1 = he & 0xff; }
                emit(ok, x);
            }
        }
        if (!FILL && lane == 0) counts[i - row_begin] = pos;
        __syncwarp();
    }
}


// ---- rank-based row builder --------------------------------------------------------------
// Same structure as k_projh2 (string lists drive the loops) but nothing is hashed and nothing
// is decoded twice.  ncu on k_projh2 (profiles/r01h): a third of its instructions and over
// half of its stall samples sit in the 128-bit hash probe, another eighth in 64-bit (n,n,n,n)
// index arithmetic.  Here
//   * a connected string is found by scanning the sorted distinct-string list, so its RANK is
//     known; the column of (alpha rank, beta rank) is one 4-byte load from the dense pair
//     table of the index (or, for sparse bases, one probe with the strings read back from
//     the lists);
//   * for alpha-beta doubles both the table offset and the sign parity split into an alpha
//     part and a beta part: g[(e0 n + h0) n^2 + (e1 n + h1)], parity = par_a ^ par_b ^ 1
//     (exc_parity_ket, class 4); the parts are computed once per single excitation and kept
//     in shared memory, and the product loop runs flat over n_a * n_b with full lanes;
//   * same-spin doubles met during the scan are queued (ring of 64 ranks per warp) and
//     evaluated 32 at a time, instead of at the scan's lane occupancy (27 % on config 4).
// Requires the scan mode for both spins (fewer distinct strings than same-spin excitations
// per row); otherwise fgk_projh_* fall back to k_projh2.
struct __align__(8) SEntry {
    unsigned short offk;   // particle * n + hole  (ket-side table offset part, < 4096)
    unsigned short offb;   // hole * n + particle  (bra side) | parities << 12:
                           // bit12 pk, bit13 pb (alpha-beta factors), bit14 s1 ket, bit15 s1 bra
    int rank;              // rank of the target string
};

template <bool FILL, bool DENSE>
__global__ void __launch_bounds__(FGK_BLOCK)
k_projh3(HamView H, IndexView I, ProjLists PL, int cap, i64 row_begin, i64 row_end, int mode,
         i64* __restrict__ counts, const i64* __restrict__ row_ptr, const i64* __restrict__ slice_ptr,
         int32_t* __restrict__ cols, double* __restrict__ vals)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SEntry* const s_single = reinterpret_cast<SEntry*>(s_dyn) + (size_t)wib * 2 * cap;
    int* const s_queue = reinterpret_cast<int*>(s_dyn + sizeof(SEntry) * 2 * (size_t)cap * FGK_WARPS_PER_BLOCK) + wib * 64;
    const unsigned lt = (1u << lane) - 1u;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const bool sym = (mode & FGK_H_SYM) != 0, drop0 = (mode & FGK_H_DROP_ZEROS) != 0;
    const bool need_sign = FILL || drop0;
    const int n = H.n_orb, n2 = n * n;
    const int nbs = (int)I.n_bstr;
    for (i64 i = row_begin + warp0; i < row_end; i += nwarps) {
        const ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + i);
        const fgk_det d = {dv.x, dv.y};
        const int ia = __ldg(I.ra + i), ib = __ldg(I.rb + i);
        const i64 rl = i - row_begin;
        const i64 row_base = !FILL ? 0 : (slice_ptr ? __ldg(slice_ptr + (rl >> 5)) + 2 * (rl & 31)
                                                    : __ldg(row_ptr + rl));
        auto at = [&](i64 k) -> i64 {
            return slice_ptr ? row_base + (k >> 1) * 64 + (k & 1) : row_base + k;
        };
        i64 pos = 0;
        if (FILL) {
            const double dg = warp_diag_element(H, d, lane);
            if (lane == 0) {
                cols[at(0)] = (int32_t)i;
                vals[at(0)] = dg;
            }
        }
        pos += 1;
        // column of (alpha rank, beta rank), or -1
        auto column = [&](int ra2, int rb2) -> int {
            if (DENSE) return __ldg(I.pair + (i64)ra2 * nbs + rb2);
            fgk_det o = {__ldg(PL.alist + ra2), __ldg(PL.blist + rb2)};
            return index_find(I, o);
        };
        // rb_ = table value seen from the ket j (gives <i|H|j>), rk = seen from the ket i
        auto emit = [&](bool valid, int j, float rb_, float rk, int parb, int park) {
            double v = 0.0;
            bool keep = false;
            if (valid) {
                const bool kij = fabsf(rb_) > 1e-12f;
                const bool kji = sym && fabsf(rk) > 1e-12f;
                keep = kij || kji;
                if (keep && need_sign) {
                    const float vij = kij ? (parb ? -rb_ : rb_) : 0.f;
                    const float vji = kji ? (park ? -rk : rk) : 0.f;
                    v = sym ? 0.5 * ((double)vij + (double)vji) : (double)vij;
                    if (drop0 && v == 0.0) keep = false;
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) {
                const i64 o = at(pos + __popc(b & lt));
                cols[o] = j;
                vals[o] = v;
            }
            pos += __popc(b);
        };
        int n_single[2] = {0, 0};
#pragma unroll 1
        for (int spin = 0; spin < 2; spin++) {
            const u64 w = spin ? d.b : d.a;
            const u64* list = spin ? PL.blist : PL.alist;
            const int nl = (int)(spin ? PL.nb : PL.na);
            SEntry* L = s_single + spin * cap;
            int cnt = 0, qh = 0, qn = 0;
            // evaluate m (<= 32) queued same-spin doubles, one per lane
            auto flush = [&](int m) {
                const bool valid0 = lane < m;
                int j = -1;
                float rb_ = 0.f, rk = 0.f;
                int parb = 0, park = 0;
                if (valid0) {
                    const int t = s_queue[(qh + lane) & 63];
                    j = spin ? column(ia, t) : column(t, ib);
                    if (j >= 0) {
                        const u64 w2 = __ldg(list + t);
                        int h0, h1, e0, e1;
                        double_from_strings(w, w2, n, h0, h1, e0, e1);
                        rb_ = __ldg(H.w + ((h0 * n + e0) * n + h1) * n + e1);
                        if (sym) rk = __ldg(H.w + ((e0 * n + h0) * n + e1) * n + h1);
                        if (need_sign) {
                            Excitation x, rx;
                            x.cls = rx.cls = 2;       // the same-spin parity only reads the changed word
                            x.h0 = h0; x.h1 = h1; x.e0 = e0; x.e1 = e1;
                            rx.h0 = e0; rx.h1 = e1; rx.e0 = h0; rx.e1 = h1;
                            const fgk_det kd = {w, 0}, ko = {w2, 0};
                            park = exc_parity_ket(kd, n, x);
                            parb = exc_parity_ket(ko, n, rx);
                        }
                    }
                }
                emit(j >= 0, j, rb_, rk, parb, park);
            };
            for (int t0 = 0; t0 < nl; t0 += 32) {
                const int t = t0 + lane;
                const u64 w2 = t < nl ? __ldg(list + t) : w;
                const int pc = __popcll(w2 ^ w);
                const unsigned bs = __ballot_sync(0xffffffffu, pc == 2);
                if (pc == 2) {
                    int hh, ee;
                    single_from_strings(w, w2, n, hh, ee);
                    unsigned pk, pb, sk, sb;
                    single_factors(w, w2, n, hh, ee, pk, pb, sk, sb);
                    const int o = cnt + __popc(bs & lt);
                    if (o < cap) {
                        SEntry e;
                        e.offk = (unsigned short)(ee * n + hh);
                        e.offb = (unsigned short)((hh * n + ee) | ((pk | (pb << 1) | (sk << 2) | (sb << 3)) << 12));
                        e.rank = t;
                        L[o] = e;
                    }
                }
                cnt += __popc(bs);
                const unsigned bd = __ballot_sync(0xffffffffu, pc == 4);
                if (bd) {
                    if (pc == 4) s_queue[(qh + qn + __popc(bd & lt)) & 63] = t;
                    qn += __popc(bd);
                    __syncwarp();
                    if (qn >= 32) {
                        flush(32);
                        qh = (qh + 32) & 63;
                        qn -= 32;
                        __syncwarp();
                    }
                }
            }
            if (qn > 0) { flush(qn); __syncwarp(); }
            n_single[spin] = cnt < cap ? cnt : cap;
        }
        __syncwarp();
        // singles: column = (target rank, own rank of the other spin)
#pragma unroll 1
        for (int spin = 0; spin < 2; spin++) {
            const SEntry* L = s_single + spin * cap;
            for (int t0 = 0; t0 < n_single[spin]; t0 += 32) {
                const int t = t0 + lane;
                int j = -1, parb = 0, park = 0;
                float rb_ = 0.f, rk = 0.f;
                if (t < n_single[spin]) {
                    const SEntry e = L[t];
                    j = spin ? column(ia, e.rank) : column(e.rank, ib);
                    if (j >= 0) {
                        rb_ = __ldg(H.h1 + (e.offb & 0xfff));
                        if (sym) rk = __ldg(H.h1 + e.offk);
                        park = (e.offb >> 14) & 1;
                        parb = (e.offb >> 15) & 1;
                    }
                }
                emit(j >= 0, j, rb_, rk, parb, park);
            }
        }
        // alpha-beta doubles: flat product of the two singles lists
        const int nsa = n_single[0], nsb = n_single[1], total = nsa * nsb;
        const SEntry* La = s_single;
        const SEntry* Lb = s_single + cap;
        const float inv_nsb = nsb ? 1.0f / (float)nsb : 0.f;
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            int j = -1, parb = 0, park = 0;
            float rb_ = 0.f, rk = 0.f;
            if (t < total) {
                // t -> (ka, kb): float reciprocal + one-step fix-up (t < 2^20) instead of an integer division
                int ka = (int)((float)t * inv_nsb), kb = t - ka * nsb;
                if (kb < 0) { ka--; kb += nsb; }
                else if (kb >= nsb) { ka++; kb -= nsb; }
                const SEntry ea = La[ka], eb = Lb[kb];
                j = column(ea.rank, eb.rank);
                if (j >= 0) {
                    rb_ = __ldg(H.g + (ea.offb & 0xfff) * n2 + (eb.offb & 0xfff));
                    if (sym) rk = __ldg(H.g + ea.offk * n2 + eb.offk);
                    const unsigned px = (unsigned)(ea.offb ^ eb.offb);
                    park = ((px >> 12) ^ 1) & 1;
                    parb = ((px >> 13) ^ 1) & 1;
                }
            }
            emit(j >= 0, j, rb_, rk, parb, park);
        }
        if (!FILL && lane == 0) counts[i - row_begin] = pos;
        __syncwarp();
    }
}

static int grid_rows(i64 rows, int device)
{
    i64 need = (rows + FGK_WARPS_PER_BLOCK - 1) / FGK_WARPS_PER_BLOCK;
    i64 cap = (i64)fgk_sm_count(device) * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// scan the distinct-string list when it is shorter than the row's own enumeration
static ProjLists proj_lists(fgk_ham_t h, fgk_index_t idx)
{
    ProjLists P;
    P.alist = idx->alist; P.na = idx->n_alpha_strings;
    P.blist = idx->blist; P.nb = idx->n_beta_strings;
    auto c2 = [](i64 m) { return m * (m - 1) / 2; };
    const i64 n = h->v.n_orb, na = h->v.n_alpha, nb = h->v.n_beta;
    const i64 enum_a = na * (n - na) + c2(na) * c2(n - na);
    const i64 enum_b = nb * (n - nb) + c2(nb) * c2(n - nb);
    P.scan_a = P.na <= enum_a ? 1 : 0;
    P.scan_b = P.nb <= enum_b ? 1 : 0;
    return P;
}


// rank-based builder: usable when both spins are in scan mode and the index has its rank form
static bool use_rank_builder(fgk_ham_t h, fgk_index_t idx, int mode, const ProjLists& P)
{
    if (mode & (FGK_H_FLAT_WALK | FGK_H_HASH_WALK)) return false;
    return P.scan_a && P.scan_b && idx->ra != nullptr && idx->rb != nullptr;
}

template <bool FILL>
static int launch_projh3(fgk_ham_t h, fgk_index_t idx, const ProjLists& P, i64 row_begin, i64 row_end,
                         int mode, i64* counts, const i64* row_ptr, const i64* slice_ptr, int32_t* cols,
                         double* vals, cudaStream_t st)
{
    const int n = h->v.n_orb, na = h->v.n_alpha, nb = h->v.n_beta;
    int cap = na * (n - na) > nb * (n - nb) ? na * (n - na) : nb * (n - nb);
    if (cap < 1) cap = 1;
    const size_t smem = sizeof(SEntry) * 2 * (size_t)cap * FGK_WARPS_PER_BLOCK + 64 * sizeof(int) * FGK_WARPS_PER_BLOCK;
    const int grid = grid_rows(row_end - row_begin, h->device);
    if (idx->pair) {
        if (smem > 48 * 1024)
            FGK_CUDA(cudaFuncSetAttribute(k_projh3<FILL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_projh3<FILL, true><<<grid, FGK_BLOCK, smem, st>>>(h->v, idx->v, P, cap, row_begin, row_end, mode,
                                                            counts, row_ptr, slice_ptr, cols, vals);
    } else {
        if (smem > 48 * 1024)
            FGK_CUDA(cudaFuncSetAttribute(k_projh3<FILL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_projh3<FILL, false><<<grid, FGK_BLOCK, smem, st>>>(h->v, idx->v, P, cap, row_begin, row_end, mode,
                                                             counts, row_ptr, slice_ptr, cols, vals);
    }
    FGK_LAUNCH_CHECK();
    return FGK_OK;
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
    const ProjLists PLs = proj_lists(h, idx);
    if (use_rank_builder(h, idx, mode, PLs))
        return launch_projh3<false>(h, idx, PLs, row_begin, row_end, mode, (i64*)counts, nullptr, nullptr,
                                    nullptr, nullptr, (cudaStream_t)stream);
    if (mode & FGK_H_FLAT_WALK)
        k_projh<false><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
            h->v, idx->v, row_begin, row_end, mode, (i64*)counts, nullptr, nullptr, nullptr);
    else
        k_projh2<false><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
            h->v, idx->v, proj_lists(h, idx), row_begin, row_end, mode, (i64*)counts, nullptr, nullptr,
            nullptr, nullptr);
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
    const ProjLists PLs = proj_lists(h, idx);
    if (use_rank_builder(h, idx, mode, PLs))
        return launch_projh3<true>(h, idx, PLs, row_begin, row_end, mode, nullptr, (const i64*)row_ptr, nullptr,
                                   cols, vals, (cudaStream_t)stream);
    if (mode & FGK_H_FLAT_WALK)
        k_projh<true><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
            h->v, idx->v, row_begin, row_end, mode, nullptr, (const i64*)row_ptr, cols, vals);
    else
        k_projh2<true><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
            h->v, idx->v, proj_lists(h, idx), row_begin, row_end, mode, nullptr, (const i64*)row_ptr,
            nullptr, cols, vals);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

// fill rows straight into SELL-32 storage (no CSR copy): slice_ptr as for fgk_sell_fill; the
// arrays must be zero-filled by the caller (padding entries stay value 0 / column 0)
extern "C" int fgk_projh_fill_sell(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end,
                                   int mode, const int64_t* slice_ptr, int32_t* sell_cols,
                                   double* sell_vals, void* stream)
{
    if (!h || !idx) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill_sell: null handle");
    if (row_begin < 0 || row_end > idx->v.n || row_begin > row_end)
        return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill_sell: bad row range");
    if (row_begin == row_end) return FGK_OK;
    if (!slice_ptr || !sell_cols || !sell_vals) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill_sell: null pointer");
    if (mode & FGK_H_FLAT_WALK) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_projh_fill_sell: not with FLAT_WALK");
    if (h->device != idx->device) return fgk_fail(FGK_ERR_ARG, "fgk_projh_fill_sell: device mismatch");
    FGK_CUDA(cudaSetDevice(h->device));
    const ProjLists PLs = proj_lists(h, idx);
    if (use_rank_builder(h, idx, mode, PLs))
        return launch_projh3<true>(h, idx, PLs, row_begin, row_end, mode, nullptr, nullptr,
                                   (const i64*)slice_ptr, sell_cols, sell_vals, (cudaStream_t)stream);
    k_projh2<true><<<grid_rows(row_end - row_begin, h->device), FGK_BLOCK, 0, (cudaStream_t)stream>>>(
        h->v, idx->v, PLs, row_begin, row_end, mode, nullptr, nullptr,
        (const i64*)slice_ptr, sell_cols, sell_vals);
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
