// fgk_pt2.cu -- K7/K8: PT2 residual expansion.
//
// One warp per SOURCE determinant j ("ket mode", exactly the connections the
// reference's get_connections(j) emits): every connection x that is not in the
// basis adds c_j * <x|H|j> to the FP64 accumulator of x in a device hash map
// (open addressing; 64-bit table entries tag<<32|slot pointing into a pool of
// 32-byte entries {key, FP64 sum}).  This is phase 1 of
// SelectedCIExpander._find_important_configs (residual_expansion.py:498-522) --
// there a Python dict keyed by hash(bytes).  fgk_pt2_export is phase 2
// (:527-548): diagonal of every candidate and coupling^2 / (|E - E_x| + 1e-10).
#include <stdlib.h>

#include "fgk_internal.cuh"

// exact accumulation: fx_from_double / fx_to_double live in fgk_core.cuh (shared with the CPU self-check)
__device__ __forceinline__ void fx_atomic_add(u64* acc, u64 lo, u64 hi)
{
    const u64 old = atomicAdd((unsigned long long*)acc, (unsigned long long)lo);
    const u64 add_hi = hi + ((old + lo) < old ? 1ull : 0ull);      // carry out of the low word
    if (add_hi) atomicAdd((unsigned long long*)(acc + 1), (unsigned long long)add_hi);
}

// insert-or-accumulate; returns false on pool overflow.
// Pool entry = 32 bytes {alpha, beta, accumulator lo, accumulator hi}: the key compare and the
// accumulation of a repeat candidate touch ONE 32-byte sector (the sweep is bound by random
// DRAM sectors).  Publication protocol: the thread that wins the table slot (CAS empty ->
// tag|LOCKED) is the only one that claims a pool entry; it writes the entry complete, first
// contribution included, and then stores tag|slot.  A thread that meets tag|LOCKED with its own
// tag waits for the publication.  So a pass needs exactly one pool entry per DISTINCT candidate
// (a loser of the race used to burn one per simultaneous duplicate), the pool needs no clearing
// between sweeps and a new candidate costs no accumulator atomic.
static const unsigned PT2_LOCKED = 0xFFFFFFFEu, PT2_TOMB = 0xFFFFFFFDu;

__device__ __forceinline__ bool pt2_upsert(const Pt2View& W, fgk_det o, u64 h, double val, int mode)
{
    const u64 tag = h >> 32;
    const u64 base = W.region_bits ? (h >> (64 - W.region_bits)) * (W.region_mask + 1) : 0;
    u64 local = (h >> 7) & W.region_mask;
    u64 slot = base + local;
    u64 probes = 0;
    u64 lo, hi = 0;
    if (mode == FGK_PT2_MAXABS) lo = (u64)__double_as_longlong(fabs(val));
    else if (!fx_from_double(val, lo, hi)) { atomicExch(W.counters + 2, 2ull); return false; }
    while (true) {
        if (++probes > W.region_mask + 1) { atomicExch(W.counters + 2, 1ull); return false; }   // region full
        volatile u64* ts = reinterpret_cast<volatile u64*>(W.table + slot);
        u64 e = *ts;
        if (e == FGK_EMPTY) {
            const u64 prev = atomicCAS((unsigned long long*)(W.table + slot), FGK_EMPTY,
                                       (tag << 32) | (u64)PT2_LOCKED);
            if (prev == FGK_EMPTY) {
                const u64 mine = atomicAdd(W.counters, 1ull);
                if (mine >= (u64)W.capacity) {
                    atomicExch(W.counters + 2, 1ull);
                    *ts = (tag << 32) | (u64)PT2_TOMB;          // release the waiters
                    return false;
                }
                ulonglong2* p = reinterpret_cast<ulonglong2*>(W.pool + 4 * mine);
                p[0] = make_ulonglong2(o.a, o.b);
                p[1] = make_ulonglong2(lo, hi);
                __threadfence();
                *ts = (tag << 32) | mine;                        // publish, contribution inside
                return true;
            }
            e = prev;               // somebody else took the slot: inspect it
        }
        if ((e >> 32) == tag) {
            while ((unsigned)e == PT2_LOCKED) e = *ts;          // entry being written
            if ((unsigned)e == PT2_TOMB) return false;
            // L2 read: the entry was completed with __threadfence before the publishing store
            u64* p = W.pool + 4 * (e & 0xffffffffull);
            const ulonglong2 k = __ldcg(reinterpret_cast<const ulonglong2*>(p));
            if (k.x == o.a && k.y == o.b) {
                if (mode == FGK_PT2_MAXABS) atomicMax((unsigned long long*)(p + 2), (unsigned long long)lo);
                else fx_atomic_add(p + 2, lo, hi);
                return true;
            }
        }
        local = (local + 1) & W.region_mask;
        slot = base + local;
    }
}

// work unit = (source, split): the 32-wide index chunks of one source's excitation
// space are dealt round-robin to `n_split` warps, so that a few thousand sources
// (each with 5e4 .. 3e5 connections) still fill 148 SMs x 64 warps.
template <class FS, class FD>
__device__ __forceinline__ void warp_enumerate_split(const DetCtx& c, int lane, int split, int n_split,
                                                     FS&& fs, FD&& fd)
{
    // chunk g of the concatenated index spaces [singles | aa | bb | ab], each padded to 32;
    // this warp takes g = split, split + n_split, ... (a strided loop: walking every chunk and
    // skipping the foreign ones cost 45 % of the kernel's instructions, ncu r01l)
    const int c0 = (c.n_s + 31) >> 5;
    const int c1 = c0 + ((c.n_aa + 31) >> 5);
    const int c2 = c1 + ((c.n_bb + 31) >> 5);
    const int c3 = c2 + ((c.n_ab + 31) >> 5);
    for (int g = split; g < c3; g += n_split) {
        if (g < c0) {
            int t = g * 32 + lane, p = 0, q = 0;
            bool va = false, vb = false;
            if (t < c.n_s) decode_single(c, t, p, q, va, vb);
            fs(va, vb, p, q);
        } else {
            const int st = g < c1 ? 2 : (g < c2 ? 3 : 4);
            const int base = st == 2 ? c0 : (st == 3 ? c1 : c2);
            const int size = st == 2 ? c.n_aa : (st == 3 ? c.n_bb : c.n_ab);
            const int t = (g - base) * 32 + lane;
            Excitation x;
            x.cls = st; x.h0 = x.h1 = x.e0 = x.e1 = 0;
            const bool valid = t < size;
            if (valid) decode_double(c, st, t, x);
            fd(valid, x);
        }
    }
}

// append (determinant, value) to the partition queue chosen by the top hash bits; called by
// all 32 lanes (push = false for idle lanes); one cursor atomic per distinct queue per warp
__device__ __forceinline__ void pt2_queue_push(const Pt2View& W, bool push, fgk_det o, u64 h, double val,
                                               int lane)
{
    const unsigned q = push ? (unsigned)(h >> (64 - W.queue_bits)) : (0x80000000u | (unsigned)lane);
    const unsigned peers = __match_any_sync(0xffffffffu, q);
    if (!push) return;
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(W.qcursors + q, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const i64 pos = (i64)base + __popc(peers & ((1u << lane) - 1u));
    if (pos >= W.qstride) { atomicExch(W.counters + 2, 1ull); return; }
    const i64 at = (i64)q * W.qstride + pos;
    reinterpret_cast<ulonglong2*>(W.qdets)[at] = make_ulonglong2(o.a, o.b);
    W.qvals[at] = val;
}

template <int BPS, bool QUEUE>      // BPS: resident CTAs per SM the register budget is sized for
__global__ void __launch_bounds__(FGK_BLOCK, BPS)
k_pt2_accumulate(HamView H, IndexView I, Pt2View W, const i64* __restrict__ src_idx,
                 const double* __restrict__ coeff, i64 n_src, int n_split, int mode, unsigned n_pass,
                 unsigned pass_id)
{
    __shared__ WarpLists s_lists[FGK_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const i64 n_units = n_src * n_split;
    LdgF ldf;
    i64 tested = 0;
    for (i64 unit = warp0; unit < n_units; unit += nwarps) {
        const i64 sidx = unit / n_split;
        const int split = (int)(unit - sidx * n_split);
        const i64 j = __ldg(src_idx + sidx);
        const double cj = __ldg(coeff + sidx);
        ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + j);
        fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, H.n_orb, d, s_lists[wib], lane);
        auto visit = [&](bool valid, const Excitation& x) {
            float el = 0.f;
            fgk_det o = d;
            u64 h = 0;
            bool push = valid && ket_element_fast(H, d, x, ldf, el);  // reference filter
            if (push) {
                o = apply_excitation(d, c.n, x);
                h = det_hash(o.a, o.b);
                if (n_pass > 1 && (unsigned)((h >> 40) % n_pass) != pass_id) push = false;
            }
            if (push) {
                tested++;
                if (index_find_filtered_h(I, o, h, x.cls) >= 0) push = false;   // in the basis (:513)
            }
            if (QUEUE) pt2_queue_push(W, push, o, h, cj * (double)el, lane);
            else if (push) pt2_upsert(W, o, h, cj * (double)el, mode);
        };
        warp_enumerate_split(
            c, lane, split, n_split,
            [&](bool va, bool vb, int p, int q) {
                Excitation x;
                x.h0 = q; x.e0 = p; x.h1 = 0; x.e1 = 0;
                x.cls = 0;
                visit(va, x);
                x.cls = 1;
                visit(vb, x);
            },
            visit);
    }
    tested = warp_sum_i64(tested);
    if (lane == 0 && tested) atomicAdd(W.counters + 1, (unsigned long long)tested);
}

// ---- accumulate, second form: buckets by the candidate's ALPHA STRING -----------------------
// The bucket of a candidate (pass of a multi-pass sweep, owner rank of a sharded sweep) is
// two-level, n_pass = n_a * n_d: first (word_hash(alpha') >> 40) % n_a on its alpha string,
// then (det_hash >> 40) % n_d on the full key (n_d = 1 up to 64 buckets; the second level only
// exists so that very many passes still split a single alpha string's candidates).  Whole runs
// of the walk share the first-level bucket and are skipped before any per-candidate work:
//   beta singles, beta-beta doubles : alpha' = the source's alpha  -> all or nothing per source
//   alpha-beta doubles (70 % of the connections): alpha' is fixed by the alpha single, so the
//     loop runs over (alpha single) x (chunks of beta singles) and skips foreign alpha singles
//   alpha singles, alpha-alpha doubles : one word hash per candidate
// With per-candidate buckets every rank / pass paid the full walk (decode, table value, parity,
// 128-bit hash) for every connection, owned or not.  As in k_projh3 the alpha-beta table offset
// and sign parity split into per-single parts, precomputed once per work unit in shared memory:
// entry = hole | particle << 8 | pk << 16 | s1 << 17 | owned << 18.
// STAGED (n_orb <= 12, every molecule of the reference): the integral tables h1 | g | w are staged in
// shared memory by one TMA bulk copy per CTA, behind the singles-list area, and read with ld.shared.
template <bool STAGED> struct Pt2Loader { typedef LdgF type; };
template <> struct Pt2Loader<true> { typedef LdsF type; };

template <int BPS, bool STAGED>
__global__ void __launch_bounds__(FGK_BLOCK, BPS)
k_pt2_accumulate2(HamView H, IndexView I, Pt2View W, const i64* __restrict__ src_idx,
                  const double* __restrict__ coeff, i64 n_src, int n_split, int mode, unsigned n_a,
                  unsigned a_id, unsigned n_d, unsigned d_id, int cap, unsigned tab_bytes, unsigned tab_offset)
{
    extern __shared__ __align__(128) unsigned char s_dyn2[];    // [tables (tab_offset bytes in)] after [warps][2][cap] entries
    unsigned* const s_ent = reinterpret_cast<unsigned*>(s_dyn2);
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ WarpLists s_lists[FGK_WARPS_PER_BLOCK];
    if (STAGED) {
        unsigned char* dst = s_dyn2 + tab_offset;
        tma_stage_table(dst, H.h1, tab_bytes, &s_mbar);
        const float* s_tab = reinterpret_cast<const float*>(dst);
        const float* g0 = H.h1;
        H.g = s_tab + (H.g - g0);
        H.w = s_tab + (H.w - g0);
        H.h1 = s_tab;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 warp0 = (i64)blockIdx.x * FGK_WARPS_PER_BLOCK + wib;
    const i64 nwarps = (i64)gridDim.x * FGK_WARPS_PER_BLOCK;
    const i64 n_units = n_src * n_split;
    const int n = H.n_orb, n2 = n * n;
    unsigned* const La = s_ent + (size_t)wib * 2 * cap;
    unsigned* const Lb = La + cap;
    typename Pt2Loader<STAGED>::type ldf;
    i64 tested = 0;
    auto mine = [&](u64 alpha_word) -> bool {
        return n_a <= 1 || (unsigned)((word_hash(alpha_word) >> 40) % n_a) == a_id;
    };
    for (i64 unit = warp0; unit < n_units; unit += nwarps) {
        const i64 sidx = unit / n_split;
        const int split = (int)(unit - sidx * n_split);
        const i64 j = __ldg(src_idx + sidx);
        const double cj = __ldg(coeff + sidx);
        const ulonglong2 dv = __ldg(reinterpret_cast<const ulonglong2*>(I.dets) + j);
        const fgk_det d = {dv.x, dv.y};
        DetCtx c;
        warp_build_ctx(c, n, d, s_lists[wib], lane);
        const int nsa = c.noa * c.nva, nsb = c.nob * c.nvb;
        // singles lists: every (occupied, virtual) pair, reference order (hole-major)
        for (int spin = 0; spin < 2; spin++) {
            const u64 w = spin ? d.b : d.a;
            const uint8_t* occ = spin ? c.occ_b : c.occ_a;
            const uint8_t* virt = spin ? c.virt_b : c.virt_a;
            const int nv = spin ? c.nvb : c.nva, ns = spin ? nsb : nsa;
            unsigned* L = spin ? Lb : La;
            for (int t = lane; t < ns; t += 32) {
                const int hh = occ[t / nv], ee = virt[t % nv];
                const u64 w2 = w ^ orb_bit(n, hh) ^ orb_bit(n, ee);
                unsigned pk, pb, s1, sb;
                single_factors(w, w2, n, hh, ee, pk, pb, s1, sb);
                unsigned own = 1u;
                if (!spin) own = mine(w2) ? 1u : 0u;
                L[t] = (unsigned)hh | ((unsigned)ee << 8) | (pk << 16) | (s1 << 17) | (own << 18);
            }
        }
        __syncwarp();
        const bool own_src = mine(d.a);
        // candidate o with value el (already filtered): count, drop basis members, accumulate
        auto push_candidate = [&](bool push, fgk_det o, float el, int cls) {
            if (push) {
                const u64 h = det_hash(o.a, o.b);
                if (n_d > 1 && (unsigned)((h >> 40) % n_d) != d_id) return;
                tested++;
                if (index_find_filtered_h(I, o, h, cls) < 0) pt2_upsert(W, o, h, cj * (double)el, mode);
            }
        };
        // chunk space: [alpha singles][beta singles][aa][bb][alpha single x beta-single chunks]
        const int cbk = (nsb + 31) >> 5;
        const int g0 = (nsa + 31) >> 5;
        const int g1 = g0 + cbk;
        const int g2 = g1 + ((c.n_aa + 31) >> 5);
        const int g3 = g2 + ((c.n_bb + 31) >> 5);
        const int g4 = g3 + nsa * cbk;
        for (int g = split; g < g4; g += n_split) {
            if (g < g1) {                                   // singles
                const bool beta = g >= g0;
                if (beta && !own_src) continue;
                const int t = (beta ? g - g0 : g) * 32 + lane;
                const int ns = beta ? nsb : nsa;
                bool push = false;
                float el = 0.f;
                fgk_det o = d;
                if (t < ns) {
                    const unsigned e = beta ? Lb[t] : La[t];
                    const int hh = e & 0xff, ee = (e >> 8) & 0xff;
                    if ((e >> 18) & 1u) {
                        const float v = ldf(H.h1 + ee * n + hh);            // molecular.py:234-251
                        if (fabsf(v) > 1e-12f) {
                            push = true;
                            el = ((e >> 17) & 1u) ? -v : v;
                            const u64 flip = orb_bit(n, hh) ^ orb_bit(n, ee);
                            if (beta) o.b ^= flip; else o.a ^= flip;
                        }
                    }
                }
                push_candidate(push, o, el, beta ? 1 : 0);
            } else if (g < g3) {                            // same-spin doubles
                const int st = g < g2 ? 2 : 3;
                if (st == 3 && !own_src) continue;
                const int size = st == 2 ? c.n_aa : c.n_bb;
                const int t = (g - (st == 2 ? g1 : g2)) * 32 + lane;
                bool push = false;
                float el = 0.f;
                fgk_det o = d;
                if (t < size) {
                    Excitation x;
                    decode_double(c, st, t, x);
                    o = apply_excitation(d, n, x);
                    if (st == 3 || mine(o.a)) push = ket_element_fast(H, d, x, ldf, el);
                }
                push_candidate(push, o, el, st);
            } else {                                        // alpha-beta doubles
                const int gg = g - g3;
                const int ka = gg / cbk, kb = (gg - ka * cbk) * 32 + lane;
                const unsigned ea = La[ka];
                if (!((ea >> 18) & 1u)) continue;           // alpha' belongs to another bucket
                bool push = false;
                float el = 0.f;
                fgk_det o = d;
                if (kb < nsb) {
                    const unsigned eb = Lb[kb];
                    const int h0 = ea & 0xff, e0 = (ea >> 8) & 0xff, h1 = eb & 0xff, e1 = (eb >> 8) & 0xff;
                    const float v = ldf(H.g + (e0 * n + h0) * n2 + e1 * n + h1);   // molecular.py:302-318
                    if (fabsf(v) > 1e-12f) {
                        push = true;
                        el = (((ea ^ eb) >> 16) & 1u) ? v : -v;               // parity = pk_a ^ pk_b ^ 1
                        o.a ^= orb_bit(n, h0) ^ orb_bit(n, e0);
                        o.b ^= orb_bit(n, h1) ^ orb_bit(n, e1);
                    }
                }
                push_candidate(push, o, el, 4);
            }
        }
        __syncwarp();
    }
    tested = warp_sum_i64(tested);
    if (lane == 0 && tested) atomicAdd(W.counters + 1, (unsigned long long)tested);
}

__global__ void __launch_bounds__(256)
k_pt2_merge(Pt2View W, const fgk_det* __restrict__ dets, const double* __restrict__ vals, i64 m, int mode)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
        fgk_det o = {d.x, d.y};
        pt2_upsert(W, o, det_hash(o.a, o.b), __ldg(vals + i), mode);
    }
}

// second half of the partitioned sweep: walk the queues in queue (= table region) order and
// upsert; the CTAs that run together work on neighbouring queues, whose table region and
// freshly allocated pool slots stay L2-resident instead of being random DRAM sectors.
__global__ void __launch_bounds__(256)
k_pt2_aggregate(Pt2View W, int mode, i64 per_block)
{
    const i64 total = ((i64)1 << W.queue_bits) * W.qstride;
    const i64 lo = (i64)blockIdx.x * per_block;
    const i64 hi = lo + per_block < total ? lo + per_block : total;
    for (i64 s0 = lo; s0 < hi; s0 += blockDim.x) {
        const i64 s = s0 + threadIdx.x;
        if (s >= hi) continue;
        const i64 q = s / W.qstride, idx = s - q * W.qstride;
        unsigned long long cnt = W.qcursors[q];
        if (cnt > (unsigned long long)W.qstride) cnt = (unsigned long long)W.qstride;
        if ((unsigned long long)idx >= cnt) continue;
        ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2*>(W.qdets) + s);
        fgk_det o = {d.x, d.y};
        pt2_upsert(W, o, det_hash(o.a, o.b), __ldg(W.qvals + s), mode);
    }
}

// live candidates are compacted to the front of the outputs (order is not
// deterministic; every consumer treats them as a set); counters[3] = live count.
__global__ void __launch_bounds__(1024)
k_pt2_export(HamView H, bool have_h, unsigned tab_bytes, Pt2View W, i64 n_slots, double energy, bool fixed_point,
             fgk_det* __restrict__ out_dets, double* __restrict__ out_coupling,
             double* __restrict__ out_diag, double* __restrict__ out_importance)
{
    // h_pp + nibble row-sum tables staged in shared memory by one TMA bulk copy (as in k_diag)
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    const bool staged = have_h && tab_bytes;
    if (staged) {
        tma_stage_table(s_raw, H.hdiag, tab_bytes, &s_mbar);
        const double* s_tab = reinterpret_cast<const double*>(s_raw);
        const double* g0 = H.hdiag;
        H.nib_jk = s_tab + (H.nib_jk - g0);
        H.nib_jab = s_tab + (H.nib_jab - g0);
        H.hdiag = s_tab;
    }
    const int lane = threadIdx.x & 31;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 rounds = (n_slots + stride - 1) / stride;
    for (i64 it = 0; it < rounds; it++) {
        const i64 k = it * stride + (i64)blockIdx.x * blockDim.x + threadIdx.x;
        bool live = false;
        ulonglong2 d = make_ulonglong2(0, 0);
        ulonglong2 acc = make_ulonglong2(0, 0);
        if (k < n_slots) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(W.pool + 4 * k);
            d = p[0];
            acc = p[1];
            live = !(d.x == FGK_EMPTY && d.y == FGK_EMPTY);
        }
        unsigned b = __ballot_sync(0xffffffffu, live);
        if (!b) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(W.counters + 3, (unsigned long long)__popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!live) continue;
        const i64 o = (i64)base + __popc(b & ((1u << lane) - 1u));
        const double cpl = fixed_point ? fx_to_double(acc.x, acc.y) : __longlong_as_double((long long)acc.x);
        if (out_dets) reinterpret_cast<ulonglong2*>(out_dets)[o] = d;
        if (out_coupling) out_coupling[o] = cpl;
        if (have_h) {
            fgk_det dd = {d.x, d.y};
            double ex = staged ? diag_element(H, dd, LdsD())
                               : diag_element(H, dd, [](const double* p) { return __ldg(p); });
            if (out_diag) out_diag[o] = ex;
            if (out_importance) out_importance[o] = cpl * cpl / (fabs(energy - ex) + 1e-10);   // :547-548
        }
    }
}

// ---- streaming selection: score in place -> exponent histogram -> gather the head ----------
// The top-k of a sweep never needs the full candidate list on the host side of the ABI:
// k_pt2_score writes every live candidate's score (importance, or max |c.H| in MAXABS mode)
// into the spare word of its pool entry and counts scores per binary exponent (2,048 bins,
// shared-memory histogram per CTA); k_pt2_threshold walks the bins from the top until k
// candidates are covered and steps one more bin down (covers the caller's relative tie band);
// k_pt2_gather compacts the candidates at or above that bound.  The caller then orders a few
// thousand candidates instead of running a top-k over tens of millions.
static const int PT2_BINS = 2048;      // sign + 11 exponent bits of a non-negative double

__global__ void __launch_bounds__(1024)
k_pt2_score(HamView H, bool have_h, unsigned tab_bytes, Pt2View W, i64 n_slots, double energy, bool fixed_point,
            unsigned* __restrict__ hist)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ unsigned s_hist[PT2_BINS];
    const bool staged = have_h && tab_bytes;
    for (int b = threadIdx.x; b < PT2_BINS; b += blockDim.x) s_hist[b] = 0;
    if (staged) {
        tma_stage_table(s_raw, H.hdiag, tab_bytes, &s_mbar);
        const double* s_tab = reinterpret_cast<const double*>(s_raw);
        const double* g0 = H.hdiag;
        H.nib_jk = s_tab + (H.nib_jk - g0);
        H.nib_jab = s_tab + (H.nib_jab - g0);
        H.hdiag = s_tab;
    }
    __syncthreads();
    i64 live_count = 0;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_slots; k += (i64)gridDim.x * blockDim.x) {
        ulonglong2* p = reinterpret_cast<ulonglong2*>(W.pool + 4 * k);
        const ulonglong2 d = p[0];
        if (d.x == FGK_EMPTY && d.y == FGK_EMPTY) continue;
        ulonglong2 acc = p[1];
        // the accumulator becomes {coupling (FP64), score}: the sweep is complete when it is scored
        const double cpl = fixed_point ? fx_to_double(acc.x, acc.y) : __longlong_as_double((long long)acc.x);
        acc.x = (u64)__double_as_longlong(cpl);
        double score = fabs(cpl);
        if (have_h) {
            fgk_det dd = {d.x, d.y};
            const double ex = staged ? diag_element(H, dd, LdsD())
                                     : diag_element(H, dd, [](const double* q) { return __ldg(q); });
            score = cpl * cpl / (fabs(energy - ex) + 1e-10);      // residual_expansion.py:547-548
        }
        acc.y = (u64)__double_as_longlong(score);
        p[1] = acc;
        atomicAdd(&s_hist[(unsigned)(acc.y >> 52) & (PT2_BINS - 1)], 1u);
        live_count++;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < PT2_BINS; b += blockDim.x)
        if (s_hist[b]) atomicAdd(hist + b, s_hist[b]);
    live_count = warp_sum_i64(live_count);
    if ((threadIdx.x & 31) == 0 && live_count) atomicAdd(W.counters + 3, (unsigned long long)live_count);
}

// thr[0] = lowest score bit pattern to keep, thr[1] = how many candidates that keeps
__global__ void k_pt2_threshold(const unsigned* __restrict__ hist, i64 k, unsigned long long* thr)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long cum = 0;
    int b = PT2_BINS - 1;
    for (; b >= 0; b--) {
        cum += hist[b];
        if ((i64)cum >= k) break;
    }
    if (b > 0) { b--; cum += hist[b]; }          // one bin of slack below the k-th score
    if (b < 0) b = 0;
    thr[0] = (unsigned long long)b << 52;
    thr[1] = cum;
}

__global__ void __launch_bounds__(256)
k_pt2_gather(Pt2View W, i64 n_slots, const unsigned long long* __restrict__ thr, fgk_det* __restrict__ out_dets,
             double* __restrict__ out_score, i64 cap, unsigned long long* cursor)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long lo = thr[0];
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 rounds = (n_slots + stride - 1) / stride;
    for (i64 it = 0; it < rounds; it++) {
        const i64 k = it * stride + (i64)blockIdx.x * blockDim.x + threadIdx.x;
        bool take = false;
        ulonglong2 d = make_ulonglong2(0, 0), acc = make_ulonglong2(0, 0);
        if (k < n_slots) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(W.pool + 4 * k);
            d = p[0];
            acc = p[1];
            take = !(d.x == FGK_EMPTY && d.y == FGK_EMPTY) && acc.y >= lo && !(acc.y >> 63);
        }
        const unsigned b = __ballot_sync(0xffffffffu, take);
        if (!b) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!take) continue;
        const i64 o = (i64)base + __popc(b & ((1u << lane) - 1u));
        if (o < cap) {
            reinterpret_cast<ulonglong2*>(out_dets)[o] = d;
            out_score[o] = __longlong_as_double((long long)acc.y);
        }
    }
}

extern "C" int fgk_pt2_score(fgk_ham_t h, fgk_pt2_t ws, int64_t n_slots, double energy, int64_t k,
                             uint32_t* hist, uint64_t* thr, int64_t* n_live, int64_t* n_keep, void* stream)
{
    if (!ws || !hist || !thr || k < 0) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_score: bad argument");
    if (n_slots < 0 || n_slots > ws->v.capacity) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_score: bad n_slots");
    if (n_live) *n_live = 0;
    if (n_keep) *n_keep = 0;
    if (n_slots == 0) return FGK_OK;
    FGK_CUDA(cudaSetDevice(ws->device));
    HamView hv;
    if (h) hv = h->v; else { hv = HamView(); }
    const unsigned tab_bytes = h ? h->dtab_bytes : 0;
    static bool attr_set[64] = {false};
    if (tab_bytes > 32 * 1024 && !attr_set[ws->device & 63]) {
        FGK_CUDA(cudaFuncSetAttribute(k_pt2_score, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        attr_set[ws->device & 63] = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FGK_CUDA(cudaMemsetAsync(ws->v.counters + 3, 0, sizeof(unsigned long long), st));
    FGK_CUDA(cudaMemsetAsync(hist, 0, PT2_BINS * sizeof(unsigned), st));
    i64 need = (n_slots + 1023) / 1024, cap = (i64)fgk_sm_count(ws->device) * 2;
    k_pt2_score<<<(int)(need < cap ? need : cap), 1024, tab_bytes, st>>>(
        hv, h != nullptr, tab_bytes, ws->v, n_slots, energy, ws->mode == FGK_PT2_SUM && !ws->scored, (unsigned*)hist);
    ws->scored = true;
    FGK_LAUNCH_CHECK();
    k_pt2_threshold<<<1, 32, 0, st>>>((const unsigned*)hist, k, (unsigned long long*)thr);
    FGK_LAUNCH_CHECK();
    unsigned long long host[3] = {0, 0, 0};
    FGK_CUDA(cudaMemcpyAsync(host, thr, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaMemcpyAsync(host + 2, ws->v.counters + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaStreamSynchronize(st));
    if (n_keep) *n_keep = (int64_t)host[1];
    if (n_live) *n_live = (int64_t)host[2];
    return FGK_OK;
}

extern "C" int fgk_pt2_gather(fgk_pt2_t ws, int64_t n_slots, const uint64_t* thr, uint64_t* out_dets,
                              double* out_score, int64_t out_cap, int64_t* n_written, void* stream)
{
    if (!ws || !thr) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_gather: bad argument");
    if (n_written) *n_written = 0;
    if (n_slots == 0 || out_cap == 0) return FGK_OK;
    if (n_slots < 0 || n_slots > ws->v.capacity || out_cap < 0 || !out_dets || !out_score)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_gather: bad argument");
    FGK_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    FGK_CUDA(cudaMemsetAsync(ws->v.counters + 3, 0, sizeof(unsigned long long), st));
    i64 need = (n_slots + 255) / 256, cap = (i64)fgk_sm_count(ws->device) * 8;
    k_pt2_gather<<<(int)(need < cap ? need : cap), 256, 0, st>>>(
        ws->v, n_slots, (const unsigned long long*)thr, (fgk_det*)out_dets, out_score, out_cap,
        ws->v.counters + 3);
    FGK_LAUNCH_CHECK();
    unsigned long long w = 0;
    FGK_CUDA(cudaMemcpyAsync(&w, ws->v.counters + 3, sizeof(w), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaStreamSynchronize(st));
    if ((int64_t)w > out_cap) return fgk_fail(FGK_ERR_CAPACITY, "fgk_pt2_gather: %lld candidates, room for %lld",
                                              (long long)w, (long long)out_cap);
    if (n_written) *n_written = (int64_t)w;
    return FGK_OK;
}

extern "C" int fgk_pt2_create(int64_t capacity, int64_t table_slots, uint64_t* table, uint64_t* pool,
                              uint64_t* counters, int device, fgk_pt2_t* out)
{
    if (!out || capacity < 1 || !table || !pool || !counters)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_create: bad argument");
    if (capacity >= (1ll << 32) - 4) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_pt2_create: capacity >= 2^32 - 4");
    if (table_slots < 2 || (table_slots & (table_slots - 1)) || table_slots < capacity)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_create: table_slots must be a power of two >= capacity");
    if ((uintptr_t)pool & 31) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_create: pool must be 32-byte aligned");
    fgk_pt2* P = new fgk_pt2();
    P->device = device;
    P->mode = -1;
    P->scored = false;
    P->v.capacity = capacity;
    P->v.mask = (u64)table_slots - 1;
    P->v.table = (u64*)table;
    P->v.pool = (u64*)pool;
    P->v.counters = (unsigned long long*)counters;
    P->v.region_bits = 0;
    P->v.region_mask = P->v.mask;
    P->v.queue_bits = 0;
    P->v.qstride = 0;
    P->v.qdets = nullptr; P->v.qvals = nullptr; P->v.qcursors = nullptr;
    *out = P;
    return FGK_OK;
}

extern "C" int fgk_pt2_set_partition(fgk_pt2_t ws, int region_bits, int queue_bits, int64_t queue_stride,
                                     uint64_t* queue_dets, double* queue_vals, uint64_t* queue_cursors)
{
    if (!ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_set_partition: null handle");
    if (queue_bits == 0) {          // back to the direct (unpartitioned) sweep
        ws->v.region_bits = 0; ws->v.region_mask = ws->v.mask;
        ws->v.queue_bits = 0; ws->v.qstride = 0;
        ws->v.qdets = nullptr; ws->v.qvals = nullptr; ws->v.qcursors = nullptr;
        return FGK_OK;
    }
    if (region_bits < 0 || queue_bits < region_bits || queue_bits > 20 || queue_stride < 1 ||
        !queue_dets || !queue_vals || !queue_cursors || ((uintptr_t)queue_dets & 15))
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_set_partition: bad argument");
    if (((ws->v.mask + 1) >> region_bits) < 64)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_set_partition: table regions would be smaller than 64 slots");
    ws->v.region_bits = region_bits;
    ws->v.region_mask = ((ws->v.mask + 1) >> region_bits) - 1;
    ws->v.queue_bits = queue_bits;
    ws->v.qstride = queue_stride;
    ws->v.qdets = (fgk_det*)queue_dets;
    ws->v.qvals = queue_vals;
    ws->v.qcursors = (unsigned long long*)queue_cursors;
    return FGK_OK;
}

extern "C" int fgk_pt2_destroy(fgk_pt2_t ws)
{
    delete ws;
    return FGK_OK;
}

extern "C" int fgk_pt2_reset(fgk_pt2_t ws, void* stream)
{
    if (!ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_reset: null handle");
    FGK_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    FGK_CUDA(cudaMemsetAsync(ws->v.table, 0xFF, (ws->v.mask + 1) * sizeof(u64), st));
    FGK_CUDA(cudaMemsetAsync(ws->v.counters, 0, 4 * sizeof(unsigned long long), st));
    if (ws->v.queue_bits)
        FGK_CUDA(cudaMemsetAsync(ws->v.qcursors, 0, sizeof(unsigned long long) << ws->v.queue_bits, st));
    ws->mode = -1;
    ws->scored = false;
    return FGK_OK;
}

extern "C" int fgk_pt2_accumulate(fgk_ham_t h, fgk_index_t idx, fgk_pt2_t ws, const int64_t* src_idx,
                                  const double* coeff, int64_t n_src, int mode, int n_pass,
                                  int pass_id, void* stream)
{
    if (!h || !idx || !ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_accumulate: null handle");
    if (n_src == 0) return FGK_OK;
    if (!src_idx || !coeff || n_src < 0 || n_pass < 1 || pass_id < 0 || pass_id >= n_pass)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_accumulate: bad argument");
    if (h->device != idx->device || h->device != ws->device)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_accumulate: device mismatch");
    if (mode != FGK_PT2_SUM && mode != FGK_PT2_MAXABS) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_accumulate: bad mode");
    if (ws->scored || (ws->mode >= 0 && ws->mode != mode))
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_accumulate: workspace already scored or used in the other mode "
                                     "(call fgk_pt2_reset first)");
    ws->mode = mode;
    FGK_CUDA(cudaSetDevice(h->device));
    const i64 resident_warps = (i64)fgk_sm_count(h->device) * 64;
    i64 n_split = (4 * resident_warps + n_src - 1) / n_src;     // >= 4 waves of work units
    if (n_split < 1) n_split = 1;
    if (n_split > 256) n_split = 256;
    {   // FGK_PT2_SPLIT: experiment knob (warps per source; more = fewer sources in flight at a time)
        static const char* e_split = getenv("FGK_PT2_SPLIT");
        if (e_split && atoi(e_split) > 0) n_split = atoi(e_split);
    }
    i64 need = (n_src * n_split + FGK_WARPS_PER_BLOCK - 1) / FGK_WARPS_PER_BLOCK;
    i64 cap = (i64)fgk_sm_count(h->device) * 8;
    int grid = (int)(need < cap ? need : cap);
    // the kernel is bound by random DRAM sectors (hash table): measured identical at 4, 6 and 8
    // resident CTAs per SM (1.9e9 candidates/s at config-5 shape); FGK_PT2_BPS selects.
    static int bps = 0;
    if (!bps) {
        const char* e = getenv("FGK_PT2_BPS");
        bps = e ? atoi(e) : 4;
        if (bps != 4 && bps != 6 && bps != 8) bps = 4;
    }
    cap = (i64)fgk_sm_count(h->device) * bps;
    grid = (int)(need < cap ? need : cap);
#define FGK_PT2_LAUNCH(B, Q)                                                                       \
    k_pt2_accumulate<B, Q><<<grid, FGK_BLOCK, 0, (cudaStream_t)stream>>>(                          \
        h->v, idx->v, ws->v, (const i64*)src_idx, coeff, n_src, (int)n_split, mode, (unsigned)n_pass, \
        (unsigned)pass_id)
    if (ws->v.queue_bits) {
        // partitioned sweep: enumerate -> append to the queue of the candidate's top hash bits,
        // then aggregate queue by queue (L2-resident table regions)
        FGK_PT2_LAUNCH(4, true);
        FGK_LAUNCH_CHECK();
        const i64 total = ((i64)1 << ws->v.queue_bits) * ws->v.qstride;
        const i64 per_block = 256 * 16;
        const i64 blocks = (total + per_block - 1) / per_block;
        if (blocks > 0x7fffffffll) return fgk_fail(FGK_ERR_UNSUPPORTED, "fgk_pt2_accumulate: queue too large");
        k_pt2_aggregate<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ws->v, mode, per_block);
        FGK_LAUNCH_CHECK();
        // the queues are transient: empty them so that a second accumulate call starts clean
        FGK_CUDA(cudaMemsetAsync(ws->v.qcursors, 0, sizeof(unsigned long long) << ws->v.queue_bits,
                                 (cudaStream_t)stream));
    }
    else {
        const int n = h->v.n_orb, na = h->v.n_alpha, nb = h->v.n_beta;
        int scap = na * (n - na) > nb * (n - nb) ? na * (n - na) : nb * (n - nb);
        if (scap < 1) scap = 1;
        const size_t smem = sizeof(unsigned) * 2 * (size_t)scap * FGK_WARPS_PER_BLOCK;
        static const bool per_candidate = getenv("FGK_PT2_PER_CANDIDATE_BUCKETS") != nullptr;
        if (!per_candidate) {
            // buckets by alpha string (see k_pt2_accumulate2)
            const unsigned tab_off = (unsigned)((smem + 127) & ~(size_t)127);
            const bool staged = h->itab_bytes && tab_off + h->itab_bytes <= 200u * 1024u;
            if (staged) {
                static bool attr = false;
                if (!attr) {
                    FGK_CUDA(cudaFuncSetAttribute(k_pt2_accumulate2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
                    attr = true;
                }
            } else if (smem > 40 * 1024)
                FGK_CUDA(cudaFuncSetAttribute(k_pt2_accumulate2<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            unsigned n_a = (unsigned)n_pass;                 // first level: largest divisor <= 64
            if (n_a > 64) {
                n_a = 1;
                for (unsigned dv = 64; dv >= 2; dv--)
                    if ((unsigned)n_pass % dv == 0) { n_a = dv; break; }
            }
            const unsigned n_d = (unsigned)n_pass / n_a;
            if (staged) {
                int per_sm = (int)((200u * 1024u) / (tab_off + h->itab_bytes + 4096u));
                per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
                const i64 cap2 = (i64)fgk_sm_count(h->device) * per_sm;
                const int grid2 = (int)(need < cap2 ? need : cap2);
                k_pt2_accumulate2<1, true><<<grid2, FGK_BLOCK, tab_off + h->itab_bytes, (cudaStream_t)stream>>>(
                    h->v, idx->v, ws->v, (const i64*)src_idx, coeff, n_src, (int)n_split, mode, n_a,
                    (unsigned)pass_id % n_a, n_d, (unsigned)pass_id / n_a, scap, h->itab_bytes, tab_off);
            } else
            k_pt2_accumulate2<4, false><<<grid, FGK_BLOCK, smem, (cudaStream_t)stream>>>(
                h->v, idx->v, ws->v, (const i64*)src_idx, coeff, n_src, (int)n_split, mode, n_a,
                (unsigned)pass_id % n_a, n_d, (unsigned)pass_id / n_a, scap, 0u, 0u);
        }
        else if (bps == 8) FGK_PT2_LAUNCH(8, false);
        else if (bps == 6) FGK_PT2_LAUNCH(6, false);
        else FGK_PT2_LAUNCH(4, false);
    }
#undef FGK_PT2_LAUNCH
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_pt2_merge(fgk_pt2_t ws, const uint64_t* dets, const double* vals, int64_t m,
                             int mode, void* stream)
{
    if (!ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_merge: null handle");
    if (m == 0) return FGK_OK;
    if (!dets || !vals || m < 0) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_merge: bad argument");
    if (mode != FGK_PT2_SUM && mode != FGK_PT2_MAXABS) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_merge: bad mode");
    if (ws->scored || (ws->mode >= 0 && ws->mode != mode))
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2_merge: workspace already scored or used in the other mode");
    ws->mode = mode;
    FGK_CUDA(cudaSetDevice(ws->device));
    i64 need = (m + 255) / 256, cap = (i64)fgk_sm_count(ws->device) * 8;
    k_pt2_merge<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(
        ws->v, (const fgk_det*)dets, vals, m, mode);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}

extern "C" int fgk_pt2_count(fgk_pt2_t ws, void* stream, int64_t* n_slots, int64_t* n_raw,
                             int* overflow)
{
    if (!ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_count: null handle");
    FGK_CUDA(cudaSetDevice(ws->device));
    unsigned long long hc[4];
    FGK_CUDA(cudaMemcpyAsync(hc, ws->v.counters, sizeof(hc), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FGK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    i64 used = (i64)hc[0] < ws->v.capacity ? (i64)hc[0] : ws->v.capacity;
    if (n_slots) *n_slots = used;
    if (n_raw) *n_raw = (i64)hc[1];
    if (overflow) *overflow = hc[2] ? 1 : 0;
    if (hc[2] == 2)
        return fgk_fail(FGK_ERR_ARG, "fgk_pt2: a coupling c_j <x|H|j> is not finite or exceeds 2^30");
    if (hc[2])
        return fgk_fail(FGK_ERR_CAPACITY, "fgk_pt2: candidate pool overflow (capacity %lld)",
                        (long long)ws->v.capacity);
    return FGK_OK;
}

extern "C" int fgk_pt2_export(fgk_ham_t h, fgk_pt2_t ws, int64_t n_slots, double energy,
                              uint64_t* out_dets, double* out_coupling, double* out_diag,
                              double* out_importance, int64_t* n_live, void* stream)
{
    if (!ws) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_export: null handle");
    if (n_slots == 0) { if (n_live) *n_live = 0; return FGK_OK; }
    if (n_slots < 0 || n_slots > ws->v.capacity) return fgk_fail(FGK_ERR_ARG, "fgk_pt2_export: bad n_slots");
    FGK_CUDA(cudaSetDevice(ws->device));
    HamView hv;
    if (h) hv = h->v; else { hv = HamView(); }
    const unsigned tab_bytes = h ? h->dtab_bytes : 0;
    static bool attr_set[64] = {false};
    if (tab_bytes > 48 * 1024 && !attr_set[ws->device & 63]) {
        FGK_CUDA(cudaFuncSetAttribute(k_pt2_export, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        attr_set[ws->device & 63] = true;
    }
    i64 need = (n_slots + 1023) / 1024, cap = (i64)fgk_sm_count(ws->device) * 2;
    cudaStream_t st = (cudaStream_t)stream;
    FGK_CUDA(cudaMemsetAsync(ws->v.counters + 3, 0, sizeof(unsigned long long), st));
    k_pt2_export<<<(int)(need < cap ? need : cap), 1024, tab_bytes, st>>>(
        hv, h != nullptr, tab_bytes, ws->v, n_slots, energy, ws->mode == FGK_PT2_SUM && !ws->scored,
        (fgk_det*)out_dets, out_coupling, out_diag,
        out_importance);
    FGK_LAUNCH_CHECK();
    unsigned long long live = 0;
    FGK_CUDA(cudaMemcpyAsync(&live, ws->v.counters + 3, sizeof(live), cudaMemcpyDeviceToHost, st));
    FGK_CUDA(cudaStreamSynchronize(st));
    if (n_live) *n_live = (int64_t)live;
    return FGK_OK;
}

// ---- dedup exchange: partition (determinant, value) pairs by owner rank ----------------------
// owner = (hash >> 24) % world (bits disjoint from the bucket-pass bits).  Count, then scatter
// into contiguous per-owner segments with warp-aggregated cursor claims; the segments are the
// send buffers of the all-to-all (dist.exchange_by_owner).
__device__ __forceinline__ int owner_of_det(u64 a, u64 b, int world)
{
    return (int)(((det_hash(a, b) >> 24) & 0xffffull) % (unsigned)world);
}

__global__ void __launch_bounds__(256)
k_partition(const fgk_det* __restrict__ dets, const double* __restrict__ vals, i64 m, int world,
            unsigned long long* cursors, fgk_det* __restrict__ out_dets, double* __restrict__ out_vals,
            bool scatter)
{
    const int lane = threadIdx.x & 31;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 rounds = (m + stride - 1) / stride;
    for (i64 it = 0; it < rounds; it++) {
        const i64 i = it * stride + (i64)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live = i < m;
        ulonglong2 d = make_ulonglong2(0, 0);
        int own = -1 - lane;                       // distinct dummy keys for idle lanes
        if (live) {
            d = __ldg(reinterpret_cast<const ulonglong2*>(dets) + i);
            own = owner_of_det(d.x, d.y, world);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, own);
        if (!live) continue;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(cursors + own, (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        if (scatter) {
            const i64 o = (i64)base + __popc(peers & ((1u << lane) - 1u));
            reinterpret_cast<ulonglong2*>(out_dets)[o] = d;
            out_vals[o] = __ldg(vals + i);
        }
    }
}

extern "C" int fgk_partition_by_owner(const uint64_t* dets, const double* vals, int64_t m, int world,
                                      uint64_t* cursors, uint64_t* out_dets, double* out_vals,
                                      int scatter, int device, void* stream)
{
    if (m == 0) return FGK_OK;
    if (!dets || !cursors || m < 0 || world < 1 || (scatter && (!vals || !out_dets || !out_vals)))
        return fgk_fail(FGK_ERR_ARG, "fgk_partition_by_owner: bad argument");
    FGK_CUDA(cudaSetDevice(device));
    i64 need = (m + 255) / 256, cap = (i64)fgk_sm_count(device) * 8;
    k_partition<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(
        (const fgk_det*)dets, vals, m, world, (unsigned long long*)cursors, (fgk_det*)out_dets, out_vals,
        scatter != 0);
    FGK_LAUNCH_CHECK();
    return FGK_OK;
}
