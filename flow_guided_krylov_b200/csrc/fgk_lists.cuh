// fgk_lists.cuh -- string-driven form of the projected-H rows (shared by fgk_projh4.cu and the
// CPU self-check fgk_hostcheck.cpp).
//
// For every distinct alpha / beta string of a basis the strings of the SAME basis that differ
// from it by one or two orbitals are listed once ("single / double replacement lists", the
// classic determinant-CI organisation), each with the matrix element factors that depend on
// the string pair only:
//   singles : rank of the target string, signed values <i|H|j>, <j|H|i> of the one-body
//             connection (molecular.py:234-251: sign1 * h_pq), and the separable alpha-beta
//             factors of fgk_core.cuh (table offsets e*n+h / h*n+e, parities pk / pb);
//   doubles : rank, signed same-spin values (molecular.py:254-300: sign2 * fp32(g - g)).
// A row of the projected H is then assembled from list entries and the (alpha rank, beta rank)
// pair table with no bit manipulation at all:
//   alpha singles  x own beta   |  own alpha x beta singles   |  alpha doubles x own beta
//   own alpha x beta doubles    |  alpha singles x beta singles (g[offa n^2 + offb], parity xor)
// Values and filters are exactly those of k_projh3 / the reference (|value| > 1e-12 on the
// float32 table value; symmetrised flavour 0.5 (<i|H|j> + <j|H|i>) in FP64).
#pragma once
#include "fgk_core.cuh"

struct alignas(16) LEntry {
    int rank;        // rank of the target string in the ascending distinct-string list
    float vij;       // signed <i|H|j>  (j = determinant with the target string; ket j)
    float vji;       // signed <j|H|i>
    unsigned info;   // singles: offk | offb << 12 | pk << 24 | pb << 25 ; doubles: 0
};

template <class Ld>
FGK_HD LEntry single_entry(const HamView& H, u64 w, u64 w2, int rank, Ld ldf)
{
    const int n = H.n_orb;
    int hh, ee;
    single_from_strings(w, w2, n, hh, ee);
    unsigned pk, pb, sk, sb;
    single_factors(w, w2, n, hh, ee, pk, pb, sk, sb);
    const float rb = ldf(H.h1 + hh * n + ee), rk = ldf(H.h1 + ee * n + hh);
    LEntry e;
    e.rank = rank;
    e.vij = sb ? -rb : rb;
    e.vji = sk ? -rk : rk;
    e.info = (unsigned)(ee * n + hh) | ((unsigned)(hh * n + ee) << 12) | (pk << 24) | (pb << 25);
    return e;
}

template <class Ld>
FGK_HD LEntry double_entry(const HamView& H, u64 w, u64 w2, int rank, Ld ldf)
{
    const int n = H.n_orb;
    int h0, h1, e0, e1;
    double_from_strings(w, w2, n, h0, h1, e0, e1);
    const float rb = ldf(H.w + idx4(n, h0, e0, h1, e1)), rk = ldf(H.w + idx4(n, e0, h0, e1, h1));
    Excitation x, rx;
    x.cls = rx.cls = 2;                 // the same-spin parity only reads the changed word
    x.h0 = h0; x.h1 = h1; x.e0 = e0; x.e1 = e1;
    rx.h0 = e0; rx.h1 = e1; rx.e0 = h0; rx.e1 = h1;
    const fgk_det kd = {w, 0}, ko = {w2, 0};
    const int park = exc_parity_ket(kd, n, x), parb = exc_parity_ket(ko, n, rx);
    LEntry e;
    e.rank = rank;
    e.vij = parb ? -rb : rb;
    e.vji = park ? -rk : rk;
    e.info = 0u;
    return e;
}

// reference filter + flavour: false if the entry is not stored
FGK_HD bool entry_value(bool sym, bool drop0, float vij, float vji, double& v)
{
    const bool kij = (vij < 0.f ? -vij : vij) > 1e-12f;
    const bool kji = sym && (vji < 0.f ? -vji : vji) > 1e-12f;
    if (!kij && !kji) return false;
    const float a = kij ? vij : 0.f, b = kji ? vji : 0.f;
    v = sym ? 0.5 * ((double)a + (double)b) : (double)a;
    return !(drop0 && v == 0.0);
}

// the same decision with the value as the float32 number the packed operator stores: when both
// directions agree (symmetric integrals: always, except the F3 sign cancellations) the average IS
// that float32 number and no FP64 arithmetic is needed; `exact` = false if 0.5 (vij + vji) is not
// a float32 number (v then carries the FP64 value)
FGK_HD bool entry_value_f32(bool sym, bool drop0, float vij, float vji, float& f, double& v, bool& exact)
{
    const bool kij = (vij < 0.f ? -vij : vij) > 1e-12f;
    const bool kji = sym && (vji < 0.f ? -vji : vji) > 1e-12f;
    if (!kij && !kji) return false;
    const float a = kij ? vij : 0.f, b = kji ? vji : 0.f;
    exact = true;
    if (!sym || a == b) {
        f = a;
        v = (double)a;
    } else {
        v = 0.5 * ((double)a + (double)b);
        f = (float)v;
        exact = (double)f == v;
    }
    return !(drop0 && v == 0.0);
}

// alpha-beta double built from one alpha single and one beta single (molecular.py:302-318)
template <class Ld>
FGK_HD void ab_values(const HamView& H, const LEntry& ea, const LEntry& eb, bool sym, Ld ldf, float& vij, float& vji)
{
    const int n2 = H.n_orb * H.n_orb;
    const float rb = ldf(H.g + (size_t)((ea.info >> 12) & 0xfffu) * n2 + ((eb.info >> 12) & 0xfffu));
    const float rk = sym ? ldf(H.g + (size_t)(ea.info & 0xfffu) * n2 + (eb.info & 0xfffu)) : 0.f;
    const unsigned px = ea.info ^ eb.info;
    vij = (((px >> 25) ^ 1u) & 1u) ? -rb : rb;          // parity (bra side) = pb_a ^ pb_b ^ 1
    vji = (((px >> 24) ^ 1u) & 1u) ? -rk : rk;          // parity (ket side) = pk_a ^ pk_b ^ 1
}
