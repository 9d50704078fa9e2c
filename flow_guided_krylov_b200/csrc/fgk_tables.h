// fgk_tables.h -- host-side construction of the derived integral tables of a
// Hamiltonian handle (shared by fgk_ham_create and the CPU self-check).
// Mirrors reference molecular.py:57-117 (constructor + precomputations).
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>

struct HostTables {
    int n;
    std::vector<float> h1, g, w;          // float32 tables (molecular.py:68-69)
    std::vector<double> hdiag, jks, jab;  // FP64 diagonal tables built from the float32 values
    int nchunk;
    std::vector<double> nib_jk, nib_jab;  // nibble row-sum tables (n * nchunk * 16 each), see fgk_core.cuh
};

// h1_in (n,n), g_in (n,n,n,n) are the caller's FP64 integrals; the reference
// casts them with .float() and never looks at the FP64 values again.
inline void build_host_tables(const double* h1_in, const double* g_in, int n, HostTables& T)
{
    const size_t n2 = (size_t)n * n, n4 = n2 * n2;
    T.n = n;
    T.h1.resize(n2); T.g.resize(n4); T.w.resize(n4);
    T.hdiag.resize(n); T.jks.resize(n2); T.jab.resize(n2);
    for (size_t i = 0; i < n2; i++) T.h1[i] = (float)h1_in[i];
    for (size_t i = 0; i < n4; i++) T.g[i] = (float)g_in[i];
    auto G = [&](int p, int q, int r, int s) -> float {
        return T.g[(((size_t)p * n + q) * n + r) * n + s];
    };
    // same-spin value: numpy float32 subtraction h2e[p,q,r,s] - h2e[p,s,r,q]
    // (molecular.py:265,287), stored at w[p,q,r,s]
    for (int p = 0; p < n; p++)
        for (int q = 0; q < n; q++)
            for (int r = 0; r < n; r++)
                for (int s = 0; s < n; s++) {
                    volatile float d = G(p, q, r, s) - G(p, s, r, q);
                    T.w[(((size_t)p * n + q) * n + r) * n + s] = d;
                }
    for (int p = 0; p < n; p++) {
        T.hdiag[p] = (double)T.h1[(size_t)p * n + p];
        for (int q = 0; q < n; q++) {
            double Jpq = G(p, p, q, q), Jqp = G(q, q, p, p);   // molecular.py:94-97
            double Kpq = G(p, q, q, p), Kqp = G(q, p, p, q);   // molecular.py:99-103
            T.jks[(size_t)p * n + q] = 0.5 * (Jpq + Jqp) - 0.5 * (Kpq + Kqp);
            T.jab[(size_t)p * n + q] = Jpq;
        }
    }
    // nibble row sums: chunk c covers bits 4c..4c+3 of an occupation word, bit b <-> orbital n-1-b
    T.nchunk = (n + 3) / 4;
    T.nib_jk.assign((size_t)n * T.nchunk * 16, 0.0);
    T.nib_jab.assign((size_t)n * T.nchunk * 16, 0.0);
    for (int p = 0; p < n; p++)
        for (int c = 0; c < T.nchunk; c++)
            for (int v = 0; v < 16; v++) {
                double sjk = 0.0, sab = 0.0;
                for (int b = 0; b < 4; b++) {
                    int bit = 4 * c + b, q = n - 1 - bit;
                    if (!(v & (1 << b)) || q < 0) continue;
                    sjk += T.jks[(size_t)p * n + q];
                    sab += T.jab[(size_t)p * n + q];
                }
                T.nib_jk[((size_t)p * T.nchunk + c) * 16 + v] = sjk;
                T.nib_jab[((size_t)p * T.nchunk + c) * 16 + v] = sab;
            }
}
