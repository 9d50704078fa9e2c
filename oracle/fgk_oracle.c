/*
 * fgk_oracle.c -- CPU restatement of the Flow-Guided-Krylov determinant-space
 * Hamiltonian path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.  The product path
 * (flow-guided-krylov_b200/) never does; it fails loudly without its CUDA
 * library.
 *
 * Every function follows a stretch of the (pure Python) reference line by
 * line, working on unpacked 0/1 occupation bytes exactly as the reference
 * works on 0/1 numpy vectors -- deliberately NOT on packed words, so that the
 * CUDA bit tricks are checked against an independent formulation.
 *
 * Parity pin: tests/golden/ *.npz were produced by importing the reference
 * itself (tests/golden/make_golden.py); tests/test_oracle_golden.py checks
 * every function here against them.
 *
 * Integrals are the reference's float32 tables (molecular.py:68-69).  Off
 * diagonal elements are exact float32 numbers, as in the reference.  The
 * diagonal is restated in FP64 on those float32 tables (the reference's own
 * float32 einsum has no defined summation order; see DESIGN.md "parity
 * tiers").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n_orb;
    int n_single;        /* entries in the singles list                        */
    double e_nuc;
    const float *h1;     /* (n,n)      float32, molecular.py:68                */
    const float *g;      /* (n,n,n,n)  float32 chemist order, molecular.py:69  */
    float *J;            /* J[p,q] = g[p,p,q,q]          molecular.py:94-97    */
    float *K;            /* K[p,q] = g[p,q,q,p]          molecular.py:99-103   */
    int *sp, *sq;        /* singles list (p,q) row-major molecular.py:106-117  */
    float *sh;           /* h_pq of the singles list                           */
} orc_ham;

#define G4(H, p, q, r, s) \
    ((H)->g[(((size_t)(p) * (H)->n_orb + (q)) * (H)->n_orb + (r)) * (H)->n_orb + (s)])

/* the harness sets the thread count itself: torchrun exports OMP_NUM_THREADS=1 to its workers */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* molecular.py:57-117 : constructor + _precompute_vectorized_integrals +
 * _precompute_single_excitation_data.  h1/g are borrowed (caller keeps them). */
orc_ham *orc_ham_create(const float *h1, const float *g, int n_orb, double e_nuc)
{
    orc_ham *H = (orc_ham *)calloc(1, sizeof(orc_ham));
    int n = n_orb;
    H->n_orb = n;
    H->e_nuc = e_nuc;
    H->h1 = h1;
    H->g = g;
    H->J = (float *)malloc(sizeof(float) * n * n);
    H->K = (float *)malloc(sizeof(float) * n * n);
    for (int p = 0; p < n; p++)
        for (int q = 0; q < n; q++) {
            H->J[p * n + q] = G4(H, p, p, q, q);
            H->K[p * n + q] = G4(H, p, q, q, p);
        }
    H->sp = (int *)malloc(sizeof(int) * n * n);
    H->sq = (int *)malloc(sizeof(int) * n * n);
    H->sh = (float *)malloc(sizeof(float) * n * n);
    int m = 0;
    /* torch.nonzero is row-major; mask = |h1|>1e-12 & ~eye (molecular.py:106-109) */
    for (int p = 0; p < n; p++)
        for (int q = 0; q < n; q++) {
            if (p == q) continue;
            if (fabsf(h1[p * n + q]) > 1e-12f) {
                H->sp[m] = p;
                H->sq[m] = q;
                H->sh[m] = h1[p * n + q];
                m++;
            }
        }
    H->n_single = m;
    return H;
}

void orc_ham_destroy(orc_ham *H)
{
    if (!H) return;
    free(H->J); free(H->K); free(H->sp); free(H->sq); free(H->sh);
    free(H);
}

/* molecular.py:379-389  _jw_sign_np */
int orc_sign_single(const uint8_t *cfg, int p, int q)
{
    if (p == q) return 1;
    int low = p < q ? p : q, high = p < q ? q : p;
    int count = 0;
    for (int i = low + 1; i < high; i++) count += cfg[i];
    return (count & 1) ? -1 : 1;
}

static int prefix_sum(const uint8_t *cfg, int x)
{
    int c = 0;
    for (int i = 0; i < x; i++) c += cfg[i];
    return c;
}

/* molecular.py:391-423  _jw_sign_double_np, statement by statement */
int orc_sign_double(const uint8_t *cfg, int p, int r, int q, int s)
{
    int total = 0;
    total += prefix_sum(cfg, p);

    int count_r = prefix_sum(cfg, r);
    if (q < r) count_r -= cfg[q];
    total += count_r;

    int count_s = prefix_sum(cfg, s);
    if (p < s) count_s += 1;
    if (r < s) count_s += 1;
    if (q < s) count_s -= cfg[q];
    total += count_s;

    int count_q = prefix_sum(cfg, q);
    if (p < q) count_q += 1;
    if (r < q) count_q += 1;
    if (s < q) count_q -= cfg[s];
    total += count_q;

    /* (-1) ** int(total) ; total may be negative in principle, parity is what counts */
    return (((total % 2) + 2) % 2) ? -1 : 1;
}

/* molecular.py:133-184 diagonal_elements_batch, restated term by term in FP64
 * on the float32 tables (h1_diag, J_tensor, K_tensor). cfgs: (n, 2*n_orb) 0/1 */
void orc_diag(const orc_ham *H, const uint8_t *cfgs, long n, double *out)
{
    int no = H->n_orb, S = 2 * no;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < n; b++) {
        const uint8_t *na = cfgs + b * S, *nb = na + no;
        double e = H->e_nuc;
        double one = 0, aJa = 0, bJb = 0, aJb = 0, aKa = 0, bKb = 0;
        double dJa = 0, dJb = 0, dKa = 0, dKb = 0;
        for (int p = 0; p < no; p++) {
            one += (double)(na[p] + nb[p]) * (double)H->h1[p * no + p];
            dJa += na[p] * (double)H->J[p * no + p];
            dJb += nb[p] * (double)H->J[p * no + p];
            dKa += na[p] * (double)H->K[p * no + p];
            dKb += nb[p] * (double)H->K[p * no + p];
            for (int q = 0; q < no; q++) {
                double j = H->J[p * no + q], k = H->K[p * no + q];
                aJa += na[p] * j * na[q];
                bJb += nb[p] * j * nb[q];
                aJb += na[p] * j * nb[q];
                aKa += na[p] * k * na[q];
                bKb += nb[p] * k * nb[q];
            }
        }
        e += one;
        e += 0.5 * (aJa - dJa) + 0.5 * (bJb - dJb) + aJb;   /* :163-173 */
        e += -0.5 * (aKa - dKa) - 0.5 * (bKb - dKb);        /* :176-182 */
        out[b] = e;
    }
}

/* molecular.py:194-327 get_connections for ONE configuration, in the
 * reference's emission order.  out_cfg (cap, S) and out_el (cap) may be NULL to
 * count only.  Returns the number of connections (may exceed cap; only the
 * first cap are stored). */
long orc_connections(const orc_ham *H, const uint8_t *cfg, uint8_t *out_cfg,
                     float *out_el, long cap)
{
    int n = H->n_orb, S = 2 * n;
    int occ_a[64], occ_b[64], virt_a[64], virt_b[64];
    int noa = 0, nob = 0, nva = 0, nvb = 0;
    for (int i = 0; i < n; i++) {            /* np.where(...) :220-223 */
        if (cfg[i] == 1) occ_a[noa++] = i;
        if (cfg[i] == 0) virt_a[nva++] = i;
        if (cfg[n + i] == 1) occ_b[nob++] = i;
        if (cfg[n + i] == 0) virt_b[nvb++] = i;
    }
    long m = 0;
#define EMIT(i0, i1, j0, j1, val)                                   \
    do {                                                            \
        if (m < cap && out_cfg) {                                   \
            uint8_t *o = out_cfg + m * S;                           \
            memcpy(o, cfg, S);                                      \
            o[i0] = 0; if ((i1) >= 0) o[i1] = 0;                    \
            o[j0] = 1; if ((j1) >= 0) o[j1] = 1;                    \
            out_el[m] = (val);                                      \
        }                                                           \
        m++;                                                        \
    } while (0)

    /* singles :234-251 */
    for (int t = 0; t < H->n_single; t++) {
        int p = H->sp[t], q = H->sq[t];
        float h = H->sh[t];
        if (cfg[q] == 1 && cfg[p] == 0) {
            int sg = orc_sign_single(cfg, p, q);
            EMIT(q, -1, p, -1, (float)sg * h);
        }
        if (cfg[n + q] == 1 && cfg[n + p] == 0) {
            int sg = orc_sign_single(cfg, p + n, q + n);
            EMIT(q + n, -1, p + n, -1, (float)sg * h);
        }
    }
    /* alpha-alpha :254-274 */
    for (int i = 0; i < noa; i++) {
        int q = occ_a[i];
        for (int j = i + 1; j < noa; j++) {
            int s = occ_a[j];
            for (int k = 0; k < nva; k++) {
                int p = virt_a[k];
                for (int l = k + 1; l < nva; l++) {
                    int r = virt_a[l];
                    float val = G4(H, p, q, r, s) - G4(H, p, s, r, q);  /* float32 subtract */
                    if (fabsf(val) > 1e-12f) {
                        int sg = orc_sign_double(cfg, p, r, q, s);
                        EMIT(q, s, p, r, (float)sg * val);
                    }
                }
            }
        }
    }
    /* beta-beta :276-300 */
    for (int i = 0; i < nob; i++) {
        int q = occ_b[i];
        for (int j = i + 1; j < nob; j++) {
            int s = occ_b[j];
            for (int k = 0; k < nvb; k++) {
                int p = virt_b[k];
                for (int l = k + 1; l < nvb; l++) {
                    int r = virt_b[l];
                    float val = G4(H, p, q, r, s) - G4(H, p, s, r, q);
                    if (fabsf(val) > 1e-12f) {
                        int sg = orc_sign_double(cfg, p + n, r + n, q + n, s + n);
                        EMIT(q + n, s + n, p + n, r + n, (float)sg * val);
                    }
                }
            }
        }
    }
    /* alpha-beta :302-318 */
    for (int i = 0; i < noa; i++) {
        int q = occ_a[i];
        for (int j = 0; j < nob; j++) {
            int s = occ_b[j];
            for (int k = 0; k < nva; k++) {
                int p = virt_a[k];
                for (int l = 0; l < nvb; l++) {
                    int r = virt_b[l];
                    float val = G4(H, p, q, r, s);
                    if (fabsf(val) > 1e-12f) {
                        int sg = orc_sign_double(cfg, p, r + n, q, s + n);
                        EMIT(q, s + n, p, r + n, (float)sg * val);
                    }
                }
            }
        }
    }
#undef EMIT
    return m;
}

/* batch count helper (molecular.py:329-377 get_all_connections_with_indices) */
void orc_connections_count(const orc_ham *H, const uint8_t *cfgs, long n, int64_t *counts)
{
    int S = 2 * H->n_orb;
#pragma omp parallel for schedule(dynamic, 4)
    for (long j = 0; j < n; j++)
        counts[j] = orc_connections(H, cfgs + j * S, NULL, NULL, 0);
}

/* batch fill: offsets[j] = exclusive scan of counts */
void orc_connections_fill(const orc_ham *H, const uint8_t *cfgs, long n,
                          const int64_t *offsets, uint8_t *out_cfg, float *out_el,
                          int64_t *out_src)
{
    int S = 2 * H->n_orb;
#pragma omp parallel for schedule(dynamic, 4)
    for (long j = 0; j < n; j++) {
        long cnt = offsets[j + 1] - offsets[j];
        orc_connections(H, cfgs + j * S, out_cfg + offsets[j] * S, out_el + offsets[j], cnt);
        for (long k = 0; k < cnt; k++) out_src[offsets[j] + k] = j;
    }
}

/* ---- sorted lookup: memcmp order on 0/1 bytes == site-0-MSB key order
 * (molecular.py:498-501) ------------------------------------------------- */
typedef struct { const uint8_t *base; int S; } cmp_ctx;
static __thread cmp_ctx g_ctx;
static int cmp_idx(const void *a, const void *b)
{
    long ia = *(const long *)a, ib = *(const long *)b;
    int c = memcmp(g_ctx.base + ia * g_ctx.S, g_ctx.base + ib * g_ctx.S, g_ctx.S);
    if (c) return c;
    return (ia > ib) - (ia < ib);
}

static long *sorted_perm(const uint8_t *cfgs, long n, int S)
{
    long *perm = (long *)malloc(sizeof(long) * (n > 0 ? n : 1));
    for (long i = 0; i < n; i++) perm[i] = i;
    g_ctx.base = cfgs; g_ctx.S = S;
    qsort(perm, n, sizeof(long), cmp_idx);
    return perm;
}

/* The reference's dict maps key -> LAST index holding that key
 * ({config_ints[i]: i ...}, molecular.py:501).  Returns -1 if absent. */
static long lookup(const uint8_t *cfgs, const long *perm, long n, int S, const uint8_t *key)
{
    long lo = 0, hi = n;
    while (lo < hi) {                      /* upper bound */
        long mid = (lo + hi) / 2;
        if (memcmp(cfgs + perm[mid] * S, key, S) <= 0) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return -1;
    long cand = perm[lo - 1];
    return memcmp(cfgs + cand * S, key, S) == 0 ? cand : -1;
}

/* molecular.py:580-638 get_sparse_matrix_elements (COO of OFF-diagonal hits,
 * row = index of connected det, col = source j), in the reference's order
 * (j ascending, emission order inside j).  Two-phase: pass rows==NULL to count.
 * Also the loop of matrix_elements_fast (:504-514) and of
 * _build_subspace_hamiltonian (skqd.py:390-410). */
static long offdiag_coo_impl(const orc_ham *H, const uint8_t *cfgs, long n,
                             const int64_t *kets, long n_kets, int64_t *rows,
                             int64_t *cols, float *vals, long cap);

long orc_offdiag_coo(const orc_ham *H, const uint8_t *cfgs, long n, int64_t *rows,
                     int64_t *cols, float *vals, long cap)
{
    return offdiag_coo_impl(H, cfgs, n, NULL, n, rows, cols, vals, cap);
}

/* same loop restricted to the kets listed in `kets` (a bounded sample of a large
 * basis); cols[] then holds the POSITION in `kets`, rows[] the basis index */
long orc_offdiag_coo_kets(const orc_ham *H, const uint8_t *cfgs, long n, const int64_t *kets,
                          long n_kets, int64_t *rows, int64_t *cols, float *vals, long cap)
{
    return offdiag_coo_impl(H, cfgs, n, kets, n_kets, rows, cols, vals, cap);
}

static long offdiag_coo_impl(const orc_ham *H, const uint8_t *cfgs, long n,
                             const int64_t *kets, long n_kets, int64_t *rows,
                             int64_t *cols, float *vals, long cap)
{
    int S = 2 * H->n_orb;
    long *perm = sorted_perm(cfgs, n, S);
    long *cnt = (long *)calloc(n_kets + 1, sizeof(long));
#pragma omp parallel
    {
        long cap_c = 0; uint8_t *bc = NULL; float *be = NULL;
#pragma omp for schedule(dynamic, 4)
        for (long jj = 0; jj < n_kets; jj++) {
            long j = kets ? kets[jj] : jj;
            long m = orc_connections(H, cfgs + j * S, NULL, NULL, 0);
            if (m > cap_c) {
                cap_c = m; free(bc); free(be);
                bc = (uint8_t *)malloc((size_t)m * S); be = (float *)malloc(sizeof(float) * m);
            }
            orc_connections(H, cfgs + j * S, bc, be, m);
            long c = 0;
            for (long k = 0; k < m; k++)
                if (lookup(cfgs, perm, n, S, bc + k * S) >= 0) c++;
            cnt[jj + 1] = c;
        }
        free(bc); free(be);
    }
    for (long j = 0; j < n_kets; j++) cnt[j + 1] += cnt[j];
    long total = cnt[n_kets];
    if (rows && total <= cap) {
#pragma omp parallel
        {
            long cap_c = 0; uint8_t *bc = NULL; float *be = NULL;
#pragma omp for schedule(dynamic, 4)
            for (long jj = 0; jj < n_kets; jj++) {
                long j = kets ? kets[jj] : jj;
                long m = orc_connections(H, cfgs + j * S, NULL, NULL, 0);
                if (m > cap_c) {
                    cap_c = m; free(bc); free(be);
                    bc = (uint8_t *)malloc((size_t)m * S); be = (float *)malloc(sizeof(float) * m);
                }
                orc_connections(H, cfgs + j * S, bc, be, m);
                long o = cnt[jj];
                for (long k = 0; k < m; k++) {
                    long i = lookup(cfgs, perm, n, S, bc + k * S);
                    if (i >= 0) { rows[o] = i; cols[o] = jj; vals[o] = be[k]; o++; }
                }
            }
            free(bc); free(be);
        }
    }
    free(perm); free(cnt);
    return total;
}

/* residual_expansion.py:498-522, phase 1 of _find_important_configs.
 *   order[0..n_src)  : basis indices in processing order (argsort |c| desc,
 *                      already cut at |c|>1e-8, :486-490)
 *   coeff32          : float32 coefficients (:481)
 * Candidates are returned in first-seen (dict insertion) order with
 *   coup32 : the reference's float32 running sum (numpy>=2: python float *
 *            np.float32 -> np.float32, accumulated in processing order)
 *   coup64 : the same sum carried in FP64 (the 1e-9 value oracle)
 * raw_out: number of generated connections tested against the basis.
 * Returns the number of unique candidates (<= cap stored). */
typedef struct { long f, i; } orc_pr;
static int cmp_pr(const void *x, const void *y)
{
    long a = ((const orc_pr *)x)->f, b = ((const orc_pr *)y)->f;
    return (a > b) - (a < b);
}

long orc_pt2_candidates(const orc_ham *H, const uint8_t *cfgs, long n,
                        const int64_t *order, long n_src, const float *coeff32,
                        uint8_t *out_cfg, float *coup32, double *coup64, long cap,
                        int64_t *raw_out)
{
    int S = 2 * H->n_orb;
    long *perm = sorted_perm(cfgs, n, S);
    /* generate every outside-basis connection, tagged with its sequence number */
    long raw_cap = 1 << 16, nraw = 0, tested = 0;
    uint8_t *rc = (uint8_t *)malloc((size_t)raw_cap * S);
    float *rv = (float *)malloc(sizeof(float) * raw_cap);       /* coupling float32 */
    double *rd = (double *)malloc(sizeof(double) * raw_cap);    /* coupling FP64    */
    long cap_c = 0; uint8_t *bc = NULL; float *be = NULL;
    for (long t = 0; t < n_src; t++) {
        long j = order[t];
        float cj = coeff32[j];
        long m = orc_connections(H, cfgs + j * S, NULL, NULL, 0);
        if (m > cap_c) {
            cap_c = m; free(bc); free(be);
            bc = (uint8_t *)malloc((size_t)m * S); be = (float *)malloc(sizeof(float) * m);
        }
        orc_connections(H, cfgs + j * S, bc, be, m);
        tested += m;
        for (long k = 0; k < m; k++) {
            if (lookup(cfgs, perm, n, S, bc + k * S) >= 0) continue;
            if (nraw == raw_cap) {
                raw_cap *= 2;
                rc = (uint8_t *)realloc(rc, (size_t)raw_cap * S);
                rv = (float *)realloc(rv, sizeof(float) * raw_cap);
                rd = (double *)realloc(rd, sizeof(double) * raw_cap);
            }
            memcpy(rc + nraw * S, bc + k * S, S);
            /* coeffs[j].item() is a python float holding a float32 value; times
             * np.float32 -> np.float32 product (:515) */
            rv[nraw] = cj * be[k];
            rd[nraw] = (double)cj * (double)be[k];
            nraw++;
        }
    }
    free(bc); free(be);
    if (raw_out) *raw_out = tested;
    /* group equal keys, keeping sequence order inside a group */
    long *rp = sorted_perm(rc, nraw, S);       /* ties broken by index = sequence */
    long nuniq = 0;
    long *first = (long *)malloc(sizeof(long) * (nraw > 0 ? nraw : 1));
    float *s32 = (float *)malloc(sizeof(float) * (nraw > 0 ? nraw : 1));
    double *s64 = (double *)malloc(sizeof(double) * (nraw > 0 ? nraw : 1));
    for (long a = 0; a < nraw;) {
        long b = a;
        float acc32 = rv[rp[a]];
        double acc64 = rd[rp[a]];
        b++;
        while (b < nraw && memcmp(rc + rp[b] * S, rc + rp[a] * S, S) == 0) {
            acc32 = acc32 + rv[rp[b]];       /* old_coupling_sum + coupling (:520) */
            acc64 = acc64 + rd[rp[b]];
            b++;
        }
        first[nuniq] = rp[a]; s32[nuniq] = acc32; s64[nuniq] = acc64;
        nuniq++;
        a = b;
    }
    /* dict insertion order = ascending first-seen sequence */
    long *ord = (long *)malloc(sizeof(long) * (nuniq > 0 ? nuniq : 1));
    for (long i = 0; i < nuniq; i++) ord[i] = i;
    {
        orc_pr *ps = (orc_pr *)malloc(sizeof(orc_pr) * (nuniq > 0 ? nuniq : 1));
        for (long i = 0; i < nuniq; i++) { ps[i].f = first[i]; ps[i].i = i; }
        qsort(ps, nuniq, sizeof(orc_pr), cmp_pr);
        for (long i = 0; i < nuniq; i++) ord[i] = ps[i].i;
        free(ps);
    }
    for (long i = 0; i < nuniq && i < cap; i++) {
        long u = ord[i];
        if (out_cfg) memcpy(out_cfg + i * S, rc + first[u] * S, S);
        if (coup32) coup32[i] = s32[u];
        if (coup64) coup64[i] = s64[u];
    }
    free(ord); free(first); free(s32); free(s64); free(rp);
    free(rc); free(rv); free(rd); free(perm);
    return nuniq;
}

/* scipy.sparse csr_matvec (sparsetools/csr.h csr_matvec: y[i] += sum_k Ax[k]*x[Aj[k]]),
 * the product behind eigsh (skqd.py:784, residual_expansion.py:435) */
void orc_csr_matvec_f64(long n_row, const int64_t *indptr, const int32_t *indices,
                        const double *data, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_row; i++) {
        double sum = 0.0;
        for (int64_t k = indptr[i]; k < indptr[i + 1]; k++) sum += data[k] * x[indices[k]];
        y[i] = sum;
    }
}

/* the complex128 product behind expm_multiply (skqd.py:291-293); H real-valued
 * (stored complex128 by the reference with zero imaginary part), x complex
 * as interleaved (re,im). */
void orc_csr_matvec_z(long n_row, const int64_t *indptr, const int32_t *indices,
                      const double *data, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_row; i++) {
        double sr = 0.0, si = 0.0;
        for (int64_t k = indptr[i]; k < indptr[i + 1]; k++) {
            sr += data[k] * x[2 * (long)indices[k]];
            si += data[k] * x[2 * (long)indices[k] + 1];
        }
        y[2 * i] = sr; y[2 * i + 1] = si;
    }
}
