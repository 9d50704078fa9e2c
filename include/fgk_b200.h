/*
 * fgk_b200.h -- C ABI of the B200-native determinant-space Hamiltonian engine.
 *
 * This is the drop-in boundary for the one hot path of
 * George930502/Flow-Guided-Krylov.  The reference is pure
 * Python and has no FFI of its own; each entry point below names the reference
 * method (file:line under reference/src) whose work it takes over.  The
 * reference-side binding (ctypes, because the reference host language is
 * Python) is shown in INTEGRATION.md and implemented in
 * flow_guided_krylov_b200/_native.py.
 *
 * Rules of the boundary
 *   - plain C: pointers, sizes, ints.  No torch/C++ types.
 *   - every function returns 0 (FGK_OK) or a negative error code;
 *     fgk_last_error() gives the thread-local message.  No exceptions cross.
 *   - the CALLER allocates every buffer (device memory unless a parameter is
 *     documented "host"); variable-size outputs are two-phase: count -> fill.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is enqueued on it and the call returns without synchronising
 *     unless documented otherwise.
 *   - handles are opaque and immutable after creation: concurrent calls on
 *     different streams are safe (the reference calls get_connections from an
 *     8-thread pool, molecular.py:554).
 *   - a determinant is two uint64 words {alpha, beta}; orbital p of a spin block
 *     is bit (n_orb-1-p).  Sorting (alpha, beta) as a 128-bit unsigned number
 *     reproduces the reference's site-0-is-MSB key order (molecular.py:498-500).
 *     n_orb <= 64.  Arrays of determinants are uint64[n][2], 16-byte aligned.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     FGK_ERR_CUDA.
 */
#ifndef FGK_B200_H
#define FGK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGK_OK 0
#define FGK_ERR_ARG (-1)         /* bad argument                                   */
#define FGK_ERR_CUDA (-2)        /* CUDA runtime error (message has the detail)    */
#define FGK_ERR_CAPACITY (-3)    /* a caller-sized workspace overflowed            */
#define FGK_ERR_UNSUPPORTED (-4) /* e.g. n_orb > 64                                */

/* projected-H flavours */
#define FGK_H_RAW 0        /* <i|H|j> exactly as get_connections(j) reports it      */
#define FGK_H_SYM 1        /* 0.5*(<i|H|j> + <j|H|i>)  (residual_expansion.py:425,  */
                           /*  skqd.py:725, molecular.py:921)                       */
#define FGK_H_DROP_ZEROS 2 /* OR-able: skip entries whose value is exactly 0.0      */
                           /*  (what scipy csr_matrix(dense) does, skqd.py:783)     */
#define FGK_H_FLAT_WALK 4  /* OR-able: build rows by the flat reference-order walk  */
                           /*  over every excitation instead of the string-set      */
                           /*  driven builder (same entries; cross-check / debug)   */
#define FGK_H_HASH_WALK 8  /* OR-able: string-set driven builder with a hash probe  */
                           /*  per candidate, instead of the rank-based builder     */
                           /*  (same entries; cross-check, and the automatic path   */
                           /*  when a basis has more distinct strings than a row    */
                           /*  has same-spin excitations)                           */

/* PT2 accumulation flavours */
#define FGK_PT2_SUM 0      /* signed coupling sum   (residual_expansion.py:515-520) */
#define FGK_PT2_MAXABS 1   /* max |c_j * elem|      (residual_expansion.py:218-236) */

typedef struct fgk_ham* fgk_ham_t;
typedef struct fgk_index* fgk_index_t;
typedef struct fgk_pt2* fgk_pt2_t;
typedef struct fgk_strlists* fgk_strlists_t;

int fgk_version(void);
const char* fgk_last_error(void);
/* sm count, L2 bytes, free and total HBM bytes of `device` (host outputs) */
int fgk_device_info(int device, int* sm_count, size_t* l2_bytes, size_t* free_bytes,
                    size_t* total_bytes);

/* ---- Hamiltonian handle ------------------------------------------------------
 * replaces MolecularHamiltonian.__init__ + _precompute_vectorized_integrals +
 * _precompute_single_excitation_data (molecular.py:57-117).
 * h1 (n,n) and g (n,n,n,n), chemist order (pq|rs), are HOST FP64 arrays; they are
 * rounded to float32 exactly like the reference's .float() (molecular.py:68-69). */
int fgk_ham_create(const double* h1_host, const double* g_host, int n_orb, int n_alpha,
                   int n_beta, double e_nuc, int device, fgk_ham_t* out);
int fgk_ham_destroy(fgk_ham_t h);

/* ---- K1 pack / unpack --------------------------------------------------------
 * replaces the integer-key encoders molecular.py:498-500,609-611 and
 * connection_cache.py:82-98.  cfg is (n, 2*n_orb) int64 0/1, row-major. */
int fgk_pack_i64(const int64_t* cfg, int64_t n, int n_orb, uint64_t* dets, int device, void* stream);
int fgk_unpack_i64(const uint64_t* dets, int64_t n, int n_orb, int64_t* cfg, int device, void* stream);

/* ---- K2 diagonal ---------------------------------------------------------------
 * replaces diagonal_elements_batch / diagonal_element (molecular.py:133-192).
 * FP64 arithmetic on the float32-rounded tables. */
int fgk_diag(fgk_ham_t h, const uint64_t* dets, int64_t n, double* out, void* stream);

/* ---- K3 connections, reference emission order ---------------------------------
 * replaces get_connections (molecular.py:194-327) and its batch wrappers
 * get_all_connections_with_indices (:329-377), get_connections_parallel (:518-578)
 * and the unimplemented get_connections_batch hook (utils/connection_cache.py:250).
 * counts[j] = number of connections of dets[j].  offsets = exclusive scan of counts
 * (n+1 entries).  Outputs: out_dets (total,2) u64, out_elems float32 (exact
 * reference values), out_src int64 (index j of the source), any of which may be NULL. */
int fgk_conn_count(fgk_ham_t h, const uint64_t* dets, int64_t n, int64_t* counts, void* stream);
int fgk_conn_fill(fgk_ham_t h, const uint64_t* dets, int64_t n, const int64_t* offsets,
                  uint64_t* out_dets, float* out_elems, int64_t* out_src, void* stream);

/* ---- K4 basis index -------------------------------------------------------------
 * replaces the Python dict / set lookups molecular.py:501,512;
 * residual_expansion.py:445-449,513; skqd.py:171-175,405-407.
 * `dets` must stay alive and unchanged while the index is used (the table stores
 * indices into it).  Duplicate determinants resolve to the LAST index, like the
 * reference's dict comprehension.  Synchronises `stream` once (to size the
 * alpha/beta string sets); the index is complete in stream order on `stream`.  Its device
 * buffers come from a stream-ordered memory pool owned by the library (the device's default
 * pool is not touched); fgk_index_destroy returns them in the order of the creation stream,
 * without a device-wide synchronisation.  On error nothing is leaked. */
int fgk_index_create(const uint64_t* dets, int64_t n, int device, void* stream, fgk_index_t* out);
int fgk_index_destroy(fgk_index_t idx);
/* out_idx[k] = position of query[k] in the basis, or -1 */
int fgk_index_lookup(fgk_index_t idx, const uint64_t* query, int64_t m, int32_t* out_idx, void* stream);
/* number of distinct alpha / beta strings of the basis (host outputs) */
int fgk_index_info(fgk_index_t idx, int64_t* n_dets, int64_t* n_alpha_strings, int64_t* n_beta_strings);
/* *dense_pairs (host) = 1 if the index also holds the dense (alpha rank, beta rank) -> position
 * table, which it builds when the basis covers at least 1/16 of its alpha x beta string product
 * (or that product is below 2^20); the projected-H builder then needs no hash probe at all. */
int fgk_index_layout(fgk_index_t idx, int* dense_pairs);

/* ---- K5 projected Hamiltonian (CSR rows) ------------------------------------------
 * replaces matrix_elements_fast (molecular.py:471-516), get_sparse_matrix_elements
 * (:580-638) and _build_subspace_hamiltonian (skqd.py:374-419).
 * Rows [row_begin,row_end) of H over the indexed basis; entry (i,j) = <i|H|j>;
 * the diagonal is always stored first in its row.  counts has row_end-row_begin
 * entries; row_ptr = exclusive scan (local, starts at 0).  cols are GLOBAL basis
 * indices.  Entries of a row are in enumeration order; fgk_csr_sort_rows orders
 * them by column. */
int fgk_projh_count(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end, int mode,
                    int64_t* counts, void* stream);
int fgk_projh_fill(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end, int mode,
                   const int64_t* row_ptr, int32_t* cols, double* vals, void* stream);
int fgk_csr_sort_rows(int64_t n_rows, const int64_t* row_ptr, int32_t* cols, double* vals,
                      int device, void* stream);
/* Same fill, but straight into SELL-32 storage (layout under K6 below; no CSR copy is made):
 * slice_ptr from the row counts as for fgk_sell_fill; sell_cols / sell_vals must be
 * zero-filled by the caller. */
int fgk_projh_fill_sell(fgk_ham_t h, fgk_index_t idx, int64_t row_begin, int64_t row_end, int mode,
                        const int64_t* slice_ptr, int32_t* sell_cols, double* sell_vals, void* stream);

/* ---- K5b projected H straight into the packed SELL-32 operator (8 B/nnz) --------------------
 * Same matrix as fgk_projh_* (molecular.py:471-516, skqd.py:374-419) but assembled string-driven,
 * one warp per 32-row slice with lane = row, directly in the layout fgk_spmv_sell_f32_* read
 * (see fgk_sell_pack_f32): no CSR arrays, no CSR -> SELL pass, no re-pack.
 * fgk_strlists_create builds, once per (Hamiltonian, index) pair, the single / double replacement
 * lists of the distinct alpha / beta strings of the basis (synchronises `stream` twice to size them;
 * memory from the library's pool).
 * Row lengths count OFF-DIAGONAL entries (the diagonal is the separate FP64 array fgk_diag
 * fills).  fgk_projh_packed_bound: upper bound from the list lengths alone (exact for product
 * bases with dense integrals).  fgk_projh_packed_count: exact lengths by the same walk, for
 * the slices s = 0, slice_stride, 2 slice_stride, ... only (other rows' counts are untouched):
 * stride 1 = every row; a larger stride is a cheap sample to check the bound against.
 * fgk_projh_packed_fill: slice_ptr (16-byte units) from the caller's row lengths (slice width
 * = ceil(max length in the slice / 2) pair-columns); writes every unit of every slice (padding
 * included), row_len[r] = entries written, *flag = 1 if a value was not float32-exact
 * (non-symmetric integrals with FGK_H_SYM) or a slice was too narrow -- rebuild through
 * fgk_projh_count / fgk_projh_fill then. */
int fgk_strlists_create(fgk_ham_t h, fgk_index_t idx, void* stream, fgk_strlists_t* out);
int fgk_strlists_destroy(fgk_strlists_t lists);
int fgk_strlists_info(fgk_strlists_t lists, int64_t* n_single_alpha, int64_t* n_double_alpha,
                      int64_t* n_single_beta, int64_t* n_double_beta);
int fgk_projh_packed_bound(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t lists, int64_t row_begin,
                           int64_t row_end, int64_t* counts, void* stream);
int fgk_projh_packed_count(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t lists, int64_t row_begin,
                           int64_t row_end, int mode, int64_t slice_stride, int64_t* counts, void* stream);
int fgk_projh_packed_fill(fgk_ham_t h, fgk_index_t idx, fgk_strlists_t lists, int64_t row_begin,
                          int64_t row_end, int mode, const int64_t* slice_ptr, void* packed,
                          int32_t* row_len, int* flag, void* stream);

/* ---- K6 sparse H.v, FP64 CSR ----------------------------------------------------------
 * replaces scipy's csr_matvec inside eigsh (skqd.py:784, residual_expansion.py:435,
 * molecular.py:936) and inside expm_multiply (skqd.py:291-293).
 * y[i] = sum_k vals[k] * x[cols[k]].  _z: x, y complex128 interleaved (re,im), H real. */
int fgk_spmv_f64(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols, const double* vals,
                 const double* x, double* y, int device, void* stream);
int fgk_spmv_z(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols, const double* vals,
               const double* x, double* y, int device, void* stream);

/* SELL-32 flavour of the same product (sliced ELLPACK, slice height 32, entries
 * stored in lane-interleaved pairs): entry k of row r = 32 s + l sits at
 *   slice_ptr[s] + ((k >> 1) * 32 + l) * 2 + (k & 1).
 * slice_ptr (n_slices + 1 entries, in elements) is sized by the caller:
 * width_s = max row length in slice s rounded up to even, slice_ptr = scan(32 * width_s).
 * fgk_sell_fill transposes CSR rows into it (padding: value 0, column 0). */
int fgk_sell_fill(int64_t n_rows, const int64_t* row_ptr, const int32_t* cols, const double* vals,
                  const int64_t* slice_ptr, int32_t* sell_cols, double* sell_vals, int device,
                  void* stream);
int fgk_spmv_sell_f64(int64_t n_rows, const int64_t* slice_ptr, const int32_t* sell_cols,
                      const double* sell_vals, const double* x, double* y, int device, void* stream);
int fgk_spmv_sell_z(int64_t n_rows, const int64_t* slice_ptr, const int32_t* sell_cols,
                    const double* sell_vals, const double* x, double* y, int device, void* stream);

/* Packed SELL-32: exact-float32 off-diagonal values + FP64 diagonal (8 B per nonzero instead
 * of 12; FP64 arithmetic on the same numbers, hence bit-identical products).  One 16-byte unit
 * {float v0, float v1, int32 c0, int32 c1} per lane per pair-column; slice_ptr counts 16-byte
 * units and is sized from (row length - 1) rounded up to even; the diagonal entry of row r
 * (column row_offset + r) goes to diag[r].  *inexact_flag (device int, zeroed by the caller)
 * is set if some off-diagonal value is not exactly representable in float32 -- then the
 * packed copy must not be used. */
int fgk_sell_pack_f32(int64_t n_rows, int64_t row_offset, const int64_t* row_ptr, const int32_t* cols,
                      const double* vals, const int64_t* slice_ptr, void* packed, double* diag,
                      int* inexact_flag, int device, void* stream);
int fgk_spmv_sell_f32_f64(int64_t n_rows, int64_t row_offset, const int64_t* slice_ptr,
                          const void* packed, const double* diag, const double* x, double* y,
                          int device, void* stream);
int fgk_spmv_sell_f32_z(int64_t n_rows, int64_t row_offset, const int64_t* slice_ptr,
                        const void* packed, const double* diag, const double* x, double* y,
                        int device, void* stream);

/* ---- multi-GPU H.v with the all-gather fused into the product (one process per GPU) ----
 * fgk_peer_alloc: cudaMalloc'ed, zeroed buffer + its 64-byte cudaIpc handle (exchange the
 * handles with any host-side collective); fgk_peer_open maps a peer's buffer into this
 * process.  Peer pointer arguments (peer_out, peer_flags, ...) are HOST arrays of `world`
 * device pointers, the rank's own buffer included.
 * fgk_peer_barrier: stand-alone flag barrier over peer-mapped arrays (peer_flags[p] = rank p's
 * uint64[world] array); epoch must increase by one per synchronising call (fgk_peer_barrier /
 * _step / _gather / _allreduce_sum share the flags); *err_flag (device) is set to the epoch if a
 * peer never arrives (bounded spin). */
int fgk_peer_alloc(size_t bytes, int device, void** dev_ptr, unsigned char* handle64);
int fgk_peer_open(const unsigned char* handle64, int device, void** dev_ptr);
int fgk_peer_close(void* dev_ptr, int device);
int fgk_peer_free(void* dev_ptr, int device);
int fgk_peer_barrier(uint64_t* const* peer_flags, int rank, int world, uint64_t epoch,
                     uint64_t* err_flag, int device, void* stream);

/* One-launch multi-GPU step: product + broadcast + barrier (no reference counterpart).
 * y = H[row block] x with the rank's rows in SELL-32 storage -- FP64 (cols_or_packed = int32
 * columns, vals) or packed exact-float32 (cols_or_packed = the uint4 array of fgk_sell_pack_f32,
 * diag; flag FGK_PEER_PACKED_F32) -- for a real or complex (FGK_PEER_COMPLEX, interleaved) x.
 * Every y_r is stored into peer_out[p][row_offset + r] on every rank p; the CTA that finishes
 * last arrives on every rank's flag array (peer_flags as for fgk_peer_barrier) with `epoch` and
 * waits for all peers: when the kernel completes the output vector is complete on this rank and
 * every peer has finished reading x.  done_counter: device uint32, zero before the first call.
 * Every rank must own at least one row.  x must not alias the outputs (ping-pong buffers). */
#define FGK_PEER_COMPLEX 1
#define FGK_PEER_PACKED_F32 2
int fgk_peer_step(int64_t n_rows, const int64_t* slice_ptr, const void* cols_or_packed,
                  const double* vals, const double* diag, const double* x, double* const* peer_out,
                  int flags, int64_t row_offset, uint64_t* const* peer_flags, int rank, int world,
                  uint64_t epoch, uint32_t* done_counter, uint64_t* err_flag, int device, void* stream);
/* All-gather of a row-sharded vector over peer memory: src_local[0 .. n_bytes) goes to
 * peer_dst[p] + dst_offset_bytes on every rank p (8-byte multiples), closed by the same
 * barrier.  Distributes the INPUT of a product: row-sharded Krylov vectors, host vectors
 * uploaded as N slices. */
int fgk_peer_gather(const void* src_local, int64_t n_bytes, void* const* peer_dst, int64_t dst_offset_bytes,
                    uint64_t* const* peer_flags, int rank, int world, uint64_t epoch,
                    uint32_t* done_counter, uint64_t* err_flag, int device, void* stream);

/* Fused vector algebra of one block-Davidson iteration (replaces np.linalg.eigh / eigsh of
 * residual_expansion.py:408-443, skqd.py:754-796 together with solvers.py).  V, W: the basis
 * vectors and their images as rows (m_max x ld, ld >= n_local).  mode 0: t = (W^T s - theta V^T s)
 * / (theta - diag) with s = host_coef[0..m), partial[cta][j] = V_j . t, partial[cta][m] = r . r;
 * mode 1: t -= V^T c (c = dev_coef), partial = {V_j . t, t . t}; mode 2: out = (t - V^T c) /
 * sqrt(*tt_dev - c . c), the norm to *nrm_out; mode 3: partial[cta][j] = V_j . w.  partial has
 * n_blocks rows of m + 1 doubles (the caller adds the rows in order, then across ranks). */
int fgk_davidson_step(int mode, int64_t n_local, int64_t ld, int m, const double* V, const double* W,
                      const double* host_coef, const double* dev_coef, const double* tt_dev, double theta,
                      const double* diag, double* t, const double* w, double* out, double* partial,
                      int n_blocks, double* nrm_out, int device, void* stream);

/* Host-buffer form of the multi-GPU step (the e2e call): x_host_slice = this rank's n_rows
 * entries of x in (pinned) host memory; copied to x_dev_slice, exchanged with fgk_peer_gather into
 * every rank's CURRENT vector (peer_cur), one fgk_peer_step into peer_next, the rank's rows of the
 * result copied to y_host; synchronises the stream.  Uses epochs `epoch` and `epoch + 1`. */
int fgk_peer_matvec_host(int64_t n_rows, const int64_t* slice_ptr, const void* cols_or_packed,
                         const double* vals, const double* diag, const void* x_host_slice, void* x_dev_slice,
                         double* const* peer_cur, double* const* peer_next, void* y_host, int flags,
                         int64_t row_offset, uint64_t* const* peer_flags, int rank, int world, uint64_t epoch,
                         uint32_t* done_counter, uint64_t* err_flag, int device, void* stream);

/* Small all-reduce (sum) of n <= slot_stride doubles over peer memory: the dot products of a
 * row-sharded Krylov iteration.  src holds src_rows rows of n doubles (the per-CTA partial rows of
 * fgk_davidson_step; 1 for a plain vector), added in row order first.  peer_scratch[p]: rank p's
 * scratch, 2 * world * slot_stride doubles; area alternates 0 / 1 between consecutive calls.
 * Partials are added in rank order: identical bits on every rank.  dst (n doubles) may alias row 0. */
int fgk_peer_allreduce_sum(const double* src, int64_t n, int64_t src_rows, double* dst, double* const* peer_scratch,
                           int64_t slot_stride, int area, uint64_t* const* peer_flags, int rank, int world,
                           uint64_t epoch, uint32_t* done_counter, uint64_t* err_flag, int device, void* stream);

/* One Taylor term of exp(t (H - mu I)) psi (scipy expm_multiply under skqd.py:291-293), complex128
 * vectors of length n: B <- c (y - mu B) with y = H B, F <- F + B; norms[0] = max |B_i|,
 * norms[1] = max |F_i| (device double[2], written by the call). */
int fgk_taylor_update_z(int64_t n, const double* y, double* B, double* F, double mu, double c_re,
                        double c_im, double* norms, int device, void* stream);

/* ---- K7/K8 PT2 residual expansion --------------------------------------------------------
 * replaces SelectedCIExpander._find_important_configs (residual_expansion.py:451-554)
 * and ResidualBasedExpander._find_residual_configs (:174-253).
 * The workspace is a device hash map determinant -> accumulator with room for
 * `capacity` distinct candidates (one pool entry per DISTINCT candidate of a pass, however
 * many sources reach it at the same time).  All of its memory is the caller's:
 *   table    uint64[table_slots]   (table_slots a power of two, >= 2*capacity advised)
 *   pool     uint64[capacity][4]   (32-byte aligned; one entry = {alpha, beta, accumulator lo,
 *                                   hi}: key and sum share a 32-byte sector)
 *   counters uint64[4]
 * SUM mode accumulates in signed 128-bit fixed point (2^-70 resolution, |addend| < 2^30):
 * integer adds are associative, so coupling sums are bit-identical from run to run and for
 * any split into passes / owner ranks; they become FP64 (one rounding) in score / export.
 * fgk_pt2_create only wraps them in a handle; call fgk_pt2_reset before use (it clears the
 * table and the counters; pool entries are written complete when they are claimed). */
int fgk_pt2_create(int64_t capacity, int64_t table_slots, uint64_t* table, uint64_t* pool,
                   uint64_t* counters, int device, fgk_pt2_t* out);
int fgk_pt2_destroy(fgk_pt2_t ws);
/* Optional radix partition in front of the hash (large sweeps, whose accumulator is far
 * bigger than L2): the table is split into 2^region_bits regions chosen by the top hash bits
 * and fgk_pt2_accumulate first appends every candidate to one of 2^queue_bits queues
 * (queue_bits >= region_bits; queue_stride pairs each: queue_dets uint64[2^queue_bits *
 * stride][2], queue_vals double[...], queue_cursors uint64[2^queue_bits], all caller memory),
 * then folds the queues into the table in queue order, so that the table region and pool
 * slots being written are L2-resident.  Call with queue_bits = 0 to go back to the direct sweep.
 * Must be followed by fgk_pt2_reset. */
int fgk_pt2_set_partition(fgk_pt2_t ws, int region_bits, int queue_bits, int64_t queue_stride,
                          uint64_t* queue_dets, double* queue_vals, uint64_t* queue_cursors);
int fgk_pt2_reset(fgk_pt2_t ws, void* stream);
/* For every source s in [0,n_src): j = src_idx[s] (basis position), coefficient
 * coeff[s]; every connection x of basis[j] with x not in the basis adds
 * coeff[s]*<x|H|j> to x's accumulator (mode SUM) or maxes |.| (mode MAXABS).
 * Only the candidates of bucket pass_id out of n_pass are handled, so that a candidate set
 * larger than the workspace can be processed exactly in n_pass sweeps (and N ranks can own
 * disjoint buckets).  Buckets partition the candidate space: by a hash of the candidate's
 * alpha string (up to 64 buckets; runs of connections that share an alpha string are skipped
 * as a whole), refined by a hash of the full determinant beyond that.  Overflow is reported
 * by fgk_pt2_count. */
int fgk_pt2_accumulate(fgk_ham_t h, fgk_index_t idx, fgk_pt2_t ws, const int64_t* src_idx,
                       const double* coeff, int64_t n_src, int mode, int n_pass, int pass_id,
                       void* stream);
/* merge externally produced (determinant, value) pairs (multi-GPU dedup exchange) */
int fgk_pt2_merge(fgk_pt2_t ws, const uint64_t* dets, const double* vals, int64_t m, int mode,
                  void* stream);
/* synchronises; host outputs: number of pool slots used (= distinct candidates), raw
 * candidates tested so far, overflow flag.  Returns FGK_ERR_CAPACITY on overflow, FGK_ERR_ARG
 * if an addend was not finite or >= 2^30 in magnitude. */
int fgk_pt2_count(fgk_pt2_t ws, void* stream, int64_t* n_slots, int64_t* n_raw, int* overflow);
/* The live candidates among the first n_slots pool slots, COMPACTED to the front of
 * the outputs (buffers sized n_slots; order unspecified): determinant, accumulated
 * coupling and, if h != NULL, out_diag = <x|H|x> and
 * out_importance = coupling^2 / (|energy - diag| + 1e-10)  (:547-548).
 * Synchronises; *n_live (host) = number of candidates written. */
int fgk_pt2_export(fgk_ham_t h, fgk_pt2_t ws, int64_t n_slots, double energy, uint64_t* out_dets,
                   double* out_coupling, double* out_diag, double* out_importance,
                   int64_t* n_live, void* stream);

/* Streaming selection (what _find_important_configs' torch.topk, :551-552, needs): instead of
 * exporting every candidate, fgk_pt2_score computes each live candidate's score in place --
 * importance coupling^2 / (|energy - diag| + 1e-10) if h != NULL, else |coupling| -- counts
 * scores per binary exponent in hist (device uint32[2048], scratch) and derives the bound
 * thr (device uint64[2]: {lowest score bit pattern to keep, number kept}) that keeps at
 * least k candidates plus one exponent bin of slack.  Synchronises; *n_live = live
 * candidates, *n_keep = candidates at or above the bound (size the gather buffers with it).
 * fgk_pt2_gather compacts those candidates (order unspecified) into out_dets / out_score;
 * synchronises; FGK_ERR_CAPACITY if out_cap is too small. */
int fgk_pt2_score(fgk_ham_t h, fgk_pt2_t ws, int64_t n_slots, double energy, int64_t k, uint32_t* hist,
                  uint64_t* thr, int64_t* n_live, int64_t* n_keep, void* stream);
int fgk_pt2_gather(fgk_pt2_t ws, int64_t n_slots, const uint64_t* thr, uint64_t* out_dets,
                   double* out_score, int64_t out_cap, int64_t* n_written, void* stream);

/* Dedup exchange helper: owner rank of a determinant = (hash >> 24) % world.
 * scatter = 0: cursors[w] += number of pairs owned by w (cursors zeroed by the caller);
 * scatter = 1: cursors[w] hold the segment starts (exclusive scan of the counts); pairs are
 * copied into out_dets / out_vals so that every owner's pairs are contiguous. */
int fgk_partition_by_owner(const uint64_t* dets, const double* vals, int64_t m, int world,
                           uint64_t* cursors, uint64_t* out_dets, double* out_vals, int scatter,
                           int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FGK_B200_H */
