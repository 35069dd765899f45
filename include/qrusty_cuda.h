/*
 * qrusty_cuda.h -- C ABI of the B200 (sm_100a) implementation of qrusty's
 * SparsePauliOp -> CSR hot path and the matrix-free Pauli-sum apply H.v.
 *
 * This is the boundary a `qrusty::cuda` Rust module binds with `extern "C"`
 * (see INTEGRATION.md and rust/cuda.rs).  Plain pointers and sizes only.
 * File:line citations are into the reference repository (chetmurthy/qrusty).
 *
 * Conventions
 *  - every function returns an int status (QR_OK = 0); on failure
 *    qr_last_error() returns a thread-local message.  Nothing aborts or throws
 *    across this boundary (the reference panics/returns QrustyErr instead,
 *    qrusty/src/lib.rs:34-52).
 *  - complex128 is two consecutive doubles (re, im), as num_complex::Complex64.
 *  - the library never frees caller memory and never keeps caller pointers
 *    after a call returns (asynchronous calls: until the stream reaches them).
 *  - a plan is not re-entrant; distinct plans are independent.  Callable from
 *    any host thread.
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with QR_ERR_CUDA.
 */
#ifndef QRUSTY_CUDA_H
#define QRUSTY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define QR_API __attribute__((visibility("default")))
#else
#define QR_API
#endif

#define QR_OK               0
#define QR_ERR_INVALID      1   /* bad argument                                 */
#define QR_ERR_CUDA         2   /* CUDA runtime / launch failure, no device     */
#define QR_ERR_NCCL         3   /* NCCL missing or failed                       */
#define QR_ERR_OOM          4   /* host or device allocation failed             */
#define QR_ERR_UNSUPPORTED  5   /* e.g. n_qubits > 32                           */

#define QR_VERSION 100          /* 0.1.0 */

/* One Pauli term as rowwise::make_params emits it (qrusty/src/accel.rs:141-157):
 * (z_indices, x_indices, coeff') with coeff' = (-i)^phase * coeff already applied
 * by the host (Pauli::{phase,x_indices,z_indices}, qrusty/src/lib.rs:161-181).
 * Bit k of x/z is qubit k = k-th label character from the right (lib.rs:144). */
typedef struct qr_term {
    uint64_t z;
    uint64_t x;
    double   re;
    double   im;
} qr_term;

typedef struct qr_plan qr_plan;     /* opaque: canonicalised operator resident on one GPU */
typedef struct qr_comm qr_comm;     /* opaque: NCCL communicator, one rank per process     */

typedef struct qr_plan_info_t {
    int32_t  n_qubits;
    int32_t  device;
    uint64_t dim;        /* 2^n_qubits rows and columns                                   */
    uint64_t n_terms;    /* T as supplied (after QR_PLAN_MERGE_DUPLICATES: see qr_plan_groups)   */
    uint64_t n_groups;   /* G = distinct X-masks = stored entries per row                 */
    uint64_t nnz;        /* G * dim: explicit zeros are kept, as accel.rs:171-210 does    */
} qr_plan_info_t;

/* ---- plan: replaces make_params' consumer side + the grouping that make_row
 * redoes per row with a sort (accel.rs:174-205).  Uploads the T terms, runs the
 * canonicalisation kernel (stable radix sort by X-mask, head-flag scan -> groups,
 * rank tables) on `device`, and waits for it.  n_qubits in [1, 32], n_terms >= 1
 * (SparsePauliOp::new, lib.rs:354-376), every x and z < 2^n_qubits. */
#define QR_PLAN_MERGE_DUPLICATES 1u  /* merge terms with identical (x, z): segmented reduce after a
                                      (x, z)-keyed sort.  Changes the summation order inside a
                                      group: data then matches the reference to 1e-12, not bit
                                      for bit.  Off by default. */
QR_API int qr_plan_create(int n_qubits, const qr_term *terms, size_t n_terms, int device,
                   uint32_t flags, qr_plan **out);
QR_API int qr_plan_destroy(qr_plan *plan);
QR_API int qr_plan_info(const qr_plan *plan, qr_plan_info_t *info);

/* Canonicalisation output, for inspection/tests: the G distinct X-masks in
 * ascending order, group_offsets[G+1] into the sorted term list, and
 * term_order[T] = original index of each sorted term (stable: original order
 * inside a group).  Any pointer may be NULL.  Synchronous. */
QR_API int qr_plan_groups(const qr_plan *plan, uint64_t *xmask, uint32_t *group_offsets,
                   uint32_t *term_order);
/* Number of terms the kernels iterate over: n_terms, or fewer after QR_PLAN_MERGE_DUPLICATES
 * (group_offsets then index the merged list and term_order[i] is the first original index of
 * merged term i, for i < the returned count). */
QR_API int qr_plan_canonical_terms(const qr_plan *plan, uint64_t *count);
/* Name of the fill kernel qr_build_rows_device launches for an aligned row window of this plan
 * ("fill_staged_kernel", "fill_staged_swz_kernel", "fill_rows_kernel", "fill_lanes_kernel", "fill_blocked_kernel" or
 * "fill_direct_kernel"); a static string, for benchmarks and logs.  Never NULL. */
QR_API const char *qr_plan_fill_kernel(const qr_plan *plan);
/* Name of the kernel qr_apply_device / qr_apply_p2p launch for rows [row_lo, row_hi) of this plan with the
 * default settings: "apply_fold_kernel" (term-rich operators: bucketed terms + in-register Walsh-Hadamard fold,
 * rows in whole aligned blocks of 1024) or "apply_direct_kernel" (gather).  Builds the fold tables on first use.
 * A static string; "" on error (qr_last_error). */
QR_API const char *qr_plan_apply_kernel(qr_plan *plan, uint64_t row_lo, uint64_t row_hi);

/* Re-runs the canonicalisation kernel from the raw term table already in HBM,
 * asynchronously on `stream` (a cudaStream_t, NULL = default stream).  Lets a
 * caller time the whole device sequence canonicalise -> fill. */
QR_API int qr_plan_canonicalise_async(qr_plan *plan, void *stream);

/* ---- CSR build: replaces rowwise::make_unsafe_vectors_chunked
 * (accel.rs:267-336) for rows [row_lo, row_hi).  Output layout is that of
 * UnsafeVectors (accel.rs:15-20) / CsMatI<Complex64,u64,u64> (lib.rs:558-571):
 *   d_indptr  u64[row_hi-row_lo+1]
 *   d_indices u64[(row_hi-row_lo)*G]   GLOBAL column ids, ascending inside a row
 *   d_data    complex128[(row_hi-row_lo)*G]
 * indptr values: with QR_INDPTR_LOCAL (default) the shard is a self-contained
 * CSR, d_indptr[i] = i*G; with QR_INDPTR_GLOBAL d_indptr[i] = (row_lo+i)*G as in
 * the full matrix.  d_indptr may be NULL (skip).  Device pointers, 16-byte
 * aligned; asynchronous on `stream`. */
#define QR_INDPTR_LOCAL   0u
#define QR_INDPTR_GLOBAL  1u
#define QR_FILL_DIRECT    2u    /* force the unstaged kernel (debug / comparison)  */
QR_API int qr_build_rows_device(qr_plan *plan, uint64_t row_lo, uint64_t row_hi,
                         uint64_t *d_indptr, uint64_t *d_indices, double *d_data,
                         uint32_t flags, void *stream);

/* Same, into caller-allocated HOST buffers (what the Rust shim hands to
 * CsMatI::new_unchecked): builds on the device in row windows and copies back.
 * Synchronous.  Page-locked buffers (qr_malloc_host) receive the windows directly; ordinary
 * pageable memory (a Rust Vec, a numpy array) is filled through pinned staging windows by a
 * few host threads, which keeps the PCIe copy at full rate (QR_HOST_NO_STAGING: plain cudaMemcpy). */
#define QR_HOST_NO_STAGING 4u
/* On the wire (PCIe is the bound of this call: 56 GB/s against a 6.4 TB/s build) the shard is 17-18 bytes per entry, not 24:
 * data as stored; the column of an entry as the id of its group (1 byte while G <= 256, 2 while G <= 65536, else the 32-bit
 * column), rebuilt into the u64 column row ^ mask by host threads while the next window is in flight; indptr written by the
 * host (affine, r*G).  The arrays the caller receives are the same, bit for bit.  QR_HOST_WIDE: copy all three arrays as
 * stored (24 + 8/G bytes per entry; comparison and tests). */
#define QR_HOST_WIDE 8u
QR_API int qr_build_host(qr_plan *plan, uint64_t row_lo, uint64_t row_hi,
                  uint64_t *indptr, uint64_t *indices, double *data, uint32_t flags);
/* Bytes the last qr_build_host of the calling thread copied device -> host. */
QR_API uint64_t qr_last_d2h_bytes(void);

/* rawio::write (qrusty/src/rawio.rs:128-148) of rows [row_lo,row_hi) as a (row_hi-row_lo) x 2^n CSR,
 * streamed from the GPU in row windows (fill -> pinned staging -> pwrite at the section offsets): the
 * matrix is never resident as a whole in HBM or in host memory.  Synchronous. */
QR_API int qr_write_rawio(qr_plan *plan, uint64_t row_lo, uint64_t row_hi, const char *path);

/* ---- matrix-free H.v: replaces build + rowwise::spmat_dot_densevec
 * (accel.rs:338-370) without reading a matrix.  d_v is the FULL vector
 * (complex128[dim]); d_y receives rows [row_lo,row_hi) (complex128[row_hi-row_lo]).
 * Asynchronous on `stream`. */
QR_API int qr_apply_device(qr_plan *plan, uint64_t row_lo, uint64_t row_hi,
                    const double *d_v, double *d_y, void *stream);
QR_API int qr_apply_host(qr_plan *plan, const double *v, double *y);   /* full vector, host buffers */

/* diag(H) for rows [row_lo,row_hi) (SpMat.diagonal, pyqrusty/src/lib.rs:118-125). */
QR_API int qr_diagonal_device(qr_plan *plan, uint64_t row_lo, uint64_t row_hi, double *d_diag, void *stream);

/* diag of a CSR shard already resident in HBM: d_diag[i] = the stored entry (i, col0 + i), 0 if none
 * (SpMat.diagonal, pyqrusty/src/lib.rs:118-125, on a matrix that was scaled or compacted after the build). */
QR_API int qr_csr_diagonal_device(uint64_t n_rows, uint64_t col0, const uint64_t *d_indptr, const uint64_t *d_indices,
                           const double *d_data, double *d_diag, void *stream);
/* CSR SpMV on a device-resident CSR shard, the reference's own H.v
 * (accel.rs:338-370): y[r] = sum_k data[k]*v[indices[k]] in stored order,
 * starting from zero; indptr may be local or global (rebased by d_indptr[0]). */
QR_API int qr_spmv_device(uint64_t n_rows, const uint64_t *d_indptr, const uint64_t *d_indices,
                   const double *d_data, const double *d_v, double *d_y, void *stream);

/* ---- zero elimination on a device-resident shard with uniform row length G (every plan-built
 * shard): util::csmatrix_nz / csmatrix_eliminate_zeroes (qrusty/src/util.rs:144-171; Python
 * SpMat.count_zeros / eliminate_zeros, pyqrusty/src/lib.rs:170-183).  An entry is kept iff
 * hypot(re, im) > tol.  qr_count_kept_device writes the compacted CSR's indptr (local, starting
 * at 0: per-row counts + prefix scan) and returns its nnz (synchronises); qr_compact_rows_device
 * then writes indices/data of the kept entries, row-major order preserved. */
QR_API int qr_count_kept_device(uint64_t n_rows, uint64_t n_groups, const double *d_data, double tol,
                         uint64_t *d_indptr_out, uint64_t *nnz_out, void *stream);
QR_API int qr_compact_rows_device(uint64_t n_rows, uint64_t n_groups, const uint64_t *d_indices,
                           const double *d_data, double tol, const uint64_t *d_indptr,
                           uint64_t *d_indices_out, double *d_data_out, void *stream);
/* The same two passes for ANY device-resident CSR (rows of any length: a matrix wrapped by SpMat::new_unchecked,
 * pyqrusty/src/lib.rs:104-116, read back from disk, or already compacted) -- util::csmatrix_nz and
 * csmatrix_eliminate_zeroes take any CsMat (util.rs:144-171).  d_indptr_in is the stored indptr (n_rows + 1 entries;
 * its first entry may be a global offset), d_indptr_out receives the new local indptr. */
QR_API int qr_csr_count_kept_device(uint64_t n_rows, const uint64_t *d_indptr_in, const double *d_data, double tol,
                                    uint64_t *d_indptr_out, uint64_t *nnz_out, void *stream);
QR_API int qr_csr_compact_device(uint64_t n_rows, const uint64_t *d_indptr_in, const uint64_t *d_indices, const double *d_data,
                                 double tol, const uint64_t *d_indptr, uint64_t *d_indices_out, double *d_data_out,
                                 void *stream);

/* Fused drop-zeros build: the CSR that csmatrix_eliminate_zeroes (util.rs:154-171) produces from
 * rows [row_lo,row_hi) of the reference's build, without writing the explicit zeros first.
 * qr_build_compact_count evaluates every (row, group) value in registers, counts the kept ones per
 * row, prefix-scans the counts into d_indptr (u64[row_hi-row_lo+1], local, d_indptr[0] = 0) and
 * returns the shard's nnz (synchronises).  qr_build_compact_fill then writes indices/data of the
 * kept entries (u64[nnz], complex128[nnz]; global column ids, ascending inside a row). */
QR_API int qr_build_compact_count(qr_plan *plan, uint64_t row_lo, uint64_t row_hi, double tol,
                           uint64_t *d_indptr, uint64_t *nnz_out, void *stream);
QR_API int qr_build_compact_fill(qr_plan *plan, uint64_t row_lo, uint64_t row_hi, double tol,
                          const uint64_t *d_indptr, uint64_t *d_indices, double *d_data, void *stream);

/* Lanczos/Davidson vector kernels (accel.rs:374-393): z = a*x + b*y, a*x + y, a*x. */
QR_API int qr_axpby_device(uint64_t n, const double a[2], const double *d_x, const double b[2],
                    const double *d_y, double *d_z, void *stream);
QR_API int qr_axpy_device(uint64_t n, const double a[2], const double *d_x, const double *d_y,
                   double *d_z, void *stream);
QR_API int qr_ax_device(uint64_t n, const double a[2], const double *d_x, double *d_z, void *stream);
/* Davidson preconditioner precond2 (pyqrusty/src/lib.rs:457-468, reg() at :436-440):
 * d_out[i] = d_dx[i] / reg(d_diag[i] - e), reg(x) = (tol, 0) if |x| < tol else x.  precond
 * (lib.rs:442-455) is the same on qr_diagonal_device's output. */
QR_API int qr_precond2_device(uint64_t n, const double *d_diag, const double *d_dx, const double e[2],
                       double tol, double *d_out, void *stream);
/* <x,y> = sum conj(x_i) y_i  -> d_out[2] (numpy.vdot; the reference leaves this to numpy). */
QR_API int qr_dotc_device(uint64_t n, const double *d_x, const double *d_y, double *d_out, void *stream);

/* Lanczos three-term recurrence fused with the norm (the vector work either side of H.v in an
 * eigensolver iteration): d_w_out = d_w - alpha*d_v - beta*d_v_prev (d_v_prev may be NULL),
 * d_norm2_out[0] = sum |w_out|^2 (one double).  d_w_out may alias d_w. */
QR_API int qr_lanczos_update_device(uint64_t n, const double alpha[2], const double beta[2], const double *d_w,
                             const double *d_v, const double *d_v_prev, double *d_w_out,
                             double *d_norm2_out, void *stream);

/* ---- multi-GPU, one process per GPU.  NCCL is dlopen'ed on first use.  The
 * CSR build needs no communication; H.v all-gathers the row-sharded vector. */
#define QR_UNIQUE_ID_BYTES 128
QR_API int qr_comm_unique_id(void *id_out /* QR_UNIQUE_ID_BYTES */);
QR_API int qr_comm_create(const void *id, int n_ranks, int rank, int device, qr_comm **out);
QR_API int qr_comm_destroy(qr_comm *comm);
/* d_v_shard: this rank's dim/n_ranks elements; d_v_full: scratch of dim elements;
 * d_y_shard: this rank's rows of H.v.  ncclAllGather then the local apply. */
QR_API int qr_apply_distributed(qr_plan *plan, qr_comm *comm, const double *d_v_shard,
                         double *d_v_full, double *d_y_shard, void *stream);
/* y = H v on rows [row_lo, row_hi) AND d_dot[0..1] = sum over those rows of conj(v_r) y_r, in one pass: the apply
 * kernels fold <v, H v> per CTA in their epilogue and one CTA folds the partials (fixed order: deterministic).  A
 * Lanczos / Davidson step then spends no separate pass on the Rayleigh quotient.  d_dot: device, 16-byte aligned.
 * The sharded form reads the peers' shards like qr_apply_p2p; d_dot is this rank's part (all-reduce it). */
QR_API int qr_apply_dot_device(qr_plan *plan, uint64_t row_lo, uint64_t row_hi, const double *d_v, double *d_y,
                               double *d_dot, void *stream);
QR_API int qr_apply_p2p_dot(qr_plan *plan, qr_comm *comm, const double *const *v_shards, double *d_y_shard,
                            double *d_dot, void *stream);
/* Lanczos with its scalars on the device -- no host round trip inside an iteration.  d_state: 16 + 2 K doubles,
 * 16-byte aligned: [0,1] <u, H u> (written by qr_apply_*_dot)  [2,3] ||w||^2 (written by qr_lanczos_update_dev)
 * [4] beta_{k-2}-scale  [5] ||u_k|| (initialise to 1 for a normalised start vector)  [6..8] coefficients
 * [16 + k] alpha_k  [16 + K + k] beta_k.  The vectors stay unnormalised (u_{k+1} = w_k), so no rescale pass:
 *   qr_lanczos_coef_device(state, k, K, 0)   after the apply (+ all-reduce of [0,1]):  alpha_k, coefficients
 *   qr_lanczos_update_dev(n, state, y, u, u_prev, w)   w = (H u)/b - (alpha/b) u - (b/b') u_prev, [2] = ||w||^2
 *   qr_lanczos_coef_device(state, k, K, 1)   after the update (+ all-reduce of [2]):   beta_k = sqrt([2]) */
QR_API int qr_lanczos_coef_device(double *d_state, uint32_t k, uint32_t K, uint32_t phase, void *stream);
QR_API int qr_lanczos_update_dev(uint64_t n, double *d_state, const double *d_y, const double *d_u, const double *d_u_prev,
                                 double *d_w_out, void *stream);
QR_API int qr_allreduce_sum_f64(qr_comm *comm, double *d_buf, size_t count, void *stream);

/* Fused distributed H.v over peer memory: no all-gather and no full copy of v.  Every rank
 * passes the device pointers of ALL ranks' v shards (its own at [rank]; the others opened from
 * CUDA IPC handles, below).  The apply kernel reads v[r ^ x_g] straight from the owning GPU
 * over NVLink: the owner is rank ^ (x_g >> log2(rows per rank)), the same for every row of the
 * caller, so only groups whose mask touches the sharded (top) row bits cross the link.  Two
 * stream-ordered NCCL barriers bracket the kernel (shards ready / shards no longer read). */
QR_API int qr_apply_p2p(qr_plan *plan, qr_comm *comm, const double *const *v_shards,
                 double *d_y_shard, void *stream);
#define QR_IPC_HANDLE_BYTES 64
QR_API int qr_ipc_get_handle(void *d_ptr, void *handle_out /* QR_IPC_HANDLE_BYTES */);
QR_API int qr_ipc_open_handle(const void *handle, void **d_ptr_out);
QR_API int qr_ipc_close_handle(void *d_ptr);

/* ---- runtime helpers for hosts without a CUDA binding of their own ---- */
QR_API int qr_device_count(int *count);
QR_API int qr_device_name(int device, char *buf, size_t buf_len);
QR_API int qr_set_device(int device);
QR_API int qr_malloc_device(void **ptr, size_t bytes);
QR_API int qr_free_device(void *ptr);
QR_API int qr_malloc_host(void **ptr, size_t bytes);     /* page-locked */
QR_API int qr_free_host(void *ptr);
QR_API int qr_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);  /* async if stream != NULL... */
QR_API int qr_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
QR_API int qr_memset_device(void *dst, int value, size_t bytes, void *stream);
QR_API int qr_stream_create(void **stream);
QR_API int qr_stream_destroy(void *stream);
QR_API int qr_stream_synchronize(void *stream);          /* NULL = whole device */
/* CUDA graphs: record the asynchronous calls issued on `stream` between begin and end (every *_device /
 * *_async entry point of this header is capturable), replay them with one launch. */
QR_API int qr_graph_begin_capture(void *stream);
QR_API int qr_graph_end_capture(void *stream, void **graph_exec);
QR_API int qr_graph_launch(void *graph_exec, void *stream);
QR_API int qr_graph_destroy(void *graph_exec);
QR_API int qr_event_create(void **event);
QR_API int qr_event_destroy(void *event);
QR_API int qr_event_record(void *event, void *stream);
QR_API int qr_event_elapsed_ms(void *start, void *stop, float *ms);   /* synchronises on `stop` */

/* Frees the per-device staging windows qr_build_host keeps between calls (2 x <= 256 MB of HBM). */
QR_API int qr_release_scratch(void);
/* Number of this library's kernels launched by the calling process so far. */
QR_API uint64_t qr_kernel_launches(void);
QR_API const char *qr_last_error(void);
QR_API int qr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QRUSTY_CUDA_H */
