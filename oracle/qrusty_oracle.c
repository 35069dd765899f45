/*
 * qrusty_oracle.c -- CPU restatement of qrusty's SparsePauliOp -> CSR path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under qrusty_b200/ may link, import or
 * call this file; it is the checker for tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py.
 *
 * The reference (chetmurthy/qrusty) is Rust and cannot be built in this image
 * (no cargo/rustc; `sprs` and `rayon-subslice` are un-vendored git
 * dependencies, Cargo.lock:1055-1057, 933-935), so this file restates the
 * algorithm function by function.  Citations are into /root/reference.
 *
 *   oracle_parse_label        qrusty/src/lib.rs:125-149 (grammar, reversal),
 *                             :80-91 (x/z per letter), :161-181 (masks, phase)
 *   oracle_make_params        qrusty/src/accel.rs:141-157
 *   oracle_make_row           qrusty/src/accel.rs:171-210
 *   oracle_build_chunked      qrusty/src/accel.rs:267-336
 *   oracle_single_pauli       qrusty/src/accel.rs:22-122
 *   oracle_spmv               qrusty/src/accel.rs:338-370
 *   oracle_axpby/axpy/ax      qrusty/src/accel.rs:374-393
 *   oracle_build_grouped      NOT the reference's algorithm: a tuned CPU variant (group-first, closed-form slots,
 *                             no per-row sort, no concat passes) shown beside it as a fairness line;
 *                             must produce the same bytes as oracle_build_chunked
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against every
 * known-answer the reference's own tests hold for the path (lib.rs:608-919,
 * pyqrusty/tests/test_it.py:60-101, test_H.py:32-52) and against an
 * independent numpy restatement of the reference's default kron-and-add
 * to_matrix (lib.rs:203-213, 401-412) on the H2/H4/H6 fixtures.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction,
 * so every add/multiply rounds exactly as the Rust code does).
 */
#include <pthread.h>
#include <stdatomic.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct { double re, im; } c128;

/* one element of make_params' output: (z_indices, x_indices, coeff')  accel.rs:154 */
typedef struct { uint64_t z, x; c128 c; } oracle_param;

/* num_complex Mul: (a+bi)(c+di) = (ac - bd) + (ad + bc)i */
static inline c128 cmul(c128 a, c128 b) {
    c128 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline c128 cadd(c128 a, c128 b) { c128 r = { a.re + b.re, a.im + b.im }; return r; }
static inline c128 cneg(c128 a) { c128 r = { -a.re, -a.im }; return r; }

/* ------------------------------------------------------------------------- */
/* Pauli label model                                                          */
/* ------------------------------------------------------------------------- */

/* Regex ^([+-]?)1?([ij]?)([IXYZ]+)$  (lib.rs:127).  base_phase = imag + 2*neg
 * (lib.rs:135-141).  Characters are consumed right-to-left: qubit k is the
 * k-th character from the right (lib.rs:144).  x bit <- X|Y, z bit <- Y|Z
 * (lib.rs:80-91, 161-176).  Returns 0 on success, -1 on a malformed label,
 * -2 if more than 64 qubits. */
int oracle_parse_label(const char *s, int *base_phase, int *n_qubits,
                       uint64_t *x, uint64_t *z, int *n_y)
{
    size_t i = 0, len = strlen(s);
    int neg = 0, imag = 0;
    if (i < len && (s[i] == '+' || s[i] == '-')) { neg = (s[i] == '-'); i++; }
    if (i < len && s[i] == '1') i++;
    if (i < len && (s[i] == 'i' || s[i] == 'j')) { imag = 1; i++; }
    if (i >= len) return -1;                     /* [IXYZ]+ needs one char */
    size_t nq = len - i;
    for (size_t k = i; k < len; k++)
        if (s[k] != 'I' && s[k] != 'X' && s[k] != 'Y' && s[k] != 'Z') return -1;
    if (nq > 64) return -2;
    uint64_t xm = 0, zm = 0; int ny = 0;
    for (size_t q = 0; q < nq; q++) {
        char c = s[len - 1 - q];
        if (c == 'X' || c == 'Y') xm |= (uint64_t)1 << q;
        if (c == 'Y' || c == 'Z') zm |= (uint64_t)1 << q;
        if (c == 'Y') ny++;
    }
    *base_phase = imag + (neg ? 2 : 0);
    *n_qubits = (int)nq;
    *x = xm; *z = zm; *n_y = ny;
    return 0;
}

/* accel.rs:141-157.  phase = (base_phase + #Y) % 4 (lib.rs:179-181) selects
 * the unit {1, -i, -1, +i} that multiplies the coefficient.
 *
 * convention 0 ("rowwise"): exactly accel.rs, unit = (-i)^(base_phase + nY).
 * convention 1 ("to_matrix"): the reference's default to_matrix scales the
 *   Kronecker product by base_coeff = (+i)^base_phase (lib.rs:182-190, 211),
 *   so unit = (+i)^base_phase (-i)^nY = (-i)^((nY - base_phase) mod 4).  The two
 *   agree for every label without an i/j prefix (base_phase in {0,2}).       */
void oracle_make_params(const int *base_phase, const int *n_y, const uint64_t *x,
                        const uint64_t *z, const c128 *coeff, size_t n_terms,
                        int convention, oracle_param *out)
{
    static const c128 unit[4] = { {1.0, 0.0}, {0.0, -1.0}, {-1.0, 0.0}, {0.0, 1.0} };
    for (size_t t = 0; t < n_terms; t++) {
        int ph = convention == 0 ? (base_phase[t] + n_y[t]) % 4
                                 : (((n_y[t] - base_phase[t]) % 4) + 4) % 4;
        out[t].z = z[t];
        out[t].x = x[t];
        out[t].c = cmul(unit[ph], coeff[t]);
    }
}

/* ------------------------------------------------------------------------- */
/* make_row                                                                   */
/* ------------------------------------------------------------------------- */

typedef struct { uint64_t col; c128 v; } pair_t;

/* Stable sort by column, as slice::sort_by (accel.rs:188): insertion sort for
 * short inputs, top-down merge sort otherwise.  Any stable sort yields the
 * same sequence. */
static void insertion_sort(pair_t *a, size_t n) {
    for (size_t i = 1; i < n; i++) {
        pair_t key = a[i];
        size_t j = i;
        while (j > 0 && a[j - 1].col > key.col) { a[j] = a[j - 1]; j--; }
        a[j] = key;
    }
}
static void merge_sort(pair_t *a, pair_t *tmp, size_t n) {
    if (n <= 20) { insertion_sort(a, n); return; }
    size_t h = n / 2;
    merge_sort(a, tmp, h);
    merge_sort(a + h, tmp, n - h);
    if (a[h - 1].col <= a[h].col) return;
    memcpy(tmp, a, h * sizeof(pair_t));
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) a[k++] = (a[j].col < tmp[i].col) ? a[j++] : tmp[i++];
    while (i < h) a[k++] = tmp[i++];
}

/* accel.rs:171-210.  Maps every term to (row ^ x, +-c'), stable-sorts by
 * column, folds equal columns left to right (the first element of a run is
 * taken as is, not added to zero).  Explicit zeros are kept.  Writes at most
 * n_terms entries; returns the count.  scratch: 2*n_terms pair_t. */
size_t oracle_make_row_scratch(const oracle_param *params, size_t n_terms, uint64_t row,
                               uint64_t *cols, c128 *vals, pair_t *scratch)
{
    pair_t *v = scratch;
    for (size_t t = 0; t < n_terms; t++) {
        v[t].col = row ^ params[t].x;
        int odd = __builtin_popcountll(row & params[t].z) & 1;     /* accel.rs:179 */
        v[t].v = odd ? cneg(params[t].c) : params[t].c;
    }
    merge_sort(v, scratch + n_terms, n_terms);
    size_t out = 0;
    uint64_t col = v[0].col; c128 sum = v[0].v;
    for (size_t t = 1; t < n_terms; t++) {
        if (v[t].col == col) sum = cadd(sum, v[t].v);
        else { cols[out] = col; vals[out] = sum; out++; col = v[t].col; sum = v[t].v; }
    }
    cols[out] = col; vals[out] = sum; out++;
    return out;
}

size_t oracle_make_row(const oracle_param *params, size_t n_terms, uint64_t row,
                       uint64_t *cols, c128 *vals)
{
    pair_t *scratch = (pair_t *)malloc(2 * n_terms * sizeof(pair_t));
    size_t n = oracle_make_row_scratch(params, n_terms, row, cols, vals, scratch);
    free(scratch);
    return n;
}

/* ------------------------------------------------------------------------- */
/* make_unsafe_vectors_chunked  (accel.rs:267-336)                             */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint64_t lo, hi;
    uint64_t *row_nnz;     /* v_nnz           accel.rs:294 */
    uint64_t *indices;     /* per-chunk concat accel.rs:300 */
    c128 *data;            /*                  accel.rs:301 */
    uint64_t nnz;
} chunk_t;

typedef struct {
    const oracle_param *params; size_t n_terms;
    chunk_t *chunks; size_t n_chunks;
    atomic_size_t next;
    int failed;
} build_job;

/* One rayon task of accel.rs:291-304: make_row for every row of the chunk into
 * per-row vectors, then concatenate them into the chunk's vectors. */
static void *build_worker(void *arg)
{
    build_job *job = (build_job *)arg;
    size_t T = job->n_terms;
    pair_t *scratch = (pair_t *)malloc(2 * T * sizeof(pair_t));
    uint64_t *rcols = (uint64_t *)malloc(T * sizeof(uint64_t));
    c128 *rvals = (c128 *)malloc(T * sizeof(c128));
    for (;;) {
        size_t ci = atomic_fetch_add(&job->next, 1);
        if (ci >= job->n_chunks) break;
        chunk_t *ch = &job->chunks[ci];
        size_t rows = (size_t)(ch->hi - ch->lo);
        uint64_t **pc = (uint64_t **)malloc(rows * sizeof(*pc));   /* v_rc: one Vec pair per row */
        c128 **pv = (c128 **)malloc(rows * sizeof(*pv));
        ch->row_nnz = (uint64_t *)malloc(rows * sizeof(uint64_t));
        uint64_t sum = 0;
        for (size_t i = 0; i < rows; i++) {
            size_t k = oracle_make_row_scratch(job->params, T, ch->lo + i, rcols, rvals, scratch);
            pc[i] = (uint64_t *)malloc(k * sizeof(uint64_t));
            pv[i] = (c128 *)malloc(k * sizeof(c128));
            memcpy(pc[i], rcols, k * sizeof(uint64_t));
            memcpy(pv[i], rvals, k * sizeof(c128));
            ch->row_nnz[i] = k; sum += k;
        }
        ch->nnz = sum;
        ch->indices = (uint64_t *)malloc((sum ? sum : 1) * sizeof(uint64_t));
        ch->data = (c128 *)malloc((sum ? sum : 1) * sizeof(c128));
        uint64_t off = 0;
        for (size_t i = 0; i < rows; i++) {                        /* unsafe_concat_slices */
            memcpy(ch->indices + off, pc[i], ch->row_nnz[i] * sizeof(uint64_t));
            memcpy(ch->data + off, pv[i], ch->row_nnz[i] * sizeof(c128));
            off += ch->row_nnz[i];
            free(pc[i]); free(pv[i]);
        }
        free(pc); free(pv);
    }
    free(scratch); free(rcols); free(rvals);
    return NULL;
}

typedef struct { chunk_t *chunks; size_t lo, hi; const uint64_t *offs; uint64_t *indices; c128 *data; } concat_job;
static void *concat_worker(void *arg)
{
    concat_job *j = (concat_job *)arg;
    for (size_t ci = j->lo; ci < j->hi; ci++) {
        memcpy(j->indices + j->offs[ci], j->chunks[ci].indices, j->chunks[ci].nnz * sizeof(uint64_t));
        memcpy(j->data + j->offs[ci], j->chunks[ci].data, j->chunks[ci].nnz * sizeof(c128));
    }
    return NULL;
}

/* Rows [row_lo, row_hi) of the 2^n x 2^n matrix, chunked by `step` rows over
 * `n_threads` workers.  indptr has (row_hi-row_lo)+1 entries and starts at 0
 * (for row_lo = 0, row_hi = 2^n this is exactly the reference's output).
 * indices/data must hold at least n_terms*(row_hi-row_lo) entries (upper
 * bound); *nnz_out receives the number written.  Returns 0, or -1 on OOM. */
int oracle_build_chunked(const oracle_param *params, size_t n_terms,
                         uint64_t row_lo, uint64_t row_hi, size_t step, int n_threads,
                         uint64_t *indptr, uint64_t *indices, c128 *data, uint64_t *nnz_out)
{
    if (step == 0) step = 1;
    if (n_threads < 1) n_threads = 1;
    uint64_t rows = row_hi - row_lo;
    size_t n_chunks = (size_t)((rows + step - 1) / step);
    chunk_t *chunks = (chunk_t *)calloc(n_chunks ? n_chunks : 1, sizeof(chunk_t));
    if (!chunks) return -1;
    for (size_t ci = 0; ci < n_chunks; ci++) {                     /* accel.rs:283-288 */
        chunks[ci].lo = row_lo + (uint64_t)ci * step;
        chunks[ci].hi = chunks[ci].lo + step < row_hi ? chunks[ci].lo + step : row_hi;
    }
    build_job job = { params, n_terms, chunks, n_chunks, 0, 0 };
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, build_worker, &job);
    build_worker(&job);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);

    /* serial indptr loop, accel.rs:309-319 */
    uint64_t nnz = 0;
    for (uint64_t r = 0; r < rows; r++) {
        size_t ci = (size_t)(r / step), co = (size_t)(r % step);
        indptr[r] = nnz;
        nnz += chunks[ci].row_nnz[co];
    }
    indptr[rows] = nnz;

    /* unsafe_par_concat_slices, accel.rs:322-325 */
    uint64_t *offs = (uint64_t *)malloc((n_chunks + 1) * sizeof(uint64_t));
    offs[0] = 0;
    for (size_t ci = 0; ci < n_chunks; ci++) offs[ci + 1] = offs[ci] + chunks[ci].nnz;
    concat_job *cj = (concat_job *)malloc((size_t)n_threads * sizeof(concat_job));
    for (int i = 0; i < n_threads; i++) {
        cj[i].chunks = chunks; cj[i].offs = offs; cj[i].indices = indices; cj[i].data = data;
        cj[i].lo = n_chunks * (size_t)i / (size_t)n_threads;
        cj[i].hi = n_chunks * (size_t)(i + 1) / (size_t)n_threads;
    }
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, concat_worker, &cj[i]);
    concat_worker(&cj[0]);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);

    for (size_t ci = 0; ci < n_chunks; ci++) { free(chunks[ci].row_nnz); free(chunks[ci].indices); free(chunks[ci].data); }
    free(cj); free(offs); free(th); free(chunks);
    *nnz_out = nnz;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Tuned CPU variant -- a fairness line, NOT a restatement of the reference.   */
/* ------------------------------------------------------------------------- */
/* What a CPU can do with the observations the CUDA path rests on: group the terms by X-mask
 * once (stable, so the summation order inside a group is the reference's), and for every row write each group's
 * entry straight to its final slot, slot(r,g) = sum_b cnt[g][b] * bit_b(x_g ^ r) with
 * cnt[g][b] = #{h != g : msb(x_g ^ x_h) == b}.  No per-row sort, no per-row allocation, no concat passes;
 * threads take contiguous row blocks.  Values are folded exactly as oracle_make_row does (first term as is, then
 * adds in term order), so the output is byte-identical to oracle_build_chunked.  G*n_qubits <= 2^24 words. */
typedef struct {
    const uint64_t *gx; const uint32_t *goff; const uint64_t *tz; const c128 *tc; const uint32_t *cnt;
    size_t G; int nq; uint64_t row_lo, lo, hi; uint64_t *indices; c128 *data;
} grouped_job;

static void *grouped_worker(void *arg)
{
    grouped_job *j = (grouped_job *)arg;
    const size_t G = j->G;
    for (uint64_t r = j->lo; r < j->hi; r++) {
        uint64_t *ri = j->indices + (r - j->row_lo) * G;
        c128 *rd = j->data + (r - j->row_lo) * G;
        for (size_t g = 0; g < G; g++) {
            const uint64_t xr = j->gx[g] ^ r;
            uint32_t slot = 0;
            const uint32_t *c = j->cnt + g * 64;
            for (uint64_t m = xr; m; m &= m - 1) slot += c[__builtin_ctzll(m)];
            uint32_t t = j->goff[g], t1 = j->goff[g + 1];
            c128 v = (__builtin_popcountll(r & j->tz[t]) & 1) ? cneg(j->tc[t]) : j->tc[t];
            for (t++; t < t1; t++)
                v = cadd(v, (__builtin_popcountll(r & j->tz[t]) & 1) ? cneg(j->tc[t]) : j->tc[t]);
            ri[slot] = xr;
            rd[slot] = v;
        }
    }
    return NULL;
}

int oracle_build_grouped(const oracle_param *params, size_t n_terms, int n_qubits,
                         uint64_t row_lo, uint64_t row_hi, int n_threads,
                         uint64_t *indptr, uint64_t *indices, c128 *data, uint64_t *nnz_out)
{
    if (n_threads < 1) n_threads = 1;
    const size_t T = n_terms;
    /* stable sort of term indices by x (insertion into a merge sort of (x, index) pairs) */
    pair_t *a = (pair_t *)malloc(T * sizeof(pair_t)), *tmp = (pair_t *)malloc(T * sizeof(pair_t));
    if (!a || !tmp) { free(a); free(tmp); return -1; }
    for (size_t t = 0; t < T; t++) { a[t].col = params[t].x; a[t].v.re = (double)t; a[t].v.im = 0.0; }
    merge_sort(a, tmp, T);                                         /* stable: equal x keep term order */
    uint64_t *gx = (uint64_t *)malloc(T * sizeof(uint64_t)), *tz = (uint64_t *)malloc(T * sizeof(uint64_t));
    uint32_t *goff = (uint32_t *)malloc((T + 1) * sizeof(uint32_t));
    c128 *tc = (c128 *)malloc(T * sizeof(c128));
    size_t G = 0;
    for (size_t i = 0; i < T; i++) {
        const size_t t = (size_t)a[i].v.re;
        if (i == 0 || a[i].col != a[i - 1].col) { gx[G] = a[i].col; goff[G] = (uint32_t)i; G++; }
        tz[i] = params[t].z; tc[i] = params[t].c;
    }
    goff[G] = (uint32_t)T;
    uint32_t *cnt = (uint32_t *)calloc(G * 64, sizeof(uint32_t));
    for (size_t g = 0; g < G; g++)
        for (size_t h = 0; h < G; h++)
            if (h != g) cnt[g * 64 + (63 - __builtin_clzll(gx[g] ^ gx[h]))]++;
    const uint64_t rows = row_hi - row_lo;
    for (uint64_t r = 0; r <= rows; r++) indptr[r] = r * G;
    grouped_job *jobs = (grouped_job *)malloc((size_t)n_threads * sizeof(grouped_job));
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    for (int i = 0; i < n_threads; i++) {
        grouped_job jb = { gx, goff, tz, tc, cnt, G, n_qubits, row_lo,
                           row_lo + rows * (uint64_t)i / (uint64_t)n_threads, row_lo + rows * (uint64_t)(i + 1) / (uint64_t)n_threads,
                           indices, data };
        jobs[i] = jb;
    }
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, grouped_worker, &jobs[i]);
    grouped_worker(&jobs[0]);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    free(jobs); free(th); free(cnt); free(tc); free(goff); free(tz); free(gx); free(tmp); free(a);
    *nnz_out = rows * G;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* single-Pauli fast path (accel.rs:22-122), kept as an independent check      */
/* ------------------------------------------------------------------------- */
/* indptr = 0..dim, indices = r ^ x, data = +-unit[phase]*coeff.  Note the
 * reference computes mut_phase for group_phase but then uses `phase`
 * (accel.rs:41-49 vs :89); callers pass group_phase = false (lib.rs:220). */
void oracle_single_pauli(uint64_t z, uint64_t x, double coeff_re, double coeff_im, int phase,
                         int n_qubits, uint64_t *indptr, uint64_t *indices, c128 *data)
{
    c128 coeff = { coeff_re, coeff_im };
    static const c128 unit[4] = { {1.0, 0.0}, {0.0, -1.0}, {-1.0, 0.0}, {0.0, 1.0} };
    uint64_t dim = (uint64_t)1 << n_qubits;
    c128 c = cmul(unit[phase % 4], coeff);
    for (uint64_t r = 0; r <= dim; r++) indptr[r] = r;
    for (uint64_t r = 0; r < dim; r++) {
        indices[r] = r ^ x;
        data[r] = (__builtin_popcountll(r & z) & 1) ? cneg(c) : c;
    }
}

/* ------------------------------------------------------------------------- */
/* CSR SpMV (accel.rs:338-370): rows in chunks of 1024, per row a sequential   */
/* dot in stored (column) order starting from zero (sprs dot_dense).           */
/* ------------------------------------------------------------------------- */
typedef struct {
    const uint64_t *indptr, *indices; const c128 *data, *v; c128 *y;
    uint64_t rows, indptr_base; atomic_size_t next; size_t n_chunks;
} spmv_job;

static void *spmv_worker(void *arg)
{
    spmv_job *j = (spmv_job *)arg;
    for (;;) {
        size_t ci = atomic_fetch_add(&j->next, 1);
        if (ci >= j->n_chunks) break;
        uint64_t lo = (uint64_t)ci * 1024, hi = lo + 1024 < j->rows ? lo + 1024 : j->rows;
        for (uint64_t r = lo; r < hi; r++) {
            c128 acc = { 0.0, 0.0 };
            for (uint64_t k = j->indptr[r] - j->indptr_base; k < j->indptr[r + 1] - j->indptr_base; k++)
                acc = cadd(acc, cmul(j->data[k], j->v[j->indices[k]]));
            j->y[r] = acc;
        }
    }
    return NULL;
}

/* y[r] = sum_k data[k] * v[indices[k]] for the `rows` rows described by indptr
 * (which may be a shard: entries are rebased by indptr[0]). */
void oracle_spmv(const uint64_t *indptr, const uint64_t *indices, const c128 *data,
                 uint64_t rows, const c128 *v, c128 *y, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    spmv_job job = { indptr, indices, data, v, y, rows, indptr[0], 0, (size_t)((rows + 1023) / 1024) };
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, spmv_worker, &job);
    spmv_worker(&job);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
}

/* Matrix-free reference for H.v on arbitrary rows: builds the row with
 * oracle_make_row and dots it with v in stored order -- i.e. "CSR built by
 * the reference algorithm, times v" without materialising the CSR. */
void oracle_apply_rows(const oracle_param *params, size_t n_terms, const uint64_t *rows,
                       size_t n_rows, const c128 *v, c128 *y)
{
    pair_t *scratch = (pair_t *)malloc(2 * n_terms * sizeof(pair_t));
    uint64_t *cols = (uint64_t *)malloc(n_terms * sizeof(uint64_t));
    c128 *vals = (c128 *)malloc(n_terms * sizeof(c128));
    for (size_t i = 0; i < n_rows; i++) {
        size_t k = oracle_make_row_scratch(params, n_terms, rows[i], cols, vals, scratch);
        c128 acc = { 0.0, 0.0 };
        for (size_t q = 0; q < k; q++) acc = cadd(acc, cmul(vals[q], v[cols[q]]));
        y[i] = acc;
    }
    free(scratch); free(cols); free(vals);
}

/* accel.rs:374-393 */
void oracle_axpby(double ar, double ai, const c128 *x, double br, double bi, const c128 *y, c128 *z, size_t n)
{ c128 a = { ar, ai }, b = { br, bi }; for (size_t i = 0; i < n; i++) z[i] = cadd(cmul(a, x[i]), cmul(b, y[i])); }
void oracle_axpy(double ar, double ai, const c128 *x, const c128 *y, c128 *z, size_t n)
{ c128 a = { ar, ai }; for (size_t i = 0; i < n; i++) z[i] = cadd(cmul(a, x[i]), y[i]); }
void oracle_ax(double ar, double ai, const c128 *x, c128 *z, size_t n)
{ c128 a = { ar, ai }; for (size_t i = 0; i < n; i++) z[i] = cmul(a, x[i]); }

/* Davidson preconditioner, pyqrusty/src/lib.rs:436-468: reg() replaces |x| < tol (norm = hypot)
 * by (tol, 0); precond2 returns dx / reg(diag - e).  Complex division as num-complex 0.4.1
 * (Cargo.lock) spells it: re = (a.re*b.re + a.im*b.im) / |b|^2, im = (a.im*b.re - a.re*b.im) / |b|^2,
 * |b|^2 = b.re*b.re + b.im*b.im. */
void oracle_precond2(const c128 *diag, const c128 *dx, double er, double ei, double tol, c128 *out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        c128 x = { diag[i].re - er, diag[i].im - ei };
        if (hypot(x.re, x.im) < tol) { x.re = tol; x.im = 0.0; }
        const double norm_sqr = x.re * x.re + x.im * x.im;
        const double re = dx[i].re * x.re + dx[i].im * x.im;
        const double im = dx[i].im * x.re - dx[i].re * x.im;
        out[i].re = re / norm_sqr; out[i].im = im / norm_sqr;
    }
}

int oracle_hardware_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
