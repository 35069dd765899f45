// apply.cuh -- K4: matrix-free H.v, plus the small kernels either side of it in an
// eigensolver iteration (CSR SpMV on a built shard, diag(H), axpby/axpy/ax, <x,y>).
//
// Reference H.v is CSR SpMV over the built matrix (qrusty/src/accel.rs:338-370):
// it reads 24 B/nnz of matrix plus the gathered v.  The matrix-free apply reads no
// matrix:  y[r] = sum_g value(r,g) * v[r ^ gx[g]]  with value(r,g) as in fill.cuh.
// Compulsory HBM traffic is 32 B/row (v once, y once).
#pragma once
#include "fill.cuh"

namespace qr {

__device__ __forceinline__ double2 ld_nc_double2(const double2 *ptr)
{
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(ptr));
    return r;
}

// ---- cross-GPU synchronisation through flags in IPC-mapped memory (fused distributed apply) -------------------
// Every rank owns {ready[P], done[P]} epoch counters; rank q's array is mapped into every peer.  "ready[me] = epoch" on
// a peer says "my shard of v is complete" (stream order: whatever produced it has finished when the apply kernel starts);
// "done[me] = epoch" says "I no longer read your shard".  No NCCL call and no host round trip on the path.
constexpr int APPLY_MAX_PEERS = 16;

__device__ __forceinline__ void st_release_sys(uint64_t *ptr, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *ptr)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= epoch; a peer that never arrives must not hang the GPU: trap after 20 s
__device__ __forceinline__ void wait_flag(const uint64_t *flag, uint64_t epoch)
{
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > 20000000000ull) __trap();
    }
}

struct ApplyPeerArgs {
    const double2 *peer[APPLY_MAX_PEERS];     // v shard of rank q, pre-offset so that it is indexed by the GLOBAL row id
    uint32_t n_peers, shard_bits;
    // CTA order (local apply on a power-of-two number of row blocks): launch index t -> row block
    // ((t & (2^swz_f - 1)) << (swz_nb - swz_f)) | (t >> swz_f): the TOP swz_f bits of the row-block index vary fastest in
    // time, so the CTAs that are resident together are each other's partners along the top row bits (which otherwise miss L2)
    uint32_t swz_f, swz_nb;
};

// One CTA before the fused apply: tell every peer that this rank's shard is complete (stream order: whatever produced it
// has finished), then wait until the ranks this one reads from have said the same.  The apply kernel follows in stream
// order, so none of its CTAs needs to look at a flag (measured: a system-scope acquire per CTA cost 5 % of the kernel).
__global__ void p2p_ready_kernel(uint64_t *flags_local, uint64_t *const *peer_flags_dev, uint32_t n_peers, uint32_t my_rank,
                                 uint32_t need_mask, uint64_t epoch)
{
    const uint32_t q = threadIdx.x;
    if (q < n_peers && q != my_rank) st_release_sys(peer_flags_dev[q] + my_rank, epoch);
    if (q < n_peers && ((need_mask >> q) & 1u)) wait_flag(flags_local + q, epoch);
}
// ... and one after it: tell every peer that this rank no longer reads their shards, then wait until the ranks that read
// this rank's shard (the same set) have said so -- whatever the stream runs next may overwrite the shard.
__global__ void p2p_done_kernel(uint64_t *flags_local, uint64_t *const *peer_flags_dev, uint32_t n_peers, uint32_t my_rank,
                                uint32_t need_mask, uint64_t epoch)
{
    const uint32_t q = threadIdx.x;
    if (q < n_peers && q != my_rank) st_release_sys(peer_flags_dev[q] + n_peers + my_rank, epoch);
    if (q < n_peers && ((need_mask >> q) & 1u)) wait_flag(flags_local + n_peers + q, epoch);
}

// v0: lane <-> row, walk the groups.  For 32 aligned consecutive rows the gather
// v[r ^ x] is one aligned 512-byte segment with lanes permuted, so every load is
// fully coalesced; re-use across groups is left to L1/L2.
constexpr int APPLY_THREADS = 256;
constexpr int APPLY_ROWS = 4;          // rows per thread: group descriptors are read once per 4 rows
constexpr int APPLY_BATCH = 128;       // descriptors staged in shared memory at a time (4 KB)

// acc += a * w, with the 2-FMA form when a is known to be real (warp-uniform flag).  The FMA sequence is spelled out so
// that every kernel that walks the groups in the same order (gather, tiled, fused distributed) rounds identically: left
// to the compiler, the contraction of `ar*w.x - ai*w.y` differs from one instantiation to the next.
__device__ __forceinline__ void cfma(double &yr, double &yi, double ar, double ai, double2 w, bool a_real)
{
    yr = __fma_rn(ar, w.x, yr);
    yi = __fma_rn(ar, w.y, yi);
    if (!a_real) { yr = __fma_rn(-ai, w.y, yr); yi = __fma_rn(ai, w.x, yi); }
}

// <v, y> over the rows of a CTA, folded in a fixed order (warp shuffles, then warp 0..W-1 in sequence): the epilogue of
// the apply kernels when the caller wants alpha = <v, H v> with the product (a Lanczos step then needs no separate pass
// over v and y).  All threads of the CTA must call; out = one double2 per CTA.
__device__ __forceinline__ void block_dot_store(double dr, double di, double2 *out)
{
    __shared__ double s_dre[32], s_dim[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { dr += __shfl_xor_sync(0xffffffffu, dr, d); di += __shfl_xor_sync(0xffffffffu, di, d); }
    if ((threadIdx.x & 31u) == 0u) { s_dre[threadIdx.x >> 5] = dr; s_dim[threadIdx.x >> 5] = di; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t w = 1; w < (blockDim.x + 31u) / 32u; w++) { dr += s_dre[w]; di += s_dim[w]; }
        *out = make_double2(dr, di);
    }
}
// conj(v) * y accumulated into (dr, di)
__device__ __forceinline__ void cdot_acc(double &dr, double &di, double2 v, double yr, double yi)
{
    dr = __fma_rn(v.x, yr, dr); dr = __fma_rn(v.y, yi, dr);
    di = __fma_rn(v.x, yi, di); di = __fma_rn(-v.y, yr, di);
}

// v0 (gather): thread <-> APPLY_ROWS rows (r, r+256, ...), so a warp's loads of v[r ^ x] stay one
// permuted, fully coalesced 512-byte segment.  Group descriptors come through shared memory.
// `diag` / `diag_re` (optional): cached values of the mask-0 group for rows [row_lo,row_hi) -- complex, or the real parts
// only when every c' of that group is real (half the bytes); when given, group 0 is not re-evaluated (an eigensolver
// applies the same operator hundreds of times).
// Row-sharded form (pa.n_peers > 1): the shard that owns v[r ^ x] is rank ^ (x >> shard_bits) for every row of this rank,
// so the base pointer is a per-group constant (staged beside the descriptor) and remote shards are read in place over
// NVLink -- the collective is fused into the apply.  p2p_ready_kernel / p2p_done_kernel bracket it.
// One kernel for both forms: the unsharded and the distributed apply run the same instruction sequence and agree bit for
// bit.  Measured and dropped (profiles/r04_summary.md): requesting the remote partners first with cp.async into shared
// memory (2.3x slower: LDGSTS to peer memory), pulling them by TMA into a tile (apply_tile.cuh, slower at 2 GPUs).
__global__ void __launch_bounds__(APPLY_THREADS, 4)
apply_direct_kernel(PlanDev p, uint32_t G, uint64_t row_lo, uint64_t row_hi,
                    const double2 *__restrict__ v, double2 *__restrict__ y,
                    const double2 *__restrict__ diag, const double *__restrict__ diag_re,
                    const __grid_constant__ ApplyPeerArgs pa,
                    const uint32_t *__restrict__ glist = nullptr, uint32_t n_list = 0,
                    double2 *__restrict__ dotp = nullptr)                 // optional: per-CTA partials of <v, y> over these rows
{
    // glist (optional): apply only these groups (ascending ids; the mask-0 group, when cached in diag, first) --
    // the NEAR pass of the two-pass apply (apply_tile.cuh)
    if (glist != nullptr) G = n_list;
    constexpr int E = APPLY_ROWS;
    __shared__ GroupDesc sd[APPLY_BATCH];
    __shared__ const double2 *sv[APPLY_BATCH];
    const bool PEERS = pa.n_peers > 1u;
    const uint32_t blk = pa.swz_f ? ((blockIdx.x & ((1u << pa.swz_f) - 1u)) << (pa.swz_nb - pa.swz_f)) | (blockIdx.x >> pa.swz_f) : blockIdx.x;
    const uint64_t cta_base = row_lo + (uint64_t)blk * (APPLY_THREADS * E);
    const uint32_t my_rank = PEERS ? (uint32_t)(row_lo >> pa.shard_bits) : 0u;
    const double2 *v_own = PEERS ? pa.peer[my_rank] : v;
    uint32_t r[E];
    bool live[E];
    double yr[E], yi[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint64_t r64 = cta_base + (uint64_t)e * APPLY_THREADS + threadIdx.x;
        live[e] = r64 < row_hi;
        r[e] = (uint32_t)(live[e] ? r64 : row_hi - 1);          // clamp: dead rows recompute a live one
        yr[e] = 0.0; yi[e] = 0.0;
    }
    const uint32_t G_main = G;
    uint32_t g_first = 0;
    if (diag_re != nullptr) {                                      // real diagonal (every c' of the mask-0 group is real): 8 B per row
        g_first = 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const double d = __ldcs(&diag_re[(uint64_t)r[e] - row_lo]);     // evict-first: keep L2 for v
            cfma(yr[e], yi[e], d, 0.0, ld_nc_double2(&v_own[r[e]]), true);
        }
    } else if (diag != nullptr) {
        g_first = 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const double2 d = __ldcs(&diag[(uint64_t)r[e] - row_lo]);      // evict-first: keep L2 for v
            cfma(yr[e], yi[e], d.x, d.y, ld_nc_double2(&v_own[r[e]]), false);
        }
    }
    for (uint32_t g0 = g_first; g0 < G_main; g0 += APPLY_BATCH) {
        const uint32_t nb = min((uint32_t)APPLY_BATCH, G_main - g0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb; i += APPLY_THREADS) {
            const GroupDesc d = p.gdesc[glist != nullptr ? __ldg(&glist[g0 + i]) : g0 + i];
            sd[i] = d;
            sv[i] = PEERS ? pa.peer[my_rank ^ (d.x >> pa.shard_bits)] : v;
        }
        __syncthreads();
        for (uint32_t k = 0; k < nb; k++) {
            const GroupDesc d = sd[k];
            const double2 *vb = sv[k];
            const bool real = (d.flag & 2u) != 0u;
            if (d.flag & 1u) {
#pragma unroll
                for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, d.cim, ld_nc_double2(&vb[r[e] ^ d.x]), real);
            } else {
                double ar[E], ai[E];
                group_values<E>(p, d.t0, d.t1, r, ar, ai);
#pragma unroll
                for (int e = 0; e < E; e++) cfma(yr[e], yi[e], ar[e], ai[e], ld_nc_double2(&vb[r[e] ^ d.x]), real);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++)
        if (live[e]) __stcs(&y[(uint64_t)r[e] - row_lo], make_double2(yr[e], yi[e]));   // streaming store
    if (dotp != nullptr) {                                         // CTA-uniform
        double dr = 0.0, di = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++)
            if (live[e]) cdot_acc(dr, di, ld_nc_double2(&v_own[r[e]]), yr[e], yi[e]);
        block_dot_store(dr, di, &dotp[blk]);
    }
}

// ---------------------------------------------------------------------------------
// v1: multi-pass shared-memory tiling.
//
// v0 moves 16*(G+1) bytes per row through L1/L2.  Here the groups are split into passes;
// pass q owns a set S_q of K bit positions (bits 0..4 plus K-5 higher bits of the
// LOCAL row index) and every group whose X-mask lies inside S_q.  A CTA takes one
// tile = the 2^K elements of v whose bits outside S_q are fixed (2^(K-5) runs of 32
// consecutive elements, 512 B each), loads it into shared memory once, and for each of
// its rows accumulates sum_g a_g(r) * tile[idx ^ cx_g] where idx is the row's index inside
// the tile and cx_g the mask compacted onto S_q.  Pass 0 writes y, later passes add to
// it, so HBM/L2 traffic is 32 + 48*(passes-1) bytes per row instead of 16*(G+1).
// Groups that fit no pass (mask wider than K-5 high bits, or touching bits above the
// local row block in a row-sharded apply) are gathered from global memory in pass 0.
// ---------------------------------------------------------------------------------
struct ApplyPass {
    const uint32_t *groups;     // group ids applied from shared memory in this pass
    const uint32_t *cmask;      // their masks compacted onto S (tile-index space)
    uint32_t n_groups;
    const uint32_t *direct;     // pass 0 only: groups gathered from global memory
    uint32_t n_direct;
    const uint32_t *expand;     // [2^(K-5)]: tile index bits >= 5 -> row bits (S positions)
    uint32_t free_mask;         // local row bits NOT in S (enumerated by blockIdx)
    uint32_t first;             // 1: y = acc, 0: y += acc
    const double2 *diag;        // pass 0 only, optional: cached mask-0 group (then not in `groups`)
};

template <int K, int THREADS>
__global__ void __launch_bounds__(THREADS)
apply_pass_kernel(PlanDev p, ApplyPass ps, uint64_t row_lo, const double2 *__restrict__ v,
                  double2 *__restrict__ y)
{
    constexpr uint32_t TILE = 1u << K, RUNS = TILE / 32u;
    constexpr int E = APPLY_ROWS;
    constexpr int CHUNKS = TILE / (THREADS * E);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);                               // [2^K]
    uint32_t *s_exp = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TILE * 16);          // [2^(K-5)]
    GroupDesc *sd = reinterpret_cast<GroupDesc *>(smem_raw + (size_t)TILE * 16 + RUNS * 4); // [APPLY_BATCH]

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // scatter blockIdx's bits onto the free (non-S) positions of the local row index
    uint32_t fixed = 0, rest = blockIdx.x, fm = ps.free_mask;
    while (fm) { const uint32_t b = fm & (0u - fm); if (rest & 1u) fixed |= b; rest >>= 1; fm ^= b; }
    const uint32_t base = (uint32_t)row_lo + fixed;                      // row_lo is block-aligned

    for (uint32_t j = threadIdx.x; j < RUNS; j += THREADS) s_exp[j] = __ldg(&ps.expand[j]);
    __syncthreads();
    for (uint32_t j = warp; j < RUNS; j += THREADS / 32)
        tile[j * 32u + lane] = ld_nc_double2(&v[(base | s_exp[j]) + lane]);

    const uint32_t n_all = ps.n_groups + ps.n_direct;
    for (int c = 0; c < CHUNKS; c++) {
        uint32_t idx[E], r[E];
        double yr[E], yi[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
            idx[e] = (uint32_t)(c * E + e) * THREADS + threadIdx.x;
            yr[e] = 0.0; yi[e] = 0.0;
        }
        for (uint32_t g0 = 0; g0 < n_all || g0 == 0; g0 += APPLY_BATCH) {
            const uint32_t nb = min((uint32_t)APPLY_BATCH, n_all - g0);
            __syncthreads();                                     // tile loaded / previous batch consumed
            for (uint32_t i = threadIdx.x; i < nb; i += THREADS) {
                const uint32_t k = g0 + i;
                GroupDesc d;
                if (k < ps.n_groups) { d = p.gdesc[__ldg(&ps.groups[k])]; d.x = __ldg(&ps.cmask[k]); }
                else { d = p.gdesc[__ldg(&ps.direct[k - ps.n_groups])]; d.flag |= 4u; }      // bit2: gather from global
                sd[i] = d;
            }
            __syncthreads();
            if (g0 == 0) {
#pragma unroll
                for (int e = 0; e < E; e++) r[e] = (base | s_exp[idx[e] >> 5]) + (idx[e] & 31u);
                if (ps.diag != nullptr) {
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double2 dg = ps.diag[(uint64_t)r[e] - row_lo];
                        cfma(yr[e], yi[e], dg.x, dg.y, tile[idx[e]], false);
                    }
                }
            }
            for (uint32_t k = 0; k < nb; k++) {
                const GroupDesc d = sd[k];
                const bool real = (d.flag & 2u) != 0u;
                double ar[E], ai[E];
                if (d.flag & 1u) {
#pragma unroll
                    for (int e = 0; e < E; e++) { ar[e] = d.cre; ai[e] = d.cim; }
                } else {
                    group_values<E>(p, d.t0, d.t1, r, ar, ai);
                }
                if (d.flag & 4u) {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], ar[e], ai[e], ld_nc_double2(&v[r[e] ^ d.x]), real);
                } else {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], ar[e], ai[e], tile[idx[e] ^ d.x], real);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            double2 *out = &y[(uint64_t)r[e] - row_lo];
            if (!ps.first) { const double2 o = *out; yr[e] += o.x; yi[e] += o.y; }
            *out = make_double2(yr[e], yi[e]);
        }
    }
}

// diag(H): only the group with X-mask 0 (gx[0], masks are ascending) touches the diagonal.
__global__ void __launch_bounds__(256)
diagonal_kernel(PlanDev p, uint64_t row_lo, uint64_t row_hi, double2 *__restrict__ diag, double *__restrict__ diag_re = nullptr)
{
    const uint64_t r64 = row_lo + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r64 >= row_hi) return;
    double2 d = make_double2(0.0, 0.0);
    if (__ldg(&p.gx[0]) == 0u) d = group_value(p.tz, p.tc, 0u, __ldg(&p.goff[1]), (uint32_t)r64);
    if (diag_re != nullptr) diag_re[r64 - row_lo] = d.x;          // real group: the imaginary part is a sum of +-0
    else diag[r64 - row_lo] = d;
}

// diag of a stored CSR shard (pyqrusty/src/lib.rs:118-125 reads it from the stored matrix: after scale() or
// eliminate_zeros() the plan no longer describes the values).  Row i of the shard is row col0 + i of the matrix:
// diag[i] = the stored entry of row i whose column is col0 + i, else 0.  Columns ascend inside a row: binary search.
__global__ void __launch_bounds__(256)
csr_diagonal_kernel(uint64_t n_rows, uint64_t col0, const uint64_t *__restrict__ indptr, const uint64_t *__restrict__ indices,
                    const double2 *__restrict__ data, double2 *__restrict__ diag)
{
    const uint64_t row = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t base = indptr[0], want = col0 + row;
    uint64_t lo = indptr[row] - base, hi = indptr[row + 1] - base;
    double2 d = make_double2(0.0, 0.0);
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1, c = indices[mid];
        if (c == want) { d = data[mid]; break; }
        if (c < want) lo = mid + 1; else hi = mid;
    }
    diag[row] = d;
}

// CSR SpMV as rowwise::spmat_dot_densevec does it (accel.rs:355-364): one row per
// thread, products accumulated sequentially in stored order starting from zero,
// with the multiply written out as num_complex does (no FMA contraction) so the
// result is bit-identical to the CPU.
__global__ void __launch_bounds__(256)
spmv_csr_kernel(uint64_t n_rows, const uint64_t *__restrict__ indptr,
                const uint64_t *__restrict__ indices, const double2 *__restrict__ data,
                const double2 *__restrict__ v, double2 *__restrict__ y)
{
    const uint64_t row = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t base = indptr[0];
    double re = 0.0, im = 0.0;
    for (uint64_t k = indptr[row] - base; k < indptr[row + 1] - base; k++) {
        const double2 a = data[k];
        const double2 w = ld_nc_double2(&v[indices[k]]);
        const double pr = __dsub_rn(__dmul_rn(a.x, w.x), __dmul_rn(a.y, w.y));
        const double pi = __dadd_rn(__dmul_rn(a.x, w.y), __dmul_rn(a.y, w.x));
        re = __dadd_rn(re, pr);
        im = __dadd_rn(im, pi);
    }
    y[row] = make_double2(re, im);
}

// Thread <-> row with U entries of the row in flight at a time (their loads are independent of the running sum; the
// additions stay sequential in stored order, so the sums round exactly as the reference's loop).  U = 4 is the default:
// C2 0.143 ms against 0.183 for the plain loop above, XXZ n = 24 1.33 / 1.70, H8 0.52 / 0.73, C3 1.32 / 1.45
// (profiles/r06_spmv3.jsonl; 3.0-3.8 TB/s of matrix bytes, 5.0-6.3 TB/s counting the gathered v).  Measured and dropped:
// tiles through shared memory (CTA- or warp-owned row blocks, lane <-> entry loads in memory order, then lane <-> row sums):
// 0.24-0.39 ms on C2 -- sixteen warps per SM do not keep enough loads in flight; sixty-four independent row streams do.
template <int U>
__global__ void __launch_bounds__(256)
spmv_csr_unrolled_kernel(uint64_t n_rows, const uint64_t *__restrict__ indptr,
                         const uint64_t *__restrict__ indices, const double2 *__restrict__ data,
                         const double2 *__restrict__ v, double2 *__restrict__ y)
{
    const uint64_t row = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t base = indptr[0], k1 = indptr[row + 1] - base;
    uint64_t k = indptr[row] - base;
    double re = 0.0, im = 0.0;
    for (; k + U <= k1; k += U) {
        double2 a[U], w[U];
        uint64_t col[U];
#pragma unroll
        for (int u = 0; u < U; u++) { col[u] = __ldg(&indices[k + u]); a[u] = ld_nc_double2(&data[k + u]); }
#pragma unroll
        for (int u = 0; u < U; u++) w[u] = ld_nc_double2(&v[col[u]]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            re = __dadd_rn(re, __dsub_rn(__dmul_rn(a[u].x, w[u].x), __dmul_rn(a[u].y, w[u].y)));
            im = __dadd_rn(im, __dadd_rn(__dmul_rn(a[u].x, w[u].y), __dmul_rn(a[u].y, w[u].x)));
        }
    }
    for (; k < k1; k++) {
        const double2 a = ld_nc_double2(&data[k]);
        const double2 w = ld_nc_double2(&v[__ldg(&indices[k])]);
        re = __dadd_rn(re, __dsub_rn(__dmul_rn(a.x, w.x), __dmul_rn(a.y, w.y)));
        im = __dadd_rn(im, __dadd_rn(__dmul_rn(a.x, w.y), __dmul_rn(a.y, w.x)));
    }
    y[row] = make_double2(re, im);
}

// accel.rs:374-393, arithmetic spelled as num_complex's so results are bit-identical.
__device__ __forceinline__ double2 cmul_rn(double2 a, double2 b)
{
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
template <int MODE>   // 0: a*x + b*y   1: a*x + y   2: a*x
__global__ void __launch_bounds__(256)
vec_axpby_kernel(uint64_t n, double2 a, const double2 *__restrict__ x, double2 b,
                 const double2 *__restrict__ y, double2 *__restrict__ z)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        double2 t = cmul_rn(a, x[i]);
        if (MODE == 0) { double2 u = cmul_rn(b, y[i]); t = make_double2(__dadd_rn(t.x, u.x), __dadd_rn(t.y, u.y)); }
        if (MODE == 1) { double2 u = y[i]; t = make_double2(__dadd_rn(t.x, u.x), __dadd_rn(t.y, u.y)); }
        z[i] = t;
    }
}

// Davidson preconditioner (pyqrusty/src/lib.rs:436-468): out = dx / reg(diag - e, tol), where reg()
// replaces a denominator with norm() < tol by (tol, 0).  Division spelled as num-complex's (no FMA
// contraction), so the result is bit-identical to the CPU for every element whose |diag - e| is not
// within an ulp of tol.  32 B read + 16 B written per element.
__global__ void __launch_bounds__(256)
precond2_kernel(uint64_t n, const double2 *__restrict__ diag, const double2 *__restrict__ dx, double2 e, double tol,
                double2 *__restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const double2 d = diag[i], a = dx[i];
        double xr = __dsub_rn(d.x, e.x), xi = __dsub_rn(d.y, e.y);
        if (hypot(xr, xi) < tol) { xr = tol; xi = 0.0; }
        const double norm_sqr = __dadd_rn(__dmul_rn(xr, xr), __dmul_rn(xi, xi));
        const double re = __dadd_rn(__dmul_rn(a.x, xr), __dmul_rn(a.y, xi));
        const double im = __dsub_rn(__dmul_rn(a.y, xr), __dmul_rn(a.x, xi));
        out[i] = make_double2(__ddiv_rn(re, norm_sqr), __ddiv_rn(im, norm_sqr));
    }
}

// Lanczos three-term update fused with the norm: w_out = w - alpha*v - beta*v_prev and
// partial[b] = (sum |w_out|^2, 0) per CTA (folded by dotc_final_kernel).  One pass over three
// vectors instead of two axpy passes and a dot product.
__global__ void __launch_bounds__(256)
lanczos_update_kernel(uint64_t n, double2 alpha, double2 beta, const double2 *w,      // w_out may alias w
                      const double2 *__restrict__ v, const double2 *__restrict__ v_prev,
                      double2 *w_out, double2 *__restrict__ partial)
{
    __shared__ double sre[8];
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const double2 a = v[i], b = v_prev ? v_prev[i] : make_double2(0.0, 0.0), c = w[i];
        const double re = c.x - (alpha.x * a.x - alpha.y * a.y) - (beta.x * b.x - beta.y * b.y);
        const double im = c.y - (alpha.x * a.y + alpha.y * a.x) - (beta.x * b.y + beta.y * b.x);
        w_out[i] = make_double2(re, im);
        acc += re * re + im * im;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) sre[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) acc += sre[k];
        partial[blockIdx.x] = make_double2(acc, 0.0);
    }
}

// ---- Lanczos with the scalars on the device (no host round trip inside an iteration) ----------------------------
// The vectors are kept UNNORMALISED: u_{k+1} = w_k, v_{k+1} = u_{k+1} / beta_k, so no pass is spent on a rescale:
//   alpha_k = <u_k, H u_k> / bcur^2                      bcur = ||u_k|| (1 for the start vector)
//   w_k     = (H u_k) / bcur - (alpha_k / bcur) u_k - (bcur / bprev) u_{k-1}
//   beta_k  = ||w_k||
// state (doubles): [0,1] <u, H u> (re, im)   [2,3] ||w||^2 (, 0)   [4] bprev   [5] bcur   [6] ca   [7] cb   [8] cc
//                  [16 + k] alpha_k          [16 + K + k] beta_k
constexpr int LZ_STATE_HEAD = 16;
__global__ void lanczos_coef_kernel(double *st, uint32_t k, uint32_t K, uint32_t phase)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (phase == 0u) {                                             // after the apply: alpha_k and the update's coefficients
        const double bcur = st[5], bprev = st[4];
        const double alpha = st[0] / (bcur * bcur);
        st[6] = 1.0 / bcur; st[7] = alpha / bcur; st[8] = k == 0u ? 0.0 : bcur / bprev;
        st[LZ_STATE_HEAD + k] = alpha;
    } else {                                                       // after the update: beta_k, and u_{k+1}'s scale
        const double beta = sqrt(st[2]);
        st[LZ_STATE_HEAD + K + k] = beta;
        st[4] = st[5]; st[5] = beta;
    }
}
// w_out = ca * y - cb * u - cc * u_prev (real coefficients read from the device state), partial[b] = sum |w_out|^2
__global__ void __launch_bounds__(256)
lanczos_update_dev_kernel(uint64_t n, const double *__restrict__ st, const double2 *__restrict__ y, const double2 *__restrict__ u,
                          const double2 *__restrict__ u_prev, double2 *__restrict__ w_out, double2 *__restrict__ partial)
{
    __shared__ double sre[8];
    const double ca = st[6], cb = st[7], cc = st[8];
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const double2 a = y[i], b = u[i], c = u_prev ? u_prev[i] : make_double2(0.0, 0.0);
        const double re = ca * a.x - cb * b.x - cc * c.x, im = ca * a.y - cb * b.y - cc * c.y;
        __stcs(&w_out[i], make_double2(re, im));
        acc += re * re + im * im;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) sre[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) acc += sre[k];
        partial[blockIdx.x] = make_double2(acc, 0.0);
    }
}

// <x,y> = sum conj(x_i) y_i : per-CTA partials, then one CTA folds them (deterministic).
constexpr int DOT_THREADS = 256;
__global__ void __launch_bounds__(DOT_THREADS)
dotc_partial_kernel(uint64_t n, const double2 *__restrict__ x, const double2 *__restrict__ y,
                    double2 *__restrict__ partial)
{
    __shared__ double sre[DOT_THREADS / 32], sim[DOT_THREADS / 32];
    double re = 0.0, im = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * DOT_THREADS) {
        const double2 a = x[i], b = y[i];
        re += a.x * b.x + a.y * b.y;
        im += a.x * b.y - a.y * b.x;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, d); im += __shfl_xor_sync(0xffffffffu, im, d); }
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < DOT_THREADS / 32; w++) { re += sre[w]; im += sim[w]; }
        partial[blockIdx.x] = make_double2(re, im);
    }
}
__global__ void __launch_bounds__(DOT_THREADS)
dotc_final_kernel(uint32_t n_partials, const double2 *__restrict__ partial, double2 *__restrict__ out)
{
    __shared__ double sre[DOT_THREADS / 32], sim[DOT_THREADS / 32];
    double re = 0.0, im = 0.0;
    for (uint32_t i = threadIdx.x; i < n_partials; i += DOT_THREADS) { re += partial[i].x; im += partial[i].y; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, d); im += __shfl_xor_sync(0xffffffffu, im, d); }
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < DOT_THREADS / 32; w++) { re += sre[w]; im += sim[w]; }
        *out = make_double2(re, im);
    }
}

}  // namespace qr
