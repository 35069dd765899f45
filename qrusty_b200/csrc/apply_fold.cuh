// apply_fold.cuh -- K4b: matrix-free H.v for term-rich operators (molecular Hamiltonians: 5-6 Pauli strings per
// X-mask, a few groups with hundreds).
//
// The gather kernel (apply.cuh) evaluates value(r, g) = sum_t c'_t (-1)^popc(r & z_t) term by term for every row:
// T evaluations per row, ~7 instructions each, and H.v on the reference's H12 fixture (T = 4 497, 2^24 rows) is
// bound by instruction issue at 2.2e12 evaluations/s.  Here a thread owns the E = 8 rows that differ in row bits
// B0, B0+1, B0+2 (B0 = log2(threads per CTA)); writing r = r0 ^ (e << B0),
//     popc(r & z) = popc(r0 & z) + popc(e & q),   q = (z >> B0) & 7,
// so with the group's terms bucketed by q (host, once per plan)
//     S_q      = sum_{t in bucket q} c'_t (-1)^popc(r0 & z_t)         one evaluation per term and THREAD, not per row
//     value(e) = sum_q S_q (-1)^popc(e & q)                           a 3-stage Walsh-Hadamard butterfly in registers
// i.e. T_g + 24 additions per group and thread instead of 8 T_g.  Groups of up to FOLD_WHT_MIN terms skip the butterfly
// (it would cost more than it saves) and add every term into the 8 row values with the sign pattern of its bucket;
// row-independent groups use gconst as before.
//
// The order of the additions differs from the reference's left-to-right fold, so the values agree to rounding
// (|d - d_ref| <= ~T_g * 2^-53 * sum|c'|), not bit for bit: this is H.v only -- its contract is the 1e-12 of
// north_star; the CSR fill kernels keep the reference's fold order and stay bit-exact.  (accel.rs:338-370 is SpMV over
// the stored matrix; there is no matrix-free apply in the reference to be bit-identical to.)
#pragma once
#include "apply_tile.cuh"

namespace qr {

// Bucketed copy of the term table (same group ranges [goff[g], goff[g+1]) as PlanDev::tz / tc):
//   zc[t]   = {z, q, lo32(re c'), hi32(re c')}      one 16-byte load per term; q = (z >> B0) & 7, its bucket
//   im[t]   = im c'                                  read for groups that are not real (gflag bit1 clear)
//   bend[g] = 8 x u16: end of bucket q relative to the group's first term (bucket q = terms [bend[q-1], bend[q]))
// B0 is fixed per table (the kernel's thread count).
struct FoldDev {
    const uint4  *zc;
    const double *im;
    const uint4  *bend;
};

constexpr int FOLD_THREADS = 128;
constexpr int FOLD_ROWS = 8;
constexpr int FOLD_B0 = 7;                 // log2(FOLD_THREADS)
constexpr int FOLD_BATCH = 64;             // group descriptors staged in shared memory at a time
constexpr uint32_t FOLD_MAX_GROUP_TERMS = 65535;

__device__ __forceinline__ double signed_re(const uint4 w, uint32_t s) { return __hiloint2double((int)(w.w ^ s), (int)w.z); }

// in-register Walsh-Hadamard transform over the 3 row bits a thread owns: S[e] <- sum_q S[q] (-1)^popc(e & q)
__device__ __forceinline__ void wht8(double (&S)[8])
{
#pragma unroll
    for (int b = 1; b < 8; b <<= 1)
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (!(i & b)) { const double a = S[i], c = S[i | b]; S[i] = a + c; S[i | b] = a - c; }
}

// Short groups (<= FOLD_WHT_MIN terms): every term is added straight into the thread's 8 row values, branch-free.  One
// POPC per term gives the parity of row e = 0; row e's parity adds the bits of z at the thread's row bits that e has set
// (XOR of pre-shifted copies of z), and fma(+-1.0, c, S) is the sign flip and the addition in one instruction.  IM selects
// the component: groups that are not real are swept twice (re, then im), so one set of 8 accumulators serves both.
constexpr uint32_t FOLD_WHT_MIN = 12;      // groups with more terms go through the bucket sums + Walsh-Hadamard butterfly
// `zcs` = the group's entries of FoldDev::zc (global memory, or the copy the fold kernel stages in shared memory),
// `ims` = its entries of FoldDev::im.
template <bool IM>
__device__ __forceinline__ void fold_short_sweep(const uint4 *zcs, const double *ims, uint32_t n_t, uint32_t r0, double (&S)[8])
{
#pragma unroll
    for (int e = 0; e < 8; e++) S[e] = 0.0;
#pragma unroll 2
    for (uint32_t t = 0; t < n_t; t++) {
        const uint4 zc = zcs[t];
        const double c = IM ? __ldg(&ims[t]) : __hiloint2double((int)zc.w, (int)zc.z);
        const uint32_t p0 = (uint32_t)__popc(r0 & zc.x);
        const uint32_t z1 = zc.x >> FOLD_B0, z2 = zc.x >> (FOLD_B0 + 1), z3 = zc.x >> (FOLD_B0 + 2);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            uint32_t par = p0;
            if (e & 1) par ^= z1;
            if (e & 2) par ^= z2;
            if (e & 4) par ^= z3;
            S[e] = __fma_rn(pm_one(par), c, S[e]);
        }
    }
}

__global__ void __launch_bounds__(FOLD_THREADS, 4)
apply_fold_kernel(PlanDev p, FoldDev f, uint32_t G, uint64_t row_lo, uint64_t row_hi,
                  const double2 *__restrict__ v, double2 *__restrict__ y,
                  const double2 *__restrict__ diag, const double *__restrict__ diag_re,
                  const __grid_constant__ ApplyPeerArgs pa, double2 *__restrict__ dotp)
{
    constexpr int E = FOLD_ROWS, TH = FOLD_THREADS, B0 = FOLD_B0;
    __shared__ GroupDesc sd[FOLD_BATCH];
    __shared__ uint4 sb[FOLD_BATCH];
    __shared__ const double2 *sv[FOLD_BATCH];
    __shared__ uint4 sterm[FOLD_BATCH * FOLD_WHT_MIN];             // the short groups' terms: warp-uniform LDS instead of an L1 round trip per term
    const bool PEERS = pa.n_peers > 1u;
    // host-checked: row_lo and row_hi - row_lo are multiples of E * TH, so bits B0..B0+2 of r0 are clear
    const uint32_t r0 = (uint32_t)(row_lo + (uint64_t)blockIdx.x * (TH * E)) + threadIdx.x;
    const uint32_t my_rank = PEERS ? (uint32_t)(row_lo >> pa.shard_bits) : 0u;
    const double2 *v_own = PEERS ? pa.peer[my_rank] : v;
    double yr[E], yi[E];
#pragma unroll
    for (int e = 0; e < E; e++) { yr[e] = 0.0; yi[e] = 0.0; }
    uint32_t g_first = 0;
    if (diag_re != nullptr) {
        g_first = 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t r = r0 + ((uint32_t)e << B0);
            cfma(yr[e], yi[e], __ldcs(&diag_re[(uint64_t)r - row_lo]), 0.0, ld_nc_double2(&v_own[r]), true);
        }
    } else if (diag != nullptr) {
        g_first = 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t r = r0 + ((uint32_t)e << B0);
            const double2 d = __ldcs(&diag[(uint64_t)r - row_lo]);
            cfma(yr[e], yi[e], d.x, d.y, ld_nc_double2(&v_own[r]), false);
        }
    }
    for (uint32_t g0 = g_first; g0 < G; g0 += FOLD_BATCH) {
        const uint32_t nb = min((uint32_t)FOLD_BATCH, G - g0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb; i += TH) {
            const GroupDesc d = p.gdesc[g0 + i];
            sd[i] = d;
            sb[i] = __ldg(&f.bend[g0 + i]);
            sv[i] = PEERS ? pa.peer[my_rank ^ (d.x >> pa.shard_bits)] : v;
        }
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < nb * FOLD_WHT_MIN; q += TH) {
            const uint32_t i = q / FOLD_WHT_MIN, j = q - i * FOLD_WHT_MIN, t0 = sd[i].t0, n = sd[i].t1 - t0;
            if (n <= FOLD_WHT_MIN && j < n && !(sd[i].flag & 1u)) sterm[q] = __ldg(&f.zc[t0 + j]);
        }
        __syncthreads();
        for (uint32_t k = 0; k < nb; k++) {
            const GroupDesc d = sd[k];
            const double2 *vb = sv[k];
            const bool real = (d.flag & 2u) != 0u;
            const uint32_t n_t = d.t1 - d.t0;
            if (d.flag & 1u) {                                     // row-independent value
                if (real) {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, 0.0, ld_nc_double2(&vb[(r0 + ((uint32_t)e << B0)) ^ d.x]), true);
                } else {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, d.cim, ld_nc_double2(&vb[(r0 + ((uint32_t)e << B0)) ^ d.x]), false);
                }
                continue;
            }
            // the partners are requested before the fold: their latency hides behind it
            double2 w[E];
#pragma unroll
            for (int e = 0; e < E; e++) w[e] = ld_nc_double2(&vb[(r0 + ((uint32_t)e << B0)) ^ d.x]);
            double S[E];
            if (n_t <= FOLD_WHT_MIN) {
                fold_short_sweep<false>(&sterm[k * FOLD_WHT_MIN], f.im + d.t0, n_t, r0, S);
#pragma unroll
                for (int e = 0; e < E; e++) { yr[e] = __fma_rn(S[e], w[e].x, yr[e]); yi[e] = __fma_rn(S[e], w[e].y, yi[e]); }
                if (!real) {
                    fold_short_sweep<true>(&sterm[k * FOLD_WHT_MIN], f.im + d.t0, n_t, r0, S);
#pragma unroll
                    for (int e = 0; e < E; e++) { yr[e] = __fma_rn(-S[e], w[e].y, yr[e]); yi[e] = __fma_rn(S[e], w[e].x, yi[e]); }
                }
                continue;
            }
            const uint4 be = sb[k];
            const uint32_t ends[8] = {be.x & 0xffffu, be.x >> 16, be.y & 0xffffu, be.y >> 16,
                                      be.z & 0xffffu, be.z >> 16, be.w & 0xffffu, be.w >> 16};
            uint32_t t = d.t0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                double acc = 0.0;
                for (const uint32_t te = d.t0 + ends[q]; t < te; t++) {
                    const uint4 zc = __ldg(&f.zc[t]);
                    acc += signed_re(zc, (uint32_t)__popc(r0 & zc.x) << 31);
                }
                S[q] = acc;
            }
            wht8(S);
#pragma unroll
            for (int e = 0; e < E; e++) { yr[e] = __fma_rn(S[e], w[e].x, yr[e]); yi[e] = __fma_rn(S[e], w[e].y, yi[e]); }
            if (!real) {                                           // imaginary parts: the same buckets, a second sweep
                t = d.t0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    double acc = 0.0;
                    for (const uint32_t te = d.t0 + ends[q]; t < te; t++) {
                        const uint32_t z = __ldg(&f.zc[t]).x;
                        acc += flip_sign(__ldg(&f.im[t]), (uint32_t)__popc(r0 & z) << 31);
                    }
                    S[q] = acc;
                }
                wht8(S);
#pragma unroll
                for (int e = 0; e < E; e++) { yr[e] = __fma_rn(-S[e], w[e].y, yr[e]); yi[e] = __fma_rn(S[e], w[e].x, yi[e]); }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) __stcs(&y[(uint64_t)(r0 + ((uint32_t)e << B0)) - row_lo], make_double2(yr[e], yi[e]));
    if (dotp != nullptr) {                                         // per-CTA partial of <v, y> (apply.cuh)
        double dr = 0.0, di = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++) cdot_acc(dr, di, ld_nc_double2(&v_own[r0 + ((uint32_t)e << B0)]), yr[e], yi[e]);
        block_dot_store(dr, di, &dotp[blockIdx.x]);
    }
}

// ---------------------------------------------------------------------------------
// K4c: partner tiles.  The gather kernels move 16 B per (row, group) through L2 (ncu, C4: 11.9 GB for 1.3 GB of
// compulsory traffic; the kernel runs at the L2's ~15 TB/s).  But the masks of real operators share their high bits:
// a lattice model's single- and two-bit masks, a molecular Hamiltonian's double excitations (H10: 2 536 masks, 163
// distinct x >> 12; H12: 811 / 138; TFIM 5x5: 26 / 14).  A CTA owns one aligned tile of 2^K rows; the groups, sorted
// by mask, fall into SEGMENTS of equal h = x >> K, and every group of a segment reads the same partner tile
// (tile ^ h): 2^K * 16 contiguous bytes, pulled ONCE per segment by the TMA (cp.async.bulk, mbarrier expect-tx) into a
// ring of NBUF shared-memory buffers, one or two segments ahead of the compute.  Inside the segment row i of the
// tile reads partner i ^ (x & (2^K - 1)) from shared memory (a warp reads a permuted 512-byte line: conflict-free).
// L2 traffic falls from 16 B * G to 16 B * S per row (S = segments).  Values are evaluated as in apply_fold_kernel
// (thread <-> 8 rows differing in row bits 7..9, bucketed Walsh-Hadamard fold for groups of >= 3 terms).
// Row-sharded form: a segment whose h reaches above the shard reads its tile from the peer that owns it -- the same
// bulk copy, over NVLink (pa.peer[], as the gather kernel); p2p_ready_kernel / p2p_done_kernel bracket the launch.
// ---------------------------------------------------------------------------------
struct PtileDev {
    const uint32_t *seg_g0;    // [n_seg + 1] segment s = sorted groups [seg_g0[s], seg_g0[s + 1])
    const uint32_t *seg_h;     // [n_seg]     their common x >> K
    uint32_t n_seg;
};

template <int K, int NBUF>
__global__ void __launch_bounds__(1 << (K - 3), K >= 12 ? 1 : (K == 11 ? 2 : 4))
apply_ptile_kernel(PlanDev p, FoldDev f, PtileDev pt, uint64_t row_lo,
                   const double2 *__restrict__ v, double2 *__restrict__ y,
                   const double2 *__restrict__ diag, const double *__restrict__ diag_re,
                   const __grid_constant__ ApplyPeerArgs pa)
{
    constexpr int E = FOLD_ROWS, B0 = FOLD_B0;
    constexpr uint32_t TILE = 1u << K, TILE_BYTES = TILE * 16u, PIECE = 16384u;
    static_assert(K >= 10 && K <= 12 && TILE_BYTES % PIECE == 0, "tile = whole 16 KB pieces, row bits 7..9 inside it");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *ring = reinterpret_cast<double2 *>(smem_raw);                       // [NBUF][TILE]
    __shared__ uint64_t full[NBUF];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // row of the tile this thread's e = 0 row is: lane -> bits 0..4, warp -> bits 5, 6, 10.., e -> bits 7..9
    const uint32_t idx0 = lane | ((warp & 3u) << 5) | ((warp >> 2) << 10);
    const uint64_t tile_row0 = row_lo + ((uint64_t)blockIdx.x << K);            // host-checked: a multiple of 2^K
    const uint32_t tile_id = (uint32_t)(tile_row0 >> K);
    const uint32_t r0 = (uint32_t)tile_row0 + idx0;
    const bool PEERS = pa.n_peers > 1u;
    const uint32_t my_rank = PEERS ? (uint32_t)(row_lo >> pa.shard_bits) : 0u;
    const uint32_t S = pt.n_seg;

    auto issue = [&](uint32_t s, uint32_t b) {                                   // thread 0: partner tile of segment s -> buffer b
        const uint32_t h = __ldg(&pt.seg_h[s]);
        const double2 *base = PEERS ? pa.peer[my_rank ^ (h >> (pa.shard_bits - (uint32_t)K))] : v;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(base + ((uint64_t)(tile_id ^ h) << K));
        unsigned char *dst = reinterpret_cast<unsigned char *>(ring + (size_t)b * TILE);
        mbar_expect_tx(&full[b], TILE_BYTES);
#pragma unroll
        for (uint32_t c = 0; c < TILE_BYTES; c += PIECE) bulk_load_global_to_smem(dst + c, src + c, PIECE, &full[b]);
    };
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; b++) mbar_init(&full[b], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (uint32_t s = 0; s < (uint32_t)(NBUF - 1) && s < S; s++) issue(s, s);

    double yr[E], yi[E];
#pragma unroll
    for (int e = 0; e < E; e++) { yr[e] = 0.0; yi[e] = 0.0; }
    const bool have_diag = diag_re != nullptr || diag != nullptr;               // then group 0 is the mask-0 group
    uint32_t b = 0, phase = 0, b_next = NBUF - 1;                                // buffer / parity of segment s, buffer of segment s + NBUF - 1
    for (uint32_t s = 0; s < S; s++) {
        if (tid == 0 && s + (uint32_t)(NBUF - 1) < S) issue(s + (uint32_t)(NBUF - 1), b_next);   // its last readers passed the barrier below
        mbar_wait(&full[b], phase);
        const double2 *tb = ring + (size_t)b * TILE;
        const uint32_t g_end = __ldg(&pt.seg_g0[s + 1]);
        for (uint32_t g = __ldg(&pt.seg_g0[s]); g < g_end; g++) {
            const GroupDesc d = p.gdesc[g];
            const uint32_t pidx = idx0 ^ (d.x & (TILE - 1u));                    // partner of the e = 0 row; e flips bits 7..9
            const bool real = (d.flag & 2u) != 0u;
            const uint32_t n_t = d.t1 - d.t0;
            if (g == 0u && have_diag) {                                          // diag(H) from the cache
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const uint64_t lr = (uint64_t)(r0 + ((uint32_t)e << B0)) - row_lo;
                    if (diag_re != nullptr) cfma(yr[e], yi[e], __ldcs(&diag_re[lr]), 0.0, tb[pidx ^ ((uint32_t)e << B0)], true);
                    else { const double2 dg = __ldcs(&diag[lr]); cfma(yr[e], yi[e], dg.x, dg.y, tb[pidx ^ ((uint32_t)e << B0)], false); }
                }
                continue;
            }
            if (d.flag & 1u) {                                                   // row-independent value
                if (real) {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, 0.0, tb[pidx ^ ((uint32_t)e << B0)], true);
                } else {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, d.cim, tb[pidx ^ ((uint32_t)e << B0)], false);
                }
                continue;
            }
            double S8[E];
            if (n_t <= FOLD_WHT_MIN) {
                fold_short_sweep<false>(f.zc + d.t0, f.im + d.t0, n_t, r0, S8);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 w = tb[pidx ^ ((uint32_t)e << B0)];
                    yr[e] = __fma_rn(S8[e], w.x, yr[e]); yi[e] = __fma_rn(S8[e], w.y, yi[e]);
                }
                if (!real) {
                    fold_short_sweep<true>(f.zc + d.t0, f.im + d.t0, n_t, r0, S8);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double2 w = tb[pidx ^ ((uint32_t)e << B0)];
                        yr[e] = __fma_rn(-S8[e], w.y, yr[e]); yi[e] = __fma_rn(S8[e], w.x, yi[e]);
                    }
                }
                continue;
            }
            const uint4 be = __ldg(&f.bend[g]);
            const uint32_t ends[8] = {be.x & 0xffffu, be.x >> 16, be.y & 0xffffu, be.y >> 16,
                                      be.z & 0xffffu, be.z >> 16, be.w & 0xffffu, be.w >> 16};
            uint32_t t = d.t0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                double acc = 0.0;
                for (const uint32_t te = d.t0 + ends[q]; t < te; t++) {
                    const uint4 zc = __ldg(&f.zc[t]);
                    acc += signed_re(zc, (uint32_t)__popc(r0 & zc.x) << 31);
                }
                S8[q] = acc;
            }
            wht8(S8);
#pragma unroll
            for (int e = 0; e < E; e++) {
                const double2 w = tb[pidx ^ ((uint32_t)e << B0)];
                yr[e] = __fma_rn(S8[e], w.x, yr[e]); yi[e] = __fma_rn(S8[e], w.y, yi[e]);
            }
            if (!real) {
                t = d.t0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    double acc = 0.0;
                    for (const uint32_t te = d.t0 + ends[q]; t < te; t++) {
                        const uint32_t z = __ldg(&f.zc[t]).x;
                        acc += flip_sign(__ldg(&f.im[t]), (uint32_t)__popc(r0 & z) << 31);
                    }
                    S8[q] = acc;
                }
                wht8(S8);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 w = tb[pidx ^ ((uint32_t)e << B0)];
                    yr[e] = __fma_rn(-S8[e], w.y, yr[e]); yi[e] = __fma_rn(S8[e], w.x, yi[e]);
                }
            }
        }
        __syncthreads();                                                         // buffer b is free for segment s + NBUF
        b_next = b;
        if (++b == (uint32_t)NBUF) { b = 0; phase ^= 1u; }
    }
#pragma unroll
    for (int e = 0; e < E; e++) __stcs(&y[(uint64_t)(r0 + ((uint32_t)e << B0)) - row_lo], make_double2(yr[e], yi[e]));
}

}  // namespace qr
