// fill.cuh -- K3: the CSR fill kernels (indices + data + indptr).
//
// Replaces the per-row work of rowwise::make_row (qrusty/src/accel.rs:171-210)
// and the assembly of make_unsafe_vectors_chunked (accel.rs:267-336): no per-row
// sort, no per-row allocation, no concat passes -- every (row, group) value is
// computed in registers and lands at its final position.
//
// Thread mapping (both kernels): lane <-> row.  A warp owns 32 consecutive,
// 32-aligned rows and walks the groups, so everything that depends only on the
// group (mask, term list, rank-table row) is warp-uniform: no divergence in the
// term loop, broadcast loads of the term table.
//
//   col   = r ^ gx[g]                                            accel.rs:176
//   value = sum over the group's terms, in original term order,  accel.rs:177-184,
//           of (popc(r & z_t) odd ? -c'_t : c'_t)                :191-205
//           -- sign applied by flipping the IEEE sign bit, adds are __dadd_rn,
//           the first term is taken as is (no 0 + x), so data is bit-identical
//           to the reference's fold, signed zeros included.
//   slot  = sum_b cnt[g][b] * bit_b(gx[g] ^ r)                   (plan.cuh)
//           bits >= 5 are warp-uniform: lane b contributes bit b, one REDUX;
//           bits < 5 come from lr5[g][(gx[g]^lane)&31].
#pragma once
#include "plan.cuh"
#include "scan.cuh"

namespace qr {

__device__ __forceinline__ double flip_sign(double d, uint32_t sign_bit)
{
    return __hiloint2double(__double2hiint(d) ^ (int)sign_bit, __double2loint(d));
}

// Value of group [t0,t1) in row r.  t1 > t0 always (a group has >= 1 term).
__device__ __forceinline__ double2 group_value(const uint32_t *__restrict__ tz,
                                               const double2 *__restrict__ tc,
                                               uint32_t t0, uint32_t t1, uint32_t r)
{
    double2 c = __ldg(&tc[t0]);
    uint32_t s = (uint32_t)(__popc(r & __ldg(&tz[t0])) & 1) << 31;
    double re = flip_sign(c.x, s), im = flip_sign(c.y, s);
    for (uint32_t t = t0 + 1; t < t1; t++) {
        c = __ldg(&tc[t]);
        s = (uint32_t)(__popc(r & __ldg(&tz[t])) & 1) << 31;
        re = __dadd_rn(re, flip_sign(c.x, s));
        im = __dadd_rn(im, flip_sign(c.y, s));
    }
    return make_double2(re, im);
}

// Values of one group for E rows at once: the term table is walked once, every (z, c') is loaded
// once (warp-uniform) and applied to the E rows held in registers.  Same fold as group_value.
template <int E>
__device__ __forceinline__ void group_values(const PlanDev &p, uint32_t t0, uint32_t t1, const uint32_t (&r)[E],
                                             double (&ar)[E], double (&ai)[E])
{
    double2 c = __ldg(&p.tc[t0]);
    uint32_t z = __ldg(&p.tz[t0]);
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t s = (uint32_t)(__popc(r[e] & z) & 1) << 31;
        ar[e] = flip_sign(c.x, s); ai[e] = flip_sign(c.y, s);
    }
    for (uint32_t t = t0 + 1; t < t1; t++) {
        c = __ldg(&p.tc[t]);
        z = __ldg(&p.tz[t]);
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t s = (uint32_t)(__popc(r[e] & z) & 1) << 31;
            ar[e] = __dadd_rn(ar[e], flip_sign(c.x, s)); ai[e] = __dadd_rn(ai[e], flip_sign(c.y, s));
        }
    }
}

// Value of group g in row r: row-independent groups (gflag bit0: every z == 0, e.g. all
// X-only strings) skip the term loop -- gconst holds the same ordered fold, bit for bit.
__device__ __forceinline__ double2 group_value_g(const PlanDev &p, uint32_t g, uint32_t r)
{
    if (__ldg(&p.gflag[g]) & 1u) return __ldg(&p.gconst[g]);          // warp-uniform branch
    return group_value(p.tz, p.tc, __ldg(&p.goff[g]), __ldg(&p.goff[g + 1]), r);
}

// Slot of group g in each of the warp's 32 rows (all 32 lanes must call).
// warp_row_base: the warp's first row (multiple of 32).
__device__ __forceinline__ uint32_t group_slot(const PlanDev &p, uint32_t g, uint32_t x,
                                               uint32_t warp_row_base, uint32_t lane)
{
    const uint32_t c = __ldg(&p.cnt[g * 32u + lane]);
    const uint32_t bit = ((x ^ warp_row_base) >> lane) & 1u;
    const uint32_t hi = __reduce_add_sync(0xffffffffu, (lane >= 5u && bit) ? c : 0u);
    return hi + __ldg(&p.lr5[g * 32u + ((x ^ lane) & 31u)]);
}

// ---------------------------------------------------------------------------------
// Direct kernel: any row range, any G.  Stores go straight to global memory: a
// warp's 32 stores of one group are G*16 B apart, so this is the correctness
// baseline and the edge/large-G path, not the fast path.
// ---------------------------------------------------------------------------------
constexpr int FILL_DIRECT_THREADS = 256;

__global__ void __launch_bounds__(FILL_DIRECT_THREADS)
fill_direct_kernel(PlanDev p, uint32_t G, uint64_t lo, uint64_t hi, uint64_t out_row0, uint64_t req_hi,
                   uint64_t indptr_base, uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices,
                   double2 *__restrict__ data)
{
    // rows [lo,hi) of a request [out_row0, req_hi): outputs are indexed relative to out_row0
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_id = (uint64_t)blockIdx.x * (FILL_DIRECT_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t wbase64 = (lo & ~(uint64_t)31) + warp_id * 32u;
    if (wbase64 >= hi) return;                                   // warp-uniform
    const uint32_t wbase = (uint32_t)wbase64;
    const uint64_t r64 = wbase64 + lane;
    const uint32_t r = (uint32_t)r64;
    const bool live = r64 >= lo && r64 < hi;
    const uint64_t out_row = (r64 - out_row0) * G;               // garbage when !live, never used

    for (uint32_t g = 0; g < G; g++) {
        const uint32_t x = __ldg(&p.gx[g]);
        const uint32_t slot = group_slot(p, g, x, wbase, lane);
        const double2 v = group_value_g(p, g, r);
        if (live) {
            indices[out_row + slot] = (uint64_t)(r ^ x);
            data[out_row + slot] = v;
        }
    }
    if (indptr != nullptr && live) {
        indptr[r64 - out_row0] = indptr_base + (r64 - out_row0) * G;
        if (r64 + 1 == req_hi) indptr[req_hi - out_row0] = indptr_base + (req_hi - out_row0) * G;
    }
}

// ---- mbarrier helpers (also used by the tiled H.v kernels, apply_tile.cuh / apply_fold.cuh) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------------------------
// Staged kernel: a CTA owns a tile of R = 32*E whole rows (aligned to R).  The
// tile's R*G entries are contiguous in BOTH output arrays, so the CTA assembles
// them in shared memory in final order and one thread hands each array to the
// TMA as a single bulk copy (cp.async.bulk.global.shared::cta): HBM sees only
// full-line, perfectly sequential writes and no LSU instruction is spent on them.
// The GW warps split the groups (g = gw, gw+GW, ...); each handles E row strips per visit.
// Requires R*G*24 B of shared memory; the host picks (E, GW) from G.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store_smem_to_global(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

// PAD: the 32 lanes of a store hold one row each, G*16 bytes apart.  When G is a multiple of 8 they all
// fall on the same shared-memory banks (8-way conflicts: 1.9x slower on XXZ n=23, l1tex 93 % busy with
// 30 M conflict cycles).  The padded layout inserts one 16-byte gap after every 2^ps rows, which spreads
// the rows over the banks, and the tile becomes R >> ps contiguous pieces per array instead of one; the
// lanes of warp 0 hand them to the TMA.  Measured (profiles/r02_pad_probe.jsonl): ps = 2 is the best
// period for G = 24 (4.65 vs 3.51 TB/s unpadded; one piece per row is bound by the TMA's request rate);
// for every other G the single copy of the unpadded tile wins, 4-way-conflict cases (G = 20, 28) included.
template <int E, int GW, bool PAD>
__global__ void __launch_bounds__(32 * GW)
fill_staged_kernel(PlanDev p, uint32_t G, uint64_t tile_row0, uint64_t row_lo, uint64_t indptr_base,
                   uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices,
                   double2 *__restrict__ data, uint64_t indptr_last_row, uint32_t ps)
{
    // tile = E strips of 32 rows; each of the GW warps visits its share of the groups and handles
    // all E strips per visit: descriptor, rank-table row and every term are read once per 32*E entries
    constexpr uint32_t R = 32u * E;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t gaps = PAD ? (R >> ps) : 0u;
    double2 *sdat = reinterpret_cast<double2 *>(smem_raw);                                         // (R*G + gaps) * 16 B
    uint64_t *sidx = reinterpret_cast<uint64_t *>(smem_raw + ((size_t)R * G + gaps) * 16u);         // (R*G + 2*gaps) * 8 B

    const uint32_t lane = threadIdx.x & 31u, gw = threadIdx.x >> 5;
    const uint64_t tile_base = tile_row0 + (uint64_t)blockIdx.x * R;
    const uint32_t tbase = (uint32_t)tile_base;
    // launched ahead of the end of the canonicalisation (programmatic dependent launch): the plan tables are not to be
    // read, nor the outputs written, before it has completed
    pdl_wait();
    pdl_launch_dependents();
    uint32_t r[E], od[E], oi[E];                                  // row, and where the row starts in sdat / sidx
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t row = 32u * e + lane;
        r[e] = tbase + row;
        od[e] = row * G + (PAD ? (row >> ps) : 0u);
        oi[e] = row * G + (PAD ? ((row >> ps) << 1) : 0u);
    }

    for (uint32_t g = gw; g < G; g += GW) {
        const GroupDesc d = p.gdesc[g];                                          // warp-uniform, 2 x 16 B
        const uint32_t c = __ldg(&p.cnt[g * 32u + lane]);
        const uint32_t lo = __ldg(&p.lr5[g * 32u + ((d.x ^ lane) & 31u)]);
        double ar[E], ai[E];
        if (d.flag & 1u) {
#pragma unroll
            for (int e = 0; e < E; e++) { ar[e] = d.cre; ai[e] = d.cim; }
        } else {
            group_values<E>(p, d.t0, d.t1, r, ar, ai);
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t bit = ((d.x ^ (tbase + 32u * e)) >> lane) & 1u;
            const uint32_t slot = __reduce_add_sync(0xffffffffu, (lane >= 5u && bit) ? c : 0u) + lo;
            sidx[oi[e] + slot] = (uint64_t)(r[e] ^ d.x);
            sdat[od[e] + slot] = make_double2(ar[e], ai[e]);
        }
    }
    if (gw == 0 && indptr != nullptr) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint64_t lr = tile_base + 32u * e + lane - row_lo;
            indptr[lr] = indptr_base + lr * G;
            if (lr + 1 == indptr_last_row) indptr[lr + 1] = indptr_base + (lr + 1) * G;
        }
    }
    // generic-proxy writes -> visible to the async proxy, then the copies are issued
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint64_t off = (tile_base - row_lo) * G;
    if (PAD) {
        if (gw == 0) {
            const uint32_t piece = G << ps;                                       // entries per contiguous piece
            for (uint32_t c = lane; c < gaps; c += 32u) {
                bulk_store_smem_to_global(data + off + (uint64_t)c * piece, sdat + c * (piece + 1u), piece * 16u);
                bulk_store_smem_to_global(indices + off + (uint64_t)c * piece, sidx + c * (piece + 2u), piece * 8u);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (threadIdx.x == 0) {
        bulk_store_smem_to_global(data + off, sdat, R * G * 16u);
        bulk_store_smem_to_global(indices + off, sidx, R * G * 8u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
    }
}

// ---------------------------------------------------------------------------------
// Staged kernel, swizzled tile (even G).  The lanes of a store hold one row each, G*16 bytes apart: with G even they
// share shared-memory banks (2-way; 4-way for G = 4 mod 8: XXZ n = 19 at 0.72 of peak; 8-way for G = 0 mod 8).  Here the
// tile is not the final byte order: it is cut into column boxes of 128 bytes per row (8 entries of data, 16 column ids)
// laid out as the TMA's 128-byte swizzle wants them -- row i of a box at i * 128, its 16-byte chunk c at c ^ (i & 7) -- so
// the eight lanes of a store phase always fall on eight different chunks, and the boxes leave through 2-D tensor maps of
// the output arrays (cp.async.bulk.tensor.2d.global.shared::cta, SASS UTMASTG), which undo the swizzle on the way out.
// The last box of a row may reach past column G: the TMA clips it.  Values, order and bytes written are the staged
// kernel's.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tensor_store_2d(const void *tmap, int32_t x, int32_t y, const void *ssrc)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(tmap), "r"(x), "r"(y), "r"((uint32_t)__cvta_generic_to_shared(ssrc)) : "memory");
}

struct __align__(64) TensorMap { unsigned char bytes[128]; };      // a CUtensorMap, opaque to the device code

template <int E, int GW>
__global__ void __launch_bounds__(32 * GW)
fill_staged_swz_kernel(PlanDev p, uint32_t G, uint64_t tile_row0, uint64_t row_lo, uint64_t indptr_base,
                       uint64_t *__restrict__ indptr, uint64_t indptr_last_row,
                       const __grid_constant__ TensorMap tm_data, const __grid_constant__ TensorMap tm_idx)
{
    constexpr uint32_t R = 32u * E, BOX = R * 128u;                // bytes of one box: R rows of 128 bytes
    extern __shared__ unsigned char smem_dyn[];
    // the swizzle pattern is a function of the shared-memory ADDRESS: boxes start on 1024-byte boundaries
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem_dyn);
    unsigned char *tile = smem_dyn + ((1024u - (s0 & 1023u)) & 1023u);
    const uint32_t nbd = (G + 7u) >> 3, nbi = (G + 15u) >> 4;      // data / column-id boxes per row
    unsigned char *sdat = tile, *sidx = tile + (size_t)nbd * BOX;

    const uint32_t lane = threadIdx.x & 31u, gw = threadIdx.x >> 5;
    const uint64_t tile_base = tile_row0 + (uint64_t)blockIdx.x * R;
    const uint32_t tbase = (uint32_t)tile_base;
    pdl_wait();
    pdl_launch_dependents();
    uint32_t r[E], ob[E];                                          // row, byte offset of the row inside a box
#pragma unroll
    for (int e = 0; e < E; e++) { r[e] = tbase + 32u * e + lane; ob[e] = (32u * e + lane) * 128u; }
    const uint32_t sw = lane & 7u;                                 // (row & 7): 32 * e does not touch it

    for (uint32_t g = gw; g < G; g += GW) {
        const GroupDesc d = p.gdesc[g];
        const uint32_t c = __ldg(&p.cnt[g * 32u + lane]);
        const uint32_t lo = __ldg(&p.lr5[g * 32u + ((d.x ^ lane) & 31u)]);
        double ar[E], ai[E];
        if (d.flag & 1u) {
#pragma unroll
            for (int e = 0; e < E; e++) { ar[e] = d.cre; ai[e] = d.cim; }
        } else {
            group_values<E>(p, d.t0, d.t1, r, ar, ai);
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t bit = ((d.x ^ (tbase + 32u * e)) >> lane) & 1u;
            const uint32_t slot = __reduce_add_sync(0xffffffffu, (lane >= 5u && bit) ? c : 0u) + lo;
            const uint32_t doff = (slot >> 3) * BOX + ob[e] + (((slot & 7u) ^ sw) << 4);
            const uint32_t ioff = (slot >> 4) * BOX + ob[e] + (((((slot & 15u) >> 1) ^ sw) << 4) | ((slot & 1u) << 3));
            *reinterpret_cast<double2 *>(sdat + doff) = make_double2(ar[e], ai[e]);
            *reinterpret_cast<uint64_t *>(sidx + ioff) = (uint64_t)(r[e] ^ d.x);
        }
    }
    if (gw == 0 && indptr != nullptr) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint64_t lr = tile_base + 32u * e + lane - row_lo;
            indptr[lr] = indptr_base + lr * G;
            if (lr + 1 == indptr_last_row) indptr[lr + 1] = indptr_base + (lr + 1) * G;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (gw == 0) {
        const int32_t y = (int32_t)(tile_base - row_lo);            // row coordinate in the output arrays (host-checked: < 2^31)
        for (uint32_t b = lane; b < nbd + nbi; b += 32u) {
            if (b < nbd) tensor_store_2d(&tm_data, (int32_t)(b * 16u), y, sdat + (size_t)b * BOX);       // 16 doubles = 8 entries
            else tensor_store_2d(&tm_idx, (int32_t)((b - nbd) * 16u), y, sidx + (size_t)(b - nbd) * BOX); // 16 column ids
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------
// Blocked kernel (large G).  Work item = (block of <= S groups, run of `strips_per_cta`
// 32-row strips).  The block's tables (cnt, lr5, masks, term offsets) are copied to
// shared memory once and reused for every strip.  For a strip:
//   base   = sum_{b >= p} cnt[g0][b] * bit_b(gx[g0] ^ rbase)      slots before the block
//   offset = sum_{5 <= b < p} cnt[g][b] * bit_b(gx[g] ^ rbase)    (REDUX over lanes [5,p))
//            + lr5[g][(gx[g] ^ lane) & 31]                        position inside the block
// Each lane drops (col, value) at stage[lane][offset]; after a barrier the 32 row
// segments [base, base + size) stream out with coalesced 16-byte / 8-byte stores.
// ---------------------------------------------------------------------------------
constexpr int FILL_BLOCKED_WARPS = 8;
constexpr int FILL_BLOCKED_TCAP = 512;      // terms of a block staged in shared memory (10 KB)

// group_values with the term table behind generic pointers (shared or global memory)
template <int E>
__device__ __forceinline__ void group_values_ptr(const uint32_t *tz, const double2 *tc, uint32_t t0, uint32_t t1,
                                                 const uint32_t (&r)[E], double (&ar)[E], double (&ai)[E])
{
    double2 c = tc[t0];
    uint32_t z = tz[t0];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t s = (uint32_t)(__popc(r[e] & z) & 1) << 31;
        ar[e] = flip_sign(c.x, s); ai[e] = flip_sign(c.y, s);
    }
    for (uint32_t t = t0 + 1; t < t1; t++) {
        c = tc[t];
        z = tz[t];
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t s = (uint32_t)(__popc(r[e] & z) & 1) << 31;
            ar[e] = __dadd_rn(ar[e], flip_sign(c.x, s)); ai[e] = __dadd_rn(ai[e], flip_sign(c.y, s));
        }
    }
}

template <int E>
__global__ void __launch_bounds__(32 * FILL_BLOCKED_WARPS)
fill_blocked_kernel(PlanDev p, uint32_t G, uint32_t S, uint32_t n_blocks, uint32_t strips_per_cta,
                    uint64_t tile_row0, uint64_t n_strips, uint64_t row_lo, uint64_t indptr_base,
                    uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices,
                    double2 *__restrict__ data, uint64_t indptr_last_row)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t pitch = S + 1;
    constexpr uint32_t ROWS = 32u * E;
    double2 *sdat = reinterpret_cast<double2 *>(smem_raw);                                // [32E][S+1]
    uint64_t *sidx = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ROWS * pitch * 16);    // [32E][S+1]
    unsigned char *tab = smem_raw + (size_t)ROWS * pitch * 24;
    GroupDesc *s_desc = reinterpret_cast<GroupDesc *>(tab);                               // [S]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(tab + (size_t)S * sizeof(GroupDesc));   // [S][32]
    uint32_t *s_lr5 = s_cnt + S * 32u;                                                    // [S][32]
    double2 *s_tc = reinterpret_cast<double2 *>(s_lr5 + S * 32u);                         // [TCAP]
    uint32_t *s_tz = reinterpret_cast<uint32_t *>(s_tc + FILL_BLOCKED_TCAP);               // [TCAP]

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t blk = blockIdx.x % n_blocks;
    const uint64_t strip0 = (uint64_t)(blockIdx.x / n_blocks) * strips_per_cta;
    const uint32_t g0 = __ldg(&p.blk_start[blk]), g1 = __ldg(&p.blk_start[blk + 1]);
    const uint32_t size = g1 - g0, plevel = __ldg(&p.blk_p[blk]);
    const uint32_t term0 = __ldg(&p.goff[g0]), n_terms = __ldg(&p.goff[g1]) - term0;
    const bool terms_in_smem = n_terms <= (uint32_t)FILL_BLOCKED_TCAP;

    for (uint32_t i = threadIdx.x; i < size * 32u; i += blockDim.x) {
        s_cnt[i] = __ldg(&p.cnt[g0 * 32u + i]);
        s_lr5[i] = __ldg(&p.lr5[g0 * 32u + i]);
    }
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) s_desc[i] = p.gdesc[g0 + i];
    if (terms_in_smem)
        for (uint32_t i = threadIdx.x; i < n_terms; i += blockDim.x) { s_tc[i] = __ldg(&p.tc[term0 + i]); s_tz[i] = __ldg(&p.tz[term0 + i]); }
    __syncthreads();
    // term table of this block: shared copy (indices relative to term0) or the global one
    const uint32_t *tzp = terms_in_smem ? s_tz - term0 : p.tz;
    const double2 *tcp = terms_in_smem ? s_tc - term0 : p.tc;

    const uint32_t in_block = (lane >= 5u && lane < plevel) ? 1u : 0u;     // bits ordering groups inside the block
    const uint32_t above = lane >= plevel ? 1u : 0u;                        // bits ordering the block among blocks
    const uint64_t strip_end = min(strip0 + strips_per_cta, n_strips);
    for (uint64_t strip = strip0; strip < strip_end; strip += E) {
        const uint64_t tile_base = tile_row0 + strip * 32u;
        const uint32_t tbase = (uint32_t)tile_base;
        const uint32_t ne = (uint32_t)min((uint64_t)E, strip_end - strip);   // strips in this visit (tail: 1)
        uint32_t r[E], base[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
            r[e] = tbase + 32u * e + lane;
            const uint32_t bit0 = ((s_desc[0].x ^ (tbase + 32u * e)) >> lane) & 1u;
            base[e] = __reduce_add_sync(0xffffffffu, (above && bit0) ? s_cnt[lane] : 0u);
        }
        for (uint32_t gi = warp; gi < size; gi += FILL_BLOCKED_WARPS) {
            const GroupDesc d = s_desc[gi];
            const uint32_t c = s_cnt[gi * 32u + lane];
            const uint32_t lo = s_lr5[gi * 32u + ((d.x ^ lane) & 31u)];
            double ar[E], ai[E];
            if (d.flag & 1u) {
#pragma unroll
                for (int e = 0; e < E; e++) { ar[e] = d.cre; ai[e] = d.cim; }
            } else {
                group_values_ptr<E>(tzp, tcp, d.t0, d.t1, r, ar, ai);
            }
#pragma unroll
            for (int e = 0; e < E; e++) {
                const uint32_t bit = ((d.x ^ (tbase + 32u * e)) >> lane) & 1u;
                const uint32_t off = __reduce_add_sync(0xffffffffu, (in_block && bit) ? c : 0u) + lo;
                const uint32_t o = (32u * e + lane) * pitch + off;
                sidx[o] = (uint64_t)(r[e] ^ d.x);
                sdat[o] = make_double2(ar[e], ai[e]);
            }
        }
        if (blk == 0 && warp == 0 && indptr != nullptr) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                if ((uint32_t)e < ne) {
                    const uint64_t lr = tile_base + 32u * e + lane - row_lo;
                    indptr[lr] = indptr_base + lr * G;
                    if (lr + 1 == indptr_last_row) indptr[lr + 1] = indptr_base + (lr + 1) * G;
                }
            }
        }
        __syncthreads();
        for (uint32_t l = warp; l < 32u * ne; l += FILL_BLOCKED_WARPS) {
            const uint64_t o = (tile_base + l - row_lo) * G + base[0];
            // base differs per strip: rows of strip e use base[e]
            const uint64_t oo = o - base[0] + (l < 32u ? base[0] : base[E - 1]);
            for (uint32_t k = lane; k < size; k += 32u) {
                data[oo + k] = sdat[l * pitch + k];
                indices[oo + k] = sidx[l * pitch + k];
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------
// Lanes kernel (large G, the default when whole rows do not fit in shared memory).
//
// lane <-> group: a warp owns 32 consecutive sorted groups and walks a run of R = 2^k
// aligned rows in GRAY-CODE order.  Everything that depends only on the group -- mask,
// up to NT (z, c') terms, the rank-table column -- lives in the lane's registers for the
// whole run, and moving to the next row flips exactly one row bit b, so
//     offset(r ^ 2^b, g) = offset(r, g) +- (cnt[g][b] +- G*2^b)     (one IADD; the sign alternates)
// and a row costs ~13 + 6*(terms-1) thread instructions per entry with no shared-memory
// staging.  The 32 entries a warp stores per row are the block's slots of that row:
// consecutive sorted masks are a union of a few trie subtrees, each of which fills one
// contiguous slot range in every row (XOR never splits a subtree), so one STG.128 + one
// STG.64 per row cover a few contiguous segments (<= 512 B + 256 B).  All warps of a CTA --
// and the sibling CTAs that own the other groups of the same rows, adjacent in blockIdx --
// walk the same row sequence and re-align every 32 rows (`resync`), so the sectors shared by
// two segments are completed in L2 within the warps' drift, long before they are evicted
// (ncu: without the re-alignment 0.5 GB of DRAM read-modify-write traffic per 4.7 GB written).
//
// Groups with more than NT terms ("heavy": the Z-only group of a molecular Hamiltonian, a
// few dozen others) are collected per CTA and handled lane <-> row at the start of every
// 32-row strip, round-robin over the CTA's warps, with a warp-uniform term loop -- in the
// same time window as their neighbours in the row.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_global_f64x2(uintptr_t addr, double a, double b)
{
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" :: "l"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_global_u64(uintptr_t addr, uint64_t v)
{
    asm volatile("st.global.u64 [%0], %1;" :: "l"(addr), "l"(v) : "memory");
}

constexpr int FILL_LANES_NT = LANE_TERMS;
constexpr int FILL_LANES_MAXLOG2R = 12;

// CTA-wide barrier that does not care which code path a warp arrives from (the warps of a CTA run
// differently specialised row walks): bar.sync on a named barrier with an explicit thread count.
__device__ __forceinline__ void lanes_barrier(uint32_t n_threads)
{
    asm volatile("bar.sync 1, %0;" :: "r"(n_threads) : "memory");
}
// The same across the thread-block cluster that holds the sibling CTAs of a row run (pure
// rendezvous: no data is exchanged, so the relaxed form is enough).
__device__ __forceinline__ void lanes_cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

struct LanesCtx {
    const PlanDev *p;
    uintptr_t dptr, iptr;          // outputs, pre-offset to the run's first row
    const uint32_t *s_heavy;       // heavy groups of this CTA
    uint32_t n_heavy, G, r0, R, resync, n_threads, lane, warp, n_warps;   // resync: 0 none, 1 CTA, 2 cluster
};

// heavy groups of the CTA for the 32-row strip that holds r: lane <-> row, round-robin over warps
__device__ __forceinline__ void lanes_heavy_strip(const LanesCtx &c, uint32_t r)
{
    const PlanDev &p = *c.p;
    const uint32_t wbase = r & ~31u, rr = wbase + c.lane;
    for (uint32_t q = c.warp; q < c.n_heavy; q += c.n_warps) {
        const uint32_t hg = c.s_heavy[q];
        const GroupDesc d = p.gdesc[hg];
        const uint32_t slot = group_slot(p, hg, d.x, wbase, c.lane);
        const double2 v = group_value(p.tz, p.tc, d.t0, d.t1, rr);
        const uint32_t o = (rr - c.r0) * c.G + slot;
        st_global_f64x2(c.dptr + ((size_t)o << 4), v.x, v.y);
        st_global_u64(c.iptr + ((size_t)o << 3), (uint64_t)(rr ^ d.x));
    }
}

// The Gray-code row walk of one warp, specialised for the number K of terms its longest light
// group holds (a warp-uniform switch picks the instance, so no slot is spent on absent terms).
// K == 0: the warp owns no groups and only takes part in the heavy strips and the barriers.
template <int K, int NT>
__device__ __forceinline__ void lanes_walk(const LanesCtx &c, bool active, uint32_t x, uint32_t off,
                                           int32_t d0, int32_t d1, int32_t d2, int32_t *s_delta /* [bit-3][32] + lane */,
                                           const uint32_t (&z)[NT], const double (&cr)[NT], const double (&ci)[NT])
{
    uint32_t r = c.r0;                                             // warp-uniform current row
    // +-1.0 from the parity of r & z: (+-1.0) * c' and fma(+-1.0, c', acc) are exactly the sign flip
    // and the __dadd_rn of the reference fold (accel.rs:191-205), signed zeros included
    auto sign_of = [](uint32_t m) { return __hiloint2double((int)(0x3ff00000u | ((uint32_t)__popc(m) << 31)), 0); };
    auto emit = [&]() {
        if (K == 0) return;
        const double s0 = sign_of(r & z[0]);
        double re = __dmul_rn(s0, cr[0]), im = __dmul_rn(s0, ci[0]);
#pragma unroll
        for (int t = 1; t < K; t++) {
            const double sg = sign_of(r & z[t]);
            re = __fma_rn(sg, cr[t], re); im = __fma_rn(sg, ci[t], im);
        }
        if (active) {
            st_global_f64x2(c.dptr + ((size_t)off << 4), re, im);
            st_global_u64(c.iptr + ((size_t)off << 3), (uint64_t)(r ^ x));
        }
    };
    for (uint32_t i = 0; i < c.R; i += 8u) {
        if (i != 0u) {                                             // Gray code: step i flips bit ctz(i) >= 3
            if ((i & 31u) == 0u) {                                 // keep every writer of these rows on the same strip
                if (c.resync == 2u) lanes_cluster_barrier(); else if (c.resync == 1u) lanes_barrier(c.n_threads);
            }
            const uint32_t b = (uint32_t)__ffs((int)i) - 1u;
            r ^= 1u << b;
            if (K != 0) {
                const int32_t st = s_delta[(b - 3u) * 32u];
                off += (uint32_t)st;
                s_delta[(b - 3u) * 32u] = -st;
            }
        }
        if ((i & 31u) == 0u && c.n_heavy != 0u) lanes_heavy_strip(c, r);
        emit();
        r ^= 1u; off += (uint32_t)d0; d0 = -d0; emit();
        r ^= 2u; off += (uint32_t)d1; d1 = -d1; emit();
        r ^= 1u; off += (uint32_t)d0; d0 = -d0; emit();
        r ^= 4u; off += (uint32_t)d2; d2 = -d2; emit();
        r ^= 1u; off += (uint32_t)d0; d0 = -d0; emit();
        r ^= 2u; off += (uint32_t)d1; d1 = -d1; emit();
        r ^= 1u; off += (uint32_t)d0; d0 = -d0; emit();
    }
}

// All runs of one persistent CTA for a fixed K (only the first K terms stay live in registers).
template <int K, int NT, int LW>
__device__ __forceinline__ void lanes_runs(LanesCtx &c, uint32_t j0, uint32_t J, uint32_t n_runs, uint32_t log2R, uint32_t beta,
                                           bool warp_live, bool active, uint32_t x, const uint32_t *s_cnt, int32_t *sd,
                                           uint64_t tile_row0, uint64_t row_lo, uint64_t indptr_base, uint64_t *indptr,
                                           uint64_t *indices, double2 *data, uint64_t indptr_last_row,
                                           const uint32_t (&z)[NT], const double (&cr)[NT], const double (&ci)[NT])
{
    const uint32_t nq = (uint32_t)c.p->n_qubits, G = c.G, R = c.R;
    for (uint32_t run = j0; run < n_runs; run += J) {
        const uint64_t r0_64 = tile_row0 + (uint64_t)run * R;      // first row of the run (aligned to R)
        const uint32_t r0 = (uint32_t)r0_64;
        if (beta == 0 && indptr != nullptr) {
            for (uint32_t i = threadIdx.x; i < R; i += 32 * LW) {
                const uint64_t lr = r0_64 + i - row_lo;
                indptr[lr] = indptr_base + lr * G;
                if (lr + 1 == indptr_last_row) indptr[lr + 1] = indptr_base + (lr + 1) * G;
            }
        }
        // entry offset (in entries, relative to the run's first row) at the first row of the run, and
        // the signed step for every row bit of the run.  Flipping row bit b moves the row by +-2^b (the
        // offset by +-G*2^b) and the slot by +-cnt[g][b]; both signs alternate with every flip of b, so
        // one signed step per (lane, bit) carries both.  R*G*16 < 2^32 (host-checked): 32-bit byte offsets.
        uint32_t off = 0;
        int32_t d0 = 0, d1 = 0, d2 = 0;
        if (warp_live) {
            const uint32_t xr = x ^ r0;
            for (uint32_t b = 0; b < nq; b++) {
                const uint32_t cb = s_cnt[b * 32u];
                const bool one = (xr >> b) & 1u;
                if (one) off += cb;
                if (b < log2R) {
                    // r0 is aligned to R: its bits < log2R are 0, the first flip of b moves the row up
                    const int32_t step = (one ? -(int32_t)cb : (int32_t)cb) + (int32_t)(G << b);
                    if (b == 0) d0 = step; else if (b == 1) d1 = step; else if (b == 2) d2 = step;
                    else sd[(b - 3u) * 32u] = step;
                }
            }
        }
        c.dptr = reinterpret_cast<uintptr_t>(data + (r0_64 - row_lo) * G);
        c.iptr = reinterpret_cast<uintptr_t>(indices + (r0_64 - row_lo) * G);
        c.r0 = r0;
        // the CTA's warps -- with a cluster launch: all CTAs that write these rows -- start the run together
        if (c.resync == 2u) lanes_cluster_barrier(); else if (c.resync == 1u) lanes_barrier(c.n_threads);
        lanes_walk<K, NT>(c, active, x, off, d0, d1, d2, sd, z, cr, ci);
    }
}

// Persistent CTAs: CTA (beta, j) owns the 32*LW groups of batch beta for good -- masks and terms
// stay in registers, the groups' rank-table columns in shared memory -- and takes the runs
// j, j + J, j + 2J, ... of the window; sibling CTAs (same j, other batches) take the same runs in
// the same order.  Per run only the entry offsets are re-derived (from shared memory, no global
// load), so short runs (R = 32: all of a row's segments are written within a few microseconds of
// each other) cost no memory latency.
template <int NT, int LW>
__global__ void __launch_bounds__(32 * LW, 32 / LW)
fill_lanes_kernel(PlanDev p, uint32_t G, uint32_t n_light, uint32_t log2R, uint32_t resync, uint32_t n_runs,
                  uint64_t tile_row0, uint64_t row_lo, uint64_t indptr_base,
                  uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices,
                  double2 *__restrict__ data, uint64_t indptr_last_row)
{
    extern __shared__ uint32_t s_cnt_all[];                        // [LW][n_qubits][32]: cnt[g][b] of the warp's groups
    __shared__ int32_t s_delta[LW][FILL_LANES_MAXLOG2R - 3][32];   // signed step of the row bits b >= 3
    __shared__ uint32_t s_heavy[LW * 32];                          // heavy groups of this CTA's blocks
    __shared__ uint32_t s_nheavy;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t beta = blockIdx.x % n_light, j0 = blockIdx.x / n_light, J = gridDim.x / n_light;
    const uint32_t R = 1u << log2R, nq = (uint32_t)p.n_qubits;

    if (threadIdx.x == 0) s_nheavy = 0;
    __syncthreads();

    // ---- once per CTA: warp <-> block of 32 groups, lane <-> group -----------------------------------
    const uint32_t g = (beta * LW + warp) * 32u + lane;
    const bool warp_live = (beta * LW + warp) * 32u < G;          // warp-uniform
    // one round of coalesced, mutually independent loads: (mask, term count), the first NT terms (padded
    // with (z = 0, c' = -0.0): x + (-0.0) == x for every x, signed zeros included, so padding never needs a
    // predicate), and -- below -- the rank-table column
    const uint32_t T = p.n_terms, gg = g < G ? g : G - 1u;
    const uint2 xn = __ldg(&p.lt_xn[gg]);
    uint32_t z[NT];
    double cr[NT], ci[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) { z[t] = __ldg(&p.lt_z[t * T + gg]); const double2 c = __ldg(&p.lt_c[t * T + gg]); cr[t] = c.x; ci[t] = c.y; }
    const uint32_t x = xn.x;
    uint32_t nt = xn.y;
    const bool active = g < G && nt <= (uint32_t)NT;
    if (g < G && !active) s_heavy[atomicAdd(&s_nheavy, 1u)] = g;
    if (!active) nt = 0;
    const uint32_t nt_max = __reduce_max_sync(0xffffffffu, nt);
    uint32_t *s_cnt = s_cnt_all + (size_t)warp * nq * 32u + lane;   // this lane's column: s_cnt[b * 32]
    if (warp_live)
        for (uint32_t b = 0; b < nq; b++) s_cnt[b * 32u] = __ldg(&p.cnt_t[b * T + gg]);        // coalesced
    __syncthreads();                                               // s_heavy complete

    LanesCtx c;
    c.p = &p;
    c.s_heavy = s_heavy; c.n_heavy = s_nheavy; c.G = G; c.R = R; c.resync = resync;
    c.n_threads = 32u * LW; c.lane = lane; c.warp = warp; c.n_warps = LW;
    int32_t *sd = &s_delta[warp][0][lane];
    const uint32_t k_sel = warp_live ? (nt_max == 0u ? 1u : nt_max) : 0u;   // all-heavy live warp: K = 1, stores off

    static_assert(NT == 6, "the switch below enumerates 0..NT terms");
#define QR_LANES_RUNS(K_, ACT_)                                                                                         \
    lanes_runs<K_, NT, LW>(c, j0, J, n_runs, log2R, beta, warp_live, ACT_, x, s_cnt, sd, tile_row0, row_lo, indptr_base, \
                           indptr, indices, data, indptr_last_row, z, cr, ci)
    switch (k_sel) {                                               // warp-uniform, once per CTA
    case 0: QR_LANES_RUNS(0, false); break;
    case 1: QR_LANES_RUNS(1, active); break;
    case 2: QR_LANES_RUNS(2, active); break;
    case 3: QR_LANES_RUNS(3, active); break;
    case 4: QR_LANES_RUNS(4, active); break;
    case 5: QR_LANES_RUNS(5, active); break;
    default: QR_LANES_RUNS(6, active); break;
    }
#undef QR_LANES_RUNS
}

// ---------------------------------------------------------------------------------
// Rows kernel: the default for every operator the staged kernel does not take (host: choose_rows, qrusty_cuda.cu).
//
// A persistent CTA owns WHOLE rows (split mode: whole row segments), so -- like the staged kernel, and
// unlike the lanes kernel -- everything it sends to HBM is a sequential stream of full lines issued by the
// TMA; no LSU store instruction touches global memory and no sector is ever half-written.
//
// thread <-> group: a thread keeps its groups (<= NG of them, TH apart) in registers for the whole kernel:
// mask, slot steps of the lowest row bits, and either the first (z, c') -- the other terms ("extras", T - G of
// them) sit in a shared-memory table -- or, REGT, up to six (z, c') per group (template parameters below).
// Rows are visited in batches of RT = 2^Q aligned consecutive rows per thread; the batches of a run of
// R = 2^log2R rows follow the Gray code of the batch index, so from one batch to the next exactly one row bit
// b flips and the slot of group g moves by +-cnt[g][b] (plan.cuh):
//     off(g) += (bit_b(row) just became 1) ? sd : -sd,     sd = bit_b(x_g) ? -cnt[g][b] : +cnt[g][b]
// (sd of the two lowest batch bits is a register, the rest one coalesced L1- or shared-memory-resident load
// per 4 batches).  Within a batch the RT rows differ in bits < Q only: slot(row j) = off + sum_{b in j} sd_b,
// and popc((r + j) & z) = popc(r & z) + popc(j & z): one POPC per term serves the RT rows.
// A batch is assembled in one of two shared-memory buffers in final order -- the lanes of a warp hold
// consecutive sorted groups, whose slots form a few contiguous runs (XOR never splits a trie subtree), so the
// 16-byte shared stores are mostly conflict-free -- and is handed to the TMA as two bulk copies (data, column
// ids) while the CTA fills the other buffer.  One barrier per batch.
//
// Small G (one group per thread, G <= TH / 2): the CTA's threads split into 2^sl sub-batches of TH >> sl
// threads, sub-batch s taking rows s*RT .. s*RT + RT - 1 of a batch of RT << sl rows, so that all threads
// work and a batch stays tens of KB however short the rows are.
//
// Groups with more than hv_thr terms ("heavy": the Z-only group of a molecular Hamiltonian and a
// few dozen others) would make their owner's warp the critical path of every batch.  They are
// evaluated by the whole CTA instead, lane <-> row, for an aligned strip of 2^hv_log2 = 32..128
// rows at a time (the next batches stay inside that strip), into a side buffer s_hv[heavy][strip]
// their owners read back; a lane folds up to 4 rows at once, which is what hides the latency of
// the one dependent FP64 chain per (group, row).
//
// Values: first term = sign-bit flip; later terms fma(+-1.0, c', acc) in original term order -- the product
// is exact, so this is the __dadd_rn fold of accel.rs:191-205 bit for bit, signed zeros included.
//
// Dynamic shared memory (host: rows_smem): [2][tile] double2 data | [2][tile] u64 column ids | term table
// (c' then z) | s_hv [hv_cap][2^hv_log2] double2 | s_h0c [hv_cap] double2 | s_hd [hv_cap] uint4 | s_ord [hv_cap] | s_cnt (CS).
// ---------------------------------------------------------------------------------
// ---- thread-block cluster helpers (rows kernel, CL > 1) ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier()           // arrive has release, wait has acquire semantics
{
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)    // same offset in CTA `rank`'s shared memory
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f64x2(uint32_t addr, double a, double b)
{
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, uint64_t v)
{
    asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}

// +-1.0 whose sign is bit 0 of p (the other bits of p are ignored)
__device__ __forceinline__ double pm_one(uint32_t p)
{
    return __hiloint2double((int)(0x3ff00000u + (p << 31)), 0);
}
// bit 0 of the result: parity of popc((rb + j) & z) given p = popc(rb & z), for rb with no bits below Q and j < 2^Q
template <int Q>
__device__ __forceinline__ uint32_t rows_parity(uint32_t p, uint32_t z, uint32_t j)
{
#pragma unroll
    for (int b = 0; b < Q; b++) if ((j >> b) & 1u) p ^= z >> b;
    return p;
}

// Folds terms [0, n) of (zs, cs) into HE rows per lane: r, r + 32, ...
template <int HE>
__device__ __forceinline__ void rows_heavy_fold(const uint32_t *zs, const double2 *cs, uint32_t n, uint32_t r,
                                                double (&hre)[4], double (&him)[4])
{
#pragma unroll 2
    for (uint32_t t = 0; t < n; t++) {
        const uint32_t z = zs[t];                                  // shared memory, warp-uniform (broadcast)
        const double2 c = cs[t];
#pragma unroll
        for (int e = 0; e < HE; e++) {
            const double sg = pm_one((uint32_t)__popc((r + 32u * e) & z));
            hre[e] = __fma_rn(sg, c.x, hre[e]); him[e] = __fma_rn(sg, c.y, him[e]);
        }
    }
}

// The CTA-wide heavy phase for the strip of HS rows that starts at row `sbase`.  s_hd[q] = {offset of the group's
// terms 1.. in the shared term table, their number, z of term 0, group}; s_h0c[q] = c' of term 0.
template <int HE>
__device__ __forceinline__ void rows_heavy_item(const uint4 d, const double2 c0, double2 *hv, const uint32_t *s_ez,
                                                const double2 *s_ec, uint32_t r, uint32_t lane)
{
    double hre[4], him[4];
#pragma unroll
    for (int e = 0; e < HE; e++) {
        const uint32_t s = (uint32_t)(__popc((r + 32u * e) & d.z) & 1) << 31;
        hre[e] = flip_sign(c0.x, s); him[e] = flip_sign(c0.y, s);
    }
    rows_heavy_fold<HE>(s_ez + d.x, s_ec + d.x, d.y, r, hre, him);
#pragma unroll
    for (int e = 0; e < HE; e++) hv[lane + 32u * e] = make_double2(hre[e], him[e]);
}

// Heavy groups round-robin over the warps, lane <-> row, up to 4 rows per lane (the rows of a lane are independent FP64
// chains that hide the latency of the group's one dependent chain per row); when there are fewer heavy groups than half
// the warps, (group, 32 rows) items are spread over the warps instead.
template <int TH>
__device__ __forceinline__ void rows_heavy_phase(const uint4 *s_hd, const double2 *s_h0c, double2 *s_hv, const uint32_t *s_ez,
                                                 const double2 *s_ec, const uint32_t *s_ord, uint32_t n_heavy, uint32_t HS,
                                                 uint32_t hv_log2, uint32_t sbase)
{
    constexpr uint32_t NW = TH / 32u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (HS < 32u) {                                                // runs shorter than a warp (tiny matrices)
        for (uint32_t q = warp; q < n_heavy; q += NW) {
            const uint4 d = s_hd[q];
            const double2 c0 = s_h0c[q];
            double hre[4], him[4];
            const uint32_t sg = (uint32_t)(__popc((sbase + lane) & d.z) & 1) << 31;
            hre[0] = flip_sign(c0.x, sg); him[0] = flip_sign(c0.y, sg);
            rows_heavy_fold<1>(s_ez + d.x, s_ec + d.x, d.y, sbase + lane, hre, him);
            if (lane < HS) s_hv[(q << hv_log2) + lane] = make_double2(hre[0], him[0]);
        }
        return;
    }
    if (2u * n_heavy >= NW) {
        // at least half as many heavy groups as warps (molecular Hamiltonians: a 100-300-term Z-only group, a few dozen of
        // 13-140 terms).  Items = (group, 64 rows: 2 rows per lane, two independent chains per component hide the FP64
        // latency), dealt to the warps longest first and in snake order (s_ord: the groups sorted by term count), so that
        // no warp is left with the two longest folds while fifteen wait at the barrier (ncu, H12: 30 % of the warp
        // cycles were barrier stalls when group q simply went to warp q mod 16).
        if (HS >= 64u) {
            const uint32_t n_chunks = HS >> 6, n_items = n_heavy * n_chunks;
            for (uint32_t round = 0, base = 0; base < n_items; round++, base += NW) {
                const uint32_t item = base + ((round & 1u) ? NW - 1u - warp : warp);
                if (item >= n_items) continue;
                const uint32_t q = s_ord[item / n_chunks], ch = item % n_chunks;
                rows_heavy_item<2>(s_hd[q], s_h0c[q], s_hv + (q << hv_log2) + (ch << 6), s_ez, s_ec, sbase + (ch << 6) + lane, lane);
            }
        } else {
            for (uint32_t round = 0, base = 0; base < n_heavy; round++, base += NW) {
                const uint32_t item = base + ((round & 1u) ? NW - 1u - warp : warp);
                if (item >= n_heavy) continue;
                const uint32_t q = s_ord[item];
                rows_heavy_item<1>(s_hd[q], s_h0c[q], s_hv + (q << hv_log2), s_ez, s_ec, sbase + lane, lane);
            }
        }
        return;
    }
    // a few heavy groups (one long diagonal group: spin chains): (group, 32 rows) items spread over the warps
    const uint32_t n_chunks = HS >> 5;
    for (uint32_t item = warp; item < n_heavy * n_chunks; item += NW) {
        const uint32_t q = item / n_chunks, ch = item - q * n_chunks;
        rows_heavy_item<1>(s_hd[q], s_h0c[q], s_hv + (q << hv_log2) + (ch << 5), s_ez, s_ec, sbase + (ch << 5) + lane, lane);
    }
}

// REGT = false: terms 1.. of every group ("extras", n_extra = T - G of them) live in shared memory.
// REGT = true : a thread keeps up to NT = LANE_TERMS (z, c') of each of its groups in registers (the lanes
//               kernel's SoA tables, padded with (0, -0.0): x + (-0.0) == x for every x, signed zeros included,
//               so padding needs no predicate); the term loop runs to the warp's longest light group; groups
//               with more terms must all be heavy (hv_thr == NT); the shared term table holds the heavy groups'
//               terms only (n_extra = their number).  For term-rich operators (molecular Hamiltonians: 5-6
//               terms per group) this removes two thirds of the kernel's shared-memory traffic.
// HEAVY = false: the plan has no heavy group; the heavy path is compiled out (its 4-rows-per-lane fold costs the
//               1024-thread instances registers they do not have: C3 6.6 TB/s without, 5.6 with).
// CS = true   : the rank-table columns are staged in shared memory (short rows; instantiated for NG = 1, REGT only).
// CL > 1      : rows too long for one CTA's shared memory.  The CTAs of a thread-block cluster of CL split the GROUPS
//               (CTA c: groups [c*Gc, (c+1)*Gc), terms in registers) and the ROWS of a batch (CTA c owns rows
//               c*RT/CL .. of every batch): each thread folds its groups for all RT rows and stores row j's entry into
//               the batch buffer of the CTA that owns row j through distributed shared memory (mapa + st.shared::cluster);
//               a cluster barrier per batch; every CTA hands its own whole rows to the TMA.  REGT, sl = 0 only.
//               (Measured slower than one CTA per row wherever both apply: distributed shared memory moves 17-21 B
//               per cycle and SM, about the SM's share of HBM; kept for term-rich operators with 1024 < G <= 2048.)
// CL == 0     : SPLIT mode, rows of any length.  The sorted masks are cut into trie subtrees of <= 1024 groups (K1b,
//               partition_kernel); a subtree's groups fill one contiguous slot range [base(r), base(r) + Gs) in every row
//               (XOR never splits a subtree), base(r) = sum_{b >= level} cnt[g0][b] * bit_b(x_g0 ^ r).  A CTA owns one
//               subtree (terms in registers) and assembles, per batch, the RT row SEGMENTS of its subtree, each handed to
//               the TMA on its own (data: always 16-byte aligned; column ids: the buffer row is shifted by one entry when
//               the segment starts at an odd entry, the aligned interior goes through the TMA and the one or two edge
//               entries through plain stores).  CTAs are shared out among the subtrees in proportion to their size; the
//               CTAs of a subtree share out the runs.  No exchange between CTAs.
struct RowsSplit {
    uint32_t n;                    // subtrees
    uint32_t g0[33];               // subtree s = sorted groups [g0[s], g0[s + 1])
    uint32_t level[32];            // its groups share the mask bits >= level[s]
    uint32_t cta0[33];             // CTAs [cta0[s], cta0[s + 1]) work on subtree s
    // whole-row variants (CL == 1), terms in registers: thread slot gl owns group perm[gl] (0xffffffff: none) instead of
    // group gl.  The host sorts the groups by term count and deals warp-sized chunks to the warps so that every warp's
    // longest groups add up to about the same (a warp folds to its LONGEST group; with the groups in mask order a few
    // warps fold 6 + 6 terms while most could do with 4 + 4, and the batch barrier waits for them).
    // Split mode: the same per subtree, perm[s * perm_n + gl] (perm_n slots per subtree).
    const uint32_t *perm;
    uint32_t perm_n;
    // EXTERNAL heavy values (split mode on term-rich operators): the groups of more than hv_thr terms are folded by
    // heavy_values_kernel for a chunk of rows before this launch -- ext_hv[h * ext_rows + (row - ext_row0)], h =
    // ext_hidx[group] -- instead of by the CTA that owns them, so that no subtree's CTAs carry the 100-300-term groups
    // of the mask-0 neighbourhood, no shared memory goes to heavy tables and there is no heavy phase (and barrier).
    const double2 *ext_hv;
    const uint32_t *ext_hidx;
    uint32_t ext_rows, ext_row0;
    uint32_t dec_block;            // DEC: 1 = the warp that hands a batch over waits for the TMA's read at once (else after its next batch)
};

// Values of the heavy groups for rows [row0, row0 + n_rows): warp <-> (heavy group, 32 * E rows), lane <-> E rows, the
// reference's left-to-right fold (group_values: bit-exact).  heavy_g is sorted longest first, so the long folds start
// first.  out[h * n_rows + (row - row0)]: a warp's stores are contiguous.
template <int E>
__global__ void __launch_bounds__(256)
heavy_values_kernel(PlanDev p, const uint32_t *__restrict__ heavy_g, uint32_t nh, uint64_t row0, uint32_t n_rows,
                    double2 *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31u, blocks_per_h = n_rows / (32u * E);
    const uint64_t item = (uint64_t)blockIdx.x * 8u + (threadIdx.x >> 5);
    if (item >= (uint64_t)nh * blocks_per_h) return;
    const uint32_t h = (uint32_t)(item / blocks_per_h), rb = (uint32_t)(item % blocks_per_h);
    const uint32_t g = __ldg(&heavy_g[h]);
    uint32_t r[E];
#pragma unroll
    for (int e = 0; e < E; e++) r[e] = (uint32_t)row0 + rb * 32u * E + 32u * e + lane;
    double ar[E], ai[E];
    group_values<E>(p, __ldg(&p.goff[g]), __ldg(&p.goff[g + 1]), r, ar, ai);
#pragma unroll
    for (int e = 0; e < E; e++) out[(size_t)h * n_rows + (r[e] - (uint32_t)row0)] = make_double2(ar[e], ai[e]);
}


__device__ __forceinline__ void mbar_arrive(uint64_t *bar)     // release.cta: the arriving thread's earlier writes are visible to the waiter
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// DEC = true  : DECOUPLED warps (CL <= 1).  The batch barrier of the plain variant makes every warp wait for the slowest one
//               of every batch (ncu, H12 / H10 split: 30-35 % of the warp cycles are barrier stalls at 16 warps per SM, and
//               while the last warps of a batch finish, their schedulers have nothing else to issue).  Here the two batch
//               buffers are handed over through mbarriers: a warp arrives on full[b] when its entries of the batch are in
//               buffer b and goes straight on to the next batch as soon as empty[b ^ 1] says that the TMA has read the other
//               buffer out.  The warp that arrives FIRST -- the one with time to spare -- takes the batch's hand-over: it
//               waits for full[b], gives the buffer to the TMA, waits for the read and arrives on empty[b] (a bulk group can
//               only be waited for by the thread that committed it, and a 17th warp would cost every thread 32 registers).
//               A fast warp runs up to one batch ahead of the slowest.  The in-CTA heavy phase keeps two named barriers per strip.
template <int NG, int Q, int TH, bool REGT, bool HEAVY, bool CS, int CL, bool DEC = false>
__global__ void __launch_bounds__(TH, 1)
fill_rows_kernel(PlanDev p, uint32_t G, uint32_t Gc, uint32_t n_extra, uint32_t log2R, uint32_t sl, uint32_t n_runs,
                 uint32_t hv_thr, uint32_t hv_cap, uint32_t hv_log2, uint64_t tile_row0, uint64_t row_lo,
                 uint64_t indptr_base, uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices,
                 double2 *__restrict__ data, uint64_t indptr_last_row, const __grid_constant__ RowsSplit sp)
{
    constexpr uint32_t RT = 1u << Q;
    constexpr bool SPLIT = CL == 0;
    constexpr int NS = Q + 2;                                      // row bits whose slot step lives in a register
    constexpr int NT = REGT ? LANE_TERMS : 1;                      // terms of a group kept in registers
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t QB = (uint32_t)Q + sl;                          // log2(rows per batch): 2^sl sub-batches of RT rows
    constexpr uint32_t CLD = CL > 1 ? CL : 1;
    static_assert(CL <= 1 || (REGT && !CS && RT % CLD == 0), "cluster variant: terms in registers, whole rows per CTA");
    static_assert(!SPLIT || (REGT && !CS), "split variant: terms in registers");
    constexpr uint32_t ROWS_OWN = RT / CLD;                        // CL > 1: rows of a batch this CTA hands to the TMA
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    // this CTA's groups [gbase, gbase + Gn) and its share jcta, jcta + n_share, ... of the runs
    uint32_t gbase = crank * Gc, Gn = CL == 1 ? G : Gc, jcta = blockIdx.x / CLD, n_share = gridDim.x / CLD, level = 32u, sid = 0u;
    if constexpr (SPLIT) {
        while (sid + 1u < sp.n && blockIdx.x >= sp.cta0[sid + 1u]) sid++;
        gbase = sp.g0[sid]; Gn = sp.g0[sid + 1u] - gbase; level = sp.level[sid];
        jcta = blockIdx.x - sp.cta0[sid]; n_share = sp.cta0[sid + 1u] - sp.cta0[sid];
    }
    const uint32_t GW = SPLIT ? Gn : G;                            // entries of one buffered row (segment)
    const uint32_t SI = SPLIT ? ((Gn + 3u) & ~1u) : G;             // pitch of a buffered row of column ids (room for the shift)
    const uint32_t tile_n = (CL > 1 ? ROWS_OWN : (RT << sl)) * (SPLIT ? SI : G);   // entries of one batch buffer (SPLIT: its id pitch)
    const uint32_t GP = (uint32_t)TH >> sl;                        // threads per sub-batch (>= G when sl > 0; host-checked)
    const uint32_t tg = threadIdx.x & (GP - 1u), sub = threadIdx.x / GP;   // this thread's group slot and sub-batch
    double2 *sdat = reinterpret_cast<double2 *>(smem_raw);                               // [2][tile_n]
    uint64_t *sidx = reinterpret_cast<uint64_t *>(smem_raw + (size_t)tile_n * 32u);      // [2][tile_n]
    double2 *s_ec = reinterpret_cast<double2 *>(smem_raw + (size_t)tile_n * 48u);        // [n_extra]
    uint32_t *s_ez = reinterpret_cast<uint32_t *>(s_ec + n_extra);                       // [n_extra]
    double2 *s_hv = reinterpret_cast<double2 *>(smem_raw + (((size_t)tile_n * 48u + (size_t)n_extra * 20u + 15u) & ~(size_t)15u));   // [hv_cap][2^hv_log2]
    double2 *s_h0c = s_hv + ((size_t)hv_cap << hv_log2);                                 // [hv_cap]
    uint4 *s_hd = reinterpret_cast<uint4 *>(s_h0c + hv_cap);                             // [hv_cap]
    uint32_t *s_ord = reinterpret_cast<uint32_t *>(s_hd + hv_cap);                       // [hv_cap] heavy groups, longest first
    uint32_t *s_cnt = s_ord + hv_cap;                                                    // [n_qubits][G] when CS
    __shared__ uint32_t s_nheavy, s_hterms;
    __shared__ uint32_t s_cb[32];                                  // SPLIT: cnt[g0][b] for b >= level (else 0), x of g0 in s_cb_x
    __shared__ uint32_t s_cb_x;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];       // DEC: buffer b written by every warp / read out by the TMA
    __shared__ uint32_t s_arrived[2];                              // DEC: warps that have arrived on full[b] (the first one hands over)
    static_assert(!DEC || CL <= 1, "decoupled variant: one CTA per row run or split mode");
    const uint32_t T = p.n_terms, nq = (uint32_t)p.n_qubits;
    constexpr uint32_t NOT_HEAVY = 0xffffffffu;
    if (threadIdx.x == 0) {
        s_nheavy = 0; s_hterms = 0;
        if constexpr (DEC) {
            s_arrived[0] = 0; s_arrived[1] = 0;
            mbar_init(&s_full[0], TH / 32u); mbar_init(&s_full[1], TH / 32u); mbar_init(&s_empty[0], 1u); mbar_init(&s_empty[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (SPLIT && threadIdx.x < 32u) {
        s_cb[threadIdx.x] = (threadIdx.x >= level && threadIdx.x < nq) ? __ldg(&p.cnt_t[threadIdx.x * T + gbase]) : 0u;
        if (threadIdx.x == 0) s_cb_x = __ldg(&p.gx[gbase]);
    }
    // short rows: the rank-table columns (n_qubits * G words) fit in shared memory, so neither the slot of a run's
    // first row (n_qubits reads per group and run) nor a Gray step waits on L2
    if constexpr (CS)
        for (uint32_t i = threadIdx.x; i < nq * G; i += TH) { const uint32_t b = i / G; s_cnt[i] = __ldg(&p.cnt_t[b * T + (i - b * G)]); }
    __syncthreads();

    // ---- once per CTA: this thread's groups ------------------------------------------------------
    uint32_t x[NG], z[NG][NT], eb[NG], ee[NG], off[NG], hidx[NG], gid[NG];
    int32_t sd[NG][NS];
    double cr[NG][NT], ci[NG][NT];
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const uint32_t gl = tg + (uint32_t)k * GP;                          // slot inside the CTA
        uint32_t g = gbase + gl;                                            // ... and its group
        bool mine = gl < Gn && g < G;
        if (CL == 1 && sp.perm != nullptr) { g = gl < sp.perm_n ? __ldg(&sp.perm[gl]) : NOT_HEAVY; mine = g < G; }
        if (SPLIT && sp.perm != nullptr) { g = gl < sp.perm_n ? __ldg(&sp.perm[sid * sp.perm_n + gl]) : NOT_HEAVY; mine = g < G; }
        gid[k] = mine ? g : NOT_HEAVY;
        const uint32_t gg = g < G ? g : G - 1u;
        const uint32_t t0 = __ldg(&p.goff[gg]), t1 = __ldg(&p.goff[gg + 1]);
        x[k] = __ldg(&p.gx[gg]);
        // heavy groups (more than hv_thr terms: the Z-only group of a molecular Hamiltonian, a few dozen others)
        // are evaluated lane <-> row for a strip of rows at a time by the whole CTA, not inside their owner's lane
        hidx[k] = NOT_HEAVY;
        if (HEAVY && sp.ext_hv != nullptr) {                       // heavy values come from heavy_values_kernel
            if (mine && t1 - t0 > hv_thr) hidx[k] = __ldg(&sp.ext_hidx[gg]);
        } else if (HEAVY && mine && sub == 0u && t1 - t0 > hv_thr) {
            const uint32_t h = atomicAdd(&s_nheavy, 1u);
            if (h < hv_cap) {
                hidx[k] = h;
                uint32_t o = t0 - gg;                              // shared-memory variant: its slice of the extras table
                if constexpr (REGT) {                              // register variant: the table holds heavy groups only
                    o = atomicAdd(&s_hterms, t1 - t0 - 1u);
                    for (uint32_t t = t0 + 1u; t < t1; t++) { s_ez[o + t - t0 - 1u] = __ldg(&p.tz[t]); s_ec[o + t - t0 - 1u] = __ldg(&p.tc[t]); }
                }
                s_hd[h] = make_uint4(o, t1 - t0 - 1u, __ldg(&p.tz[t0]), g);
                s_h0c[h] = __ldg(&p.tc[t0]);
            }
        }
        if constexpr (REGT) {
#pragma unroll
            for (int t = 0; t < NT; t++) {
                z[k][t] = __ldg(&p.lt_z[(uint32_t)t * T + gg]);
                const double2 c = __ldg(&p.lt_c[(uint32_t)t * T + gg]);
                cr[k][t] = c.x; ci[k][t] = c.y;
            }
            // ee: terms the warp folds for this group slot = its longest light group (warp-uniform)
            const uint32_t nt = (mine && !(HEAVY && t1 - t0 > hv_thr)) ? min(t1 - t0, (uint32_t)NT) : 1u;
            eb[k] = 0; ee[k] = __reduce_max_sync(0xffffffffu, nt);
        } else {
            z[k][0] = __ldg(&p.tz[t0]);
            const double2 c = __ldg(&p.tc[t0]);
            cr[k][0] = c.x; ci[k][0] = c.y;
            eb[k] = t0 - gg; ee[k] = t1 - gg - 1u;                 // sorted term t > t0 of group g is extra t - g - 1
            if (mine)
                for (uint32_t t = t0 + 1u; t < t1; t++) { s_ez[t - gg - 1u] = __ldg(&p.tz[t]); s_ec[t - gg - 1u] = __ldg(&p.tc[t]); }
        }
#pragma unroll
        for (int b = 0; b < NS; b++) {                             // bits 0..Q-1 (rows of the thread), then QB and QB + 1
            const uint32_t bit = b < Q ? (uint32_t)b : QB + (uint32_t)(b - Q);
            const int32_t cb = (int32_t)__ldg(&p.cnt_t[bit * T + gg]);
            sd[k][b] = ((x[k] >> bit) & 1u) ? -cb : cb;
        }
        off[k] = 0;
    }
    __syncthreads();

    if constexpr (HEAVY) {
        const uint32_t nh = min(s_nheavy, hv_cap);
        for (uint32_t h = threadIdx.x; h < nh; h += TH) {
            const uint32_t mine_t = s_hd[h].y;
            uint32_t rank = 0;
            for (uint32_t o = 0; o < nh; o++) { const uint32_t t = s_hd[o].y; rank += (t > mine_t || (t == mine_t && o < h)) ? 1u : 0u; }
            s_ord[rank] = h;
        }
        __syncthreads();
    }
    if constexpr (CL > 1) cluster_barrier();                       // every CTA of the cluster is resident before the first remote store
    if (HEAVY && sl != 0u && sp.ext_hv == nullptr) {
        // sub-batches > 0 own the same groups as sub-batch 0: look the heavy index up in the descriptor table
        const uint32_t nh = min(s_nheavy, hv_cap);
#pragma unroll
        for (int k = 0; k < NG; k++)
            if (sub != 0u) {
                hidx[k] = NOT_HEAVY;
                for (uint32_t h = 0; h < nh; h++) if (s_hd[h].w == gid[k]) hidx[k] = h;
            }
    }
    const uint32_t R = 1u << log2R, n_batches = R >> QB;
    const uint32_t n_heavy = HEAVY ? min(s_nheavy, hv_cap) : 0u;
    const uint32_t HS = min(R, 1u << hv_log2), SB = HS >> QB;      // rows / batches per heavy strip (HS >= 2^QB; host-checked)
    uint32_t parity = 0;                                           // buffer of the current batch
    uint32_t use = 0;                                              // DEC: batches done so far (buffer use = use >> 1)
    uint32_t pending = 2u;                                         // DEC: buffer this warp handed to the TMA and has not yet declared empty
    for (uint32_t run = jcta; run < n_runs; run += n_share) {
        const uint64_t r0_64 = tile_row0 + ((uint64_t)run << log2R);   // first row of the run (aligned to R)
        const uint32_t r0 = (uint32_t)r0_64;
        if (indptr != nullptr && crank == 0u && sid == 0u)
            for (uint32_t i = threadIdx.x; i < R; i += TH) {
                const uint64_t lr = r0_64 + i - row_lo;
                indptr[lr] = indptr_base + lr * G;
                if (lr + 1 == indptr_last_row) indptr[lr + 1] = indptr_base + (lr + 1) * G;
            }
        // slot of every group in the first row this thread handles in the run's first batch
        // (summing the NEXT run's slots one rank-table word per batch during the current run was measured: H10 / H11 -4 %,
        // H8 -1 %, profiles/r06_summary.md -- the loop below is not where the time goes)
#pragma unroll
        for (int k = 0; k < NG; k++) {
            const uint32_t g = gid[k], gg = g < G ? g : G - 1u, xr = x[k] ^ (r0 + (sub << Q));   // the thread's first row
            uint32_t o = 0;
            for (uint32_t b = 0; b < nq; b++) {
                const uint32_t cb = CS ? s_cnt[b * G + gg] : __ldg(&p.cnt_t[b * T + gg]);
                o += ((xr >> b) & 1u) ? cb : 0u;
            }
            off[k] = o;
        }
        uint32_t rb = r0;                                          // first row of the batch (bits < QB are 0)
        uint32_t base = 0;                                         // SPLIT: first slot of the subtree in the rows of the batch
        if constexpr (SPLIT) {
            const uint32_t xr = s_cb_x ^ r0;
            for (uint32_t b = level; b < nq; b++) base += ((xr >> b) & 1u) ? s_cb[b] : 0u;
        }
        for (uint32_t i = 0; i < n_batches; i++) {
            if (i != 0u) {                                         // Gray code over batches: step i flips row bit QB + ctz(i)
                const uint32_t b = QB + (uint32_t)__ffs((int)i) - 1u;
                rb ^= 1u << b;
                const bool up = (rb >> b) & 1u;                    // CTA-uniform
                if constexpr (SPLIT) base += (((s_cb_x ^ rb) >> b) & 1u) ? s_cb[b] : 0u - s_cb[b];   // 0 for b < level
#pragma unroll
                for (int k = 0; k < NG; k++) {
                    int32_t s;
                    if (b == QB) s = sd[k][Q];
                    else if (b == QB + 1u) s = sd[k][Q + 1];
                    else {
                        const uint32_t g = gid[k], gg = g < G ? g : G - 1u;
                        const int32_t cb = (int32_t)(CS ? s_cnt[b * G + gg] : __ldg(&p.cnt_t[b * T + gg]));
                        s = ((x[k] >> b) & 1u) ? -cb : cb;
                    }
                    off[k] += (uint32_t)(up ? s : -s);
                }
            }
            if (HEAVY && n_heavy != 0u && (i & (SB - 1u)) == 0u) {
                // the next SB batches stay inside the aligned strip of HS rows that holds rb: every warp takes heavy
                // groups round-robin, lane <-> row, warp-uniform term loop (terms broadcast to the warp).
                // The barrier that ended the previous batch also retired the last reader of s_hv.
                // A lane folds rows lane, lane + 32, ... of the strip together: the fold is one dependent chain per
                // row, so the rows of a lane are what hides the FP64 latency of the longest group.
                if constexpr (DEC) lanes_barrier(TH);              // no batch barrier here: retire the readers of the previous strip
                rows_heavy_phase<TH>(s_hd, s_h0c, s_hv, s_ez, s_ec, s_ord, n_heavy, HS, hv_log2, rb & ~(HS - 1u));
                if constexpr (DEC) lanes_barrier(TH); else __syncthreads();
            }
            if (HEAVY && sp.ext_hv != nullptr && i + 1u < n_batches) {
                // external heavy values of the NEXT batch -> L1 (RT rows = one 16 * RT-byte piece per heavy group): the load
                // that feeds the store below otherwise waits a full L2 round trip in every batch
                const uint32_t rn = (rb ^ (1u << (QB + (uint32_t)__ffs((int)(i + 1u)) - 1u))) + (sub << Q);
#pragma unroll
                for (int k = 0; k < NG; k++)
                    if (hidx[k] != NOT_HEAVY)
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(sp.ext_hv + (size_t)hidx[k] * sp.ext_rows + (rn - sp.ext_row0)));
            }
            if constexpr (DEC) mbar_wait(&s_empty[parity], ((use >> 1) & 1u) ^ 1u);   // the TMA has read this buffer's previous batch out
            double2 *bd = sdat + parity * tile_n;
            uint64_t *bi = sidx + parity * tile_n;
            const uint32_t rt = rb + (sub << Q);                   // first of this thread's RT rows (bits < Q are 0)
            // SPLIT: where slot 0 of the row would sit in segment j of the batch buffer (CTA-uniform, once per batch):
            // data j * GW - base; ids j * SI - base + the parity of the segment's first entry ((rt + j - row_lo) * G + base)
            uint32_t seg_d[RT], seg_i[RT];
            if constexpr (SPLIT) {
                const uint32_t par0 = (uint32_t)(((uint64_t)rt - row_lo) * G + base) & 1u;
#pragma unroll
                for (uint32_t j = 0; j < RT; j++) { seg_d[j] = j * GW - base; seg_i[j] = j * SI - base + (par0 ^ (j & G & 1u)); }
            }
#pragma unroll
            for (int k = 0; k < NG; k++) {
                if (gid[k] != NOT_HEAVY) {
                    double re[RT], im[RT];
                    if (HEAVY && hidx[k] != NOT_HEAVY) {
                        if (sp.ext_hv != nullptr) {
                            const double2 *src = sp.ext_hv + (size_t)hidx[k] * sp.ext_rows + (rt - sp.ext_row0);
#pragma unroll
                            for (uint32_t j = 0; j < RT; j++) { const double2 v = __ldg(&src[j]); re[j] = v.x; im[j] = v.y; }
                        } else {
#pragma unroll
                            for (uint32_t j = 0; j < RT; j++) {
                                const double2 v = s_hv[(hidx[k] << hv_log2) + ((rt + j) & (HS - 1u))];
                                re[j] = v.x; im[j] = v.y;
                            }
                        }
                    } else {
                        // rt has no bits below Q, so popc((rt + j) & z) = popc(rt & z) + popc(j & z): one POPC per term
                        // serves the RT rows; bit 0 of p0 ^ (z >> b) ^ ... is row j's parity
                        const uint32_t p0 = (uint32_t)__popc(rt & z[k][0]);
#pragma unroll
                        for (uint32_t j = 0; j < RT; j++) {
                            const uint32_t s = rows_parity<Q>(p0, z[k][0], j) << 31;
                            re[j] = flip_sign(cr[k][0], s); im[j] = flip_sign(ci[k][0], s);
                        }
                        // later terms: fma(+-1.0, c', acc) is the sign flip and the __dadd_rn of the fold in one
                        // instruction per component (the product is exact, signed zeros included)
                        if constexpr (REGT) {
#pragma unroll
                            for (int t = 1; t < NT; t++) {
                                if ((uint32_t)t < ee[k]) {             // warp-uniform
                                    const uint32_t pt = (uint32_t)__popc(rt & z[k][t]);
#pragma unroll
                                    for (uint32_t j = 0; j < RT; j++) {
                                        const double sg = pm_one(rows_parity<Q>(pt, z[k][t], j));
                                        re[j] = __fma_rn(sg, cr[k][t], re[j]); im[j] = __fma_rn(sg, ci[k][t], im[j]);
                                    }
                                }
                            }
                        } else {
#pragma unroll 1
                            for (uint32_t e = eb[k]; e < ee[k]; e++) {
                                const uint32_t ze = s_ez[e];
                                const double2 c = s_ec[e];
                                const uint32_t pe = (uint32_t)__popc(rt & ze);
#pragma unroll
                                for (uint32_t j = 0; j < RT; j++) {
                                    const double sg = pm_one(rows_parity<Q>(pe, ze, j));
                                    re[j] = __fma_rn(sg, c.x, re[j]); im[j] = __fma_rn(sg, c.y, im[j]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (uint32_t j = 0; j < RT; j++) {
                        uint32_t o = (CL > 1 ? (j % ROWS_OWN) : ((sub << Q) + j)) * G + off[k];
#pragma unroll
                        for (int b = 0; b < Q; b++) if ((j >> b) & 1u) o += (uint32_t)sd[k][b];
                        if constexpr (SPLIT) {                     // segment-local slot; ids shifted by the parity of the segment start
                            const uint32_t slot = o - j * G;       // = off[k] + the steps of row j's low bits (sub == 0 here)
                            bd[seg_d[j] + slot] = make_double2(re[j], im[j]);
                            bi[seg_i[j] + slot] = (uint64_t)((rt + j) ^ x[k]);
                        } else if constexpr (CL > 1) {                    // row j belongs to CTA j / ROWS_OWN of the cluster
                            const uint32_t owner = j / ROWS_OWN;
                            st_cluster_f64x2(mapa_shared((uint32_t)__cvta_generic_to_shared(bd + o), owner), re[j], im[j]);
                            st_cluster_u64(mapa_shared((uint32_t)__cvta_generic_to_shared(bi + o), owner), (uint64_t)((rt + j) ^ x[k]));
                        } else {
                            bd[o] = make_double2(re[j], im[j]);
                            bi[o] = (uint64_t)((rt + j) ^ x[k]);
                        }
                    }
                }
            }
            // generic-proxy writes -> visible to the async proxy; the copy that last read the OTHER buffer
            // (issued one batch ago) must have drained before anyone refills it after the barrier
            if constexpr (CL > 1) {
                asm volatile("fence.proxy.async;" ::: "memory");           // this thread's (remote) writes -> async proxy
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                cluster_barrier();                                         // release / acquire across the cluster
            } else if constexpr (DEC) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t lane = threadIdx.x & 31u;
                if (pending != 2u) {                               // warp-uniform: the batch this warp handed over last time
                    if (lane < (SPLIT ? RT : 1u)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    if (lane == 0u) mbar_arrive(&s_empty[pending]);
                    pending = 2u;
                }
                __syncwarp();
                uint32_t first = 0;
                if (lane == 0u) { first = atomicAdd(&s_arrived[parity], 1u) == 0u ? 1u : 0u; mbar_arrive(&s_full[parity]); }
                first = __shfl_sync(0xffffffffu, first, 0);
                if (first) {                                       // warp-uniform: this warp hands the batch over
                    mbar_wait(&s_full[parity], (use >> 1) & 1u);
                    if (lane == 0u) s_arrived[parity] = 0u;        // nobody arrives here again before empty[parity]
                    if constexpr (SPLIT) {
                        if (lane < RT) {                           // one lane per row segment
                            const uint32_t j = lane;
                            const uint64_t gs = ((uint64_t)(rb + j) - row_lo) * G + base;
                            const uint32_t par = (uint32_t)gs & 1u, n_al = (Gn - par) & ~1u;
                            bulk_store_smem_to_global(data + gs, bd + j * GW, Gn * 16u);
                            if (n_al) bulk_store_smem_to_global(indices + gs + par, bi + j * SI + 2u * par, n_al * 8u);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            if (par) indices[gs] = bi[j * SI + 1u];
                            if ((Gn - par) & 1u) indices[gs + Gn - 1u] = bi[j * SI + par + Gn - 1u];
                        }
                    } else if (lane == 0u) {
                        const uint64_t o = ((uint64_t)rb - row_lo) * G;
                        bulk_store_smem_to_global(data + o, bd, tile_n * 16u);
                        bulk_store_smem_to_global(indices + o, bi, tile_n * 8u);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    // the buffer is free once the TMA has read it: waited for after this warp's NEXT batch (by then the read
                    // is normally over; waiting here would serialise the copy with this warp's share of the next batch)
                    pending = parity;
                    if (sp.dec_block) {
                        if (lane < (SPLIT ? RT : 1u)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                        if (lane == 0u) mbar_arrive(&s_empty[parity]);
                        pending = 2u;
                    }
                }
                parity ^= 1u; use++;
                continue;
            } else {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (threadIdx.x < (SPLIT ? RT : 1u)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
            }
            if constexpr (SPLIT) {
                if (threadIdx.x < RT) {                            // one thread per row segment
                    const uint32_t j = threadIdx.x;
                    const uint64_t gs = ((uint64_t)(rb + j) - row_lo) * G + base;     // first entry of the segment in the outputs
                    const uint32_t par = (uint32_t)gs & 1u, n_al = (Gn - par) & ~1u;
                    bulk_store_smem_to_global(data + gs, bd + j * GW, Gn * 16u);
                    if (n_al) bulk_store_smem_to_global(indices + gs + par, bi + j * SI + 2u * par, n_al * 8u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (par) indices[gs] = bi[j * SI + 1u];
                    if ((Gn - par) & 1u) indices[gs + Gn - 1u] = bi[j * SI + par + Gn - 1u];
                }
            } else if (threadIdx.x == 0) {
                if constexpr (CL > 1) asm volatile("fence.proxy.async;" ::: "memory");
                const uint64_t o = ((uint64_t)rb + crank * ROWS_OWN - row_lo) * G;
                bulk_store_smem_to_global(data + o, bd, tile_n * 16u);
                bulk_store_smem_to_global(indices + o, bi, tile_n * 8u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            parity ^= 1u;
        }
    }
    if (DEC || threadIdx.x < (SPLIT ? RT : 1u)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if constexpr (CL > 1) cluster_barrier();                       // nobody leaves while a peer could still address its memory
}

}  // namespace qr
