// compact.cuh -- K2: per-row nnz count -> prefix scan -> indptr, and the compaction that uses it.
//
// The reference keeps explicit zeros in the build (accel.rs:171-210) and removes them in a
// separate opt-in pass, util::csmatrix_eliminate_zeroes (qrusty/src/util.rs:154-171): keep an
// entry iff norm() > tolerance (norm = hypot(re, im)), row-major order preserved;
// csmatrix_nz (util.rs:144-152) counts the complement.  Here that pass runs on the device-
// resident CSR shard (uniform row length G, as every plan-built shard has):
//   count_kept_kernel    kept entries per row                       (reads 16 B/entry)
//   scan_*               exclusive prefix sum of the counts -> indptr, built on the warp scan
//                        of scan.cuh: tile sums -> one CTA scans the sums -> tiles re-scanned
//   compact_rows_kernel  warp per row: ballot + popc rank, entries written at indptr[row] + rank
#pragma once
#include "fill.cuh"
#include "scan.cuh"
#include <cuda_runtime.h>

namespace qr {

constexpr int K2_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                                // per thread
constexpr int SCAN_TILE = K2_THREADS * SCAN_ITEMS;

// norm() > tol with norm = hypot(re, im) (util.rs:159).  max(|re|,|im|) <= hypot <= |re| + |im| decides
// almost every entry (exact zeros, values far above tol) without evaluating hypot.
__device__ __forceinline__ bool keep_entry(double2 d, double tol)
{
    const double a = fabs(d.x), b = fabs(d.y);
    if (fmax(a, b) > tol) return true;
    if (a + b <= tol) return false;
    return hypot(d.x, d.y) > tol;
}

// counts[r + 1] = #{entries of row r with norm > tol}; counts[0] = 0.  CTA b owns rows [b*R, (b+1)*R).
__global__ void __launch_bounds__(K2_THREADS)
count_kept_kernel(uint64_t n_rows, uint32_t G, uint32_t R, const double2 *__restrict__ data, double tol,
                  uint64_t *__restrict__ counts, uint32_t write_zero = 1u)     // write_zero = 0: a later window of a windowed count
{
    extern __shared__ uint32_t s_cnt[];                       // [R]
    const uint64_t row0 = (uint64_t)blockIdx.x * R;
    const uint32_t nr = (uint32_t)min((uint64_t)R, n_rows - row0);
    for (uint32_t l = threadIdx.x; l < nr; l += K2_THREADS) s_cnt[l] = 0;
    __syncthreads();
    const uint64_t base = row0 * G, total = (uint64_t)nr * G;
    for (uint64_t i = threadIdx.x; i < total; i += K2_THREADS)
        if (keep_entry(data[base + i], tol)) atomicAdd(&s_cnt[(uint32_t)(i / G)], 1u);
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < nr; l += K2_THREADS) counts[row0 + l + 1] = s_cnt[l];
    if (write_zero && blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
}

// ---- three-phase scan over a[1..n] (a[0] = 0 stays): inclusive, in place ----------------------
__global__ void __launch_bounds__(K2_THREADS)
scan_tile_sums_kernel(uint64_t n, const uint64_t *__restrict__ a, uint64_t *__restrict__ tile_sums)
{
    __shared__ uint64_t scratch[33];
    const uint64_t i0 = 1 + (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (i0 + k <= n) s += a[i0 + k];
    uint64_t total;
    block_exclusive_scan(s, scratch, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(K2_THREADS)
scan_tile_offsets_kernel(uint64_t n_tiles, uint64_t *__restrict__ tile_sums, uint64_t *__restrict__ total_out)
{
    __shared__ uint64_t scratch[33];
    uint64_t carry = 0;
    for (uint64_t t0 = 0; t0 < n_tiles; t0 += K2_THREADS) {
        const uint64_t t = t0 + threadIdx.x;
        const uint64_t v = t < n_tiles ? tile_sums[t] : 0;
        uint64_t total;
        const uint64_t excl = block_exclusive_scan(v, scratch, &total);
        if (t < n_tiles) tile_sums[t] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(K2_THREADS)
scan_apply_kernel(uint64_t n, uint64_t *__restrict__ a, const uint64_t *__restrict__ tile_offsets)
{
    __shared__ uint64_t scratch[33];
    const uint64_t i0 = 1 + (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = i0 + k <= n ? a[i0 + k] : 0; s += v[k]; }
    uint64_t total;
    uint64_t run = tile_offsets[blockIdx.x] + block_exclusive_scan(s, scratch, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { run += v[k]; if (i0 + k <= n) a[i0 + k] = run; }
}

// Warp per row: entries of a row are read in column order 32 at a time; a kept entry lands at
// indptr[row] + (number of kept entries before it) -- ballot + popc, no shared memory.
__global__ void __launch_bounds__(K2_THREADS)
compact_rows_kernel(uint64_t n_rows, uint32_t G, const uint64_t *__restrict__ indices_in,
                    const double2 *__restrict__ data_in, double tol, const uint64_t *__restrict__ indptr,
                    uint64_t *__restrict__ indices_out, double2 *__restrict__ data_out)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (K2_THREADS / 32);
    for (uint64_t row = (uint64_t)blockIdx.x * (K2_THREADS / 32) + (threadIdx.x >> 5); row < n_rows; row += warps) {
        uint64_t out = indptr[row];
        const uint64_t in0 = row * G;
        for (uint32_t j0 = 0; j0 < G; j0 += 32u) {
            const uint32_t j = j0 + lane;
            double2 d = make_double2(0.0, 0.0);
            if (j < G) d = data_in[in0 + j];
            const bool keep = j < G && keep_entry(d, tol);
            const unsigned mask = __ballot_sync(FULL_MASK, keep);
            if (keep) {
                const uint64_t pos = out + __popc(mask & ((1u << lane) - 1u));
                data_out[pos] = d;
                indices_out[pos] = indices_in[in0 + j];
            }
            out += __popc(mask);
        }
    }
}


// ---------------------------------------------------------------------------------
// Fused drop-zeros build: the CSR util::csmatrix_eliminate_zeroes (util.rs:154-171) would
// produce from the reference's build, without ever writing the explicit zeros.
//   count_rows_kernel    evaluates every (row, group) value in registers, counts the kept
//                        ones per row (no matrix traffic: 8 B written per row)
//   scan_*               counts -> indptr (above)
//   fill_compact_kernel  assembles a 32-row tile in shared memory in column order exactly as
//                        fill_staged_kernel does, then compacts the tile CTA-wide.
//                        24 B written per KEPT entry.
// ---------------------------------------------------------------------------------
constexpr int COUNT_ROWS_WARPS = 8;

__global__ void __launch_bounds__(32 * COUNT_ROWS_WARPS)
count_rows_kernel(PlanDev p, uint32_t G, uint64_t row_lo, uint64_t row_hi, double tol, uint64_t *__restrict__ counts)
{
    constexpr int E = 2;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t base = row_lo + ((uint64_t)blockIdx.x * COUNT_ROWS_WARPS + warp) * (32u * E);
    if (base >= row_hi) return;
    uint32_t r[E], cnt[E];
    bool live[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint64_t r64 = base + 32u * e + lane;
        live[e] = r64 < row_hi;
        r[e] = (uint32_t)(live[e] ? r64 : row_hi - 1);
        cnt[e] = 0;
    }
    // gridDim.y > 1 (few rows, many groups: the host zeroes `counts` first): this CTA counts its slice of the groups only
    const uint32_t g_lo = (uint32_t)((uint64_t)G * blockIdx.y / gridDim.y), g_hi = (uint32_t)((uint64_t)G * (blockIdx.y + 1) / gridDim.y);
    for (uint32_t g = g_lo; g < g_hi; g++) {
        const GroupDesc d = p.gdesc[g];
        double ar[E], ai[E];
        if (d.flag & 1u) {
#pragma unroll
            for (int e = 0; e < E; e++) { ar[e] = d.cre; ai[e] = d.cim; }
        } else {
            group_values<E>(p, d.t0, d.t1, r, ar, ai);
        }
#pragma unroll
        for (int e = 0; e < E; e++) cnt[e] += keep_entry(make_double2(ar[e], ai[e]), tol) ? 1u : 0u;
    }
    if (gridDim.y > 1) {
#pragma unroll
        for (int e = 0; e < E; e++)
            if (live[e]) atomicAdd(reinterpret_cast<unsigned long long *>(&counts[base + 32u * e + lane - row_lo + 1]), (unsigned long long)cnt[e]);
        return;
    }
#pragma unroll
    for (int e = 0; e < E; e++)
        if (live[e]) counts[base + 32u * e + lane - row_lo + 1] = cnt[e];
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
}

// The tile's kept entries, taken in tile order (row-major = output order), fill the contiguous output
// range that starts at indptr[first row of the tile].  Each warp takes one contiguous share of the
// tile's entries: it counts its kept entries (ballot + popc), the 8 counts are prefixed through shared
// memory behind ONE barrier, and the warp then stores its kept entries in order -- consecutive lanes
// write consecutive output entries.
template <int E, int GW>
__global__ void __launch_bounds__(32 * GW)
fill_compact_kernel(PlanDev p, uint32_t G, uint64_t tile_row0, uint64_t row_lo, uint64_t row_hi, double tol,
                    const uint64_t *__restrict__ indptr, uint64_t *__restrict__ indices, double2 *__restrict__ data)
{
    constexpr uint32_t R = 32u * E;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sdat = reinterpret_cast<double2 *>(smem_raw);                         // R*G * 16 B
    uint64_t *sidx = reinterpret_cast<uint64_t *>(smem_raw + (size_t)R * G * 16u);  // R*G *  8 B
    __shared__ uint32_t s_wcnt[GW];
    const uint32_t lane = threadIdx.x & 31u, gw = threadIdx.x >> 5;
    const uint64_t tile_base = tile_row0 + (uint64_t)blockIdx.x * R;               // multiple of R
    const uint32_t tbase = (uint32_t)tile_base;
    uint32_t r[E];
#pragma unroll
    for (int e = 0; e < E; e++) r[e] = tbase + 32u * e + lane;
    for (uint32_t g = gw; g < G; g += GW) {                                         // as fill_staged_kernel
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
            const uint32_t o = (32u * e + lane) * G + slot;
            sidx[o] = (uint64_t)(r[e] ^ d.x);
            sdat[o] = make_double2(ar[e], ai[e]);
        }
    }
    __syncthreads();
    // the tile's rows inside the request (ragged first / last tile), as a range of tile entries
    const uint64_t first = tile_base < row_lo ? row_lo : tile_base;
    const uint64_t last = tile_base + R < row_hi ? tile_base + R : row_hi;             // exclusive
    const uint32_t e_lo = (uint32_t)(first - tile_base) * G, e_hi = (uint32_t)(last - tile_base) * G;
    const uint32_t per = ((e_hi - e_lo + GW - 1) / GW + 31u) & ~31u;                  // entries per warp, whole chunks
    const uint32_t w_lo = min(e_hi, e_lo + gw * per), w_hi = min(e_hi, w_lo + per);
    uint32_t mine = 0;
    for (uint32_t e0 = w_lo; e0 < w_hi; e0 += 32u) {
        const uint32_t e = e0 + lane;
        const bool keep = e < w_hi && keep_entry(sdat[e < w_hi ? e : w_lo], tol);
        mine += __popc(__ballot_sync(FULL_MASK, keep));
    }
    if (lane == 0) s_wcnt[gw] = mine;
    __syncthreads();
    uint64_t out = indptr[first - row_lo];
#pragma unroll
    for (int w = 0; w < GW; w++) out += w < (int)gw ? s_wcnt[w] : 0u;
    for (uint32_t e0 = w_lo; e0 < w_hi; e0 += 32u) {
        const uint32_t e = e0 + lane;
        double2 d = make_double2(0.0, 0.0);
        if (e < w_hi) d = sdat[e];
        const bool keep = e < w_hi && keep_entry(d, tol);
        const unsigned mask = __ballot_sync(FULL_MASK, keep);
        if (keep) {
            const uint64_t pos = out + __popc(mask & ((1u << lane) - 1u));
            data[pos] = d;
            indices[pos] = sidx[e];
        }
        out += __popc(mask);
    }
}

// ---- any CSR (rows of any length: a matrix wrapped by SpMat::new_unchecked, read back from disk, or already compacted) --
// util::csmatrix_nz / csmatrix_eliminate_zeroes (util.rs:144-171) take any CsMat.  Warp per row, driven by the stored
// indptr (whose first entry may be a global offset): same keep rule, same order.
__global__ void __launch_bounds__(K2_THREADS)
csr_count_kept_kernel(uint64_t n_rows, const uint64_t *__restrict__ indptr_in, const double2 *__restrict__ data, double tol,
                      uint64_t *__restrict__ counts)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (K2_THREADS / 32), base = indptr_in[0];
    for (uint64_t row = (uint64_t)blockIdx.x * (K2_THREADS / 32) + (threadIdx.x >> 5); row < n_rows; row += warps) {
        const uint64_t k1 = indptr_in[row + 1] - base;
        uint32_t n = 0;
        for (uint64_t k = indptr_in[row] - base + lane; k < k1; k += 32u) n += keep_entry(data[k], tol) ? 1u : 0u;
        n = __reduce_add_sync(FULL_MASK, n);
        if (lane == 0) counts[row + 1] = n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
}

__global__ void __launch_bounds__(K2_THREADS)
csr_compact_rows_kernel(uint64_t n_rows, const uint64_t *__restrict__ indptr_in, const uint64_t *__restrict__ indices_in,
                        const double2 *__restrict__ data_in, double tol, const uint64_t *__restrict__ indptr,
                        uint64_t *__restrict__ indices_out, double2 *__restrict__ data_out)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (K2_THREADS / 32), base = indptr_in[0];
    for (uint64_t row = (uint64_t)blockIdx.x * (K2_THREADS / 32) + (threadIdx.x >> 5); row < n_rows; row += warps) {
        uint64_t out = indptr[row];
        const uint64_t k1 = indptr_in[row + 1] - base;
        for (uint64_t k0 = indptr_in[row] - base; k0 < k1; k0 += 32u) {
            const uint64_t k = k0 + lane;
            double2 d = make_double2(0.0, 0.0);
            if (k < k1) d = data_in[k];
            const bool keep = k < k1 && keep_entry(d, tol);
            const unsigned mask = __ballot_sync(FULL_MASK, keep);
            if (keep) {
                const uint64_t pos = out + __popc(mask & ((1u << lane) - 1u));
                data_out[pos] = d;
                indices_out[pos] = indices_in[k];
            }
            out += __popc(mask);
        }
    }
}

}  // namespace qr
