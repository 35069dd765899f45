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
#include "scan.cuh"
#include <cuda_runtime.h>

namespace qr {

constexpr int K2_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                                // per thread
constexpr int SCAN_TILE = K2_THREADS * SCAN_ITEMS;

__device__ __forceinline__ bool keep_entry(double2 d, double tol) { return hypot(d.x, d.y) > tol; }

// counts[r + 1] = #{entries of row r with norm > tol}; counts[0] = 0.  CTA b owns rows [b*R, (b+1)*R).
__global__ void __launch_bounds__(K2_THREADS)
count_kept_kernel(uint64_t n_rows, uint32_t G, uint32_t R, const double2 *__restrict__ data, double tol,
                  uint64_t *__restrict__ counts)
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
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
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

}  // namespace qr
