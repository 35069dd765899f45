// scan.cuh -- the one warp-level scan every prefix sum in this library is built on
// (K1: head-flag scan -> group ids, digit-bin scan of the radix sort; K2: per-row
// nnz counts -> indptr in drop-zeros mode).
#pragma once
#include <cstdint>

namespace qr {

constexpr unsigned FULL_MASK = 0xffffffffu;

// Inclusive scan of one value per lane (Kogge-Stone over shuffles, 5 steps).
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v)
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T up = __shfl_up_sync(FULL_MASK, v, d);
        if (lane >= (unsigned)d) v += up;
    }
    return v;
}

// Exclusive scan across a whole CTA (blockDim.x a multiple of 32, <= 1024).
// `warp_sums` is shared scratch of >= 33 entries.  Returns the exclusive prefix
// of v; *total receives the CTA-wide sum.  Contains three __syncthreads(); the
// last one makes back-to-back calls on the same scratch safe.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *warp_sums, T *total)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    T incl = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T s = lane < n_warps ? warp_sums[lane] : T(0);
        T si = warp_inclusive_scan(s);
        warp_sums[lane] = si - s;               // exclusive prefix of each warp
        if (lane == 31) warp_sums[32] = si;     // grand total
    }
    __syncthreads();
    T base = warp_sums[warp];
    *total = warp_sums[32];
    __syncthreads();
    return base + incl - v;
}

}  // namespace qr
