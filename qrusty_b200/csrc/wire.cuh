// wire.cuh -- the CSR shard as it crosses PCIe (qr_build_host).
//
// The reference hands its caller indptr u64[rows+1], indices u64[nnz], data complex128[nnz] in host memory
// (UnsafeVectors, qrusty/src/accel.rs:15-20; moved into numpy by pyqrusty/src/lib.rs:190-214): 24 + 8/G bytes per entry,
// and the export is bound by the PCIe link (56 GB/s), not by the build (6.4 TB/s).  Two of the three arrays are
// redundant on the wire:
//   indptr   is affine, r * G -- written by the host;
//   indices  are r ^ gx[g] for one of the G masks: the device sends the group id g of every stored entry (one byte while
//            G <= 256, two while G <= 65536, else the 32-bit column) and the host threads that would otherwise idle
//            during the DMA rebuild the 64-bit column while the next window is in flight.
// data crosses as it is, straight into the caller's array when that is page-locked.  17-18 bytes per entry instead of 24.
#pragma once
#include "plan.cuh"

namespace qr {

// group id of the entry whose column is `col` in row `row`: the masks are ascending, so a binary search
__device__ __forceinline__ uint32_t group_of(const uint32_t *__restrict__ gx, uint32_t G, uint32_t x)
{
    uint32_t lo = 0, hi = G;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&gx[mid]) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// T = uint8_t / uint16_t: group ids; T = uint32_t: the column itself.  entries [0, n) of a window whose first row is row0.
template <typename T>
__global__ void __launch_bounds__(256)
wire_columns_kernel(PlanDev p, uint32_t G, uint64_t row0, uint64_t n, const uint64_t *__restrict__ indices, T *__restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const uint32_t col = (uint32_t)indices[i];
        if (sizeof(T) == 4) { out[i] = (T)col; continue; }
        const uint32_t row = (uint32_t)(row0 + i / G);
        out[i] = (T)group_of(p.gx, G, col ^ row);
    }
}

}  // namespace qr
