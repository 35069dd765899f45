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

// v0: lane <-> row, walk the groups.  For 32 aligned consecutive rows the gather
// v[r ^ x] is one aligned 512-byte segment with lanes permuted, so every load is
// fully coalesced; re-use across groups is left to L1/L2.
constexpr int APPLY_THREADS = 256;

__global__ void __launch_bounds__(APPLY_THREADS)
apply_direct_kernel(PlanDev p, uint32_t G, uint64_t row_lo, uint64_t row_hi,
                    const double2 *__restrict__ v, double2 *__restrict__ y)
{
    const uint64_t r64 = row_lo + (uint64_t)blockIdx.x * APPLY_THREADS + threadIdx.x;
    if (r64 >= row_hi) return;
    const uint32_t r = (uint32_t)r64;
    double yr = 0.0, yi = 0.0;
    for (uint32_t g = 0; g < G; g++) {
        const uint32_t x = __ldg(&p.gx[g]);
        const double2 a = group_value(p.tz, p.tc, __ldg(&p.goff[g]), __ldg(&p.goff[g + 1]), r);
        const double2 w = ld_nc_double2(&v[r ^ x]);
        yr += a.x * w.x - a.y * w.y;
        yi += a.x * w.y + a.y * w.x;
    }
    y[r64 - row_lo] = make_double2(yr, yi);
}

// diag(H): only the group with X-mask 0 (gx[0], masks are ascending) touches the diagonal.
__global__ void __launch_bounds__(256)
diagonal_kernel(PlanDev p, uint64_t row_lo, uint64_t row_hi, double2 *__restrict__ diag)
{
    const uint64_t r64 = row_lo + (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r64 >= row_hi) return;
    double2 d = make_double2(0.0, 0.0);
    if (__ldg(&p.gx[0]) == 0u) d = group_value(p.tz, p.tc, 0u, __ldg(&p.goff[1]), (uint32_t)r64);
    diag[r64 - row_lo] = d;
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
