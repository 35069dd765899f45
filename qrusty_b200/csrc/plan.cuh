// plan.cuh -- device-side view of a canonicalised operator (what every kernel reads).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/qrusty_cuda.h"

namespace qr {

// HBM layout of one plan.  All tables are tiny next to the CSR (<= 300 B per
// term) and stay L2-resident during a build.
//
//   raw   qr_term[T]      the caller's terms, original order (make_params output)
//   tz    u32[T]          Z-masks, sorted by X-mask, original order inside a group
//   tc    double2[T]      coefficients c', same order
//   perm  u32[T]          original index of each sorted term
//   gx    u32[G]          distinct X-masks, ascending
//   goff  u32[G+1]        group g owns sorted terms [goff[g], goff[g+1])
//   cnt   u32[G][32]      cnt[g][b] = #{h != g : msb(gx[g]^gx[h]) == b}
//   cnt_t u32[32][T]      cnt transposed (cnt_t[b*T + g]): lane <-> group kernels read it coalesced
//   lt_xn, lt_z, lt_c     the lanes kernel's view of the groups, structure-of-arrays with stride T: mask
//                         and term count, then the first LANE_TERMS (z, c') of every group padded with
//                         (0, -0.0) -- everything a lane needs, readable coalesced in one round of loads
//   lr5   u32[G][32]      lr5[g][j] = sum_{b<5} cnt[g][b] * bit_b(j)
//   gflag u32[G]          bit0: every term of the group has z == 0 (value is row-independent)
//                         bit1: every c' of the group is real (im == +-0)
//   gconst double2[G]     the group's ordered sum of c' (its value when bit0 is set)
//   gdesc GroupDesc[G]    {x, flag, t0, t1, gconst} packed in 32 B for the H.v kernels
//   meta  u32[8]          {G, max terms in a group, B, S, #row-independent groups, #terms after merging, 0, 0}
//   blk_start u32[B+1]    large-G path: the sorted groups cut into B trie subtrees ("blocks")
//   blk_p     u32[B]      of <= S groups; block b = groups [blk_start[b], blk_start[b+1]), all
//                         sharing the mask bits >= blk_p[b] (>= 5).  A subtree's groups fill
//                         one contiguous slot range in every row (XOR keeps subtrees together).
//
// Slot of group g in row r (columns ascending, accel.rs:188):
//   slot(r,g) = sum_b cnt[g][b] * bit_b(gx[g] ^ r)
// because h precedes g in row r  <=>  (r^gx[h]) < (r^gx[g])  <=>  at the most
// significant bit where gx[h] and gx[g] differ, r^gx[g] has a 1.
struct __align__(16) GroupDesc {
    uint32_t x, flag, t0, t1;      // mask (or tile-compacted mask), gflag, term range
    double   cre, cim;             // gconst
};

struct PlanDev {
    int       n_qubits;
    uint32_t  n_terms;
    const qr_term *raw;
    uint32_t *key_a, *key_b, *idx_a, *idx_b;   // radix-sort ping-pong, u32[T]
    uint32_t *tz;
    double2  *tc;
    uint32_t *perm;
    uint32_t *gx;
    uint32_t *goff;
    uint32_t *cnt;
    uint32_t *lr5;
    uint32_t *gflag;
    double2  *gconst;
    GroupDesc *gdesc;
    uint32_t *meta;
    uint32_t *blk_start;
    uint32_t *blk_p;
    uint32_t *cnt_t;
    uint2    *lt_xn;     // lanes kernel: per group (mask, term count)
    uint32_t *lt_z;      // lanes kernel: z of the group's first LANE_TERMS terms, [t][T], padded with 0
    double2  *lt_c;      // lanes kernel: c' of those terms, [t][T], padded with -0.0
};

// Programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization): a kernel launched with the attribute
// may start while its predecessor in the stream still runs; pdl_wait() blocks until that predecessor has completed and its
// writes are visible, pdl_launch_dependents() lets the successor start early.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int LANE_TERMS = 6;   // terms a lane of the lanes kernel keeps in registers

}  // namespace qr
