// canonicalise.cuh -- K1: term canonicalisation on the GPU.
//
// Replaces what rowwise::make_row redoes for EVERY row with a stable sort of all
// T terms (qrusty/src/accel.rs:174-205): group the terms by X-mask once, keeping
// the original term order inside a group (that order is the reference's
// left-to-right summation order, accel.rs:191-205), and precompute the tables the
// fill kernel needs to place a group in a row without sorting (plan.cuh).
//
// One CTA of 512 threads; T is at most tens of thousands (a Hamiltonian's term
// list), the output of this kernel is what 2^n rows then share.
//   1. stable LSD radix sort of (x, original index), 8-bit digits, ceil(n/8) passes
//      (preceded by a z-keyed sort when duplicate merging is requested):
//      histogram (shared atomics) -> bin scan -> per-tile stable ranking with
//      __match_any_sync + per-warp digit counts -> scatter.
//   2. head flags + CTA scan (scan.cuh) -> group ids, gx[], goff[], G.
//   2b. (opt-in) runs of identical (x, z) merged by a segmented reduce.
//   3. rank tables cnt[g][b] by two binary searches on the sorted masks, lr5[g][j].
#pragma once
#include "plan.cuh"
#include "scan.cuh"

namespace qr {

constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t key)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t *a, uint32_t n, uint32_t key)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] <= key) lo = mid + 1; else hi = mid; }
    return lo;
}

// One stable LSD radix sort of (key, payload) over `bits` key bits; returns with the sorted
// arrays in (kin, iin) (the pointers are swapped per pass).  All threads of the CTA call it.
struct RadixSmem {
    uint32_t hist[256];                  // digit histogram, then running bin base
    uint16_t wcount[K1_WARPS][256];      // per-warp digit counts of the current tile
    uint32_t woffset[K1_WARPS][256];     // per-warp digit start positions
    uint32_t scan_scratch[33];
};

__device__ __forceinline__ void radix_sort_cta(RadixSmem &sm, uint32_t T, int bits, uint32_t *&kin, uint32_t *&kout,
                                               uint32_t *&iin, uint32_t *&iout)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (int shift = 0; shift < bits; shift += 8) {
        if (tid < 256) sm.hist[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < T; i += K1_THREADS) atomicAdd(&sm.hist[(kin[i] >> shift) & 255u], 1u);
        __syncthreads();
        uint32_t total;
        const uint32_t h = tid < 256 ? sm.hist[tid] : 0u;
        const uint32_t excl = block_exclusive_scan(h, sm.scan_scratch, &total);
        if (tid < 256) sm.hist[tid] = excl;         // hist[d] = next free output position of digit d
        __syncthreads();
        for (uint32_t tile = 0; tile < T; tile += K1_THREADS) {
            const uint32_t i = tile + tid;
            const bool valid = i < T;
            uint32_t key = 0, idx = 0, d = 0xffffffffu;     // invalid lanes form their own match group
            if (valid) { key = kin[i]; idx = iin[i]; d = (key >> shift) & 255u; }
            const unsigned peers = __match_any_sync(FULL_MASK, d);
            const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank_in_warp == 0) sm.wcount[warp][d] = (uint16_t)__popc(peers);
            __syncthreads();
            if (tid < 256) {                        // digit tid: prefix over warps, in element order
                uint32_t run = sm.hist[tid];
#pragma unroll 4
                for (int w = 0; w < K1_WARPS; w++) {
                    const uint32_t c = sm.wcount[w][tid];
                    sm.wcount[w][tid] = 0;
                    sm.woffset[w][tid] = run;
                    run += c;
                }
                sm.hist[tid] = run;
            }
            __syncthreads();
            if (valid) { const uint32_t pos = sm.woffset[warp][d] + rank_in_warp; kout[pos] = key; iout[pos] = idx; }
            __syncthreads();
        }
        uint32_t *t = kin; kin = kout; kout = t;
        t = iin; iin = iout; iout = t;
    }
}

// merge_dups = 0: terms keep their original order inside a group (bit-exact data).
// merge_dups = 1: terms are ordered by (x, z) -- a z-keyed sort followed by the stable x-keyed
// sort -- and every run of identical (x, z) is replaced by one term whose coefficient is the
// run's sum in original order (segmented reduce).  This changes the summation order inside a
// group, so data then agrees with the reference to rounding (1e-12), not bit for bit.
__global__ void __launch_bounds__(K1_THREADS, 1) canonicalise_kernel(PlanDev p, uint32_t merge_dups)
{
    __shared__ RadixSmem sm;
    __shared__ uint32_t max_group, n_const;
    // small operators (the usual case: tens of terms) sort entirely in shared memory
    constexpr uint32_t SMALL_T = 1024;
    constexpr uint32_t RANK_T = 512;                 // up to here one rank-sort pass beats the radix passes
    __shared__ uint32_t s_sort[4][SMALL_T];

    const uint32_t tid = threadIdx.x;
    const uint32_t T = p.n_terms;
    // programmatic dependent launch: this kernel rewrites the tables the previous fill may still be reading, so it
    // waits for its predecessor first; the fill that follows may be launched right away (it waits in turn)
    pdl_wait();
    pdl_launch_dependents();

    // ---- 0. payload = original index ------------------------------------------------------
    for (uint32_t i = tid; i < K1_WARPS * 256; i += K1_THREADS) (&sm.wcount[0][0])[i] = 0;
    if (tid == 0) { max_group = 0; n_const = 0; }
    __syncthreads();

    // ---- 1. stable sort of (x, original index) -- with merge_dups: of ((x, z), original index) --------
    uint32_t *kin = p.key_a, *kout = p.key_b, *iin = p.idx_a, *iout = p.idx_b;
    if (T <= SMALL_T) { kin = s_sort[0]; kout = s_sort[1]; iin = s_sort[2]; iout = s_sort[3]; }
    if (T <= RANK_T) {
        // short term lists (the usual Hamiltonian: tens of terms): one rank-sort pass in shared memory.
        // rank(i) = #{j : key_j < key_i or (key_j == key_i and j < i)} -- stable by construction; every
        // thread scans the same j, so the shared-memory reads are broadcasts.
        for (uint32_t i = tid; i < T; i += K1_THREADS) { kout[i] = (uint32_t)p.raw[i].x; iout[i] = merge_dups ? (uint32_t)p.raw[i].z : 0u; }
        __syncthreads();
        for (uint32_t i = tid; i < T; i += K1_THREADS) {
            const uint64_t ki = ((uint64_t)kout[i] << 32) | iout[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < T; j++) {
                const uint64_t kj = ((uint64_t)kout[j] << 32) | iout[j];
                rank += (kj < ki || (kj == ki && j < i)) ? 1u : 0u;
            }
            kin[rank] = kout[i]; iin[rank] = i;
        }
        __syncthreads();
    } else {
        // stable LSD radix sort(s)
        for (uint32_t i = tid; i < T; i += K1_THREADS) iin[i] = i;
        if (merge_dups) {
            for (uint32_t i = tid; i < T; i += K1_THREADS) kin[i] = (uint32_t)p.raw[i].z;
            __syncthreads();
            radix_sort_cta(sm, T, p.n_qubits, kin, kout, iin, iout);
        }
        for (uint32_t i = tid; i < T; i += K1_THREADS) kin[i] = (uint32_t)p.raw[iin[i]].x;
        __syncthreads();
        radix_sort_cta(sm, T, p.n_qubits, kin, kout, iin, iout);
    }
    // sorted (x, idx) now in (kin, iin)

    // ---- 2. head flags -> groups; gather the sorted term table ------------------------------
    // the masks also stay in shared memory (the sort's spare buffer) for the binary searches of step 3: done on the global
    // copy each probe is an L2 round trip, 10 dependent ones per table entry -- a third of this kernel's 7 us on C2
    uint32_t *sgx = (!merge_dups && T <= SMALL_T) ? kout : nullptr;
    uint32_t carry = 0;
    for (uint32_t tile = 0; tile < T; tile += K1_THREADS) {
        const uint32_t i = tile + tid;
        const bool valid = i < T;
        const uint32_t key = valid ? kin[i] : 0u;
        const uint32_t head = (valid && (i == 0 || kin[i - 1] != key)) ? 1u : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(head, sm.scan_scratch, &total);
        if (valid) {
            if (head) { p.gx[carry + excl] = key; p.goff[carry + excl] = i; if (sgx) sgx[carry + excl] = key; }
            const uint32_t src = iin[i];
            p.perm[i] = src;
            p.tz[i] = (uint32_t)p.raw[src].z;
            p.tc[i] = make_double2(p.raw[src].re, p.raw[src].im);
        }
        carry += total;
    }
    const uint32_t G = carry;
    if (tid == 0) { p.goff[G] = T; p.meta[0] = G; }
    __syncthreads();

    // ---- 2b. optional: merge runs of identical (x, z) (segmented reduce) ----------------------
    uint32_t T2 = T;
    if (merge_dups) {
        // kout/iout are free now: kout[i] = compact position of term i, iout[] = merged perm
        uint32_t c2 = 0;
        for (uint32_t tile = 0; tile < T; tile += K1_THREADS) {
            const uint32_t i = tile + tid;
            const bool valid = i < T;
            const uint32_t head = (valid && (i == 0 || kin[i - 1] != kin[i] || p.tz[i - 1] != p.tz[i])) ? 1u : 0u;
            uint32_t total;
            const uint32_t excl = block_exclusive_scan(head, sm.scan_scratch, &total);
            if (valid) kout[i] = head ? (c2 + excl) : 0xffffffffu;
            c2 += total;
        }
        T2 = c2;
        __syncthreads();
        // run heads sum their run in original order and write the compact term; double2 staging
        // goes through p.gconst (G <= T entries are not enough) -> reuse raw-order scratch in cnt
        double2 *tc2 = reinterpret_cast<double2 *>(p.cnt);         // >= T*128 B, free until step 3
        uint32_t *tz2 = p.lr5;                                      // >= T*128 B, free until step 3
        for (uint32_t i = tid; i < T; i += K1_THREADS) {
            const uint32_t pos = kout[i];
            if (pos == 0xffffffffu) continue;
            double2 c = p.tc[i];
            double re = c.x, im = c.y;
            for (uint32_t j = i + 1; j < T && kout[j] == 0xffffffffu; j++) {
                c = p.tc[j];
                re = __dadd_rn(re, c.x); im = __dadd_rn(im, c.y);
            }
            tc2[pos] = make_double2(re, im);
            tz2[pos] = p.tz[i];
            iout[pos] = p.perm[i];
        }
        __syncthreads();
        for (uint32_t g = tid; g <= G; g += K1_THREADS) {            // group offsets in the merged list
            const uint32_t o = p.goff[g];
            p.goff[g] = o < T ? kout[o] : T2;                        // a group's first term heads a run
        }
        __syncthreads();
        for (uint32_t i = tid; i < T2; i += K1_THREADS) { p.tc[i] = tc2[i]; p.tz[i] = tz2[i]; p.perm[i] = iout[i]; }
        __syncthreads();
    }

    // ---- 3. rank tables -----------------------------------------------------------------------
    const uint32_t nq = (uint32_t)p.n_qubits;
    const uint32_t *gxs = sgx ? sgx : p.gx;
    for (uint32_t q = tid; q < G * 32u; q += K1_THREADS) {
        const uint32_t g = q >> 5, b = q & 31u;
        uint32_t c = 0;
        if (b < nq) {
            const uint32_t low = (1u << b) - 1u;
            const uint32_t lo = (gxs[g] ^ (1u << b)) & ~low;      // same prefix above b, bit b flipped
            c = upper_bound_u32(gxs, G, lo | low) - lower_bound_u32(gxs, G, lo);
        }
        p.cnt[q] = c;
        p.cnt_t[b * T + g] = c;
    }
    for (uint32_t g = tid; g < G; g += K1_THREADS) {
        const uint32_t t0 = p.goff[g], t1 = p.goff[g + 1];
        atomicMax(&max_group, t1 - t0);
        // row-independent groups (all z == 0: X-only strings) and real-valued groups
        uint32_t zor = p.tz[t0];
        double2 c = p.tc[t0];
        double re = c.x, im = c.y;
        bool real = c.y == 0.0;
        for (uint32_t t = t0 + 1; t < t1; t++) {
            zor |= p.tz[t];
            c = p.tc[t];
            re = __dadd_rn(re, c.x); im = __dadd_rn(im, c.y);     // same fold as the fill kernels
            real = real && c.y == 0.0;
        }
        p.gflag[g] = (zor == 0u ? 1u : 0u) | (real ? 2u : 0u);
        if (zor == 0u) atomicAdd(&n_const, 1u);
        p.gconst[g] = make_double2(re, im);
        GroupDesc d; d.x = p.gx[g]; d.flag = p.gflag[g]; d.t0 = t0; d.t1 = t1; d.cre = re; d.cim = im;
        p.gdesc[g] = d;
        p.lt_xn[g] = make_uint2(d.x, t1 - t0);
        for (uint32_t t = 0; t < (uint32_t)LANE_TERMS; t++) {
            const bool has = t0 + t < t1;
            p.lt_z[t * T + g] = has ? p.tz[t0 + t] : 0u;
            p.lt_c[t * T + g] = has ? p.tc[t0 + t] : make_double2(-0.0, -0.0);
        }
    }
    __syncthreads();
    for (uint32_t q = tid; q < G * 32u; q += K1_THREADS) {
        const uint32_t g = q >> 5, j = q & 31u;
        uint32_t s = 0;
#pragma unroll
        for (uint32_t b = 0; b < 5; b++) s += ((j >> b) & 1u) ? p.cnt[g * 32u + b] : 0u;
        p.lr5[q] = s;
    }
    if (tid == 0) { p.meta[1] = max_group; p.meta[2] = 0; p.meta[3] = 0; p.meta[4] = n_const; p.meta[5] = T2; }
}

// K1b: cut the sorted masks into maximal trie subtrees of at most S groups (S >= 32).
// For group g, p_g = the largest prefix level p in [5,32] whose subtree
// {h : gx[h] >> p == gx[g] >> p} holds <= S groups (two binary searches per level); g heads a
// block iff it is the first group of that subtree; a CTA scan numbers the blocks.
__global__ void __launch_bounds__(K1_THREADS, 1) partition_kernel(PlanDev p, uint32_t S)
{
    __shared__ uint32_t scan_scratch[33];
    const uint32_t tid = threadIdx.x;
    const uint32_t G = p.meta[0];
    uint32_t carry = 0;
    for (uint32_t tile = 0; tile < G; tile += K1_THREADS) {
        const uint32_t g = tile + tid;
        uint32_t head = 0, level = 0;
        if (g < G) {
            const uint32_t x = p.gx[g];
            uint32_t lo_idx = 0;
            for (level = 32; level > 5; level--) {
                const uint32_t low = level == 32 ? 0xffffffffu : ((1u << level) - 1u);
                lo_idx = lower_bound_u32(p.gx, G, x & ~low);
                if (upper_bound_u32(p.gx, G, x | low) - lo_idx <= S) break;
            }
            if (level == 5) lo_idx = lower_bound_u32(p.gx, G, x & ~31u);
            head = lo_idx == g ? 1u : 0u;
        }
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(head, scan_scratch, &total);
        if (head) { p.blk_start[carry + excl] = g; p.blk_p[carry + excl] = level; }
        carry += total;
    }
    if (tid == 0) { p.blk_start[carry] = G; p.meta[2] = carry; p.meta[3] = S; }
}

}  // namespace qr
