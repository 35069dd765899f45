// apply_tile.cuh -- K4 tiled: matrix-free H.v with the FAR groups served from shared memory.
//
// The gather kernel (apply.cuh) leaves all re-use of v to L1/L2.  That works for masks whose partners
// v[r ^ x] lie inside the window L2 can hold; partners along the top row bits miss (ncu, C4 n = 25:
// v read 6.6 times) and partners on another GPU cross NVLink one 16-byte load at a time.  Here those
// FAR groups are served from shared memory instead:
//
//   * the row space is cut at bit d: a row is (slice = bits >= d, offset = bits < d); a tile is one
//     run of R = 2^log2_run consecutive offsets in EVERY slice the CTA needs;
//   * a producer warp pulls each needed run -- R*16 contiguous bytes, local HBM or a peer's shard over
//     NVLink -- with cp.async.bulk.shared::cluster.global (TMA, SASS UBLKCP) into a ring of stages and
//     signals an mbarrier with the byte count; the consumers never issue a load for FAR data;
//   * a FAR group with mask x maps row (slice s, run t, offset i) to chunk (s ^ x_hi, t ^ x_mid) at
//     offset i ^ x_lo: one shared-memory read, conflict-free (a warp reads a permuted 512-byte line).
//
// Two uses (host side: make_tile_plan, qrusty_cuda.cu):
//   NEAR = false  second pass of the single-GPU apply: the row slots are all 2^(m-d) slices of the
//                 local block (closed under the FAR masks), every slot is loaded once per tile and
//                 y += sum_far.  Traffic 48 B/row whatever the number of FAR groups, against 16 B/row
//                 PER far group for the gather.
//   NEAR = true   fused distributed apply (qr_apply_p2p): one row slot (the rank's shard), the chunks
//                 are runs of the PEERS' shards, and the same CTA gathers the NEAR (local) groups from
//                 L1/L2 while its bulk loads are in flight over NVLink.  y is written once.
//                 Ranks synchronise through flags in IPC-mapped memory (st.release.sys / ld.acquire.sys):
//                 "my shard is complete" before the first remote load, "I no longer read yours" after
//                 the last -- no NCCL call on the path.
#pragma once
#include "apply.cuh"

namespace qr {

constexpr int TILE_THREADS = 512;
constexpr int TILE_MAX_LOAD = 128;        // chunks per tile
constexpr int TILE_MAX_ROWSLOTS = 128;    // row slices per tile
constexpr int TILE_MAX_PEERS = APPLY_MAX_PEERS;
constexpr int TILE_MAX_FAR = 64;          // groups served from shared memory (descriptors in shared memory)
constexpr int TILE_MAX_NEAR = 256;        // groups gathered by the same kernel (descriptors in shared memory)

struct ApplyTileArgs {
    const double2 *base[TILE_MAX_LOAD];   // chunk j lives in the slice starting here (element 0 = offset 0)
    uint32_t xi[TILE_MAX_LOAD];           // ... at run (t ^ xi[j])
    uint32_t row0[TILE_MAX_ROWSLOTS];     // global row id of offset 0 of row slot s
    const double2 *peer[TILE_MAX_PEERS];  // NEAR: v of block q, pre-offset so that it is indexed by the GLOBAL row id
                                          // (n_peers <= 1: peer[0] is the whole vector)
    uint32_t n_load, n_row_slots, log2_run, n_tiles, stages;
    uint32_t n_far, n_near, block_bits, my_block, accumulate;
    uint32_t own_loaded;                  // row slot s is chunk s of the tile (the mask-0 group reads v[r] there)
    const uint32_t *far_groups;           // [n_far] group ids
    const uint8_t *far_part;              // [n_far][n_row_slots]: chunk serving row slot s
    const uint32_t *near_groups;          // [n_near] group ids
    // cross-GPU flags (n_peers > 1): flags_local = {ready[P], done[P]} in this rank's memory,
    // peer_flags[q] = the same array of rank q (IPC-mapped)
    uint64_t *flags_local;
    uint64_t *peer_flags[TILE_MAX_PEERS];
    uint32_t *cta_counter;
    uint64_t epoch;
    uint32_t n_peers, need_mask;          // ranks whose shards this rank reads
};

// smem_u32, mbar_init, mbar_expect_tx, mbar_wait: fill.cuh (shared with the rows kernel)
// global (local HBM or a peer's memory behind NVLink) -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load_global_to_smem(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// E = rows per thread per chunk of a tile; the host picks it so that a tile is a whole number of chunks
template <bool NEAR, int E>
__global__ void __launch_bounds__(TILE_THREADS, 1)
apply_tile_kernel(PlanDev p, const __grid_constant__ ApplyTileArgs a, uint64_t row_lo, double2 *__restrict__ y,
                  const double2 *__restrict__ diag, const double *__restrict__ diag_re)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t R = 1u << a.log2_run, rmask = R - 1u;
    const uint32_t stage_elems = a.n_load * R;
    double2 *ring = reinterpret_cast<double2 *>(smem_raw);                                 // [stages][n_load][R]
    unsigned char *q = smem_raw + (size_t)a.stages * stage_elems * 16u;
    uint64_t *full = reinterpret_cast<uint64_t *>(q); q += 64;                             // one mbarrier per stage
    GroupDesc *sfar = reinterpret_cast<GroupDesc *>(q); q += (size_t)a.n_far * sizeof(GroupDesc);
    GroupDesc *snear = reinterpret_cast<GroupDesc *>(q); q += (size_t)(NEAR ? a.n_near : 0u) * sizeof(GroupDesc);
    const double2 **snearv = reinterpret_cast<const double2 **>(q); q += (size_t)(NEAR ? a.n_near : 0u) * 8u;
    uint8_t *spart = q;                                                                    // [n_far][n_row_slots]

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) {
        for (uint32_t s = 0; s < a.stages; s++) mbar_init(&full[s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t k = tid; k < a.n_far; k += TILE_THREADS) sfar[k] = p.gdesc[__ldg(&a.far_groups[k])];
    for (uint32_t k = tid; k < a.n_far * a.n_row_slots; k += TILE_THREADS) spart[k] = __ldg(&a.far_part[k]);
    if (NEAR) {
        for (uint32_t k = tid; k < a.n_near; k += TILE_THREADS) {
            const GroupDesc d = p.gdesc[__ldg(&a.near_groups[k])];
            snear[k] = d;
            snearv[k] = a.n_peers > 1u ? a.peer[a.my_block ^ (d.x >> a.block_bits)] : a.peer[0];
        }
    }
    if (a.n_peers > 1u) {
        // "my shard is complete" (stream order: whatever produced it has finished) -> every peer ...
        if (blockIdx.x == 0 && tid < a.n_peers && tid != a.my_block) st_release_sys(a.peer_flags[tid] + a.my_block, a.epoch);
        // ... and nobody here touches a peer's shard before that peer has said the same
        if (tid < a.n_peers && ((a.need_mask >> tid) & 1u)) wait_flag(a.flags_local + tid, a.epoch);
    }
    __syncthreads();

    auto issue = [&](uint32_t tau, uint32_t stage) {          // warp 0: one bulk copy per chunk
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&full[stage], stage_elems * 16u);
            __syncwarp();
            for (uint32_t j = lane; j < a.n_load; j += 32u)
                bulk_load_global_to_smem(ring + (size_t)stage * stage_elems + (size_t)j * R,
                                         a.base[j] + ((uint64_t)(tau ^ a.xi[j]) << a.log2_run), R * 16u, &full[stage]);
        }
    };

    const uint32_t tile_rows = a.n_row_slots << a.log2_run;
    for (uint32_t s = 0; s + 1u < a.stages; s++) {
        const uint64_t tau = (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x;
        if (tau < a.n_tiles) issue((uint32_t)tau, s);
    }
    uint32_t it = 0;
    for (uint32_t tau = blockIdx.x; tau < a.n_tiles; tau += gridDim.x, it++) {
        const uint32_t stage = it % a.stages, parity = (it / a.stages) & 1u;
        {   // refill the stage the previous iteration released (its trailing __syncthreads)
            const uint64_t tau_n = (uint64_t)tau + (uint64_t)(a.stages - 1u) * gridDim.x;
            if (tau_n < a.n_tiles) issue((uint32_t)tau_n, (it + a.stages - 1u) % a.stages);
        }
        const double2 *st = ring + (size_t)stage * stage_elems;
        bool waited = false;
        for (uint32_t c0 = 0; c0 < tile_rows; c0 += TILE_THREADS * E) {
            uint32_t r[E], sig[E], off[E];
            double yr[E], yi[E];
#pragma unroll
            for (int e = 0; e < E; e++) {
                const uint32_t idx = c0 + (uint32_t)e * TILE_THREADS + tid;  // tile_rows is a multiple of TILE_THREADS * E
                sig[e] = idx >> a.log2_run; off[e] = idx & rmask;
                r[e] = a.row0[sig[e]] + (tau << a.log2_run) + off[e];
                yr[e] = 0.0; yi[e] = 0.0;
            }
            if (NEAR) {
                // the groups this kernel gathers: issued while the tile's bulk loads are still in flight
                for (uint32_t k = 0; k < a.n_near; k++) {
                    const GroupDesc d = snear[k];
                    const double2 *vb = snearv[k];
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
            } else if (a.accumulate) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 o = __ldcs(&y[(uint64_t)r[e] - row_lo]);
                    yr[e] = o.x; yi[e] = o.y;
                }
            }
            if (!waited) { mbar_wait(&full[stage], parity); waited = true; }
            if (NEAR && (diag_re != nullptr || diag != nullptr)) {
                // cached diag(H) (the mask-0 group is in neither list then): v[r] from the tile when the row slots are loaded
                const double2 *v_own = a.n_peers > 1u ? a.peer[a.my_block] : a.peer[0];
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 w = a.own_loaded ? st[(sig[e] << a.log2_run) + off[e]] : ld_nc_double2(&v_own[r[e]]);
                    if (diag_re != nullptr) cfma(yr[e], yi[e], __ldcs(&diag_re[(uint64_t)r[e] - row_lo]), 0.0, w, true);
                    else { const double2 dg = __ldcs(&diag[(uint64_t)r[e] - row_lo]); cfma(yr[e], yi[e], dg.x, dg.y, w, false); }
                }
            }
            for (uint32_t k = 0; k < a.n_far; k++) {
                const GroupDesc d = sfar[k];
                const uint8_t *part = spart + k * a.n_row_slots;
                const uint32_t xl = d.x & rmask;
                const bool real = (d.flag & 2u) != 0u;
                if (d.flag & 1u) {
#pragma unroll
                    for (int e = 0; e < E; e++)
                        cfma(yr[e], yi[e], d.cre, d.cim, st[((uint32_t)part[sig[e]] << a.log2_run) + (off[e] ^ xl)], real);
                } else {
                    double ar[E], ai[E];
                    group_values<E>(p, d.t0, d.t1, r, ar, ai);
#pragma unroll
                    for (int e = 0; e < E; e++)
                        cfma(yr[e], yi[e], ar[e], ai[e], st[((uint32_t)part[sig[e]] << a.log2_run) + (off[e] ^ xl)], real);
                }
            }
#pragma unroll
            for (int e = 0; e < E; e++)
                __stcs(&y[(uint64_t)r[e] - row_lo], make_double2(yr[e], yi[e]));
        }
        if (!waited) mbar_wait(&full[stage], parity);          // tile_rows == 0 never happens; keeps the phases in step
        __syncthreads();                                       // every read of this stage is done: it may be refilled
    }

    if (a.n_peers > 1u) {
        // last CTA out tells every peer that this rank no longer reads their shards
        __shared__ uint32_t s_last;
        if (tid == 0) {
            __threadfence();
            s_last = atomicAdd(a.cta_counter, 1u) == gridDim.x - 1u;
        }
        __syncthreads();
        if (s_last) {
            if (tid == 0) *a.cta_counter = 0u;
            if (tid < a.n_peers && tid != a.my_block) st_release_sys(a.peer_flags[tid] + a.n_peers + a.my_block, a.epoch);
        }
    }
}

// ---------------------------------------------------------------------------------
// One contiguous run per tile (n_row_slots == 1): the local apply and the fused distributed apply.
// The gathers of the NEAR groups are bound by L2 latency and throughput, so the SM must hold many warps in DIFFERENT phases
// (gather / wait for the TMA / shared-memory groups): CTAs of 256 or 512 threads, 1024 threads per SM at 64 registers, one
// tile per CTA (stages = 1, grid = tiles) or a few per CTA through the ring.  Rows r0 + e * THREADS per thread.  Chunk 0 is the CTA's own run; the mask-0 group
// (cached diag) and every group whose mask lies inside the run read it; further chunks are runs of other blocks
// (peers' shards over NVLink).  sfoff[k] = first element of group k's chunk inside a stage.
// ---------------------------------------------------------------------------------
constexpr int RUN_E = 4;

template <int RUN_THREADS, int E, int MINB>
__global__ void __launch_bounds__(RUN_THREADS, MINB)
apply_run_kernel(PlanDev p, const __grid_constant__ ApplyTileArgs a, uint64_t row_lo, double2 *__restrict__ y,
                 const double2 *__restrict__ diag, const double *__restrict__ diag_re)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t R = 1u << a.log2_run, rmask = R - 1u;
    const uint32_t stage_elems = a.n_load * R;
    double2 *ring = reinterpret_cast<double2 *>(smem_raw);                                 // [stages][n_load][R]
    unsigned char *q = smem_raw + (size_t)a.stages * stage_elems * 16u;
    uint64_t *full = reinterpret_cast<uint64_t *>(q); q += 64;
    GroupDesc *sfar = reinterpret_cast<GroupDesc *>(q); q += (size_t)a.n_far * sizeof(GroupDesc);
    GroupDesc *snear = reinterpret_cast<GroupDesc *>(q); q += (size_t)a.n_near * sizeof(GroupDesc);
    const double2 **snearv = reinterpret_cast<const double2 **>(q); q += (size_t)a.n_near * 8u;
    uint32_t *sfoff = reinterpret_cast<uint32_t *>(q);                                     // [n_far]

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) {
        for (uint32_t s = 0; s < a.stages; s++) mbar_init(&full[s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t k = tid; k < a.n_far; k += RUN_THREADS) {
        GroupDesc d = p.gdesc[__ldg(&a.far_groups[k])];
        d.x &= rmask;                                                                      // offset inside the chunk
        sfar[k] = d;
        sfoff[k] = (uint32_t)__ldg(&a.far_part[k]) << a.log2_run;
    }
    for (uint32_t k = tid; k < a.n_near; k += RUN_THREADS) {
        const GroupDesc d = p.gdesc[__ldg(&a.near_groups[k])];
        snear[k] = d;
        snearv[k] = a.n_peers > 1u ? a.peer[a.my_block ^ (d.x >> a.block_bits)] : a.peer[0];
    }
    if (a.n_peers > 1u) {
        if (blockIdx.x == 0 && tid < a.n_peers && tid != a.my_block) st_release_sys(a.peer_flags[tid] + a.my_block, a.epoch);
        if (tid < a.n_peers && ((a.need_mask >> tid) & 1u)) wait_flag(a.flags_local + tid, a.epoch);
    }
    __syncthreads();

    auto issue = [&](uint32_t tau, uint32_t stage) {
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&full[stage], stage_elems * 16u);
            __syncwarp();
            for (uint32_t j = lane; j < a.n_load; j += 32u)
                bulk_load_global_to_smem(ring + (size_t)stage * stage_elems + (size_t)j * R,
                                         a.base[j] + ((uint64_t)(tau ^ a.xi[j]) << a.log2_run), R * 16u, &full[stage]);
        }
    };

    for (uint32_t s = 0; s + 1u < a.stages; s++) {
        const uint64_t tau = (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x;
        if (tau < a.n_tiles) issue((uint32_t)tau, s);
    }
    const bool has_diag = diag_re != nullptr || diag != nullptr;
    const double2 *v_own = a.n_peers > 1u ? a.peer[a.my_block] : a.peer[0];
    uint32_t it = 0;
    for (uint32_t tau = blockIdx.x; tau < a.n_tiles; tau += gridDim.x, it++) {
        const uint32_t stage = it % a.stages, parity = (it / a.stages) & 1u;
        {
            const uint64_t tau_n = (uint64_t)tau + (uint64_t)(a.stages - 1u) * gridDim.x;
            if (tau_n < a.n_tiles) issue((uint32_t)tau_n, (it + a.stages - 1u) % a.stages);
        }
        const double2 *st = ring + (size_t)stage * stage_elems;
        bool waited = false;
        for (uint32_t c0 = 0; c0 < R; c0 += RUN_THREADS * E) {                   // R is a multiple of RUN_THREADS * E
            const uint32_t i0 = c0 + tid;
            uint32_t r[E];
            double yr[E], yi[E];
#pragma unroll
            for (int e = 0; e < E; e++) { r[e] = a.row0[0] + (tau << a.log2_run) + i0 + (uint32_t)e * RUN_THREADS; yr[e] = 0.0; yi[e] = 0.0; }
            // Summation order when the own run is not a chunk (fused distributed apply): cached diagonal, local groups by
            // ascending mask, remote groups by ascending mask -- every remote mask is above every local one, so this is the
            // gather kernel's order and the result equals the all-gather form bit for bit.
            if (has_diag && !a.own_loaded) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 w = ld_nc_double2(&v_own[r[e]]);
                    if (diag_re != nullptr) cfma(yr[e], yi[e], __ldcs(&diag_re[(uint64_t)r[e] - row_lo]), 0.0, w, true);
                    else { const double2 dg = __ldcs(&diag[(uint64_t)r[e] - row_lo]); cfma(yr[e], yi[e], dg.x, dg.y, w, false); }
                }
            }
            // the groups gathered through L1/L2 (peers' memory for the few remote groups that found no chunk): in flight
            // while the tile's bulk loads land
#pragma unroll 1
            for (uint32_t k = 0; k < a.n_near; k++) {
                const GroupDesc d = snear[k];
                const double2 *vb = snearv[k];
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
            if (!waited) { mbar_wait(&full[stage], parity); waited = true; }
            if (has_diag && a.own_loaded) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 w = st[i0 + (uint32_t)e * RUN_THREADS];
                    if (diag_re != nullptr) cfma(yr[e], yi[e], __ldcs(&diag_re[(uint64_t)r[e] - row_lo]), 0.0, w, true);
                    else { const double2 dg = __ldcs(&diag[(uint64_t)r[e] - row_lo]); cfma(yr[e], yi[e], dg.x, dg.y, w, false); }
                }
            }
#pragma unroll 1
            for (uint32_t k = 0; k < a.n_far; k++) {
                const GroupDesc d = sfar[k];
                const double2 *ch = st + sfoff[k];
                const bool real = (d.flag & 2u) != 0u;
                if (d.flag & 1u) {
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], d.cre, d.cim, ch[(i0 + (uint32_t)e * RUN_THREADS) ^ d.x], real);
                } else {
                    double ar[E], ai[E];
                    group_values<E>(p, d.t0, d.t1, r, ar, ai);
#pragma unroll
                    for (int e = 0; e < E; e++) cfma(yr[e], yi[e], ar[e], ai[e], ch[(i0 + (uint32_t)e * RUN_THREADS) ^ d.x], real);
                }
            }
#pragma unroll
            for (int e = 0; e < E; e++) __stcs(&y[(uint64_t)r[e] - row_lo], make_double2(yr[e], yi[e]));
        }
        if (!waited) mbar_wait(&full[stage], parity);
        __syncthreads();
    }

    if (a.n_peers > 1u) {
        __shared__ uint32_t s_last;
        if (tid == 0) {
            __threadfence();
            s_last = atomicAdd(a.cta_counter, 1u) == gridDim.x - 1u;
        }
        __syncthreads();
        if (s_last) {
            if (tid == 0) *a.cta_counter = 0u;
            if (tid < a.n_peers && tid != a.my_block) st_release_sys(a.peer_flags[tid] + a.n_peers + a.my_block, a.epoch);
        }
    }
}

// after the fused apply: wait until every peer has finished reading this rank's shard (done[q] >= epoch),
// so that whatever the stream runs next may overwrite it
__global__ void p2p_wait_done_kernel(const uint64_t *flags_local, uint32_t n_peers, uint32_t my_rank, uint32_t readers_mask, uint64_t epoch)
{
    const uint32_t q = threadIdx.x;
    if (q < n_peers && q != my_rank && ((readers_mask >> q) & 1u)) wait_flag(flags_local + n_peers + q, epoch);
}

}  // namespace qr
