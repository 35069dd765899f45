// qrusty_cuda.cu -- the extern "C" boundary (include/qrusty_cuda.h) over the sm_100a kernels.
// No torch, no CPU compute path: every compute entry point launches kernels or fails.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <chrono>
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <new>
#include <map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <functional>
#include <string>
#include <vector>

#include <cuda.h>           // CUtensorMap types only: cuTensorMapEncodeTiled is looked up through the runtime
#include <cuda_runtime.h>
#include <nccl.h>   // types only; the library is dlopen'ed (see NcclApi)

#include "../../include/qrusty_cuda.h"
#include "apply.cuh"
#include "apply_tile.cuh"
#include "apply_fold.cuh"
#include "canonicalise.cuh"
#include "compact.cuh"
#include "fill.cuh"
#include "plan.cuh"
#include "wire.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define QR_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? QR_ERR_OOM : QR_ERR_CUDA,          \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
    } while (0)

#define QR_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return fail(QR_ERR_CUDA, std::string("launch ") + name + ": " + cudaGetErrorString(e__)); \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr size_t MAX_SMEM = 227 * 1024;

// qr_build_host's device-side staging (two row windows, two streams), one set per device
constexpr int WIN_RING = 4;                                        // windows in flight: the DMA queue never runs dry while the host finishes one
struct WinScratch {
    void *buf[WIN_RING] = {nullptr}; size_t bytes = 0; cudaStream_t stream[WIN_RING] = {nullptr};
    cudaEvent_t done[WIN_RING] = {nullptr};                        // window w's copies have landed
    void *host[WIN_RING] = {nullptr}; size_t host_bytes = 0;       // pinned staging: compact columns, data bound for pageable memory
    std::mutex busy;                                               // one qr_build_host at a time per device
};
std::mutex g_win_mutex;                                            // guards the map, not the transfers
std::map<int, WinScratch> g_win;

// small device scratch of the scan and the reductions, one set per device (a host thread may drive several GPUs in turn:
// thread-local pointers would leak the previous device's block on every switch)
struct DevScratch {
    uint64_t *tiles = nullptr; uint64_t tiles_cap = 0; double2 *partials = nullptr; double2 *tmp = nullptr; int n_sm = 0;
    // row windows of the drop-zeros build when a row is too long for a shared-memory tile (qr_build_compact_count / _fill):
    // kept per device -- a cudaMalloc + cudaFree of 256 MB per call cost 6 ms of the 6.8 ms on H8
    void *cwin = nullptr; size_t cwin_bytes = 0; std::mutex cwin_busy;
};
std::mutex g_dev_mutex;
std::map<int, DevScratch> g_dev;
DevScratch *dev_scratch(int dev)
{
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DevScratch &d = g_dev[dev];
    if (d.n_sm == 0 && (cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.n_sm < 1)) { cudaGetLastError(); d.n_sm = 148; }
    return &d;
}
// QR_PLAN_TRACE: wall-clock milestones of plan creation on stderr (where the host-visible latency of a plan goes)
struct PlanTrace {
    bool on = getenv("QR_PLAN_TRACE") != nullptr;
    double t0 = now();
    static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    void mark(const char *what) { if (on) { const double t = now(); fprintf(stderr, "qr_plan_create: %-28s +%.3f ms\n", what, t - t0); t0 = t; } }
};
int current_sm_count()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    return dev_scratch(dev)->n_sm;
}

}  // namespace

// H.v pass plan for a local row block of 2^m rows (see apply.cuh)
struct ApplyPlan {
    std::vector<qr::ApplyPass> passes;
    void *slab = nullptr;
};

// Tiled H.v plan for one aligned block of 2^m rows cut at bit d (apply_tile.cuh): which groups are FAR
// (served from shared memory), the chunks a tile loads, the NEAR groups left to the gather
struct TilePlan {
    bool ok = false;
    uint32_t m = 0, d = 0, log2_run = 0, n_row_slots = 0, stages = 0, n_far = 0, n_near = 0, e = 4;
    uint32_t need_mask = 0;                     // other blocks (ranks) whose memory this block reads
    bool own_loaded = true;                     // the row slots are chunks 0.. of the tile
    struct Load { uint32_t block, lambda, xi; };
    std::vector<Load> loads;
    std::vector<uint32_t> far_host;             // the groups this plan serves from shared memory
    void *slab = nullptr;
    const uint32_t *d_far = nullptr, *d_near = nullptr;
    const uint8_t *d_part = nullptr;
    size_t smem = 0;
};

struct qr_plan {
    qr::PlanDev dev{};
    std::map<int, ApplyPlan> apply_plans;       // keyed by log2(rows of the block)
    std::map<uint64_t, TilePlan> tile_plans;    // keyed by (m, block, d, fused)
    int n_sm = 0;                               // multiprocessors of the plan's device
    std::vector<uint32_t> host_gx;
    // cached values of the mask-0 group (diag(H)) for one row range, reused by every apply
    double2 *diag_cache = nullptr;              // complex values, or (diag_cache_real) doubles in the same allocation
    bool diag_cache_real = false;
    int diag_is_real = -1;                      // every c' of the mask-0 group is real (gflag bit 1)
    uint64_t diag_lo = 0, diag_hi = 0;
    int diag_terms = -1;                        // number of terms in group 0 if its mask is 0, else 0
    int device = 0;
    int n_qubits = 0;
    uint64_t dim = 0, n_terms = 0, n_groups = 0;
    void *slab = nullptr;             // one allocation holding every table of PlanDev
    // fill configuration chosen from G: staged whole-row tiles (rw,gw) or subtree blocks
    int rw = 0, gw = 0;
    uint32_t block_s = 0, n_blocks = 0;    // blocked kernel: S and the number of subtree blocks
    // lanes kernel (large G, default): rows per run = 2^lanes_log2r, warps per CTA, heavy groups
    int lanes = 0, lanes_log2r = 0, lanes_warps = 0, lanes_resync = 2, lanes_persist = 0;
    // rows kernel (large G, whole rows staged a few at a time): threads per CTA, groups per thread,
    // log2(rows per batch), log2(rows per run); rows_th == 0: not used
    int rows_th = 0, rows_ng = 0, rows_q = 0, rows_log2r = 0;
    uint32_t rows_hv_thr = 0xffffffffu, rows_hv_cap = 0;   // heavy groups: more than hv_thr terms, hv_cap of them
    int rows_hv_log2 = 5;                                  // log2(rows per heavy strip)
    int rows_regt = 0;                                     // 1: terms in registers (fill_rows_kernel<.., REGT = true>)
    int rows_sl = 0;                                       // log2(sub-batches): a batch holds 2^(rows_q + rows_sl) rows
    size_t rows_smem_bytes = 0;
    uint64_t rows_table_terms = 0;                         // entries of the kernel's shared term table
    uint32_t rows_cnt_smem = 0;                            // 1: rank-table columns staged in shared memory
    int rows_cl = 1;                                       // > 1: thread-block cluster of rows_cl CTAs per run of rows; 0: split mode
    qr::RowsSplit rows_split{};                            // split mode: trie subtrees of <= 1024 groups, CTAs per subtree
    uint32_t rows_gc = 0;                                  // groups per CTA of the cluster
    uint32_t *rows_perm = nullptr;                         // thread slot -> group table of the whole-row register variant (RowsSplit::perm)
    // split mode with EXTERNAL heavy values (heavy_values_kernel, fill.cuh): the heavy groups longest first, group -> index,
    // and the values of one chunk of rows
    bool rows_ext_heavy = false;
    uint32_t ext_nh = 0;
    uint32_t *ext_heavy_g = nullptr, *ext_hidx = nullptr;
    double2 *ext_hv = nullptr;
    uint64_t ext_hv_rows = 0;
    uint32_t n_const = 0;                  // groups whose value does not depend on the row
    uint32_t max_group_terms = 0;          // longest term list of a group
    uint32_t merge_dups = 0;               // QR_PLAN_MERGE_DUPLICATES
    uint64_t n_terms_canonical = 0;
    // lazily allocated scratch
    double2 *dot_partials = nullptr;       // per-CTA partials of <v, H v> (qr_apply_dot_device / qr_apply_p2p_dot)
    uint64_t dot_cap = 0;
    // term-rich H.v (apply_fold.cuh): bucketed copy of the term table, built at the first apply
    int fold_state = -1;                   // -1: not looked at yet, 0: the gather kernel serves this operator, 1: the fold kernel does
    bool fold_blocked = false;             // a group too long for the bucket table: gather kernel only
    void *fold_slab = nullptr;             // the tables (also read by the partner-tile kernel)
    qr::FoldDev fold{};
    // partner-tile H.v (apply_fold.cuh): segments of groups sharing x >> K, per tile size K
    struct Ptile { void *slab = nullptr; qr::PtileDev dev{}; uint32_t n_seg = 0; };
    std::map<int, Ptile> ptiles;
};

struct qr_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0, device = 0;
    // qr_apply_p2p without flags: scratch of the two NCCL all-reduce barriers
    double *d_scratch = nullptr;
    // qr_apply_p2p, tiled path: {ready[P], done[P]} epoch flags of every rank, IPC-mapped (apply_tile.cuh)
    uint64_t *d_flags = nullptr;                // this rank's flags + a CTA counter behind them
    uint64_t *peer_flags[qr::TILE_MAX_PEERS] = {nullptr};
    uint64_t **d_peer_flags = nullptr;          // the same table in device memory (p2p_ready_kernel / p2p_done_kernel)
    bool flags_ok = false;
    uint64_t epoch = 0;
};

namespace {

// ---------------------------------------------------------------------------------
// staged fill: template dispatch
// ---------------------------------------------------------------------------------
using StagedFn = void (*)(qr::PlanDev, uint32_t, uint64_t, uint64_t, uint64_t, uint64_t *, uint64_t *,
                          double2 *, uint64_t, uint32_t);
struct StagedCfg { int rw, gw; StagedFn fn, fn_pad; };      // rw = E (row strips per warp visit), gw = warps
#define QR_STAGED(E, GW) {E, GW, qr::fill_staged_kernel<E, GW, false>, qr::fill_staged_kernel<E, GW, true>}
const StagedCfg kStaged[] = {
    QR_STAGED(1, 4), QR_STAGED(1, 8), QR_STAGED(1, 16), QR_STAGED(2, 4), QR_STAGED(2, 8), QR_STAGED(2, 16),
    QR_STAGED(4, 4), QR_STAGED(4, 8), QR_STAGED(4, 16),
};

// ---- 2-D tensor maps of the output arrays (fill_staged_swz_kernel) --------------------------------------------
// cuTensorMapEncodeTiled lives in libcuda; the library links only the runtime, so the entry point is asked for once.
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            ptr = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}
// rows x (inner elements of 8 bytes), boxes of 16 elements (128 bytes) x box_rows, 128-byte swizzle
bool make_tile_map(qr::TensorMap *out, void *base, CUtensorMapDataType type, uint64_t inner, uint64_t rows, uint32_t box_rows)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(qr::TensorMap), "CUtensorMap is 128 bytes");
    const cuuint64_t dims[2] = {inner, rows}, strides[1] = {inner * 8};
    const cuuint32_t box[2] = {16, box_rows}, estr[2] = {1, 1};
    return enc(reinterpret_cast<CUtensorMap *>(out), type, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the swizzled tile is the default where it measured ahead of the plain tile, the padded tile and the rows kernel
bool swz_default(uint64_t G)
{
    if (const char *env = getenv("QR_FILL_SWZ")) return env[0] == '1';
    return G % 4 == 0 && G <= 24 && encode_tiled() != nullptr;
}
using SwzFn = void (*)(qr::PlanDev, uint32_t, uint64_t, uint64_t, uint64_t, uint64_t *, uint64_t, const qr::TensorMap, const qr::TensorMap);
SwzFn find_swz(int rw, int gw)
{
#define QR_SWZ(E, GW) if (rw == E && gw == GW) return qr::fill_staged_swz_kernel<E, GW>;
    QR_SWZ(1, 4) QR_SWZ(1, 8) QR_SWZ(1, 16) QR_SWZ(2, 4) QR_SWZ(2, 8) QR_SWZ(2, 16) QR_SWZ(4, 4) QR_SWZ(4, 8) QR_SWZ(4, 16)
#undef QR_SWZ
    return nullptr;
}

const StagedCfg *find_staged(int rw, int gw)
{
    for (const auto &c : kStaged) if (c.rw == rw && c.gw == gw) return &c;
    return nullptr;
}

// Pick (RW, GW) from G: as many resident warps per SM as the R*G*24-byte tile allows.
size_t blocked_smem(uint32_t S, int E)
{
    return (size_t)32 * E * (S + 1) * 24 + (size_t)S * (256 + sizeof(qr::GroupDesc)) + qr::FILL_BLOCKED_TCAP * 20;
}
// strips per group visit: 2 amortises the term loads of long groups (H8 2.33 vs 2.14 TB/s), 1 keeps
// occupancy for short ones (C3 4.26 vs 3.98 TB/s); profiles/r01_fill_sweep_largeG.jsonl
int blocked_strips(const qr_plan *pl)
{
    if (const char *e = getenv("QR_FILL_BLOCK_E")) return atoi(e) == 1 ? 1 : 2;
    return pl->n_terms >= 3 * pl->n_groups ? 2 : 1;
}

// Rows kernel: eligible when two batches of 2^q whole rows (plus, variant (b), the groups' extra terms) fit in
// shared memory and a thread keeps at most 2 (a) / 3 (b) groups.  QR_FILL_ROWS_REGT / _Q / _R / _HV / _HVS override
// the variant, the batch, the run length, the heavy threshold and the heavy strip (tests, sweeps).
constexpr int ROWS_MAX_NG_1024 = 3;
constexpr size_t ROWS_SMEM_CAP = MAX_SMEM - 1024;           // the kernel's static __shared__ variables count too
constexpr uint32_t ROWS_HEAVY_TERMS = 6;                 // groups with more terms than this are "heavy"
size_t rows_smem(uint64_t G, uint64_t n_extra, int q, uint64_t n_heavy = 0, int hv_log2 = 5)
{
    return (size_t)align_up((G << q) * 48 + n_extra * 20, 16) + (size_t)n_heavy * ((16ull << hv_log2) + 36);
}

bool choose_rows_shape(qr_plan *pl);

// choose_rows_shape picks the variant; the rank-table columns are staged in shared memory on top when they are small
// (short rows: G * n_qubits words <= 24 KB) and still fit
// Whole-row variant with the terms in registers: which group each thread slot owns.  A warp folds to its LONGEST light
// group, so the groups are sorted by term count (heavy ones, folded elsewhere, count 1) and dealt in warp-sized chunks:
// every chunk to the warp whose chunks so far add up to the least (longest chunk first).  H12: 16 warps x 2 chunks, the five
// chunks of 6-term groups end up alone or with a 4-term chunk instead of two to a warp.  QR_FILL_ROWS_PERM=0: mask order.
void build_rows_perm(qr_plan *pl)
{
    if (pl->rows_perm) { cudaFree(pl->rows_perm); pl->rows_perm = nullptr; }
    pl->rows_split.perm = nullptr; pl->rows_split.perm_n = 0;
    const char *env = getenv("QR_FILL_ROWS_PERM");
    const bool split = pl->rows_cl == 0;
    if ((env && env[0] == '0') || !pl->rows_regt || (pl->rows_cl != 1 && !split) || pl->rows_th != 512) return;
    const uint32_t G = (uint32_t)pl->n_groups, GP = 512u >> pl->rows_sl, NG = (uint32_t)pl->rows_ng, n_warps = GP / 32u;
    if (n_warps == 0 || NG == 0 || (!split && (uint64_t)NG * GP < G)) return;
    std::vector<uint32_t> goff(G + 1);
    if (cudaMemcpy(goff.data(), pl->dev.goff, (G + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
    auto cost = [&](uint32_t g) { const uint32_t t = goff[g + 1] - goff[g]; return t > pl->rows_hv_thr ? 1u : t; };
    // split mode: one table per subtree (groups [g0[s], g0[s + 1])), the same dealing inside each
    const uint32_t n_sub = split ? pl->rows_split.n : 1u, slots = NG * GP;
    std::vector<uint32_t> perm((size_t)n_sub * slots, 0xffffffffu);
    uint64_t sum_perm = 0, sum_id = 0;                              // longest warp, summed over the subtrees: dealt / mask order
    for (uint32_t sb = 0; sb < n_sub; sb++) {
        const uint32_t g_lo = split ? pl->rows_split.g0[sb] : 0u, g_hi = split ? pl->rows_split.g0[sb + 1] : G, Gs = g_hi - g_lo;
        if ((uint64_t)slots < Gs) return;
        std::vector<uint32_t> order(Gs);
        for (uint32_t g = 0; g < Gs; g++) order[g] = g_lo + g;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost(a) > cost(b); });
        const uint32_t n_chunks = (Gs + 31u) / 32u;
        std::vector<uint32_t> load(n_warps, 0), used(n_warps, 0);
        for (uint32_t c = 0; c < n_chunks; c++) {                       // chunks come longest first
            uint32_t w = n_warps;
            for (uint32_t v = 0; v < n_warps; v++) if (used[v] < NG && (w == n_warps || load[v] < load[w])) w = v;
            if (w == n_warps) return;                                    // cannot happen: NG * GP >= Gs
            const uint32_t k = used[w]++;
            load[w] += cost(order[c * 32u]);
            for (uint32_t l = 0; l < 32u && c * 32u + l < Gs; l++) perm[(size_t)sb * slots + (size_t)k * GP + w * 32u + l] = order[c * 32u + l];
        }
        uint32_t max_perm = 0, max_id = 0;
        for (uint32_t w = 0; w < n_warps; w++) {
            max_perm = std::max(max_perm, load[w]);
            uint32_t id = 0;
            for (uint32_t k = 0; k < NG; k++) {
                uint32_t m = 0;
                for (uint32_t l = 0; l < 32u; l++) { const uint64_t g = (uint64_t)k * GP + w * 32u + l; if (g < Gs) m = std::max(m, cost(g_lo + (uint32_t)g)); }
                id += m;
            }
            max_id = std::max(max_id, id);
        }
        sum_perm += max_perm; sum_id += max_id;
    }
    // worth it only when it shortens the longest warp by 15 % or more against mask order (molecular Hamiltonians: H8 4.21 ->
    // 4.61 TB/s); with uniform groups it only perturbs the lane <-> slot pattern of the stores (XXZ n = 27: 6.06 -> 5.47)
    if (!(env && env[0] == '1') && 100u * sum_perm > 85u * sum_id) return;
    if (cudaMalloc(reinterpret_cast<void **>(&pl->rows_perm), perm.size() * 4) != cudaSuccess) { cudaGetLastError(); pl->rows_perm = nullptr; return; }
    if (cudaMemcpy(pl->rows_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); cudaFree(pl->rows_perm); pl->rows_perm = nullptr; return;
    }
    pl->rows_split.perm = pl->rows_perm; pl->rows_split.perm_n = slots;
}

// tables of the external-heavy path: heavy groups longest first, group -> heavy index
bool setup_ext_heavy(qr_plan *pl, const std::vector<uint32_t> &goff, uint32_t thr)
{
    const uint32_t G = (uint32_t)pl->n_groups;
    std::vector<uint32_t> hg, hidx(G, 0xffffffffu);
    for (uint32_t g = 0; g < G; g++) if (goff[g + 1] - goff[g] > thr) hg.push_back(g);
    std::stable_sort(hg.begin(), hg.end(), [&](uint32_t a, uint32_t b) { return goff[a + 1] - goff[a] > goff[b + 1] - goff[b]; });
    for (uint32_t h = 0; h < hg.size(); h++) hidx[hg[h]] = h;
    if (pl->ext_heavy_g) { cudaFree(pl->ext_heavy_g); pl->ext_heavy_g = nullptr; }
    if (pl->ext_hidx) { cudaFree(pl->ext_hidx); pl->ext_hidx = nullptr; }
    if (hg.empty()) return false;
    bool ok = cudaMalloc(reinterpret_cast<void **>(&pl->ext_heavy_g), hg.size() * 4) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void **>(&pl->ext_hidx), (size_t)G * 4) == cudaSuccess &&
              cudaMemcpy(pl->ext_heavy_g, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(pl->ext_hidx, hidx.data(), (size_t)G * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return false; }
    pl->ext_nh = (uint32_t)hg.size();
    pl->rows_ext_heavy = true;
    return true;
}

bool choose_rows_shape(qr_plan *pl);
bool choose_rows_inner(qr_plan *pl);
bool choose_rows(qr_plan *pl)
{
    PlanTrace tr;
    const bool ok = choose_rows_inner(pl);
    tr.mark("  choose_rows_inner");
    if (ok) build_rows_perm(pl);
    tr.mark("  build_rows_perm");
    return ok;
}
bool choose_rows_inner(qr_plan *pl)
{
    pl->rows_cnt_smem = 0; pl->rows_cl = 1; pl->rows_gc = 0; pl->rows_split = qr::RowsSplit{}; pl->rows_ext_heavy = false;
    if (!choose_rows_shape(pl)) return false;
    const size_t cnt_bytes = (size_t)pl->n_groups * pl->n_qubits * 4;
    const char *env = getenv("QR_FILL_ROWS_CNT");
    if (pl->rows_regt && pl->rows_ng == 1 && cnt_bytes <= 24 * 1024 && pl->rows_smem_bytes + cnt_bytes <= ROWS_SMEM_CAP && !(env && env[0] == '0')) {
        pl->rows_cnt_smem = 1;
        pl->rows_smem_bytes += cnt_bytes;
    }
    return true;
}

bool choose_rows_shape(qr_plan *pl)
{
    const uint64_t G = pl->n_groups, n_extra = pl->n_terms_canonical - G;
    int q_forced = 0, r = 0;
    if ((G << 2) * 16 >= (1ull << 31)) return false;
    if (const char *env = getenv("QR_FILL_ROWS_Q")) { int v = atoi(env); if (v >= 1 && v <= 3) q_forced = v; }
    if (const char *env = getenv("QR_FILL_ROWS_R")) { int v = atoi(env); if (v >= 1 && v <= 16) r = v; }
    // group sizes decide which groups leave their owner's lane for the CTA-wide heavy path
    std::vector<uint32_t> goff(G + 1);
    if (cudaMemcpy(goff.data(), pl->dev.goff, (G + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    uint32_t thr0 = ROWS_HEAVY_TERMS;
    if (const char *env = getenv("QR_FILL_ROWS_HV")) { int v = atoi(env); thr0 = v <= 0 ? 0xffffffffu : (uint32_t)v; }
    auto heavy_count = [&](uint32_t thr) { uint64_t n = 0; for (uint64_t g = 0; g < G; g++) n += goff[g + 1] - goff[g] > thr; return n; };
    auto strip_log2 = [&](uint64_t extra, int q, uint64_t nh) {
        // strips of 128 / 64 rows (4 / 2 rows per lane in the heavy fold) when the side buffer still fits
        int hl = 5;
        while (hl < 7 && nh != 0 && rows_smem(G, extra, q, nh, hl + 1) <= ROWS_SMEM_CAP) hl++;
        if (const char *env = getenv("QR_FILL_ROWS_HVS")) { int v = atoi(env); if (v >= 5 && v <= 7 && rows_smem(G, extra, q, nh, v) <= ROWS_SMEM_CAP) hl = v; }
        return hl;
    };

    // (a) terms in registers: 512 threads x <= 2 groups x <= 6 terms; every longer group must fit the heavy path.
    //     Chosen for term-rich operators (molecular Hamiltonians, 5-6 terms per group), where the extras table of
    //     variant (b) makes shared memory the busiest unit (ncu: l1tex 59 %, issue 44 %, H12).
    //     Also for every G <= 512 (one group per thread): measured 6.67 vs 5.52 TB/s at G = 300, equal at G = 400;
    //     at G = 1000 with 1.3 terms per group variant (b) is ahead, 6.36 vs 6.20 (profiles/r03_rows_sweep.jsonl).
    int regt = ((G <= 512 || (G <= 1024 && n_extra >= G)) && thr0 == (uint32_t)qr::LANE_TERMS) ? 1 : 0;
    if (const char *env = getenv("QR_FILL_ROWS_CL")) if (atoi(env) > 1) regt = 0;
    if (const char *env = getenv("QR_FILL_ROWS_SPLIT")) if (atoi(env) > 0) regt = 0;
    if (const char *env = getenv("QR_FILL_ROWS_REGT")) regt = (env[0] == '1' && G <= 1024 && thr0 == (uint32_t)qr::LANE_TERMS) ? 1 : 0;
    if (regt) {
        const uint64_t nh = heavy_count(thr0);
        uint64_t hx = 0;                                           // terms 1.. of the heavy groups: the shared term table
        for (uint64_t g = 0; g < G; g++) if (goff[g + 1] - goff[g] > thr0) hx += goff[g + 1] - goff[g] - 1;
        // 8-row batches (one group per thread only) pay while a 4-row batch is small: G = 160 / 200: 6.7 TB/s against
        // 5.3 / 5.7; G = 300: 6.3 against 6.7; equal from 400 up
        const int q_hi = (G <= 256 || (q_forced == 3 && G <= 512)) ? 3 : 2;
        if (q_forced > q_hi) q_forced = q_hi;
        // small G: 2^sl sub-batches of 512 >> sl >= G threads each, so that every thread has a group (G = 64: 8 sub-
        // batches of 8 rows = 64-row batches of 98 KB)
        int sl_hi = 0;
        while (sl_hi < 4 && (512u >> (sl_hi + 1)) >= G) sl_hi++;
        if (const char *env = getenv("QR_FILL_ROWS_SL")) { int v = atoi(env); if (v >= 0 && v <= sl_hi) sl_hi = v; }
        // EXTERNAL heavy values (heavy_values_kernel, as in split mode): no heavy tables and no side buffer in shared memory,
        // so a batch can hold more rows -- H8 (G = 981, 57 heavy groups with 1 817 terms): 4-row batches instead of 2-row ones.
        // Default: when that is what it buys; QR_FILL_ROWS_EXTHV = 1 / 0: whenever there is a heavy group / never.
        const char *xenv = getenv("QR_FILL_ROWS_EXTHV");
        int in_q = -1, in_sl = -1, in_hl = 5;
        for (int sl = sl_hi; sl >= 0 && in_q < 0; sl--)
            for (int q = q_forced ? q_forced : q_hi; q >= (q_forced ? q_forced : 1); q--) {
                const int qb = q + sl;
                if (rows_smem(G, hx, qb, nh) > ROWS_SMEM_CAP) continue;
                int hl = std::max(strip_log2(hx, qb, nh), qb);             // a heavy strip holds whole batches
                if (hl > 7 || rows_smem(G, hx, qb, nh, hl) > ROWS_SMEM_CAP) continue;
                in_q = q; in_sl = sl; in_hl = hl;
                break;
            }
        int ex_q = -1, ex_sl = -1;
        if (nh != 0 && !(xenv && xenv[0] == '0'))
            for (int sl = sl_hi; sl >= 0 && ex_q < 0; sl--)
                for (int q = q_forced ? q_forced : q_hi; q >= (q_forced ? q_forced : 1); q--)
                    if (rows_smem(G, 0, q + sl) <= ROWS_SMEM_CAP) { ex_q = q; ex_sl = sl; break; }
        const bool use_ext = ex_q >= 0 && ((xenv && xenv[0] == '1') || in_q < 0 || ex_q + ex_sl > in_q + in_sl);
        if (use_ext || in_q >= 0) {
            const int q = use_ext ? ex_q : in_q, sl = use_ext ? ex_sl : in_sl, qb = q + sl;
            pl->rows_th = 512; pl->rows_ng = (int)((G + 511) / 512); pl->rows_q = q; pl->rows_sl = sl;
            pl->rows_log2r = r ? std::max(r, qb) : 0;
            pl->rows_hv_thr = thr0; pl->rows_regt = 1;
            if (use_ext) {
                pl->rows_hv_cap = 0; pl->rows_hv_log2 = std::max(5, qb); pl->rows_smem_bytes = rows_smem(G, 0, qb); pl->rows_table_terms = 0;
                if (!setup_ext_heavy(pl, goff, thr0)) return false;
            } else {
                pl->rows_hv_cap = (uint32_t)nh; pl->rows_hv_log2 = in_hl;
                pl->rows_smem_bytes = rows_smem(G, hx, qb, nh, in_hl); pl->rows_table_terms = hx;
            }
            return true;
        }
    }

    // (d) SPLIT: rows of any length.  K1b cuts the sorted masks into trie subtrees of <= 1024 groups; a CTA owns one subtree
    //     (terms in registers) and writes its contiguous segment of every row.  Chosen when one CTA cannot hold whole rows
    //     (G > ~2400, or (b)'s term table does not fit) and for term-rich operators with more than 1024 groups.
    const bool b_fits = G <= 1024 * (uint64_t)ROWS_MAX_NG_1024 && rows_smem(G, n_extra, 1) <= ROWS_SMEM_CAP;
    int want_split = (G > 1024 && thr0 == (uint32_t)qr::LANE_TERMS && (n_extra >= G || !b_fits)) ? 1 : 0;
    uint32_t split_s = 1024;
    if (const char *env = getenv("QR_FILL_ROWS_SPLIT")) {          // "0": never; "<S>": force, subtrees of <= S groups (tests)
        const int v = atoi(env);
        want_split = (v > 0 && thr0 == (uint32_t)qr::LANE_TERMS) ? 1 : 0;
        if (v >= 32 && v <= 1024) split_s = (uint32_t)v;
    }
    if (const char *env = getenv("QR_FILL_ROWS_CL")) if (atoi(env) > 1) want_split = 0;
    bool ext_heavy = false;
    if (want_split) {
        qr::partition_kernel<<<1, qr::K1_THREADS>>>(pl->dev, split_s);
        uint32_t meta[8] = {0};
        std::vector<uint32_t> bs, bp;
        bool ok = cudaGetLastError() == cudaSuccess && cudaMemcpy(meta, pl->dev.meta, sizeof(meta), cudaMemcpyDeviceToHost) == cudaSuccess;
        const uint32_t nb = meta[2];
        ok = ok && nb >= 1 && nb <= 32;
        if (ok) {
            bs.resize(nb + 1); bp.resize(nb);
            ok = cudaMemcpy(bs.data(), pl->dev.blk_start, (nb + 1) * 4, cudaMemcpyDeviceToHost) == cudaSuccess &&
                 cudaMemcpy(bp.data(), pl->dev.blk_p, nb * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        if (!ok) cudaGetLastError();
        int n_sm = 0;
        if (ok) ok = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pl->device) == cudaSuccess && n_sm >= (int)nb;
        if (ok) {
            uint64_t gmax = 0, gmin = G, nh_max = 0, hx_max = 0, hx_all = 0;
            for (uint32_t b = 0; b < nb; b++) {
                uint64_t nh = 0, hx = 0;
                for (uint64_t g = bs[b]; g < bs[b + 1]; g++)
                    if (goff[g + 1] - goff[g] > thr0) { nh++; hx += goff[g + 1] - goff[g] - 1; }
                gmax = std::max<uint64_t>(gmax, bs[b + 1] - bs[b]); gmin = std::min<uint64_t>(gmin, bs[b + 1] - bs[b]);
                nh_max = std::max(nh_max, nh); hx_max = std::max(hx_max, hx); hx_all += hx;
            }
            // Where it pays (profiles/r03_rows_sweep.jsonl): balanced tries of light groups -- random operators with
            // G = 1500..4000: 4.1-5.3 TB/s against 1.9-4.8 with the lanes kernel.  Molecular Hamiltonians keep their long
            // groups in the subtree around mask 0: with the heavy groups folded inside the CTAs that own them a third of
            // the CTAs folds all the heavy terms of every row (H10 1.9-2.2 TB/s against 2.5 lanes).  Their heavy groups
            // are therefore folded by a kernel of their own, one chunk of rows ahead of the fill (heavy_values_kernel:
            // EXTERNAL heavy values), and the split kernel reads them like any other value.  QR_FILL_ROWS_EXTHV=0: in-CTA.
            const char *xenv = getenv("QR_FILL_ROWS_EXTHV");
            ext_heavy = hx_all != 0 && !(xenv && xenv[0] == '0');
            if (!getenv("QR_FILL_ROWS_SPLIT") && (gmin < 256 || (!ext_heavy && hx_all * 100 > pl->n_terms_canonical * 20))) ok = false;
        }
        if (ok) {
            uint64_t gmax = 0, nh_max = 0, hx_max = 0;
            for (uint32_t b = 0; b < nb; b++) {
                uint64_t nh = 0, hx = 0;
                for (uint64_t g = bs[b]; g < bs[b + 1]; g++)
                    if (!ext_heavy && goff[g + 1] - goff[g] > thr0) { nh++; hx += goff[g + 1] - goff[g] - 1; }
                gmax = std::max<uint64_t>(gmax, bs[b + 1] - bs[b]); nh_max = std::max(nh_max, nh); hx_max = std::max(hx_max, hx);
            }
            const uint64_t si = (gmax + 3) & ~1ull;                // pitch of a buffered row of column ids
            for (int q = q_forced ? std::min(q_forced, 2) : 2; ok && q >= (q_forced ? std::min(q_forced, 2) : 1); q--) {
                auto smem_d = [&](int hl) { return (size_t)align_up((si << q) * 48 + hx_max * 20, 16) + (size_t)nh_max * ((16ull << hl) + 36); };
                int hl = 5;
                if (smem_d(hl) > ROWS_SMEM_CAP) continue;
                while (hl < 7 && nh_max != 0 && smem_d(hl + 1) <= ROWS_SMEM_CAP) hl++;
                qr::RowsSplit &sp = pl->rows_split;
                sp = qr::RowsSplit{};
                sp.n = nb;
                // CTAs in proportion to the subtrees' work (terms folded + entries written per row), at least one each,
                // n_sm in all
                std::vector<uint64_t> w(nb);
                uint64_t w_all = 0;
                for (uint32_t b = 0; b < nb; b++) {
                    w[b] = bs[b + 1] - bs[b];
                    for (uint64_t g = bs[b]; g < bs[b + 1]; g++) { const uint32_t t = goff[g + 1] - goff[g]; w[b] += (ext_heavy && t > thr0) ? 1u : t; }
                    w_all += w[b];
                }
                uint32_t given = 0;
                std::vector<uint32_t> cnt(nb);
                for (uint32_t b = 0; b < nb; b++) { cnt[b] = std::max<uint32_t>(1, (uint32_t)(w[b] * n_sm / w_all)); given += cnt[b]; }
                while (given > (uint32_t)n_sm) { uint32_t m = 0; for (uint32_t b = 1; b < nb; b++) if (cnt[b] > cnt[m]) m = b; cnt[m]--; given--; }
                while (given < (uint32_t)n_sm) {                   // hand the rest to the subtrees with the most work per CTA
                    uint32_t m = 0;
                    for (uint32_t b = 1; b < nb; b++) if (w[b] * cnt[m] > w[m] * cnt[b]) m = b;
                    cnt[m]++; given++;
                }
                sp.cta0[0] = 0;
                for (uint32_t b = 0; b < nb; b++) { sp.g0[b] = bs[b]; sp.level[b] = bp[b]; sp.cta0[b + 1] = sp.cta0[b] + cnt[b]; }
                sp.g0[nb] = bs[nb];
                pl->rows_th = 512; pl->rows_ng = (int)((gmax + 511) / 512); pl->rows_q = q; pl->rows_sl = 0; pl->rows_log2r = r ? std::max(r, q) : 0;
                pl->rows_hv_thr = thr0; pl->rows_hv_cap = (uint32_t)nh_max; pl->rows_hv_log2 = hl; pl->rows_regt = 1;
                pl->rows_smem_bytes = smem_d(hl); pl->rows_table_terms = hx_max; pl->rows_cl = 0; pl->rows_gc = (uint32_t)gmax;
                if (ext_heavy && !setup_ext_heavy(pl, goff, thr0)) return false;
                return true;
            }
        }
    }

    // (c) rows too long for one CTA (or term-rich with more than 1024 groups): a cluster of 2 or 4 CTAs splits the groups
    //     (terms in registers, <= 2 groups per thread) and the rows of every 4-row batch; entries travel to the row's
    //     owner through distributed shared memory.  Needs an even number of entries per owned row block (16-byte TMA
    //     alignment of the column ids).
    //     Measured slower than (b), (d) and mostly the lanes kernel (distributed shared memory moves 17-21 B per cycle and
    //     SM, about an SM's share of HBM: C3 3.6 TB/s with a cluster of 2 against 6.6 alone; H10 2.0 against 2.5 lanes), so
    //     it is only taken on request (QR_FILL_ROWS_CL=2|4; tests, sweeps).
    int want_cl = 0;
    if (const char *env = getenv("QR_FILL_ROWS_CL")) want_cl = (env[0] != '0' && thr0 == (uint32_t)qr::LANE_TERMS) ? atoi(env) : 0;
    for (int cl = 2; want_cl && cl <= 4; cl *= 2) {
        if (want_cl > 1 && cl != want_cl) continue;                // forced size (tests)
        const uint64_t gc = align_up((G + cl - 1) / cl, 32);
        const uint64_t rows_own = 4 / cl;
        if (gc > 1024 || ((rows_own * G) & 1)) continue;
        uint64_t nh_max = 0, hx_max = 0;
        for (int c = 0; c < cl; c++) {
            uint64_t nh = 0, hx = 0;
            for (uint64_t g = c * gc; g < std::min<uint64_t>(G, (c + 1) * gc); g++)
                if (goff[g + 1] - goff[g] > thr0) { nh++; hx += goff[g + 1] - goff[g] - 1; }
            nh_max = std::max(nh_max, nh); hx_max = std::max(hx_max, hx);
        }
        auto smem_c = [&](int hl) { return (size_t)align_up(rows_own * G * 48 + hx_max * 20, 16) + (size_t)nh_max * ((16ull << hl) + 36); };
        int hl = 5;
        if (smem_c(hl) > ROWS_SMEM_CAP) continue;
        while (hl < 7 && nh_max != 0 && smem_c(hl + 1) <= ROWS_SMEM_CAP) hl++;
        pl->rows_th = 512; pl->rows_ng = (int)((gc + 511) / 512); pl->rows_q = 2; pl->rows_sl = 0; pl->rows_log2r = r ? std::max(r, 2) : 0;
        pl->rows_hv_thr = thr0; pl->rows_hv_cap = (uint32_t)nh_max; pl->rows_hv_log2 = hl; pl->rows_regt = 1;
        pl->rows_smem_bytes = smem_c(hl); pl->rows_table_terms = hx_max; pl->rows_cl = cl; pl->rows_gc = (uint32_t)gc;
        return true;
    }

    // (b) first term in registers, the others in shared memory.  1024 threads (32 warps hide the shared-memory and
    //     FP64 latency of the term loops), up to 3 groups per thread in its 64 registers: C3 6.63 TB/s against 6.34
    //     with 512 threads (profiles/r03_rows_sweep.jsonl)
    const int th = 1024, ng = (int)((G + th - 1) / th);
    if (ng > ROWS_MAX_NG_1024) return false;
    if (q_forced > 2) q_forced = 2;
    for (int q = q_forced ? q_forced : 2; q >= (q_forced ? q_forced : 1); q--) {
        if (rows_smem(G, n_extra, q) > ROWS_SMEM_CAP) continue;
        uint32_t thr = thr0;
        uint64_t nh = heavy_count(thr);
        // 4-row batches only if every heavy group fits beside them; 2-row batches shed heavy groups (raise
        // the threshold) until the side buffer fits
        while (nh != 0 && rows_smem(G, n_extra, q, nh) > ROWS_SMEM_CAP) {
            if (q == 2 && !q_forced) break;
            thr = thr > 0x7fffffffu ? 0xffffffffu : thr * 2;
            nh = thr == 0xffffffffu ? 0 : heavy_count(thr);
        }
        if (rows_smem(G, n_extra, q, nh) > ROWS_SMEM_CAP) continue;
        const int hl = strip_log2(n_extra, q, nh);
        pl->rows_th = th; pl->rows_ng = ng; pl->rows_q = q; pl->rows_sl = 0; pl->rows_log2r = r ? std::max(r, q) : 0;   // 0: per window
        pl->rows_hv_thr = nh ? thr : 0xffffffffu; pl->rows_hv_cap = (uint32_t)nh; pl->rows_hv_log2 = hl; pl->rows_regt = 0;
        pl->rows_smem_bytes = rows_smem(G, n_extra, q, nh, hl); pl->rows_table_terms = n_extra;
        return true;
    }
    return false;
}

void choose_staged(qr_plan *pl)
{
    pl->rw = pl->gw = 0;
    pl->rows_th = 0;
    if (const char *env = getenv("QR_FILL_ROWS")) if (env[0] == '1') choose_rows(pl);   // force (tests, sweeps)
    const uint64_t G = pl->n_groups;
    pl->block_s = 0;
    pl->lanes = 0;
    // measured (profiles/r02_lanes_sweep.jsonl): runs of 32 rows, one run per CTA, 8 warps up to 1024 groups
    // and 16 beyond, sibling CTAs as one cluster: C3 4.85 TB/s, H10 2.53, H12 2.82, H8 2.15
    pl->lanes_log2r = 5; pl->lanes_warps = G <= 1024 ? 8 : 16;
    if (const char *env = getenv("QR_FILL_LANES_R")) { int k = atoi(env); if (k >= 5 && k <= qr::FILL_LANES_MAXLOG2R) pl->lanes_log2r = k; }
    if (const char *env = getenv("QR_FILL_LANES_W")) { int w = atoi(env); if (w == 8 || w == 16 || w == 32) pl->lanes_warps = w; }
    pl->lanes_resync = 2;                                            // 0 none, 1 per CTA, 2 per cluster of sibling CTAs
    if (const char *env = getenv("QR_FILL_LANES_SYNC")) { int v = atoi(env); if (v >= 0 && v <= 2) pl->lanes_resync = v; }
    pl->lanes_persist = 0;                                           // 1: persistent CTAs looping over runs (measured slower)
    if (const char *env = getenv("QR_FILL_LANES_PERSIST")) pl->lanes_persist = env[0] != '0';
    if (const char *env = getenv("QR_FILL_LANES")) if (env[0] == '1') { pl->lanes = 1; return; }   // force (tests, sweeps)
    if (const char *env = getenv("QR_FILL_BLOCK")) {         // "S" override: force the blocked kernel
        int S = atoi(env);
        if (S >= 32 && blocked_smem(S, 2) <= MAX_SMEM) { pl->block_s = (uint32_t)S; return; }
    }
    if (const char *env = getenv("QR_FILL_CFG")) {           // "RW,GW" override for experiments
        int rw = 0, gw = 0;
        if (sscanf(env, "%d,%d", &rw, &gw) == 2 && find_staged(rw, gw) && (uint64_t)32 * rw * G * 24 <= MAX_SMEM) {
            pl->rw = rw; pl->gw = gw; return;
        }
    }
    // Measured on B200 (profiles/r01_fill_sweep.jsonl): 8 warps splitting the groups, each visiting
    // two 32-row strips per group: C2 6.56 TB/s, C4 6.72 TB/s (one strip: 5.45 / 6.39; four: 5.95 / 4.48).
    const uint64_t row_bytes = G * 24;
    if (G < 8 && 64 * row_bytes <= MAX_SMEM)  { pl->rw = 2; pl->gw = 4; }
    else if (64 * row_bytes <= 113 * 1024)    { pl->rw = 2; pl->gw = 8; }   // 2+ CTAs/SM
    else if (32 * row_bytes <= 113 * 1024)    { pl->rw = 1; pl->gw = 8; }
    if (pl->rw) {
        // The staged kernel (lane <-> row) is at its best on short rows with a long diagonal group (spin chains and
        // lattices: G = n + 1, 97-100 % of peak).  On operators without such a group it pays per-group overheads and,
        // from G = 76, runs one 8-warp CTA per SM: 2.3-4.4 TB/s for G = 32..150, where the rows kernel (thread <->
        // group, sub-batches) holds 6.2-6.9 TB/s (profiles/r03_rows_sweep.jsonl).  With heavy groups the rows kernel
        // takes over from G = 76 only (C2 / C4 / XXZ n = 24: 4.8-5.1 TB/s against 6.0-6.6 staged).
        const char *env = getenv("QR_FILL_ROWS");
        // Chains whose G = n + 1 is a multiple of 4 hit 4- and 8-way bank conflicts in the plain staged tile (lanes G * 16 B
        // apart).  Up to G = 24 the swizzled tile (fill_staged_swz_kernel) removes them: XXZ n = 19 / 23 (G = 20 / 24)
        // 5.19 / 6.86 TB/s against 4.62 / 5.25 plain and 5.47 rows; from G = 28 the rows kernel is ahead of both
        // (n = 27 / 31: 6.05 / 6.65 rows, 5.84 / 6.09 swizzled, 5.40 plain; profiles/r05_swz_sweep*.jsonl).
        if (G >= 24 && !(env && (env[0] == '0' || env[0] == '1')) && choose_rows(pl)) {
            const bool no_long_group = G >= 32 && pl->rows_hv_cap == 0, long_rows = G > 75, conflicts = G % 4 == 0 && (G >= 28 || !swz_default(G));
            if (!(pl->rows_regt && (no_long_group || long_rows || conflicts))) pl->rows_th = 0;
        }
    } else {
        // whole rows do not fit (twice) in shared memory: the lanes kernel (lane <-> group, rows in
        // Gray-code order).  QR_FILL_LANES=0 falls back to subtree blocks of <= 32 groups through
        // shared memory (profiles/r01_fill_sweep_largeG.jsonl: C3 4.52 TB/s at S=32, 3.73 at 64).
        pl->lanes = 1;
        if (const char *env = getenv("QR_FILL_LANES")) if (env[0] == '0') { pl->lanes = 0; pl->block_s = 32; }
        // rows that fit shared memory two batches at a time (G <= ~2400 unless the term table is huge): whole
        // rows through the TMA (fill_rows_kernel), measured ahead of the lanes kernel for every G from 160 up
        // (profiles/r03_rows_sweep.jsonl); the lanes kernel stays the path for everything else and for the
        // ragged windows build_rows cannot align
        const char *env = getenv("QR_FILL_ROWS");
        if (!(env && env[0] == '0')) choose_rows(pl);
    }
}

int launch_direct(const qr_plan *pl, uint64_t lo, uint64_t hi, uint64_t out_row0, uint64_t req_hi,
                  uint64_t indptr_base, uint64_t *indptr, uint64_t *indices, double2 *data, cudaStream_t st);

}  // namespace

// =====================================================================================
// error / misc
// =====================================================================================
extern "C" const char *qr_last_error(void) { return g_err.c_str(); }
extern "C" int qr_version(void) { return QR_VERSION; }
extern "C" uint64_t qr_kernel_launches(void) { return g_launches.load(); }

// =====================================================================================
// plan
// =====================================================================================
// Programmatic dependent launch for the canonicalise -> fill -> canonicalise -> ... chain (QR_PDL=1): each kernel is launched
// while its predecessor still runs and waits for it on the device (griddepcontrol.wait, plan.cuh).  Off by default: measured
// on the bench step inside a CUDA graph it changes nothing (85.5 against 85.0 us; profiles/r04_summary.md) -- the 10 us
// round 1 attributed to the canonicalisation were the event-record nodes its bench put between the kernels.
static bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("QR_PDL"); return e && e[0] == '1'; }();
    return on;
}
static void pdl_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute &attr)
{
    if (!pdl_enabled()) return;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
}

static int run_canonicalise(qr_plan *pl, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr;
    cfg.gridDim = dim3(1, 1, 1); cfg.blockDim = dim3(qr::K1_THREADS, 1, 1); cfg.stream = st;
    pdl_config(cfg, attr);
    QR_CUDA(cudaLaunchKernelEx(&cfg, qr::canonicalise_kernel, pl->dev, pl->merge_dups));
    QR_LAUNCH_CHECK("canonicalise_kernel");
    return QR_OK;
}

static int run_partition(qr_plan *pl, cudaStream_t st)
{
    if (!pl->block_s) return QR_OK;
    qr::partition_kernel<<<1, qr::K1_THREADS, 0, st>>>(pl->dev, pl->block_s);
    QR_LAUNCH_CHECK("partition_kernel");
    return QR_OK;
}

extern "C" int qr_plan_create(int n_qubits, const qr_term *terms, size_t n_terms, int device,
                              uint32_t flags, qr_plan **out)
{
    if (!out) return fail(QR_ERR_INVALID, "qr_plan_create: out is NULL");
    *out = nullptr;
    if (!terms || n_terms == 0)
        return fail(QR_ERR_INVALID, "qr_plan_create: at least one term must be supplied");   // lib.rs:358-360
    if (n_qubits < 1) return fail(QR_ERR_INVALID, "qr_plan_create: n_qubits must be >= 1");
    if (n_qubits > 32) return fail(QR_ERR_UNSUPPORTED, "qr_plan_create: n_qubits > 32 is not supported");
    if (n_terms > 0x7fffffffu / 32) return fail(QR_ERR_UNSUPPORTED, "qr_plan_create: too many terms");
    const uint64_t dim = 1ull << n_qubits;
    for (size_t t = 0; t < n_terms; t++)
        if (terms[t].x >= dim || terms[t].z >= dim)
            return fail(QR_ERR_INVALID, "qr_plan_create: term mask has bits outside n_qubits");

    QR_CUDA(cudaSetDevice(device));
    qr_plan *pl = new (std::nothrow) qr_plan();
    if (!pl) return fail(QR_ERR_OOM, "qr_plan_create: host allocation failed");
    pl->device = device; pl->n_qubits = n_qubits; pl->dim = dim; pl->n_terms = n_terms;
    if (cudaDeviceGetAttribute(&pl->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || pl->n_sm < 1) { cudaGetLastError(); pl->n_sm = 148; }
    pl->merge_dups = (flags & QR_PLAN_MERGE_DUPLICATES) ? 1u : 0u;

    const size_t T = n_terms;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_raw = carve(T * sizeof(qr_term));
    const size_t o_ka = carve(T * 4), o_kb = carve(T * 4), o_ia = carve(T * 4), o_ib = carve(T * 4);
    const size_t o_tz = carve(T * 4), o_tc = carve(T * 16), o_perm = carve(T * 4);
    const size_t o_gx = carve(T * 4), o_goff = carve((T + 1) * 4);
    const size_t o_cnt = carve(T * 128), o_lr5 = carve(T * 128), o_meta = carve(32);
    const size_t o_bs = carve((T + 1) * 4), o_bp = carve(T * 4);
    const size_t o_gf = carve(T * 4), o_gc = carve(T * 16), o_gd = carve(T * sizeof(qr::GroupDesc));
    const size_t o_ct = carve(T * 128);
    const size_t o_lx = carve(T * 8), o_lz = carve(T * 4 * qr::LANE_TERMS), o_lc = carve(T * 16 * qr::LANE_TERMS);
    PlanTrace tr;
    cudaError_t e = cudaMalloc(&pl->slab, off);
    if (e != cudaSuccess) { delete pl; return fail(QR_ERR_OOM, std::string("qr_plan_create: cudaMalloc: ") + cudaGetErrorString(e)); }
    tr.mark("cudaMalloc slab");
    char *b = static_cast<char *>(pl->slab);
    qr::PlanDev &d = pl->dev;
    d.n_qubits = n_qubits; d.n_terms = (uint32_t)T;
    d.raw = reinterpret_cast<const qr_term *>(b + o_raw);
    d.key_a = reinterpret_cast<uint32_t *>(b + o_ka); d.key_b = reinterpret_cast<uint32_t *>(b + o_kb);
    d.idx_a = reinterpret_cast<uint32_t *>(b + o_ia); d.idx_b = reinterpret_cast<uint32_t *>(b + o_ib);
    d.tz = reinterpret_cast<uint32_t *>(b + o_tz); d.tc = reinterpret_cast<double2 *>(b + o_tc);
    d.perm = reinterpret_cast<uint32_t *>(b + o_perm);
    d.gx = reinterpret_cast<uint32_t *>(b + o_gx); d.goff = reinterpret_cast<uint32_t *>(b + o_goff);
    d.cnt = reinterpret_cast<uint32_t *>(b + o_cnt); d.lr5 = reinterpret_cast<uint32_t *>(b + o_lr5);
    d.meta = reinterpret_cast<uint32_t *>(b + o_meta);
    d.blk_start = reinterpret_cast<uint32_t *>(b + o_bs); d.blk_p = reinterpret_cast<uint32_t *>(b + o_bp);
    d.gflag = reinterpret_cast<uint32_t *>(b + o_gf); d.gconst = reinterpret_cast<double2 *>(b + o_gc);
    d.gdesc = reinterpret_cast<qr::GroupDesc *>(b + o_gd);
    d.cnt_t = reinterpret_cast<uint32_t *>(b + o_ct);
    d.lt_xn = reinterpret_cast<uint2 *>(b + o_lx); d.lt_z = reinterpret_cast<uint32_t *>(b + o_lz);
    d.lt_c = reinterpret_cast<double2 *>(b + o_lc);

    auto bail = [&](int code) { cudaFree(pl->slab); delete pl; return code; };
    e = cudaMemcpy(b + o_raw, terms, T * sizeof(qr_term), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(QR_ERR_CUDA, std::string("qr_plan_create: H2D: ") + cudaGetErrorString(e)));
    tr.mark("H2D terms");
    int rc = run_canonicalise(pl, nullptr);
    if (rc != QR_OK) return bail(rc);
    uint32_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpy(meta, d.meta, sizeof(meta), cudaMemcpyDeviceToHost);
    tr.mark("K1 canonicalise + meta D2H");
    if (e != cudaSuccess) return bail(fail(QR_ERR_CUDA, std::string("qr_plan_create: canonicalise: ") + cudaGetErrorString(e)));
    pl->n_groups = meta[0];
    pl->n_const = meta[4];
    pl->max_group_terms = meta[1];
    pl->n_terms_canonical = meta[5];
    if (pl->n_groups == 0 || pl->n_groups > T) return bail(fail(QR_ERR_CUDA, "qr_plan_create: canonicalisation produced no groups"));
    choose_staged(pl);
    tr.mark("choose kernels");
    if (pl->block_s) {
        rc = run_partition(pl, nullptr);
        if (rc != QR_OK) return bail(rc);
        e = cudaMemcpy(meta, d.meta, sizeof(meta), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return bail(fail(QR_ERR_CUDA, std::string("qr_plan_create: partition: ") + cudaGetErrorString(e)));
        pl->n_blocks = meta[2];
        if (pl->n_blocks == 0) return bail(fail(QR_ERR_CUDA, "qr_plan_create: partition produced no blocks"));
    }
    *out = pl;
    return QR_OK;
}

extern "C" int qr_plan_destroy(qr_plan *pl)
{
    if (!pl) return QR_OK;
    cudaSetDevice(pl->device);
    for (auto &kv : pl->apply_plans) if (kv.second.slab) cudaFree(kv.second.slab);
    for (auto &kv : pl->tile_plans) if (kv.second.slab) cudaFree(kv.second.slab);
    if (pl->diag_cache) cudaFree(pl->diag_cache);
    if (pl->dot_partials) cudaFree(pl->dot_partials);
    if (pl->fold_slab) cudaFree(pl->fold_slab);
    if (pl->rows_perm) cudaFree(pl->rows_perm);
    if (pl->ext_heavy_g) cudaFree(pl->ext_heavy_g);
    if (pl->ext_hidx) cudaFree(pl->ext_hidx);
    if (pl->ext_hv) cudaFree(pl->ext_hv);
    for (auto &kv : pl->ptiles) if (kv.second.slab) cudaFree(kv.second.slab);
    if (pl->slab) cudaFree(pl->slab);
    delete pl;
    return QR_OK;
}

extern "C" int qr_plan_info(const qr_plan *pl, qr_plan_info_t *info)
{
    if (!pl || !info) return fail(QR_ERR_INVALID, "qr_plan_info: NULL argument");
    info->n_qubits = pl->n_qubits; info->device = pl->device; info->dim = pl->dim;
    info->n_terms = pl->n_terms; info->n_groups = pl->n_groups; info->nnz = pl->n_groups * pl->dim;
    return QR_OK;
}

extern "C" int qr_plan_groups(const qr_plan *pl, uint64_t *xmask, uint32_t *group_offsets, uint32_t *term_order)
{
    if (!pl) return fail(QR_ERR_INVALID, "qr_plan_groups: NULL plan");
    QR_CUDA(cudaSetDevice(pl->device));
    const size_t G = pl->n_groups, T = pl->n_terms;
    if (xmask) {
        std::vector<uint32_t> tmp(G);
        QR_CUDA(cudaMemcpy(tmp.data(), pl->dev.gx, G * 4, cudaMemcpyDeviceToHost));
        for (size_t g = 0; g < G; g++) xmask[g] = tmp[g];
    }
    if (group_offsets) QR_CUDA(cudaMemcpy(group_offsets, pl->dev.goff, (G + 1) * 4, cudaMemcpyDeviceToHost));
    if (term_order) QR_CUDA(cudaMemcpy(term_order, pl->dev.perm, T * 4, cudaMemcpyDeviceToHost));
    return QR_OK;
}

extern "C" int qr_plan_canonical_terms(const qr_plan *pl, uint64_t *count)
{
    if (!pl || !count) return fail(QR_ERR_INVALID, "qr_plan_canonical_terms: NULL argument");
    *count = pl->n_terms_canonical;
    return QR_OK;
}

extern "C" const char *qr_plan_fill_kernel(const qr_plan *pl)
{
    if (!pl) return "";
    if (pl->rows_th) return "fill_rows_kernel";
    if (pl->lanes) return "fill_lanes_kernel";
    if (pl->block_s) return "fill_blocked_kernel";
    if (pl->rw) return swz_default(pl->n_groups) && pl->n_groups % 2 == 0 ? "fill_staged_swz_kernel" : "fill_staged_kernel";
    return "fill_direct_kernel";
}

extern "C" int qr_plan_canonicalise_async(qr_plan *pl, void *stream)
{
    if (!pl) return fail(QR_ERR_INVALID, "qr_plan_canonicalise_async: NULL plan");
    QR_CUDA(cudaSetDevice(pl->device));
    int rc = run_canonicalise(pl, as_stream(stream));
    return rc != QR_OK ? rc : run_partition(pl, as_stream(stream));
}

// =====================================================================================
// CSR build
// =====================================================================================
namespace {

int launch_direct(const qr_plan *pl, uint64_t lo, uint64_t hi, uint64_t out_row0, uint64_t req_hi,
                  uint64_t indptr_base, uint64_t *indptr, uint64_t *indices, double2 *data, cudaStream_t st)
{
    if (hi <= lo) return QR_OK;
    const uint64_t first = lo & ~(uint64_t)31;
    const uint64_t warps = (hi - first + 31) / 32;
    const uint64_t per_cta = qr::FILL_DIRECT_THREADS / 32;
    const uint64_t ctas = (warps + per_cta - 1) / per_cta;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_direct: row window too large for one launch");
    qr::fill_direct_kernel<<<(unsigned)ctas, qr::FILL_DIRECT_THREADS, 0, st>>>(
        pl->dev, (uint32_t)pl->n_groups, lo, hi, out_row0, req_hi, indptr_base, indptr, indices, data);
    QR_LAUNCH_CHECK("fill_direct_kernel");
    return QR_OK;
}

int build_rows(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, uint64_t *d_indptr, uint64_t *d_indices,
               double2 *d_data, uint32_t flags, cudaStream_t st)
{
    const uint64_t G = pl->n_groups;
    const uint64_t indptr_base = (flags & QR_INDPTR_GLOBAL) ? row_lo * G : 0;
    uint64_t lo = row_lo, hi = row_hi;
    if (!(flags & QR_FILL_DIRECT) && pl->rows_th) {
        const int q = pl->rows_q, qb = pl->rows_q + pl->rows_sl;   // log2 rows per thread / per batch
        // rows per run: 256 (C3 6.27 TB/s; 128: 6.04, 64: 5.66), shorter when the window does not hold enough runs
        // to spread evenly over the persistent CTAs (2^16 rows of H8: 32-row runs 3.28 TB/s, 128-row runs 3.02)
        int k = pl->rows_log2r ? pl->rows_log2r : std::max(8, qb + 3);   // big batches (short rows): 8 batches per run
        if (!pl->rows_log2r) {
            const uint64_t span = row_hi - row_lo;
            while (k > 5 && k > qb) {
                const uint64_t runs = span >> k;
                const uint64_t sm = (uint64_t)pl->n_sm;
                if (runs >= sm && (runs + sm - 1) / sm * sm * 100 <= runs * 104) break;    // <= 4 % idle tail on one persistent CTA per SM
                k--;
            }
        }
        k = std::max(k, qb);
        while (k > qb && align_up(row_lo, 1ull << k) + (1ull << k) > row_hi) k--;     // short windows: shorter runs
        const uint64_t R = 1ull << k;
        const uint64_t s0 = align_up(row_lo, R), s1 = row_hi / R * R;
        // the bulk copies need 16-byte aligned global addresses: indices + (s0 - row_lo) * G * 8
        const bool aligned = pl->rows_cl == 0 || (((s0 - row_lo) * G) & 1) == 0;    // split mode handles odd segment starts itself
        if (s1 > s0 && aligned && (s1 - s0) / R <= 0xffffffffull && (!pl->rows_ext_heavy || R % 32 == 0)) {
            const uint64_t n_runs = (s1 - s0) / R, n_extra = pl->rows_table_terms;
            const size_t smem = pl->rows_smem_bytes;
            using RowsFn = void (*)(qr::PlanDev, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                                    uint64_t, uint64_t, uint64_t, uint64_t *, uint64_t *, double2 *, uint64_t, const qr::RowsSplit);
            RowsFn kern = nullptr;
            const bool hv_any = pl->rows_hv_cap != 0 || pl->rows_ext_heavy;    // HEAVY instances: in-CTA heavy phase or external values
#define QR_ROWS_Q(NG_, TH_, RG_, HV_) \
            (q == 2 ? (RowsFn)qr::fill_rows_kernel<NG_, 2, TH_, RG_, HV_, false, 1> : (RowsFn)qr::fill_rows_kernel<NG_, 1, TH_, RG_, HV_, false, 1>)
#define QR_ROWS_CASE(NG_, TH_, RG_) \
            if (q <= 2 && pl->rows_cl == 1 && !pl->rows_cnt_smem && pl->rows_ng == NG_ && pl->rows_th == TH_ && pl->rows_regt == (RG_ ? 1 : 0)) \
                kern = hv_any ? QR_ROWS_Q(NG_, TH_, RG_, true) : QR_ROWS_Q(NG_, TH_, RG_, false);
            QR_ROWS_CASE(1, 1024, false) QR_ROWS_CASE(2, 1024, false) QR_ROWS_CASE(3, 1024, false)
            QR_ROWS_CASE(1, 512, true) QR_ROWS_CASE(2, 512, true)
            // one group per thread, terms in registers: 8-row batches and the rank table in shared memory exist here only
#define QR_ROWS_ONE(Q_, HV_, CS_) \
            if (pl->rows_cl == 1 && pl->rows_ng == 1 && pl->rows_regt && q == Q_ && hv_any == HV_ && (pl->rows_cnt_smem != 0) == CS_) \
                kern = (RowsFn)qr::fill_rows_kernel<1, Q_, 512, true, HV_, CS_, 1>;
            QR_ROWS_ONE(3, false, false) QR_ROWS_ONE(3, true, false)
            QR_ROWS_ONE(1, false, true) QR_ROWS_ONE(1, true, true) QR_ROWS_ONE(2, false, true) QR_ROWS_ONE(2, true, true)
            QR_ROWS_ONE(3, false, true) QR_ROWS_ONE(3, true, true)
#undef QR_ROWS_ONE
            // clusters: terms in registers, 4-row batches
#define QR_ROWS_CL(NG_, HV_, CL_) \
            if (pl->rows_cl == CL_ && pl->rows_ng == NG_ && (pl->rows_hv_cap != 0) == HV_) kern = (RowsFn)qr::fill_rows_kernel<NG_, 2, 512, true, HV_, false, CL_>;
            QR_ROWS_CL(1, false, 2) QR_ROWS_CL(1, true, 2) QR_ROWS_CL(2, false, 2) QR_ROWS_CL(2, true, 2)
            QR_ROWS_CL(1, false, 4) QR_ROWS_CL(1, true, 4) QR_ROWS_CL(2, false, 4) QR_ROWS_CL(2, true, 4)
#undef QR_ROWS_CL
            // split mode: a CTA per trie subtree of <= 1024 groups
#define QR_ROWS_SPLIT(NG_, HV_) \
            if (pl->rows_cl == 0 && pl->rows_ng == NG_ && (pl->rows_hv_cap != 0 || pl->rows_ext_heavy) == HV_) \
                kern = q == 2 ? (RowsFn)qr::fill_rows_kernel<NG_, 2, 512, true, HV_, false, 0> : (RowsFn)qr::fill_rows_kernel<NG_, 1, 512, true, HV_, false, 0>;
            QR_ROWS_SPLIT(1, false) QR_ROWS_SPLIT(1, true) QR_ROWS_SPLIT(2, false) QR_ROWS_SPLIT(2, true)
#undef QR_ROWS_SPLIT
#undef QR_ROWS_Q
#undef QR_ROWS_CASE
            // decoupled warps (fill.cuh, DEC): whole rows (NG = 1, 2) and split mode, Q = 1, 2
            {
                const char *denv = getenv("QR_FILL_ROWS_DEC");
                // measured (profiles/r06_summary.md): split mode H10 2.49 -> 3.06 TB/s, H11 2.37 -> 2.89, random G = 3000 4.96 -> 5.4;
                // whole rows H8 4.52 -> 4.37, H12 5.17 -> 5.05 (their limit is the slowest warp's chain, not the barrier): split only
                const bool want = denv ? denv[0] != '0' : pl->rows_cl == 0;
                pl->rows_split.dec_block = denv && denv[0] == '2';
                const bool hv = pl->rows_hv_cap != 0 || pl->rows_ext_heavy;
                if (want && q >= 1 && q <= 2 && pl->rows_th == 512 && pl->rows_regt && !pl->rows_cnt_smem && (pl->rows_cl == 0 || pl->rows_cl == 1) &&
                    (pl->rows_ng == 1 || pl->rows_ng == 2)) {
#define QR_ROWS_DEC(NG_, Q_, HV_, CL_) \
                    if (pl->rows_ng == NG_ && q == Q_ && hv == HV_ && pl->rows_cl == CL_) kern = (RowsFn)qr::fill_rows_kernel<NG_, Q_, 512, true, HV_, false, CL_, true>;
                    QR_ROWS_DEC(1, 1, false, 1) QR_ROWS_DEC(1, 1, true, 1) QR_ROWS_DEC(1, 2, false, 1) QR_ROWS_DEC(1, 2, true, 1)
                    QR_ROWS_DEC(2, 1, false, 1) QR_ROWS_DEC(2, 1, true, 1) QR_ROWS_DEC(2, 2, false, 1) QR_ROWS_DEC(2, 2, true, 1)
                    QR_ROWS_DEC(1, 1, false, 0) QR_ROWS_DEC(1, 1, true, 0) QR_ROWS_DEC(1, 2, false, 0) QR_ROWS_DEC(1, 2, true, 0)
                    QR_ROWS_DEC(2, 1, false, 0) QR_ROWS_DEC(2, 1, true, 0) QR_ROWS_DEC(2, 2, false, 0) QR_ROWS_DEC(2, 2, true, 0)
#undef QR_ROWS_DEC
                }
            }
            if (!kern) return fail(QR_ERR_UNSUPPORTED, "fill_rows: no kernel instance for this plan");
            QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const uint32_t cl = (uint32_t)pl->rows_cl, gc = cl != 1 ? pl->rows_gc : (uint32_t)G;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3((unsigned)pl->rows_th, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            uint64_t ctas = 0;
            if (cl > 1) {                                          // persistent clusters: as many as are resident at once
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cfg.gridDim = dim3(cl, 1, 1);
                int max_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
                    cudaGetLastError();
                    return fail(QR_ERR_CUDA, "fill_rows: the cluster does not fit on the device");
                }
                ctas = std::min<uint64_t>(n_runs, (uint64_t)max_clusters) * cl;
            } else if (cl == 0) {
                ctas = pl->rows_split.cta0[pl->rows_split.n];      // fixed at plan creation: CTAs per subtree
            } else {
                int per_sm = 0, n_sm = 0;
                QR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pl->rows_th, smem));
                QR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, pl->device));
                if (per_sm < 1) return fail(QR_ERR_CUDA, "fill_rows: kernel does not fit on an SM");
                ctas = std::min<uint64_t>(n_runs, (uint64_t)per_sm * (uint64_t)n_sm);   // persistent CTAs
            }
            cfg.gridDim = dim3((unsigned)ctas, 1, 1);
            int rc = launch_direct(pl, row_lo, s0, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            if (rc != QR_OK) return rc;
            if (pl->rows_ext_heavy) {
                // chunks of rows: the heavy groups' values for the chunk (heavy_values_kernel), then the split fill reads them.
                // 2^16 rows (or the window) per chunk: enough runs for every CTA of every subtree, nh MB of values
                const uint64_t chunk = std::max<uint64_t>(R, std::min<uint64_t>(s1 - s0, 1ull << 16)) / R * R;
                if (!pl->ext_hv || pl->ext_hv_rows < chunk) {
                    if (pl->ext_hv) { QR_CUDA(cudaStreamSynchronize(st)); cudaFree(pl->ext_hv); pl->ext_hv = nullptr; pl->ext_hv_rows = 0; }
                    QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&pl->ext_hv), (size_t)pl->ext_nh * chunk * sizeof(double2)));
                    pl->ext_hv_rows = chunk;
                }
                qr::RowsSplit sp = pl->rows_split;
                sp.ext_hv = pl->ext_hv; sp.ext_hidx = pl->ext_hidx;
                for (uint64_t c0 = s0; c0 < s1; c0 += chunk) {
                    const uint64_t len = std::min(chunk, s1 - c0);
                    const bool e4 = len % 128 == 0;
                    const uint64_t items = (uint64_t)pl->ext_nh * (len / (e4 ? 128 : 32));
                    const uint64_t hctas = (items + 7) / 8;
                    if (hctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_rows: heavy-value chunk too large");
                    if (e4) qr::heavy_values_kernel<4><<<(unsigned)hctas, 256, 0, st>>>(pl->dev, pl->ext_heavy_g, pl->ext_nh, c0, (uint32_t)len, pl->ext_hv);
                    else qr::heavy_values_kernel<1><<<(unsigned)hctas, 256, 0, st>>>(pl->dev, pl->ext_heavy_g, pl->ext_nh, c0, (uint32_t)len, pl->ext_hv);
                    QR_LAUNCH_CHECK("heavy_values_kernel");
                    sp.ext_rows = (uint32_t)len; sp.ext_row0 = (uint32_t)c0;
                    QR_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->dev, (uint32_t)G, gc, (uint32_t)n_extra, (uint32_t)k, (uint32_t)pl->rows_sl, (uint32_t)(len / R),
                                               pl->rows_hv_thr, pl->rows_hv_cap, (uint32_t)pl->rows_hv_log2, c0, row_lo, indptr_base, d_indptr,
                                               d_indices, d_data, row_hi - row_lo, sp));
                    QR_LAUNCH_CHECK("fill_rows_kernel");
                }
                return launch_direct(pl, s1, row_hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            }
            QR_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->dev, (uint32_t)G, gc, (uint32_t)n_extra, (uint32_t)k, (uint32_t)pl->rows_sl, (uint32_t)n_runs,
                                       pl->rows_hv_thr, pl->rows_hv_cap, (uint32_t)pl->rows_hv_log2, s0, row_lo, indptr_base, d_indptr,
                                       d_indices, d_data, row_hi - row_lo, pl->rows_split));
            QR_LAUNCH_CHECK("fill_rows_kernel");
            return launch_direct(pl, s1, row_hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
        }
    }
    if (!(flags & QR_FILL_DIRECT) && pl->lanes) {
        int k = pl->lanes_log2r;
        while (k > 5 && align_up(row_lo, 1ull << k) + (1ull << k) > row_hi) k--;     // short windows: shorter runs
        while (k > 5 && (G << k) >= (1ull << 28)) k--;                                // 32-bit byte offsets inside a run
        const uint64_t R = 1ull << k;
        const uint64_t s0 = align_up(row_lo, R), s1 = row_hi / R * R;
        if (s1 > s0 && (G << k) < (1ull << 28)) {
            const uint32_t LW = (uint32_t)pl->lanes_warps;
            const uint32_t n_light = (uint32_t)((G + 32ull * LW - 1) / (32ull * LW));
            // persistent CTAs: J groups of n_light sibling CTAs, as many as are resident at once
            const uint64_t n_runs = (s1 - s0) / R;
            if (n_runs > 0xffffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_lanes: row window too large for one launch");
            const uint32_t resident = (uint32_t)pl->n_sm * (32u / LW);
            const uint64_t J = std::min<uint64_t>(n_runs, std::max<uint32_t>(1u, (resident + n_light - 1) / n_light));
            uint64_t J2 = pl->lanes_persist ? J : n_runs;
            if (J2 * n_light > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_lanes: row window too large for one launch");
            const size_t smem = (size_t)LW * pl->n_qubits * 32 * 4;
            int rc = launch_direct(pl, row_lo, s0, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            if (rc != QR_OK) return rc;
            using LanesFn = void (*)(qr::PlanDev, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint64_t, uint64_t,
                                     uint64_t, uint64_t *, uint64_t *, double2 *, uint64_t);
            const LanesFn kern = LW == 8 ? (LanesFn)qr::fill_lanes_kernel<qr::FILL_LANES_NT, 8>
                               : LW == 32 ? (LanesFn)qr::fill_lanes_kernel<qr::FILL_LANES_NT, 32>
                                          : (LanesFn)qr::fill_lanes_kernel<qr::FILL_LANES_NT, 16>;
            QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // sibling CTAs (the other group batches of the same rows) form one thread-block cluster and
            // rendezvous every 32 rows: every segment of a row is written in the same short time window
            const bool cluster = pl->lanes_resync == 2 && n_light > 1 && n_light <= 8;
            uint32_t resync = (uint32_t)pl->lanes_resync;
            if (resync == 2 && !cluster) resync = 1;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(32 * LW, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            if (cluster) {
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = n_light; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cfg.gridDim = dim3(n_light, 1, 1);
                int max_clusters = 0;
                if (!pl->lanes_persist) { /* one cluster per run */ }
                else if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0)
                    J2 = std::min<uint64_t>(n_runs, (uint64_t)max_clusters);
                else
                    cudaGetLastError();
            }
            cfg.gridDim = dim3((unsigned)(J2 * n_light), 1, 1);
            QR_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->dev, (uint32_t)G, n_light, (uint32_t)k, resync, (uint32_t)n_runs, s0, row_lo,
                                       indptr_base, d_indptr, d_indices, d_data, row_hi - row_lo));
            QR_LAUNCH_CHECK("fill_lanes_kernel");
            return launch_direct(pl, s1, row_hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
        }
    }
    if (!(flags & QR_FILL_DIRECT) && pl->block_s) {
        const uint64_t s0 = align_up(row_lo, 32), s1 = row_hi / 32 * 32;
        if (s1 > s0) {
            const uint64_t strips = (s1 - s0) / 32;
            const int E = blocked_strips(pl);
            const size_t smem = blocked_smem(pl->block_s, E);
            // enough CTAs to fill the chip several times over, else as many strips per CTA as
            // amortise the table load (10 KB at S=32 against ~17 KB of output per strip)
            uint32_t per_cta = 8;
            while (per_cta > (uint32_t)E && (strips + per_cta - 1) / per_cta * pl->n_blocks < (uint64_t)pl->n_sm * 16) per_cta /= 2;
            const uint64_t ctas = (strips + per_cta - 1) / per_cta * pl->n_blocks;
            if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_blocked: row window too large for one launch");
            auto kern = E == 2 ? qr::fill_blocked_kernel<2> : qr::fill_blocked_kernel<1>;
            QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int rc = launch_direct(pl, row_lo, s0, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            if (rc != QR_OK) return rc;
            kern<<<(unsigned)ctas, 32 * qr::FILL_BLOCKED_WARPS, smem, st>>>(
                pl->dev, (uint32_t)G, pl->block_s, pl->n_blocks, per_cta, s0, strips, row_lo, indptr_base,
                d_indptr, d_indices, d_data, row_hi - row_lo);
            QR_LAUNCH_CHECK("fill_blocked_kernel");
            return launch_direct(pl, s1, row_hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
        }
    }
    const StagedCfg *cfg = (flags & QR_FILL_DIRECT) || pl->rw == 0 ? nullptr : find_staged(pl->rw, pl->gw);
    if (cfg) {
        const uint64_t R = 32ull * cfg->rw;
        const uint64_t s0 = align_up(row_lo, R), s1 = row_hi / R * R;
        // Even G: lanes G*16 B apart share banks in the plain tile.  The swizzled tile leaves through 2-D tensor maps
        // (fill_staged_swz_kernel); it needs an even G (16-byte row pitch of the column ids) and the driver's encoder.
        bool swz = swz_default(G);
        if (const char *env = getenv("QR_FILL_SWZ")) swz = env[0] == '1';
        const size_t swz_smem = (size_t)(((G + 7) / 8 + (G + 15) / 16) * R * 128 + 1024);
        if (swz && s1 > s0 && G % 2 == 0 && swz_smem <= MAX_SMEM && row_hi - row_lo < (1ull << 31) && (s1 - s0) / R <= 0x7fffffffull) {
            qr::TensorMap tmd, tmi;
            SwzFn fn = find_swz(cfg->rw, cfg->gw);
            if (fn && make_tile_map(&tmd, d_data, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2 * G, row_hi - row_lo, (uint32_t)R) &&
                make_tile_map(&tmi, d_indices, CU_TENSOR_MAP_DATA_TYPE_UINT64, G, row_hi - row_lo, (uint32_t)R)) {
                QR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)swz_smem));
                int rc = launch_direct(pl, row_lo, s0, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
                if (rc != QR_OK) return rc;
                cudaLaunchConfig_t lc = {};
                cudaLaunchAttribute attr;
                lc.gridDim = dim3((unsigned)((s1 - s0) / R), 1, 1); lc.blockDim = dim3(32 * cfg->gw, 1, 1); lc.dynamicSmemBytes = swz_smem; lc.stream = st;
                if (s0 == row_lo) pdl_config(lc, attr);
                QR_CUDA(cudaLaunchKernelEx(&lc, fn, pl->dev, (uint32_t)G, s0, row_lo, indptr_base, d_indptr, (uint64_t)(row_hi - row_lo), tmd, tmi));
                QR_LAUNCH_CHECK("fill_staged_swz_kernel");
                return launch_direct(pl, s1, row_hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            }
        }
        // the bulk copies need 16-byte aligned global addresses: indices + (s0-row_lo)*G*8
        const bool aligned = (((s0 - row_lo) * G) & 1) == 0;
        if (s1 > s0 && aligned) {
            // G a multiple of 8: the lanes of a store (one row each, G*16 B apart) all hit the same banks
            // (8-way conflicts, 3.5 TB/s on XXZ n=23).  A 16-byte gap after every 4 rows leaves 2-way
            // conflicts at the price of R/4 copies per array instead of one: 4.65 TB/s.  For G = 4 mod 8
            // (4-way conflicts, 4.7-5.0 TB/s) every gap period measured slower than none, as it is for
            // every other G (profiles/r02_pad_probe.jsonl).
            uint32_t ps = 2u;
            bool pad = G % 8 == 0;
            if (const char *env = getenv("QR_FILL_PAD")) {             // "0": never; "1".."5": force this gap period
                const int v = atoi(env);
                pad = v >= 1 && v <= 5;
                if (pad) ps = (uint32_t)v;
            }
            size_t smem = (size_t)(R * G * 24);
            if (pad) smem += (size_t)(R >> ps) * 32;
            if (smem > MAX_SMEM) { pad = false; smem = (size_t)(R * G * 24); }
            const uint64_t tiles = (s1 - s0) / R;
            if (tiles > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "fill_staged: row window too large for one launch");
            StagedFn fn = pad ? cfg->fn_pad : cfg->fn;
            QR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // prefix / suffix rows that do not fill a tile
            int rc = launch_direct(pl, row_lo, s0, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
            if (rc != QR_OK) return rc;
            cudaLaunchConfig_t lc = {};
            cudaLaunchAttribute attr;
            lc.gridDim = dim3((unsigned)tiles, 1, 1); lc.blockDim = dim3(32 * cfg->gw, 1, 1); lc.dynamicSmemBytes = smem; lc.stream = st;
            if (s0 == row_lo) pdl_config(lc, attr);                  // no edge kernel between the canonicalisation and this launch
            QR_CUDA(cudaLaunchKernelEx(&lc, fn, pl->dev, (uint32_t)G, s0, row_lo, indptr_base, d_indptr, d_indices, d_data,
                                       (uint64_t)(row_hi - row_lo), ps));
            QR_LAUNCH_CHECK("fill_staged_kernel");
            lo = s1; hi = row_hi;
        }
    }
    return launch_direct(pl, lo, hi, row_lo, row_hi, indptr_base, d_indptr, d_indices, d_data, st);
}

}  // namespace

extern "C" int qr_build_rows_device(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, uint64_t *d_indptr,
                                    uint64_t *d_indices, double *d_data, uint32_t flags, void *stream)
{
    if (!pl || !d_indices || !d_data) return fail(QR_ERR_INVALID, "qr_build_rows_device: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_build_rows_device: bad row range");
    if (((uintptr_t)d_indices | (uintptr_t)d_data) & 15) return fail(QR_ERR_INVALID, "qr_build_rows_device: outputs must be 16-byte aligned");
    QR_CUDA(cudaSetDevice(pl->device));
    return build_rows(pl, row_lo, row_hi, d_indptr, d_indices, reinterpret_cast<double2 *>(d_data), flags, as_stream(stream));
}

namespace {

// Host threads for the work qr_build_host leaves to the CPU while the DMA engines run: moving staged windows into pageable
// destinations and rebuilding 64-bit columns from the compact wire form.  One pool per process, created on first use
// (spawning 16 threads costs about a millisecond -- too much to pay per export).
class HostPool {
public:
    explicit HostPool(unsigned n_threads) : n_(n_threads)
    {
        for (unsigned i = 1; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    unsigned threads() const { return n_; }
    // runs fn(0..n_tasks-1) on the pool (the calling thread included); returns when all are done
    void run(size_t n_tasks, const std::function<void(size_t)> &fn)
    {
        std::lock_guard<std::mutex> serial(run_m_);
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; n_tasks_ = n_tasks; next_.store(0); busy_ = (unsigned)workers_.size(); gen_++;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return busy_ == 0; });
    }
private:
    void drain()
    {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= n_tasks_) return;
            (*fn_)(i);
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            drain();
            { std::lock_guard<std::mutex> l(m_); if (--busy_ == 0) done_.notify_one(); }
        }
    }
    unsigned n_;
    std::vector<std::thread> workers_;
    std::mutex m_, run_m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_tasks_ = 0;
    std::atomic<size_t> next_{0};
    unsigned busy_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

HostPool &host_pool()
{
    static HostPool *pool = [] {
        unsigned n = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char *env = getenv("QR_HOST_COPY_THREADS")) { int t = atoi(env); if (t >= 1 && t <= 64) n = (unsigned)t; }
        return new HostPool(n);                                   // lives for the process: no static-destruction order to get wrong
    }();
    return *pool;
}

bool is_page_locked(const void *ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

thread_local uint64_t g_last_d2h_bytes = 0;

}  // namespace

extern "C" uint64_t qr_last_d2h_bytes(void) { return g_last_d2h_bytes; }

extern "C" int qr_build_host(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, uint64_t *indptr,
                             uint64_t *indices, double *data, uint32_t flags)
{
    if (!pl || !indices || !data) return fail(QR_ERR_INVALID, "qr_build_host: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_build_host: bad row range");
    QR_CUDA(cudaSetDevice(pl->device));
    const uint64_t G = pl->n_groups, rows = row_hi - row_lo;
    g_last_d2h_bytes = 0;
    // Row windows through a ring of WIN_RING device buffers on ONE stream: the copy engine always has the next windows queued
    // and they land in order (on separate streams the copies share the link and finish together: measured, the host then
    // sits idle for the first four windows and works through them while the link idles), so the host's share of window w
    // (below) runs while window w+1 crosses.
    // Wire form (wire.cuh): data as it is; the columns as group ids of 1 / 2 bytes (or the 32-bit column beyond 65536
    // groups), widened by the host pool while the next window is in flight; indptr is written by the host (r * G).
    // QR_HOST_WIDE: the 24-byte form of round 1, every array copied as stored (comparison, tests).
    // A page-locked `data` receives its windows directly; pageable destinations go through pinned staging.
    const bool no_staging = (flags & QR_HOST_NO_STAGING) != 0;
    const bool wide = (flags & QR_HOST_WIDE) != 0 || no_staging;
    const bool data_direct = no_staging || is_page_locked(data);
    const bool idx_direct = wide && (no_staging || is_page_locked(indices));
    uint32_t wcol = wide ? 8u : G <= 256 ? 1u : G <= 65536 ? 2u : 4u;            // bytes per column on the wire
    if (const char *env = getenv("QR_HOST_WIRE_COL")) { const int v = atoi(env); if (!wide && (v == 2 || v == 4) && (uint32_t)v > wcol) wcol = (uint32_t)v; }   // tests
    const uint64_t per_row = G * (16 + wcol);
    uint64_t win_mb = 64;
    if (const char *env = getenv("QR_HOST_WIN_MB")) { const int v = atoi(env); if (v >= 1 && v <= 1024) win_mb = (uint64_t)v; }
    uint64_t win_rows = (win_mb << 20) / per_row;
    if (win_rows >= rows) win_rows = rows;
    else { win_rows = win_rows / 256 * 256; if (win_rows == 0) win_rows = 32; }
    const size_t dat_bytes = align_up(win_rows * G * 16, 256), idx_bytes = align_up(win_rows * G * 8, 256);
    const size_t col_bytes = align_up(win_rows * G * (wide ? 0 : wcol), 256);
    const size_t need = dat_bytes + idx_bytes + col_bytes;
    // the staging windows and their streams belong to the device, not to the plan: a caller that
    // builds one matrix per plan (the reference's to_matrix call pattern) does not pay two
    // cudaMalloc/cudaFree of 256 MB per call
    WinScratch *wsp = nullptr;
    { std::lock_guard<std::mutex> lock(g_win_mutex); wsp = &g_win[pl->device]; }   // std::map nodes never move
    WinScratch &ws = *wsp;
    std::lock_guard<std::mutex> busy(ws.busy);                      // shards on different GPUs copy concurrently
    if (ws.bytes < need) {
        for (int i = 0; i < WIN_RING; i++) { if (ws.buf[i]) cudaFree(ws.buf[i]); ws.buf[i] = nullptr; }
        ws.bytes = 0;
        for (int i = 0; i < WIN_RING; i++) QR_CUDA(cudaMalloc(&ws.buf[i], need));
        ws.bytes = need;
    }
    if (!ws.stream[0]) QR_CUDA(cudaStreamCreateWithFlags(&ws.stream[0], cudaStreamNonBlocking));
    for (int i = 0; i < WIN_RING; i++)
        if (!ws.done[i]) QR_CUDA(cudaEventCreateWithFlags(&ws.done[i], cudaEventDisableTiming));
    // pinned staging for whatever does not go straight to its destination
    const size_t h_dat = data_direct ? 0 : dat_bytes, h_idx = wide ? (idx_direct ? 0 : idx_bytes) : col_bytes;
    const size_t host_need = h_dat + h_idx;
    if (host_need && ws.host_bytes < host_need) {
        for (int i = 0; i < WIN_RING; i++) { if (ws.host[i]) cudaFreeHost(ws.host[i]); ws.host[i] = nullptr; }
        ws.host_bytes = 0;
        for (int i = 0; i < WIN_RING; i++) QR_CUDA(cudaMallocHost(&ws.host[i], host_need));
        ws.host_bytes = host_need;
    }
    if (wcol < 8 && pl->host_gx.empty()) {
        pl->host_gx.resize(G);
        QR_CUDA(cudaMemcpy(pl->host_gx.data(), pl->dev.gx, G * 4, cudaMemcpyDeviceToHost));
    }
    const uint32_t *gx = pl->host_gx.data();
    HostPool &pool = host_pool();
    // window boundaries (cutting the last window into quarters so that less host work trails the final copy was measured:
    // no difference on C2, the copies land 1.2-1.4 ms apart and the host needs 0.5-0.7 ms per window)
    std::vector<uint64_t> wb;
    for (uint64_t r = 0; r < rows; r += win_rows) wb.push_back(r);
    wb.push_back(rows);
    const uint64_t n_win = wb.size() - 1;
    uint64_t d2h = 0;

    auto submit = [&](uint64_t w) -> int {                          // fill window w, start its copies
        const uint64_t w0 = row_lo + wb[w], w1 = row_lo + wb[w + 1], n = w1 - w0, o = (w0 - row_lo) * G;
        const int k = (int)(w % WIN_RING);
        cudaStream_t st = ws.stream[0];
        char *buf = static_cast<char *>(ws.buf[k]), *h = static_cast<char *>(ws.host[k]);
        double2 *dd = reinterpret_cast<double2 *>(buf);
        uint64_t *di = reinterpret_cast<uint64_t *>(buf + dat_bytes);
        int rc = build_rows(pl, w0, w1, nullptr, di, dd, 0, st);
        if (rc != QR_OK) return rc;
        QR_CUDA(cudaMemcpyAsync(data_direct ? reinterpret_cast<char *>(data) + o * 16 : h, dd, n * G * 16, cudaMemcpyDeviceToHost, st));
        if (wide) {
            QR_CUDA(cudaMemcpyAsync(idx_direct ? reinterpret_cast<char *>(indices + o) : h + h_dat, di, n * G * 8, cudaMemcpyDeviceToHost, st));
        } else {
            char *dc = buf + dat_bytes + idx_bytes;
            const uint64_t ne = n * G;
            const unsigned grid = (unsigned)std::min<uint64_t>((ne + 255) / 256, (uint64_t)pl->n_sm * 16);
            if (wcol == 1) qr::wire_columns_kernel<uint8_t><<<grid, 256, 0, st>>>(pl->dev, (uint32_t)G, w0, ne, di, reinterpret_cast<uint8_t *>(dc));
            else if (wcol == 2) qr::wire_columns_kernel<uint16_t><<<grid, 256, 0, st>>>(pl->dev, (uint32_t)G, w0, ne, di, reinterpret_cast<uint16_t *>(dc));
            else qr::wire_columns_kernel<uint32_t><<<grid, 256, 0, st>>>(pl->dev, (uint32_t)G, w0, ne, di, reinterpret_cast<uint32_t *>(dc));
            QR_LAUNCH_CHECK("wire_columns_kernel");
            QR_CUDA(cudaMemcpyAsync(h + h_dat, dc, ne * wcol, cudaMemcpyDeviceToHost, st));
        }
        d2h += n * G * (16 + wcol);
        QR_CUDA(cudaEventRecord(ws.done[k], st));
        return QR_OK;
    };
    // host side of window w once its copies have landed: widen the columns / move staged pieces to their destination
    auto finish = [&](uint64_t w) {
        const uint64_t w0 = row_lo + wb[w], n = wb[w + 1] - wb[w], o = (w0 - row_lo) * G;
        const char *h = static_cast<const char *>(ws.host[(int)(w % WIN_RING)]);
        const bool move_dat = !data_direct, move_idx = wide ? !idx_direct : true;
        if (!move_dat && !move_idx) return;
        const uint64_t rows_per_task = std::max<uint64_t>(1, (1u << 16) / G);
        const uint64_t n_tasks = (n + rows_per_task - 1) / rows_per_task;
        pool.run(n_tasks, [&](size_t t) {
            const uint64_t r0 = t * rows_per_task, r1 = std::min(n, r0 + rows_per_task), e0 = r0 * G, e1 = r1 * G;
            if (move_dat) memcpy(reinterpret_cast<char *>(data) + (o + e0) * 16, h + e0 * 16, (e1 - e0) * 16);
            if (!move_idx) return;
            // plain stores: streaming (non-temporal) ones were measured slower here -- they contend with the DMA engine's
            // writes of `data` in DRAM, where ordinary stores are absorbed by the last-level cache (9.4 against 7.6 ms on C2)
            uint64_t *out = indices + o;
            const char *src = h + h_dat;
            if (wide) { memcpy(out + e0, src + e0 * 8, (e1 - e0) * 8); return; }
            if (wcol == 4) {
                const uint32_t *c = reinterpret_cast<const uint32_t *>(src);
                for (uint64_t e = e0; e < e1; e++) out[e] = c[e];
            } else if (wcol == 2) {
                const uint16_t *c = reinterpret_cast<const uint16_t *>(src);
                for (uint64_t r = r0; r < r1; r++) {
                    const uint64_t row = w0 + r;
                    for (uint64_t e = r * G; e < (r + 1) * G; e++) out[e] = row ^ (uint64_t)gx[c[e]];
                }
            } else {
                const uint8_t *c = reinterpret_cast<const uint8_t *>(src);
                for (uint64_t r = r0; r < r1; r++) {
                    const uint64_t row = w0 + r;
                    for (uint64_t e = r * G; e < (r + 1) * G; e++) out[e] = row ^ (uint64_t)gx[c[e]];
                }
            }
        });
    };

    const bool trace = getenv("QR_HOST_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    int rc = QR_OK;
    uint64_t submitted = 0;
    for (uint64_t w = 0; w < n_win && rc == QR_OK; w++) {
        const double t0 = now();
        // windows w .. w + WIN_RING - 1 in flight: the buffers of every earlier window are free (finish() has returned)
        while (rc == QR_OK && submitted < n_win && submitted < w + WIN_RING) rc = submit(submitted++);
        if (rc != QR_OK) break;
        if (w == 0 && indptr) {                                     // affine: never crosses the link; written while the first windows fly
            const uint64_t base = (flags & QR_INDPTR_GLOBAL) ? row_lo * G : 0;
            const uint64_t per = 1u << 16, n_tasks = (rows + 1 + per - 1) / per;
            pool.run(n_tasks, [&](size_t t) {
                const uint64_t i1 = std::min(rows + 1, (t + 1) * per);
                for (uint64_t i = t * per; i < i1; i++) indptr[i] = base + i * G;
            });
        }
        const double t1 = now();
        QR_CUDA(cudaEventSynchronize(ws.done[(int)(w % WIN_RING)]));
        const double t2 = now();
        finish(w);
        if (trace) fprintf(stderr, "qr_build_host window %llu/%llu: at %.3f ms submit %.3f wait %.3f finish %.3f\n", (unsigned long long)w,
                           (unsigned long long)n_win, t0 - t_begin, t1 - t0, t2 - t1, now() - t2);
    }
    cudaStreamSynchronize(ws.stream[0]);
    if (rc != QR_OK) return rc;
    g_last_d2h_bytes = d2h;
    return QR_OK;
}

// rawio::write (qrusty/src/rawio.rs:128-148) streamed from the GPU: the file is
//   "MI" (native-endian u16 mark, rawio.rs:59-68) | u64 storage (0 = CSR) | u64 rows | u64 cols |
//   u64 len, indptr u64[] | u64 len, indices u64[] | u64 len, data complex128[]
// Every section's offset is known up front (nnz = G * rows), so row windows are filled on the
// device, copied to pinned staging on two streams and written with pwrite at their final offsets;
// the matrix is never resident as a whole on either side.
extern "C" int qr_write_rawio(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, const char *path)
{
    if (!pl || !path) return fail(QR_ERR_INVALID, "qr_write_rawio: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_write_rawio: bad row range");
    QR_CUDA(cudaSetDevice(pl->device));
    const uint64_t G = pl->n_groups, rows = row_hi - row_lo, nnz = rows * G;
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(QR_ERR_INVALID, std::string("qr_write_rawio: cannot open ") + path + ": " + strerror(errno));
    auto put = [&](const void *buf, size_t bytes, uint64_t off) {
        const char *b = static_cast<const char *>(buf);
        while (bytes) {
            const ssize_t w = pwrite(fd, b, bytes, (off_t)off);
            if (w <= 0) return false;
            b += w; off += (uint64_t)w; bytes -= (size_t)w;
        }
        return true;
    };
    const unsigned char mark_bytes[2] = {'M', 'I'};                 // u16::from_ne_bytes(['M','I']).to_ne_bytes()
    const uint64_t o_ptr = 2 + 4 * 8, o_idx = o_ptr + (rows + 1) * 8 + 8, o_dat = o_idx + nnz * 8 + 8;
    const uint64_t head[4] = {0 /* CSR */, rows, pl->dim, rows + 1};
    bool ok = put(mark_bytes, 2, 0) && put(head, sizeof(head), 2);
    ok = ok && put(&nnz, 8, o_idx - 8) && put(&nnz, 8, o_dat - 8);

    uint64_t win = std::max<uint64_t>(32, (64ull << 20) / (G * 24 + 8) / 32 * 32);
    if (win > rows) win = rows;
    const size_t dat_b = align_up(win * G * 16, 256), idx_b = align_up(win * G * 8, 256), ptr_b = align_up((win + 1) * 8, 256);
    void *dbuf[2] = {nullptr, nullptr}, *hbuf[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMalloc(&dbuf[i], dat_b + idx_b + ptr_b);
        if (e == cudaSuccess) e = cudaMallocHost(&hbuf[i], dat_b + idx_b + ptr_b);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    }
    int rc = e == cudaSuccess ? QR_OK : fail(QR_ERR_OOM, std::string("qr_write_rawio: ") + cudaGetErrorString(e));
    struct Win { uint64_t w0 = 0, n = 0; bool live = false; } pending[2];
    auto flush = [&](int k) {                                       // window k: wait for its copy, write it out
        if (!pending[k].live) return true;
        pending[k].live = false;
        if (cudaStreamSynchronize(st[k]) != cudaSuccess) return false;
        char *h = static_cast<char *>(hbuf[k]);
        const uint64_t w0 = pending[k].w0, n = pending[k].n, o = (w0 - row_lo) * G;
        const bool last = w0 + n == row_hi;
        if (row_lo) {                                               // the file's indptr starts at 0
            uint64_t *ip = reinterpret_cast<uint64_t *>(h + dat_b + idx_b);
            for (uint64_t i = 0; i <= n; i++) ip[i] -= row_lo * G;
        }
        return put(h, n * G * 16, o_dat + o * 16) && put(h + dat_b, n * G * 8, o_idx + o * 8) &&
               put(h + dat_b + idx_b, (n + (last ? 1 : 0)) * 8, o_ptr + (w0 - row_lo) * 8);
    };
    int k = 0;
    for (uint64_t w0 = row_lo; w0 < row_hi && rc == QR_OK && ok; w0 += win, k ^= 1) {
        ok = flush(k);
        if (!ok) break;
        const uint64_t n = std::min(win, row_hi - w0);
        char *d = static_cast<char *>(dbuf[k]);
        rc = build_rows(pl, w0, w0 + n, reinterpret_cast<uint64_t *>(d + dat_b + idx_b), reinterpret_cast<uint64_t *>(d + dat_b),
                        reinterpret_cast<double2 *>(d), QR_INDPTR_GLOBAL, st[k]);
        if (rc != QR_OK) break;
        e = cudaMemcpyAsync(hbuf[k], dbuf[k], dat_b + idx_b + (n + 1) * 8, cudaMemcpyDeviceToHost, st[k]);
        if (e != cudaSuccess) { rc = fail(QR_ERR_CUDA, std::string("qr_write_rawio: D2H: ") + cudaGetErrorString(e)); break; }
        pending[k].w0 = w0; pending[k].n = n; pending[k].live = true;
    }
    // the windows still in flight, in submission order
    for (int j = 0; j < 2 && rc == QR_OK && ok; j++) { ok = flush(k); k ^= 1; }
    for (int i = 0; i < 2; i++) {
        if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
        if (dbuf[i]) cudaFree(dbuf[i]);
        if (hbuf[i]) cudaFreeHost(hbuf[i]);
    }
    if (close(fd) != 0) ok = false;
    if (rc != QR_OK) return rc;
    if (!ok) return fail(QR_ERR_INVALID, std::string("qr_write_rawio: write to ") + path + " failed: " + strerror(errno));
    return QR_OK;
}

extern "C" int qr_release_scratch(void)
{
    std::lock_guard<std::mutex> lock(g_win_mutex);
    for (auto &kv : g_win) {
        if (cudaSetDevice(kv.first) != cudaSuccess) continue;
        for (int i = 0; i < WIN_RING; i++) {
            if (kv.second.buf[i]) cudaFree(kv.second.buf[i]);
            if (kv.second.host[i]) cudaFreeHost(kv.second.host[i]);
            if (kv.second.stream[i]) cudaStreamDestroy(kv.second.stream[i]);
            if (kv.second.done[i]) cudaEventDestroy(kv.second.done[i]);
        }
    }
    g_win.clear();
    std::lock_guard<std::mutex> lock2(g_dev_mutex);
    for (auto &kv : g_dev) {
        if (cudaSetDevice(kv.first) != cudaSuccess) continue;
        if (kv.second.tiles) cudaFree(kv.second.tiles);
        if (kv.second.partials) cudaFree(kv.second.partials);
        if (kv.second.cwin) cudaFree(kv.second.cwin);
    }
    g_dev.clear();
    return QR_OK;
}

// =====================================================================================
// H.v and friends
// =====================================================================================
constexpr int APPLY_THREADS_V1 = 512;

static int apply_tile_bits()
{
    if (const char *env = getenv("QR_APPLY_K")) { int k = atoi(env); if (k >= 11 && k <= 13) return k; }
    return 12;                                   // tile = 2^12 elements = 64 KB of shared memory
}

// diag(H) of rows [row_lo,row_hi), computed once and reused by every later apply on that range
// (the mask-0 group is by far the most expensive one: all Z-only strings land in it).
static int ensure_diag_cache(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, cudaStream_t st, const double2 **out,
                             const double **out_re = nullptr)
{
    *out = nullptr;
    if (out_re) *out_re = nullptr;
    const char *off = getenv("QR_APPLY_NO_DIAG_CACHE");
    if (off && off[0] == '1') return QR_OK;
    if (pl->diag_terms < 0) {
        uint32_t x0 = 1, goff1 = 0, flag0 = 0;
        QR_CUDA(cudaMemcpy(&x0, pl->dev.gx, 4, cudaMemcpyDeviceToHost));
        QR_CUDA(cudaMemcpy(&goff1, pl->dev.goff + 1, 4, cudaMemcpyDeviceToHost));
        QR_CUDA(cudaMemcpy(&flag0, pl->dev.gflag, 4, cudaMemcpyDeviceToHost));
        pl->diag_terms = x0 == 0 ? (int)goff1 : 0;
        pl->diag_is_real = (flag0 & 2u) ? 1 : 0;
    }
    if (pl->diag_terms < 3) return QR_OK;        // cheaper to recompute than to read 16 B/row
    // callers that can take it get the real parts only (8 B per row instead of 16) when the group is real
    const char *re_off = getenv("QR_APPLY_DIAG_REAL");
    const bool want_real = out_re != nullptr && pl->diag_is_real == 1 && !(re_off && re_off[0] == '0');
    if (!pl->diag_cache || pl->diag_lo != row_lo || pl->diag_hi != row_hi || pl->diag_cache_real != want_real) {
        if (pl->diag_cache) { QR_CUDA(cudaFree(pl->diag_cache)); pl->diag_cache = nullptr; }
        QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&pl->diag_cache), (row_hi - row_lo) * (want_real ? 8 : 16)));
        const uint64_t ctas = (row_hi - row_lo + 255) / 256;
        qr::diagonal_kernel<<<(unsigned)ctas, 256, 0, st>>>(pl->dev, row_lo, row_hi, want_real ? nullptr : pl->diag_cache,
                                                            want_real ? reinterpret_cast<double *>(pl->diag_cache) : nullptr);
        QR_LAUNCH_CHECK("diagonal_kernel");
        QR_CUDA(cudaStreamSynchronize(st));      // one-time; later applies may use any stream
        pl->diag_lo = row_lo; pl->diag_hi = row_hi; pl->diag_cache_real = want_real;
    }
    if (want_real) *out_re = reinterpret_cast<const double *>(pl->diag_cache);
    else *out = pl->diag_cache;
    return QR_OK;
}

// Greedy cover of the groups by passes (apply.cuh).  m = log2(rows of the local block).
static int make_apply_plan(qr_plan *pl, int m, int K, ApplyPlan **out)
{
    auto it = pl->apply_plans.find(m * 100 + K);
    if (it != pl->apply_plans.end()) { *out = &it->second; return QR_OK; }
    const uint32_t G = (uint32_t)pl->n_groups;
    if (pl->host_gx.empty()) {
        pl->host_gx.resize(G);
        QR_CUDA(cudaMemcpy(pl->host_gx.data(), pl->dev.gx, G * 4, cudaMemcpyDeviceToHost));
    }
    const std::vector<uint32_t> &gx = pl->host_gx;
    const int HB = K - 5;
    const uint32_t local_mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
    struct HostPass { uint32_t smask; std::vector<uint32_t> groups, direct; };
    std::vector<HostPass> hp;
    std::vector<char> done(G, 0);
    std::vector<uint32_t> direct;
    // groups that can never sit in a tile: bits outside the local block, or too many high bits
    for (uint32_t g = 0; g < G; g++)
        if ((gx[g] & ~local_mask) || __builtin_popcount(gx[g] & ~31u) > HB) { direct.push_back(g); done[g] = 1; }
    auto fill_to_k = [&](uint32_t smask) {          // pad S with the lowest unused local bits
        for (int b = 5; b < m && __builtin_popcount(smask) < K; b++) smask |= 1u << b;
        return smask;
    };
    bool first = true;
    for (;;) {
        uint32_t smask = 31u;
        if (first) smask = fill_to_k(31u);           // pass 0: the contiguous low tile
        else {
            // order the uncovered groups by their highest bit; add greedily while S has room
            std::vector<uint32_t> order;
            for (uint32_t g = 0; g < G; g++) if (!done[g]) order.push_back(g);
            if (order.empty()) break;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                const int ha = 31 - __builtin_clz(gx[a] | 1u), hb = 31 - __builtin_clz(gx[b] | 1u);
                return ha != hb ? ha < hb : gx[a] < gx[b];
            });
            for (uint32_t g : order) {
                const uint32_t cand = smask | gx[g];
                if (__builtin_popcount(cand & ~31u) <= HB) smask = cand;
            }
            smask = fill_to_k(smask);
        }
        HostPass ps; ps.smask = smask;
        for (uint32_t g = 0; g < G; g++)
            if (!done[g] && (gx[g] & ~smask) == 0) { ps.groups.push_back(g); done[g] = 1; }
        if (first) ps.direct = direct;
        if (!first && ps.groups.empty()) break;      // cannot happen: every group left fits some S
        hp.push_back(ps);
        first = false;
        bool any = false;
        for (uint32_t g = 0; g < G; g++) any |= !done[g];
        if (!any) break;
    }
    // device tables
    size_t words = 0;
    for (auto &ps : hp) words += 2 * ps.groups.size() + ps.direct.size() + (1u << HB);
    ApplyPlan ap;
    std::vector<uint32_t> host(words ? words : 1);
    QR_CUDA(cudaMalloc(&ap.slab, host.size() * 4));
    uint32_t *dbase = static_cast<uint32_t *>(ap.slab);
    size_t off = 0;
    for (size_t q = 0; q < hp.size(); q++) {
        const HostPass &ps = hp[q];
        std::vector<int> sbits;
        for (int b = 0; b < 32; b++) if (ps.smask >> b & 1u) sbits.push_back(b);
        auto compact = [&](uint32_t x) { uint32_t c = 0; for (size_t j = 0; j < sbits.size(); j++) if (x >> sbits[j] & 1u) c |= 1u << j; return c; };
        qr::ApplyPass dp{};
        dp.groups = dbase + off; for (uint32_t g : ps.groups) host[off++] = g;
        dp.cmask = dbase + off;  for (uint32_t g : ps.groups) host[off++] = compact(gx[g]);
        dp.n_groups = (uint32_t)ps.groups.size();
        dp.direct = dbase + off; for (uint32_t g : ps.direct) host[off++] = g;
        dp.n_direct = (uint32_t)ps.direct.size();
        dp.expand = dbase + off;
        for (uint32_t j = 0; j < (1u << HB); j++) {            // tile index bits >= 5 -> row bits
            uint32_t rbits = 0;
            for (size_t t = 5; t < sbits.size(); t++) if (j >> (t - 5) & 1u) rbits |= 1u << sbits[t];
            host[off++] = rbits;
        }
        dp.free_mask = local_mask & ~ps.smask;
        dp.first = q == 0 ? 1u : 0u;
        ap.passes.push_back(dp);
    }
    cudaError_t e = cudaMemcpy(ap.slab, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(ap.slab); return fail(QR_ERR_CUDA, std::string("make_apply_plan: ") + cudaGetErrorString(e)); }
    *out = &(pl->apply_plans[m * 100 + K] = ap);
    return QR_OK;
}

template <int K>
static int launch_apply_passes(qr_plan *pl, ApplyPlan *ap, uint64_t row_lo, uint64_t rows, const double2 *v,
                               double2 *y, const double2 *diag, cudaStream_t st)
{
    auto kern = qr::apply_pass_kernel<K, APPLY_THREADS_V1>;
    const size_t smem = ((size_t)16 << K) + ((size_t)4 << (K - 5)) + qr::APPLY_BATCH * sizeof(qr::GroupDesc);
    QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t tiles = rows >> K;
    for (qr::ApplyPass ps : ap->passes) {
        if (ps.first && diag != nullptr && ps.n_groups > 0) {
            // group 0 (mask 0) is first in pass 0's list: served from the cache instead
            ps.groups += 1; ps.cmask += 1; ps.n_groups -= 1; ps.diag = diag;
        }
        kern<<<(unsigned)tiles, APPLY_THREADS_V1, smem, st>>>(pl->dev, ps, row_lo, v, y);
        QR_LAUNCH_CHECK("apply_pass_kernel");
    }
    return QR_OK;
}


// ---- tiled apply (apply_tile.cuh): plan ------------------------------------------------------------
static int host_masks(qr_plan *pl)
{
    if (pl->host_gx.empty()) {
        pl->host_gx.resize(pl->n_groups);
        QR_CUDA(cudaMemcpy(pl->host_gx.data(), pl->dev.gx, pl->n_groups * 4, cudaMemcpyDeviceToHost));
    }
    return QR_OK;
}

constexpr size_t TILE_SMEM_CAP = MAX_SMEM - 2048;
// run kernel shape: threads per CTA, rows per thread and chunk, CTAs per SM (registers: 65536 / (threads * CTAs))
struct RunShape { int threads, e, minb; };
static RunShape run_shape(uint32_t R)
{
    RunShape s = R >= 2048 ? RunShape{512, 4, 1} : RunShape{256, 4, 3};
    if (const char *env = getenv("QR_APPLY_RUN_SHAPE")) {                 // "threads,e,ctas" (sweeps)
        int t = 0, e = 0, b = 0;
        if (sscanf(env, "%d,%d,%d", &t, &e, &b) == 3 && (uint32_t)(t * e) <= R) s = RunShape{t, e, b};
    }
    return s;
}


// What a tile kernel launch does with the groups of a block of 2^m rows:
//   TILE_FAR    rows cut at bit d: a tile = one run in each of the 2^(m-d) slices; groups whose mask has a bit >= d (and
//               none between the run and d) come from shared memory, y += ...; nothing is gathered  (second pass)
//   TILE_LOCAL  a tile = one contiguous run of 2^lr rows; groups whose mask lies inside the run come from shared memory,
//               every other group that is not `excluded` is gathered through L1/L2 by the same kernel  (first / only pass)
//   TILE_P2P    TILE_LOCAL on a rank's shard, plus the runs of the peers' shards pulled over NVLink by the TMA
enum TileRole { TILE_FAR = 0, TILE_LOCAL = 1, TILE_P2P = 2 };

static int make_tile_plan(qr_plan *pl, uint32_t m, uint32_t blk, uint32_t d, TileRole role, bool skip_diag,
                          const TilePlan *excluded, TilePlan **out)
{
    const uint64_t key = (uint64_t)m | ((uint64_t)d << 8) | ((uint64_t)role << 16) | ((uint64_t)(skip_diag ? 1 : 0) << 20) |
                         ((uint64_t)(excluded ? excluded->d : 0) << 24) | ((uint64_t)blk << 32);
    auto it = pl->tile_plans.find(key);
    if (it != pl->tile_plans.end()) { *out = &it->second; return QR_OK; }
    int rc = host_masks(pl);
    if (rc != QR_OK) return rc;
    const std::vector<uint32_t> &gx = pl->host_gx;
    const uint32_t G = (uint32_t)pl->n_groups;
    TilePlan tp;
    tp.m = m; tp.d = d;
    const uint32_t nrs = 1u << (m - d);
    auto hi_of = [&](uint32_t x) { return (uint32_t)((uint64_t)x >> d); };          // d may be 32
    auto blk_of = [&](uint32_t x) { return (uint32_t)((uint64_t)x >> m); };
    std::vector<char> skip(G, 0);
    if (excluded) for (uint32_t g : excluded->far_host) skip[g] = 1;
    if (skip_diag && gx[0] == 0u) skip[0] = 1;
    std::vector<uint32_t> far, near;
    std::vector<uint8_t> part;
    int lr_hi = (int)std::min<uint32_t>(11, d);
    if (const char *env = getenv("QR_APPLY_LR")) { int v = atoi(env); if (v >= 5 && v <= 13 && role != TILE_FAR) lr_hi = std::min<int>(v, (int)d); }
    if (nrs <= (uint32_t)qr::TILE_MAX_ROWSLOTS && (role == TILE_FAR || nrs == 1)) {
        for (int lr = lr_hi; lr >= 5 && !tp.ok; lr--) {
            const uint32_t R = 1u << lr;
            if ((uint64_t)nrs * R < (uint64_t)(role == TILE_FAR ? qr::TILE_THREADS : 1024)) break;   // a tile is at least one chunk of rows
            std::vector<TilePlan::Load> loads;
            // the row slots themselves -- except in the fused distributed apply, where the local groups are gathered through
            // L1/L2 (measured: the gather kernel beats shared-memory tiles for local partners) and only peers' runs are chunks
            const bool own = role != TILE_P2P || getenv("QR_P2P_OWN_RUN") != nullptr;
            if (own) for (uint32_t l = 0; l < nrs; l++) loads.push_back({blk, l, 0u});
            far.clear(); part.clear(); near.clear();
            std::vector<char> is_far(G, 0);
            for (uint32_t g = 0; g < G; g++) {
                if (skip[g] || far.size() >= (size_t)qr::TILE_MAX_FAR) continue;
                const uint32_t x = gx[g], xb = blk_of(x), xl = hi_of(x) & (nrs - 1u);
                const uint32_t xi = (uint32_t)(((uint64_t)x & ((1ull << d) - 1ull)) >> lr);
                if (role == TILE_FAR && hi_of(x) == 0) continue;                    // the first pass's group
                if (role == TILE_FAR && xi != 0) continue;                          // a bit between the run and the cut: gathered by the first pass
                if (role == TILE_LOCAL && (xb != 0 || xi != 0)) continue;           // outside the run: gathered
                if (role == TILE_P2P && xb == 0 && (xi != 0 || !own)) continue;     // local: gathered (always, unless the own run is a chunk)
                const size_t keep = loads.size();
                std::vector<uint8_t> pr(nrs);
                bool ok = true;
                for (uint32_t l = 0; l < nrs && ok; l++) {
                    const TilePlan::Load want{blk ^ xb, l ^ xl, xi};
                    size_t j = 0;
                    for (; j < loads.size(); j++)
                        if (loads[j].block == want.block && loads[j].lambda == want.lambda && loads[j].xi == want.xi) break;
                    if (j == loads.size()) {
                        if (loads.size() >= (size_t)qr::TILE_MAX_LOAD) { ok = false; break; }
                        loads.push_back(want);
                    }
                    pr[l] = (uint8_t)j;
                }
                // a remote run costs a chunk of its own: stop taking them when two stages no longer fit (the rest is gathered)
                if (ok && loads.size() > keep && loads.size() * (size_t)R * 16 + 16384 > (role == TILE_FAR ? TILE_SMEM_CAP / 2 : (TILE_SMEM_CAP - 2048) / (size_t)run_shape(R).minb - 1024)) ok = false;
                if (!ok) { loads.resize(keep); continue; }
                far.push_back(g); is_far[g] = 1;
                part.insert(part.end(), pr.begin(), pr.end());
            }
            if (far.empty()) continue;
            if (role != TILE_FAR) {
                for (uint32_t g = 0; g < G; g++) if (!skip[g] && !is_far[g]) near.push_back(g);
                if (near.size() > (size_t)qr::TILE_MAX_NEAR) break;
            }
            const size_t fixed = 64 + far.size() * sizeof(qr::GroupDesc) + near.size() * (sizeof(qr::GroupDesc) + 8) + far.size() * std::max<size_t>(nrs, 4) + 64;
            const size_t stage_bytes = loads.size() * (size_t)R * 16;
            size_t stages;
            if (role == TILE_FAR) {
                if (fixed + 2 * stage_bytes > TILE_SMEM_CAP) continue;              // try shorter runs
                stages = std::min<size_t>(3, (TILE_SMEM_CAP - fixed) / stage_bytes);
                if (const char *env = getenv("QR_APPLY_TILE_STAGES")) { int v = atoi(env); if (v >= 1 && (size_t)v <= stages) stages = (size_t)v; }
            } else {
                // run kernel: one tile per CTA, two CTAs (512 threads) or four (256) per SM must fit
                const size_t per_cta = (TILE_SMEM_CAP - 2048) / (size_t)run_shape(R).minb - 1024;
                if (fixed + stage_bytes > per_cta) continue;
                stages = 1;
                if (const char *env = getenv("QR_APPLY_RUN_STAGES")) { int v = atoi(env); if (v >= 1 && v <= 3 && fixed + (size_t)v * stage_bytes <= per_cta) stages = (size_t)v; }
            }
            const uint64_t tile_rows = (uint64_t)nrs * R;
            tp.ok = true; tp.log2_run = (uint32_t)lr; tp.n_row_slots = nrs; tp.stages = (uint32_t)stages;
            tp.n_far = (uint32_t)far.size(); tp.n_near = (uint32_t)near.size();
            tp.e = tile_rows % (qr::TILE_THREADS * 4) == 0 ? 4 : tile_rows % (qr::TILE_THREADS * 2) == 0 ? 2 : 1;
            tp.loads = loads;
            tp.own_loaded = own;
            tp.far_host = far;
            tp.smem = fixed + stages * stage_bytes;
        }
    }
    if (tp.ok) {
        for (uint32_t g = 0; g < G; g++) if (blk_of(gx[g]) != 0) tp.need_mask |= 1u << ((blk ^ blk_of(gx[g])) & 31u);
        const size_t b_far = align_up(far.size() * 4, 16), b_near = align_up(near.size() * 4 + 4, 16), b_part = align_up(part.size(), 16);
        std::vector<unsigned char> host(b_far + b_near + b_part, 0);
        memcpy(host.data(), far.data(), far.size() * 4);
        memcpy(host.data() + b_far, near.data(), near.size() * 4);
        memcpy(host.data() + b_far + b_near, part.data(), part.size());
        QR_CUDA(cudaMalloc(&tp.slab, host.size()));
        cudaError_t e = cudaMemcpy(tp.slab, host.data(), host.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(tp.slab); return fail(QR_ERR_CUDA, std::string("make_tile_plan: ") + cudaGetErrorString(e)); }
        unsigned char *b = static_cast<unsigned char *>(tp.slab);
        tp.d_far = reinterpret_cast<const uint32_t *>(b);
        tp.d_near = reinterpret_cast<const uint32_t *>(b + b_far);
        tp.d_part = b + b_far + b_near;
    }
    *out = &(pl->tile_plans[key] = tp);
    return QR_OK;
}

// fills the per-call part of the kernel arguments; block_ptr(q) = element 0 of block q of v
template <typename BlockPtr>
static void tile_args(const qr_plan *pl, const TilePlan &tp, uint32_t blk, BlockPtr block_ptr, qr::ApplyTileArgs &a)
{
    a = qr::ApplyTileArgs{};
    for (size_t j = 0; j < tp.loads.size(); j++) {
        a.base[j] = block_ptr(tp.loads[j].block) + ((uint64_t)tp.loads[j].lambda << tp.d);
        a.xi[j] = tp.loads[j].xi;
    }
    for (uint32_t l = 0; l < tp.n_row_slots; l++) a.row0[l] = (uint32_t)(((uint64_t)blk << tp.m) + ((uint64_t)l << tp.d));
    a.n_load = (uint32_t)tp.loads.size(); a.n_row_slots = tp.n_row_slots; a.log2_run = tp.log2_run;
    a.n_tiles = (uint32_t)((1ull << tp.d) >> tp.log2_run); a.stages = tp.stages;
    a.n_far = tp.n_far; a.n_near = tp.n_near; a.block_bits = tp.m; a.my_block = blk;
    a.far_groups = tp.d_far; a.far_part = tp.d_part; a.near_groups = tp.d_near;
    a.own_loaded = tp.own_loaded ? 1u : 0u;
    (void)pl;
}

template <bool NEAR>
static int launch_tile(qr_plan *pl, const TilePlan &tp, const qr::ApplyTileArgs &a, uint64_t row_lo, double2 *y,
                       const double2 *diag, const double *diag_re, cudaStream_t st)
{
    using Fn = void (*)(qr::PlanDev, const qr::ApplyTileArgs, uint64_t, double2 *, const double2 *, const double *);
    Fn kern = tp.e == 4 ? (Fn)qr::apply_tile_kernel<NEAR, 4> : tp.e == 2 ? (Fn)qr::apply_tile_kernel<NEAR, 2> : (Fn)qr::apply_tile_kernel<NEAR, 1>;
    QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(a.n_tiles, (uint64_t)pl->n_sm);
    kern<<<grid, qr::TILE_THREADS, tp.smem, st>>>(pl->dev, a, row_lo, y, diag, diag_re);
    QR_LAUNCH_CHECK("apply_tile_kernel");
    return QR_OK;
}

static int launch_run(qr_plan *pl, const TilePlan &tp, const qr::ApplyTileArgs &a, uint64_t row_lo, double2 *y,
                      const double2 *diag, const double *diag_re, cudaStream_t st)
{
    using Fn = void (*)(qr::PlanDev, const qr::ApplyTileArgs, uint64_t, double2 *, const double2 *, const double *);
    const RunShape sh = run_shape(1u << tp.log2_run);
    Fn kern = nullptr;
#define QR_RUN(T_, E_, B_) if (sh.threads == T_ && sh.e == E_ && sh.minb == B_) kern = (Fn)qr::apply_run_kernel<T_, E_, B_>;
    QR_RUN(256, 4, 3) QR_RUN(256, 4, 2) QR_RUN(256, 2, 4) QR_RUN(256, 4, 4) QR_RUN(512, 4, 1) QR_RUN(512, 2, 2) QR_RUN(512, 4, 2) QR_RUN(1024, 2, 1) QR_RUN(1024, 1, 1)
#undef QR_RUN
    if (!kern) return fail(QR_ERR_INVALID, "apply_run: no kernel instance for QR_APPLY_RUN_SHAPE");
    QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
    // one tile per CTA unless the plan holds a ring of stages (then one persistent CTA per resident slot)
    const uint64_t resident = (uint64_t)pl->n_sm * sh.minb;
    const unsigned grid = (unsigned)(tp.stages > 1 ? std::min<uint64_t>(a.n_tiles, resident) : a.n_tiles);
    kern<<<grid, sh.threads, tp.smem, st>>>(pl->dev, a, row_lo, y, diag, diag_re);
    QR_LAUNCH_CHECK("apply_run_kernel");
    return QR_OK;
}

// bit at which the two-pass apply cuts the rows: masks below it are gathered through L2 (the window they span,
// 2^d * 16 B, must stay resident beside the streaming y and diag), masks at or above it go through shared memory
static uint32_t apply_cut_bit()
{
    if (const char *env = getenv("QR_APPLY_D")) { int v = atoi(env); if (v == 0 || (v >= 10 && v <= 32)) return (uint32_t)v; }
    return 21;
}


// ---- term-rich H.v (apply_fold.cuh): tables and kernel choice --------------------------------------
// The fold kernel pays when the groups a row has to evaluate carry three or more terms on average (molecular
// Hamiltonians: 5-6); spin chains and lattices (1-2 terms per off-diagonal group, diag(H) cached) stay with the gather
// kernel.  QR_APPLY_FOLD=0 / 1 overrides (tests, A/B).  Tables: the plan's (z, c') bucketed per group by the three row
// bits a thread of the kernel owns -- built on the host once per plan (T log T, T <= a few 10^4).
static int ensure_fold_tables(qr_plan *pl, uint64_t *t_eval_out, uint64_t *g_eval_out)
{
    const size_t G = pl->n_groups, T = pl->n_terms_canonical;
    std::vector<uint32_t> goff(G + 1), gflag(G), gx0(1);
    QR_CUDA(cudaMemcpy(goff.data(), pl->dev.goff, (G + 1) * 4, cudaMemcpyDeviceToHost));
    QR_CUDA(cudaMemcpy(gflag.data(), pl->dev.gflag, G * 4, cudaMemcpyDeviceToHost));
    QR_CUDA(cudaMemcpy(gx0.data(), pl->dev.gx, 4, cudaMemcpyDeviceToHost));
    uint64_t t_eval = 0, g_eval = 0;
    for (size_t g = 0; g < G; g++) {
        if (gflag[g] & 1u) continue;                              // row-independent: gconst
        if (g == 0 && gx0[0] == 0 && goff[1] >= 3) continue;       // diag(H): cached (ensure_diag_cache)
        t_eval += goff[g + 1] - goff[g]; g_eval++;
    }
    if (t_eval_out) *t_eval_out = t_eval;
    if (g_eval_out) *g_eval_out = g_eval;
    if (pl->fold_slab || pl->fold_blocked) return QR_OK;
    for (size_t g = 0; g < G; g++)
        if (goff[g + 1] - goff[g] > qr::FOLD_MAX_GROUP_TERMS) { pl->fold_blocked = true; pl->fold_state = 0; return QR_OK; }   // u16 bucket table: gather kernel
    std::vector<uint32_t> tz(T);
    std::vector<double2> tc(T);
    QR_CUDA(cudaMemcpy(tz.data(), pl->dev.tz, T * 4, cudaMemcpyDeviceToHost));
    QR_CUDA(cudaMemcpy(tc.data(), pl->dev.tc, T * 16, cudaMemcpyDeviceToHost));
    const size_t o_im = T * 16, o_be = align_up(o_im + T * 8, 16), bytes = o_be + G * 16;
    std::vector<unsigned char> host(bytes ? bytes : 16, 0);
    uint32_t *zc = reinterpret_cast<uint32_t *>(host.data());
    double *im = reinterpret_cast<double *>(host.data() + o_im);
    uint16_t *be = reinterpret_cast<uint16_t *>(host.data() + o_be);
    for (size_t g = 0; g < G; g++) {
        const uint32_t t0 = goff[g], t1 = goff[g + 1];
        uint32_t cnt[9] = {0};
        for (uint32_t t = t0; t < t1; t++) cnt[((tz[t] >> qr::FOLD_B0) & 7u) + 1]++;
        for (int q = 0; q < 8; q++) cnt[q + 1] += cnt[q];
        for (int q = 0; q < 8; q++) be[g * 8 + q] = (uint16_t)cnt[q + 1];
        uint32_t pos[8];
        for (int q = 0; q < 8; q++) pos[q] = t0 + cnt[q];
        for (uint32_t t = t0; t < t1; t++) {                       // stable: original order inside a bucket
            const uint32_t d = pos[(tz[t] >> qr::FOLD_B0) & 7u]++;
            uint64_t re_bits; memcpy(&re_bits, &tc[t].x, 8);
            zc[4 * (size_t)d] = tz[t]; zc[4 * (size_t)d + 1] = (tz[t] >> qr::FOLD_B0) & 7u;
            zc[4 * (size_t)d + 2] = (uint32_t)re_bits; zc[4 * (size_t)d + 3] = (uint32_t)(re_bits >> 32);
            im[d] = tc[t].y;
        }
    }
    QR_CUDA(cudaMalloc(&pl->fold_slab, host.size()));
    QR_CUDA(cudaMemcpy(pl->fold_slab, host.data(), host.size(), cudaMemcpyHostToDevice));
    unsigned char *base = static_cast<unsigned char *>(pl->fold_slab);
    pl->fold.zc = reinterpret_cast<const uint4 *>(base);
    pl->fold.im = reinterpret_cast<const double *>(base + o_im);
    pl->fold.bend = reinterpret_cast<const uint4 *>(base + o_be);
    return QR_OK;
}
static int ensure_fold(qr_plan *pl, bool *use)
{
    *use = false;
    const char *env = getenv("QR_APPLY_FOLD");
    if (env && env[0] == '0') return QR_OK;
    const bool force = env && env[0] == '1';
    if (pl->fold_state == 0 && !force) return QR_OK;
    if (pl->fold_state == 1) { *use = true; return QR_OK; }
    uint64_t t_eval = 0, g_eval = 0;
    const int rc = ensure_fold_tables(pl, &t_eval, &g_eval);
    if (rc != QR_OK) return rc;
    if (pl->fold_blocked) return QR_OK;
    if (!force && (g_eval == 0 || t_eval < 3 * g_eval)) { pl->fold_state = 0; return QR_OK; }
    if (!force) pl->fold_state = 1;
    *use = true;
    return QR_OK;
}
// the fold kernel's rows: whole CTAs of FOLD_THREADS * FOLD_ROWS rows, aligned so that the three in-thread bits are free
static bool fold_rows_ok(uint64_t row_lo, uint64_t row_hi)
{
    const uint64_t per_cta = (uint64_t)qr::FOLD_THREADS * qr::FOLD_ROWS;
    return row_hi > row_lo && row_lo % per_cta == 0 && (row_hi - row_lo) % per_cta == 0 && (row_hi - row_lo) / per_cta <= 0x7fffffffull;
}


// ---- partner-tile H.v (apply_fold.cuh, K4c) ---------------------------------------------------------
// Chosen when the rows are whole aligned tiles, the launch fills the GPU and the masks share their high bits: S segments
// for G groups cost 16 B * S per row through L2 instead of 16 B * G.  QR_APPLY_PTILE=0 / 1 overrides the choice,
// QR_APPLY_PTILE_K the tile (10..12 -> 16 / 32 / 64 KB).
static int ptile_k()
{
    if (const char *env = getenv("QR_APPLY_PTILE_K")) { int k = atoi(env); if (k >= 10 && k <= 12) return k; }
    return 12;
}
static int ensure_ptile(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, uint32_t shard_bits, int *K_out, qr_plan::Ptile **out)
{
    *out = nullptr;
    const char *env = getenv("QR_APPLY_PTILE");
    const bool force = env && env[0] == '1';
    if (!force) return QR_OK;                  // opt-in: measured slower than the gather / fold kernels (profiles/r05_summary.md)
    const int K = ptile_k();
    *K_out = K;
    const uint64_t tile = 1ull << K, rows = row_hi - row_lo;
    if (row_hi <= row_lo || row_lo % tile || rows % tile || (rows >> K) > 0x7fffffffull || shard_bits < (uint32_t)K) return QR_OK;
    if (!force && (rows >> K) < 2ull * (uint64_t)std::max(pl->n_sm, 1)) return QR_OK;      // too few tiles to fill the GPU
    auto it = pl->ptiles.find(K);
    if (it == pl->ptiles.end()) {
        int rc = host_masks(pl);
        if (rc != QR_OK) return rc;
        rc = ensure_fold_tables(pl, nullptr, nullptr);
        if (rc != QR_OK) return rc;
        if (pl->fold_blocked) return QR_OK;
        const uint32_t G = (uint32_t)pl->n_groups;
        std::vector<uint32_t> g0, hs;
        for (uint32_t g = 0; g < G; g++)
            if (g == 0 || (pl->host_gx[g] >> K) != hs.back()) { g0.push_back(g); hs.push_back(pl->host_gx[g] >> K); }
        g0.push_back(G);
        qr_plan::Ptile pt;
        pt.n_seg = (uint32_t)hs.size();
        std::vector<uint32_t> host(g0);
        host.insert(host.end(), hs.begin(), hs.end());
        QR_CUDA(cudaMalloc(&pt.slab, host.size() * 4));
        cudaError_t e = cudaMemcpy(pt.slab, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(pt.slab); return fail(QR_ERR_CUDA, std::string("ensure_ptile: ") + cudaGetErrorString(e)); }
        pt.dev.seg_g0 = static_cast<const uint32_t *>(pt.slab);
        pt.dev.seg_h = pt.dev.seg_g0 + g0.size();
        pt.dev.n_seg = pt.n_seg;
        it = pl->ptiles.emplace(K, pt).first;
    }
    // worth it when at least a quarter of the gathers disappear (C3: 1 259 segments for 1 500 groups -> gather)
    if (!force && 4ull * it->second.n_seg > 3ull * pl->n_groups) return QR_OK;
    *out = &it->second;
    return QR_OK;
}
template <int K, int NBUF>
static int launch_ptile_k(qr_plan *pl, const qr_plan::Ptile &pt, uint64_t row_lo, uint64_t row_hi, const double2 *v, double2 *y,
                          const double2 *diag, const double *diag_re, const qr::ApplyPeerArgs &pa, cudaStream_t st)
{
    auto kern = qr::apply_ptile_kernel<K, NBUF>;
    const size_t smem = (size_t)NBUF * (16u << K);
    QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((row_hi - row_lo) >> K), 1 << (K - 3), smem, st>>>(pl->dev, pl->fold, pt.dev, row_lo, v, y, diag, diag_re, pa);
    QR_LAUNCH_CHECK("apply_ptile_kernel");
    return QR_OK;
}
static int launch_ptile(qr_plan *pl, int K, const qr_plan::Ptile &pt, uint64_t row_lo, uint64_t row_hi, const double2 *v, double2 *y,
                        const double2 *diag, const double *diag_re, const qr::ApplyPeerArgs &pa, cudaStream_t st)
{
    int nbuf = 3;
    if (const char *env = getenv("QR_APPLY_PTILE_NBUF")) { int b = atoi(env); if (b >= 2 && b <= 4) nbuf = b; }
    if (K == 12) return nbuf == 2 ? launch_ptile_k<12, 2>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st)
                                  : launch_ptile_k<12, 3>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st);
    if (K == 11) return nbuf == 2 ? launch_ptile_k<11, 2>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st)
                                  : launch_ptile_k<11, 3>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st);
    return nbuf == 2 ? launch_ptile_k<10, 2>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st)
         : nbuf == 4 ? launch_ptile_k<10, 4>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st)
                     : launch_ptile_k<10, 3>(pl, pt, row_lo, row_hi, v, y, diag, diag_re, pa, st);
}

// <v, H v> with the product: the apply kernels leave one partial per CTA in pl->dot_partials, one CTA folds them
static int dot_partials(qr_plan *pl, uint64_t n_ctas, cudaStream_t st)
{
    if (pl->dot_cap >= n_ctas) return QR_OK;
    if (pl->dot_partials) { QR_CUDA(cudaStreamSynchronize(st)); cudaFree(pl->dot_partials); pl->dot_partials = nullptr; pl->dot_cap = 0; }
    QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&pl->dot_partials), n_ctas * sizeof(double2)));
    pl->dot_cap = n_ctas;
    return QR_OK;
}
static int dot_fold(qr_plan *pl, uint64_t n_ctas, double2 *dot_out, cudaStream_t st)
{
    qr::dotc_final_kernel<<<1, qr::DOT_THREADS, 0, st>>>((uint32_t)n_ctas, pl->dot_partials, dot_out);
    QR_LAUNCH_CHECK("dotc_final_kernel");
    return QR_OK;
}

static int apply_rows(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, const double2 *v, double2 *y, cudaStream_t st,
                      double2 *dot_out = nullptr)
{
    const uint64_t rows = row_hi - row_lo;
    const double2 *diag = nullptr;
    const double *diag_re = nullptr;
    // tiled path: the rows form an aligned power-of-two block of at least one tile
    const bool pow2 = (rows & (rows - 1)) == 0 && (row_lo & (rows - 1)) == 0;
    const char *mode = getenv("QR_APPLY_V0");
    const int K = apply_tile_bits();
    const bool tiled = pow2 && rows >= (1ull << K) && mode && mode[0] == '0' && !dot_out;   // the dot epilogue lives in the default kernels
    int rc = ensure_diag_cache(pl, row_lo, row_hi, st, &diag, tiled ? nullptr : &diag_re);
    if (rc != QR_OK) return rc;
    if (tiled) {
        const int m = 63 - __builtin_clzll(rows);
        ApplyPlan *ap = nullptr;
        rc = make_apply_plan(pl, m, K, &ap);
        if (rc != QR_OK) return rc;
        if (K == 11) return launch_apply_passes<11>(pl, ap, row_lo, rows, v, y, diag, st);
        if (K == 13) return launch_apply_passes<13>(pl, ap, row_lo, rows, v, y, diag, st);
        return launch_apply_passes<12>(pl, ap, row_lo, rows, v, y, diag, st);
    }
    const uint64_t per_cta = (uint64_t)qr::APPLY_THREADS * qr::APPLY_ROWS;
    const uint64_t ctas = (rows + per_cta - 1) / per_cta;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "apply: row window too large for one launch");
    // Tiled apply (apply_tile.cuh).  The gather kernel is bound by L2 throughput (ncu, C4: 11.9 GB through L2 for 1.3 GB
    // compulsory), so from 2^20 rows up a CTA takes a contiguous run of v into shared memory with the TMA and serves every
    // group whose mask lies inside the run from there; the others are gathered by the same kernel.  When the block is larger
    // than the window L2 can hold (m > cut bit) a second pass adds the groups with a bit at or above the cut from
    // shared-memory tiles that span the top row bits (y += ...) instead of missing L2 once per group and row.
    const char *tile_env = getenv("QR_APPLY_TILE");
    const uint32_t m_blk = pow2 ? (uint32_t)(63 - __builtin_clzll(rows)) : 0u;
    uint32_t min_m = 20;
    if (const char *env = getenv("QR_APPLY_TILE_MIN")) { int v = atoi(env); if (v >= 9 && v <= 32) min_m = (uint32_t)v; }   // tests
    if (pow2 && m_blk >= min_m && !(mode && mode[0] == '1') && tile_env && tile_env[0] == '1' && !dot_out) {
        const uint32_t blk = (uint32_t)(row_lo >> m_blk), d0 = apply_cut_bit();
        const bool use_diag = diag != nullptr || diag_re != nullptr;
        TilePlan *tb = nullptr, *ta = nullptr;
        if (d0 != 0 && m_blk > d0) {
            rc = make_tile_plan(pl, m_blk, blk, std::max(d0, m_blk - 7u), TILE_FAR, false, nullptr, &tb);
            if (rc != QR_OK) return rc;
            if (!tb->ok) tb = nullptr;
        }
        rc = make_tile_plan(pl, m_blk, blk, m_blk, TILE_LOCAL, use_diag, tb, &ta);
        if (rc != QR_OK) return rc;
        if (ta->ok) {
            auto block_ptr = [&](uint32_t q) { return v + ((uint64_t)q << m_blk); };
            qr::ApplyTileArgs a;
            tile_args(pl, *ta, blk, block_ptr, a);
            a.peer[0] = v; a.n_peers = 1;
            rc = launch_run(pl, *ta, a, row_lo, y, diag, diag_re, st);
            if (rc != QR_OK || tb == nullptr) return rc;
            tile_args(pl, *tb, blk, block_ptr, a);
            a.accumulate = 1u;
            return launch_tile<false>(pl, *tb, a, row_lo, y, nullptr, nullptr, st);
        }
    }
    if (!dot_out) {
        int K = 0; qr_plan::Ptile *pt = nullptr;
        rc = ensure_ptile(pl, row_lo, row_hi, 32u, &K, &pt);
        if (rc != QR_OK) return rc;
        if (pt) return launch_ptile(pl, K, *pt, row_lo, row_hi, v, y, diag, diag_re, qr::ApplyPeerArgs{}, st);
    }
    bool fold = false;
    if (fold_rows_ok(row_lo, row_hi)) { rc = ensure_fold(pl, &fold); if (rc != QR_OK) return rc; }
    if (fold) {
        const uint64_t fctas = rows / ((uint64_t)qr::FOLD_THREADS * qr::FOLD_ROWS);
        if (dot_out) { rc = dot_partials(pl, fctas, st); if (rc != QR_OK) return rc; }
        qr::apply_fold_kernel<<<(unsigned)fctas, qr::FOLD_THREADS, 0, st>>>(pl->dev, pl->fold, (uint32_t)pl->n_groups, row_lo, row_hi, v, y, diag, diag_re,
                                                                            qr::ApplyPeerArgs{}, dot_out ? pl->dot_partials : nullptr);
        QR_LAUNCH_CHECK("apply_fold_kernel");
        return dot_out ? dot_fold(pl, fctas, dot_out, st) : QR_OK;
    }
    if (dot_out) { rc = dot_partials(pl, ctas, st); if (rc != QR_OK) return rc; }
    qr::ApplyPeerArgs pa_local{};
    if (const char *env = getenv("QR_APPLY_SWZ")) {
        const int f = atoi(env);
        const uint32_t nb = (ctas & (ctas - 1)) == 0 ? (uint32_t)(63 - __builtin_clzll(ctas)) : 0u;
        if (f > 0 && nb >= (uint32_t)f + 5u) { pa_local.swz_f = (uint32_t)f; pa_local.swz_nb = nb; }
    }
    qr::apply_direct_kernel<<<(unsigned)ctas, qr::APPLY_THREADS, 0, st>>>(pl->dev, (uint32_t)pl->n_groups, row_lo, row_hi, v, y, diag, diag_re, pa_local,
                                                                          nullptr, 0u, dot_out ? pl->dot_partials : nullptr);
    QR_LAUNCH_CHECK("apply_direct_kernel");
    return dot_out ? dot_fold(pl, ctas, dot_out, st) : QR_OK;
}

extern "C" const char *qr_plan_apply_kernel(qr_plan *pl, uint64_t row_lo, uint64_t row_hi)
{
    if (!pl || row_lo >= row_hi || row_hi > pl->dim) { fail(QR_ERR_INVALID, "qr_plan_apply_kernel: bad argument"); return ""; }
    if (cudaSetDevice(pl->device) != cudaSuccess) { fail(QR_ERR_CUDA, "qr_plan_apply_kernel: cudaSetDevice"); return ""; }
    int K = 0; qr_plan::Ptile *pt = nullptr;
    if (ensure_ptile(pl, row_lo, row_hi, 32u, &K, &pt) != QR_OK) return "";
    if (pt) return "apply_ptile_kernel";
    bool fold = false;
    if (fold_rows_ok(row_lo, row_hi) && ensure_fold(pl, &fold) != QR_OK) return "";
    return fold ? "apply_fold_kernel" : "apply_direct_kernel";
}

extern "C" int qr_apply_device(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, const double *d_v, double *d_y, void *stream)
{
    if (!pl || !d_v || !d_y) return fail(QR_ERR_INVALID, "qr_apply_device: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_apply_device: bad row range");
    if (((uintptr_t)d_v | (uintptr_t)d_y) & 15) return fail(QR_ERR_INVALID, "qr_apply_device: vectors must be 16-byte aligned");
    QR_CUDA(cudaSetDevice(pl->device));
    return apply_rows(pl, row_lo, row_hi, reinterpret_cast<const double2 *>(d_v), reinterpret_cast<double2 *>(d_y), as_stream(stream));
}

extern "C" int qr_apply_dot_device(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, const double *d_v, double *d_y, double *d_dot, void *stream)
{
    if (!pl || !d_v || !d_y || !d_dot) return fail(QR_ERR_INVALID, "qr_apply_dot_device: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_apply_dot_device: bad row range");
    if (((uintptr_t)d_v | (uintptr_t)d_y | (uintptr_t)d_dot) & 15) return fail(QR_ERR_INVALID, "qr_apply_dot_device: vectors and the result must be 16-byte aligned");
    QR_CUDA(cudaSetDevice(pl->device));
    return apply_rows(pl, row_lo, row_hi, reinterpret_cast<const double2 *>(d_v), reinterpret_cast<double2 *>(d_y), as_stream(stream),
                      reinterpret_cast<double2 *>(d_dot));
}

extern "C" int qr_apply_host(qr_plan *pl, const double *v, double *y)
{
    if (!pl || !v || !y) return fail(QR_ERR_INVALID, "qr_apply_host: NULL argument");
    QR_CUDA(cudaSetDevice(pl->device));
    void *dv = nullptr, *dy = nullptr;
    const size_t bytes = pl->dim * 16;
    QR_CUDA(cudaMalloc(&dv, bytes));
    cudaError_t e = cudaMalloc(&dy, bytes);
    if (e != cudaSuccess) { cudaFree(dv); return fail(QR_ERR_OOM, "qr_apply_host: cudaMalloc failed"); }
    int rc = QR_OK;
    e = cudaMemcpy(dv, v, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = apply_rows(pl, 0, pl->dim, static_cast<const double2 *>(dv), static_cast<double2 *>(dy), nullptr);
        if (rc == QR_OK) e = cudaMemcpy(y, dy, bytes, cudaMemcpyDeviceToHost);
    }
    cudaFree(dv); cudaFree(dy);
    if (rc != QR_OK) return rc;
    if (e != cudaSuccess) return fail(QR_ERR_CUDA, std::string("qr_apply_host: ") + cudaGetErrorString(e));
    return QR_OK;
}

extern "C" int qr_diagonal_device(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, double *d_diag, void *stream)
{
    if (!pl || !d_diag) return fail(QR_ERR_INVALID, "qr_diagonal_device: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_diagonal_device: bad row range");
    QR_CUDA(cudaSetDevice(pl->device));
    const uint64_t ctas = (row_hi - row_lo + 255) / 256;
    qr::diagonal_kernel<<<(unsigned)ctas, 256, 0, as_stream(stream)>>>(pl->dev, row_lo, row_hi, reinterpret_cast<double2 *>(d_diag));
    QR_LAUNCH_CHECK("diagonal_kernel");
    return QR_OK;
}

extern "C" int qr_csr_diagonal_device(uint64_t n_rows, uint64_t col0, const uint64_t *d_indptr, const uint64_t *d_indices,
                                      const double *d_data, double *d_diag, void *stream)
{
    if (!d_indptr || !d_indices || !d_data || !d_diag) return fail(QR_ERR_INVALID, "qr_csr_diagonal_device: NULL argument");
    if (n_rows == 0) return QR_OK;
    const uint64_t ctas = (n_rows + 255) / 256;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "qr_csr_diagonal_device: shard too large for one launch");
    qr::csr_diagonal_kernel<<<(unsigned)ctas, 256, 0, as_stream(stream)>>>(
        n_rows, col0, d_indptr, d_indices, reinterpret_cast<const double2 *>(d_data), reinterpret_cast<double2 *>(d_diag));
    QR_LAUNCH_CHECK("csr_diagonal_kernel");
    return QR_OK;
}

extern "C" int qr_spmv_device(uint64_t n_rows, const uint64_t *d_indptr, const uint64_t *d_indices,
                              const double *d_data, const double *d_v, double *d_y, void *stream)
{
    if (!d_indptr || !d_indices || !d_data || !d_v || !d_y) return fail(QR_ERR_INVALID, "qr_spmv_device: NULL argument");
    if (n_rows == 0) return QR_OK;
    const uint64_t ctas = (n_rows + 255) / 256;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "qr_spmv_device: too many rows for one launch");
    int unroll = 4;                                                 // entries of a row in flight per thread (QR_SPMV_UNROLL: sweeps, tests)
    if (const char *env = getenv("QR_SPMV_UNROLL")) { const int u = atoi(env); if (u == 1 || u == 2 || u == 4 || u == 8) unroll = u; }
    auto kern = unroll == 1 ? qr::spmv_csr_kernel : unroll == 2 ? qr::spmv_csr_unrolled_kernel<2> :
                unroll == 4 ? qr::spmv_csr_unrolled_kernel<4> : qr::spmv_csr_unrolled_kernel<8>;
    kern<<<(unsigned)ctas, 256, 0, as_stream(stream)>>>(n_rows, d_indptr, d_indices, reinterpret_cast<const double2 *>(d_data),
                                                        reinterpret_cast<const double2 *>(d_v), reinterpret_cast<double2 *>(d_y));
    QR_LAUNCH_CHECK("spmv_csr_kernel");
    return QR_OK;
}

// =====================================================================================
// K2: count + scan + compaction (eliminate_zeros)
// =====================================================================================
// counts in d_indptr[1..rows] (d_indptr[0] = 0) -> inclusive prefix sums in place; total -> *nnz_out
static int scan_counts(uint64_t n_rows, uint64_t *d_indptr, uint64_t *nnz_out, cudaStream_t st, const char *who)
{
    const uint64_t n_tiles = (n_rows + qr::SCAN_TILE - 1) / qr::SCAN_TILE;
    // tile sums: a small scratch kept per device (cudaFree would serialise the device)
    int dev = 0;
    QR_CUDA(cudaGetDevice(&dev));
    DevScratch *ds = dev_scratch(dev);
    if (!ds->tiles || ds->tiles_cap < n_tiles + 1) {
        if (ds->tiles) { QR_CUDA(cudaStreamSynchronize(st)); cudaFree(ds->tiles); }
        ds->tiles = nullptr;
        ds->tiles_cap = std::max<uint64_t>(n_tiles + 1, 4096);
        QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&ds->tiles), ds->tiles_cap * 8));
    }
    uint64_t *d_tiles = ds->tiles;
    qr::scan_tile_sums_kernel<<<(unsigned)n_tiles, qr::K2_THREADS, 0, st>>>(n_rows, d_indptr, d_tiles);
    g_launches.fetch_add(1);
    qr::scan_tile_offsets_kernel<<<1, qr::K2_THREADS, 0, st>>>(n_tiles, d_tiles, d_tiles + n_tiles);
    g_launches.fetch_add(1);
    qr::scan_apply_kernel<<<(unsigned)n_tiles, qr::K2_THREADS, 0, st>>>(n_rows, d_indptr, d_tiles);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(nnz_out, d_tiles + n_tiles, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(QR_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
    return QR_OK;
}

extern "C" int qr_count_kept_device(uint64_t n_rows, uint64_t G, const double *d_data, double tol,
                                    uint64_t *d_indptr_out, uint64_t *nnz_out, void *stream)
{
    if (!d_data || !d_indptr_out || !nnz_out) return fail(QR_ERR_INVALID, "qr_count_kept_device: NULL argument");
    if (n_rows == 0 || G == 0 || G > 0xffffffffull) return fail(QR_ERR_INVALID, "qr_count_kept_device: bad shape");
    cudaStream_t st = as_stream(stream);
    // rows per CTA: about 8192 entries, at most 2048 rows (8 KB of counters)
    uint32_t R = (uint32_t)std::min<uint64_t>(2048, std::max<uint64_t>(1, 8192 / G));
    const uint64_t ctas = (n_rows + R - 1) / R;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "qr_count_kept_device: shard too large for one launch");
    qr::count_kept_kernel<<<(unsigned)ctas, qr::K2_THREADS, R * 4, st>>>(
        n_rows, (uint32_t)G, R, reinterpret_cast<const double2 *>(d_data), tol, d_indptr_out);
    QR_LAUNCH_CHECK("count_kept_kernel");
    return scan_counts(n_rows, d_indptr_out, nnz_out, st, "qr_count_kept_device");
}

extern "C" int qr_compact_rows_device(uint64_t n_rows, uint64_t G, const uint64_t *d_indices, const double *d_data,
                                      double tol, const uint64_t *d_indptr, uint64_t *d_indices_out,
                                      double *d_data_out, void *stream)
{
    if (!d_indices || !d_data || !d_indptr || !d_indices_out || !d_data_out)
        return fail(QR_ERR_INVALID, "qr_compact_rows_device: NULL argument");
    if (n_rows == 0 || G == 0 || G > 0xffffffffull) return fail(QR_ERR_INVALID, "qr_compact_rows_device: bad shape");
    const uint64_t per_cta = qr::K2_THREADS / 32;
    const uint64_t ctas = std::min<uint64_t>((n_rows + per_cta - 1) / per_cta, (uint64_t)current_sm_count() * 64);
    qr::compact_rows_kernel<<<(unsigned)ctas, qr::K2_THREADS, 0, as_stream(stream)>>>(
        n_rows, (uint32_t)G, d_indices, reinterpret_cast<const double2 *>(d_data), tol, d_indptr, d_indices_out,
        reinterpret_cast<double2 *>(d_data_out));
    QR_LAUNCH_CHECK("compact_rows_kernel");
    return QR_OK;
}

extern "C" int qr_csr_count_kept_device(uint64_t n_rows, const uint64_t *d_indptr_in, const double *d_data, double tol,
                                        uint64_t *d_indptr_out, uint64_t *nnz_out, void *stream)
{
    if (!d_indptr_in || !d_data || !d_indptr_out || !nnz_out) return fail(QR_ERR_INVALID, "qr_csr_count_kept_device: NULL argument");
    if (n_rows == 0) return fail(QR_ERR_INVALID, "qr_csr_count_kept_device: bad shape");
    cudaStream_t st = as_stream(stream);
    const uint64_t per_cta = qr::K2_THREADS / 32;
    const uint64_t ctas = std::min<uint64_t>((n_rows + per_cta - 1) / per_cta, (uint64_t)current_sm_count() * 64);
    qr::csr_count_kept_kernel<<<(unsigned)ctas, qr::K2_THREADS, 0, st>>>(n_rows, d_indptr_in, reinterpret_cast<const double2 *>(d_data), tol, d_indptr_out);
    QR_LAUNCH_CHECK("csr_count_kept_kernel");
    return scan_counts(n_rows, d_indptr_out, nnz_out, st, "qr_csr_count_kept_device");
}

extern "C" int qr_csr_compact_device(uint64_t n_rows, const uint64_t *d_indptr_in, const uint64_t *d_indices, const double *d_data,
                                     double tol, const uint64_t *d_indptr, uint64_t *d_indices_out, double *d_data_out, void *stream)
{
    if (!d_indptr_in || !d_indices || !d_data || !d_indptr || !d_indices_out || !d_data_out)
        return fail(QR_ERR_INVALID, "qr_csr_compact_device: NULL argument");
    if (n_rows == 0) return fail(QR_ERR_INVALID, "qr_csr_compact_device: bad shape");
    const uint64_t per_cta = qr::K2_THREADS / 32;
    const uint64_t ctas = std::min<uint64_t>((n_rows + per_cta - 1) / per_cta, (uint64_t)current_sm_count() * 64);
    qr::csr_compact_rows_kernel<<<(unsigned)ctas, qr::K2_THREADS, 0, as_stream(stream)>>>(
        n_rows, d_indptr_in, d_indices, reinterpret_cast<const double2 *>(d_data), tol, d_indptr, d_indices_out,
        reinterpret_cast<double2 *>(d_data_out));
    QR_LAUNCH_CHECK("csr_compact_rows_kernel");
    return QR_OK;
}

// the per-device scratch for row windows of <= 256 MB (G * 24 bytes per row), grown on demand; caller holds ds->cwin_busy
static int compact_window(DevScratch *ds, uint64_t G, uint64_t rows, cudaStream_t st, uint64_t *win_out, void **scratch_out)
{
    uint64_t win_bytes = 256ull << 20;
    if (const char *env = getenv("QR_COMPACT_WIN_MB")) { const int v = atoi(env); if (v >= 1 && v <= 4096) win_bytes = (uint64_t)v << 20; }   // tests
    uint64_t win = std::max<uint64_t>(32, win_bytes / (G * 24) / 32 * 32);
    if (win > rows) win = rows;
    const size_t need = (size_t)win * G * 24;
    if (ds->cwin_bytes < need) {
        if (ds->cwin) { QR_CUDA(cudaStreamSynchronize(st)); cudaFree(ds->cwin); ds->cwin = nullptr; ds->cwin_bytes = 0; }
        QR_CUDA(cudaMalloc(&ds->cwin, need));
        ds->cwin_bytes = need;
    }
    *win_out = win; *scratch_out = ds->cwin;
    return QR_OK;
}

extern "C" int qr_build_compact_count(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, double tol,
                                      uint64_t *d_indptr, uint64_t *nnz_out, void *stream)
{
    if (!pl || !d_indptr || !nnz_out) return fail(QR_ERR_INVALID, "qr_build_compact_count: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_build_compact_count: bad row range");
    QR_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = as_stream(stream);
    const uint64_t rows = row_hi - row_lo, per_cta = 64ull * qr::COUNT_ROWS_WARPS;
    const uint64_t G = pl->n_groups;
    if ((size_t)32 * G * 24 > MAX_SMEM && getenv("QR_COMPACT_COUNT_WINDOWED")) {
        // Opt-in alternative for rows too long for the tiled fill: the fill kernels build row windows into scratch and
        // count_kept_kernel counts them.  Measured against count_rows_kernel (lane <-> row, values in registers): H8 0.92
        // against 1.16 ms, random G = 400 1.08 against 0.29 ms -- no clear winner, count_rows_kernel stays the default.
        int dev = 0;
        QR_CUDA(cudaGetDevice(&dev));
        DevScratch *ds = dev_scratch(dev);
        std::lock_guard<std::mutex> busy(ds->cwin_busy);
        uint64_t win = 0; void *scratch = nullptr;
        int rc = compact_window(ds, G, rows, st, &win, &scratch);
        if (rc != QR_OK) return rc;
        double2 *t_dat = static_cast<double2 *>(scratch);
        uint64_t *t_idx = reinterpret_cast<uint64_t *>(static_cast<char *>(scratch) + win * G * 16);
        const uint32_t R = (uint32_t)std::min<uint64_t>(2048, std::max<uint64_t>(1, 8192 / G));
        for (uint64_t w0 = row_lo; w0 < row_hi; w0 += win) {
            const uint64_t w1 = std::min(row_hi, w0 + win), n = w1 - w0;
            rc = build_rows(pl, w0, w1, nullptr, t_idx, t_dat, 0, st);
            if (rc != QR_OK) return rc;
            qr::count_kept_kernel<<<(unsigned)((n + R - 1) / R), qr::K2_THREADS, R * 4, st>>>(n, (uint32_t)G, R, t_dat, tol, d_indptr + (w0 - row_lo),
                                                                                             w0 == row_lo ? 1u : 0u);
            QR_LAUNCH_CHECK("count_kept_kernel");
        }
        return scan_counts(rows, d_indptr, nnz_out, st, "qr_build_compact_count");   // synchronises: the scratch is free again
    }
    const uint64_t ctas = (rows + per_cta - 1) / per_cta;
    if (ctas > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "qr_build_compact_count: row window too large for one launch");
    // few rows and many groups (molecular Hamiltonians: 2^16 rows x 981 groups = 128 CTAs of 8 warps): the groups are cut
    // into slices along gridDim.y and the slices add their counts atomically (H8: 1.16 -> 0.3 ms)
    uint32_t slices = 1;
    {
        const uint64_t want_ctas = (uint64_t)pl->n_sm * 8;
        if (ctas < want_ctas && G >= 64) slices = (uint32_t)std::min<uint64_t>({(want_ctas + ctas - 1) / ctas, G / 32, (uint64_t)64});
        if (const char *env = getenv("QR_COUNT_ROWS_SLICES")) { const int v = atoi(env); if (v >= 1 && (uint64_t)v <= G && v <= 65535) slices = (uint32_t)v; }
        if (slices > 1) QR_CUDA(cudaMemsetAsync(d_indptr, 0, (rows + 1) * 8, st));
    }
    qr::count_rows_kernel<<<dim3((unsigned)ctas, slices, 1), 32 * qr::COUNT_ROWS_WARPS, 0, st>>>(pl->dev, (uint32_t)pl->n_groups, row_lo, row_hi, tol, d_indptr);
    QR_LAUNCH_CHECK("count_rows_kernel");
    return scan_counts(rows, d_indptr, nnz_out, st, "qr_build_compact_count");
}

extern "C" int qr_build_compact_fill(qr_plan *pl, uint64_t row_lo, uint64_t row_hi, double tol,
                                     const uint64_t *d_indptr, uint64_t *d_indices, double *d_data, void *stream)
{
    if (!pl || !d_indptr || !d_indices || !d_data) return fail(QR_ERR_INVALID, "qr_build_compact_fill: NULL argument");
    if (row_lo >= row_hi || row_hi > pl->dim) return fail(QR_ERR_INVALID, "qr_build_compact_fill: bad row range");
    QR_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = as_stream(stream);
    const uint64_t G = pl->n_groups, rows = row_hi - row_lo;
    const size_t smem32 = (size_t)32 * G * 24;
    if (smem32 <= MAX_SMEM) {
        // 64-row tiles (two strips per warp visit, as the staged fill) while two CTAs still fit per SM
        const bool two = 2 * smem32 <= 113 * 1024;
        const uint64_t R = two ? 64 : 32;
        const size_t smem = two ? 2 * smem32 : smem32;
        const uint64_t t0 = row_lo / R * R, tiles = (row_hi - t0 + R - 1) / R;
        if (tiles > 0x7fffffffull) return fail(QR_ERR_UNSUPPORTED, "qr_build_compact_fill: row window too large for one launch");
        auto kern = two ? qr::fill_compact_kernel<2, 8> : qr::fill_compact_kernel<1, 8>;
        QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)tiles, 256, smem, st>>>(pl->dev, (uint32_t)G, t0, row_lo, row_hi, tol, d_indptr, d_indices,
                                                 reinterpret_cast<double2 *>(d_data));
        QR_LAUNCH_CHECK("fill_compact_kernel");
        return QR_OK;
    }
    // rows too long for a shared-memory tile: build row windows into scratch, compact each window
    int dev = 0;
    QR_CUDA(cudaGetDevice(&dev));
    DevScratch *ds = dev_scratch(dev);
    std::lock_guard<std::mutex> busy(ds->cwin_busy);
    uint64_t win = 0; void *scratch = nullptr;
    { const int rc0 = compact_window(ds, G, rows, st, &win, &scratch); if (rc0 != QR_OK) return rc0; }
    double2 *t_dat = static_cast<double2 *>(scratch);
    uint64_t *t_idx = reinterpret_cast<uint64_t *>(static_cast<char *>(scratch) + win * G * 16);
    int rc = QR_OK;
    for (uint64_t w0 = row_lo; w0 < row_hi && rc == QR_OK; w0 += win) {
        const uint64_t w1 = std::min(row_hi, w0 + win), n = w1 - w0;
        rc = build_rows(pl, w0, w1, nullptr, t_idx, t_dat, 0, st);
        if (rc != QR_OK) break;
        const uint64_t per_cta = qr::K2_THREADS / 32;
        const uint64_t ctas = std::min<uint64_t>((n + per_cta - 1) / per_cta, (uint64_t)pl->n_sm * 64);
        qr::compact_rows_kernel<<<(unsigned)ctas, qr::K2_THREADS, 0, st>>>(
            n, (uint32_t)G, t_idx, t_dat, tol, d_indptr + (w0 - row_lo), d_indices, reinterpret_cast<double2 *>(d_data));
        g_launches.fetch_add(1);
        if (cudaGetLastError() != cudaSuccess) rc = fail(QR_ERR_CUDA, "qr_build_compact_fill: compact_rows_kernel launch failed");
    }
    cudaStreamSynchronize(st);                                      // the scratch belongs to the device: free for the next caller
    return rc;
}

static unsigned vec_grid(uint64_t n)
{
    uint64_t ctas = (n + 255) / 256;
    const uint64_t cap = (uint64_t)current_sm_count() * 16;
    return (unsigned)(ctas < cap ? (ctas ? ctas : 1) : cap);
}

extern "C" int qr_axpby_device(uint64_t n, const double a[2], const double *x, const double b[2],
                               const double *y, double *z, void *stream)
{
    if (!a || !b || !x || !y || !z) return fail(QR_ERR_INVALID, "qr_axpby_device: NULL argument");
    if (n == 0) return QR_OK;
    qr::vec_axpby_kernel<0><<<vec_grid(n), 256, 0, as_stream(stream)>>>(
        n, make_double2(a[0], a[1]), reinterpret_cast<const double2 *>(x), make_double2(b[0], b[1]),
        reinterpret_cast<const double2 *>(y), reinterpret_cast<double2 *>(z));
    QR_LAUNCH_CHECK("vec_axpby_kernel");
    return QR_OK;
}
extern "C" int qr_axpy_device(uint64_t n, const double a[2], const double *x, const double *y, double *z, void *stream)
{
    if (!a || !x || !y || !z) return fail(QR_ERR_INVALID, "qr_axpy_device: NULL argument");
    if (n == 0) return QR_OK;
    qr::vec_axpby_kernel<1><<<vec_grid(n), 256, 0, as_stream(stream)>>>(
        n, make_double2(a[0], a[1]), reinterpret_cast<const double2 *>(x), make_double2(0, 0),
        reinterpret_cast<const double2 *>(y), reinterpret_cast<double2 *>(z));
    QR_LAUNCH_CHECK("vec_axpy_kernel");
    return QR_OK;
}
extern "C" int qr_ax_device(uint64_t n, const double a[2], const double *x, double *z, void *stream)
{
    if (!a || !x || !z) return fail(QR_ERR_INVALID, "qr_ax_device: NULL argument");
    if (n == 0) return QR_OK;
    qr::vec_axpby_kernel<2><<<vec_grid(n), 256, 0, as_stream(stream)>>>(
        n, make_double2(a[0], a[1]), reinterpret_cast<const double2 *>(x), make_double2(0, 0),
        nullptr, reinterpret_cast<double2 *>(z));
    QR_LAUNCH_CHECK("vec_ax_kernel");
    return QR_OK;
}

extern "C" int qr_precond2_device(uint64_t n, const double *d_diag, const double *d_dx, const double e[2], double tol,
                                  double *d_out, void *stream)
{
    if (!d_diag || !d_dx || !e || !d_out) return fail(QR_ERR_INVALID, "qr_precond2_device: NULL argument");
    if (n == 0) return QR_OK;
    qr::precond2_kernel<<<vec_grid(n), 256, 0, as_stream(stream)>>>(
        n, reinterpret_cast<const double2 *>(d_diag), reinterpret_cast<const double2 *>(d_dx), make_double2(e[0], e[1]),
        tol, reinterpret_cast<double2 *>(d_out));
    QR_LAUNCH_CHECK("precond2_kernel");
    return QR_OK;
}

constexpr unsigned kReduceGrid = 148 * 8;             // partial sums per reduction (a fixed count keeps the fold order, hence the result, device-independent)
static int reduce_scratch(double2 **partials, double2 **tmp)  // per device: kReduceGrid partial sums + one folded value
{
    int dev = 0;
    QR_CUDA(cudaGetDevice(&dev));
    DevScratch *ds = dev_scratch(dev);
    if (!ds->partials) {
        QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&ds->partials), (kReduceGrid + 1) * sizeof(double2)));
        ds->tmp = ds->partials + kReduceGrid;
    }
    *partials = ds->partials;
    if (tmp) *tmp = ds->tmp;
    return QR_OK;
}

extern "C" int qr_lanczos_update_device(uint64_t n, const double alpha[2], const double beta[2], const double *w,
                                        const double *v, const double *v_prev, double *w_out, double *d_norm2_out,
                                        void *stream)
{
    if (!alpha || !beta || !w || !v || !w_out || !d_norm2_out) return fail(QR_ERR_INVALID, "qr_lanczos_update_device: NULL argument");
    double2 *partials = nullptr, *tmp = nullptr;
    int rc = reduce_scratch(&partials, &tmp);
    if (rc != QR_OK) return rc;
    qr::lanczos_update_kernel<<<kReduceGrid, 256, 0, as_stream(stream)>>>(
        n, make_double2(alpha[0], alpha[1]), make_double2(beta[0], beta[1]), reinterpret_cast<const double2 *>(w),
        reinterpret_cast<const double2 *>(v), reinterpret_cast<const double2 *>(v_prev),
        reinterpret_cast<double2 *>(w_out), partials);
    QR_LAUNCH_CHECK("lanczos_update_kernel");
    // fold the partials; the result is (norm2, 0): only the first double is the caller's
    qr::dotc_final_kernel<<<1, qr::DOT_THREADS, 0, as_stream(stream)>>>(kReduceGrid, partials, tmp);
    QR_LAUNCH_CHECK("dotc_final_kernel");
    QR_CUDA(cudaMemcpyAsync(d_norm2_out, tmp, sizeof(double), cudaMemcpyDeviceToDevice, as_stream(stream)));
    return QR_OK;
}

extern "C" int qr_lanczos_coef_device(double *d_state, uint32_t k, uint32_t K, uint32_t phase, void *stream)
{
    if (!d_state || k >= K || phase > 1u) return fail(QR_ERR_INVALID, "qr_lanczos_coef_device: bad argument");
    qr::lanczos_coef_kernel<<<1, 32, 0, as_stream(stream)>>>(d_state, k, K, phase);
    QR_LAUNCH_CHECK("lanczos_coef_kernel");
    return QR_OK;
}

extern "C" int qr_lanczos_update_dev(uint64_t n, double *d_state, const double *d_y, const double *d_u, const double *d_u_prev,
                                     double *d_w_out, void *stream)
{
    if (!d_state || !d_y || !d_u || !d_w_out) return fail(QR_ERR_INVALID, "qr_lanczos_update_dev: NULL argument");
    if ((uintptr_t)d_state & 15) return fail(QR_ERR_INVALID, "qr_lanczos_update_dev: the state must be 16-byte aligned");
    double2 *partials = nullptr;
    int rc = reduce_scratch(&partials, nullptr);
    if (rc != QR_OK) return rc;
    qr::lanczos_update_dev_kernel<<<kReduceGrid, 256, 0, as_stream(stream)>>>(
        n, d_state, reinterpret_cast<const double2 *>(d_y), reinterpret_cast<const double2 *>(d_u),
        reinterpret_cast<const double2 *>(d_u_prev), reinterpret_cast<double2 *>(d_w_out), partials);
    QR_LAUNCH_CHECK("lanczos_update_dev_kernel");
    qr::dotc_final_kernel<<<1, qr::DOT_THREADS, 0, as_stream(stream)>>>(kReduceGrid, partials, reinterpret_cast<double2 *>(d_state + 2));
    QR_LAUNCH_CHECK("dotc_final_kernel");
    return QR_OK;
}

extern "C" int qr_dotc_device(uint64_t n, const double *x, const double *y, double *d_out, void *stream)
{
    if (!x || !y || !d_out) return fail(QR_ERR_INVALID, "qr_dotc_device: NULL argument");
    const unsigned grid = kReduceGrid;
    double2 *partials = nullptr;
    { int rc = reduce_scratch(&partials, nullptr); if (rc != QR_OK) return rc; }
    qr::dotc_partial_kernel<<<grid, qr::DOT_THREADS, 0, as_stream(stream)>>>(
        n, reinterpret_cast<const double2 *>(x), reinterpret_cast<const double2 *>(y), partials);
    QR_LAUNCH_CHECK("dotc_partial_kernel");
    qr::dotc_final_kernel<<<1, qr::DOT_THREADS, 0, as_stream(stream)>>>(grid, partials, reinterpret_cast<double2 *>(d_out));
    QR_LAUNCH_CHECK("dotc_final_kernel");
    return QR_OK;
}

// =====================================================================================
// NCCL (dlopen'ed: the library loads and the 1-GPU path runs without it)
// =====================================================================================
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        const char *names[] = {"libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
        if (!a.handle) { a.why = "cannot dlopen libnccl.so.2"; return a; }
#define QR_SYM(field, name)                                                       \
        a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name));     \
        if (!a.field) { a.why = std::string("missing NCCL symbol ") + name; return a; }
        QR_SYM(GetUniqueId, "ncclGetUniqueId")
        QR_SYM(CommInitRank, "ncclCommInitRank")
        QR_SYM(CommDestroy, "ncclCommDestroy")
        QR_SYM(AllGather, "ncclAllGather")
        QR_SYM(AllReduce, "ncclAllReduce")
        QR_SYM(GetErrorString, "ncclGetErrorString")
#undef QR_SYM
        a.ok = true;
        return a;
    }();
    return api;
}

#define QR_NCCL(expr)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (expr);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return fail(QR_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(r__));  \
    } while (0)
}  // namespace

extern "C" int qr_comm_unique_id(void *id_out)
{
    static_assert(sizeof(ncclUniqueId) == QR_UNIQUE_ID_BYTES, "ncclUniqueId size");
    if (!id_out) return fail(QR_ERR_INVALID, "qr_comm_unique_id: NULL argument");
    if (!nccl().ok) return fail(QR_ERR_NCCL, nccl().why);
    ncclUniqueId id;
    QR_NCCL(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return QR_OK;
}

extern "C" int qr_comm_create(const void *id, int n_ranks, int rank, int device, qr_comm **out)
{
    if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(QR_ERR_INVALID, "qr_comm_create: bad argument");
    *out = nullptr;
    if (!nccl().ok) return fail(QR_ERR_NCCL, nccl().why);
    QR_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    QR_NCCL(nccl().CommInitRank(&c, n_ranks, uid, rank));
    qr_comm *cm = new (std::nothrow) qr_comm();
    if (!cm) return fail(QR_ERR_OOM, "qr_comm_create: host allocation failed");
    cm->comm = c; cm->n_ranks = n_ranks; cm->rank = rank; cm->device = device;
    *out = cm;
    // Epoch flags for the fused apply, exchanged over the communicator itself: every rank allocates
    // {ready[P], done[P], counter}, all-gathers the CUDA IPC handles and maps its peers' arrays.  If any step
    // fails (IPC unavailable, peers in one process) the fused apply keeps its NCCL-barrier form.
    if (n_ranks > 1 && n_ranks <= qr::TILE_MAX_PEERS) {
        const size_t flag_bytes = 4096;
        cudaIpcMemHandle_t *d_handles = nullptr;
        std::vector<cudaIpcMemHandle_t> handles((size_t)n_ranks);
        bool ok = cudaMalloc(reinterpret_cast<void **>(&cm->d_flags), flag_bytes) == cudaSuccess &&
                  cudaMemset(cm->d_flags, 0, flag_bytes) == cudaSuccess &&
                  cudaIpcGetMemHandle(&handles[rank], cm->d_flags) == cudaSuccess &&
                  cudaMalloc(reinterpret_cast<void **>(&d_handles), sizeof(cudaIpcMemHandle_t) * n_ranks) == cudaSuccess &&
                  cudaMemcpy(d_handles + rank, &handles[rank], sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice) == cudaSuccess;
        // every rank takes part in the collective whatever its local outcome (a rank that failed sends zeros)
        if (!ok && d_handles == nullptr) cudaMalloc(reinterpret_cast<void **>(&d_handles), sizeof(cudaIpcMemHandle_t) * n_ranks);
        if (d_handles != nullptr) {
            if (!ok) cudaMemset(d_handles + rank, 0, sizeof(cudaIpcMemHandle_t));
            const bool sent = nccl().AllGather(d_handles + rank, d_handles, sizeof(cudaIpcMemHandle_t), ncclChar, c, nullptr) == ncclSuccess &&
                              cudaStreamSynchronize(nullptr) == cudaSuccess &&
                              cudaMemcpy(handles.data(), d_handles, sizeof(cudaIpcMemHandle_t) * n_ranks, cudaMemcpyDeviceToHost) == cudaSuccess;
            ok = ok && sent;
            cudaFree(d_handles);
        }
        const cudaIpcMemHandle_t zero{};
        for (int q = 0; ok && q < n_ranks; q++) {
            if (q == rank) { cm->peer_flags[q] = cm->d_flags; continue; }
            void *ptr = nullptr;
            ok = memcmp(&handles[q], &zero, sizeof(zero)) != 0 &&
                 cudaIpcOpenMemHandle(&ptr, handles[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            cm->peer_flags[q] = static_cast<uint64_t *>(ptr);
        }
        if (ok) ok = cudaMalloc(reinterpret_cast<void **>(&cm->d_peer_flags), sizeof(uint64_t *) * n_ranks) == cudaSuccess &&
                     cudaMemcpy(cm->d_peer_flags, cm->peer_flags, sizeof(uint64_t *) * n_ranks, cudaMemcpyHostToDevice) == cudaSuccess;
        if (const char *env = getenv("QR_P2P_FLAGS")) if (env[0] == '0') ok = false;
        cm->flags_ok = ok;
        cudaGetLastError();
    }
    return QR_OK;
}

extern "C" int qr_comm_destroy(qr_comm *cm)
{
    if (!cm) return QR_OK;
    cudaSetDevice(cm->device);
    if (cm->d_scratch) cudaFree(cm->d_scratch);
    for (int q = 0; q < cm->n_ranks && q < qr::TILE_MAX_PEERS; q++)
        if (q != cm->rank && cm->peer_flags[q]) cudaIpcCloseMemHandle(cm->peer_flags[q]);
    if (cm->d_flags) cudaFree(cm->d_flags);
    if (cm->d_peer_flags) cudaFree(cm->d_peer_flags);
    if (cm->comm && nccl().ok) nccl().CommDestroy(cm->comm);
    delete cm;
    return QR_OK;
}

extern "C" int qr_apply_distributed(qr_plan *pl, qr_comm *cm, const double *d_v_shard, double *d_v_full,
                                    double *d_y_shard, void *stream)
{
    if (!pl || !cm || !d_v_shard || !d_v_full || !d_y_shard) return fail(QR_ERR_INVALID, "qr_apply_distributed: NULL argument");
    const uint64_t P = (uint64_t)cm->n_ranks;
    if (pl->dim % P) return fail(QR_ERR_INVALID, "qr_apply_distributed: dim is not divisible by n_ranks");
    const uint64_t shard = pl->dim / P;
    QR_CUDA(cudaSetDevice(pl->device));
    // accel.rs has no counterpart: the reference is single-process.  Rows are
    // block-sharded; the only exchange the path needs is this all-gather of v.
    QR_NCCL(nccl().AllGather(d_v_shard, d_v_full, 2 * shard, ncclDouble, cm->comm, as_stream(stream)));
    return apply_rows(pl, shard * cm->rank, shard * (cm->rank + 1), reinterpret_cast<const double2 *>(d_v_full),
                      reinterpret_cast<double2 *>(d_y_shard), as_stream(stream));
}

static int apply_p2p(qr_plan *pl, qr_comm *cm, const double *const *v_shards, double *d_y_shard, double2 *dot_out, void *stream);
extern "C" int qr_apply_p2p(qr_plan *pl, qr_comm *cm, const double *const *v_shards, double *d_y_shard, void *stream)
{
    return apply_p2p(pl, cm, v_shards, d_y_shard, nullptr, stream);
}
extern "C" int qr_apply_p2p_dot(qr_plan *pl, qr_comm *cm, const double *const *v_shards, double *d_y_shard, double *d_dot, void *stream)
{
    if (!d_dot || ((uintptr_t)d_dot & 15)) return fail(QR_ERR_INVALID, "qr_apply_p2p_dot: d_dot must be a 16-byte aligned device pointer");
    return apply_p2p(pl, cm, v_shards, d_y_shard, reinterpret_cast<double2 *>(d_dot), stream);
}
static int apply_p2p(qr_plan *pl, qr_comm *cm, const double *const *v_shards, double *d_y_shard, double2 *dot_out, void *stream)
{
    if (!pl || !cm || !v_shards || !d_y_shard) return fail(QR_ERR_INVALID, "qr_apply_p2p: NULL argument");
    const uint64_t P = (uint64_t)cm->n_ranks;
    if ((P & (P - 1)) || pl->dim % P) return fail(QR_ERR_INVALID, "qr_apply_p2p: n_ranks must be a power of two dividing dim");
    const uint64_t shard = pl->dim / P;
    const uint32_t m = (uint32_t)(63 - __builtin_clzll(shard));
    for (uint64_t o = 0; o < P; o++)
        if (!v_shards[o] || ((uintptr_t)v_shards[o] & 15)) return fail(QR_ERR_INVALID, "qr_apply_p2p: bad shard pointer");
    QR_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = as_stream(stream);
    const uint64_t row_lo = shard * cm->rank, row_hi = row_lo + shard;
    const double2 *diag = nullptr;
    const double *diag_re = nullptr;
    int rc = ensure_diag_cache(pl, row_lo, row_hi, st, &diag, &diag_re);
    if (rc != QR_OK) return rc;

    // Tiled path: remote runs pulled by TMA into shared memory while the local groups are gathered; ranks
    // synchronise through epoch flags in IPC-mapped memory, no NCCL call (apply_tile.cuh).
    const char *tile_env = getenv("QR_P2P_TILE");
    if (cm->flags_ok && P <= (uint64_t)qr::TILE_MAX_PEERS && tile_env && tile_env[0] == '1' && !dot_out) {
        TilePlan *tp = nullptr;
        rc = make_tile_plan(pl, m, (uint32_t)cm->rank, m, TILE_P2P, diag != nullptr || diag_re != nullptr, nullptr, &tp);
        if (rc != QR_OK) return rc;
        if (tp->ok) {
            qr::ApplyTileArgs a;
            tile_args(pl, *tp, (uint32_t)cm->rank, [&](uint32_t q) { return reinterpret_cast<const double2 *>(v_shards[q]); }, a);
            for (uint64_t o = 0; o < P; o++) {
                a.peer[o] = reinterpret_cast<const double2 *>(v_shards[o]) - (o << m);      // indexable with the GLOBAL row id
                a.peer_flags[o] = cm->peer_flags[o];
            }
            a.flags_local = cm->d_flags;
            a.cta_counter = reinterpret_cast<uint32_t *>(cm->d_flags + 2 * P);
            a.epoch = ++cm->epoch;
            a.n_peers = (uint32_t)P; a.need_mask = tp->need_mask;
            rc = launch_run(pl, *tp, a, row_lo, reinterpret_cast<double2 *>(d_y_shard), diag, diag_re, st);
            if (rc != QR_OK) return rc;
            // nobody overwrites a shard while a peer may still be reading it: the ranks that read mine are the ones I read
            qr::p2p_wait_done_kernel<<<1, 32, 0, st>>>(cm->d_flags, (uint32_t)P, (uint32_t)cm->rank, tp->need_mask, a.epoch);
            QR_LAUNCH_CHECK("p2p_wait_done_kernel");
            return QR_OK;
        }
    }

    // Gather path (default): the gather kernel with the peers' shards read in place over NVLink.  Ranks synchronise through
    // epoch flags in IPC-mapped memory, set and awaited by two one-CTA kernels around the apply; without flags (IPC
    // unavailable) two NCCL all-reduces of one double do.
    rc = host_masks(pl);
    if (rc != QR_OK) return rc;
    uint32_t need_mask = 0;
    for (uint32_t x : pl->host_gx) if (((uint64_t)x >> m) != 0) need_mask |= 1u << (((uint32_t)cm->rank ^ (uint32_t)((uint64_t)x >> m)) & 31u);
    qr::ApplyPeerArgs pa{};
    for (uint64_t o = 0; o < P; o++) pa.peer[o] = reinterpret_cast<const double2 *>(v_shards[o]) - (o << m);   // indexable with the GLOBAL row id
    pa.n_peers = (uint32_t)P; pa.shard_bits = m;
    const bool flags = cm->flags_ok && P <= (uint64_t)qr::APPLY_MAX_PEERS;
    uint64_t epoch = 0;
    if (flags) {
        epoch = ++cm->epoch;
        qr::p2p_ready_kernel<<<1, 32, 0, st>>>(cm->d_flags, cm->d_peer_flags, (uint32_t)P, (uint32_t)cm->rank, need_mask, epoch);
        QR_LAUNCH_CHECK("p2p_ready_kernel");
    } else {
        if (!cm->d_scratch) {
            QR_CUDA(cudaMalloc(reinterpret_cast<void **>(&cm->d_scratch), 16));
            QR_CUDA(cudaMemset(cm->d_scratch, 0, 16));
        }
        // every rank's shard is complete before anyone reads it ...
        QR_NCCL(nccl().AllReduce(cm->d_scratch, cm->d_scratch, 1, ncclDouble, ncclSum, cm->comm, st));
    }
    bool fold = false;
    int ptK = 0; qr_plan::Ptile *ptp = nullptr;
    if (!dot_out) rc = ensure_ptile(pl, row_lo, row_hi, m, &ptK, &ptp);
    if (rc != QR_OK) return rc;
    if (!ptp && fold_rows_ok(row_lo, row_hi)) { rc = ensure_fold(pl, &fold); if (rc != QR_OK) return rc; }
    if (ptp) {                                                     // same choice, same kernel as apply_rows: the two forms agree bit for bit
        rc = launch_ptile(pl, ptK, *ptp, row_lo, row_hi, nullptr, reinterpret_cast<double2 *>(d_y_shard), diag, diag_re, pa, st);
        if (rc != QR_OK) return rc;
    } else if (fold) {
        const uint64_t fctas = shard / ((uint64_t)qr::FOLD_THREADS * qr::FOLD_ROWS);
        if (dot_out) { rc = dot_partials(pl, fctas, st); if (rc != QR_OK) return rc; }
        qr::apply_fold_kernel<<<(unsigned)fctas, qr::FOLD_THREADS, 0, st>>>(
            pl->dev, pl->fold, (uint32_t)pl->n_groups, row_lo, row_hi, nullptr, reinterpret_cast<double2 *>(d_y_shard), diag, diag_re, pa,
            dot_out ? pl->dot_partials : nullptr);
        QR_LAUNCH_CHECK("apply_fold_kernel(p2p)");
        if (dot_out) { rc = dot_fold(pl, fctas, dot_out, st); if (rc != QR_OK) return rc; }
    } else {
        const uint64_t per_cta = (uint64_t)qr::APPLY_THREADS * qr::APPLY_ROWS;
        const uint64_t ctas = (shard + per_cta - 1) / per_cta;
        if (dot_out) { rc = dot_partials(pl, ctas, st); if (rc != QR_OK) return rc; }
        qr::apply_direct_kernel<<<(unsigned)ctas, qr::APPLY_THREADS, 0, st>>>(
            pl->dev, (uint32_t)pl->n_groups, row_lo, row_hi, nullptr, reinterpret_cast<double2 *>(d_y_shard), diag, diag_re, pa,
            nullptr, 0u, dot_out ? pl->dot_partials : nullptr);
        QR_LAUNCH_CHECK("apply_direct_kernel(p2p)");
        if (dot_out) { rc = dot_fold(pl, ctas, dot_out, st); if (rc != QR_OK) return rc; }
    }
    // ... and nobody overwrites a shard while a peer may still be reading it
    if (flags) {
        qr::p2p_done_kernel<<<1, 32, 0, st>>>(cm->d_flags, cm->d_peer_flags, (uint32_t)P, (uint32_t)cm->rank, need_mask, epoch);
        QR_LAUNCH_CHECK("p2p_done_kernel");
    } else {
        QR_NCCL(nccl().AllReduce(cm->d_scratch, cm->d_scratch, 1, ncclDouble, ncclSum, cm->comm, st));
    }
    return QR_OK;
}

extern "C" int qr_ipc_get_handle(void *d_ptr, void *handle_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == QR_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    if (!d_ptr || !handle_out) return fail(QR_ERR_INVALID, "qr_ipc_get_handle: NULL argument");
    cudaIpcMemHandle_t h;
    QR_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle_out, &h, sizeof(h));
    return QR_OK;
}
extern "C" int qr_ipc_open_handle(const void *handle, void **d_ptr_out)
{
    if (!handle || !d_ptr_out) return fail(QR_ERR_INVALID, "qr_ipc_open_handle: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *d_ptr_out = nullptr;
    QR_CUDA(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return QR_OK;
}
extern "C" int qr_ipc_close_handle(void *d_ptr)
{
    if (d_ptr) QR_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return QR_OK;
}

extern "C" int qr_allreduce_sum_f64(qr_comm *cm, double *d_buf, size_t count, void *stream)
{
    if (!cm || !d_buf) return fail(QR_ERR_INVALID, "qr_allreduce_sum_f64: NULL argument");
    QR_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, cm->comm, as_stream(stream)));
    return QR_OK;
}

// =====================================================================================
// runtime helpers
// =====================================================================================
extern "C" int qr_device_count(int *count)
{
    if (!count) return fail(QR_ERR_INVALID, "qr_device_count: NULL argument");
    *count = 0;
    QR_CUDA(cudaGetDeviceCount(count));
    return QR_OK;
}
extern "C" int qr_device_name(int device, char *buf, size_t buf_len)
{
    if (!buf || buf_len == 0) return fail(QR_ERR_INVALID, "qr_device_name: NULL argument");
    cudaDeviceProp prop;
    QR_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, buf_len, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    return QR_OK;
}
extern "C" int qr_set_device(int device) { QR_CUDA(cudaSetDevice(device)); return QR_OK; }
extern "C" int qr_malloc_device(void **ptr, size_t bytes)
{
    if (!ptr) return fail(QR_ERR_INVALID, "qr_malloc_device: NULL argument");
    *ptr = nullptr;
    QR_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
    return QR_OK;
}
extern "C" int qr_free_device(void *ptr) { if (ptr) QR_CUDA(cudaFree(ptr)); return QR_OK; }
extern "C" int qr_malloc_host(void **ptr, size_t bytes)
{
    if (!ptr) return fail(QR_ERR_INVALID, "qr_malloc_host: NULL argument");
    *ptr = nullptr;
    QR_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return QR_OK;
}
extern "C" int qr_free_host(void *ptr) { if (ptr) QR_CUDA(cudaFreeHost(ptr)); return QR_OK; }
extern "C" int qr_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    if (stream) QR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    else QR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return QR_OK;
}
extern "C" int qr_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    if (stream) QR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    else QR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return QR_OK;
}
extern "C" int qr_memset_device(void *dst, int value, size_t bytes, void *stream)
{
    QR_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(stream)));
    return QR_OK;
}
extern "C" int qr_stream_create(void **stream)
{
    if (!stream) return fail(QR_ERR_INVALID, "qr_stream_create: NULL argument");
    cudaStream_t s;
    QR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return QR_OK;
}
extern "C" int qr_stream_destroy(void *stream) { if (stream) QR_CUDA(cudaStreamDestroy(as_stream(stream))); return QR_OK; }
extern "C" int qr_stream_synchronize(void *stream)
{
    if (stream) QR_CUDA(cudaStreamSynchronize(as_stream(stream)));
    else QR_CUDA(cudaDeviceSynchronize());
    return QR_OK;
}
// ---- CUDA graphs: a caller that repeats the same device sequence (canonicalise -> fill, an H.v
// iteration) records it once and replays it with one launch, so the host's launch rate stops mattering.
extern "C" int qr_graph_begin_capture(void *stream)
{
    if (!stream) return fail(QR_ERR_INVALID, "qr_graph_begin_capture: capture needs an explicit stream");
    QR_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
    return QR_OK;
}
extern "C" int qr_graph_end_capture(void *stream, void **graph_exec)
{
    if (!stream || !graph_exec) return fail(QR_ERR_INVALID, "qr_graph_end_capture: NULL argument");
    *graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    QR_CUDA(cudaStreamEndCapture(as_stream(stream), &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(QR_ERR_CUDA, std::string("qr_graph_end_capture: ") + cudaGetErrorString(e));
    *graph_exec = exec;
    return QR_OK;
}
extern "C" int qr_graph_launch(void *graph_exec, void *stream)
{
    if (!graph_exec) return fail(QR_ERR_INVALID, "qr_graph_launch: NULL graph");
    QR_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
    return QR_OK;
}
extern "C" int qr_graph_destroy(void *graph_exec)
{
    if (graph_exec) QR_CUDA(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)));
    return QR_OK;
}

extern "C" int qr_event_create(void **event)
{
    if (!event) return fail(QR_ERR_INVALID, "qr_event_create: NULL argument");
    cudaEvent_t e;
    QR_CUDA(cudaEventCreate(&e));
    *event = e;
    return QR_OK;
}
extern "C" int qr_event_destroy(void *event) { if (event) QR_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event))); return QR_OK; }
extern "C" int qr_event_record(void *event, void *stream)
{
    // inside a stream capture the record becomes a graph node of its own (timestamps usable after replay)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (stream && cudaStreamIsCapturing(as_stream(stream), &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) {
        QR_CUDA(cudaEventRecordWithFlags(reinterpret_cast<cudaEvent_t>(event), as_stream(stream), cudaEventRecordExternal));
        return QR_OK;
    }
    QR_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
    return QR_OK;
}
extern "C" int qr_event_elapsed_ms(void *start, void *stop, float *ms)
{
    if (!ms) return fail(QR_ERR_INVALID, "qr_event_elapsed_ms: NULL argument");
    QR_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(stop)));
    QR_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(stop)));
    return QR_OK;
}
