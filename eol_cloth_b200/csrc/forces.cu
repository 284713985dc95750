// forces.cu — Forces::fill on the GPU (include/eolc.h, "Forces::fill" section).
//
// Replaces /root/reference/src/Forces.cpp:912-930 (fill), :331-520 (faceBasedF), :685-910 (edgeBasedF) and Eigen's
// setFromTriplets.  The Lagrangian 3N x 3N part of M / MDK and the face forces come from the pipelines below; meshes with EoL
// nodes add the two small kernels at the end of this file (eol_elements_kernel, eol_gather_kernel — design in forces_eol.h).
//
// Default pipeline "tiles" (one persistent kernel, no HBM scratch) — assemble_tiles_kernel:
//   the nodes are partitioned into spatially compact tiles (forces_plan.h).  One CTA per SM walks its tiles; for each tile
//   phase 1 evaluates every face / bending stencil touching the tile ONCE (one element per thread, FP64, elements.cuh) and
//   parks the element blocks in shared memory; after one barrier, phase 2 lets every output 3x3 block of the tile's CSR rows
//   pull its contributions from shared memory in a fixed order and writes it ONCE to its fixed CSR slot (tile_exec.cuh).
//   M and f come out of the same kernel.  No atomics, no scratch traffic: HBM sees x, X, the plan's index streams and each
//   output value exactly once.  The inputs of the NEXT tile (node table, x / X gathers, template) are staged with cp.async
//   while the current tile computes.  Warp roles: 8 compute warps (216 registers, setmaxnreg) run phases 1 and 2; the service
//   warpgroup (4 warps, 72 registers) stages the next tiles, takes its share of the phase-2 groups, sums the diagonal blocks
//   (phase 3, for plans of mostly full tiles: assemble_tiles_kernel<true>) and issues the bulk copy-out (cp.async.bulk).
// Alternative pipeline "rows" (EOLC_FORCES_PIPELINE=rows, the previous design, kept for A/B measurements): a CTA owns a run
//   of consecutive nodes, one thread per (node, incident element) evaluates that element's block ROW — every element is
//   re-evaluated once per node it touches (3x / 4x), which made the kernel FP64-issue bound (profiles/r01).
// Both are bit-reproducible run to run.
#include "common.h"
#include "elements.cuh"
#include "forces_plan.h"
#include "tile_exec.cuh"
#include "forces_eol.h"
#include "solve.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <numeric>

using namespace eolc;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct eolc_forces_plan {
    eolc_ctx *ctx = nullptr;
    int device = 0;                               // own copy: the plan may be destroyed after its ctx (thread_local order)
    int32_t N = 0, F = 0, E = 0, Ei = 0, dof = 0;
    int64_t nnzM = 0, nnzK = 0;
    // host topology
    std::vector<int32_t> h_face_nodes, h_iedge;   // h_iedge: 4 per INTERIOR edge, ascending mesh edge order
    Pattern pat;                                  // block pattern of M and MDK
    std::vector<int32_t> h_outerM, h_innerM, h_outerK, h_innerK;  // lazily built Eigen-style arrays
    int pipeline = 0;                             // 0 = tiles, 1 = rows
    int max_degM = -1;                            // largest number of column blocks in a row of M (computed on first use by the rhs kernel)
    // EOL branch (forces_eol.h); n_eol == 0: Lagrangian mesh, nothing of this is used
    int32_t n_eol = 0, eol_faces = 0, eol_edges = 0, eol_targets = 0;
    int64_t eol_scratch = 0;
    DevBuf<int32_t> d_eol_faces, d_eol_edges;
    DevBuf<eol::Target> d_eol_targets;
    DevBuf<uint32_t> d_eol_sources;
    DevBuf<double> d_eol_scratch;                 // eol_scratch doubles per scene
    bool smem_attr_set = false;
    bool host_M_valid = false;
    // exact symmetry of MDK (EOLC_FILL_EXACT_SYMMETRY): the pairs (a, b), a < b, whose two nodes are owned by different tiles
    std::vector<int32_t> h_tile_of;               // owner tile per node ("tiles" pipeline)
    DevBuf<uint4> d_sym_pairs;                    // per pair: offset of block (a,b), offset of block (b,a), row strides (a | b << 16), 0
    int64_t n_sym_pairs = -1;                     // -1: not built yet
    std::vector<int64_t> h_dstK;                  // EOL plans: value index of scalar row 3a of every node (rows carry Eulerian columns)
    std::vector<int32_t> h_extraK;                // EOL plans: Eulerian columns per scalar row of the node                    // eolc_forces_fill has produced M on this plan (EOLC_FILL_M_UNCHANGED may skip it)
    // "tiles" pipeline
    int32_t n_tiles = 0, n_templates = 0;
    bool service_p3 = false;
    uint32_t geo16 = 0, tmplA16 = 0, tmplB16 = 0, loc_max = 0, scr_doubles = 0, kstage = 0, mstage = 0;
    int64_t elem_evals = 0, geo_bytes = 0, tmpl_bytes = 0;
    DevBuf<uint4> d_geo, d_tmpl;
    // "rows" pipeline
    int32_t n_cta = 0;
    DevBuf<uint4> d_cta_hdr;                      // 2 x uint4 per CTA (CtaHeader)
    DevBuf<unsigned long long> d_slot;            // per CTA slot (aligned with the block index): packed SlotRec
    DevBuf<uint32_t> d_items;                     // NT slots per CTA: pack_item(local ids, pos), 0 = empty
    DevBuf<int32_t> d_tile_nodes;                 // TILE_NODES slots per CTA: the tile's distinct global node ids
    DevBuf<uint16_t> d_pl;                        // item_local << 2 | column block j
    DevBuf<uint16_t> d_node_f;                    // first face item (CTA-local) | count << 8
    // staging for the host entry point
    DevBuf<double> d_x, d_X, d_f, d_Mv, d_Kv;
    PinnedBuf<double> p_in, p_out;
    // block structure on the device + CG work vectors (solve.cuh), built on first use
    DevBuf<int32_t> d_blkM, d_nbrM, d_blkK, d_nbrK;
    std::vector<int32_t> h_eol_index;             // N, EOL plans only
    DevBuf<int32_t> d_outerM, d_innerM, d_outerK, d_innerK, d_eol_index;   // scalar structure of an EOL plan for the device consumers, built on first use
    DevBuf<int32_t> d_fn, d_nfp;                  // face nodes and node -> incident faces (CSR) for the normals, built on first use
    DevBuf<uint32_t> d_nfl;
    DevBuf<double> d_fnorm, d_nnorm;              // staging of the host entry point
    DevBuf<double> d_cg;                          // r, p, Ap/z, dinv (dof each), partials, scalars
    PinnedBuf<double> p_sc;
#ifdef EOLC_TILE_CLOCKS
    DevBuf<unsigned long long> d_dbg;
    int dbg_grid = 0;
#endif
};

namespace {

// ------------------------------------------------------------------------------------------------
// "tiles" pipeline
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Shared memory: [scr: parked element blocks][staged MDK rows][staged M rows][staged f][template A x 2][template B]
// [geometry x 4][x x 2][X x 2]; work item w = scene * n_tiles + tile.
//   at the start of tile k:  geometry(k+2), template A(k+1) and B(k) (only if they differ from the resident ones), and the
//                            x / X gathers of tile k+1 are issued (cp.async, one group)
//   phase 1(k) -> wait (inputs landed, bulk copy-out(k-1) has read its staging) + barrier -> phase 2(k) -> proxy fence + barrier
//   -> one warp issues the bulk copies of tile k's staged runs (cp.async.bulk shared -> global); they drain during phase 1(k+1)
struct TilesSmem {
    uint32_t scr, kst, mst, fst, tmplA, tmplB, geo, x, X, total;   // byte offsets
};
__host__ __device__ inline TilesSmem tiles_smem_layout(uint32_t scr_doubles, uint32_t kstage, uint32_t mstage, uint32_t tmplA16, uint32_t tmplB16,
                                                       uint32_t geo16, uint32_t loc_max) {
    TilesSmem L;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t at = o; o += (bytes + 15u) & ~15u; return at; };
    // + 2 doubles per staging array: the arrays are shifted by the 16-byte phase of the scene's output pointers
    L.scr = take(8 * scr_doubles); L.kst = take(8 * (kstage + 2)); L.mst = take(8 * (mstage + 2)); L.fst = take(8 * (tiles::MAX_FSTAGE + 2));
    L.tmplA = take(2 * 16 * tmplA16); L.tmplB = take(16 * tmplB16); L.geo = take(4 * 16 * geo16);
    L.x = take(2 * 8 * 3 * loc_max); L.X = take(2 * 8 * 2 * loc_max);
    L.total = o;
    return L;
}

struct TilesArgs {
    uint32_t n_work, n_tiles, step_q, step_r;   // step = gridDim.x = step_q * n_tiles + step_r
    uint32_t geo16, tmplA16, tmplB16, loc_max, scr_doubles, kstage, mstage;
    const uint4 *geo, *tmpl;
    const double *x, *X;
    double *f, *Mv, *Kv;
    size_t x_stride, X_stride, f_stride, M_stride, K_stride;
    tiles::FillParams prm;
    uint32_t skip_m;   // EOLC_FILL_M_UNCHANGED: M rows are neither recomputed (apart from M_aa, which MDK_aa needs) nor written
#ifdef EOLC_TILE_CLOCKS
    unsigned long long *dbg;   // per (CTA, warp): clocks spent in [prefetch issue, phase 1, wait+barrier, phase 2, barrier, copy-out], tiles
#endif
};
#ifdef EOLC_TILE_CLOCKS
#define EOLC_CLK(i) { const long long now_ = clock64(); clk_acc[i] += now_ - clk_last; clk_last = now_; }
// a barrier whose completion the NEXT instruction depends on (a plain __syncthreads() defers the blocking, which would move the wait
// into whatever region follows)
#define EOLC_SYNC() { if (__syncthreads_or(0)) clk_acc[6] += 1000000; }
#else
#define EOLC_CLK(i)
#define EOLC_SYNC() __syncthreads()
#endif

struct BulkStore {   // one asynchronous bulk copy shared -> global; both addresses and the size are multiples of 16 bytes
    __device__ __forceinline__ void operator()(double *dst, const double *src, uint32_t bytes) const {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    }
};

// Register budget: the kernel is launched with 384 threads (ptxas cap 168 registers); the two compute warpgroups then take 216
// registers per thread and the service warpgroup drops to 72 (setmaxnreg): the pool is the CTA's own 384 x 168 allocation, and
// 256 x 216 + 128 x 72 = 64512 is all of it.  ptxas allocates each region within its setmaxnreg value (checked in the SASS: the
// service region, which also runs phase-2 groups, stays below R72 without spills).
#ifndef EOLC_COMPUTE_REGS
#define EOLC_COMPUTE_REGS 216
#endif
#ifndef EOLC_SERVICE_REGS
#define EOLC_SERVICE_REGS 72
#endif
constexpr int COMPUTE_REGS = EOLC_COMPUTE_REGS, SERVICE_REGS = EOLC_SERVICE_REGS;
constexpr int LAUNCH_REGS = (65536 / (tiles::CTA_THREADS * tiles::CTAS_PER_SM)) / 8 * 8;   // what ptxas may use under the launch bounds: 168 for one 384-thread CTA per SM
static_assert(tiles::NTHREADS * COMPUTE_REGS + 128 * SERVICE_REGS <= tiles::CTA_THREADS * LAUNCH_REGS, "register pool of the CTA");
constexpr bool SERVICE_P2 = tiles::P2THREADS > tiles::NTHREADS;   // the service warpgroup runs phase-2 groups too
constexpr int BAR1_THREADS = SERVICE_P2 ? tiles::CTA_THREADS : tiles::NTHREADS;
// Phase 3 (the diagonal blocks from the staged rows) on the service warpgroup: the compute warps go from the end of phase 2 straight
// to phase 1 of the next tile.  Needs the service warps in phase 2 (they then pass the same barrier 1).
// Decided per plan (eolc_forces_plan::service_p3 picks the instantiation): it pays when nearly every tile is a full one (1024^2 sheet: -1.7 %, 0.736 -> 0.724 ms)
// and costs when many tiles are light boundary tiles whose phase 1 is shorter than the service warps' chain (4096 x 64^2 ensemble:
// +2.5 %); profiles/r01/experiments.md p3.  EOLC_FORCES_P3=0|1 overrides.
#if defined(EOLC_TILE_CLOCKS) || defined(EOLC_B2_FULL)
constexpr bool SERVICE_P3_BUILD = false;
#else
constexpr bool SERVICE_P3_BUILD = SERVICE_P2;
#endif

template <bool P3_ARG>
__global__ void __launch_bounds__(tiles::CTA_THREADS, tiles::CTAS_PER_SM) assemble_tiles_kernel(const __grid_constant__ TilesArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t ctl[2];          // service -> compute warps: [0] template A buffer of the NEXT tile, [1] m_full of the current tile
    const TilesSmem L = tiles_smem_layout(A.scr_doubles, A.kstage, A.mstage, A.tmplA16, A.tmplB16, A.geo16, A.loc_max);
    double *scr = reinterpret_cast<double *>(smem_raw + L.scr);
    uint4 *TAst = reinterpret_cast<uint4 *>(smem_raw + L.tmplA);
    uint4 *TBst = reinterpret_cast<uint4 *>(smem_raw + L.tmplB);
    uint4 *Gst = reinterpret_cast<uint4 *>(smem_raw + L.geo);
    double *xst = reinterpret_cast<double *>(smem_raw + L.x);
    double *Xst = reinterpret_cast<double *>(smem_raw + L.X);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t geo16 = A.geo16, tmplA16 = A.tmplA16, loc_max = A.loc_max;
    const uint32_t n_tiles = A.n_tiles, n_work = A.n_work, step = gridDim.x;
    constexpr uint32_t NC = tiles::NTHREADS, NCW = NC / 32;   // compute threads / warps; warps NCW .. NCW+3 are the service warpgroup
    constexpr bool SERVICE_P3 = SERVICE_P3_BUILD && P3_ARG;   // two instantiations; a run-time flag cost the gain (experiments.md p3)
    constexpr int GSTAGES = 4;

    auto geo_hdr = [&](int stage) { return reinterpret_cast<const uint32_t *>(Gst + (uint32_t)stage * geo16); };
    // work items of this CTA: w = blockIdx.x + k * step; (tile, scene) advance without divisions
    auto advance = [&](uint32_t &tile, uint32_t &scene) {
        tile += A.step_r; scene += A.step_q;
        if (tile >= n_tiles) { tile -= n_tiles; ++scene; }
    };
    if (blockIdx.x >= n_work) return;
    uint32_t w = blockIdx.x;
    uint32_t tile0 = w % n_tiles, scene0 = w / n_tiles;      // once per CTA
    tiles::TileView V;
    V.scr = scr;
    V.tmplB = reinterpret_cast<const uint32_t *>(TBst);
    double *const kst16 = reinterpret_cast<double *>(smem_raw + L.kst), *const mst16 = reinterpret_cast<double *>(smem_raw + L.mst),
                 *const fst16 = reinterpret_cast<double *>(smem_raw + L.fst);
    int gs = 0, ta = 0, xs_ = 0;
#ifdef EOLC_TILE_CLOCKS
    long long clk_acc[7] = {0, 0, 0, 0, 0, 0, 0}, clk_last = clock64();
#endif
    auto set_view = [&](double *fs, double *Ms, double *Ks) {
        V.geo = geo_hdr(gs);
        V.tmpl = reinterpret_cast<const uint32_t *>(TAst + (uint32_t)ta * tmplA16);
        V.xs = xst + (uint32_t)xs_ * 3 * loc_max;
        V.Xs = Xst + (uint32_t)xs_ * 2 * loc_max;
        V.kst = kst16 + ((uint32_t)(reinterpret_cast<uintptr_t>(Ks) >> 3) & 1u);
        V.mst = mst16 + ((uint32_t)(reinterpret_cast<uintptr_t>(Ms) >> 3) & 1u);
        V.fst = fst16 + ((uint32_t)(reinterpret_cast<uintptr_t>(fs) >> 3) & 1u);
    };

    if (warp >= NCW) {
        // =============================== service warpgroup ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SERVICE_REGS));
        const uint32_t role = warp - NCW;    // 0: geometry / template blobs + control words, 1: x / X gathers, 2 and 3: bulk copy-out
        auto load_geo = [&](uint32_t tile, int stage) {
            const uint4 *src = A.geo + (size_t)tile * geo16;
            uint4 *dst = Gst + (uint32_t)stage * geo16;
            for (uint32_t i = lane; i < geo16; i += 32) cp16(dst + i, src + i);
        };
        auto load_tmplA = [&](int gstage, int tstage) {
            const uint32_t *g = geo_hdr(gstage);
            const uint4 *src = A.tmpl + g[0];
            uint4 *dst = TAst + (uint32_t)tstage * tmplA16;
            const uint32_t n = g[2] & 0xffffu;
            for (uint32_t i = lane; i < n; i += 32) cp16(dst + i, src + i);
        };
        auto load_tmplB = [&](int gstage) {
            const uint32_t *g = geo_hdr(gstage);
            const uint4 *src = A.tmpl + g[0] + (g[2] & 0xffffu);
            const uint32_t n = g[2] >> 16;
            for (uint32_t i = lane; i < n; i += 32) cp16(TBst + i, src + i);
        };
        auto gather = [&](uint32_t scene, int gstage, int xstage) {
            const uint32_t *g = geo_hdr(gstage);
            const uint32_t nLoc = (g[1] >> 8) & 255u;
            const uint32_t *loc = g + 4;
            const double *xs = A.x + (size_t)scene * A.x_stride, *Xs = A.X + (size_t)scene * A.X_stride;
            double *xd = xst + (uint32_t)xstage * 3 * loc_max, *Xd = Xst + (uint32_t)xstage * 2 * loc_max;
            for (uint32_t l = lane; l < nLoc; l += 32) {
                const size_t gid = loc[l];
                cp8(xd + 3 * l, xs + 3 * gid); cp8(xd + 3 * l + 1, xs + 3 * gid + 1); cp8(xd + 3 * l + 2, xs + 3 * gid + 2);
                cp16(Xd + 2 * l, Xs + 2 * gid);
            }
        };
        uint32_t tile1 = tile0, scene1 = scene0;
        advance(tile1, scene1);
        uint32_t tile2 = tile1, scene2 = scene1;
        advance(tile2, scene2);
        if (role == 0) {
            load_geo(tile0, 0);
            if (w + step < n_work) load_geo(tile1, 1);
            cp_commit(); cp_wait_all();
            __syncwarp();
            load_tmplA(0, 0);
            gather(scene0, 0, 0);
            cp_commit(); cp_wait_all();
        }
        __syncthreads();                 // [P] first tile staged
        // role 0 state: template held by A buffer 0 / 1 and by the B buffer; template / pointer phase the M staging zeros belong to
        uint32_t tA_id0 = geo_hdr(0)[0], tA_id1 = 0xffffffffu, tB_id = 0xffffffffu, tM_id = 0xffffffffu, tM_pb = 2;
        for (; w < n_work; w += step) {
            const int gs1 = (gs + 1) & (GSTAGES - 1), gs2 = (gs + 2) & (GSTAGES - 1);
            double *const fs = A.f + (size_t)scene0 * A.f_stride, *const Ms = A.Mv + (size_t)scene0 * A.M_stride, *const Ks = A.Kv + (size_t)scene0 * A.K_stride;
            int ta1 = ta;
            if (role == 0) {
                // stage reuse: geometry(k+2) overwrites the stage of tile k-2 (its copy-out was issued before the barriers of tile k-1);
                // template B(k) overwrites B(k-1), last read in phase 2(k-1), i.e. before a barrier every thread has passed
                const uint32_t pbM = (uint32_t)(reinterpret_cast<uintptr_t>(Ms) >> 3) & 1u;
                if (w + 2 * step < n_work) load_geo(tile2, gs2);
                if (w + step < n_work) {
                    const uint32_t id1 = geo_hdr(gs1)[0];
                    if (id1 != (ta ? tA_id1 : tA_id0)) {
                        ta1 = ta ^ 1;
                        if (id1 != (ta1 ? tA_id1 : tA_id0)) { load_tmplA(gs1, ta1); if (ta1) tA_id1 = id1; else tA_id0 = id1; }
                    }
                }
                const uint32_t id0 = geo_hdr(gs)[0];
                if (id0 != tB_id) { load_tmplB(gs); tB_id = id0; }
                cp_commit();
                if (lane == 0) { ctl[0] = (uint32_t)ta1; ctl[1] = (id0 != tM_id || pbM != tM_pb) ? 1u : 0u; }
                tM_id = id0; tM_pb = pbM;
                cp_wait_all();
            } else if (role == 1) {
#ifndef EOLC_DEBUG_NO_GATHER
                if (w + step < n_work) gather(scene1, gs1, xs_ ^ 1);
#endif
                cp_commit(); cp_wait_all();
            } else {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // copy-out(k-1) has read its staging
            }
            EOLC_CLK(0)
            EOLC_SYNC();                 // [B1] elements parked, staged inputs landed, staging free
            EOLC_CLK(2)
            ta1 = (int)ctl[0];
            if (SERVICE_P2) {            // this warp's share of the phase-2 groups (forces_plan.h spreads them over all 12 warps)
                set_view(fs, Ms, Ks);
                tiles::phase2((int)tid, tiles::P2THREADS, V, ctl[1] != 0u, A.skip_m != 0u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp's staged rows -> visible to the bulk-copy engine
#ifndef EOLC_TILE_CLOCKS
                asm volatile("bar.sync 1, %0;" ::"n"(BAR1_THREADS) : "memory");
#endif
            }
#ifdef EOLC_TILE_CLOCKS
            if (__syncthreads_or(0) == 12345) clk_acc[6] += 1000000;   // the instrumented build's blocking barrier between phases 2 and 3
#endif
#if defined(EOLC_TILE_CLOCKS) || defined(EOLC_B2_FULL)
            EOLC_SYNC();                 // [B2] staged rows complete
#else
            if (SERVICE_P3) {
                tiles::phase3((int)(tid - NC), tiles::CTA_THREADS - (int)NC, V);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 2, %0;" ::"n"(tiles::CTA_THREADS - tiles::NTHREADS) : "memory");   // [B2] the service warps: diagonal blocks staged
            } else {
                asm volatile("bar.sync 2, %0;" ::"n"(tiles::CTA_THREADS) : "memory");   // [B2] every compute thread has arrived: staged rows complete
            }
#endif
            EOLC_CLK(4)
#ifndef EOLC_DEBUG_NO_COPYOUT
            if (role >= 2) {
                set_view(fs, Ms, Ks);
                tiles::copy_out_runs((int)(2 * lane + (role - 2)), 64, V, fs, Ms, Ks, A.skip_m ? 5u : 7u, BulkStore());
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#endif
            EOLC_CLK(5)
#ifdef EOLC_TILE_CLOCKS
            clk_acc[6] += 1;
#endif
            gs = gs1; ta = ta1; xs_ ^= 1;
            tile0 = tile1; scene0 = scene1; tile1 = tile2; scene1 = scene2;
            advance(tile2, scene2);
        }
        if (role >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last bulk copies
    } else {
        // =============================== compute warpgroups ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(COMPUTE_REGS));
        if (tid < tiles::ZPAD) scr[tid] = 0.0;   // the zero block padded pull entries point at (never written again)
        __syncthreads();                 // [P]
        for (; w < n_work; w += step) {
            double *const fs = A.f + (size_t)scene0 * A.f_stride, *const Ms = A.Mv + (size_t)scene0 * A.M_stride, *const Ks = A.Kv + (size_t)scene0 * A.K_stride;
            set_view(fs, Ms, Ks);
            tiles::phase1((int)tid, (int)NC, V, A.prm);
            EOLC_CLK(1)
            EOLC_SYNC();                 // [B1]
            EOLC_CLK(2)
            const int ta1 = (int)ctl[0];
            tiles::phase2((int)tid, tiles::P2THREADS, V, ctl[1] != 0u, A.skip_m != 0u);
            EOLC_CLK(0)
#ifdef EOLC_TILE_CLOCKS
            if (__syncthreads_or(0) == 12345) clk_acc[6] += 1000000;   // never true; see EOLC_SYNC (service warps take part in this build)
            EOLC_CLK(5)
#else
            // with phase 3 on the service warps nothing orders this warp's staged rows before their bulk copies after this barrier
            if (SERVICE_P3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(BAR1_THREADS) : "memory");   // every warp that ran phase-2 groups: off-diagonal and mass blocks staged
#endif
            if (!SERVICE_P3) tiles::phase3((int)tid, (int)NC, V);
            if (!SERVICE_P3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the bulk-copy engine
            EOLC_CLK(3)
#if defined(EOLC_TILE_CLOCKS) || defined(EOLC_B2_FULL)
            EOLC_SYNC();                 // [B2]
#else
            if (!SERVICE_P3)
            // [B2] arrive only: the compute warps go straight to phase 1 of the next tile (the scratch is free since the named barrier
            // above, the next tile's inputs landed before [B1]); the service warps wait here for the staged rows
            asm volatile("bar.arrive 2, %0;" ::"n"(tiles::CTA_THREADS) : "memory");
#endif
            EOLC_CLK(4)
#ifdef EOLC_TILE_CLOCKS
            clk_acc[6] += 1;
#endif
            gs = (gs + 1) & (GSTAGES - 1); ta = ta1; xs_ ^= 1;
            advance(tile0, scene0);
        }
    }
#ifdef EOLC_TILE_CLOCKS
    if (lane == 0)
        for (int i = 0; i < 7; ++i) A.dbg[((size_t)blockIdx.x * (tiles::CTA_THREADS / 32) + warp) * 7 + i] = (unsigned long long)clk_acc[i];
#endif
}

size_t tiles_smem_bytes(const eolc_forces_plan *P) {
    return tiles_smem_layout(P->scr_doubles, P->kstage, P->mstage, P->tmplA16, P->tmplB16, P->geo16, P->loc_max).total;
}

int build_tiles_plan(eolc_forces_plan *P, const double *X_hint, cudaStream_t st, const tiles::RowLayout *rows, const NodeCSR *csr) {
    tiles::Plan tp;
    const char *dd = getenv("EOLC_FORCES_DEDUP");
    const bool dedup = !(dd && strcmp(dd, "0") == 0);
    if (!tiles::build(P->N, P->F, P->h_face_nodes.data(), P->Ei, P->h_iedge.data(), P->pat, X_hint, dedup, tp, rows, csr)) {
        set_error("tile plan: %s", tp.error.c_str());
        return EOLC_ERR_UNSUPPORTED;
    }
    P->n_tiles = tp.n_tiles; P->n_templates = tp.n_templates;
    P->geo16 = tp.max_geo16; P->tmplA16 = tp.max_tmplA16; P->tmplB16 = tp.max_tmplB16;
    P->loc_max = (tp.max_loc + 1) & ~1u;
    P->scr_doubles = (tp.max_scratch + 1) & ~1u;
    P->kstage = (tp.max_kstage + 1) & ~1u; P->mstage = (tp.max_mstage + 1) & ~1u;
    P->elem_evals = tp.elem_evals;
    {
        P->service_p3 = tiles::mostly_full_tiles(tp);
        const char *ev = getenv("EOLC_FORCES_P3");
        if (ev) P->service_p3 = atoi(ev) != 0;
    }
    P->geo_bytes = (int64_t)tp.geo.size() * 4; P->tmpl_bytes = (int64_t)tp.tmpl.size() * 4;
    P->h_tile_of.assign((size_t)P->N, -1);
    for (int32_t t = 0; t < tp.n_tiles; ++t) {
        const uint32_t *g = tp.geo.data() + (size_t)t * 4 * tp.max_geo16;
        for (uint32_t l = 0, n_own = g[1] & 255u; l < n_own; ++l) P->h_tile_of[g[4 + l]] = t;
    }
    if (tiles_smem_bytes(P) > (size_t)(227 / tiles::CTAS_PER_SM - (tiles::CTAS_PER_SM > 1 ? 1 : 0)) * 1024) { set_error("tile plan needs %zu bytes of shared memory", tiles_smem_bytes(P)); return EOLC_ERR_UNSUPPORTED; }
    static_assert(sizeof(uint4) == 16, "uint4");
    EOLC_CUDA(P->d_geo.alloc(tp.geo.size() / 4));
    EOLC_CUDA(P->d_tmpl.alloc(tp.tmpl.size() / 4));
    if (!tp.geo.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_geo.p, tp.geo.data(), tp.geo.size() * 4, cudaMemcpyHostToDevice, st));
    if (!tp.tmpl.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_tmpl.p, tp.tmpl.data(), tp.tmpl.size() * 4, cudaMemcpyHostToDevice, st));
    EOLC_CUDA(cudaStreamSynchronize(st));   // tp dies at scope exit
    return EOLC_OK;
}


// ------------------------------------------------------------------------------------------------
// "rows" pipeline
// ------------------------------------------------------------------------------------------------
constexpr int FACE_SLOTS = 32;                     // warp 0: (node, face) items
constexpr int EDGE_SLOTS = 64;                     // warps 1-2: (node, interior edge) items
constexpr int NT = FACE_SLOTS + EDGE_SLOTS;        // threads per CTA; kinds are warp aligned, so no warp diverges
constexpr int ITEM_STRIDE = 42;                    // doubles parked per item: block j at 10*j (16 B aligned, 9 used);
                                                   // faces: f at 30..32, t8/12 at 33, t8/24 at 34.  Stride 42 makes the
                                                   // quarter-warp STS.128 pattern bank-conflict free (84 words = 20 mod 32).
#ifndef ROWS_MIN_CTAS
#define ROWS_MIN_CTAS 4
#endif

__device__ __forceinline__ void park_block(double *dst, const blk3 &B) {
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    d2[0] = make_double2(B.m[0], B.m[1]); d2[1] = make_double2(B.m[2], B.m[3]);
    d2[2] = make_double2(B.m[4], B.m[5]); d2[3] = make_double2(B.m[6], B.m[7]);
    dst[8] = B.m[8];
}

// One 32-byte header per CTA: everything later loads depend on, fetched with two 128-bit loads.
struct CtaHeader {
    int32_t node0, nnodes;       // the CTA's run of consecutive nodes
    uint32_t plbase;             // first pull entry in d_pl
    uint16_t npl, nbc;           // pull entries / output blocks of the CTA
    long long kbase, mbase;      // offsets (in doubles) of the CTA's first MDK / M block row
};
static_assert(sizeof(CtaHeader) == 32, "CtaHeader must be 32 bytes");
// One 64-bit record per output block, in the order threads take them (descending contribution count):
//  q0:9 first pull entry | cnt:7 entries | nf:6 leading face entries | diag:1 | koff:11 | deg:8 | hasm:1 | moff:11 | degM:8
// koff/moff: offset of the block's first row from kbase/mbase in units of 3 doubles; rows are 3*deg (3*degM) apart.
__host__ __device__ inline unsigned long long pack_slot(unsigned q0, unsigned cnt, unsigned nf, unsigned diag, unsigned koff,
                                                        unsigned deg, unsigned hasm, unsigned moff, unsigned degM) {
    return (unsigned long long)q0 | ((unsigned long long)cnt << 9) | ((unsigned long long)nf << 16) | ((unsigned long long)diag << 22) |
           ((unsigned long long)koff << 23) | ((unsigned long long)deg << 34) | ((unsigned long long)hasm << 42) |
           ((unsigned long long)moff << 43) | ((unsigned long long)degM << 54);
}

// Persistent, warp-specialised: warps 0-2 (NT threads) compute and assemble; warp 3 is a loader that stages the inputs
// of the tiles AHEAD (ring of RING stages) in shared memory, so the FP64 pipe never waits on HBM/L2 latency.
// A tile's inputs are its (<= TILE_NODES) distinct nodes: the loader reads the tile's node table, then x / X of those
// nodes once (two dependent global latencies); items address them with 6-bit tile-local ids.
//   loader :  for each tile: wait EMPTY[s] -> fetch -> arrive FULL[s]
//   compute:  wait FULL[s] -> phase 1 (rows -> scr) -> bar(compute) -> phase 2 (pull + store) -> arrive EMPTY[s] -> bar(compute)
constexpr int NTHREADS = NT + 32;
constexpr int TILE_NODES = 64;
constexpr int RING = 4;
// item word: local ids n0..n3 (6 bits each, bits 0-23) | pos (bits 24-25) | valid (bit 26)
__host__ __device__ inline uint32_t pack_item(int n0, int n1, int n2, int n3, int pos) {
    return (uint32_t)n0 | ((uint32_t)n1 << 6) | ((uint32_t)n2 << 12) | ((uint32_t)n3 << 18) | ((uint32_t)pos << 24) | (1u << 26);
}

struct TileStage {
    CtaHeader hdr;
    double x[TILE_NODES * 3];
    double X[TILE_NODES * 2];
    uint32_t item[NT];
    uint16_t pl[3 * FACE_SLOTS + 4 * EDGE_SLOTS];
    uint16_t nf[NT];
};

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int BAR_COMPUTE = 1, BAR_FULL0 = 2, BAR_EMPTY0 = 2 + RING;

__device__ __forceinline__ void loader_fetch(int lane, long long tile, int n_cta, const uint4 *__restrict__ cta_hdr,
                                             const uint32_t *__restrict__ items, const int32_t *__restrict__ tile_nodes,
                                             const uint16_t *__restrict__ pl, const uint16_t *__restrict__ node_f,
                                             const double *__restrict__ x, const double *__restrict__ X, size_t x_stride,
                                             size_t X_stride, TileStage *__restrict__ S) {
    const int c = (int)(tile % n_cta);
    const size_t sc = (size_t)(tile / n_cta);
    x += sc * x_stride; X += sc * X_stride;
    // level 1: header, node table, items (all independent)
    const uint4 h0 = cta_hdr[2 * (size_t)c], h1 = cta_hdr[2 * (size_t)c + 1];
    const int ntab = (int)(h0.y >> 8) & 255;
    int g0 = -1, g1 = -1;
    if (lane < ntab) g0 = tile_nodes[(size_t)c * TILE_NODES + lane];
    if (lane + 32 < ntab) g1 = tile_nodes[(size_t)c * TILE_NODES + 32 + lane];
    uint32_t it[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) it[k] = items[(size_t)c * NT + lane + 32 * k];
    CtaHeader H;
    H.node0 = (int)h0.x; H.nnodes = (int)(h0.y & 255); H.plbase = h0.z; H.npl = (uint16_t)(h0.w & 0xffff); H.nbc = (uint16_t)(h0.w >> 16);
    H.kbase = (long long)(((unsigned long long)h1.y << 32) | h1.x); H.mbase = (long long)(((unsigned long long)h1.w << 32) | h1.z);
    // level 2: positions of the tile's distinct nodes; pull entries and per-node face ranges ride along
    double xa[3] = {0, 0, 0}, xb[3] = {0, 0, 0};
    double2 Xa = make_double2(0, 0), Xb = make_double2(0, 0);
    if (g0 >= 0) { const double *p = x + 3 * (size_t)g0; xa[0] = p[0]; xa[1] = p[1]; xa[2] = p[2]; Xa = *reinterpret_cast<const double2 *>(X + 2 * (size_t)g0); }
    if (g1 >= 0) { const double *p = x + 3 * (size_t)g1; xb[0] = p[0]; xb[1] = p[1]; xb[2] = p[2]; Xb = *reinterpret_cast<const double2 *>(X + 2 * (size_t)g1); }
    // pull entries: each tile's segment is padded to 16 bytes by the plan -> two 128-bit loads per lane, issued together
    const uint4 *plsrc = reinterpret_cast<const uint4 *>(pl + H.plbase);
    const int nchunk = (H.npl + 7) >> 3;
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    if (lane < nchunk) c0 = plsrc[lane];
    if (lane + 32 < nchunk) c1 = plsrc[lane + 32];
    uint16_t nf0 = 0, nf1 = 0, nf2 = 0;
    if (lane < H.nnodes) nf0 = node_f[H.node0 + lane];
    if (lane + 32 < H.nnodes) nf1 = node_f[H.node0 + lane + 32];
    if (lane + 64 < H.nnodes) nf2 = node_f[H.node0 + lane + 64];
    if (lane == 0) S->hdr = H;
    reinterpret_cast<uint4 *>(S->pl)[lane] = c0;
    if (lane + 32 < (3 * FACE_SLOTS + 4 * EDGE_SLOTS) / 8) reinterpret_cast<uint4 *>(S->pl)[lane + 32] = c1;
    S->nf[lane] = nf0; S->nf[lane + 32] = nf1; S->nf[lane + 64] = nf2;
#pragma unroll
    for (int k = 0; k < 3; ++k) S->item[lane + 32 * k] = it[k];
    S->x[3 * lane] = xa[0]; S->x[3 * lane + 1] = xa[1]; S->x[3 * lane + 2] = xa[2];
    S->x[3 * (lane + 32)] = xb[0]; S->x[3 * (lane + 32) + 1] = xb[1]; S->x[3 * (lane + 32) + 2] = xb[2];
    reinterpret_cast<double2 *>(S->X)[lane] = Xa; reinterpret_cast<double2 *>(S->X)[lane + 32] = Xb;
}

__global__ void __launch_bounds__(NTHREADS, ROWS_MIN_CTAS) assemble_rows_kernel(
    long long n_tiles, int n_cta, const uint4 *__restrict__ cta_hdr, const uint32_t *__restrict__ items,
    const int32_t *__restrict__ tile_nodes, const unsigned long long *__restrict__ slot, const uint16_t *__restrict__ pl,
    const uint16_t *__restrict__ node_f, const double *__restrict__ x, const double *__restrict__ X, double mu, double lam,
    double rho, double beta, double gx, double gy, double gz, double dhh, double *__restrict__ f, double *__restrict__ Mv,
    double *__restrict__ Kv, size_t x_stride, size_t X_stride, size_t f_stride, size_t M_stride, size_t K_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];      // > 48 KB: dynamic shared memory
    double *scr = reinterpret_cast<double *>(smem_raw);
    TileStage *ring = reinterpret_cast<TileStage *>(scr + ITEM_STRIDE * NT);
    const int t = threadIdx.x;

    if (t >= NT) {
        // ================= loader warp =================
        const int lane = t - NT;
        int s = 0;
        long long k = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            if (k >= RING) bar_sync(BAR_EMPTY0 + s, NTHREADS);          // consumers released this stage
            loader_fetch(lane, tile, n_cta, cta_hdr, items, tile_nodes, pl, node_f, x, X, x_stride, X_stride, &ring[s]);
            __threadfence_block();
            bar_arrive(BAR_FULL0 + s, NTHREADS);
            s = s + 1 == RING ? 0 : s + 1;
        }
        return;
    }

    // ================= compute warps =================
    int s = 0;
    long long kt = 0;
    long long my_tiles = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my_tiles;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++kt) {
        bar_sync(BAR_FULL0 + s, NTHREADS);                              // stage s holds this tile
        const TileStage &St = ring[s];
        const CtaHeader H = St.hdr;
        const uint32_t it = St.item[t];
        const long long sbase = H.kbase / 9;
        unsigned long long rec = 0;
        if (t < H.nbc) rec = slot[sbase + t];                           // latency hides under phase 1
        const size_t sc = (size_t)(tile / n_cta);
        double *fs = f + sc * f_stride, *Ms = Mv + sc * M_stride, *Ks = Kv + sc * K_stride;
        // ---- phase 1: one (node, element) block row per thread -> shared memory
        if (it & (1u << 26)) {
            const int pos = (it >> 24) & 3;
            const double *q0 = St.x + 3 * (it & 63), *q1 = St.x + 3 * ((it >> 6) & 63), *q2 = St.x + 3 * ((it >> 12) & 63);
            const double *Q0 = St.X + 2 * (it & 63), *Q1 = St.X + 2 * ((it >> 6) & 63), *Q2 = St.X + 2 * ((it >> 12) & 63);
            double *dst = scr + ITEM_STRIDE * t;
            if (t >= FACE_SLOTS) {
                const double *q3 = St.x + 3 * ((it >> 18) & 63), *Q3 = St.X + 2 * ((it >> 18) & 63);
                edge_row_emit(pos, mk3(q0[0], q0[1], q0[2]), mk3(q1[0], q1[1], q1[2]), mk3(q2[0], q2[1], q2[2]), mk3(q3[0], q3[1], q3[2]),
                              Q0[0], Q0[1], Q1[0], Q1[1], Q2[0], Q2[1], Q3[0], Q3[1], beta, dhh,
                              [dst](int j, const blk3 &B) { park_block(dst + 10 * j, B); });
            } else {
                FaceRowOut o;
                face_row(pos, mk3(q0[0], q0[1], q0[2]), mk3(q1[0], q1[1], q1[2]), mk3(q2[0], q2[1], q2[2]), Q0[0], Q0[1], Q1[0], Q1[1],
                         Q2[0], Q2[1], mu, lam, rho, mk3(gx, gy, gz), dhh, o);
                park_block(dst, o.K[0]); park_block(dst + 10, o.K[1]); park_block(dst + 20, o.K[2]);
                dst[30] = o.f[0]; dst[31] = o.f[1]; dst[32] = o.f[2]; dst[33] = o.md; dst[34] = o.mo;
            }
        }
        bar_sync(BAR_COMPUTE, NT);
        // ---- phase 2: every output block of the tile's rows pulls its contributions in reference order.
        // Slots are handed out by descending contribution count, so the lanes of a warp run similar trip counts.
        const uint16_t *spl = St.pl;
        for (int k = t; k < H.nbc; k += NT) {
            if (k != t) rec = slot[sbase + k];
            const uint32_t q0 = (uint32_t)(rec & 511), cnt = (uint32_t)(rec >> 9) & 127;
            double a0, a1, a2, a3, a4, a5, a6, a7, a8;
            {
                const uint32_t en = spl[q0];
                const double *src = scr + ITEM_STRIDE * (en >> 2) + 10 * (en & 3);
                const double2 *s2 = reinterpret_cast<const double2 *>(src);
                double2 v0 = s2[0], v1 = s2[1], v2 = s2[2], v3_ = s2[3];
                a0 = v0.x; a1 = v0.y; a2 = v1.x; a3 = v1.y; a4 = v2.x; a5 = v2.y; a6 = v3_.x; a7 = v3_.y; a8 = src[8];
            }
            for (uint32_t p = q0 + 1; p < q0 + cnt; ++p) {
                const uint32_t en = spl[p];
                const double *src = scr + ITEM_STRIDE * (en >> 2) + 10 * (en & 3);
                const double2 *s2 = reinterpret_cast<const double2 *>(src);
                double2 v0 = s2[0], v1 = s2[1], v2 = s2[2], v3_ = s2[3];
                double v8 = src[8];
                a0 = a0 + v0.x; a1 = a1 + v0.y; a2 = a2 + v1.x; a3 = a3 + v1.y; a4 = a4 + v2.x; a5 = a5 + v2.y;
                a6 = a6 + v3_.x; a7 = a7 + v3_.y; a8 = a8 + v8;
            }
            {
                const uint32_t koff = (uint32_t)(rec >> 23) & 2047, deg = (uint32_t)(rec >> 34) & 255;
                double *row = Ks + H.kbase + 3 * (size_t)koff;
                __stcs(row, a0); __stcs(row + 1, a1); __stcs(row + 2, a2);
                row += 3 * deg;
                __stcs(row, a3); __stcs(row + 1, a4); __stcs(row + 2, a5);
                row += 3 * deg;
                __stcs(row, a6); __stcs(row + 1, a7); __stcs(row + 2, a8);
            }
            if ((rec >> 42) & 1) {   // mass: the block's face contributions only (they lead the list); t8/12 on the diagonal
                const int nf = (int)(rec >> 16) & 63, mo_ = ((rec >> 22) & 1) ? 33 : 34;
                double m = scr[ITEM_STRIDE * (spl[q0] >> 2) + mo_];
                for (int r = 1; r < nf; ++r) m = m + scr[ITEM_STRIDE * (spl[q0 + r] >> 2) + mo_];
                const uint32_t moff = (uint32_t)(rec >> 43) & 2047, degM = (uint32_t)(rec >> 54) & 255;
                double *row = Ms + H.mbase + 3 * (size_t)moff;
                __stcs(row, m); __stcs(row + 1, 0.0); __stcs(row + 2, 0.0);
                row += 3 * degM;
                __stcs(row, 0.0); __stcs(row + 1, m); __stcs(row + 2, 0.0);
                row += 3 * degM;
                __stcs(row, 0.0); __stcs(row + 1, 0.0); __stcs(row + 2, m);
            }
        }
        // ---- f: per node, its face items in ascending face order (f.setZero() then +=, Forces.cpp:915,500-502)
        if (t < H.nnodes) {
            const uint32_t nfm = St.nf[t];
            const int first = nfm & 255, cnt = nfm >> 8;
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            for (int k = 0; k < cnt; ++k) {
                const double *src = scr + ITEM_STRIDE * (first + k) + 30;
                f0 += src[0]; f1 += src[1]; f2 += src[2];
            }
            double *dst = fs + 3 * (size_t)(H.node0 + t);
            dst[0] = f0; dst[1] = f1; dst[2] = f2;
        }
        // release the stage to the loader only if it will be refilled (keeps arrive/sync counts matched)
        if (kt + RING < my_tiles) bar_arrive(BAR_EMPTY0 + s, NTHREADS);
        bar_sync(BAR_COMPUTE, NT);                                      // scr is free for the next tile
        s = s + 1 == RING ? 0 : s + 1;
    }
}

int build_rows_plan(eolc_forces_plan *P, cudaStream_t st) {
    const int32_t N = P->N, F = P->F, Ei = P->Ei;
    const int32_t *fn = P->h_face_nodes.data();
    const int32_t *ie = P->h_iedge.data();
    // node -> incident faces / interior edges (ascending element index), with the node's position in the element
    std::vector<int32_t> nfp(N + 1, 0), nep(N + 1, 0);
    for (int64_t i = 0; i < 3 * (int64_t)F; ++i) nfp[fn[i] + 1]++;
    for (int64_t i = 0; i < 4 * (int64_t)Ei; ++i) nep[ie[i] + 1]++;
    for (int32_t a = 0; a < N; ++a) { nfp[a + 1] += nfp[a]; nep[a + 1] += nep[a]; }
    std::vector<uint32_t> nfl(nfp[N]), nel(nep[N]);   // elem << 2 | pos
    {
        std::vector<int32_t> pf(nfp.begin(), nfp.end() - 1), pe(nep.begin(), nep.end() - 1);
        for (int32_t i = 0; i < F; ++i) for (int v = 0; v < 3; ++v) nfl[pf[fn[3 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
        for (int32_t i = 0; i < Ei; ++i) for (int v = 0; v < 4; ++v) nel[pe[ie[4 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
    }
    // ---- tiles: runs of consecutive nodes whose face items fit warp 0, edge items fit warps 1-2 and whose elements touch
    // at most TILE_NODES distinct nodes
    std::vector<int32_t> cta_node0;
    std::vector<int32_t> stamp(N, -1);          // stamp[g] == tile -> g is in the tile's node table
    {
        int cf = 0, ce = 0, nodes = 0, ndist = 0, tile = 0;
        std::vector<int32_t> fresh;
        cta_node0.push_back(0);
        for (int32_t a = 0; a < N; ++a) {
            const int nf = nfp[a + 1] - nfp[a], ne = nep[a + 1] - nep[a];
            auto collect = [&]() {
                fresh.clear();
                auto touch = [&](int32_t g) { if (stamp[g] != tile) { stamp[g] = tile; fresh.push_back(g); } };
                touch(a);
                for (int32_t k = nfp[a]; k < nfp[a + 1]; ++k) for (int j = 0; j < 3; ++j) touch(fn[3 * (size_t)(nfl[k] >> 2) + j]);
                for (int32_t k = nep[a]; k < nep[a + 1]; ++k) for (int j = 0; j < 4; ++j) touch(ie[4 * (size_t)(nel[k] >> 2) + j]);
            };
            collect();
            if (cf + nf > FACE_SLOTS || ce + ne > EDGE_SLOTS || nodes == NT || ndist + (int)fresh.size() > TILE_NODES) {
                if (nodes == 0) {
                    set_error("node %d: %d faces / %d bending stencils / %d stencil nodes exceed the tile limits (%d / %d / %d)", a, nf, ne,
                              (int)fresh.size(), FACE_SLOTS, EDGE_SLOTS, TILE_NODES);
                    return EOLC_ERR_UNSUPPORTED;
                }
                cta_node0.push_back(a);
                ++tile; cf = ce = nodes = ndist = 0;
                collect();
                if (nf > FACE_SLOTS || ne > EDGE_SLOTS || (int)fresh.size() > TILE_NODES) {
                    set_error("node %d: %d faces / %d bending stencils / %d stencil nodes exceed the tile limits (%d / %d / %d)", a, nf, ne,
                              (int)fresh.size(), FACE_SLOTS, EDGE_SLOTS, TILE_NODES);
                    return EOLC_ERR_UNSUPPORTED;
                }
            }
            cf += nf; ce += ne; ++nodes; ndist += (int)fresh.size();
        }
        cta_node0.push_back(N);
    }
    const int32_t nc = N > 0 ? (int32_t)cta_node0.size() - 1 : 0;
    P->n_cta = nc;
    std::vector<uint32_t> items((size_t)nc * NT, 0u);
    std::vector<int32_t> tile_nodes((size_t)nc * TILE_NODES, 0);
    std::vector<uint16_t> pl;
    pl.reserve(9 * (size_t)F + 16 * (size_t)Ei);
    std::vector<uint16_t> node_f(N, 0);
    std::vector<unsigned long long> slots((size_t)P->pat.nblkK, 0);
    std::vector<CtaHeader> hdr((size_t)nc);
    std::vector<std::vector<uint16_t>> tmp;   // per block of the current node
    struct SlotTmp { int cnt; unsigned long long rec; };
    std::vector<SlotTmp> cslots;              // slot records of the current tile
    std::vector<int32_t> local(N, -1);        // global -> tile-local id, valid while lstamp[g] == c
    std::vector<int32_t> lstamp(N, -1);
    for (int32_t c = 0; c < nc; ++c) {
        const int32_t n0 = cta_node0[c], n1 = cta_node0[c + 1];
        int32_t fpos = 0, epos = FACE_SLOTS, ntab = 0;
        auto lid = [&](int32_t g) {
            if (lstamp[g] != c) { lstamp[g] = c; local[g] = ntab; tile_nodes[(size_t)c * TILE_NODES + ntab] = g; ++ntab; }
            return local[g];
        };
        cslots.clear();
        const int64_t cb0 = P->pat.blkptrK[n0], cm0 = P->pat.blkptrM[n0];
        const size_t plbase = pl.size();
        for (int32_t a = n0; a < n1; ++a) {
            const int64_t b0 = P->pat.blkptrK[a], b1 = P->pat.blkptrK[a + 1];
            const int deg = (int)(b1 - b0);
            if (deg > 255) { set_error("node %d has %d neighbours (limit 255)", a, deg); return EOLC_ERR_UNSUPPORTED; }
            tmp.assign(deg, {});
            std::vector<int> nfaces(deg, 0);
            node_f[a] = (uint16_t)(fpos | ((nfp[a + 1] - nfp[a]) << 8));
            for (int32_t k = nfp[a]; k < nfp[a + 1]; ++k) {      // faces ascending
                const int32_t face = nfl[k] >> 2;
                const int32_t *v = fn + 3 * (size_t)face;
                items[(size_t)c * NT + fpos] = pack_item(lid(v[0]), lid(v[1]), lid(v[2]), 0, nfl[k] & 3);
                for (int j = 0; j < 3; ++j) {
                    int p = (int)(find_block(P->pat.blkptrK, P->pat.nbrK, a, v[j]) - b0);
                    tmp[p].push_back((uint16_t)((fpos << 2) | j));
                    nfaces[p]++;
                }
                ++fpos;
            }
            for (int32_t k = nep[a]; k < nep[a + 1]; ++k) {      // then interior edges ascending
                const int32_t ed = nel[k] >> 2;
                const int32_t *v = ie + 4 * (size_t)ed;
                items[(size_t)c * NT + epos] = pack_item(lid(v[0]), lid(v[1]), lid(v[2]), lid(v[3]), nel[k] & 3);
                for (int j = 0; j < 4; ++j) {
                    int p = (int)(find_block(P->pat.blkptrK, P->pat.nbrK, a, v[j]) - b0);
                    tmp[p].push_back((uint16_t)((epos << 2) | j));
                }
                ++epos;
            }
            const int64_t m0 = P->pat.blkptrM[a];
            const int degM = (int)(P->pat.blkptrM[a + 1] - m0);
            for (int p = 0; p < deg; ++p) {
                const int32_t b = P->pat.nbrK[b0 + p];
                const unsigned q0 = (unsigned)(pl.size() - plbase);
                pl.insert(pl.end(), tmp[p].begin(), tmp[p].end());
                unsigned hasm = 0, moff = 0;
                if (nfaces[p] > 0) { hasm = 1; moff = (unsigned)(3 * (m0 - cm0) + (find_block(P->pat.blkptrM, P->pat.nbrM, a, b) - m0)); }
                const unsigned koff = (unsigned)(3 * (b0 - cb0) + p);
                if (q0 > 511 || tmp[p].size() > 127 || nfaces[p] > 63 || koff > 2047 || moff > 2047 || degM > 255) {
                    set_error("internal: slot record overflow at node %d", a);
                    return EOLC_ERR_UNSUPPORTED;
                }
                cslots.push_back({(int)tmp[p].size(), pack_slot(q0, (unsigned)tmp[p].size(), (unsigned)nfaces[p], a == b ? 1u : 0u, koff,
                                                                (unsigned)deg, hasm, moff, hasm ? (unsigned)degM : 0u)});
            }
        }
        if (ntab > TILE_NODES) { set_error("internal: tile node table overflow"); return EOLC_ERR_UNSUPPORTED; }
        const uint16_t npl_true = (uint16_t)(pl.size() - plbase);
        while ((pl.size() - plbase) % 8) pl.push_back(0);   // 16-byte granules for the loader's 128-bit copies
        // descending contribution count, ties keep block order: deterministic
        std::stable_sort(cslots.begin(), cslots.end(), [](const SlotTmp &u, const SlotTmp &v) { return u.cnt > v.cnt; });
        for (size_t k = 0; k < cslots.size(); ++k) slots[cb0 + k] = cslots[k].rec;
        CtaHeader &H = hdr[c];
        H.node0 = n0; H.nnodes = (n1 - n0) | (ntab << 8); H.plbase = (uint32_t)plbase; H.npl = npl_true;
        H.nbc = (uint16_t)cslots.size(); H.kbase = 9 * cb0; H.mbase = 9 * cm0;
        if (pl.size() >= ((size_t)1 << 32)) { set_error("pull list too long"); return EOLC_ERR_UNSUPPORTED; }
    }
    {
        std::vector<uint4> hraw(2 * (size_t)nc);
        static_assert(sizeof(CtaHeader) == 2 * sizeof(uint4), "header layout");
        if (nc) memcpy(hraw.data(), hdr.data(), sizeof(CtaHeader) * (size_t)nc);
        EOLC_CUDA(P->d_cta_hdr.upload(hraw, st));
    }
    EOLC_CUDA(P->d_items.upload(items, st)); EOLC_CUDA(P->d_tile_nodes.upload(tile_nodes, st)); EOLC_CUDA(P->d_pl.upload(pl, st));
    EOLC_CUDA(P->d_slot.upload(slots, st));
    EOLC_CUDA(P->d_node_f.upload(node_f, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    return EOLC_OK;
}


// ------------------------------------------------------------------------------------------------
// EOL branch (forces_eol.h): the elements that touch an EoL node, and the Eulerian rows / columns they feed
// ------------------------------------------------------------------------------------------------
struct EolArgs {
    int32_t n_faces, n_edges, n_targets, n_scenes;
    uint32_t lag_dof;
    const int32_t *faces, *edges;
    const eol::Target *targets;
    const uint32_t *sources;
    const double *x, *X;
    double *scratch, *f, *Mv, *Kv;
    size_t x_stride, X_stride, f_stride, M_stride, K_stride, scratch_stride;
    eol::Params prm;
};

// one thread per (scene, EOL element): the element's Eulerian expansion into its scratch record
__global__ void __launch_bounds__(128) eol_elements_kernel(EolArgs A) {
    const long long per = (long long)A.n_faces + A.n_edges;
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= per * A.n_scenes) return;
    const int s = (int)(w / per), i = (int)(w % per);
    const double *x = A.x + (size_t)s * A.x_stride, *X = A.X + (size_t)s * A.X_stride;
    double *scr = A.scratch + (size_t)s * A.scratch_stride;
    if (i < A.n_faces) eol::face_record(A.faces + 4 * (size_t)i, x, X, A.prm, scr + (size_t)i * eol::FACE_REC);
    else eol::edge_record(A.edges + 8 * (size_t)(i - A.n_faces), x, X, A.prm, scr + (size_t)A.n_faces * eol::FACE_REC + (size_t)(i - A.n_faces) * eol::EDGE_REC);
}

// one thread per (scene, Eulerian entry): fixed-order sum of its contributions, written once
__global__ void __launch_bounds__(128) eol_gather_kernel(EolArgs A) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)A.n_targets * A.n_scenes) return;
    const int s = (int)(w / A.n_targets), t = (int)(w % A.n_targets);
    eol::gather_target(A.targets[t], A.sources, A.scratch + (size_t)s * A.scratch_stride, A.lag_dof, A.f + (size_t)s * A.f_stride,
                       A.Mv + (size_t)s * A.M_stride, A.Kv + (size_t)s * A.K_stride);
}

int build_eol_plan(eolc_forces_plan *P, const int32_t *eol_index, eol::Plan &ep, cudaStream_t st) {
    if (!eol::build(P->N, P->F, P->h_face_nodes.data(), P->Ei, P->h_iedge.data(), eol_index, P->pat.blkptrM, P->pat.nbrM, P->pat.blkptrK, P->pat.nbrK, ep)) {
        set_error("EOL plan: %s", ep.error.c_str());
        return EOLC_ERR_UNSUPPORTED;
    }
    P->n_eol = ep.n_eol; P->dof = ep.dof; P->nnzM = ep.nnzM; P->nnzK = ep.nnzK;
    P->h_eol_index.assign(eol_index, eol_index + P->N);
    P->eol_faces = ep.n_faces(); P->eol_edges = ep.n_edges(); P->eol_targets = (int32_t)ep.targets.size(); P->eol_scratch = ep.scratch_doubles;
    EOLC_CUDA(P->d_eol_faces.alloc(ep.faces.size())); EOLC_CUDA(P->d_eol_edges.alloc(ep.edges.size()));
    EOLC_CUDA(P->d_eol_targets.alloc(ep.targets.size())); EOLC_CUDA(P->d_eol_sources.alloc(ep.sources.size()));
    if (!ep.faces.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_eol_faces.p, ep.faces.data(), ep.faces.size() * 4, cudaMemcpyHostToDevice, st));
    if (!ep.edges.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_eol_edges.p, ep.edges.data(), ep.edges.size() * 4, cudaMemcpyHostToDevice, st));
    if (!ep.targets.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_eol_targets.p, ep.targets.data(), ep.targets.size() * sizeof(eol::Target), cudaMemcpyHostToDevice, st));
    if (!ep.sources.empty()) EOLC_CUDA(cudaMemcpyAsync(P->d_eol_sources.p, ep.sources.data(), ep.sources.size() * 4, cudaMemcpyHostToDevice, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    P->h_outerM.swap(ep.outerM); P->h_innerM.swap(ep.innerM); P->h_outerK.swap(ep.outerK); P->h_innerK.swap(ep.innerK);
    return EOLC_OK;
}

// after the Lagrangian kernel, same stream: records, then the gather (which also adds the bending force onto f)
int launch_eol(eolc_forces_plan *P, int32_t S, const double *x, const double *X, const eolc_material *mat, const double *grav, double dhh,
               double *f, double *Mv, double *Kv) {
    cudaStream_t st = P->ctx->stream;
    EOLC_CUDA(P->d_eol_scratch.ensure((size_t)std::max<int64_t>(P->eol_scratch, 1) * S));
    EolArgs A;
    A.n_faces = P->eol_faces; A.n_edges = P->eol_edges; A.n_targets = P->eol_targets; A.n_scenes = S; A.lag_dof = 3u * (uint32_t)P->N;
    A.faces = P->d_eol_faces.p; A.edges = P->d_eol_edges.p; A.targets = P->d_eol_targets.p; A.sources = P->d_eol_sources.p;
    A.x = x; A.X = X; A.scratch = P->d_eol_scratch.p; A.f = f; A.Mv = Mv; A.Kv = Kv;
    A.x_stride = (size_t)3 * P->N; A.X_stride = (size_t)2 * P->N; A.f_stride = (size_t)P->dof; A.M_stride = (size_t)P->nnzM; A.K_stride = (size_t)P->nnzK;
    A.scratch_stride = (size_t)P->eol_scratch;
    A.prm.e = mat->e; A.prm.nu = mat->nu; A.prm.rho = mat->density; A.prm.beta = mat->beta; A.prm.gx = grav[0]; A.prm.gy = grav[1]; A.prm.gz = grav[2];
    A.prm.dhh = dhh;
    const long long n_el = ((long long)P->eol_faces + P->eol_edges) * S, n_t = (long long)P->eol_targets * S;
    if (n_el > 0) eol_elements_kernel<<<(unsigned)((n_el + 127) / 128), 128, 0, st>>>(A);
    if (n_t > 0) eol_gather_kernel<<<(unsigned)((n_t + 127) / 128), 128, 0, st>>>(A);
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

// ---- exact symmetry of MDK ------------------------------------------------------------------------------------------------------
// The reference pushes every off-diagonal element block twice, (i, j) and mirrored (j, i) (fillxxMI / fillxxB, Forces.cpp:114-125,
// :531-539), so its MDK is symmetric bit for bit.  Here a pair of nodes owned by ONE tile is summed once and mirrored (exact), but the
// two blocks of a pair that straddles two tiles are summed by two CTAs in two orders and agree to rounding only.  This pass copies
// the block of the lower node onto the transposed block of the higher one for those pairs (~30 % of the pairs, 0.1 ms at 1024^2).
__global__ void __launch_bounds__(256) symmetrize_kernel(long long n_pairs, const uint4 *__restrict__ pairs, double *__restrict__ Kv, size_t K_stride) {
    double *K = Kv + (size_t)blockIdx.y * K_stride;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * 256) {
        const uint4 p = pairs[i];
        const double *src = K + p.x;
        double *dst = K + p.y;
        const uint32_t ss = p.z & 0xffffu, ds = p.z >> 16;
        double b[9];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) b[3 * j + k] = src[j * ss + k];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) dst[j * ds + k] = b[3 * k + j];
    }
}

static int ensure_sym_pairs(eolc_forces_plan *P) {
    if (P->n_sym_pairs >= 0) return EOLC_OK;
    // EOL plans: the Lagrangian blocks sit in rows that also carry Eulerian columns (forces_eol.h): first scalar row of node a at
    // h_dstK[a], rows 3 deg + extra apart.  The Eulerian rows / columns come from one source per mirrored pair and are symmetric already.
    const bool eol = P->n_eol != 0;
    if (P->nnzK >= ((int64_t)1 << 32)) { set_error("EOLC_FILL_EXACT_SYMMETRY: matrix too large for 32-bit block offsets"); return EOLC_ERR_UNSUPPORTED; }
    const Pattern &pat = P->pat;
    std::vector<uint4> pairs;
    for (int32_t a = 0; a < P->N; ++a) {
        const int64_t a0 = pat.blkptrK[a], a1 = pat.blkptrK[a + 1];
        const uint32_t dega = (uint32_t)(a1 - a0);
        for (int64_t q = a0; q < a1; ++q) {
            const int32_t b = pat.nbrK[q];
            if (b <= a) continue;
            // "rows" pipeline: every row is summed on its own; "tiles": only pairs across two tiles differ
            if (P->pipeline == 0 && P->h_tile_of[a] == P->h_tile_of[b]) continue;
            const uint32_t degb = (uint32_t)(pat.blkptrK[b + 1] - pat.blkptrK[b]);
            const int64_t qb = find_block(pat.blkptrK, pat.nbrK, b, a);
            const int64_t rowa = eol ? P->h_dstK[a] : 9 * a0, rowb = eol ? P->h_dstK[b] : 9 * pat.blkptrK[b];
            const uint32_t stra = 3u * dega + (eol ? (uint32_t)P->h_extraK[a] : 0u), strb = 3u * degb + (eol ? (uint32_t)P->h_extraK[b] : 0u);
            pairs.push_back(make_uint4((uint32_t)(rowa + 3 * (q - a0)), (uint32_t)(rowb + 3 * (qb - pat.blkptrK[b])), stra | (strb << 16), 0u));
        }
    }
    EOLC_CUDA(P->d_sym_pairs.upload(pairs, P->ctx->stream));
    EOLC_CUDA(cudaStreamSynchronize(P->ctx->stream));
    P->n_sym_pairs = (int64_t)pairs.size();
    return EOLC_OK;
}

static int launch_symmetrize(eolc_forces_plan *P, int32_t S, double *Kv) {
    int rc = ensure_sym_pairs(P);
    if (rc) return rc;
    if (P->n_sym_pairs == 0) return EOLC_OK;
    const int grid = (int)std::min<int64_t>((P->n_sym_pairs + 255) / 256, (int64_t)P->ctx->sm_count * 8);
    symmetrize_kernel<<<dim3(grid, S), 256, 0, P->ctx->stream>>>(P->n_sym_pairs, P->d_sym_pairs.p, Kv, (size_t)P->nnzK);
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

int launch_fill(eolc_forces_plan *P, int32_t S, const double *x, const double *X, const eolc_material *mat,
                const double *grav, double h, double *f, double *Mv, double *Kv, bool skip_m = false) {
    cudaStream_t st = P->ctx->stream;
    const double dhh = mat->dampingB * h * h;   // damping(1)*h*h, Forces.cpp:105
    if (P->N == 0) return EOLC_OK;
    if (P->pipeline == 0) {
        EOLC_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0, "x must be 8-byte and X 16-byte aligned");
        EOLC_REQUIRE(((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(Mv) | reinterpret_cast<uintptr_t>(Kv)) & 7) == 0, "outputs must be 8-byte aligned");
        const long long n_work = (long long)P->n_tiles * S;
        EOLC_REQUIRE(n_work < (1ll << 31), "too many (scene, tile) work items for one launch");
        const int grid = (int)std::min<long long>(n_work, (long long)P->ctx->sm_count * tiles::CTAS_PER_SM);
        const size_t smem = tiles_smem_bytes(P);
        if (!P->smem_attr_set) {
            EOLC_CUDA(cudaFuncSetAttribute(assemble_tiles_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            EOLC_CUDA(cudaFuncSetAttribute(assemble_tiles_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            P->smem_attr_set = true;
        }
        TilesArgs A;
        A.n_work = (uint32_t)n_work; A.n_tiles = (uint32_t)P->n_tiles; A.step_q = (uint32_t)grid / (uint32_t)P->n_tiles; A.step_r = (uint32_t)grid % (uint32_t)P->n_tiles;
        A.geo16 = P->geo16; A.tmplA16 = P->tmplA16; A.tmplB16 = P->tmplB16; A.loc_max = P->loc_max;
        A.scr_doubles = P->scr_doubles; A.kstage = P->kstage; A.mstage = P->mstage; A.geo = P->d_geo.p; A.tmpl = P->d_tmpl.p; A.x = x; A.X = X;
        A.f = f; A.Mv = Mv; A.Kv = Kv; A.x_stride = (size_t)3 * P->N; A.X_stride = (size_t)2 * P->N; A.f_stride = (size_t)P->dof;
        A.M_stride = (size_t)P->nnzM; A.K_stride = (size_t)P->nnzK;
        A.prm.mu = membrane_mu(mat->e, mat->nu); A.prm.lam = membrane_lambda(mat->e, mat->nu); A.prm.rho = mat->density; A.prm.beta = mat->beta;
        A.prm.gx = grav[0]; A.prm.gy = grav[1]; A.prm.gz = grav[2]; A.prm.dhh = dhh;
        A.skip_m = skip_m ? 1u : 0u;
#ifdef EOLC_TILE_CLOCKS
        EOLC_CUDA(P->d_dbg.ensure((size_t)grid * (tiles::CTA_THREADS / 32) * 7));
        P->dbg_grid = grid;
        A.dbg = P->d_dbg.p;
#endif
        if (P->service_p3) assemble_tiles_kernel<true><<<grid, tiles::CTA_THREADS, smem, st>>>(A);
        else assemble_tiles_kernel<false><<<grid, tiles::CTA_THREADS, smem, st>>>(A);
        EOLC_CUDA(cudaGetLastError());
        return P->n_eol ? launch_eol(P, S, x, X, mat, grav, dhh, f, Mv, Kv) : EOLC_OK;
    }
    {
        const long long n_tiles = (long long)P->n_cta * S;
        const int grid = (int)std::min<long long>(n_tiles, (long long)P->ctx->sm_count * ROWS_MIN_CTAS);
        const size_t smem = sizeof(double) * ITEM_STRIDE * NT + RING * sizeof(TileStage);
        if (!P->smem_attr_set) {   // per device; plans are per ctx/device
            EOLC_CUDA(cudaFuncSetAttribute(assemble_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            P->smem_attr_set = true;
        }
        assemble_rows_kernel<<<grid, NTHREADS, smem, st>>>(
            n_tiles, P->n_cta, P->d_cta_hdr.p, P->d_items.p, P->d_tile_nodes.p, P->d_slot.p, P->d_pl.p, P->d_node_f.p, x, X,
            membrane_mu(mat->e, mat->nu), membrane_lambda(mat->e, mat->nu), mat->density, mat->beta, grav[0], grav[1], grav[2],
            dhh, f, Mv, Kv, (size_t)3 * P->N, (size_t)2 * P->N, (size_t)P->dof, (size_t)P->nnzM, (size_t)P->nnzK);
    }
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

}  // namespace

extern "C" {

int eolc_forces_plan_create(eolc_ctx *ctx, int32_t N, int32_t F, const int32_t *face_nodes, int32_t E,
                            const int32_t *edge_stencil, const int32_t *eol_index, const double *X_hint,
                            eolc_forces_plan **out) {
    EOLC_REQUIRE(ctx && out, "ctx/out is NULL");
    *out = nullptr;
    EOLC_REQUIRE(N >= 0 && F >= 0 && E >= 0, "negative size");
    EOLC_REQUIRE(F == 0 || face_nodes, "face_nodes is NULL");
    EOLC_REQUIRE(E == 0 || edge_stencil, "edge_stencil is NULL");
    EOLC_REQUIRE((int64_t)N * 3 < INT32_MAX && (int64_t)F < (1 << 25) && (int64_t)E < (1 << 25), "mesh too large for int32 indexing");
    bool has_eol = false;
    if (eol_index)
        for (int32_t a = 0; a < N; ++a) {
            EOLC_REQUIRE(eol_index[a] >= -1 && eol_index[a] < N, "eol_index out of range");
            has_eol |= eol_index[a] >= 0;
        }
    for (int64_t i = 0; i < 3 * (int64_t)F; ++i) EOLC_REQUIRE(face_nodes[i] >= 0 && face_nodes[i] < N, "face node index out of range");
    for (int32_t i = 0; i < F; ++i) {
        const int32_t *v = face_nodes + 3 * (size_t)i;
        EOLC_REQUIRE(v[0] != v[1] && v[1] != v[2] && v[0] != v[2], "degenerate face (repeated node)");
    }
    EOLC_CUDA(cudaSetDevice(ctx->device));
    const bool timing = getenv("EOLC_PLAN_TIMING") != nullptr;
    auto tlast = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[plan_create] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tlast).count());
        tlast = t;
    };
    eolc_forces_plan *P = new eolc_forces_plan;
    P->ctx = ctx; P->device = ctx->device; P->N = N; P->F = F; P->E = E; P->dof = 3 * N;
    P->h_face_nodes.assign(face_nodes, face_nodes + 3 * (size_t)F);
    P->h_iedge.reserve(4 * (size_t)E);
    for (int32_t e = 0; e < E; ++e) {
        const int32_t *s = edge_stencil + 4 * (size_t)e;
        if (s[2] < 0 || s[3] < 0) continue;  // boundary edge, Forces.cpp:688-690
        for (int v = 0; v < 4; ++v)
            if (s[v] >= N) { delete P; set_error("edge stencil index out of range"); return EOLC_ERR_ARG; }
        if (s[0] < 0 || s[1] < 0) { delete P; set_error("edge stencil index out of range"); return EOLC_ERR_ARG; }
        P->h_iedge.insert(P->h_iedge.end(), s, s + 4);
    }
    P->Ei = (int32_t)(P->h_iedge.size() / 4);
    lap("copy + validate");
    // node -> incident elements, built once: the pattern, the tiles and (later, on demand) the normals read it
    const NodeCSR csr(N, F, P->h_face_nodes.data(), P->Ei, P->h_iedge.data());
    lap("node csr");
    build_pattern(N, F, P->h_face_nodes.data(), P->Ei, P->h_iedge.data(), P->pat, &csr);
    lap("pattern");
    P->nnzM = 9 * P->pat.nblkM; P->nnzK = 9 * P->pat.nblkK;
    if (P->nnzK > (int64_t)INT32_MAX) { delete P; set_error("nnz(MDK) exceeds int32 (Eigen StorageIndex is int)"); return EOLC_ERR_UNSUPPORTED; }
    cudaStream_t st = ctx->stream;
    const char *pe = getenv("EOLC_FORCES_PIPELINE");
    P->pipeline = (pe && strcmp(pe, "rows") == 0) ? 1 : 0;
    int rc;
    if (has_eol) {
        if (P->pipeline != 0) { delete P; set_error("EOL nodes need the tiles pipeline"); return EOLC_ERR_UNSUPPORTED; }
        eol::Plan ep;
        rc = build_eol_plan(P, eol_index, ep, st);
        if (!rc) {
            const tiles::RowLayout rows{ep.dstM.data(), ep.dstK.data(), ep.extraM.data(), ep.extraK.data()};
            P->h_dstK = ep.dstK; P->h_extraK = ep.extraK;
            rc = build_tiles_plan(P, X_hint, st, &rows, &csr);
        }
    } else {
        rc = P->pipeline == 0 ? build_tiles_plan(P, X_hint, st, nullptr, &csr) : build_rows_plan(P, st);
    }
    lap("tiles + upload");
    if (rc) { delete P; return rc; }
    *out = P;
    return EOLC_OK;
}

void eolc_forces_plan_destroy(eolc_forces_plan *plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    delete plan;
}

int eolc_forces_pattern(const eolc_forces_plan *plan, int which, int32_t *dof, int64_t *nnz, const int32_t **outer,
                        const int32_t **inner) {
    EOLC_REQUIRE(plan && (which == 0 || which == 1), "bad arguments");
    eolc_forces_plan *P = const_cast<eolc_forces_plan *>(plan);
    if (dof) *dof = P->dof;
    if (nnz) *nnz = which ? P->nnzK : P->nnzM;
    if (outer || inner) {
        std::vector<int32_t> &o = which ? P->h_outerK : P->h_outerM, &in = which ? P->h_innerK : P->h_innerM;
        if (o.empty()) build_eigen_arrays(P->N, which ? P->pat.blkptrK : P->pat.blkptrM, which ? P->pat.nbrK : P->pat.nbrM, o, in);
        if (outer) *outer = o.data();
        if (inner) *inner = in.data();
    }
    return EOLC_OK;
}

int eolc_forces_counts(const eolc_forces_plan *plan, int32_t *n_faces, int32_t *n_interior_edges) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (n_faces) *n_faces = plan->F;
    if (n_interior_edges) *n_interior_edges = plan->Ei;
    return EOLC_OK;
}

#ifdef EOLC_TILE_CLOCKS
// developer build only (scripts/tile_clocks.py): per (CTA, warp) clock totals of the last fill; returns the number of rows of 7
int eolc_debug_tile_clocks(eolc_forces_plan *plan, unsigned long long *out, int max_rows) {
    const int rows = plan->dbg_grid * (tiles::CTA_THREADS / 32);
    if (rows > max_rows) return -1;
    cudaStreamSynchronize(plan->ctx->stream);
    cudaMemcpy(out, plan->d_dbg.p, (size_t)rows * 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return rows;
}
#endif

// ---- consumer of the fill on the device: right-hand side and the collision-free CG branch (solve.cuh) ----
static int ensure_block_structure(eolc_forces_plan *P) {
    if (P->n_eol) {      // EOL plan: the consumers walk the scalar arrays (solve.cuh, *_csr kernels)
        if (P->d_outerK.p) return EOLC_OK;
        cudaStream_t st = P->ctx->stream;
        EOLC_CUDA(P->d_outerM.upload(P->h_outerM, st)); EOLC_CUDA(P->d_innerM.upload(P->h_innerM, st));
        EOLC_CUDA(P->d_outerK.upload(P->h_outerK, st)); EOLC_CUDA(P->d_innerK.upload(P->h_innerK, st));
        EOLC_CUDA(P->d_eol_index.upload(P->h_eol_index, st));
        EOLC_CUDA(cudaStreamSynchronize(st));
        return EOLC_OK;
    }
    if (P->d_blkK.p || P->N == 0) return EOLC_OK;
    cudaStream_t st = P->ctx->stream;
    std::vector<int32_t> bm(P->pat.blkptrM.begin(), P->pat.blkptrM.end()), bk(P->pat.blkptrK.begin(), P->pat.blkptrK.end());
    EOLC_CUDA(P->d_blkM.upload(bm, st)); EOLC_CUDA(P->d_blkK.upload(bk, st));
    EOLC_CUDA(P->d_nbrM.upload(P->pat.nbrM, st)); EOLC_CUDA(P->d_nbrK.upload(P->pat.nbrK, st));
    EOLC_CUDA(cudaStreamSynchronize(st));      // bm / bk die at scope exit
    return EOLC_OK;
}

int eolc_forces_rhs_dev(eolc_forces_plan *plan, const double *M_vals_dev, const double *f_dev, const double *v_dev, double h, double *b_dev) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (plan->N == 0) return EOLC_OK;
    EOLC_REQUIRE(M_vals_dev && f_dev && v_dev && b_dev, "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    int rc = ensure_block_structure(plan);
    if (rc) return rc;
    if (plan->n_eol) {
        const int gridr = std::min((plan->dof + solve::WARPS - 1) / solve::WARPS, 16 * plan->ctx->sm_count);
        solve::k_rhs_csr<<<gridr, solve::THREADS, 0, plan->ctx->stream>>>(plan->dof, plan->d_outerM.p, plan->d_innerM.p, M_vals_dev, f_dev, v_dev, h, b_dev);
        EOLC_CUDA(cudaGetLastError());
        return EOLC_OK;
    }
    if (plan->max_degM < 0) {
        int m = 0;
        for (int32_t a = 0; a < plan->N; ++a) m = std::max(m, (int)(plan->pat.blkptrM[a + 1] - plan->pat.blkptrM[a]));
        plan->max_degM = m;
    }
    if (plan->max_degM <= 8) {     // 8 lanes per node: four nodes per warp instead of two, the same bits
        const int grid = std::min((plan->N + 4 * solve::WARPS - 1) / (4 * solve::WARPS), 16 * plan->ctx->sm_count);
        solve::k_rhs<8><<<grid, solve::THREADS, 0, plan->ctx->stream>>>(plan->N, plan->d_blkM.p, plan->d_nbrM.p, M_vals_dev, f_dev, v_dev, h, b_dev);
    } else {
        const int grid = std::min((plan->N + 2 * solve::WARPS - 1) / (2 * solve::WARPS), 16 * plan->ctx->sm_count);
        solve::k_rhs<16><<<grid, solve::THREADS, 0, plan->ctx->stream>>>(plan->N, plan->d_blkM.p, plan->d_nbrM.p, M_vals_dev, f_dev, v_dev, h, b_dev);
    }
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

int eolc_forces_integrate_dev(eolc_forces_plan *plan, const double *v_dev, double h, double *x_dev) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (plan->N == 0) return EOLC_OK;
    EOLC_REQUIRE(v_dev && x_dev, "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    const size_t n = (size_t)3 * plan->N;
    const int grid = (int)std::min<size_t>((n + solve::THREADS - 1) / solve::THREADS, (size_t)8 * plan->ctx->sm_count);
    solve::k_integrate<<<grid, solve::THREADS, 0, plan->ctx->stream>>>(n, x_dev, v_dev, h);
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

int eolc_forces_integrate_X_dev(eolc_forces_plan *plan, const double *v_dev, double h, double *X_dev) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (plan->N == 0 || plan->n_eol == 0) return EOLC_OK;
    EOLC_REQUIRE(v_dev && X_dev, "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    int rc = ensure_block_structure(plan);
    if (rc) return rc;
    const int grid = (int)std::min<size_t>(((size_t)plan->N + solve::THREADS - 1) / solve::THREADS, (size_t)8 * plan->ctx->sm_count);
    solve::k_integrate_X<<<grid, solve::THREADS, 0, plan->ctx->stream>>>(plan->N, plan->d_eol_index.p, X_dev, v_dev, h);
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

int eolc_solve_cg_dev(eolc_forces_plan *plan, const double *MDK_vals_dev, const double *b_dev, const unsigned char *fixed_dev, double *v_dev,
                      double tol, int32_t max_iter, int32_t *iters_out, double *rel_resid_out) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (iters_out) *iters_out = 0;
    if (rel_resid_out) *rel_resid_out = 0.0;
    if (plan->N == 0) return EOLC_OK;
    EOLC_REQUIRE(MDK_vals_dev && b_dev && v_dev, "NULL device pointer");
    EOLC_REQUIRE(tol > 0.0 && max_iter >= 0, "tol must be positive and max_iter non-negative");
    eolc_forces_plan *P = plan;
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    int rc = ensure_block_structure(P);
    if (rc) return rc;
    cudaStream_t st = P->ctx->stream;
    const size_t n = (size_t)P->dof;
    const int gridv = (int)std::min<size_t>((n + solve::THREADS - 1) / solve::THREADS, (size_t)8 * P->ctx->sm_count);
    const bool csr = P->n_eol != 0;
    const int gridn = csr ? std::min((P->dof + solve::WARPS - 1) / solve::WARPS, 16 * P->ctx->sm_count)
                          : std::min((P->N + 2 * solve::WARPS - 1) / (2 * solve::WARPS), 16 * P->ctx->sm_count);
    const size_t nparts = (size_t)2 * std::max(gridv, gridn);
    EOLC_CUDA(P->d_cg.ensure(4 * n + nparts + 8));
    EOLC_CUDA(P->p_sc.ensure(8));
    double *r = P->d_cg.p, *p = r + n, *Ap = p + n, *dinv = Ap + n, *part = dinv + n, *sc = part + nparts;
    if (csr) solve::k_cg_init_csr<<<gridv, solve::THREADS, 0, st>>>(P->dof, P->d_outerK.p, P->d_innerK.p, MDK_vals_dev, b_dev, fixed_dev, v_dev, r, p, dinv, part);
    else solve::k_cg_init<<<gridv, solve::THREADS, 0, st>>>(P->N, P->d_blkK.p, P->d_nbrK.p, MDK_vals_dev, b_dev, fixed_dev, v_dev, r, p, dinv, part);
    solve::k_cg_scalars<<<1, solve::THREADS, 0, st>>>(gridv, 2, part, sc, 0, tol);
    int it = 0;
    const int check_every = 8;      // the convergence flag lives on the device; the host looks at it every few iterations
    EOLC_CUDA(cudaMemcpyAsync(P->p_sc.p, sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    bool done = P->p_sc.p[6] != 0.0;   // zero (or already tiny) right-hand side: x = 0, like Eigen's early return
    while (!done && it < max_iter) {
        const int batch = std::min(check_every, max_iter - it);
        for (int k = 0; k < batch; ++k) {
            if (csr) solve::k_cg_ap_csr<<<gridn, solve::THREADS, 0, st>>>(P->dof, P->d_outerK.p, P->d_innerK.p, MDK_vals_dev, p, dinv, Ap, part, sc);
            else solve::k_cg_ap<<<gridn, solve::THREADS, 0, st>>>(P->N, P->d_blkK.p, P->d_nbrK.p, MDK_vals_dev, p, dinv, Ap, part, sc);
            solve::k_cg_scalars<<<1, solve::THREADS, 0, st>>>(gridn, 1, part, sc, 1, tol);
            solve::k_cg_update<<<gridv, solve::THREADS, 0, st>>>(n, v_dev, r, p, Ap, dinv, part, sc);
            solve::k_cg_scalars<<<1, solve::THREADS, 0, st>>>(gridv, 2, part, sc, 2, tol);
            solve::k_cg_dir<<<gridv, solve::THREADS, 0, st>>>(n, p, Ap, sc);
        }
        it += batch;
        EOLC_CUDA(cudaMemcpyAsync(P->p_sc.p, sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        EOLC_CUDA(cudaStreamSynchronize(st));
        done = P->p_sc.p[6] != 0.0;
    }
    EOLC_CUDA(cudaGetLastError());
    if (iters_out) *iters_out = it;      // upper bound to a multiple of the check interval: iterations after convergence are no-ops
    const double rhs2 = P->p_sc.p[5] / (tol * tol);
    if (rel_resid_out) *rel_resid_out = rhs2 > 0.0 ? std::sqrt(P->p_sc.p[2] / rhs2) : 0.0;
    return EOLC_OK;
}

// ---- per-step derived mesh data: face and node normals, kernels in solve.cuh ----
static int ensure_face_csr(eolc_forces_plan *P) {
    if (P->d_nfp.p || P->N == 0) return EOLC_OK;
    cudaStream_t st = P->ctx->stream;
    const tiles::NodeCSR csr(P->N, P->F, P->h_face_nodes.data(), 0, nullptr);
    EOLC_CUDA(P->d_fn.upload(P->h_face_nodes, st)); EOLC_CUDA(P->d_nfp.upload(csr.nfp, st)); EOLC_CUDA(P->d_nfl.upload(csr.nfl, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    return EOLC_OK;
}

int eolc_mesh_normals_dev(eolc_forces_plan *plan, const double *x_dev, double *face_n_dev, double *node_n_dev) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (plan->N == 0) return EOLC_OK;
    EOLC_REQUIRE(x_dev, "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    int rc = ensure_face_csr(plan);
    if (rc) return rc;
    cudaStream_t st = plan->ctx->stream;
    const size_t cap = (size_t)8 * plan->ctx->sm_count;
    if (face_n_dev && plan->F > 0)
        solve::k_face_normals<<<(int)std::min<size_t>(((size_t)plan->F + solve::THREADS - 1) / solve::THREADS, cap), solve::THREADS, 0, st>>>(plan->F, plan->d_fn.p, x_dev, face_n_dev);
    if (node_n_dev)
        solve::k_node_normals<<<(int)std::min<size_t>(((size_t)plan->N + solve::THREADS - 1) / solve::THREADS, cap), solve::THREADS, 0, st>>>(plan->N, plan->d_nfp.p, plan->d_nfl.p, plan->d_fn.p, x_dev, node_n_dev);
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

int eolc_mesh_normals(eolc_forces_plan *plan, const double *x, double *face_n, double *node_n) {
    EOLC_REQUIRE(plan, "plan is NULL");
    eolc_forces_plan *P = plan;
    if (P->N == 0) return EOLC_OK;
    EOLC_REQUIRE(x, "NULL host pointer");
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    const size_t N = P->N, F = P->F;
    EOLC_CUDA(P->d_x.ensure(3 * N)); EOLC_CUDA(P->d_fnorm.ensure(3 * std::max<size_t>(F, 1))); EOLC_CUDA(P->d_nnorm.ensure(3 * N));
    EOLC_CUDA(cudaMemcpyAsync(P->d_x.p, x, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = eolc_mesh_normals_dev(P, P->d_x.p, face_n ? P->d_fnorm.p : nullptr, node_n ? P->d_nnorm.p : nullptr);
    if (rc) return rc;
    if (face_n && F) EOLC_CUDA(cudaMemcpyAsync(face_n, P->d_fnorm.p, 3 * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (node_n) EOLC_CUDA(cudaMemcpyAsync(node_n, P->d_nnorm.p, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    return EOLC_OK;
}

int eolc_forces_launches_per_fill(const eolc_forces_plan *plan) { return !plan ? 0 : plan->n_eol ? 3 : 1; }

// M depends only on X and the density (ComputeInertial.cpp:33,44-47): for a Lagrangian mesh it is constant between remeshes.
static bool honours_m_unchanged(const eolc_forces_plan *P, uint32_t flags) {
    return (flags & EOLC_FILL_M_UNCHANGED) && P->pipeline == 0 && P->n_eol == 0;
}

int eolc_forces_fill_batched_dev_ex(eolc_forces_plan *plan, int32_t n_scenes, const double *x_dev, const double *X_dev,
                                    const eolc_material *mat, const double grav[3], double h, double *f_dev,
                                    double *M_vals_dev, double *MDK_vals_dev, uint32_t flags) {
    EOLC_REQUIRE(plan && mat && grav, "NULL argument");
    EOLC_REQUIRE(n_scenes >= 1, "n_scenes must be >= 1");
    EOLC_REQUIRE((flags & ~(EOLC_FILL_M_UNCHANGED | EOLC_FILL_EXACT_SYMMETRY)) == 0, "unknown flag");
    EOLC_REQUIRE(plan->N == 0 || (x_dev && X_dev && f_dev), "NULL device pointer");
    EOLC_REQUIRE(plan->nnzM == 0 || (M_vals_dev && MDK_vals_dev), "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    int rc = launch_fill(plan, n_scenes, x_dev, X_dev, mat, grav, h, f_dev, M_vals_dev, MDK_vals_dev, honours_m_unchanged(plan, flags));
    if (rc == EOLC_OK && (flags & EOLC_FILL_EXACT_SYMMETRY) && plan->N) rc = launch_symmetrize(plan, n_scenes, MDK_vals_dev);
    return rc;
}

int eolc_forces_fill_batched_dev(eolc_forces_plan *plan, int32_t n_scenes, const double *x_dev, const double *X_dev,
                                 const eolc_material *mat, const double grav[3], double h, double *f_dev,
                                 double *M_vals_dev, double *MDK_vals_dev) {
    return eolc_forces_fill_batched_dev_ex(plan, n_scenes, x_dev, X_dev, mat, grav, h, f_dev, M_vals_dev, MDK_vals_dev, 0u);
}

int eolc_forces_fill_dev(eolc_forces_plan *plan, const double *x_dev, const double *X_dev, const eolc_material *mat,
                         const double grav[3], double h, double *f_dev, double *M_vals_dev, double *MDK_vals_dev) {
    return eolc_forces_fill_batched_dev(plan, 1, x_dev, X_dev, mat, grav, h, f_dev, M_vals_dev, MDK_vals_dev);
}

int eolc_forces_fill(eolc_forces_plan *plan, const double *x, const double *X, const eolc_material *mat,
                     const double grav[3], double h, double *f, double *M_vals, double *MDK_vals) {
    return eolc_forces_fill_ex(plan, x, X, mat, grav, h, f, M_vals, MDK_vals, 0u);
}

int eolc_forces_fill_ex(eolc_forces_plan *plan, const double *x, const double *X, const eolc_material *mat,
                        const double grav[3], double h, double *f, double *M_vals, double *MDK_vals, uint32_t flags) {
    EOLC_REQUIRE(plan && mat && grav, "NULL argument");
    EOLC_REQUIRE((flags & ~(EOLC_FILL_M_UNCHANGED | EOLC_FILL_EXACT_SYMMETRY)) == 0, "unknown flag");
    eolc_forces_plan *P = plan;
    // M unchanged: valid only if this plan's device copy of M was produced by an earlier full fill
    const bool skip_m = honours_m_unchanged(P, flags) && P->host_M_valid;
    EOLC_REQUIRE(P->N == 0 || (x && X && f), "NULL host pointer");
    EOLC_REQUIRE(P->nnzM == 0 || (M_vals && MDK_vals), "NULL host pointer");
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    const size_t N = P->N, nf = (size_t)P->dof;
    if (N == 0) return EOLC_OK;
    EOLC_CUDA(P->d_x.ensure(3 * N)); EOLC_CUDA(P->d_X.ensure(2 * N)); EOLC_CUDA(P->d_f.ensure(nf));
    EOLC_CUDA(P->d_Mv.ensure(P->nnzM)); EOLC_CUDA(P->d_Kv.ensure(P->nnzK));
    // pinned (or registered) caller buffers are DMA'd directly; pageable ones go through the plan's pinned staging
    auto pinned = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool in_pinned = pinned(x) && pinned(X);
    const bool out_pinned = pinned(f) && (skip_m || pinned(M_vals)) && pinned(MDK_vals);
    const double *hx = x, *hX = X;
    if (!in_pinned) {
        EOLC_CUDA(P->p_in.ensure(5 * N));
        memcpy(P->p_in.p, x, 3 * N * sizeof(double));
        if (!skip_m) memcpy(P->p_in.p + 3 * N, X, 2 * N * sizeof(double));
        hx = P->p_in.p; hX = P->p_in.p + 3 * N;
    }
    EOLC_CUDA(cudaMemcpyAsync(P->d_x.p, hx, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
    // M unchanged == X unchanged (the flag's contract): the plan's device copy of X is the previous host fill's, no second upload
    if (!skip_m) EOLC_CUDA(cudaMemcpyAsync(P->d_X.p, hX, 2 * N * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = launch_fill(P, 1, P->d_x.p, P->d_X.p, mat, grav, h, P->d_f.p, P->d_Mv.p, P->d_Kv.p, skip_m);
    if (rc) return rc;
    // the host entry always hands out an exactly symmetric MDK, like the reference's (0.3 ms next to the copies)
    rc = launch_symmetrize(P, 1, P->d_Kv.p);
    if (rc) return rc;
    double *hf = f, *hM = M_vals, *hK = MDK_vals;
    if (!out_pinned) {
        EOLC_CUDA(P->p_out.ensure(nf + P->nnzM + P->nnzK));
        hf = P->p_out.p; hM = P->p_out.p + nf; hK = P->p_out.p + nf + P->nnzM;
    }
    EOLC_CUDA(cudaMemcpyAsync(hf, P->d_f.p, nf * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!skip_m) EOLC_CUDA(cudaMemcpyAsync(hM, P->d_Mv.p, P->nnzM * sizeof(double), cudaMemcpyDeviceToHost, st));
    // the big copy goes out in two halves on two streams: two copy engines keep the link a few per cent busier than one (54.7 ->
    // 57.1 GB/s on the 980 MB of the 1024^2 sheet, scripts/micro/d2h_bw.py)
    cudaStream_t st2 = P->ctx->copy_stream;
    const size_t half = (P->nnzK > ((size_t)1 << 22) && st2 && P->ctx->copy_event) ? (P->nnzK / 2) & ~(size_t)511 : 0;
    if (half) {
        EOLC_CUDA(cudaEventRecord(P->ctx->copy_event, st));
        EOLC_CUDA(cudaStreamWaitEvent(st2, P->ctx->copy_event, 0));
        EOLC_CUDA(cudaMemcpyAsync(hK + half, P->d_Kv.p + half, (P->nnzK - half) * sizeof(double), cudaMemcpyDeviceToHost, st2));
    }
    EOLC_CUDA(cudaMemcpyAsync(hK, P->d_Kv.p, (half ? half : P->nnzK) * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    if (half) EOLC_CUDA(cudaStreamSynchronize(st2));
    if (!out_pinned) {
        memcpy(f, hf, nf * sizeof(double));
        if (!skip_m) memcpy(M_vals, hM, P->nnzM * sizeof(double));
        memcpy(MDK_vals, hK, P->nnzK * sizeof(double));
    }
    P->host_M_valid = true;
    return EOLC_OK;
}

}  // extern "C"