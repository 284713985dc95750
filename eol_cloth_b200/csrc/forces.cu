// forces.cu — Forces::fill on the GPU (include/eolc.h, "Forces::fill" section).
//
// Replaces /root/reference/src/Forces.cpp:912-930 (fill), :331-520 (faceBasedF, non-EOL branch),
// :685-910 (edgeBasedF, non-EOL branch) and Eigen's setFromTriplets.
//
// Default pipeline "rows" (one kernel, no HBM scratch) — assemble_rows_kernel:
//   owner-computes: a CTA owns a run of consecutive nodes and their CSR rows. One thread per (node, incident element)
//   item evaluates that element's block ROW for the node (face_row / edge_row, elements.cuh) and parks it in shared
//   memory; after one barrier every output 3x3 block of the CTA's rows sums its contributions from shared memory in the
//   reference's triplet insertion order (faces ascending, then edges ascending, Forces.cpp:922-923), left to right like
//   Eigen's collapseDuplicates, and is written ONCE to its fixed CSR slot. M and f come out of the same kernel.
//   No atomics, no scratch traffic: HBM sees x, X, the plan's index streams and each output value exactly once.
// Legacy pipeline "scratch" (EOLC_FORCES_PIPELINE=scratch, kept for A/B measurements): element-per-thread kernels write
//   element blocks to an HBM scratch, three gather kernels pull them into the CSR slots (3.5x the algorithmic traffic).
// Both are bit-reproducible run to run.
#include "common.h"
#include "elements.cuh"
#include <algorithm>
#include <cstdlib>
#include <numeric>

using namespace eolc;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct eolc_forces_plan {
    eolc_ctx *ctx = nullptr;
    int32_t N = 0, F = 0, E = 0, Ei = 0, dof = 0;
    int64_t nnzM = 0, nnzK = 0, nblkM = 0, nblkK = 0;
    // host topology
    std::vector<int32_t> h_face_nodes, h_iedge;   // h_iedge: 4 per INTERIOR edge, ascending mesh edge order
    std::vector<int64_t> h_blkptrM, h_blkptrK;    // N+1: first block of node a
    std::vector<int32_t> h_nbrM, h_nbrK;          // neighbour node per block (ascending within a node)
    std::vector<int32_t> h_outerM, h_innerM, h_outerK, h_innerK;  // lazily built Eigen-style arrays
    // device topology
    DevBuf<int32_t> d_face_nodes, d_iedge;
    DevBuf<int64_t> d_blkptrM, d_blkptrK;
    DevBuf<int32_t> d_blknodeM, d_blknodeK;       // owning (row) node of each block
    DevBuf<int32_t> d_nbrM;                       // column node of each M block
    DevBuf<int64_t> d_cptrM, d_cptrK, d_cptrF;    // contribution list offsets per block / per node
    DevBuf<int32_t> d_contribM, d_contribK, d_contribF;
    // "rows" pipeline
    int pipeline = 0;                             // 0 = rows, 1 = scratch
    int32_t n_cta = 0;
    bool smem_attr_set = false;
    DevBuf<uint4> d_cta_hdr;                      // 2 x uint4 per CTA (CtaHeader)
    DevBuf<unsigned long long> d_slot;            // per CTA slot (aligned with the block index): packed SlotRec
    DevBuf<uint32_t> d_items;                     // NT slots per CTA: pack_item(local ids, pos), 0 = empty
    DevBuf<int32_t> d_tile_nodes;                 // TILE_NODES slots per CTA: the tile's distinct global node ids
    DevBuf<uint16_t> d_pl;                        // item_local << 2 | column block j
    DevBuf<uint16_t> d_node_f;                    // first face item (CTA-local) | count << 8
    // scratch (per scene chunk)
    DevBuf<double> d_face_scr, d_edge_scr;
    int32_t scratch_scenes = 0;
    // staging for the host entry point
    DevBuf<double> d_x, d_X, d_f, d_Mv, d_Kv;
    PinnedBuf<double> p_in, p_out;
};

namespace {

constexpr int FACE_SCR = 64;   // doubles per face: 54 K + 9 f + 1 t8
constexpr int EDGE_SCR = 90;   // doubles per interior edge: 10 blocks x 9

// contribution code: bits 0-3 local block id, bit 4 transposed, bit 5 edge (else face), bits 6.. element index
__host__ __device__ inline int32_t mk_code(int32_t elem, int is_edge, int blk, int tr) {
    return (elem << 6) | (is_edge << 5) | (tr << 4) | blk;
}

__global__ void __launch_bounds__(128) face_kernel(int F, const int32_t *__restrict__ fn, const double *__restrict__ x,
                                                   const double *__restrict__ X, double e, double nu, double rho,
                                                   double gx, double gy, double gz, double dhh, double *__restrict__ scr,
                                                   size_t x_stride, size_t X_stride, size_t scr_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    const int s = blockIdx.y;
    x += s * x_stride; X += s * X_stride; scr += s * scr_stride;
    int a = fn[3 * i], b = fn[3 * i + 1], c = fn[3 * i + 2];
    FaceOut o;
    face_element(mk3(x[3 * a], x[3 * a + 1], x[3 * a + 2]), mk3(x[3 * b], x[3 * b + 1], x[3 * b + 2]),
                 mk3(x[3 * c], x[3 * c + 1], x[3 * c + 2]), X[2 * a], X[2 * a + 1], X[2 * b], X[2 * b + 1], X[2 * c],
                 X[2 * c + 1], e, nu, rho, mk3(gx, gy, gz), dhh, o);
    // SoA: scr[k*F + i]
#pragma unroll
    for (int bk = 0; bk < 6; ++bk)
#pragma unroll
        for (int k = 0; k < 9; ++k) scr[(size_t)(bk * 9 + k) * F + i] = o.K[bk].m[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        scr[(size_t)(54 + k) * F + i] = o.fa[k];
        scr[(size_t)(57 + k) * F + i] = o.fb[k];
        scr[(size_t)(60 + k) * F + i] = o.fc[k];
    }
    scr[(size_t)63 * F + i] = o.t8;
}

__global__ void __launch_bounds__(128) edge_kernel(int Ei, const int32_t *__restrict__ st, const double *__restrict__ x,
                                                   const double *__restrict__ X, double beta, double dhh,
                                                   double *__restrict__ scr, size_t x_stride, size_t X_stride,
                                                   size_t scr_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ei) return;
    const int s = blockIdx.y;
    x += s * x_stride; X += s * X_stride; scr += s * scr_stride;
    int n0 = st[4 * i], n1 = st[4 * i + 1], n2 = st[4 * i + 2], n3 = st[4 * i + 3];
    EdgeOut o;
    edge_element(mk3(x[3 * n0], x[3 * n0 + 1], x[3 * n0 + 2]), mk3(x[3 * n1], x[3 * n1 + 1], x[3 * n1 + 2]),
                 mk3(x[3 * n2], x[3 * n2 + 1], x[3 * n2 + 2]), mk3(x[3 * n3], x[3 * n3 + 1], x[3 * n3 + 2]), X[2 * n0],
                 X[2 * n0 + 1], X[2 * n1], X[2 * n1 + 1], X[2 * n2], X[2 * n2 + 1], X[2 * n3], X[2 * n3 + 1], beta, dhh, o);
#pragma unroll
    for (int bk = 0; bk < 10; ++bk)
#pragma unroll
        for (int k = 0; k < 9; ++k) scr[(size_t)(bk * 9 + k) * Ei + i] = o.K[bk].m[k];
}

// one thread per MDK block (a, p): vals[rowstart(3a+j) + 3p + k], rowstart(3a+j) = 9*blkptr[a] + j*3*deg(a)
__global__ void __launch_bounds__(256) gather_mdk(int64_t nblk, const int32_t *__restrict__ blknode,
                                                  const int64_t *__restrict__ blkptr, const int64_t *__restrict__ cptr,
                                                  const int32_t *__restrict__ contrib, int F, int Ei,
                                                  const double *__restrict__ fscr, const double *__restrict__ escr,
                                                  double *__restrict__ vals, size_t fscr_stride, size_t escr_stride,
                                                  size_t vals_stride) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblk) return;
    const int s = blockIdx.y;
    fscr += s * fscr_stride; escr += s * escr_stride; vals += s * vals_stride;
    double acc[9];
    bool first = true;
    for (int64_t c = cptr[t]; c < cptr[t + 1]; ++c) {
        int32_t code = contrib[c];
        int blk = code & 15, tr = (code >> 4) & 1, is_edge = (code >> 5) & 1;
        int32_t el = code >> 6;
        const double *src = is_edge ? escr + (size_t)(blk * 9) * Ei + el : fscr + (size_t)(blk * 9) * F + el;
        size_t st = is_edge ? (size_t)Ei : (size_t)F;
        double v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = src[k * st];
        if (tr) { double q; q = v[1]; v[1] = v[3]; v[3] = q; q = v[2]; v[2] = v[6]; v[6] = q; q = v[5]; v[5] = v[7]; v[7] = q; }
        if (first) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = v[k];
            first = false;
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = acc[k] + v[k];
        }
    }
    int a = blknode[t];
    int64_t b0 = blkptr[a];
    int deg = (int)(blkptr[a + 1] - b0);
    int p = (int)(t - b0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double *row = vals + 9 * b0 + (int64_t)j * 3 * deg + 3 * p;
        row[0] = acc[3 * j]; row[1] = acc[3 * j + 1]; row[2] = acc[3 * j + 2];
    }
}

// one thread per M block: sum over faces of t8/12 (a == b) or t8/24, on the block diagonal; explicit zeros elsewhere
__global__ void __launch_bounds__(256) gather_m(int64_t nblk, const int32_t *__restrict__ blknode,
                                                const int64_t *__restrict__ blkptr, const int32_t *__restrict__ nbr,
                                                const int64_t *__restrict__ cptr, const int32_t *__restrict__ contrib,
                                                int F, const double *__restrict__ fscr, double *__restrict__ vals,
                                                size_t fscr_stride, size_t vals_stride) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblk) return;
    const int s = blockIdx.y;
    fscr += s * fscr_stride; vals += s * vals_stride;
    int a = blknode[t];
    const bool diag = nbr[t] == a;
    double acc = 0.0;
    bool first = true;
    for (int64_t c = cptr[t]; c < cptr[t + 1]; ++c) {
        double t8 = fscr[(size_t)63 * F + contrib[c]];
        double m = diag ? t8 / 12.0 : t8 / 24.0;
        acc = first ? m : acc + m;
        first = false;
    }
    int64_t b0 = blkptr[a];
    int deg = (int)(blkptr[a + 1] - b0);
    int p = (int)(t - b0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double *row = vals + 9 * b0 + (int64_t)j * 3 * deg + 3 * p;
        row[0] = j == 0 ? acc : 0.0; row[1] = j == 1 ? acc : 0.0; row[2] = j == 2 ? acc : 0.0;
    }
}

// one thread per node: f[3a..3a+2] = sum over incident faces (ascending) of (fm + fi) of that vertex
__global__ void __launch_bounds__(256) gather_f(int N, const int64_t *__restrict__ cptr, const int32_t *__restrict__ contrib,
                                                int F, const double *__restrict__ fscr, double *__restrict__ f,
                                                size_t fscr_stride, size_t f_stride) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const int s = blockIdx.y;
    fscr += s * fscr_stride; f += s * f_stride;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;   // f.setZero() then +=  (Forces.cpp:915, :500-502)
    for (int64_t c = cptr[a]; c < cptr[a + 1]; ++c) {
        int32_t code = contrib[c];
        int lv = code & 3;
        int32_t face = code >> 2;
        const double *src = fscr + (size_t)(54 + 3 * lv) * F + face;
        f0 += src[0]; f1 += src[(size_t)F]; f2 += src[2 * (size_t)F];
    }
    f[3 * a] = f0; f[3 * a + 1] = f1; f[3 * a + 2] = f2;
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

inline int64_t find_block(const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr, int32_t a, int32_t b);

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
    std::vector<unsigned long long> slots((size_t)P->nblkK, 0);
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
        const int64_t cb0 = P->h_blkptrK[n0], cm0 = P->h_blkptrM[n0];
        const size_t plbase = pl.size();
        for (int32_t a = n0; a < n1; ++a) {
            const int64_t b0 = P->h_blkptrK[a], b1 = P->h_blkptrK[a + 1];
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
                    int p = (int)(find_block(P->h_blkptrK, P->h_nbrK, a, v[j]) - b0);
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
                    int p = (int)(find_block(P->h_blkptrK, P->h_nbrK, a, v[j]) - b0);
                    tmp[p].push_back((uint16_t)((epos << 2) | j));
                }
                ++epos;
            }
            const int64_t m0 = P->h_blkptrM[a];
            const int degM = (int)(P->h_blkptrM[a + 1] - m0);
            for (int p = 0; p < deg; ++p) {
                const int32_t b = P->h_nbrK[b0 + p];
                const unsigned q0 = (unsigned)(pl.size() - plbase);
                pl.insert(pl.end(), tmp[p].begin(), tmp[p].end());
                unsigned hasm = 0, moff = 0;
                if (nfaces[p] > 0) { hasm = 1; moff = (unsigned)(3 * (m0 - cm0) + (find_block(P->h_blkptrM, P->h_nbrM, a, b) - m0)); }
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

int build_pattern(eolc_forces_plan *P) {
    const int32_t N = P->N, F = P->F;
    const int32_t Ei = P->Ei;
    const int32_t *fn = P->h_face_nodes.data();
    const int32_t *ie = P->h_iedge.data();
    // adjacency via counting
    std::vector<int64_t> cntM(N + 1, 0), cntK(N + 1, 0);
    for (int32_t a = 0; a < N; ++a) { cntM[a + 1] = 1; cntK[a + 1] = 1; }  // self (isolated nodes get no block: fix below)
    std::vector<char> used(N, 0);
    for (int32_t i = 0; i < F; ++i)
        for (int v = 0; v < 3; ++v) { used[fn[3 * i + v]] = 1; cntM[fn[3 * i + v] + 1] += 2; cntK[fn[3 * i + v] + 1] += 2; }
    for (int32_t i = 0; i < Ei; ++i)
        for (int v = 0; v < 4; ++v) cntK[ie[4 * i + v] + 1] += 3;
    for (int32_t a = 0; a < N; ++a) if (!used[a]) { cntM[a + 1] = 0; cntK[a + 1] = 0; }
    for (int32_t a = 0; a < N; ++a) { cntM[a + 1] += cntM[a]; cntK[a + 1] += cntK[a]; }
    std::vector<int32_t> rawM(cntM[N]), rawK(cntK[N]);
    std::vector<int64_t> pM(cntM.begin(), cntM.end() - 1), pK(cntK.begin(), cntK.end() - 1);
    for (int32_t a = 0; a < N; ++a) if (used[a]) { rawM[pM[a]++] = a; rawK[pK[a]++] = a; }
    for (int32_t i = 0; i < F; ++i)
        for (int v = 0; v < 3; ++v) {
            int32_t a = fn[3 * i + v];
            for (int w = 0; w < 3; ++w) if (w != v) { rawM[pM[a]++] = fn[3 * i + w]; rawK[pK[a]++] = fn[3 * i + w]; }
        }
    for (int32_t i = 0; i < Ei; ++i)
        for (int v = 0; v < 4; ++v) {
            int32_t a = ie[4 * i + v];
            for (int w = 0; w < 4; ++w) if (w != v) rawK[pK[a]++] = ie[4 * i + w];
        }
    auto compress = [&](std::vector<int64_t> &cnt, std::vector<int32_t> &raw, std::vector<int64_t> &blkptr, std::vector<int32_t> &nbr) {
        blkptr.assign(N + 1, 0);
        nbr.clear();
        nbr.reserve(raw.size() / 2);
        for (int32_t a = 0; a < N; ++a) {
            auto b = raw.begin() + cnt[a], e = raw.begin() + cnt[a + 1];
            std::sort(b, e);
            auto u = std::unique(b, e);
            nbr.insert(nbr.end(), b, u);
            blkptr[a + 1] = (int64_t)nbr.size();
        }
    };
    compress(cntM, rawM, P->h_blkptrM, P->h_nbrM);
    compress(cntK, rawK, P->h_blkptrK, P->h_nbrK);
    P->nblkM = P->h_blkptrM[N]; P->nblkK = P->h_blkptrK[N];
    P->nnzM = 9 * P->nblkM; P->nnzK = 9 * P->nblkK;
    if (P->nnzK > (int64_t)INT32_MAX) { set_error("nnz(MDK) exceeds int32 (Eigen StorageIndex is int)"); return EOLC_ERR_UNSUPPORTED; }
    return EOLC_OK;
}

inline int64_t find_block(const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr, int32_t a, int32_t b) {
    auto beg = nbr.begin() + blkptr[a], end = nbr.begin() + blkptr[a + 1];
    return std::lower_bound(beg, end, b) - nbr.begin();
}

// local block id of (vi, vj), vi <= vj, in the element's emitted block order
const int kFaceBlk[3][3] = {{0, 3, 4}, {3, 1, 5}, {4, 5, 2}};
const int kEdgeBlk[4][4] = {{0, 4, 5, 6}, {4, 1, 7, 8}, {5, 7, 2, 9}, {6, 8, 9, 3}};

int build_contribs(eolc_forces_plan *P, cudaStream_t st) {
    const int32_t N = P->N, F = P->F, Ei = P->Ei;
    const int32_t *fn = P->h_face_nodes.data();
    const int32_t *ie = P->h_iedge.data();
    // ---- MDK: count then fill; element order = faces ascending then interior edges ascending
    std::vector<int64_t> cK(P->nblkK + 1, 0), cM(P->nblkM + 1, 0), cF(N + 1, 0);
    for (int32_t i = 0; i < F; ++i)
        for (int v = 0; v < 3; ++v) {
            cF[fn[3 * i + v] + 1]++;
            for (int w = 0; w < 3; ++w) {
                cK[find_block(P->h_blkptrK, P->h_nbrK, fn[3 * i + v], fn[3 * i + w]) + 1]++;
                cM[find_block(P->h_blkptrM, P->h_nbrM, fn[3 * i + v], fn[3 * i + w]) + 1]++;
            }
        }
    for (int32_t i = 0; i < Ei; ++i)
        for (int v = 0; v < 4; ++v)
            for (int w = 0; w < 4; ++w) cK[find_block(P->h_blkptrK, P->h_nbrK, ie[4 * i + v], ie[4 * i + w]) + 1]++;
    for (int64_t b = 0; b < P->nblkK; ++b) cK[b + 1] += cK[b];
    for (int64_t b = 0; b < P->nblkM; ++b) cM[b + 1] += cM[b];
    for (int32_t a = 0; a < N; ++a) cF[a + 1] += cF[a];
    std::vector<int32_t> lK(cK[P->nblkK]), lM(cM[P->nblkM]), lF(cF[N]);
    std::vector<int64_t> pK(cK.begin(), cK.end() - 1), pM(cM.begin(), cM.end() - 1), pF(cF.begin(), cF.end() - 1);
    for (int32_t i = 0; i < F; ++i)
        for (int v = 0; v < 3; ++v) {
            lF[pF[fn[3 * i + v]]++] = (i << 2) | v;
            for (int w = 0; w < 3; ++w) {
                int lo = v < w ? v : w, hi = v < w ? w : v;
                int64_t bk = find_block(P->h_blkptrK, P->h_nbrK, fn[3 * i + v], fn[3 * i + w]);
                lK[pK[bk]++] = mk_code(i, 0, kFaceBlk[lo][hi], v > w ? 1 : 0);
                int64_t bm = find_block(P->h_blkptrM, P->h_nbrM, fn[3 * i + v], fn[3 * i + w]);
                lM[pM[bm]++] = i;
            }
        }
    for (int32_t i = 0; i < Ei; ++i)
        for (int v = 0; v < 4; ++v)
            for (int w = 0; w < 4; ++w) {
                int lo = v < w ? v : w, hi = v < w ? w : v;
                int64_t bk = find_block(P->h_blkptrK, P->h_nbrK, ie[4 * i + v], ie[4 * i + w]);
                lK[pK[bk]++] = mk_code(i, 1, kEdgeBlk[lo][hi], v > w ? 1 : 0);
            }
    std::vector<int32_t> bnM(P->nblkM), bnK(P->nblkK);
    for (int32_t a = 0; a < N; ++a) {
        for (int64_t b = P->h_blkptrM[a]; b < P->h_blkptrM[a + 1]; ++b) bnM[b] = a;
        for (int64_t b = P->h_blkptrK[a]; b < P->h_blkptrK[a + 1]; ++b) bnK[b] = a;
    }
    EOLC_CUDA(P->d_cptrK.upload(cK, st)); EOLC_CUDA(P->d_contribK.upload(lK, st));
    EOLC_CUDA(P->d_cptrM.upload(cM, st)); EOLC_CUDA(P->d_contribM.upload(lM, st));
    EOLC_CUDA(P->d_cptrF.upload(cF, st)); EOLC_CUDA(P->d_contribF.upload(lF, st));
    EOLC_CUDA(P->d_blknodeM.upload(bnM, st)); EOLC_CUDA(P->d_blknodeK.upload(bnK, st));
    EOLC_CUDA(P->d_blkptrM.upload(P->h_blkptrM, st)); EOLC_CUDA(P->d_blkptrK.upload(P->h_blkptrK, st));
    EOLC_CUDA(P->d_nbrM.upload(P->h_nbrM, st));
    EOLC_CUDA(cudaStreamSynchronize(st));  // host vectors die at scope exit
    return EOLC_OK;
}

void build_eigen_arrays(int32_t N, const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr,
                        std::vector<int32_t> &outer, std::vector<int32_t> &inner) {
    outer.assign(3 * (size_t)N + 1, 0);
    inner.resize(9 * (size_t)blkptr[N]);
    for (int32_t a = 0; a < N; ++a) {
        int64_t b0 = blkptr[a];
        int deg = (int)(blkptr[a + 1] - b0);
        for (int j = 0; j < 3; ++j) {
            int64_t rs = 9 * b0 + (int64_t)j * 3 * deg;
            outer[3 * (size_t)a + j] = (int32_t)rs;
            for (int p = 0; p < deg; ++p)
                for (int k = 0; k < 3; ++k) inner[rs + 3 * p + k] = 3 * nbr[b0 + p] + k;
        }
    }
    outer[3 * (size_t)N] = (int32_t)(9 * blkptr[N]);
}

int ensure_scratch(eolc_forces_plan *P, int32_t scenes) {
    if (scenes <= P->scratch_scenes) return EOLC_OK;
    EOLC_CUDA(P->d_face_scr.alloc((size_t)scenes * FACE_SCR * P->F));
    EOLC_CUDA(P->d_edge_scr.alloc((size_t)scenes * EDGE_SCR * P->Ei));
    P->scratch_scenes = scenes;
    return EOLC_OK;
}

int launch_fill(eolc_forces_plan *P, int32_t S, const double *x, const double *X, const eolc_material *mat,
                const double *grav, double h, double *f, double *Mv, double *Kv) {
    cudaStream_t st = P->ctx->stream;
    const double dhh = mat->dampingB * h * h;   // damping(1)*h*h, Forces.cpp:105
    if (P->pipeline == 0) {
        if (P->N == 0) return EOLC_OK;
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
    // scene chunks bounded by the scratch budget (~4 GB)
    size_t per_scene = ((size_t)FACE_SCR * P->F + (size_t)EDGE_SCR * P->Ei) * sizeof(double);
    int32_t chunk = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)S, ((size_t)4 << 30) / std::max<size_t>(per_scene, 1)));
    chunk = std::min<int32_t>(chunk, 65535);
    int rc = ensure_scratch(P, chunk);
    if (rc) return rc;
    const size_t fs = (size_t)FACE_SCR * P->F, es = (size_t)EDGE_SCR * P->Ei;
    for (int32_t s0 = 0; s0 < S; s0 += chunk) {
        int32_t sc = std::min(chunk, S - s0);
        const double *xs = x + (size_t)s0 * 3 * P->N, *Xs = X + (size_t)s0 * 2 * P->N;
        if (P->F > 0)
            face_kernel<<<dim3((P->F + 127) / 128, sc), 128, 0, st>>>(P->F, P->d_face_nodes.p, xs, Xs, mat->e, mat->nu, mat->density,
                                                                      grav[0], grav[1], grav[2], dhh, P->d_face_scr.p,
                                                                      (size_t)3 * P->N, (size_t)2 * P->N, fs);
        if (P->Ei > 0)
            edge_kernel<<<dim3((P->Ei + 127) / 128, sc), 128, 0, st>>>(P->Ei, P->d_iedge.p, xs, Xs, mat->beta, dhh, P->d_edge_scr.p,
                                                                       (size_t)3 * P->N, (size_t)2 * P->N, es);
        if (P->nblkK > 0)
            gather_mdk<<<dim3((unsigned)((P->nblkK + 255) / 256), sc), 256, 0, st>>>(
                P->nblkK, P->d_blknodeK.p, P->d_blkptrK.p, P->d_cptrK.p, P->d_contribK.p, P->F, P->Ei, P->d_face_scr.p,
                P->d_edge_scr.p, Kv + (size_t)s0 * P->nnzK, fs, es, (size_t)P->nnzK);
        if (P->nblkM > 0)
            gather_m<<<dim3((unsigned)((P->nblkM + 255) / 256), sc), 256, 0, st>>>(
                P->nblkM, P->d_blknodeM.p, P->d_blkptrM.p, P->d_nbrM.p, P->d_cptrM.p, P->d_contribM.p, P->F, P->d_face_scr.p,
                Mv + (size_t)s0 * P->nnzM, fs, (size_t)P->nnzM);
        if (P->N > 0)
            gather_f<<<dim3((P->N + 255) / 256, sc), 256, 0, st>>>(P->N, P->d_cptrF.p, P->d_contribF.p, P->F, P->d_face_scr.p,
                                                                   f + (size_t)s0 * P->dof, fs, (size_t)P->dof);
    }
    EOLC_CUDA(cudaGetLastError());
    return EOLC_OK;
}

}  // namespace

extern "C" {

int eolc_forces_plan_create(eolc_ctx *ctx, int32_t N, int32_t F, const int32_t *face_nodes, int32_t E,
                            const int32_t *edge_stencil, const int32_t *eol_index, const double *X_hint,
                            eolc_forces_plan **out) {
    (void)X_hint;
    EOLC_REQUIRE(ctx && out, "ctx/out is NULL");
    *out = nullptr;
    EOLC_REQUIRE(N >= 0 && F >= 0 && E >= 0, "negative size");
    EOLC_REQUIRE(F == 0 || face_nodes, "face_nodes is NULL");
    EOLC_REQUIRE(E == 0 || edge_stencil, "edge_stencil is NULL");
    EOLC_REQUIRE((int64_t)N * 3 < INT32_MAX && (int64_t)F < (1 << 25) && (int64_t)E < (1 << 25), "mesh too large for int32 indexing");
    if (eol_index)
        for (int32_t a = 0; a < N; ++a)
            if (eol_index[a] >= 0) {
                set_error("node %d is an EOL node: only the Lagrangian branch of Forces::fill is implemented", a);
                return EOLC_ERR_UNSUPPORTED;
            }
    for (int64_t i = 0; i < 3 * (int64_t)F; ++i) EOLC_REQUIRE(face_nodes[i] >= 0 && face_nodes[i] < N, "face node index out of range");
    for (int32_t i = 0; i < F; ++i) {
        const int32_t *v = face_nodes + 3 * (size_t)i;
        EOLC_REQUIRE(v[0] != v[1] && v[1] != v[2] && v[0] != v[2], "degenerate face (repeated node)");
    }
    EOLC_CUDA(cudaSetDevice(ctx->device));
    eolc_forces_plan *P = new eolc_forces_plan;
    P->ctx = ctx; P->N = N; P->F = F; P->E = E; P->dof = 3 * N;
    P->h_face_nodes.assign(face_nodes, face_nodes + 3 * (size_t)F);
    for (int32_t e = 0; e < E; ++e) {
        const int32_t *s = edge_stencil + 4 * (size_t)e;
        if (s[2] < 0 || s[3] < 0) continue;  // boundary edge, Forces.cpp:688-690
        for (int v = 0; v < 4; ++v)
            if (s[v] >= N) { delete P; set_error("edge stencil index out of range"); return EOLC_ERR_ARG; }
        if (s[0] < 0 || s[1] < 0) { delete P; set_error("edge stencil index out of range"); return EOLC_ERR_ARG; }
        P->h_iedge.insert(P->h_iedge.end(), s, s + 4);
    }
    P->Ei = (int32_t)(P->h_iedge.size() / 4);
    int rc = build_pattern(P);
    if (rc) { delete P; return rc; }
    cudaStream_t st = ctx->stream;
    cudaError_t ce = P->d_face_nodes.upload(P->h_face_nodes, st);
    if (ce == cudaSuccess) ce = P->d_iedge.upload(P->h_iedge, st);
    if (ce != cudaSuccess) { delete P; set_error("upload failed: %s", cudaGetErrorString(ce)); return EOLC_ERR_CUDA; }
    const char *pe = getenv("EOLC_FORCES_PIPELINE");
    P->pipeline = (pe && strcmp(pe, "scratch") == 0) ? 1 : 0;
    rc = P->pipeline == 0 ? build_rows_plan(P, st) : build_contribs(P, st);
    if (rc) { delete P; return rc; }
    *out = P;
    return EOLC_OK;
}

void eolc_forces_plan_destroy(eolc_forces_plan *plan) {
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
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
        if (o.empty()) build_eigen_arrays(P->N, which ? P->h_blkptrK : P->h_blkptrM, which ? P->h_nbrK : P->h_nbrM, o, in);
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

int eolc_forces_launches_per_fill(const eolc_forces_plan *plan) { return plan ? (plan->pipeline == 0 ? 1 : 5) : 0; }

int eolc_forces_fill_batched_dev(eolc_forces_plan *plan, int32_t n_scenes, const double *x_dev, const double *X_dev,
                                 const eolc_material *mat, const double grav[3], double h, double *f_dev,
                                 double *M_vals_dev, double *MDK_vals_dev) {
    EOLC_REQUIRE(plan && mat && grav, "NULL argument");
    EOLC_REQUIRE(n_scenes >= 1, "n_scenes must be >= 1");
    EOLC_REQUIRE(plan->N == 0 || (x_dev && X_dev && f_dev), "NULL device pointer");
    EOLC_REQUIRE(plan->nnzM == 0 || (M_vals_dev && MDK_vals_dev), "NULL device pointer");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    return launch_fill(plan, n_scenes, x_dev, X_dev, mat, grav, h, f_dev, M_vals_dev, MDK_vals_dev);
}

int eolc_forces_fill_dev(eolc_forces_plan *plan, const double *x_dev, const double *X_dev, const eolc_material *mat,
                         const double grav[3], double h, double *f_dev, double *M_vals_dev, double *MDK_vals_dev) {
    return eolc_forces_fill_batched_dev(plan, 1, x_dev, X_dev, mat, grav, h, f_dev, M_vals_dev, MDK_vals_dev);
}

int eolc_forces_fill(eolc_forces_plan *plan, const double *x, const double *X, const eolc_material *mat,
                     const double grav[3], double h, double *f, double *M_vals, double *MDK_vals) {
    EOLC_REQUIRE(plan && mat && grav, "NULL argument");
    eolc_forces_plan *P = plan;
    EOLC_REQUIRE(P->N == 0 || (x && X && f), "NULL host pointer");
    EOLC_REQUIRE(P->nnzM == 0 || (M_vals && MDK_vals), "NULL host pointer");
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    const size_t N = P->N;
    if (N == 0) return EOLC_OK;
    EOLC_CUDA(P->d_x.ensure(3 * N)); EOLC_CUDA(P->d_X.ensure(2 * N)); EOLC_CUDA(P->d_f.ensure(3 * N));
    EOLC_CUDA(P->d_Mv.ensure(P->nnzM)); EOLC_CUDA(P->d_Kv.ensure(P->nnzK));
    // pinned (or registered) caller buffers are DMA'd directly; pageable ones go through the plan's pinned staging
    auto pinned = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool in_pinned = pinned(x) && pinned(X);
    const bool out_pinned = pinned(f) && pinned(M_vals) && pinned(MDK_vals);
    const double *hx = x, *hX = X;
    if (!in_pinned) {
        EOLC_CUDA(P->p_in.ensure(5 * N));
        memcpy(P->p_in.p, x, 3 * N * sizeof(double));
        memcpy(P->p_in.p + 3 * N, X, 2 * N * sizeof(double));
        hx = P->p_in.p; hX = P->p_in.p + 3 * N;
    }
    EOLC_CUDA(cudaMemcpyAsync(P->d_x.p, hx, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
    EOLC_CUDA(cudaMemcpyAsync(P->d_X.p, hX, 2 * N * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = launch_fill(P, 1, P->d_x.p, P->d_X.p, mat, grav, h, P->d_f.p, P->d_Mv.p, P->d_Kv.p);
    if (rc) return rc;
    double *hf = f, *hM = M_vals, *hK = MDK_vals;
    if (!out_pinned) {
        EOLC_CUDA(P->p_out.ensure(3 * N + P->nnzM + P->nnzK));
        hf = P->p_out.p; hM = P->p_out.p + 3 * N; hK = P->p_out.p + 3 * N + P->nnzM;
    }
    EOLC_CUDA(cudaMemcpyAsync(hf, P->d_f.p, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaMemcpyAsync(hM, P->d_Mv.p, P->nnzM * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaMemcpyAsync(hK, P->d_Kv.p, P->nnzK * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    if (!out_pinned) {
        memcpy(f, hf, 3 * N * sizeof(double));
        memcpy(M_vals, hM, P->nnzM * sizeof(double));
        memcpy(MDK_vals, hK, P->nnzK * sizeof(double));
    }
    return EOLC_OK;
}

}  // extern "C"
