// cd.cu — collision narrow phase on the GPU (include/eolc.h, "CD / CD2" section).
//
// Replaces /root/reference/src/Collisions.cpp:11-78 (CD, CD2) and src/boxTriCollision.cpp:617-1063
// (boxTriCollision), :1067-1224 (pointTriCollision), :141-231 (createEdges), raytri.cpp:260-316.
//
// Structure of one run (S scenes, P obstacle points, B boxes):
//   cd_prepare      perturbed verts (:648-659), perturbed + unperturbed face normals (:233-247, :193-214),
//                   per-scene cloth AABB (:425-433)
//   "sections"      every ordered piece of the reference's contact list is a section of items:
//                     PE  cloth vertex  vs obstacle points   (pointTriCollision :1099-1139, CD only)
//                     PT  obstacle point vs cloth triangles  (:1142-1213)  warp-cooperative lexicographic argmin
//                     A   cloth vertex  vs 24 box triangles  (:675-764)
//                     Bc  box corner    vs cloth triangles   (:771-845)    warp-cooperative lexicographic argmin
//                     C   cloth edge    vs 12 box edges      (:849-1017)
//                   pass 1 stores a small per-item result (winner id / 12-bit hit mask) and per-256-item counts,
//                   one exclusive scan over the counts gives every item's slot in the FINAL list order
//                   (scene, PE, PT, box0{A,Bc,C}, box1{...}), pass 2 re-derives the records of the hits and writes
//                   them to their slots (ballot/popc + warp-shuffle prefix inside a block). No atomics.
//   host post-pass  step (D) corner de-duplication (:1022-1052, literal forward swap-delete) and the CD index
//                   remap (Collisions.cpp:39-48) run over the (small) contact list while it is copied out.
//
// O(1) obstacle setup (createBox: 14 verts, 24 face normals, vertex normals with acos, :400-420) is evaluated on
// the host per call with libm, exactly as the reference does, and shipped to the device as a constant table.
#include "common.h"
#include "cd_math.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace eolc;
using namespace cdm;

namespace {

// ---- static box tables (data of boxTriCollision.cpp:289-398) --------------------------------------
__constant__ int c_faces1[24][3] = {{0, 8, 2},  {1, 8, 0},  {3, 8, 1},  {2, 8, 3},  {4, 10, 0}, {5, 10, 4}, {1, 10, 5}, {0, 10, 1},
                                    {6, 9, 4},  {7, 9, 6},  {5, 9, 7},  {4, 9, 5},  {2, 11, 6}, {3, 11, 2}, {7, 11, 3}, {6, 11, 7},
                                    {1, 13, 3}, {5, 13, 1}, {7, 13, 5}, {3, 13, 7}, {0, 12, 4}, {2, 12, 0}, {6, 12, 2}, {4, 12, 6}};
__constant__ int c_edgeVerts1[12][4] = {{0, 1, 8, 10}, {2, 0, 8, 12}, {1, 3, 8, 13}, {3, 2, 8, 11}, {0, 4, 10, 12}, {5, 1, 10, 13},
                                        {4, 5, 10, 9}, {6, 2, 11, 12}, {4, 6, 9, 12}, {3, 7, 11, 13}, {7, 5, 9, 13}, {6, 7, 9, 11}};
__constant__ int c_vertEdges1[8][3] = {{0, 1, 4}, {0, 2, 5}, {1, 3, 7}, {2, 3, 9}, {4, 6, 8}, {5, 6, 10}, {7, 8, 11}, {9, 10, 11}};
__constant__ int c_edgeFaces1[12][2] = {{1, 7}, {0, 21}, {2, 16}, {3, 13}, {4, 20}, {6, 17}, {5, 11}, {12, 22}, {8, 23}, {14, 19}, {10, 18}, {9, 15}};

const double h_unitVerts[14][3] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1},
                                   {1, 1, 1},    {-1, 0, 0},  {1, 0, 0},   {0, -1, 0}, {0, 1, 0},   {0, 0, -1}, {0, 0, 1}};
const int h_faces1[24][3] = {{0, 8, 2},  {1, 8, 0},  {3, 8, 1},  {2, 8, 3},  {4, 10, 0}, {5, 10, 4}, {1, 10, 5}, {0, 10, 1},
                             {6, 9, 4},  {7, 9, 6},  {5, 9, 7},  {4, 9, 5},  {2, 11, 6}, {3, 11, 2}, {7, 11, 3}, {6, 11, 7},
                             {1, 13, 3}, {5, 13, 1}, {7, 13, 5}, {3, 13, 7}, {0, 12, 4}, {2, 12, 0}, {6, 12, 2}, {4, 12, 6}};
const int h_edgeVerts1[12][4] = {{0, 1, 8, 10}, {2, 0, 8, 12}, {1, 3, 8, 13}, {3, 2, 8, 11}, {0, 4, 10, 12}, {5, 1, 10, 13},
                                 {4, 5, 10, 9}, {6, 2, 11, 12}, {4, 6, 9, 12}, {3, 7, 11, 13}, {7, 5, 9, 13}, {6, 7, 9, 11}};
const int h_edgeFaces1[12][2] = {{1, 7}, {0, 21}, {2, 16}, {3, 13}, {4, 20}, {6, 17}, {5, 11}, {12, 22}, {8, 23}, {14, 19}, {10, 18}, {9, 15}};

// Per-box constants produced by createBox (:400-420) and hoisted loop invariants of :849-915.
struct BoxData {
    double verts1[14][3];
    double faceNors1[24][3];
    double vertNors1[14][3];
    double aabbB1[6];
    double aabbF1[24][6];
    double edgeAngle[12];    // e1->angle
    double angleCD[12];      // acos(n1c.n1d)  (:906)
    double dx1[12][3], len1[12], tan1[12][3], nor1e[12][3];   // x1b-x1a, |dx1|, dx1/len1, normalized(n1c+n1d)
    // The three acos() of section C (:879-887, :907-915) only feed threshold comparisons.  Their outcome is decided on the
    // device in cosine space against critical doubles found on the host WITH THE HOST'S libm acos (the one a reference
    // build on this machine calls), see angle_cuts(): identical decisions by construction, and no FP64 acos on the device.
    double cosParHi, cosParLo;   // |acos c| < 2 deg  <=>  cosParHi <= c <= 1 ;  |pi - acos c| < 2 deg  <=>  -1 <= c <= cosParLo
    double cosWedge[12];         // acos c - angleCD[k] > 2 deg  <=>  -1 <= c < cosWedge[k]  (c <= 1)
    double aabbEdge[12][6];      // AABB of box edge k (its two corners), for the conservative pair cull of section C
};

// ---- libm-exact angle decisions in cosine space ----------------------------------------------------
// order-preserving map between doubles and integers (-0.0 and +0.0 share key 0)
inline int64_t dkey(double d) { int64_t i; std::memcpy(&i, &d, 8); return i >= 0 ? i : INT64_MIN - i; }
inline double dunkey(int64_t k) { int64_t i = k >= 0 ? k : INT64_MIN - k; double d; std::memcpy(&d, &i, 8); return d; }

// pred is false ... false true ... true over the doubles of [-1, 1] (acos is monotone; libm's rounding cannot break that
// further away from the switch than its error bound, and the WINDOW doubles on either side of it are checked one by one).
// Returns the smallest c with pred(c); 2.0 when there is none.  ok = false if the window check fails.
template <class P> double first_true(P pred, bool &ok) {
    const int64_t WINDOW = 512;
    int64_t lo = dkey(-1.0), hi = dkey(1.0);
    if (pred(-1.0)) return -1.0;
    if (!pred(1.0)) return 2.0;
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        if (pred(dunkey(mid))) hi = mid; else lo = mid;
    }
    for (int64_t k = std::max(dkey(-1.0), hi - WINDOW); k <= std::min(dkey(1.0), hi + WINDOW); ++k)
        if (pred(dunkey(k)) != (k >= hi)) ok = false;
    return dunkey(hi);
}

const double kThreshAng = 2.0 * M_PI / 180.0;   // :877

// global cuts (depend on libm only): computed once per process
bool parallel_cuts(double &hi, double &lo) {
    static bool done = false, good = true;
    static double s_hi, s_lo;
    if (!done) {
        volatile double T = kThreshAng;   // keep the calls run-time libm calls (no compile-time folding)
        s_hi = first_true([&](double c) { double angle = acos(c); return fabs(angle) < T; }, good);
        double nf = first_true([&](double c) { double angle = acos(c); return !(fabs(M_PI - angle) < T); }, good);
        s_lo = nf > 1.5 ? 1.0 : dunkey(dkey(nf) - 1);   // last c that still satisfies |pi - acos c| < T
        done = true;
    }
    hi = s_hi; lo = s_lo;
    return good;
}

V3 hcol(const double (*M)[3], int i) { return mk(M[i][0], M[i][1], M[i][2]); }

// createBox + createFaceNormals + createVertNormals on the host (libm acos, as the reference)
bool make_box(BoxData &B, const double *whd, const double *E1) {
    double S[4] = {0.5 * whd[0], 0.5 * whd[1], 0.5 * whd[2], 1.0};
    double E[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            double acc = E1[0 * 4 + r] * (c == 0 ? S[0] : 0.0);
            acc = acc + E1[1 * 4 + r] * (c == 1 ? S[1] : 0.0);
            acc = acc + E1[2 * 4 + r] * (c == 2 ? S[2] : 0.0);
            acc = acc + E1[3 * 4 + r] * (c == 3 ? S[3] : 0.0);
            E[c * 4 + r] = acc;
        }
    for (int i = 0; i < 14; ++i)
        for (int r = 0; r < 3; ++r) {
            double acc = E[0 * 4 + r] * h_unitVerts[i][0];
            acc = acc + E[1 * 4 + r] * h_unitVerts[i][1];
            acc = acc + E[2 * 4 + r] * h_unitVerts[i][2];
            acc = acc + E[3 * 4 + r] * 1.0;
            B.verts1[i][r] = acc;
        }
    double vn[14][3] = {}, angles[14] = {};
    for (int k = 0; k < 24; ++k) {
        const int *f = h_faces1[k];
        V3 xa = hcol(B.verts1, f[0]), xb = hcol(B.verts1, f[1]), xc = hcol(B.verts1, f[2]);
        V3 dba = xb - xa, dcb = xc - xb, dac = xa - xc;
        V3 nor = normalized(cross(dba, neg(dac)));
        B.faceNors1[k][0] = nor.x; B.faceNors1[k][1] = nor.y; B.faceNors1[k][2] = nor.z;
        dba = normalized(dba); dcb = normalized(dcb); dac = normalized(dac);
        double a1 = acos(dot(dba, neg(dac))), a2 = acos(dot(dcb, neg(dba))), a3 = acos(dot(dac, neg(dcb)));
        double an[3] = {a1, a2, a3};
        for (int v = 0; v < 3; ++v) {
            vn[f[v]][0] += an[v] * nor.x; vn[f[v]][1] += an[v] * nor.y; vn[f[v]][2] += an[v] * nor.z;
            angles[f[v]] += an[v];
        }
        for (int i = 0; i < 3; ++i) {   // build_AABB_F :436-452
            double a = B.verts1[f[0]][i], b = B.verts1[f[1]][i], c = B.verts1[f[2]][i];
            B.aabbF1[k][i] = std::min(c, std::min(b, a));
            B.aabbF1[k][i + 3] = std::max(c, std::max(b, a));
        }
    }
    for (int k = 0; k < 14; ++k) {
        V3 nor = normalized(divs(mk(vn[k][0], vn[k][1], vn[k][2]), angles[k]));
        B.vertNors1[k][0] = nor.x; B.vertNors1[k][1] = nor.y; B.vertNors1[k][2] = nor.z;
    }
    for (int r = 0; r < 3; ++r) {
        double mn = B.verts1[0][r], mx = B.verts1[0][r];
        for (int i = 1; i < 14; ++i) { mn = std::min(mn, B.verts1[i][r]); mx = std::max(mx, B.verts1[i][r]); }
        B.aabbB1[r] = mn; B.aabbB1[3 + r] = mx;
    }
    for (int k = 0; k < 12; ++k) {
        V3 n1c = hcol(B.faceNors1, h_edgeFaces1[k][0]), n1d = hcol(B.faceNors1, h_edgeFaces1[k][1]);
        B.edgeAngle[k] = acos(dot(n1c, n1d));
        B.angleCD[k] = acos(dot(n1c, n1d));
        V3 x1a = hcol(B.verts1, h_edgeVerts1[k][0]), x1b = hcol(B.verts1, h_edgeVerts1[k][1]);
        V3 dx1 = x1b - x1a;
        double len1 = norm(dx1);
        V3 tan1 = divs(dx1, len1);
        V3 nor1 = normalized(n1c + n1d);
        B.dx1[k][0] = dx1.x; B.dx1[k][1] = dx1.y; B.dx1[k][2] = dx1.z;
        B.len1[k] = len1;
        B.tan1[k][0] = tan1.x; B.tan1[k][1] = tan1.y; B.tan1[k][2] = tan1.z;
        B.nor1e[k][0] = nor1.x; B.nor1e[k][1] = nor1.y; B.nor1e[k][2] = nor1.z;
    }
    for (int k = 0; k < 12; ++k)
        for (int r = 0; r < 3; ++r) {
            const double a = B.verts1[h_edgeVerts1[k][0]][r], b = B.verts1[h_edgeVerts1[k][1]][r];
            B.aabbEdge[k][r] = std::min(a, b); B.aabbEdge[k][3 + r] = std::max(a, b);
        }
    bool ok = parallel_cuts(B.cosParHi, B.cosParLo);
    for (int k = 0; k < 12; ++k) {
        // :906-915 with angleCD >= 0 (an acos; the `angleCD < 0` branch cannot be taken): reject iff angleCN - angleCD > threshAng
        volatile double angleCD = B.angleCD[k], T = kThreshAng;
        B.cosWedge[k] = first_true([&](double c) { double angleCN = acos(c); return !(angleCN - angleCD > T); }, ok);
    }
    return ok;
}

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ V3 dcol(const double *M, int i) { return mk(M[3 * (size_t)i], M[3 * (size_t)i + 1], M[3 * (size_t)i + 2]); }
__device__ __forceinline__ V3 bcol(const double (*M)[3], int i) { return mk(M[i][0], M[i][1], M[i][2]); }

__device__ __forceinline__ void zero_contact(eolc_contact &c) {
    c.dist = 0;
    for (int i = 0; i < 3; ++i) {
        c.nor1[i] = c.nor2[i] = c.pos1[i] = c.pos2[i] = c.pos1_[i] = c.weights1[i] = c.weights2[i] = c.edgeDir[i] = 0.0;
        c.verts1[i] = c.verts2[i] = 0; c.edge1[i] = -1;
    }
    c.count1 = c.count2 = 0; c.tri1 = c.tri2 = -1; c.n_edge1 = 0; c.edge2 = -1; c.reserved = 0;
}
// step (E): pos1_ = pos1 - snapDepth*nor1 (:1057-1060, :1218-1221)
__device__ __forceinline__ void finish_contact(eolc_contact &c, double threshold) {
    double snap = mul(0.1, threshold);
    for (int i = 0; i < 3; ++i) c.pos1_[i] = sub(c.pos1[i], mul(snap, c.nor1[i]));
}
// CD only (Collisions.cpp:39-48): box feature ids -> constraint-table columns; Box::num_points = 8, num_edges = 12.
// remap_base < 0 disables it (CD2, and the obstacle-point sections).
__device__ __forceinline__ void remap_contact(eolc_contact &c, int remap_base) {
    if (remap_base < 0) return;
    if (c.count1 == 1 && c.count2 == 3) c.verts1[0] = remap_base + c.verts1[0];
    for (int e = 0; e < c.n_edge1; ++e) c.edge1[e] = remap_base + (8 + c.edge1[e]);
}

// ---- prepare -------------------------------------------------------------------------------------
__global__ void k_perturb(int N, const double *__restrict__ x, const double *__restrict__ r, double *__restrict__ xp, size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * (size_t)N) return;
    size_t o = blockIdx.y * stride;
    xp[o + i] = add(x[o + i], r[i]);   // verts2.block<3,1>(0,i2) += r
}

// Face normals.  createFaceNormals (:233-247) forms dba.cross(-dac) on the perturbed verts, createEdges (:193-214)
// (xb-xa).cross(xc-xa) on the unperturbed ones; both are kept as written (they differ at most in the sign of a zero).
// Box scenes evaluate them where they are needed (section A's and C's records, every face once in section Bc) instead of storing
// two normals per face first: on the 4096 x 64^2 batch that pass wrote 1.56 GB and took 0.51 ms of the 4.1 ms narrow phase.
__device__ __forceinline__ V3 face_cross_p(V3 xa, V3 xb, V3 xc) { return cross(xb - xa, neg(xa - xc)); }
__device__ __forceinline__ V3 face_normal_p(const int32_t *__restrict__ fn, const double *__restrict__ xp, int k) {
    return normalized(face_cross_p(dcol(xp, fn[3 * (size_t)k]), dcol(xp, fn[3 * (size_t)k + 1]), dcol(xp, fn[3 * (size_t)k + 2])));
}
__device__ __forceinline__ V3 face_normal_0(const int32_t *__restrict__ fn, const double *__restrict__ x, int k) {
    V3 xa = dcol(x, fn[3 * (size_t)k]), xb = dcol(x, fn[3 * (size_t)k + 1]), xc = dcol(x, fn[3 * (size_t)k + 2]);
    return normalized(cross(xb - xa, xc - xa));
}
// Stored normals of the perturbed verts: only for obstacle POINTS (section PT tests every point against every face, so a face's
// normal would otherwise be formed once per point).
__global__ void k_face_normals(int F, const int32_t *__restrict__ fn, const double *__restrict__ xp, double *__restrict__ fnp,
                               size_t xstride, size_t fstride) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    st(fnp + blockIdx.y * fstride + 3 * (size_t)k, face_normal_p(fn, xp + blockIdx.y * xstride, k));
}

// per-scene AABB of the perturbed verts (build_AABB_B :425-433); min/max are order independent.
// stage 1: grid (nb, S) -> partial[(s*nb + blk)*6]; stage 2 (same kernel, nb = 1 over the partials): one block per scene.
__global__ void __launch_bounds__(256) k_aabb(int n, int ld, const double *__restrict__ v_, double *__restrict__ out, size_t stride) {
    // items are rows of `ld` doubles: ld == 3 -> points (min == max source), ld == 6 -> partial boxes
    const double *v = v_ + blockIdx.y * stride;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int r = 0; r < 3; ++r) {
            mn[r] = fmin(mn[r], v[(size_t)ld * i + r]);
            mx[r] = fmax(mx[r], v[(size_t)ld * i + (ld == 6 ? 3 : 0) + r]);
        }
    __shared__ double s[6][8];
    for (int r = 0; r < 3; ++r)
        for (int o = 16; o > 0; o >>= 1) {
            mn[r] = fmin(mn[r], __shfl_xor_sync(0xffffffffu, mn[r], o));
            mx[r] = fmax(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
        }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) for (int r = 0; r < 3; ++r) { s[r][w] = mn[r]; s[3 + r][w] = mx[r]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double *o = out + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 6;
        for (int r = 0; r < 3; ++r) {
            double a = s[r][0], b = s[3 + r][0];
            for (int i = 1; i < 8; ++i) { a = fmin(a, s[r][i]); b = fmax(b, s[3 + r][i]); }
            o[r] = a; o[3 + r] = b;
        }
    }
}

// ---- section bookkeeping ---------------------------------------------------------------------------
// info[] holds one int per item, sections are padded to 256 items; blocksum[] one count per 256 items.
__device__ __forceinline__ void block_count_store(int cnt, int *blocksum, size_t blk) {
    // 256 threads: warp reduce + smem
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 8; ++i) t += s[i]; blocksum[blk] = t; }
}
// exclusive prefix of cnt over the 256 threads of the block (thread order), warp-shuffle scan
__device__ __forceinline__ int block_excl_prefix(int cnt) {
    __shared__ int s[8];
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += t; }
    if (l == 31) s[w] = inc;
    __syncthreads();
    int base = 0;
    for (int i = 0; i < w; ++i) base += s[i];
    return base + inc - cnt;
}

// single-block exclusive scan of blocksum[0..n) -> blockoff[0..n], blockoff[n] = total
__global__ void __launch_bounds__(1024) k_scan(size_t n, const int *__restrict__ in, int *__restrict__ out) {
    // one block, 8 consecutive entries per thread and round: 8192 entries per round (a batched ensemble has ~10^5 block counts)
    constexpr int PER = 8;
    __shared__ int s[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t base = 0; base < n; base += 1024 * PER) {
        const size_t i0 = base + (size_t)threadIdx.x * PER;
        int v[PER], tot = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { v[q] = i0 + q < n ? in[i0 + q] : 0; tot += v[q]; }
        int inc = tot;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += t; }
        if (l == 31) s[w] = inc;
        __syncthreads();
        if (w == 0) {
            int t = s[l];
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, t, o); if (l >= o) t += u; }
            s[l] = t;
        }
        __syncthreads();
        const int wbase = w > 0 ? s[w - 1] : 0;
        const int c = carry;
        int run = c + wbase + inc - tot;
#pragma unroll
        for (int q = 0; q < PER; ++q) { if (i0 + q < n) out[i0 + q] = run; run += v[q]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// The same scan on many blocks for long inputs (a batched ensemble has ~3 x 10^5 block counts: 0.2 ms on one block): chunks of 8192
// entries are scanned locally, their totals by the single-block kernel above, and the chunk offsets are added.
constexpr int SCAN_CHUNK = 1024 * 8;
__global__ void __launch_bounds__(1024) k_scan_local(size_t n, const int *__restrict__ in, int *__restrict__ out, int *__restrict__ chunk_total) {
    constexpr int PER = 8;
    __shared__ int s[32];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t i0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * PER;
    int v[PER], tot = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { v[q] = i0 + q < n ? in[i0 + q] : 0; tot += v[q]; }
    int inc = tot;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += t; }
    if (l == 31) s[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = s[l];
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, t, o); if (l >= o) t += u; }
        s[l] = t;
    }
    __syncthreads();
    int run = (w > 0 ? s[w - 1] : 0) + inc - tot;
#pragma unroll
    for (int q = 0; q < PER; ++q) { if (i0 + q < n) out[i0 + q] = run; run += v[q]; }
    if (threadIdx.x == 1023) chunk_total[blockIdx.x] = run;
}
__global__ void __launch_bounds__(1024) k_scan_add(size_t n, int *__restrict__ out, const int *__restrict__ chunk_off, int n_chunks) {
    const size_t i0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * 8;
    const int add = chunk_off[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 8; ++q) if (i0 + q < n) out[i0 + q] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = chunk_off[n_chunks];
}

// Division-free screen of barycentric()'s verdict.  nb, ng, denom are the SAME floats barycentric() divides (beta = nb / denom,
// gamma = ng / denom; same operations in the same order), and IEEE division is monotone with a relative error of 2^-53: a quotient
// beyond [-1e-6, 1 + 1e-6] stays beyond [0, 1] when rounded, and u = (1 - beta) - gamma, w = (1 - u) - beta follow beta and gamma
// to a few 1e-16 while both are O(1).  So `true` means the reference's `u < 0 || 1 < u || v < 0 || ...` rejects for certain,
// whatever the triangle's shape; zero or NaN denominators are left to the reference's own expressions.
__device__ __forceinline__ bool barycentric_rejects(double nb, double ng, double denom) {
    const double e = 1e-6;
    if (denom > 0.0) { const double lo = -e * denom, hi = denom + e * denom; return nb < lo || nb > hi || ng < lo || ng > hi || nb + ng > hi; }
    if (denom < 0.0) { const double lo = -e * denom, hi = denom + e * denom; return nb > lo || nb < hi || ng > lo || ng < hi || nb + ng < hi; }
    return false;
}

// What pass 1 of section A reads of one box, as a kernel PARAMETER (constant bank: the 24 plane tests take their operands from it
// without a load; the pass used to read the box from global memory).  faceA[j] = verts1[faces1[j][0]].
struct ABox { double aabbB1[6]; double faceA[24][3]; double faceNors1[24][3]; double verts1[14][3]; };
inline void make_a_box(ABox &A, const BoxData &B) {
    for (int r = 0; r < 6; ++r) A.aabbB1[r] = B.aabbB1[r];
    for (int j = 0; j < 24; ++j) for (int r = 0; r < 3; ++r) { A.faceA[j][r] = B.verts1[h_faces1[j][0]][r]; A.faceNors1[j][r] = B.faceNors1[j][r]; }
    for (int v = 0; v < 14; ++v) for (int r = 0; r < 3; ++r) A.verts1[v][r] = B.verts1[v][r];
}
// ---- section A: cloth vertex vs box (boxTriCollision.cpp:675-764) -----------------------------------
// returns the winning j1 (or -1); the record of the winner is formed in pass 2 (vertex_box_record)
__device__ __forceinline__ int test_vertex_box(V3 x2, const ABox &B, double threshold) {
    if (!check_aabb_point(x2, B.aabbB1)) return -1;
    // :689-697 — inside all 24 half-spaces; the projections are kept: the second loop (:699-) forms the same expression again
    double pj[24];
#pragma unroll
    for (int j1 = 0; j1 < 24; ++j1) {
        pj[j1] = dot(bcol(B.faceNors1, j1), x2 - bcol(B.faceA, j1));
        if (pj[j1] > 0.0) return -1;
    }
    int best = -1;
    double bestd = 0.0;
    const double lim = mul(5.0, threshold);
    // A conservative screen the reference does not have and that cannot change a result: dist = |x2 - (x2 - proj nor1)| equals |proj|
    // up to a few roundings of the coordinates (nor1 is a unit vector), so a triangle whose plane is farther than lim plus a margin a
    // million times those roundings fails `dist > lim` (:712) for certain; only the near planes (the face the vertex rests on) run the
    // reference's expressions, which decide exactly as before.  Removes ~20 of the 24 sqrt / barycentric evaluations per vertex.
    const double guard = lim + (1e-6 * lim + 1e-9 + 1e-10 * (fabs(x2.x) + fabs(x2.y) + fabs(x2.z)));
#pragma unroll 1
    for (int j1 = 0; j1 < 24; ++j1) {
        const double proj = pj[j1];
        if (-proj > guard) continue;
        V3 x1a = bcol(B.verts1, c_faces1[j1][0]), x1b = bcol(B.verts1, c_faces1[j1][1]), x1c = bcol(B.verts1, c_faces1[j1][2]);
        V3 nor1 = bcol(B.faceNors1, j1);
        if (proj > 0.0) continue;
        V3 x1 = x2 - scale(proj, nor1);
        {   // a vertex resting on a box face is near the planes of that face's four triangles and inside one of them: the other three
            // are turned away here without the square root and the two divisions (barycentric_rejects: same floats, same verdict)
            const V3 v0 = x1b - x1a, v1 = x1c - x1a, v2 = x1 - x1a;
            const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
            if (barycentric_rejects(sub(mul(d11, d20), mul(d01, d21)), sub(mul(d00, d21), mul(d01, d20)), sub(mul(d00, d11), mul(d01, d01)))) continue;
        }
        double dist = norm(x2 - x1);
        if (dist > lim) continue;
        double u, v;
        barycentric(u, v, x1a, x1b, x1c, x1);
        double w = sub(sub(1.0, u), v);
        if (u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w) continue;
        if (best < 0 || dist < bestd) { best = j1; bestd = dist; }
    }
    return best;
}

// The record of cloth vertex i2 against its winning box triangle j1 (what the loop above leaves in *rec for the winner): the same
// expressions on the same inputs, so the same bits, without the other 23 triangles.
__device__ __forceinline__ void vertex_box_record(int i2, int j1, int F, V3 x2, const int32_t *__restrict__ fn, const double *__restrict__ xs, const BoxData &B, eolc_contact *rec) {
    V3 x1a = bcol(B.verts1, c_faces1[j1][0]), x1b = bcol(B.verts1, c_faces1[j1][1]), x1c = bcol(B.verts1, c_faces1[j1][2]);
    V3 nor1 = bcol(B.faceNors1, j1);
    double proj = dot(nor1, x2 - x1a);
    V3 x1 = x2 - scale(proj, nor1);
    double dist = norm(x2 - x1);
    double u, v;
    barycentric(u, v, x1a, x1b, x1c, x1);
    double w = sub(sub(1.0, u), v);
    V3 nor2 = i2 < F ? face_normal_p(fn, xs, i2) : mk(0, 0, 0);
    if (dot(nor2, nor1) < 0.0) nor2 = neg(nor2);
    zero_contact(*rec);
    rec->dist = dist;
    st(rec->nor1, nor1); st(rec->nor2, nor2); st(rec->pos1, x1); st(rec->pos2, x2);
    rec->count1 = 3; rec->count2 = 1;
    rec->verts1[0] = c_faces1[j1][0]; rec->verts1[1] = c_faces1[j1][1]; rec->verts1[2] = c_faces1[j1][2];
    rec->verts2[0] = i2; rec->verts2[1] = -1; rec->verts2[2] = -1;
    rec->weights1[0] = u; rec->weights1[1] = v; rec->weights1[2] = w;
    rec->weights2[0] = 1.0; rec->weights2[1] = 0.0; rec->weights2[2] = 0.0;
    rec->tri1 = j1; rec->tri2 = -1;
}

// grid: (ceil(N/256), S), one launch per box
__global__ void __launch_bounds__(256) k_A_count(int N, int b, const __grid_constant__ ABox B, const double *__restrict__ xp, double threshold,
                                                 int *__restrict__ info, size_t xstride, size_t scene_items, size_t box_items, size_t secA_off) {
    const int s = blockIdx.y;
    const int i2 = blockIdx.x * 256 + threadIdx.x;
    int j1 = -1;
    if (i2 < N) j1 = test_vertex_box(dcol(xp + s * xstride, i2), B, threshold);
    info[s * scene_items + secA_off + b * box_items + i2] = j1;
    // no block-wide count here: the warps of a block finish far apart (a vertex outside the box returns after one AABB test, one on a
    // face runs the barycentric path), and a barrier kept the early ones — and the CTA's slot — waiting (46 % of the kernel's stall
    // samples, ncu r02k).  k_A_sum counts the hits of every 256-item block afterwards, one warp per block.
}
__global__ void __launch_bounds__(256) k_A_sum(int nbx, long long nblk, const int *__restrict__ info, int *__restrict__ blocksum,
                                               size_t scene_items, size_t box_items, size_t secA_off, int nB) {
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= nblk) return;
    const int bx = (int)(w % nbx);
    const long long sb = w / nbx;
    const size_t item0 = (size_t)(sb / nB) * scene_items + secA_off + (size_t)(sb % nB) * box_items;
    const int4 *p = reinterpret_cast<const int4 *>(info + item0 + (size_t)bx * 256 + 8 * (threadIdx.x & 31));
    const int4 m0 = p[0], m1 = p[1];
    int t = (m0.x >= 0) + (m0.y >= 0) + (m0.z >= 0) + (m0.w >= 0) + (m1.x >= 0) + (m1.y >= 0) + (m1.z >= 0) + (m1.w >= 0);
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) blocksum[item0 / 256 + bx] = t;
}
// The hits of 32 consecutive items occupy consecutive output slots (prefix order), so a WARP's records form one contiguous range: they
// are staged in the warp's slice of shared memory and leave with coalesced 8-byte stores (a record is 33 doubles; writing it from one
// thread puts 32 lanes on 32 different 264-byte-strided lines per store).  Every warp works on its own — work item = (scene, box,
// 256-item block, warp of it), no block barrier: the slot of a warp's first hit is the block's offset plus the hits of the block's
// earlier items, which the warp counts itself from the block's 256 info words (one coalesced kilobyte).  (The first version staged per
// CTA with three barriers per item block and ran at a third of the warp slots: 0.55 ms on the 4096 x 64^2 batch.)
__global__ void __launch_bounds__(256) k_A_write(int N, int F, int nB, int S, const double *__restrict__ xp, const int32_t *__restrict__ fn,
                                                 const BoxData *__restrict__ boxes, double threshold, const int *__restrict__ info,
                                                 const int *__restrict__ blockoff, eolc_contact *__restrict__ out, size_t xstride,
                                                 size_t scene_items, size_t box_items, size_t secA_off, int out_cap) {
    extern __shared__ __align__(16) unsigned char stage_raw[];
    static_assert(sizeof(eolc_contact) % 8 == 0, "records are copied as 8-byte words");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    eolc_contact *stage = reinterpret_cast<eolc_contact *>(stage_raw) + 32 * wib;
    const int nbx = (N + 255) / 256;
    const long long nwork = (long long)nbx * nB * S * 8;
    // 32 of a warp's work items at a time: lane l reads the two block offsets of its l-th next item, and the warp then works through the items whose
    // block has hits (most have none).  One item after the other, each empty item cost a full memory round trip of the whole warp:
    // the kernel was a chain of ~150 dependent latencies per warp (0.57 ms on the 4096 x 64^2 batch at 25 % of the warp slots).
    // A big batch hands every warp runs of 32 consecutive items (neighbouring items share their block's info words and offsets);
    // with little work per warp the items go round the warps one by one, so that no warp is left with a run of 32 busy ones.
    const long long W = (long long)gridDim.x * 8, w0 = (long long)blockIdx.x * 8 + wib;
    const bool runs = nwork >= 64 * W;
    const long long lstride = runs ? 1 : W;
    for (long long wk0 = runs ? 32 * w0 : w0; wk0 < nwork; wk0 += 32 * W) {
      unsigned todo;
      {
        const long long wkl = wk0 + lane * lstride;
        bool has = false;
        if (wkl < nwork) {
            const long long vbl = wkl >> 3;
            const int bxl = (int)(vbl % nbx), sbl = (int)(vbl / nbx);
            const size_t blkl = ((size_t)(sbl / nB) * scene_items + secA_off + (size_t)(sbl % nB) * box_items) / 256 + bxl;
            has = blockoff[blkl + 1] != blockoff[blkl];
        }
        todo = __ballot_sync(0xffffffffu, has);
      }
      while (todo) {
        const long long wk = wk0 + (__ffs(todo) - 1) * lstride;
        todo &= todo - 1;
        const int sub = (int)(wk & 7);                 // which 32 items of the 256-item block
        const long long vb = wk >> 3;
        const int bx = (int)(vb % nbx);
        const int sb = (int)(vb / nbx), s = sb / nB, b = sb % nB;
        const size_t item0 = s * scene_items + secA_off + b * box_items;
        const size_t blk = item0 / 256 + bx;
        const int base = blockoff[blk];
        // hits of the block's earlier items: lane l counts the items 8 l .. 8 l + 7 of the block, a warp scan gives every prefix
        const int4 *ip = reinterpret_cast<const int4 *>(info + item0 + (size_t)bx * 256 + 8 * lane);
        const int4 m0 = ip[0], m1 = ip[1];
        const int c8 = (m0.x >= 0) + (m0.y >= 0) + (m0.z >= 0) + (m0.w >= 0) + (m1.x >= 0) + (m1.y >= 0) + (m1.z >= 0) + (m1.w >= 0);
        int inc8 = c8;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc8, o); if (lane >= o) inc8 += t; }
        // items before this warp's 32: the first 4 sub lanes' worth of 8-item groups
        const int before = sub ? __shfl_sync(0xffffffffu, inc8, 4 * sub - 1) : 0;
        const int i2 = bx * 256 + 32 * sub + lane;
        const int j1 = info[item0 + i2];
        const unsigned hits = __ballot_sync(0xffffffffu, j1 >= 0);
        const int cnt = __popc(hits);
        if (cnt == 0) continue;                        // warp-uniform
        const int pre = __popc(hits & ((1u << lane) - 1u));
        if (j1 >= 0) {                                 // the record is built in place in the warp's stage
            vertex_box_record(i2, j1, F, dcol(xp + s * xstride, i2), fn, xp + s * xstride, boxes[b], &stage[pre]);
            finish_contact(stage[pre], threshold);
        }
        __syncwarp();
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(stage);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(out + base + before);
        const int nw = max(0, min(cnt, out_cap - (base + before))) * (int)(sizeof(eolc_contact) / 8);     // out_cap: see k_PT_write
        for (int w = lane; w < nw; w += 32) dst[w] = src[w];
        __syncwarp();                                  // the warp's stage is reused by its next work item
      }
    }
}

// ---- sections Bc / PT: point (box corner or obstacle point) vs all cloth triangles -------------------
// (:771-845 and :1142-1213): strict-min over j2 ascending == lexicographic min of (dist, j2)
struct Cand { double dist; int j2; };
__device__ __forceinline__ bool better(double d, int j, const Cand &c) { return c.j2 < 0 || d < c.dist || (d == c.dist && j < c.j2); }

__device__ __forceinline__ bool test_point_tri(V3 x1, V3 nor1, int j2, const int32_t *__restrict__ fn, const double *__restrict__ xp,
                                               const double *__restrict__ fnp, double lim, double &dist, V3 &nor2, V3 &x2,
                                               double &u, double &v, double &w) {
    V3 x2a = dcol(xp, fn[3 * (size_t)j2]), x2b = dcol(xp, fn[3 * (size_t)j2 + 1]), x2c = dcol(xp, fn[3 * (size_t)j2 + 2]);
    nor2 = fnp ? dcol(fnp, j2) : normalized(face_cross_p(x2a, x2b, x2c));   // NULL: box scenes, no stored normals
    if (dot(nor1, nor2) < 0.0) nor2 = neg(nor2);
    double proj = dot(x1 - x2a, nor2);
    if (proj < 0.0) return false;
    x2 = x1 - scale(proj, nor2);
    dist = norm(x2 - x1);
    if (dist > lim) return false;
    barycentric(u, v, x2a, x2b, x2c, x1);   // the unprojected point is passed (:808)
    w = sub(sub(1.0, u), v);
    if (u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w) return false;
    return true;
}

// which point: box mode (boxes != NULL): point id = blockIdx.y % 8 of box (blockIdx.y / 8) % nB ; point mode: pxyz/pnorms
// grid: (nchunk, S * npts) ; partial[(s*npts + pt) * nchunk + chunk]
__global__ void __launch_bounds__(256) k_PT_partial(int F, int npts_per_scene, int nB, const BoxData *__restrict__ boxes,
                                                    const double *__restrict__ pxyz, const double *__restrict__ pnorms,
                                                    const int32_t *__restrict__ fn, const double *__restrict__ xp,
                                                    const double *__restrict__ fnp, const double *__restrict__ aabbB2,
                                                    double threshold, Cand *__restrict__ partial, size_t xstride, size_t fstride) {
    int s = blockIdx.y / npts_per_scene, pt = blockIdx.y % npts_per_scene;
    V3 x1, nor1;
    if (boxes) { const BoxData &B = boxes[pt / 8]; x1 = bcol(B.verts1, pt % 8); nor1 = bcol(B.vertNors1, pt % 8); }
    else { x1 = dcol(pxyz, pt); nor1 = dcol(pnorms, pt); }
    Cand best; best.dist = 0.0; best.j2 = -1;
    if (check_aabb_point(x1, aabbB2 + 6 * s)) {
        const double lim = mul(5.0, threshold);
        const double *xs = xp + s * xstride, *fs = fnp ? fnp + s * fstride : nullptr;
        for (int j2 = blockIdx.x * 256 + threadIdx.x; j2 < F; j2 += gridDim.x * 256) {
            double dist, u, v, w; V3 nor2, x2;
            if (test_point_tri(x1, nor1, j2, fn, xs, fs, lim, dist, nor2, x2, u, v, w) && better(dist, j2, best)) { best.dist = dist; best.j2 = j2; }
        }
    }
    // warp-cooperative lexicographic (dist, j2) min
    for (int o = 16; o > 0; o >>= 1) {
        double d = __shfl_xor_sync(0xffffffffu, best.dist, o);
        int j = __shfl_xor_sync(0xffffffffu, best.j2, o);
        if (j >= 0 && better(d, j, best)) { best.dist = d; best.j2 = j; }
    }
    __shared__ Cand sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        Cand b0 = sm[0];
        for (int i = 1; i < 8; ++i) if (sm[i].j2 >= 0 && better(sm[i].dist, sm[i].j2, b0)) b0 = sm[i];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = b0;
    }
}
#ifndef EOLC_PT8_WAVES
#define EOLC_PT8_WAVES 4LL     // CTAs of k_PT_partial8 per SM a run aims at (chunks of a scene's faces are added until there are that many)
#endif
#ifndef EOLC_PT8_CTAS
#define EOLC_PT8_CTAS 3      // resident CTAs per SM the register allocation of k_PT_partial8 aims at (80 registers, no spills: 352 us on the batch; 4 CTAs with 64 registers and 64 bytes of spills: 378 us)
#endif
// Box mode, all 8 corners of a box against a chunk of the cloth's faces in ONE pass: a face's indices, its three vertices and its
// normal are loaded once and tested against the eight corners (the per-corner kernel above read every face eight times: 1.2 GB of DRAM
// and eight times the gather instructions on the 4096-scene batch).  Same expressions on the same inputs as test_point_tri, so the
// same bits; partial[] has the layout of the per-corner kernel.  grid: (nchunk, S * nB)
// The reference's test of one point against one triangle (test_point_tri above, line by line) for the few pairs that pass the screens
// of k_PT_partial8; out of line so that the screening loop stays small.
__device__ __noinline__ bool point_tri_exact(V3 x1c, V3 nor1c, V3 x2a, V3 x2b, V3 x2c, double lim, double &dist) {
    V3 nor2 = normalized(face_cross_p(x2a, x2b, x2c));
    if (dot(nor1c, nor2) < 0.0) nor2 = neg(nor2);
    const double proj = dot(x1c - x2a, nor2);
    if (proj < 0.0) return false;
    const V3 x2 = x1c - scale(proj, nor2);
    dist = norm(x2 - x1c);
    if (dist > lim) return false;
    double u, v;
    barycentric(u, v, x2a, x2b, x2c, x1c);
    const double w = sub(sub(1.0, u), v);
    return !(u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w);
}
__global__ void __launch_bounds__(256, EOLC_PT8_CTAS) k_PT_partial8(int F, int nB, const BoxData *__restrict__ boxes, const int32_t *__restrict__ fn,
                                                     const double *__restrict__ xp,
                                                     const double *__restrict__ aabbB2, double threshold, Cand *__restrict__ partial,
                                                     size_t xstride) {
    const int s = blockIdx.y / nB, b = blockIdx.y % nB;
    const BoxData &B = boxes[b];
    // the corners and their normals live in shared memory (uniform reads): 48 doubles in registers would leave one CTA per SM
    __shared__ double cx[8][3], cn[8][3], cl1[8];
    __shared__ int lc[8], nlive_s;                       // the corners inside the cloth's AABB (:774), ascending
    if (threadIdx.x < 24) { cx[threadIdx.x / 3][threadIdx.x % 3] = B.verts1[threadIdx.x / 3][threadIdx.x % 3]; cn[threadIdx.x / 3][threadIdx.x % 3] = B.vertNors1[threadIdx.x / 3][threadIdx.x % 3]; }
    if (threadIdx.x == 0) {
        int n = 0;
        for (int c = 0; c < 8; ++c) {
            const V3 p = bcol(B.verts1, c);
            cl1[c] = fabs(p.x) + fabs(p.y) + fabs(p.z);
            if (check_aabb_point(p, aabbB2 + 6 * s)) lc[n++] = c;
        }
        nlive_s = n;
    }
    __syncthreads();
    const int nlive = nlive_s;
    // a thread's best candidate per corner: shared memory, [corner][thread] (rarely written: only a hit closer than the best so far)
    __shared__ double bd[8][256];
    __shared__ int bj[8][256];
    for (int q = 0; q < nlive; ++q) { bd[lc[q]][threadIdx.x] = 0.0; bj[lc[q]][threadIdx.x] = -1; }
    if (nlive) {
        const double lim = mul(5.0, threshold);
        const double *xs = xp + s * xstride;
        // Two conservative screens the reference does not have and that cannot change a result; only pairs that pass both run the
        // reference's expressions (point_tri_exact), which decide exactly as before.
        // (1) Plane distance.  With m = dba x (-dac) (the vector the reference normalises into faceNors2) a record needs 0 <= proj and
        //     |x2 - x1| <= lim, where proj = (x1 - x2a) . m / |m| and |x2 - x1| = |proj| up to a few roundings of the coordinates.  A
        //     corner whose distance to the face's plane, formed WITHOUT the normalisation, exceeds lim by a margin a million times
        //     those roundings fails one of the two tests for certain, whichever way nor2 is flipped.
        // (2) barycentric_rejects: the cloth lying flat on the box's top is in the plane of four corners with all its faces.
        // The unit normal (a square root and three divisions) is formed for the faces that come that close to a corner only — the
        // kernel used to read it from a 24 B-per-face array that a separate pass over all faces had written (0.51 ms on the batch).
        const double guard0 = lim + 1e-3 * lim + 1e-9;
        // the indices of the thread's next face are fetched while it works on this one: with one CTA per scene a thread walks ~30
        // faces, and index -> vertices -> arithmetic was two dependent round trips per face
        int j2 = blockIdx.x * 256 + threadIdx.x;
        int ia = 0, ib = 0, ic = 0;
        if (j2 < F) { ia = fn[3 * (size_t)j2]; ib = fn[3 * (size_t)j2 + 1]; ic = fn[3 * (size_t)j2 + 2]; }
        for (; j2 < F; j2 += gridDim.x * 256) {
            const int ja = ia, jb = ib, jc = ic;
            const V3 x2a = dcol(xs, ja);
            V3 v0, v1;                                             // barycentric()'s v0, v1; v1 == -(x2a - x2c) exactly
            {   // the other two vertices are not kept: the few pairs that reach the reference's expressions load the face again
                const V3 x2b = dcol(xs, jb), x2c = dcol(xs, jc);
                v0 = x2b - x2a; v1 = x2c - x2a;
            }
            {
                const int jn = j2 + gridDim.x * 256;
                if (jn < F) { ia = fn[3 * (size_t)jn]; ib = fn[3 * (size_t)jn + 1]; ic = fn[3 * (size_t)jn + 2]; }
            }
            const V3 m = cross(v0, v1);                            // == face_cross_p up to the sign of a zero
            const double mm = m.x * m.x + m.y * m.y + m.z * m.z;
            const double t = x2a.x * m.x + x2a.y * m.y + x2a.z * m.z;
            const double gbase = guard0 + 1e-9 * (fabs(x2a.x) + fabs(x2a.y) + fabs(x2a.z));
            const bool screen = mm > 1e-200 && mm < 1e200;         // degenerate faces take the reference's path unscreened
            bool have_lat = false;
            double d00 = 0.0, d01 = 0.0, d11 = 0.0, denom = 0.0;
            unsigned pass = 0;                                     // corners (positions in lc[]) that the screens let through
#pragma unroll 1
            for (int q = 0; q < nlive; ++q) {
                if (screen) {
                    const int c = lc[q];
                    const V3 x1c = mk(cx[c][0], cx[c][1], cx[c][2]);
                    const double pm = (x1c.x * m.x + x1c.y * m.y + x1c.z * m.z) - t;
                    const double g = gbase + 1e-9 * cl1[c];
                    if (pm * pm > g * g * mm) continue;
                    if (!have_lat) { d00 = dot(v0, v0); d01 = dot(v0, v1); d11 = dot(v1, v1); denom = sub(mul(d00, d11), mul(d01, d01)); have_lat = true; }
                    const V3 v2 = x1c - x2a;
                    const double d20 = dot(v2, v0), d21 = dot(v2, v1);
                    if (barycentric_rejects(sub(mul(d11, d20), mul(d01, d21)), sub(mul(d00, d21), mul(d01, d20)), denom)) continue;
                }
                pass |= 1u << q;
            }
            while (pass) {                                         // rare: a handful of faces per corner
                const int q = __ffs(pass) - 1;
                pass &= pass - 1;
                const int c = lc[q];
                double dist;
                if (!point_tri_exact(mk(cx[c][0], cx[c][1], cx[c][2]), mk(cn[c][0], cn[c][1], cn[c][2]), dcol(xs, ja), dcol(xs, jb), dcol(xs, jc), lim, dist)) continue;
                Cand cur; cur.dist = bd[c][threadIdx.x]; cur.j2 = bj[c][threadIdx.x];
                if (better(dist, j2, cur)) { bd[c][threadIdx.x] = dist; bj[c][threadIdx.x] = j2; }
            }
        }
    }
    __shared__ Cand sm[8][8];
    if (threadIdx.x < 64) { sm[threadIdx.x >> 3][threadIdx.x & 7].dist = 0.0; sm[threadIdx.x >> 3][threadIdx.x & 7].j2 = -1; }   // corners that are not live: no candidate
    __syncthreads();
    for (int q = 0; q < nlive; ++q) {
        const int c = lc[q];
        Cand bc; bc.dist = bd[c][threadIdx.x]; bc.j2 = bj[c][threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            const double d = __shfl_xor_sync(0xffffffffu, bc.dist, o);
            const int j = __shfl_xor_sync(0xffffffffu, bc.j2, o);
            if (j >= 0 && better(d, j, bc)) { bc.dist = d; bc.j2 = j; }
        }
        if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = bc;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int c = threadIdx.x;
        Cand b0 = sm[c][0];
        for (int i = 1; i < 8; ++i) if (sm[c][i].j2 >= 0 && better(sm[c][i].dist, sm[c][i].j2, b0)) b0 = sm[c][i];
        partial[((size_t)blockIdx.y * 8 + c) * gridDim.x + blockIdx.x] = b0;
    }
}
// one warp per (scene, point): reduce the partials, store winner j2 in info[], count in blocksum (section <= 256 items)
// grid: S * nsec blocks of 256 threads; section sec of scene s holds pts_per_sec points (8 per box, or P)
__global__ void __launch_bounds__(256) k_PT_final(int nchunk, int pts_per_sec, int nsec, const Cand *__restrict__ partial,
                                                  int *__restrict__ info, int *__restrict__ blocksum, size_t scene_items,
                                                  size_t sec0_off, size_t sec_stride) {
    int s = blockIdx.x / nsec, sec = blockIdx.x % nsec;
    size_t item0 = s * scene_items + sec0_off + sec * sec_stride;
    int total = 0;
    for (int base = 0; base < pts_per_sec; base += 256) {
        int pt = base + threadIdx.x;
        int j2 = -1;
        if (pt < pts_per_sec) {
            const Cand *p = partial + ((size_t)(s * nsec + sec) * pts_per_sec + pt) * nchunk;
            Cand best; best.dist = 0.0; best.j2 = -1;
            for (int c = 0; c < nchunk; ++c) if (p[c].j2 >= 0 && better(p[c].dist, p[c].j2, best)) best = p[c];
            j2 = best.j2;
        }
        info[item0 + base + threadIdx.x] = j2;
        int cnt = j2 >= 0 ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        __shared__ int sm[8];
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 8; ++i) t += sm[i]; blocksum[item0 / 256 + base / 256] = t; total += t; }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_PT_write(int F, int pts_per_sec, int nsec, const BoxData *__restrict__ boxes,
                                                  const double *__restrict__ pxyz, const double *__restrict__ pnorms,
                                                  const int32_t *__restrict__ fn, const double *__restrict__ xp,
                                                  const double *__restrict__ fnp, double threshold, const int *__restrict__ info,
                                                  const int *__restrict__ blockoff, eolc_contact *__restrict__ out, size_t xstride,
                                                  size_t fstride, size_t scene_items, size_t sec0_off, size_t sec_stride,
                                                  int remap, int nP, int out_cap) {
    int s = blockIdx.y / nsec, sec = blockIdx.y % nsec;
    size_t item0 = s * scene_items + sec0_off + sec * sec_stride;
    int pt = blockIdx.x * 256 + threadIdx.x;
    int j2 = info[item0 + pt];
    int pre = block_excl_prefix(j2 >= 0 ? 1 : 0);
    if (j2 < 0) return;
    V3 x1, nor1;
    if (boxes) { const BoxData &B = boxes[sec]; x1 = bcol(B.verts1, pt); nor1 = bcol(B.vertNors1, pt); }
    else { x1 = dcol(pxyz, pt); nor1 = dcol(pnorms, pt); }
    double dist, u, v, w; V3 nor2, x2;
    test_point_tri(x1, nor1, j2, fn, xp + s * xstride, fnp ? fnp + s * fstride : nullptr, mul(5.0, threshold), dist, nor2, x2, u, v, w);
    eolc_contact rec;
    zero_contact(rec);
    rec.dist = dist;
    st(rec.nor1, nor1); st(rec.nor2, nor2); st(rec.pos1, x1); st(rec.pos2, x2);
    rec.count1 = 1; rec.count2 = 3;
    rec.verts1[0] = pt; rec.verts1[1] = -1; rec.verts1[2] = -1;
    rec.verts2[0] = fn[3 * (size_t)j2]; rec.verts2[1] = fn[3 * (size_t)j2 + 1]; rec.verts2[2] = fn[3 * (size_t)j2 + 2];
    rec.weights1[0] = 1.0; rec.weights1[1] = 0.0; rec.weights1[2] = 0.0;
    rec.weights2[0] = u; rec.weights2[1] = v; rec.weights2[2] = w;
    if (boxes) { rec.edge1[0] = c_vertEdges1[pt][0]; rec.edge1[1] = c_vertEdges1[pt][1]; rec.edge1[2] = c_vertEdges1[pt][2]; rec.n_edge1 = 3; }
    rec.tri1 = -1; rec.tri2 = j2;
    finish_contact(rec, threshold);
    if (boxes && remap) remap_contact(rec, nP + sec * 8 + sec * 12);
    const int slot = blockoff[item0 / 256 + blockIdx.x] + pre;
    if (slot < out_cap) out[slot] = rec;          // out_cap: a speculative pass 2 runs before the host knows the total (eolc_cd_run)
}

// ---- section PE: cloth vertex vs obstacle points (pointTriCollision :1099-1139, EOL == true) ----------
__device__ __forceinline__ int test_vertex_points(V3 x2, int P, const double *__restrict__ pxyz, double threshold, double &bestd) {
    int best = -1;
    bestd = 0.0;
    for (int i1 = 0; i1 < P; ++i1) {
        double dist = norm(x2 - dcol(pxyz, i1));
        if (dist < threshold && (best < 0 || dist < bestd)) { best = i1; bestd = dist; }
    }
    return best;
}
__global__ void __launch_bounds__(256) k_PE_count(int N, int P, const double *__restrict__ pxyz, const double *__restrict__ xp,
                                                  double threshold, int *__restrict__ info, int *__restrict__ blocksum,
                                                  size_t xstride, size_t scene_items, size_t sec_off) {
    int s = blockIdx.y;
    int i2 = blockIdx.x * 256 + threadIdx.x;
    int i1 = -1;
    double d;
    if (i2 < N) i1 = test_vertex_points(dcol(xp + s * xstride, i2), P, pxyz, threshold, d);
    size_t item0 = s * scene_items + sec_off;
    info[item0 + i2] = i1;
    block_count_store(i1 >= 0 ? 1 : 0, blocksum, item0 / 256 + blockIdx.x);
}
__global__ void __launch_bounds__(256) k_PE_write(int N, int P, const double *__restrict__ pxyz, const double *__restrict__ pnorms,
                                                  const double *__restrict__ xp, double threshold, const int *__restrict__ info,
                                                  const int *__restrict__ blockoff, eolc_contact *__restrict__ out, size_t xstride,
                                                  size_t scene_items, size_t sec_off, int out_cap) {
    int s = blockIdx.y;
    size_t item0 = s * scene_items + sec_off;
    int i2 = blockIdx.x * 256 + threadIdx.x;
    int i1 = info[item0 + i2];
    int pre = block_excl_prefix(i1 >= 0 ? 1 : 0);
    if (i1 < 0) return;
    V3 x2 = dcol(xp + s * xstride, i2), x1 = dcol(pxyz, i1), nor1 = dcol(pnorms, i1);
    eolc_contact rec;
    zero_contact(rec);
    rec.dist = norm(x2 - x1);
    st(rec.nor1, nor1); st(rec.nor2, nor1); st(rec.pos1, x1); st(rec.pos2, x2);
    rec.count1 = 3; rec.count2 = 1;
    rec.verts1[0] = i1; rec.verts1[1] = -1; rec.verts1[2] = -1;
    rec.verts2[0] = i2; rec.verts2[1] = -1; rec.verts2[2] = -1;
    rec.weights1[0] = 1.0; rec.weights2[0] = 1.0;
    rec.tri1 = -1; rec.tri2 = -1;
    finish_contact(rec, threshold);
    const int slot = blockoff[item0 / 256 + blockIdx.x] + pre;
    if (slot < out_cap) out[slot] = rec;          // out_cap: a speculative pass 2 runs before the host knows the total (eolc_cd_run)
}

// ---- section C: cloth edge vs box edge (boxTriCollision.cpp:849-1017) ----------------------------------
struct EdgeRec { int32_t v[4]; int32_t f[2]; };

// The rejections of one (cloth edge, box edge) pair that need nothing but the two AABBs: the reference's soft-edge and face-AABB
// tests (:858-871), then a conservative cull that is not in the reference and cannot change a result: a record needs u1 within
// thr/len1 of [0, 1] and u2 within thr/len2 of [0, 1] (:933, :981), i.e. x1 within thr of the box edge's AABB and x2 within thr of
// the cloth edge's, and |x2 - x1| <= 2 thr (:988-999): the two AABBs are then at most 4 thr apart in every coordinate.  Pairs
// further apart than 6 thr are rejected before any arithmetic.
struct CullBox {            // what k_C_cull reads of one box: a kernel PARAMETER (constant bank: the operands of the culls need no loads)
    double aabbB1[6];
    double edgeAngle[12];
    double aabbFc[12][6], aabbFd[12][6];   // AABBs of the two box triangles next to box edge k (aabbF1[edgeFaces1[k][0 / 1]])
    double cullLo[12][3], cullHi[12][3];   // AABB of box edge k widened by 6 thr: the conservative pair cull is six comparisons
};
inline void make_cull_box(CullBox &C, const BoxData &B, double threshold) {
    for (int r = 0; r < 6; ++r) C.aabbB1[r] = B.aabbB1[r];
    for (int k = 0; k < 12; ++k) {
        C.edgeAngle[k] = B.edgeAngle[k];
        for (int r = 0; r < 6; ++r) { C.aabbFc[k][r] = B.aabbF1[h_edgeFaces1[k][0]][r]; C.aabbFd[k][r] = B.aabbF1[h_edgeFaces1[k][1]][r]; }
        for (int r = 0; r < 3; ++r) { C.cullLo[k][r] = B.aabbEdge[k][r] - 6.0 * threshold; C.cullHi[k][r] = B.aabbEdge[k][3 + r] + 6.0 * threshold; }
    }
}
__device__ __forceinline__ bool pair_culled(int k1, const CullBox &C, const double *aabbE2k, double threshold) {
    // the conservative test first: it removes nearly every pair with six comparisons; the reference's own (exact) tests then run
    // on the few survivors — the order of rejections does not matter, a pair is dropped if any of them fires
    const double *lo = C.cullLo[k1], *hi = C.cullHi[k1];
    if (aabbE2k[3] < lo[0] || aabbE2k[0] > hi[0] || aabbE2k[4] < lo[1] || aabbE2k[1] > hi[1] || aabbE2k[5] < lo[2] || aabbE2k[2] > hi[2]) return true;
    if (C.edgeAngle[k1] < M_PI / 6.0) return true;   // soft edge
    return !check_aabb(C.aabbFc[k1], aabbE2k) && !check_aabb(C.aabbFd[k1], aabbE2k);
}

// e2->normal (:851-856 reads what createEdges :193-214 stored): the normalised sum of the normals of the edge's faces on the
// UNPERTURBED verts.  Only the first rejection of :877-881 reads it: pass 1 forms it for the pairs that reach that test, pass 2 never.
__device__ __forceinline__ V3 edge_normal(const EdgeRec &e2, const int32_t *__restrict__ fn, const double *__restrict__ x0) {
    V3 n0 = face_normal_0(fn, x0, e2.f[0]);
    V3 n1 = e2.f[1] >= 0 ? face_normal_0(fn, x0, e2.f[1]) : mk(0, 0, 0);   // boundary: normals[1] stays zero (:94-95)
    return normalized(n0 + n1);
}

// the rest of :873-1017 for a pair that survived pair_culled(); out of line: the survivors are few and the callers stay light
// rec == NULL (pass 1): the verdict.  rec != NULL (pass 2, a pair that passed): the record; fn / x0 are not read.
__device__ __noinline__ bool test_edge_edge(int k1, const BoxData &B, V3 x2a, V3 x2b, const int32_t *__restrict__ fn, const double *__restrict__ x0,
                               double threshold, eolc_contact *rec, const EdgeRec &e2, int k2) {
    const V3 dx2 = x2b - x2a;
    const double len2 = norm(dx2);
    V3 x1a = bcol(B.verts1, c_edgeVerts1[k1][0]), x1b = bcol(B.verts1, c_edgeVerts1[k1][1]);
    V3 dx1 = bcol(B.dx1, k1);
    double len1 = B.len1[k1];
    V3 tan1 = bcol(B.tan1, k1);
    // :877-887: acos(c) within 2 deg of 0 or pi, decided in cosine space against the host libm's critical doubles (BoxData);
    // c outside [-1, 1] or NaN gives a NaN angle in the reference, which fails both comparisons
    // The first of the two (tan1 against the cloth edge's normal, :877-881) is the only use of that normal — two face normals and
    // three normalisations behind two more dependent gathers.  Every rejection here is a pure function of the inputs, so the order is
    // free: the cheap one runs first, and pass 2 (whose pairs have passed) does not form the normal at all.
    double c = dv(dot(tan1, dx2), len2);
    if ((c >= B.cosParHi && c <= 1.0) || (c <= B.cosParLo && c >= -1.0)) return false;
    if (!rec) {
        c = dot(tan1, edge_normal(e2, fn, x0));
        if ((c >= B.cosParHi && c <= 1.0) || (c <= B.cosParLo && c >= -1.0)) return false;
    }
    V3 nor = normalized(cross(dx1, dx2));
    V3 x1c = bcol(B.verts1, c_edgeVerts1[k1][2]), x1d = bcol(B.verts1, c_edgeVerts1[k1][3]);
    V3 n1c = bcol(B.faceNors1, c_edgeFaces1[k1][0]), n1d = bcol(B.faceNors1, c_edgeFaces1[k1][1]);
    V3 nor1 = bcol(B.nor1e, k1);
    if (dot(nor, nor1) < 0.0) nor = neg(nor);
    c = dot(n1c, nor);   // :906-915: angleCN - angleCD > 2 deg (angleCN < -2 deg cannot happen: an acos)
    if (c >= -1.0 && c <= 1.0 && c < B.cosWedge[k1]) return false;
    // the reference intersects first (:920-924) and then rejects on the line-line parameters (:927-936); both are pure functions of
    // the same inputs, so the cheap rejection runs first and the eight ray / triangle tests only for its survivors
    double u1, u2;
    lineline(u1, u2, x1a, x1b, x2a, x2b);
    double thresh1 = dv(mul(1.0, threshold), len1), thresh2 = dv(mul(1.0, threshold), len2);
    if (u1 < -thresh1 || u1 > add(1.0, thresh1) || u2 < -thresh2 || u2 > add(1.0, thresh2)) return false;
    double u2c, u2d;
    int i2c = intersect_square(x2a, dx2, x1a, x1b, x1c, u2c);
    int i2d = intersect_square(x2a, dx2, x1b, x1a, x1d, u2d);
    i2c = i2c && (0.0 <= u2c && u2c <= 1.0);
    i2d = i2d && (0.0 <= u2d && u2d <= 1.0);
    V3 x1, x2;
    if (i2c && i2d) {
        x1 = scale(sub(1.0, u1), x1a) + scale(u1, x1b);
        x2 = scale(sub(1.0, u2), x2a) + scale(u2, x2b);
    } else if (i2c && !i2d) {
        if (dot(x2a - x1a, n1c) < 0.0) u2 = fmax(0.0, fmin(u2c, u2));
        else u2 = fmax(u2c, fmin(1.0, u2));
        x2 = scale(sub(1.0, u2), x2a) + scale(u2, x2b);
        u1 = linepoint(x1a, x1b, x2);
        x1 = scale(sub(1.0, u1), x1a) + scale(u1, x1b);
    } else if (!i2c && i2d) {
        if (dot(x2a - x1b, n1d) < 0.0) u2 = fmax(0.0, fmin(u2d, u2));
        else u2 = fmax(u2d, fmin(1.0, u2));
        x2 = scale(sub(1.0, u2), x2a) + scale(u2, x2b);
        u1 = linepoint(x1a, x1b, x2);
        x1 = scale(sub(1.0, u1), x1a) + scale(u1, x1b);
    } else {
        return false;
    }
    if (u1 < -thresh1 || u1 > add(1.0, thresh1) || u2 < -thresh2 || u2 > add(1.0, thresh2)) return false;
    V3 dx = x2 - x1;
    double thresh = mul(2.0, threshold);
    if (dot(dx, dx) > mul(thresh, thresh)) return false;
    if (rec) {
        zero_contact(*rec);
        rec->dist = norm(dx);
        st(rec->nor1, nor1); st(rec->nor2, nor); st(rec->pos1, x1); st(rec->pos2, x2);
        rec->count1 = 2; rec->count2 = 2;
        rec->verts1[0] = c_edgeVerts1[k1][0]; rec->verts1[1] = c_edgeVerts1[k1][1]; rec->verts1[2] = -1;
        rec->verts2[0] = e2.v[0]; rec->verts2[1] = e2.v[1]; rec->verts2[2] = -1;
        rec->weights1[0] = sub(1.0, u1); rec->weights1[1] = u1; rec->weights1[2] = 0.0;
        rec->weights2[0] = sub(1.0, u2); rec->weights2[1] = u2; rec->weights2[2] = 0.0;
        rec->edge1[0] = k1; rec->n_edge1 = 1;
        rec->edge2 = k2;
        st(rec->edgeDir, tan1);
    }
    return true;
}

// Pass 1 of section C, in three light / dense kernels instead of one divergent one:
//   k_C_cull   one thread per cloth edge: the AABB rejections of its 12 pairs; survivors are appended to a work list
//   k_C_test   one thread per surviving pair (all lanes busy): the rest of :873-1017; a hit sets its bit in the edge's mask
//   (the hits per 256-item block, for the scan, are counted by k_C_test with an integer atomic per hit)
// The work list's order does not matter (integer atomics): a pair's outcome lands in its own bit.  If the list overflows, the
// pairs that did not fit are tested in place by k_C_cull.
// One launch per box, the box's cull data a __grid_constant__ parameter: the first version staged it in shared memory per 256-edge
// CTA (a global load, a barrier and ~200 shared-memory loads per thread in front of a kernel that is a chain of dependent latencies:
// record -> end points -> list slot).  grid: (nC / 256, S)
#ifndef EOLC_CULL_CTAS
#define EOLC_CULL_CTAS 4
#endif
#ifndef EOLC_CULL_EPT
#define EOLC_CULL_EPT 1      // edges per thread (items 256 apart)
#endif
__global__ void __launch_bounds__(256, EOLC_CULL_CTAS) k_C_cull(int E, int nC, int b, const __grid_constant__ CullBox C, const EdgeRec *__restrict__ edges,
                                                const double *__restrict__ xp, double threshold,
                                                int *__restrict__ info, int *__restrict__ blocksum, unsigned long long *__restrict__ cand_list,
                                                int *__restrict__ counter, int capacity, size_t xstride, size_t scene_items, size_t box_items,
                                                size_t secC_off) {
    const int s = blockIdx.y;
    const size_t item0 = s * scene_items + secC_off + b * box_items;
    const double *xs = xp + s * xstride;
    int k2[EOLC_CULL_EPT], cand[EOLC_CULL_EPT];
    EdgeRec e2[EOLC_CULL_EPT];
#pragma unroll
    for (int i = 0; i < EOLC_CULL_EPT; ++i) {
        k2[i] = (blockIdx.x * EOLC_CULL_EPT + i) * 256 + threadIdx.x;
        if (k2[i] < E) e2[i] = edges[k2[i]];
        else { e2[i].v[0] = e2[i].v[1] = 0; }
        if (threadIdx.x == 0 && k2[i] < nC) blocksum[(item0 + k2[i]) / 256] = 0;          // k_C_test counts the block's hits
    }
    V3 xa[EOLC_CULL_EPT], xb[EOLC_CULL_EPT];
#pragma unroll
    for (int i = 0; i < EOLC_CULL_EPT; ++i) { xa[i] = dcol(xs, e2[i].v[0]); xb[i] = dcol(xs, e2[i].v[1]); }
    int n = 0;
#pragma unroll
    for (int i = 0; i < EOLC_CULL_EPT; ++i) {
        cand[i] = 0;
        if (k2[i] < E) {
            double aabbE[6];                                                      // build_AABB_E :455-468
            aabbE[0] = fmin(xb[i].x, xa[i].x); aabbE[1] = fmin(xb[i].y, xa[i].y); aabbE[2] = fmin(xb[i].z, xa[i].z);
            aabbE[3] = fmax(xb[i].x, xa[i].x); aabbE[4] = fmax(xb[i].y, xa[i].y); aabbE[5] = fmax(xb[i].z, xa[i].z);
            // whole-box cull first: an edge outside the padded box AABB fails all 24 face-AABB tests
            if (check_aabb(C.aabbB1, aabbE)) {
#pragma unroll
                for (int k1 = 0; k1 < 12; ++k1)
                    if (!pair_culled(k1, C, aabbE, threshold)) cand[i] |= 1 << k1;
            }
        }
        n += __popc(cand[i]);
        if (k2[i] < nC) info[item0 + k2[i]] = 0;
    }
    // one atomic per WARP (the per-lane atomics on the single counter were 11 % of the kernel's stall samples, ncu r02k): the lanes'
    // counts are scanned in the warp, lane 31 reserves the warp's range.  The counter keeps counting past the capacity: the host
    // then grows the list and repeats the pass.
    const int lane = threadIdx.x & 31;
    if (!__any_sync(0xffffffffu, n)) return;
    int inc = n;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int at = base + inc - n;
#pragma unroll
    for (int i = 0; i < EOLC_CULL_EPT; ++i)
        for (int k1 = 0; k1 < 12; ++k1)
            if ((cand[i] >> k1) & 1) {
                if (at < capacity) cand_list[at] = (unsigned long long)(item0 + k2[i]) | ((unsigned long long)k1 << 60);
                ++at;
            }
}
#ifndef EOLC_CTEST_CTAS
#define EOLC_CTEST_CTAS 2
#endif
__global__ void __launch_bounds__(256, EOLC_CTEST_CTAS) k_C_test(int nB, const EdgeRec *__restrict__ edges, const double *__restrict__ xp,
                                                const int32_t *__restrict__ fn, const double *__restrict__ x0, const BoxData *__restrict__ boxes, double threshold,
                                                int *__restrict__ info, int *__restrict__ blocksum, const unsigned long long *__restrict__ cand_list,
                                                const int *__restrict__ counter, int capacity, size_t xstride,
                                                size_t scene_items, size_t box_items, size_t secC_off) {
    const int n = min(*counter, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const unsigned long long c = cand_list[i];
        const size_t item = (size_t)(c & ((1ull << 60) - 1));
        const int k1 = (int)(c >> 60);
        const size_t s = item / scene_items, rem = item - s * scene_items - secC_off;
        const int b = (int)(rem / box_items), k2 = (int)(rem - (size_t)b * box_items);
        const EdgeRec e2 = edges[k2];
        const V3 x2a = dcol(xp + s * xstride, e2.v[0]), x2b = dcol(xp + s * xstride, e2.v[1]);
        if (test_edge_edge(k1, boxes[b], x2a, x2b, fn, x0 + s * xstride, threshold, nullptr, e2, k2)) {
            atomicOr(info + item, 1 << k1);
            atomicAdd(blocksum + item / 256, 1);      // hits per 256-item block, for the scan (integer atomics: the same counts every run)
        }
    }
}
// Pass 2 of section C.  Hits are few (a band of cloth edges along the box edges) and scattered over the item blocks; re-deriving a
// record is ~2 k FP64 instructions.  k_C_expand (light, full occupancy) turns every hit bit into a work item carrying its final
// output slot; k_C_write then runs one thread per hit, all lanes busy.  The order of the work list is irrelevant (appended with an
// integer atomic): every item writes its own, predetermined slot, so the output is the same bits every run.
struct CHit { int32_t k2; int32_t sb_k1; int32_t slot; int32_t pad; };   // sb_k1 = (scene * nB + box) | k1 << 16
// One WARP per 256-item block (8 per CTA; a CTA per item block spent its time being scheduled: 196 k CTAs that mostly read two offsets
// and return): lane l takes the items 8 l .. 8 l + 7, so the hits keep their item order.
__global__ void __launch_bounds__(256) k_C_expand(int nbx, long long nblk, int nB, const int *__restrict__ info, const int *__restrict__ blockoff,
                                                  CHit *__restrict__ work, int *__restrict__ counter, int capacity, size_t scene_items,
                                                  size_t box_items, size_t secC_off) {
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= nblk) return;
    const int lane = threadIdx.x & 31;
    const int bx = (int)(w % nbx);
    const long long sb = w / nbx;
    const size_t item0 = (size_t)(sb / nB) * scene_items + secC_off + (size_t)(sb % nB) * box_items;
    const size_t blk = item0 / 256 + bx;
    const int base = blockoff[blk];
    if (blockoff[blk + 1] == base) return;            // warp-uniform
    const int4 *p = reinterpret_cast<const int4 *>(info + item0 + (size_t)bx * 256 + 8 * lane);
    const int4 m0 = p[0], m1 = p[1];
    const int m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) cnt += __popc(m[q]);
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);     // > 0: the block has hits
    int wbase = 0;
    if (lane == 31) wbase = atomicAdd(counter, total);          // one atomic per warp
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    if (!cnt) return;
    int slot = base + inc - cnt;
    int at = wbase + inc - cnt;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k2 = bx * 256 + 8 * lane + q;
        for (int k1 = 0; k1 < 12; ++k1)
            if (m[q] & (1 << k1)) {
                if (at < capacity) work[at] = CHit{k2, (int)sb | (k1 << 16), slot, 0};
                ++at; ++slot;
            }
    }
}
__global__ void __launch_bounds__(256, 2) k_C_write(int nB, const EdgeRec *__restrict__ edges, const double *__restrict__ xp,
                                                 const int32_t *__restrict__ fn, const double *__restrict__ x0, const BoxData *__restrict__ boxes, double threshold,
                                                 const CHit *__restrict__ work, const int *__restrict__ counter, int capacity,
                                                 eolc_contact *__restrict__ out, size_t xstride, int remap, int nP, int out_cap) {
    const int n = min(*counter, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const CHit h = work[i];
        const int sb = h.sb_k1 & 0xffff, k1 = h.sb_k1 >> 16, s = sb / nB, b = sb % nB;
        const EdgeRec e2 = edges[h.k2];
        const V3 x2a = dcol(xp + s * xstride, e2.v[0]), x2b = dcol(xp + s * xstride, e2.v[1]);
        eolc_contact rec;
        test_edge_edge(k1, boxes[b], x2a, x2b, nullptr, nullptr, threshold, &rec, e2, h.k2);
        finish_contact(rec, threshold);
        if (remap) remap_contact(rec, nP + b * 8 + b * 12);
        if (h.slot < out_cap) out[h.slot] = rec;
    }
}

inline size_t pad256(size_t n) { return (n + 255) / 256 * 256; }

}  // namespace

struct eolc_cd_plan {
    eolc_ctx *ctx = nullptr;
    int device = 0;                               // own copy: the plan may be destroyed after its ctx (thread_local order)
    int32_t N = 0, F = 0, E = 0;
    double threshold = 0;
    std::vector<EdgeRec> h_edges;
    DevBuf<int32_t> d_fn;
    DevBuf<EdgeRec> d_edges;
    DevBuf<double> d_r;                 // perturbation table, 3N
    // per-run buffers
    DevBuf<double> d_x, d_xp, d_fnp, d_aabb, d_aabb_part, d_pxyz, d_pnorms;
    DevBuf<BoxData> d_boxes;
    DevBuf<int> d_info, d_blocksum, d_blockoff;
    DevBuf<Cand> d_partial;
    DevBuf<eolc_contact> d_out;
    PinnedBuf<eolc_contact> p_out;
    PinnedBuf<int> p_blockoff, p_counter;
    PinnedBuf<double> p_x;
    int64_t last_pair_tests = 0;
    int32_t last_launches = 0;
    int32_t last_total = 0;             // contacts of the last run, still in d_out
    bool smem_attr_set = false;
    DevBuf<CHit> d_chits;               // work list of section C's write pass
    DevBuf<int> d_counter;              // [0]: pairs in d_cands, [1]: hits in d_chits
    DevBuf<unsigned long long> d_cands; // work list of section C's test pass: item | k1 << 60
    DevBuf<int> d_scan_tmp;             // multi-block scan: chunk totals and their offsets
    DevBuf<int32_t> d_rows_i;           // contact rows (eolc_cd_contact_rows): row_nnz [n] + cols [9 n]
    DevBuf<double> d_rows_v;            // vals [9 n]
    DevBuf<unsigned char> d_eol;
    PinnedBuf<int32_t> p_rows_i;
    PinnedBuf<double> p_rows_v;
};

extern "C" {

int eolc_cd_plan_create(eolc_ctx *ctx, int32_t N, int32_t F, const int32_t *face_nodes, double threshold, eolc_cd_plan **out) {
    EOLC_REQUIRE(ctx && out, "ctx/out is NULL");
    *out = nullptr;
    EOLC_REQUIRE(N >= 0 && F >= 0 && (F == 0 || face_nodes), "bad arguments");
    for (int64_t i = 0; i < 3 * (int64_t)F; ++i) EOLC_REQUIRE(face_nodes[i] >= 0 && face_nodes[i] < N, "face node index out of range");
    EOLC_CUDA(cudaSetDevice(ctx->device));
    eolc_cd_plan *P = new eolc_cd_plan;
    P->ctx = ctx; P->device = ctx->device; P->N = N; P->F = F; P->threshold = threshold;
    // createEdges (:141-231) with an int64 key: stable sort of the 3F half-edges by (max+1)*(3F+1) + (min+1)
    {
        struct HE { int64_t key; int32_t face; int8_t i; };
        const int64_t n = 3 * (int64_t)F;
        std::vector<HE> tmp((size_t)n);
        for (int32_t k = 0; k < F; ++k)
            for (int i = 0; i < 3; ++i) {
                int32_t a = face_nodes[3 * (size_t)k + i], b = face_nodes[3 * (size_t)k + (i + 1) % 3];
                int64_t kmin = std::min(a, b) + 1, kmax = std::max(a, b) + 1;
                tmp[3 * (size_t)k + i] = {kmin + (n + 1) * kmax, k, (int8_t)i};
            }
        std::stable_sort(tmp.begin(), tmp.end(), [](const HE &a, const HE &b) { return a.key < b.key; });
        int64_t k = 0;
        while (k < n) {
            const HE &h0 = tmp[k];
            const int32_t *f0 = face_nodes + 3 * (size_t)h0.face;
            EdgeRec e;
            e.v[0] = f0[h0.i]; e.v[1] = f0[(h0.i + 1) % 3]; e.v[2] = f0[(h0.i + 2) % 3];
            e.f[0] = h0.face;
            if (k < n - 1 && tmp[k].key == tmp[k + 1].key) {
                const HE &h1 = tmp[k + 1];
                e.v[3] = face_nodes[3 * (size_t)h1.face + (h1.i + 2) % 3];
                e.f[1] = h1.face;
                k += 2;
            } else {
                e.v[3] = -1; e.f[1] = -1;
                k += 1;
            }
            P->h_edges.push_back(e);
        }
        P->E = (int32_t)P->h_edges.size();
    }
    // perturbation table (:648-659): libstdc++ mt19937 seed 1 + uniform_real_distribution(-1,1), r = dis*thr*1e-3
    std::vector<double> r(3 * (size_t)N);
    {
        std::mt19937 gen;
        std::uniform_real_distribution<> dis(-1.0, 1.0);
        gen.seed(1);
        for (size_t i = 0; i < r.size(); ++i) r[i] = dis(gen) * threshold * 1e-3;
    }
    cudaStream_t st = ctx->stream;
    std::vector<int32_t> fnv(face_nodes, face_nodes + 3 * (size_t)F);
    cudaError_t ce = P->d_fn.upload(fnv, st);
    if (ce == cudaSuccess) ce = P->d_edges.upload(P->h_edges, st);
    if (ce == cudaSuccess) ce = P->d_r.upload(r, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { delete P; set_error("cd plan upload failed: %s", cudaGetErrorString(ce)); return EOLC_ERR_CUDA; }
    *out = P;
    return EOLC_OK;
}

void eolc_cd_plan_destroy(eolc_cd_plan *plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    delete plan;
}

int eolc_cd_edge_count(const eolc_cd_plan *plan) { return plan ? plan->E : 0; }

int eolc_cd_edge_table(const eolc_cd_plan *plan, int32_t *out6E) {
    EOLC_REQUIRE(plan && out6E, "NULL argument");
    for (int32_t k = 0; k < plan->E; ++k) {
        for (int i = 0; i < 4; ++i) out6E[6 * (size_t)k + i] = plan->h_edges[k].v[i];
        out6E[6 * (size_t)k + 4] = plan->h_edges[k].f[0]; out6E[6 * (size_t)k + 5] = plan->h_edges[k].f[1];
    }
    return EOLC_OK;
}

int eolc_cd_last_stats(const eolc_cd_plan *plan, int64_t *pair_tests, int32_t *launches) {
    EOLC_REQUIRE(plan, "plan is NULL");
    if (pair_tests) *pair_tests = plan->last_pair_tests;
    if (launches) *launches = plan->last_launches;
    return EOLC_OK;
}

int eolc_cd_angle_cuts(const double *box_whd, const double *box_E, double *cuts14) {
    EOLC_REQUIRE(box_whd && box_E && cuts14, "NULL argument");
    BoxData B;
    bool ok = make_box(B, box_whd, box_E);
    cuts14[0] = B.cosParHi; cuts14[1] = B.cosParLo;
    for (int k = 0; k < 12; ++k) cuts14[2 + k] = B.cosWedge[k];
    if (!ok) { set_error("eolc_cd_angle_cuts: this libm's acos is not monotone next to a decision threshold"); return EOLC_ERR_UNSUPPORTED; }
    return EOLC_OK;
}

}  // extern "C" (reopened below)

// One run.  resident: the records stay in the plan's device buffer (eolc_cd_contacts_dev / eolc_cd_contact_rows read them there);
// only the per-scene offsets come back to the host.
static int cd_run_impl(eolc_cd_plan *plan, int32_t S, const double *x_dev, int32_t n_points, const double *pxyz,
                       const double *pnorms, int32_t n_boxes, const double *box_whd, const double *box_E,
                       int point_eol_flag, int remap_box_indices, eolc_contact *out, int32_t capacity,
                       int32_t *scene_offset, bool resident) {
    EOLC_REQUIRE(plan && scene_offset, "NULL argument");
    eolc_cd_plan *P = plan;
    EOLC_REQUIRE(S >= 1 && n_points >= 0 && n_boxes >= 0 && capacity >= 0, "bad sizes");
    EOLC_REQUIRE(P->N == 0 || x_dev, "x_dev is NULL");
    EOLC_REQUIRE(n_points == 0 || (pxyz && pnorms), "pxyz/pnorms is NULL");
    EOLC_REQUIRE(n_boxes == 0 || (box_whd && box_E), "box_whd/box_E is NULL");
    EOLC_REQUIRE(capacity == 0 || out, "out is NULL");
    EOLC_REQUIRE((int64_t)S * std::max(std::max(n_boxes, n_points), 1) <= 65535,      // a grid dimension, and 16 bits of CHit::sb_k1
                 "too many scenes * max(boxes, points) for one run; split the batch");
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    const int N = P->N, F = P->F, E = P->E, nB = n_boxes, nP = n_points;
    const double thr = P->threshold;
    int launches = 0;
    // ---- item layout of one scene: [PE: pad(N)] [PT: pad(P)] then per box [A: pad(N)] [Bc: 256] [C: pad(E)]
    const bool doPE = point_eol_flag && nP > 0;   // with no points the PE loop emits nothing
    const size_t secPE = 0, nPE = doPE ? pad256(N) : 0;
    const size_t secPT = secPE + nPE, nPT = nP > 0 ? pad256(nP) : 0;
    const size_t secBox = secPT + nPT;
    const size_t nA = pad256(N), nBc = 256, nC = pad256(E);
    const size_t box_items = nA + nBc + nC;
    const size_t scene_items = secBox + (size_t)nB * box_items;
    const size_t total_items = scene_items * S;
    const size_t nblocks = total_items / 256;
    for (int s = 0; s <= S; ++s) scene_offset[s] = 0;
    if (total_items == 0 || N == 0) { P->last_pair_tests = 0; P->last_launches = 0; P->last_total = 0; return EOLC_OK; }
    EOLC_REQUIRE(nblocks < ((size_t)1 << 31), "batch too large");

    // ---- constants to the device
    std::vector<BoxData> hb((size_t)nB);
    for (int b = 0; b < nB; ++b)
        if (!make_box(hb[b], box_whd + 3 * b, box_E + 16 * b))
        { set_error("eolc_cd_run: this libm's acos is not monotone next to a decision threshold"); return EOLC_ERR_UNSUPPORTED; }
    EOLC_CUDA(P->d_boxes.ensure(std::max(nB, 1)));
    if (nB) EOLC_CUDA(cudaMemcpyAsync(P->d_boxes.p, hb.data(), sizeof(BoxData) * nB, cudaMemcpyHostToDevice, st));
    EOLC_CUDA(P->d_pxyz.ensure(3 * (size_t)std::max(nP, 1))); EOLC_CUDA(P->d_pnorms.ensure(3 * (size_t)std::max(nP, 1)));
    if (nP) {
        EOLC_CUDA(cudaMemcpyAsync(P->d_pxyz.p, pxyz, sizeof(double) * 3 * nP, cudaMemcpyHostToDevice, st));
        EOLC_CUDA(cudaMemcpyAsync(P->d_pnorms.p, pnorms, sizeof(double) * 3 * nP, cudaMemcpyHostToDevice, st));
    }
    const size_t xs = 3 * (size_t)N, fs = 3 * (size_t)F;
    EOLC_CUDA(P->d_xp.ensure(xs * S));
    if (nP) EOLC_CUDA(P->d_fnp.ensure(std::max<size_t>(fs * S, 1)));   // stored face normals: obstacle points only
    EOLC_CUDA(P->d_aabb.ensure(6 * (size_t)S));
    EOLC_CUDA(P->d_info.ensure(total_items)); EOLC_CUDA(P->d_blocksum.ensure(nblocks)); EOLC_CUDA(P->d_blockoff.ensure(nblocks + 1));

    // ---- prepare
    k_perturb<<<dim3((unsigned)((xs + 255) / 256), S), 256, 0, st>>>(N, x_dev, P->d_r.p, P->d_xp.p, xs); ++launches;
    if (F && nP) { k_face_normals<<<dim3((F + 255) / 256, S), 256, 0, st>>>(F, P->d_fn.p, P->d_xp.p, P->d_fnp.p, xs, fs); ++launches; }
    {
        const int nb = std::max(1, std::min((N + 2047) / 2048, 64));
        EOLC_CUDA(P->d_aabb_part.ensure(6 * (size_t)S * nb));
        k_aabb<<<dim3(nb, S), 256, 0, st>>>(N, 3, P->d_xp.p, P->d_aabb_part.p, xs);
        k_aabb<<<dim3(1, S), 256, 0, st>>>(nb, 6, P->d_aabb_part.p, P->d_aabb.p, (size_t)6 * nb);
        launches += 2;
    }

    // ---- pass 1
    const int nchunk = std::max(1, std::min((F + 256 * 8 - 1) / (256 * 8), 2 * P->ctx->sm_count));
    const size_t npart = (size_t)S * std::max(8 * nB, nP) * nchunk;
    EOLC_CUDA(P->d_partial.ensure(std::max<size_t>(npart, 1)));
    if (doPE) { k_PE_count<<<dim3((unsigned)(nPE / 256), S), 256, 0, st>>>(N, nP, P->d_pxyz.p, P->d_xp.p, thr, P->d_info.p, P->d_blocksum.p, xs, scene_items, secPE); ++launches; }
    if (nP) {
        k_PT_partial<<<dim3(nchunk, S * nP), 256, 0, st>>>(F, nP, 0, nullptr, P->d_pxyz.p, P->d_pnorms.p, P->d_fn.p, P->d_xp.p, P->d_fnp.p, P->d_aabb.p, thr, P->d_partial.p, xs, fs);
        k_PT_final<<<S, 256, 0, st>>>(nchunk, nP, 1, P->d_partial.p, P->d_info.p, P->d_blocksum.p, scene_items, secPT, 0);
        launches += 2;
    }
    if (nB) {
        for (int b = 0; b < nB; ++b) {
            ABox ab;
            make_a_box(ab, hb[b]);
            k_A_count<<<dim3((unsigned)(nA / 256), S), 256, 0, st>>>(N, b, ab, P->d_xp.p, thr, P->d_info.p, xs, scene_items, box_items, secBox);
        }
        {
            const long long nblkA = (long long)(nA / 256) * S * nB;
            k_A_sum<<<(unsigned)((nblkA + 7) / 8), 256, 0, st>>>((int)(nA / 256), nblkA, P->d_info.p, P->d_blocksum.p, scene_items, box_items, secBox, nB);
            ++launches;
        }
        // chunks of a scene's faces per (scene, box): enough CTAs to fill the GPU and no more — every CTA pays the staging of the corners
        // and an eight-corner reduction, which a batch of thousands of small scenes would otherwise pay four times per scene
        const int nchunk8 = (int)std::max<long long>(1, std::min<long long>(nchunk, (EOLC_PT8_WAVES * P->ctx->sm_count + (long long)S * nB - 1) / ((long long)S * nB)));
        k_PT_partial8<<<dim3(nchunk8, S * nB), 256, 0, st>>>(F, nB, P->d_boxes.p, P->d_fn.p, P->d_xp.p, P->d_aabb.p, thr, P->d_partial.p, xs);
        k_PT_final<<<S * nB, 256, 0, st>>>(nchunk8, 8, nB, P->d_partial.p, P->d_info.p, P->d_blocksum.p, scene_items, secBox + nA, box_items);
        launches += 2 + nB;
    }
    EOLC_CUDA(P->p_blockoff.ensure(nblocks + 1));
    EOLC_CUDA(P->d_counter.ensure(2)); EOLC_CUDA(P->p_counter.ensure(2));
    const long long nblkC = (long long)(nC / 256) * S * nB;
    EOLC_REQUIRE(nblkC * 256 * 12 < (long long)INT32_MAX, "too many (cloth edge, box edge) pairs for one run; split the batch");   // the pair counter is an int
    long long cap = std::min<long long>(std::max<long long>((long long)P->d_cands.n, nblkC * 128 + 65536), (long long)1 << 30);   // half a pair per edge, and then some
    if (const char *ev = getenv("EOLC_CD_PAIR_CAP")) cap = std::max<long long>(1, atoll(ev));   // test knob: forces the overflow path
    // ---- pass 2 (records incl. step (E) pos1_ and the CD index remap, written on the device in final order) as a callable: the
    // device-resident entry launches it SPECULATIVELY, right behind the scan and before the host has seen any size, whenever the
    // plan's record / hit buffers exist from an earlier run — every write is guarded by the buffers' capacities — and the host then
    // only confirms that the totals fitted.  A simulation's contact count changes slowly from step to step, so the usual run has no
    // point at which the GPU waits for the host (the round trip in the middle cost ~35 us: a quarter of a 512-scene batch's call).
    auto pass2 = [&](int out_cap, int chit_cap) -> int {
        if (doPE) { k_PE_write<<<dim3((unsigned)(nPE / 256), S), 256, 0, st>>>(N, nP, P->d_pxyz.p, P->d_pnorms.p, P->d_xp.p, thr, P->d_info.p, P->d_blockoff.p, P->d_out.p, xs, scene_items, secPE, out_cap); ++launches; }
        if (nP) { k_PT_write<<<dim3((unsigned)(nPT / 256), S), 256, 0, st>>>(F, nP, 1, nullptr, P->d_pxyz.p, P->d_pnorms.p, P->d_fn.p, P->d_xp.p, P->d_fnp.p, thr, P->d_info.p, P->d_blockoff.p, P->d_out.p, xs, fs, scene_items, secPT, 0, 0, nP, out_cap); ++launches; }
        if (nB) {
            {
                const long long nvb = (long long)(nA / 256) * S * nB;
                const int gridA = (int)std::min<long long>(nvb, (long long)P->ctx->sm_count * 3);
                const size_t smemA = 256 * sizeof(eolc_contact);
                if (!P->smem_attr_set) {
                    EOLC_CUDA(cudaFuncSetAttribute(k_A_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
                    P->smem_attr_set = true;
                }
                k_A_write<<<gridA, 256, smemA, st>>>(N, F, nB, S, P->d_xp.p, P->d_fn.p, P->d_boxes.p, thr, P->d_info.p, P->d_blockoff.p, P->d_out.p, xs, scene_items, box_items, secBox, out_cap);
            }
            k_PT_write<<<dim3(1, S * nB), 256, 0, st>>>(F, 8, nB, P->d_boxes.p, nullptr, nullptr, P->d_fn.p, P->d_xp.p, nullptr, thr, P->d_info.p, P->d_blockoff.p, P->d_out.p, xs, fs, scene_items, secBox + nA, box_items, remap_box_indices, nP, out_cap);
            {
                // every record has its slot: the work list of section C holds at most as many items as there are records
                EOLC_CUDA(cudaMemsetAsync(P->d_counter.p + 1, 0, sizeof(int), st));
                k_C_expand<<<(unsigned)((nblkC + 7) / 8), 256, 0, st>>>((int)(nC / 256), nblkC, nB, P->d_info.p, P->d_blockoff.p, P->d_chits.p, P->d_counter.p + 1, chit_cap, scene_items, box_items, secBox + nA + nBc);
                const int gridC = std::max(1, std::min((chit_cap + 255) / 256, P->ctx->sm_count * 2));
                k_C_write<<<gridC, 256, 0, st>>>(nB, P->d_edges.p, P->d_xp.p, P->d_fn.p, x_dev, P->d_boxes.p, thr, P->d_chits.p, P->d_counter.p + 1, chit_cap, P->d_out.p, xs, remap_box_indices, nP, out_cap);
                ++launches;
            }
            launches += 3;
        }
        return EOLC_OK;
    };
    bool speculated = false;
    int spec_cap = 0;
    for (int attempt = 0;; ++attempt) {
        if (nB) {
            EOLC_CUDA(P->d_cands.ensure((size_t)cap));
            EOLC_CUDA(cudaMemsetAsync(P->d_counter.p, 0, 2 * sizeof(int), st));
            for (int b = 0; b < nB; ++b) {
                CullBox cb;
                make_cull_box(cb, hb[b], thr);
                k_C_cull<<<dim3((unsigned)((nC / 256 + EOLC_CULL_EPT - 1) / EOLC_CULL_EPT), S), 256, 0, st>>>(E, (int)nC, b, cb, P->d_edges.p, P->d_xp.p, thr, P->d_info.p, P->d_blocksum.p, P->d_cands.p, P->d_counter.p, (int)cap, xs, scene_items, box_items, secBox + nA + nBc);
            }
            const int gridT = (int)std::max<long long>(1, std::min<long long>(nblkC, (long long)P->ctx->sm_count * EOLC_CTEST_CTAS));
            k_C_test<<<gridT, 256, 0, st>>>(nB, P->d_edges.p, P->d_xp.p, P->d_fn.p, x_dev, P->d_boxes.p, thr, P->d_info.p, P->d_blocksum.p, P->d_cands.p, P->d_counter.p, (int)cap, xs, scene_items, box_items, secBox + nA + nBc);
            launches += 1 + nB;
        }
        if (nblocks <= (size_t)4 * SCAN_CHUNK) { k_scan<<<1, 1024, 0, st>>>(nblocks, P->d_blocksum.p, P->d_blockoff.p); ++launches; }
        else {
            const int nch = (int)((nblocks + SCAN_CHUNK - 1) / SCAN_CHUNK);
            EOLC_CUDA(P->d_scan_tmp.ensure(2 * (size_t)nch + 1));
            k_scan_local<<<nch, 1024, 0, st>>>(nblocks, P->d_blocksum.p, P->d_blockoff.p, P->d_scan_tmp.p);
            k_scan<<<1, 1024, 0, st>>>((size_t)nch, P->d_scan_tmp.p, P->d_scan_tmp.p + nch);
            k_scan_add<<<nch, 1024, 0, st>>>(nblocks, P->d_blockoff.p, P->d_scan_tmp.p + nch, nch);
            launches += 3;
        }
        speculated = false;
        cudaStream_t cs = P->ctx->copy_stream;
        if (resident && attempt == 0 && P->d_out.n > 1 && P->d_chits.n > 1 && cs && P->ctx->copy_event && !getenv("EOLC_CD_NO_SPECULATION")) {
            // the sizes travel on the second stream while pass 2 is already queued behind the scan; the host waits for the copy only
            // and returns with pass 2 still running, as it always did
            EOLC_CUDA(cudaEventRecord(P->ctx->copy_event, st));
            EOLC_CUDA(cudaStreamWaitEvent(cs, P->ctx->copy_event, 0));
            EOLC_CUDA(cudaMemcpyAsync(P->p_blockoff.p, P->d_blockoff.p, sizeof(int) * (nblocks + 1), cudaMemcpyDeviceToHost, cs));
            EOLC_CUDA(cudaMemcpyAsync(P->p_counter.p, P->d_counter.p, sizeof(int), cudaMemcpyDeviceToHost, cs));   // [1] is pass 2's
            spec_cap = (int)std::min<size_t>(std::min(P->d_out.n, P->d_chits.n), (size_t)INT32_MAX);
            const int rc2 = pass2(spec_cap, spec_cap);
            if (rc2) return rc2;
            speculated = true;
            EOLC_CUDA(cudaStreamSynchronize(cs));
        } else {
            EOLC_CUDA(cudaMemcpyAsync(P->p_blockoff.p, P->d_blockoff.p, sizeof(int) * (nblocks + 1), cudaMemcpyDeviceToHost, st));
            EOLC_CUDA(cudaMemcpyAsync(P->p_counter.p, P->d_counter.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            EOLC_CUDA(cudaStreamSynchronize(st));
        }
        if (!nB || P->p_counter.p[0] <= cap) break;
        if (attempt >= 1) { set_error("internal: the pair list of section C overflowed twice"); return EOLC_ERR_UNSUPPORTED; }
        cap = P->p_counter.p[0];        // the pass counted every surviving pair: this is the exact size
        speculated = false;             // a speculative pass 2 saw an incomplete pair list: it runs again below
    }
    const int total = P->p_blockoff.p[nblocks];

    // ---- capacity check before anything is written to the caller
    const int *bo = P->p_blockoff.p;
    for (int s = 0; s <= S; ++s) scene_offset[s] = bo[(size_t)s * scene_items / 256];
    if (!resident && total > capacity) {
        set_error("contact buffer too small: need %d, capacity %d", total, capacity);
        return EOLC_ERR_CAPACITY;
    }

    // ---- pass 2, unless the speculative one has done it (its capacities held)
    if (!(speculated && total <= spec_cap)) {
        // the device-resident entry keeps a quarter of headroom, so that the next steps of a simulation can speculate
        const size_t want = resident ? (size_t)total + (size_t)total / 4 + 1024 : (size_t)std::max(total, 1);
        if ((size_t)std::max(total, 1) > P->d_out.n) EOLC_CUDA(P->d_out.ensure(want));
        if ((size_t)total + 1 > P->d_chits.n) EOLC_CUDA(P->d_chits.ensure(want + 1));
        if (total > 0) {
            const int rc2 = pass2(total, total);
            if (rc2) return rc2;
        }
    }
    if (total > 0) {
        if (!resident) {
        // pinned caller buffer: DMA straight into it; pageable: via the plan's pinned staging
        cudaPointerAttributes pa;
        bool out_pinned = cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        eolc_contact *dst = out;
        if (!out_pinned) { EOLC_CUDA(P->p_out.ensure(total)); dst = P->p_out.p; }
        EOLC_CUDA(cudaMemcpyAsync(dst, P->d_out.p, sizeof(eolc_contact) * total, cudaMemcpyDeviceToHost, st));
        EOLC_CUDA(cudaStreamSynchronize(st));
        if (!out_pinned) memcpy(out, dst, sizeof(eolc_contact) * total);
        }
    }
    EOLC_CUDA(cudaGetLastError());
    P->last_launches = launches;
    P->last_total = total;
    P->last_pair_tests = (int64_t)S * ((doPE ? (int64_t)N * nP : 0) + (int64_t)nP * F + (int64_t)nB * ((int64_t)N * 24 + 8 * (int64_t)F + (int64_t)E * 12));

    // ---- step (D) (:1022-1052): per box corner keep only the closest count1 == 1 record.  Only section Bc emits
    // count1 == 1 records and it emits at most one per corner, so the reference's delete list is always empty and the
    // pass is the identity; this is verified per box from the (<= 8 record) corner section, O(1) per box (host copy only).
    if (!resident)
    for (int s = 0; s < S; ++s)
        for (int b = 0; b < nB; ++b) {
            const size_t bb = ((size_t)s * scene_items + secBox + (size_t)b * box_items + nA) / 256;
            const int begin = bo[bb], end = bo[bb + 1];
            int seen = 0;
            for (int k = begin; k < end; ++k) {
                int corner = out[k].verts1[0] - (remap_box_indices ? nP + b * 20 : 0);
                if (out[k].count1 != 1 || corner < 0 || corner > 7 || (seen & (1 << corner))) {
                    set_error("internal: corner section of box %d is not one-record-per-corner", b);
                    return EOLC_ERR_UNSUPPORTED;
                }
                seen |= 1 << corner;
            }
        }
    return EOLC_OK;
}

extern "C" {

int eolc_cd_run_batched_dev(eolc_cd_plan *plan, int32_t S, const double *x_dev, int32_t n_points, const double *pxyz,
                            const double *pnorms, int32_t n_boxes, const double *box_whd, const double *box_E,
                            int point_eol_flag, int remap_box_indices, eolc_contact *out, int32_t capacity,
                            int32_t *scene_offset) {
    return cd_run_impl(plan, S, x_dev, n_points, pxyz, pnorms, n_boxes, box_whd, box_E, point_eol_flag, remap_box_indices, out, capacity,
                       scene_offset, false);
}

int eolc_cd_run_batched_resident_dev(eolc_cd_plan *plan, int32_t S, const double *x_dev, int32_t n_points, const double *pxyz,
                                     const double *pnorms, int32_t n_boxes, const double *box_whd, const double *box_E,
                                     int point_eol_flag, int remap_box_indices, int32_t *scene_offset) {
    return cd_run_impl(plan, S, x_dev, n_points, pxyz, pnorms, n_boxes, box_whd, box_E, point_eol_flag, remap_box_indices, nullptr, 0,
                       scene_offset, true);
}

int eolc_cd_contacts_dev(const eolc_cd_plan *plan, const eolc_contact **contacts_dev, int32_t *count) {
    EOLC_REQUIRE(plan && contacts_dev && count, "NULL argument");
    *contacts_dev = plan->last_total > 0 ? plan->d_out.p : nullptr;
    *count = plan->last_total;
    return EOLC_OK;
}

int eolc_cd_run_dev(eolc_cd_plan *plan, const double *x_dev, int32_t n_points, const double *pxyz, const double *pnorms,
                    int32_t n_boxes, const double *box_whd, const double *box_E, int point_eol_flag, int remap_box_indices,
                    eolc_contact *out, int32_t capacity, int32_t *n_out) {
    EOLC_REQUIRE(n_out, "n_out is NULL");
    int32_t off[2] = {0, 0};
    int rc = eolc_cd_run_batched_dev(plan, 1, x_dev, n_points, pxyz, pnorms, n_boxes, box_whd, box_E, point_eol_flag,
                                     remap_box_indices, out, capacity, off);
    *n_out = off[1];
    return rc;
}

int eolc_cd_run(eolc_cd_plan *plan, const double *x, int32_t n_points, const double *pxyz, const double *pnorms,
                int32_t n_boxes, const double *box_whd, const double *box_E, int point_eol_flag, int remap_box_indices,
                eolc_contact *out, int32_t capacity, int32_t *n_out) {
    EOLC_REQUIRE(plan && n_out, "NULL argument");
    EOLC_REQUIRE(plan->N == 0 || x, "x is NULL");
    EOLC_CUDA(cudaSetDevice(plan->ctx->device));
    const size_t n = 3 * (size_t)plan->N;
    EOLC_CUDA(plan->d_x.ensure(std::max<size_t>(n, 1)));
    EOLC_CUDA(plan->p_x.ensure(std::max<size_t>(n, 1)));
    if (n) {
        memcpy(plan->p_x.p, x, n * sizeof(double));
        EOLC_CUDA(cudaMemcpyAsync(plan->d_x.p, plan->p_x.p, n * sizeof(double), cudaMemcpyHostToDevice, plan->ctx->stream));
    }
    return eolc_cd_run_dev(plan, plan->d_x.p, n_points, pxyz, pnorms, n_boxes, box_whd, box_E, point_eol_flag,
                           remap_box_indices, out, capacity, n_out);
}


// ---- Constraints::fill, contact part (src/Constraints.cpp:424-468): one shared routine for the host list and the device kernel
}  // extern "C" (reopened below)
namespace {
__host__ __device__ inline int contact_row(const eolc_contact &c, const unsigned char *node_eol, int32_t *cols, double *vals) {
    int nv;
    if (c.count1 == 3 && c.count2 == 1) nv = 1;
    else if (c.count1 == 2 && c.count2 == 2) nv = 2;
    else if (c.count1 == 1 && c.count2 == 3) nv = 3;
    else nv = 0;
    if (node_eol) for (int j = 0; j < nv; ++j) if (node_eol[c.verts2[j]]) nv = 0;   // `continue` in the reference: no row
    int n = 0;
    for (int j = 0; j < nv; ++j)
        for (int k = 0; k < 3; ++k) {
            cols[n] = c.verts2[j] * 3 + k;
            // (3,1): -nor1(k)   (Constraints.cpp:427-429);   (2,2), (1,3): -nor2(k) * weights2(j)   (:442-444, :458-460)
            vals[n] = nv == 1 ? -c.nor1[k] : -c.nor2[k] * c.weights2[j];
            ++n;
        }
    for (int q = n; q < 9; ++q) { cols[q] = -1; vals[q] = 0.0; }
    return n;
}
__global__ void k_contact_rows(int n, const eolc_contact *__restrict__ c, const unsigned char *__restrict__ node_eol, int32_t *__restrict__ row_nnz,
                               int32_t *__restrict__ cols, double *__restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t cc[9]; double vv[9];
    row_nnz[i] = contact_row(c[i], node_eol, cc, vv);
    for (int q = 0; q < 9; ++q) { cols[9 * (size_t)i + q] = cc[q]; vals[9 * (size_t)i + q] = vv[q]; }
}
// compact (CSR) form of the same rows.  Per contact: 3 nv entries and (nv > 0) rows, packed as entries | rows << 16 inside a block
// (at most 2304 / 256 per 256 contacts); pass 1 stores the two sums of every block (bs[blk], bs[nblk + blk]), ONE scan of the 2 nblk
// sums gives the block offsets (rows relative to off[nblk] = the entry total), pass 2 adds the in-block prefix and writes.
__device__ __forceinline__ int rows_nv(const eolc_contact &c, const unsigned char *__restrict__ node_eol) {
    const int c1 = c.count1, c2 = c.count2;
    int nv = (c1 == 3 && c2 == 1) ? 1 : (c1 == 2 && c2 == 2) ? 2 : (c1 == 1 && c2 == 3) ? 3 : 0;
    if (node_eol) for (int j = 0; j < nv; ++j) if (node_eol[c.verts2[j]]) nv = 0;
    return nv;
}
__global__ void __launch_bounds__(256) k_rows_count(int n, int nblk, const eolc_contact *__restrict__ c, const unsigned char *__restrict__ node_eol,
                                                    int32_t *__restrict__ bs) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int nv = i < n ? rows_nv(c[i], node_eol) : 0;
    int cnt = 3 * nv | (nv > 0 ? 1 << 16 : 0);
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s[k];
        bs[blockIdx.x] = t & 0xffff; bs[nblk + blockIdx.x] = t >> 16;
    }
}
__global__ void __launch_bounds__(256) k_rows_write(int n, int nblk, const eolc_contact *__restrict__ c, const unsigned char *__restrict__ node_eol,
                                                    const int32_t *__restrict__ off, int32_t *__restrict__ row_ptr, int32_t *__restrict__ cols,
                                                    double *__restrict__ vals) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    int32_t cc[9]; double vv[9];
    const int m = i < n ? contact_row(c[i], node_eol, cc, vv) : 0;
    const int pre = block_excl_prefix(m | (m > 0 ? 1 << 16 : 0));
    const int32_t at = off[blockIdx.x] + (pre & 0xffff), row = off[nblk + blockIdx.x] - off[nblk] + (pre >> 16);
    if (m) {
        row_ptr[row] = at;
        for (int q = 0; q < m; ++q) { cols[at + q] = cc[q]; vals[at + q] = vv[q]; }
    }
    if (i == n - 1) row_ptr[off[2 * nblk] - off[nblk]] = off[nblk];
}
}  // namespace
extern "C" {

int eolc_cd_last_count(const eolc_cd_plan *plan) { return plan ? plan->last_total : 0; }

int eolc_constraints_contact_rows(const eolc_contact *contacts, int32_t n, const uint8_t *node_eol, int32_t *n_rows, int32_t *row_nnz,
                                  int32_t *cols, double *vals) {
    EOLC_REQUIRE(n_rows && n >= 0, "bad arguments");
    EOLC_REQUIRE(n == 0 || (contacts && row_nnz && cols && vals), "NULL argument");
    int32_t r = 0;
    for (int32_t i = 0; i < n; ++i) {
        const int m = contact_row(contacts[i], node_eol, cols + 9 * (size_t)r, vals + 9 * (size_t)r);
        if (m == 0) continue;                  // skipped contact: ineqsize does not advance
        row_nnz[r++] = m;
    }
    *n_rows = r;
    return EOLC_OK;
}

int eolc_constraints_fixed_rows(const double *c, const int32_t *ci, const double *v, int32_t n_nodes, int32_t eq_row0, int32_t *n_rows,
                                int32_t *rows, int32_t *cols, double *vals, double *beq) {
    EOLC_REQUIRE(c && ci && n_rows && rows && cols && vals && beq && n_nodes >= 0, "bad arguments");
    int32_t n = 0;
    for (int k = 0; k < 4; ++k) {
        const double *ck = c + 6 * k;
        if (ck[0] == -1.0) continue;                                   // `if (fs->c1(0) != -1)`, Constraints.cpp:470
        EOLC_REQUIRE(ci[k] >= 0 && ci[k] < n_nodes && v, "fixed corner names a node the mesh does not have");
        for (int j = 0; j < 3; ++j) {
            if (ck[j] != 1.0) continue;                                // `if (fs->c1(j) == 1.0) addFixed(...)`
            rows[n] = eq_row0 + n;
            cols[n] = ci[k] * 3 + j;
            vals[n] = ck[j];                                           // Aeq_.push_back(T(eqsize, ci, c(i)))          :116
            beq[n] = (1 - 0.01) * v[3 * (size_t)ci[k] + j] + ck[j + 3]; // (1 - 0.01) * v + c(i + 3)                    :117
            ++n;
        }
    }
    *n_rows = n;
    return EOLC_OK;
}

}  // extern "C" (reopened below)
// device -> caller array: straight DMA when the caller's array is page-locked, else through the plan's pinned staging (one memcpy
// after the stream has been synchronised: `Pending` remembers it)
namespace {
struct Pending { void *dst; const void *src; size_t bytes; };
template <typename T>
int fetch(T *dst, const T *src_dev, size_t count, PinnedBuf<T> &stage, size_t stage_off, size_t stage_total, cudaStream_t st,
          std::vector<Pending> &later) {
    if (!count) return EOLC_OK;
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    T *to = dst;
    if (!pinned) {
        EOLC_CUDA(stage.ensure(stage_total));   // sized for every array that shares it, so an earlier slice is never moved
        to = stage.p + stage_off;
        later.push_back({dst, to, count * sizeof(T)});
    }
    EOLC_CUDA(cudaMemcpyAsync(to, src_dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    return EOLC_OK;
}
}  // namespace
extern "C" {

int eolc_cd_contact_rows(eolc_cd_plan *plan, const uint8_t *node_eol, int32_t capacity_rows, int32_t *n_rows, int32_t *row_nnz, int32_t *cols,
                         double *vals) {
    EOLC_REQUIRE(plan && n_rows, "NULL argument");
    eolc_cd_plan *P = plan;
    const int n = P->last_total;
    *n_rows = 0;
    if (n == 0) return EOLC_OK;
    EOLC_REQUIRE(row_nnz && cols && vals, "NULL argument");
    if (n > capacity_rows) { *n_rows = n; set_error("row buffers too small: need %d rows, capacity %d", n, capacity_rows); return EOLC_ERR_CAPACITY; }
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    EOLC_CUDA(P->d_rows_i.ensure(10 * (size_t)n)); EOLC_CUDA(P->d_rows_v.ensure(9 * (size_t)n));
    const unsigned char *eol_dev = nullptr;
    if (node_eol) {
        EOLC_CUDA(P->d_eol.ensure((size_t)P->N));
        EOLC_CUDA(cudaMemcpyAsync(P->d_eol.p, node_eol, (size_t)P->N, cudaMemcpyHostToDevice, st));
        eol_dev = P->d_eol.p;
    }
    k_contact_rows<<<(n + 255) / 256, 256, 0, st>>>(n, P->d_out.p, eol_dev, P->d_rows_i.p, P->d_rows_i.p + n, P->d_rows_v.p);
    EOLC_CUDA(cudaGetLastError());
    if (!node_eol) {
        // no contact can be skipped: the device arrays ARE the result; they go straight into the caller's arrays
        std::vector<Pending> later;
        int rc = fetch(row_nnz, P->d_rows_i.p, (size_t)n, P->p_rows_i, 0, 10 * (size_t)n, st, later);
        if (!rc) rc = fetch(cols, P->d_rows_i.p + n, 9 * (size_t)n, P->p_rows_i, (size_t)n, 10 * (size_t)n, st, later);
        if (!rc) rc = fetch(vals, P->d_rows_v.p, 9 * (size_t)n, P->p_rows_v, 0, 9 * (size_t)n, st, later);
        if (rc) return rc;
        EOLC_CUDA(cudaStreamSynchronize(st));
        for (const Pending &c : later) memcpy(c.dst, c.src, c.bytes);
        *n_rows = n;
        return EOLC_OK;
    }
    EOLC_CUDA(P->p_rows_i.ensure(10 * (size_t)n)); EOLC_CUDA(P->p_rows_v.ensure(9 * (size_t)n));
    EOLC_CUDA(cudaMemcpyAsync(P->p_rows_i.p, P->d_rows_i.p, 10 * (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaMemcpyAsync(P->p_rows_v.p, P->d_rows_v.p, 9 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    // compaction of the skipped contacts (EoL nodes) on the host
    int32_t r = 0;
    const int32_t *hn = P->p_rows_i.p, *hc = P->p_rows_i.p + n;
    for (int i = 0; i < n; ++i) {
        if (hn[i] == 0) continue;
        row_nnz[r] = hn[i];
        memcpy(cols + 9 * (size_t)r, hc + 9 * (size_t)i, 9 * sizeof(int32_t));
        memcpy(vals + 9 * (size_t)r, P->p_rows_v.p + 9 * (size_t)i, 9 * sizeof(double));
        ++r;
    }
    *n_rows = r;
    return EOLC_OK;
}

int eolc_cd_contact_rows_csr(eolc_cd_plan *plan, const uint8_t *node_eol, int32_t capacity_rows, int32_t capacity_nnz, int32_t *n_rows,
                             int32_t *nnz, int32_t *row_ptr, int32_t *cols, double *vals) {
    EOLC_REQUIRE(plan && n_rows && nnz, "NULL argument");
    eolc_cd_plan *P = plan;
    const int n = P->last_total;
    *n_rows = 0; *nnz = 0;
    if (capacity_rows >= 0 && row_ptr) row_ptr[0] = 0;
    if (n == 0) return EOLC_OK;
    EOLC_REQUIRE((int64_t)n * 9 < INT32_MAX, "too many contacts for int32 row pointers");
    EOLC_CUDA(cudaSetDevice(P->ctx->device));
    cudaStream_t st = P->ctx->stream;
    // device: per-block sums -> one exclusive scan -> compact write (3 launches)
    const int nblk = (n + 255) / 256;
    EOLC_CUDA(P->d_rows_i.ensure(4 * (size_t)nblk + 2 + (size_t)n + 1 + 9 * (size_t)n)); EOLC_CUDA(P->d_rows_v.ensure(9 * (size_t)n));
    int32_t *bs = P->d_rows_i.p, *off = bs + 2 * nblk, *rp = off + (2 * nblk + 1), *cl = rp + (n + 1);   // cl: 9 n
    const unsigned char *eol_dev = nullptr;
    if (node_eol) {
        EOLC_CUDA(P->d_eol.ensure((size_t)P->N));
        EOLC_CUDA(cudaMemcpyAsync(P->d_eol.p, node_eol, (size_t)P->N, cudaMemcpyHostToDevice, st));
        eol_dev = P->d_eol.p;
    }
    k_rows_count<<<nblk, 256, 0, st>>>(n, nblk, P->d_out.p, eol_dev, bs);
    k_scan<<<1, 1024, 0, st>>>((size_t)2 * nblk, bs, off);
    k_rows_write<<<nblk, 256, 0, st>>>(n, nblk, P->d_out.p, eol_dev, off, rp, cl, P->d_rows_v.p);
    EOLC_CUDA(cudaGetLastError());
    EOLC_CUDA(P->p_counter.ensure(2));
    EOLC_CUDA(cudaMemcpyAsync(P->p_counter.p, off + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaMemcpyAsync(P->p_counter.p + 1, off + 2 * nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    EOLC_CUDA(cudaStreamSynchronize(st));
    const int32_t total = P->p_counter.p[0], rows = P->p_counter.p[1] - total;
    *n_rows = rows; *nnz = total;
    if (rows > capacity_rows || total > capacity_nnz) {
        set_error("row buffers too small: need %d rows / %d entries, capacity %d / %d", rows, total, capacity_rows, capacity_nnz);
        return EOLC_ERR_CAPACITY;
    }
    EOLC_REQUIRE(row_ptr && (total == 0 || (cols && vals)), "NULL argument");
    std::vector<Pending> later;
    const size_t stage_i = (size_t)rows + 1 + (size_t)total;
    int rc = fetch(row_ptr, rp, (size_t)rows + 1, P->p_rows_i, 0, stage_i, st, later);
    if (!rc) rc = fetch(cols, cl, (size_t)total, P->p_rows_i, (size_t)rows + 1, stage_i, st, later);
    if (!rc) rc = fetch(vals, P->d_rows_v.p, (size_t)total, P->p_rows_v, 0, (size_t)total, st, later);
    if (rc) return rc;
    EOLC_CUDA(cudaStreamSynchronize(st));
    for (const Pending &c : later) memcpy(c.dst, c.src, c.bytes);
    return EOLC_OK;
}

}  // extern "C"
