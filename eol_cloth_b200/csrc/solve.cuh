// solve.cuh — the consumer of Forces::fill on the device, so that M / MDK need not cross PCIe:
//   Cloth::solve right-hand side      b = -(M v + h f)                         /root/reference/src/Cloth.cpp:345
//   GeneralizedSolver::velocitySolve, collision-free branch without fixed points: ConjugateGradient<SparseMatrix<double>, Lower|Upper>
//       cg.compute(MDK); v = cg.solve(-b)                                     /root/reference/src/GeneralizedSolver.cpp:120-126
//   (that branch is dead code in the reference, `if (!collisions && false)`, but it is the only solver a build without Mosek / Gurobi
//   could run: BASELINE configs[0].)  The iteration restates Eigen 3.3's conjugate_gradient() (IterativeLinearSolvers/
//   ConjugateGradient.h — external, not under /root/reference) with its default DiagonalPreconditioner, started from x = 0.
//
// Kernels work on the block structure of the plan's pattern (node a owns 3 rows; row j of node a is the contiguous segment
// vals[9 blkptr[a] + 3 deg j ...] of 3 deg doubles, column block p belongs to node nbr[blkptr[a] + p]): 4 index bytes per 72 value
// bytes, coalesced value reads.  All reductions run in a fixed order (warp tree, block tree, one-block final pass): bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eolc {
namespace solve {

constexpr int WARPS = 8, THREADS = 32 * WARPS;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Rows of a node times x, a HALF warp per node (a structured cloth node has 13 column blocks in MDK, 7 in M): lane hl of the half
// takes column blocks hl, hl + 16, ... — one index load, the block's three x values, its 3 x 3 values (three runs of 3 consecutive
// doubles, neighbouring lanes on neighbouring runs) and 9 FMAs — then a fixed 16-lane tree.  y[0..2] valid in every lane of the half.
// LPN lanes per node: 16 in general; 8 when no row has more than 8 column blocks (the mass matrix of a triangle mesh of valence <= 7):
// twice the nodes per warp, and bit for bit the same sums (with at most one block per lane the 16-lane tree only adds zeros on top).
template <int LPN = 16>
__device__ __forceinline__ void node_rows_times(int hl, bool active, const int32_t *__restrict__ blkptr, const int32_t *__restrict__ nbr,
                                                const double *__restrict__ vals, const double *__restrict__ x, int a, double y[3]) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (active) {
        const int b0 = blkptr[a], deg = blkptr[a + 1] - b0;
        const unsigned n3 = 3u * (unsigned)deg;
        const double *row = vals + 9 * (size_t)b0;
        for (int p = hl; p < deg; p += LPN) {
            const double *xp = x + 3 * (size_t)nbr[b0 + p];
            const double x0 = xp[0], x1 = xp[1], x2 = xp[2];
            const double *r = row + 3u * (unsigned)p;
            s0 += r[0] * x0 + r[1] * x1 + r[2] * x2;
            s1 += r[n3] * x0 + r[n3 + 1] * x1 + r[n3 + 2] * x2;
            s2 += r[2 * n3] * x0 + r[2 * n3 + 1] * x1 + r[2 * n3 + 2] * x2;
        }
    }
#pragma unroll
    for (int o = LPN / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    y[0] = s0; y[1] = s1; y[2] = s2;
}

// fixed-order block reduction of one value per warp; result valid in thread 0
__device__ __forceinline__ double block_sum(double warp_value, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = warp_value;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < WARPS; ++w) t += sh[w];
    __syncthreads();
    return t;
}

// b = -(M v + h f)
template <int LPN>
__global__ void __launch_bounds__(THREADS) k_rhs(int N, const int32_t *__restrict__ blkptr, const int32_t *__restrict__ nbr,
                                                 const double *__restrict__ Mv, const double *__restrict__ f, const double *__restrict__ v,
                                                 double h, double *__restrict__ b) {
    constexpr int NPW = 32 / LPN;                       // nodes per warp
    const int hl = threadIdx.x & (LPN - 1);
    for (int a0 = NPW * (blockIdx.x * WARPS + (threadIdx.x >> 5)); a0 < N; a0 += NPW * gridDim.x * WARPS) {
        const int a = a0 + ((threadIdx.x & 31) / LPN);
        double y[3];
        node_rows_times<LPN>(hl, a < N, blkptr, nbr, Mv, v, a, y);
        if (hl < 3 && a < N) b[3 * (size_t)a + hl] = -((hl == 0 ? y[0] : hl == 1 ? y[1] : y[2]) + h * f[3 * (size_t)a + hl]);
    }
}

// CG state scalars on the device: [0] absNew = r.z, [1] p.Ap, [2] ||r||^2, [3] alpha, [4] beta, [5] threshold, [6] converged flag (as double)
// init: x = 0, r = rhs = -b, z = r / diag(A), p = z; partials of r.z and rhs.rhs.  A dof with fixed[i] != 0 keeps x = 0: its
// row and column drop out (dinv = 0 -> r = z = p = 0 there, and k_cg_ap zeroes its row), i.e. CG runs on the free-free block.
__global__ void __launch_bounds__(THREADS) k_cg_init(int N, const int32_t *__restrict__ blkptr, const int32_t *__restrict__ nbr,
                                                     const double *__restrict__ Kv, const double *__restrict__ b, const unsigned char *__restrict__ fixed,
                                                     double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, double *__restrict__ dinv,
                                                     double *__restrict__ part) {
    __shared__ double sh[WARPS];
    double rz = 0.0, rr = 0.0;
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < 3 * (size_t)N; i += (size_t)gridDim.x * THREADS) {
        const int a = (int)(i / 3), j = (int)(i - 3 * (size_t)a);
        const int b0 = blkptr[a], deg = blkptr[a + 1] - b0;
        double d = 1.0;
        if (deg > 0) {
            int pd = 0;
            while (pd < deg && nbr[b0 + pd] != a) ++pd;      // position of the diagonal block in the row
            const double akk = pd < deg ? Kv[9 * (size_t)b0 + (size_t)3 * deg * j + 3 * pd + j] : 0.0;
            d = akk != 0.0 ? 1.0 / akk : 1.0;                // Eigen's DiagonalPreconditioner: 1 where the diagonal is zero
        }
        if (fixed && fixed[i]) d = 0.0;                          // prescribed (zero) velocity: the dof drops out of the system
        const double ri = d != 0.0 ? -b[i] : 0.0, zi = d * ri;
        dinv[i] = d; x[i] = 0.0; r[i] = ri; p[i] = zi;
        rz += ri * zi; rr += ri * ri;
    }
    const double t0 = block_sum(warp_sum(rz), sh), t1 = block_sum(warp_sum(rr), sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = t0; part[2 * blockIdx.x + 1] = t1; }
}

// one block: sums `stride` interleaved partial columns in a fixed order and updates the scalars
//   mode 0 (after init):  absNew = sum0, threshold = tol^2 * sum1, ||r||^2 = sum1
//   mode 1 (after Ap):    p.Ap = sum0, alpha = absNew / p.Ap
//   mode 2 (after update): absOld = absNew, absNew = sum0, ||r||^2 = sum1, beta = absNew / absOld, converged = ||r||^2 < threshold
__global__ void __launch_bounds__(THREADS) k_cg_scalars(int nparts, int stride, const double *__restrict__ part, double *__restrict__ sc, int mode, double tol) {
    __shared__ double sh[WARPS];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < nparts; i += THREADS) { s0 += part[(size_t)stride * i]; if (stride > 1) s1 += part[(size_t)stride * i + 1]; }
    const double t0 = block_sum(warp_sum(s0), sh), t1 = block_sum(warp_sum(s1), sh);
    if (threadIdx.x == 0) {
        if (mode == 0) { sc[0] = t0; sc[2] = t1; sc[5] = tol * tol * t1; sc[6] = t1 <= sc[5] || t1 == 0.0 ? 1.0 : 0.0; }
        else if (mode == 1) { sc[1] = t0; sc[3] = sc[0] / t0; }
        else { const double old = sc[0]; sc[0] = t0; sc[2] = t1; sc[4] = t0 / old; if (t1 < sc[5]) sc[6] = 1.0; }
    }
}

// Ap = A p (rows of fixed dofs zeroed: dinv == 0 marks them) and the partials of p.Ap
__global__ void __launch_bounds__(THREADS) k_cg_ap(int N, const int32_t *__restrict__ blkptr, const int32_t *__restrict__ nbr,
                                                   const double *__restrict__ Kv, const double *__restrict__ p, const double *__restrict__ dinv,
                                                   double *__restrict__ Ap, double *__restrict__ part, const double *__restrict__ sc) {
    __shared__ double sh[WARPS];
    if (sc[6] != 0.0) return;                                  // converged: the remaining launches of the batch are no-ops
    const int hl = threadIdx.x & 15;
    double acc = 0.0;
    for (int a0 = 2 * (blockIdx.x * WARPS + (threadIdx.x >> 5)); a0 < N; a0 += 2 * gridDim.x * WARPS) {
        const int a = a0 + ((threadIdx.x >> 4) & 1);
        double y[3];
        node_rows_times(hl, a < N, blkptr, nbr, Kv, p, a, y);
        if (hl < 3 && a < N) {
            const double yi = dinv[3 * (size_t)a + hl] != 0.0 ? (hl == 0 ? y[0] : hl == 1 ? y[1] : y[2]) : 0.0;
            Ap[3 * (size_t)a + hl] = yi;
            acc += p[3 * (size_t)a + hl] * yi;
        }
    }
    const double t = block_sum(warp_sum(acc), sh);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// x += alpha p; r -= alpha Ap; z = dinv r (kept in Ap's storage); partials of r.z and r.r
__global__ void __launch_bounds__(THREADS) k_cg_update(size_t n, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                                       double *__restrict__ Ap_z, const double *__restrict__ dinv, double *__restrict__ part,
                                                       const double *__restrict__ sc) {
    __shared__ double sh[WARPS];
    if (sc[6] != 0.0) return;
    const double alpha = sc[3];
    double rz = 0.0, rr = 0.0;
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * THREADS) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap_z[i], zi = dinv[i] * ri;
        r[i] = ri; Ap_z[i] = zi;
        rz += ri * zi; rr += ri * ri;
    }
    const double t0 = block_sum(warp_sum(rz), sh), t1 = block_sum(warp_sum(rr), sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = t0; part[2 * blockIdx.x + 1] = t1; }
}

// p = z + beta p
__global__ void __launch_bounds__(THREADS) k_cg_dir(size_t n, double *__restrict__ p, const double *__restrict__ z, const double *__restrict__ sc) {
    if (sc[6] != 0.0) return;
    const double beta = sc[4];
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * THREADS) p[i] = z[i] + beta * p[i];
}

// x += h v (Cloth::step, Cloth.cpp:394-400: node->x = node->x + h * node->v for every node)
__global__ void __launch_bounds__(THREADS) k_integrate(size_t n, double *__restrict__ x, const double *__restrict__ v, double h) {
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * THREADS) x[i] = x[i] + h * v[i];
}

// ---- scalar-CSR forms of the same three kernels for plans with Eulerian dofs (EOL meshes, forces_eol.h): the rows of a node coupled to
// EoL nodes carry extra columns and the Eulerian rows have no block structure, so these walk Eigen's outer / inner arrays, a warp per
// scalar row (4 index bytes per value instead of 4 per 72: EOL plans only).  Same fixed-order reductions.
__global__ void __launch_bounds__(THREADS) k_rhs_csr(int dof, const int32_t *__restrict__ outer, const int32_t *__restrict__ inner,
                                                     const double *__restrict__ Mv, const double *__restrict__ f, const double *__restrict__ v,
                                                     double h, double *__restrict__ b) {
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * WARPS + (threadIdx.x >> 5); row < dof; row += gridDim.x * WARPS) {
        double s = 0.0;
        for (int k = outer[row] + lane; k < outer[row + 1]; k += 32) s += Mv[k] * v[inner[k]];
        s = warp_sum(s);
        if (lane == 0) b[row] = -(s + h * f[row]);
    }
}
__global__ void __launch_bounds__(THREADS) k_cg_init_csr(int dof, const int32_t *__restrict__ outer, const int32_t *__restrict__ inner,
                                                         const double *__restrict__ Kv, const double *__restrict__ b, const unsigned char *__restrict__ fixed,
                                                         double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, double *__restrict__ dinv,
                                                         double *__restrict__ part) {
    __shared__ double sh[WARPS];
    double rz = 0.0, rr = 0.0;
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < (size_t)dof; i += (size_t)gridDim.x * THREADS) {
        int lo = outer[i], hi = outer[i + 1];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (inner[mid] < (int)i) lo = mid + 1; else hi = mid; }   // the diagonal entry, if stored
        const double akk = (lo < outer[i + 1] && inner[lo] == (int)i) ? Kv[lo] : 0.0;
        double d = akk != 0.0 ? 1.0 / akk : 1.0;                 // Eigen's DiagonalPreconditioner: 1 where the diagonal is zero
        if (fixed && fixed[i]) d = 0.0;
        const double ri = d != 0.0 ? -b[i] : 0.0, zi = d * ri;
        dinv[i] = d; x[i] = 0.0; r[i] = ri; p[i] = zi;
        rz += ri * zi; rr += ri * ri;
    }
    const double t0 = block_sum(warp_sum(rz), sh), t1 = block_sum(warp_sum(rr), sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = t0; part[2 * blockIdx.x + 1] = t1; }
}
__global__ void __launch_bounds__(THREADS) k_cg_ap_csr(int dof, const int32_t *__restrict__ outer, const int32_t *__restrict__ inner,
                                                       const double *__restrict__ Kv, const double *__restrict__ p, const double *__restrict__ dinv,
                                                       double *__restrict__ Ap, double *__restrict__ part, const double *__restrict__ sc) {
    __shared__ double sh[WARPS];
    if (sc[6] != 0.0) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int row = blockIdx.x * WARPS + (threadIdx.x >> 5); row < dof; row += gridDim.x * WARPS) {
        double s = 0.0;
        for (int k = outer[row] + lane; k < outer[row + 1]; k += 32) s += Kv[k] * p[inner[k]];
        s = warp_sum(s);
        if (lane == 0) {
            const double yi = dinv[row] != 0.0 ? s : 0.0;
            Ap[row] = yi;
            acc += p[row] * yi;
        }
    }
    const double t = block_sum(warp_sum(acc), sh);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
// Eulerian part of the position update (Cloth.cpp:401-407): vert->u += h * vert->v for the EoL nodes, v taken at 3N + 2 EoL_index
__global__ void __launch_bounds__(THREADS) k_integrate_X(int N, const int32_t *__restrict__ eol_index, double *__restrict__ X, const double *__restrict__ v, double h) {
    for (size_t a = blockIdx.x * (size_t)THREADS + threadIdx.x; a < (size_t)N; a += (size_t)gridDim.x * THREADS) {
        const int k = eol_index[a];
        if (k < 0) continue;
        X[2 * a] = X[2 * a] + h * v[3 * (size_t)N + 2 * (size_t)k];
        X[2 * a + 1] = X[2 * a + 1] + h * v[3 * (size_t)N + 2 * (size_t)k + 1];
    }
}

// ---- per-step derived mesh data ------------------------------------------------------------------------------
// face->n of compute_ws_data(Face*), /root/reference/src/external/ArcSim/mesh.cpp:135-140: normalize(cross(x1 - x0, x2 - x0)), with
// ArcSim's normalize (vectors.hpp:111: the zero vector stays zero) and its sequential dot (vectors.hpp:108).  No FMA contraction here
// (explicit _rn intrinsics; sqrt and the division are IEEE) so that the normals match the host arithmetic bit for bit.
// A vector divided by a scalar is u * (1 / a) in ArcSim (vectors.hpp:104; the _mm256_div_pd specialisation at :130 sits behind
// `#if defined(_AVX)`, which no compiler and no line of the reference's CMakeLists.txt defines): one reciprocal, three products —
// NOT three divisions, which differ in the last bit (found by tests/test_forces_ref_pin.py against the reference's own code).
__device__ __forceinline__ void ws_normalize(double &a, double &b, double &c) {
    const double m = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    if (m == 0.0) { a = b = c = 0.0; return; }
    const double inv = __ddiv_rn(1.0, m);
    a = __dmul_rn(inv, a); b = __dmul_rn(inv, b); c = __dmul_rn(inv, c);
}
__device__ __forceinline__ void ws_cross(const double *u, const double *v, double *w) {
    w[0] = __dsub_rn(__dmul_rn(u[1], v[2]), __dmul_rn(u[2], v[1]));
    w[1] = __dsub_rn(__dmul_rn(u[2], v[0]), __dmul_rn(u[0], v[2]));
    w[2] = __dsub_rn(__dmul_rn(u[0], v[1]), __dmul_rn(u[1], v[0]));
}
__global__ void __launch_bounds__(THREADS) k_face_normals(int32_t F, const int32_t *__restrict__ fn, const double *__restrict__ x, double *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < (size_t)F; i += (size_t)gridDim.x * THREADS) {
        const int32_t a = fn[3 * i], b = fn[3 * i + 1], c = fn[3 * i + 2];
        double e1[3], e2[3], n[3];
        for (int k = 0; k < 3; ++k) { e1[k] = __dsub_rn(x[3 * (size_t)b + k], x[3 * (size_t)a + k]); e2[k] = __dsub_rn(x[3 * (size_t)c + k], x[3 * (size_t)a + k]); }
        ws_cross(e1, e2, n);
        ws_normalize(n[0], n[1], n[2]);
        out[3 * i] = n[0]; out[3 * i + 1] = n[1]; out[3 * i + 2] = n[2];
    }
}
// node->n = normal<WS>(node), /root/reference/src/external/ArcSim/geometry.cpp:302-316: over the node's faces (vert->adjf order = the order
// the faces were added, mesh.cpp:372, i.e. ascending face index in the flattened mesh) n += cross(e1, e2) / (2 |e1|^2 |e2|^2) with
// e1, e2 the edges to the next / next-but-one vertex of the face, then normalize.  nfp / nfl: node -> incident faces (CSR), face << 2 | position.
__global__ void __launch_bounds__(THREADS) k_node_normals(int32_t N, const int32_t *__restrict__ nfp, const uint32_t *__restrict__ nfl,
                                                          const int32_t *__restrict__ fn, const double *__restrict__ x, double *__restrict__ out) {
    for (size_t a = blockIdx.x * (size_t)THREADS + threadIdx.x; a < (size_t)N; a += (size_t)gridDim.x * THREADS) {
        double n[3] = {0.0, 0.0, 0.0};
        const double p[3] = {x[3 * a], x[3 * a + 1], x[3 * a + 2]};
        for (int32_t q = nfp[a]; q < nfp[a + 1]; ++q) {
            const uint32_t fl = nfl[q];
            const size_t f = fl >> 2;
            const int j = (int)(fl & 3u), j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const size_t b = (size_t)fn[3 * f + j1], c = (size_t)fn[3 * f + j2];
            double e1[3], e2[3], w[3];
            for (int k = 0; k < 3; ++k) { e1[k] = __dsub_rn(x[3 * b + k], p[k]); e2[k] = __dsub_rn(x[3 * c + k], p[k]); }
            ws_cross(e1, e2, w);
            const double l1 = __dadd_rn(__dadd_rn(__dmul_rn(e1[0], e1[0]), __dmul_rn(e1[1], e1[1])), __dmul_rn(e1[2], e1[2]));
            const double l2 = __dadd_rn(__dadd_rn(__dmul_rn(e2[0], e2[0]), __dmul_rn(e2[1], e2[1])), __dmul_rn(e2[2], e2[2]));
            const double den = __dmul_rn(__dmul_rn(2.0, l1), l2);
            const double inv = __ddiv_rn(1.0, den);                       // cross(e1, e2) / den = cross * (1 / den), vectors.hpp:104
            for (int k = 0; k < 3; ++k) n[k] = __dadd_rn(n[k], __dmul_rn(inv, w[k]));
        }
        ws_normalize(n[0], n[1], n[2]);
        out[3 * a] = n[0]; out[3 * a + 1] = n[1]; out[3 * a + 2] = n[2];
    }
}

}  // namespace solve
}  // namespace eolc
