// tile_exec.cuh — the two per-thread phases of the "tiles" assembly of Forces::fill, written once for the device
// (forces.cu: assemble_tiles_kernel) and for the host emulation the CPU tests run (tests/hostmath/hostmath.cpp).
//
//   phase 1   each producer thread evaluates one bending stencil and one face of the tile (more if the tile has more than 128)
//             (elements.cuh: edge_element_tile / face_element_tile) and parks the element blocks in `scr`
//   phase 2   every output block of the tile's nodes pulls its contributions from `scr` in the plan's fixed order into staged
//             rows (D: f, O: off-diagonal MDK block, M: mass block)
//   phase 3   diagonal MDK blocks from the row sums of the staged rows
//   copy-out  runs of staged rows leave for their CSR slots with one bulk copy each
// Data layouts: forces_plan.h.  Replaces Forces.cpp:333-397,498-518 (faces), :687-744,885-908 (edges), :925-929 (assembly).
#pragma once
#include <stdint.h>
#include "elements.cuh"
#include "forces_plan.h"

#ifndef EOLC_P2_UNROLL
#define EOLC_P2_UNROLL 1
#endif
#define EOLC_PRAGMA_(x) _Pragma(#x)
#define EOLC_UNROLL_P2_(n) EOLC_PRAGMA_(unroll n)
#define EOLC_UNROLL_P2 EOLC_UNROLL_P2_(EOLC_P2_UNROLL)

namespace eolc {
namespace tiles {

struct FillParams { double mu, lam, rho, beta, gx, gy, gz, dhh; };

struct TileView {
    const uint32_t *geo;    // geometry blob: header, local -> global node table, copy-out runs
    const uint32_t *tmpl;   // template part A: header, items
    const uint32_t *tmplB;  // template part B: header, row sizes, staging offsets, groups, records, pull entries
    const double *xs;       // staged Node::x of the tile's local nodes, 3 per node
    const double *Xs;       // staged material coordinates, 2 per node
    double *scr;            // parked element blocks
    double *kst;            // staged MDK rows of the tile's nodes, laid out like the global value array (same 16-byte phase)
    double *mst;            // staged M rows, same
    double *fst;            // staged f, same
};

EOLC_HD void st2(double *p, double a, double b) {
#ifdef __CUDA_ARCH__
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
EOLC_HD void ld2(const double *p, double &a, double &b) {
#ifdef __CUDA_ARCH__
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
// output values are written once and never re-read by this kernel: streaming stores
EOLC_HD void stg(double *p, double v) {
#ifdef __CUDA_ARCH__
    __stcs(p, v);
#else
    *p = v;
#endif
}

struct ParkEdge {   // slot layout: off-diagonal block k (K_01, K_02, K_03, K_12, K_13, K_23) at 10 k (8 + 1, 1 pad).  The diagonal
    double *s;      // blocks are not parked: phase 3 rebuilds a node's diagonal block from the row sums (forces_plan.h)
    EOLC_HD void diag(int, const sym3 &) {}
    EOLC_HD void off(int k, const blk3 &B) {
        double *d = s + 10 * k;
        st2(d, B.m[0], B.m[1]); st2(d + 2, B.m[2], B.m[3]); st2(d + 4, B.m[4], B.m[5]); st2(d + 6, B.m[6], B.m[7]); d[8] = B.m[8];
    }
};
struct ParkFace {   // off-diagonal block k (K_ab, K_ac, K_bc) at 10 k; force of vertex v at 30 + 4 v; t8 at 33
    double *s;
    EOLC_HD void diag(int, const sym3 &) {}
    EOLC_HD void off(int k, const blk3 &B) {
        double *d = s + 10 * k;
        st2(d, B.m[0], B.m[1]); st2(d + 2, B.m[2], B.m[3]); st2(d + 4, B.m[4], B.m[5]); st2(d + 6, B.m[6], B.m[7]); d[8] = B.m[8];
    }
    EOLC_HD void force(int i, v3 v) { st2(s + 30 + 4 * i, v.x, v.y); s[30 + 4 * i + 2] = v.z; }
    EOLC_HD void mass(double m) { s[FACE_T8] = m; }
};

EOLC_HD v3 ldx(const double *xs, uint32_t l) { return mk3(xs[3 * l], xs[3 * l + 1], xs[3 * l + 2]); }

// Phase 1: thread tid of nthreads evaluates bending stencils tid, tid + nthreads, ... and then faces nthreads - 1 - tid, ...
// (a regular tile has at most one of each per thread) and parks the off-diagonal element blocks in V.scr.
EOLC_HD void phase1(int tid, int nthreads, const TileView &V, const FillParams &P) {
    const uint32_t w0 = V.tmpl[0];
    const int nE = (int)(w0 & 0xffffu), nF = (int)(w0 >> 16);
    for (int e = tid; e < nE; e += nthreads) {
        const uint32_t it = V.tmpl[4 + e];
        const uint32_t l0 = it & 255u, l1 = (it >> 8) & 255u, l2 = (it >> 16) & 255u, l3 = it >> 24;
        double X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y;
        ld2(V.Xs + 2 * l0, X0x, X0y); ld2(V.Xs + 2 * l1, X1x, X1y); ld2(V.Xs + 2 * l2, X2x, X2y); ld2(V.Xs + 2 * l3, X3x, X3y);
        ParkEdge park{V.scr + ZPAD + EDGE_STRIDE * e};
        edge_element_tile(ldx(V.xs, l0), ldx(V.xs, l1), ldx(V.xs, l2), ldx(V.xs, l3), X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y, P.beta, P.dhh, park);
    }
    for (int f = nthreads - 1 - tid; f < nF; f += nthreads) {   // faces from the last thread down: other warps than the stencils'
        const uint32_t it = V.tmpl[4 + nE + f];
        const uint32_t l0 = it & 255u, l1 = (it >> 8) & 255u, l2 = (it >> 16) & 255u;
        double Xax, Xay, Xbx, Xby, Xcx, Xcy;
        ld2(V.Xs + 2 * l0, Xax, Xay); ld2(V.Xs + 2 * l1, Xbx, Xby); ld2(V.Xs + 2 * l2, Xcx, Xcy);
        ParkFace park{V.scr + ZPAD + EDGE_STRIDE * nE + FACE_STRIDE * f};
        face_element_tile(ldx(V.xs, l0), ldx(V.xs, l1), ldx(V.xs, l2), Xax, Xay, Xbx, Xby, Xcx, Xcy, P.mu, P.lam, P.rho, mk3(P.gx, P.gy, P.gz),
                          P.dhh, park);
    }
}

// Unpacks two 16-bit pull offsets (in doubles into scr; 0 = the zero block).
EOLC_HD void pull2(uint32_t e, const double *scr, const double *&s0, const double *&s1) {
    s0 = scr + (e & 0xffffu);
    s1 = scr + (e >> 16);
}

// Phase 2: the tile's records, in groups of 32 of one kind (forces_plan.h); every group names the warp that runs it.
// Every lane of a group runs the same trip counts; the sums land in the staging rows (shared memory), laid out like the
// global rows.  m_full: the M staging rows do not hold this template's explicit zeros yet (first tile / template or parity
// changed), so mass records write whole 3x3 blocks; otherwise only the three diagonal entries change.
// skip_m: the caller declared M unchanged since the last fill (EOLC_FILL_M_UNCHANGED): only the diagonal mass records run (phase 3
// needs M_aa for the diagonal MDK block); off-diagonal mass groups are skipped and the M rows are not copied out.
EOLC_HD void phase2(int tid, int nthreads, const TileView &V, bool m_full, bool skip_m = false) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const uint32_t h0 = V.tmplB[0];
    const int nOwn = (int)(h0 & 255u), nG = (int)((h0 >> 8) & 255u), n4 = (nOwn + 3) & ~3;
    const uint32_t *grp = V.tmplB + B_HDR + 3 * n4;      // after the per-node tables (host side / phase-3 item construction only)
    const unsigned long long *recs = reinterpret_cast<const unsigned long long *>(grp + 4 * nG);
    const uint32_t *pulls = reinterpret_cast<const uint32_t *>(recs + (size_t)GROUP * nG);
    const double *scr = V.scr;
    {
        const uint32_t wr = V.tmplB[4 + warp % nwarps];           // this warp's groups (forces_plan.h balances them)
        for (int g = (int)(wr & 0xffffu), gend = g + (int)(wr >> 16); g < gend; ++g) {
            const uint32_t gw = grp[4 * g];
            const int kind = (int)(gw & 255u), nA = (int)((gw >> 8) & 255u), nB = (int)((gw >> 16) & 255u);
            if (skip_m && ((gw >> 24) & 1u)) continue;
            const uint32_t *pl = pulls + grp[4 * g + 1] + lane;
            const unsigned long long rec = recs[(size_t)GROUP * g + lane];
            const uint32_t off1 = (uint32_t)rec & 0xffffu, st1 = (uint32_t)(rec >> 16) & 1023u;
            const uint32_t off2 = (uint32_t)(rec >> 26) & 0xffffu, st2 = (uint32_t)(rec >> 42) & 1023u;
            const bool valid = (rec >> 54) & 1ull, has2 = (rec >> 52) & 1ull;
            if (kind == KIND_D) {
                // ---- f of the node: its faces in ascending order (f.setZero() then +=, Forces.cpp:915,500-502)
                double f0 = 0.0, f1 = 0.0, f2 = 0.0;
                EOLC_UNROLL_P2
                for (int k = 0; k < nA; ++k, pl += GROUP) {
                    if (!valid) continue;
                    const double *s0, *s1;
                    pull2(*pl, scr, s0, s1);
                    double h, j, h1, j1;
                    ld2(s0, h, j); ld2(s1, h1, j1);
                    const double q = s0[2], q1 = s1[2];
                    f0 += h; f1 += j; f2 += q;
                    f0 += h1; f1 += j1; f2 += q1;
                }
                if (valid) {
                    double *fo = V.fst + off1;
                    fo[0] = f0; fo[1] = f1; fo[2] = f2;
                    // clear the node's diagonal block: phase 3 then sums whole rows (no stale value of an earlier tile enters the sum)
                    if (has2) {
                        double *row = V.kst + off2;
                        row[0] = 0.0; row[1] = 0.0; row[2] = 0.0;
                        row += st2;
                        row[0] = 0.0; row[1] = 0.0; row[2] = 0.0;
                        row += st2;
                        row[0] = 0.0; row[1] = 0.0; row[2] = 0.0;
                    }
                }
            } else if (kind == KIND_O) {
                // ---- off-diagonal MDK block: contributions are parked as K_(lo,hi) of the element; the ones whose row vertex comes
                //      after the column vertex are added transposed (the transposition is just the choice of accumulator)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0, a8 = 0.0;
                EOLC_UNROLL_P2
                for (int k = 0; k < nA; ++k, pl += GROUP) {
                    if (!valid) continue;
                    const double *s0, *s1;
                    pull2(*pl, scr, s0, s1);
                    double b0, b1, b2, b3, b4, b5, b6, b7, d0, d1, d2, d3, d4, d5, d6, d7;
                    ld2(s0, b0, b1); ld2(s0 + 2, b2, b3); ld2(s0 + 4, b4, b5); ld2(s0 + 6, b6, b7);
                    ld2(s1, d0, d1); ld2(s1 + 2, d2, d3); ld2(s1 + 4, d4, d5); ld2(s1 + 6, d6, d7);
                    const double b8 = s0[8], d8 = s1[8];
                    a0 += b0; a1 += b1; a2 += b2; a3 += b3; a4 += b4; a5 += b5; a6 += b6; a7 += b7; a8 += b8;
                    a0 += d0; a1 += d1; a2 += d2; a3 += d3; a4 += d4; a5 += d5; a6 += d6; a7 += d7; a8 += d8;
                }
                EOLC_UNROLL_P2
                for (int k = 0; k < nB; ++k, pl += GROUP) {
                    if (!valid) continue;
                    const double *s0, *s1;
                    pull2(*pl, scr, s0, s1);
                    double b0, b1, b2, b3, b4, b5, b6, b7, d0, d1, d2, d3, d4, d5, d6, d7;
                    ld2(s0, b0, b1); ld2(s0 + 2, b2, b3); ld2(s0 + 4, b4, b5); ld2(s0 + 6, b6, b7);
                    ld2(s1, d0, d1); ld2(s1 + 2, d2, d3); ld2(s1 + 4, d4, d5); ld2(s1 + 6, d6, d7);
                    const double b8 = s0[8], d8 = s1[8];
                    a0 += b0; a3 += b1; a6 += b2; a1 += b3; a4 += b4; a7 += b5; a2 += b6; a5 += b7; a8 += b8;
                    a0 += d0; a3 += d1; a6 += d2; a1 += d3; a4 += d4; a7 += d5; a2 += d6; a5 += d7; a8 += d8;
                }
                if (valid) {
                    double *row = V.kst + off1;
                    row[0] = a0; row[1] = a1; row[2] = a2;
                    row += st1;
                    row[0] = a3; row[1] = a4; row[2] = a5;
                    row += st1;
                    row[0] = a6; row[1] = a7; row[2] = a8;
                    if (has2) {   // the column node is owned too: its row gets the transposed block
                        double *r2 = V.kst + off2;
                        r2[0] = a0; r2[1] = a3; r2[2] = a6;
                        r2 += st2;
                        r2[0] = a1; r2[1] = a4; r2[2] = a7;
                        r2 += st2;
                        r2[0] = a2; r2[1] = a5; r2[2] = a8;
                    }
                }
            } else {
                // ---- mass block: sum of rho 2A over the faces shared by the pair; /12 on the diagonal block, /24 off it
                //      (ComputeInertial.cpp:33,44-47)
                double m = 0.0;
                for (int k = 0; k < nA; ++k, pl += GROUP) {
                    if (!valid) continue;
                    const double *s0, *s1;
                    pull2(*pl, scr, s0, s1);
                    m += *s0;
                    m += *s1;
                }
                if (valid && !(skip_m && !((rec >> 53) & 1ull))) {
                    m *= ((rec >> 53) & 1ull) ? (1.0 / 12.0) : (1.0 / 24.0);
                    for (int side = 0; side < 2; ++side) {
                        if (side && !has2) break;
                        double *row = V.mst + (side ? off2 : off1);
                        const uint32_t st = side ? st2 : st1;
                        if (m_full) {
                            row[0] = m; row[1] = 0.0; row[2] = 0.0;
                            row += st;
                            row[0] = 0.0; row[1] = m; row[2] = 0.0;
                            row += st;
                            row[0] = 0.0; row[1] = 0.0; row[2] = m;
                        } else {
                            row[0] = m; row[st + 1] = m; row[2 * st + 2] = m;
                        }
                    }
                }
            }
        }
    }
}

// Phase 3 (after a barrier: all off-diagonal blocks and mass blocks of the tile's rows are staged): the diagonal MDK block of every
// owned node.  Each element matrix has zero row sums apart from the mass part (t8/12 + 2 t8/24 = t8/6 per row), so
//     MDK_aa = 2 M_aa I - sum over b != a of MDK_ab           (M_aa = sum of t8/12 over the node's faces)
// Work item t handles entry (j, k), j <= k, of node t / 6 and mirrors it: exactly symmetric like the reference's diagonal blocks.
EOLC_HD void phase3(int tid, int nthreads, const TileView &V) {
    const uint32_t *items = V.tmplB + V.tmplB[1];
    const int n_items = (int)V.tmplB[2];
    for (int t = tid; t < n_items; t += nthreads) {
        const uint32_t w0 = items[3 * t], w1 = items[3 * t + 1];
        const int deg = (int)((w0 >> 16) & 255u);
        const double *r = V.kst + (w0 & 0xffffu);
        // entry (j, k) of the sum of the node's blocks (the diagonal block was cleared in phase 2): the first 16 blocks are loaded
        // together and summed as a fixed tree (one round trip to shared memory instead of one per block), the rest in order
        double v16[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v16[q] = q < deg ? r[3 * q] : 0.0;
        double s0 = ((v16[0] + v16[1]) + (v16[2] + v16[3])) + ((v16[4] + v16[5]) + (v16[6] + v16[7]));
        double s1 = ((v16[8] + v16[9]) + (v16[10] + v16[11])) + ((v16[12] + v16[13]) + (v16[14] + v16[15]));
        for (int p = 16; p < deg; ++p) s1 += r[3 * p];
        const double v = ((w0 >> 24) & 1u ? 2.0 * V.mst[items[3 * t + 2]] : 0.0) - (s0 + s1);
        V.kst[w1 & 0xffffu] = v;
        V.kst[w1 >> 16] = v;
    }
}

// Copy-out: every run of staged rows (rows of owned nodes with consecutive ids) goes to its CSR slot with ONE bulk copy
// (device: cp.async.bulk shared -> global, asynchronous; host emulation: memcpy).  Staged and global rows share their
// 16-byte phase (forces_plan.h + the parity of the output pointers added to V.kst / V.mst / V.fst by the caller), so only a
// misaligned first / last double is stored separately.  The mass rows were staged expanded, (m, 0, 0; 0, m, 0; 0, 0, m): the six
// off-axis entries are the reference's explicit zeros (ComputeInertial.cpp:45-46, kept by setFromTriplets).
// Lane `lane` of `nlanes` takes runs lane, lane + nlanes, ...; f, Mv, Kv: outputs of the tile's scene.
// kinds: bit mask of the run kinds to copy now (1: MDK rows, 2: M rows, 4: f) — M rows and f are final after phase 2 already.
template <typename Bulk>
EOLC_HD void copy_out_runs(int lane, int nlanes, const TileView &V, double *__restrict__ f, double *__restrict__ Mv, double *__restrict__ Kv, uint32_t kinds, Bulk bulk) {
    const uint32_t g1 = V.geo[1];
    const uint32_t *runs = V.geo + 4 + ((((g1 >> 8) & 255u) + 3u) & ~3u);
    const int nruns = (int)V.geo[3];
    for (int r = lane; r < nruns; r += nlanes) {
        const uint32_t *c = runs + 4 * r;
        const uint32_t hi = c[1], kind = hi >> 30;
        if (!((kinds >> kind) & 1u)) continue;
        double *dst = (kind == 0 ? Kv : kind == 1 ? Mv : f) + (((unsigned long long)(hi & 0x3fffffffu) << 32) | c[0]);
        const double *src = (kind == 0 ? V.kst : kind == 1 ? V.mst : V.fst) + c[2];
        uint32_t n = c[3];
        if ((uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 3) & 1u) { stg(dst, *src); ++dst; ++src; --n; }
        const uint32_t nb = n & ~1u;
        if (nb) bulk(dst, src, nb * 8u);
        if (n & 1u) stg(dst + nb, src[nb]);
    }
}

}  // namespace tiles
}  // namespace eolc
