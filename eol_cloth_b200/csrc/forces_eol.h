// forces_eol.h — the EOL (Eulerian-on-Lagrangian) branch of Forces::fill: plan (host) and per-element / per-entry work
// (host + device, the same source compiles for nvcc and for the unit-test shim, like elements.cuh).
//
// Replaces, for meshes with mesh.EoL_Count > 0:
//   deform_grad                               /root/reference/src/UtilEOL.cpp:13-28
//   fillEOLInertia / fillEOLMembrane          /root/reference/src/Forces.cpp:177-329
//   faceBasedF, EOL scatter                   /root/reference/src/Forces.cpp:399-497  (fill{X,XX,Xx,xX}MI :127-175)
//   fillEOLBending                            /root/reference/src/Forces.cpp:580-683
//   edgeBasedF, EOL scatter                   /root/reference/src/Forces.cpp:746-883  (fill{X,XX,Xx,xX}B :541-578)
//   the EoL rows / columns of setFromTriplets /root/reference/src/Forces.cpp:925-929
//
// What the reference's expansion is: every vertex v of an element that is an EoL node gains two Eulerian dofs X_v at
// 3N + 2 EoL_index, and the element matrix / force are extended by the congruence G^T K G, G^T f with G_v = [I, -F]:
//     (x_u, x_w) = K_uw (unchanged)   (X_v, x_w) = -F^T K_vw   (x_u, X_v) = -K_uv F   (X_v, X_w) = F^T K_vw F   f_X = -F^T f_v
// F = deform_grad(face) for the face terms and the mean of the two adjacent faces' F for a bending stencil.  Two consequences
// shape the design:
//   * the 3N x 3N Lagrangian part of M and MDK is the same as without EoL nodes, so the tiles kernel (forces.cu) still
//     produces all of it, straight into the final arrays — only its row destinations move (rows of nodes that gained Eulerian
//     columns are copied out per scalar row, see forces_plan.h `RowLayout`);
//   * everything else is O(#EoL nodes): the elements touching an EoL node ("EOL elements") are evaluated once more by
//     eol_elements_kernel into a small scratch record each, and eol_gather_kernel sums, per extra matrix / force entry,
//     its contributions in the reference's insertion order (faces ascending, then stencils ascending) — deterministic,
//     no atomics.  The bending FORCE, which the Lagrangian branch discards (Forces.cpp:885-908), enters f for the stencils
//     that hold an EoL node (:750-760): the gather adds it onto the tiles kernel's f.
#pragma once
#include "elements.cuh"
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace eolc {
namespace eol {

struct Params { double e, nu, rho, beta, gx, gy, gz, dhh; };

// ---- scratch record layouts (doubles) --------------------------------------------------------------------------------
// face:    [fX: 3 x 2] [per ordered pair (P, q): KXx 2x3, MXx 2x3] [per ordered pair (P, Q): KXX 2x2, MXX 2x2]
// stencil: [fx: 4 x 3] [fX: 4 x 2] [per ordered pair (P, q): KXx 2x3] [per ordered pair (P, Q): KXX 2x2]
// K = the MDK contribution (Mi + dhh Km for a face, dhh Kb for a stencil); only the parts of EoL vertices P are written / read.
constexpr int FACE_FX = 0, FACE_XX_PAIR = 12, FACE_Xx = 6, FACE_XX = FACE_Xx + 9 * 12, FACE_REC = FACE_XX + 9 * 8;   // 186
constexpr int EDGE_FX_LAG = 0, EDGE_FX = 12, EDGE_Xx = 20, EDGE_XX = EDGE_Xx + 16 * 6, EDGE_REC = EDGE_XX + 16 * 4;   // 180
static_assert(FACE_REC == 186 && EDGE_REC == 180 && FACE_XX_PAIR == 12, "record layout");

// F = Dx DX^-1 (UtilEOL.cpp:13-28): columns F0, F1
EOLC_HD void deform_grad(v3 xa, v3 xb, v3 xc, double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy, v3 &F0, v3 &F1) {
    const v3 d1 = xb - xa, d2 = xc - xa;
    const double D00 = Xbx - Xax, D01 = Xcx - Xax, D10 = Xby - Xay, D11 = Xcy - Xay;
    const double invdet = 1.0 / (D00 * D11 - D10 * D01);
    F0 = (D11 * invdet) * d1 + (-D10 * invdet) * d2;
    F1 = (-D01 * invdet) * d1 + (D00 * invdet) * d2;
}

// out (2x3, row-major) = -F^T B ; B row-major 3x3
EOLC_HD void neg_Ft_B(v3 F0, v3 F1, const double *B, double *out) {
    for (int k = 0; k < 3; ++k) {
        out[k] = -(F0.x * B[k] + F0.y * B[3 + k] + F0.z * B[6 + k]);
        out[3 + k] = -(F1.x * B[k] + F1.y * B[3 + k] + F1.z * B[6 + k]);
    }
}
// out (2x2, row-major) = F^T B F
EOLC_HD void Ft_B_F(v3 F0, v3 F1, const double *B, double *out) {
    double T[6];
    for (int k = 0; k < 3; ++k) {
        T[k] = F0.x * B[k] + F0.y * B[3 + k] + F0.z * B[6 + k];
        T[3 + k] = F1.x * B[k] + F1.y * B[3 + k] + F1.z * B[6 + k];
    }
    out[0] = T[0] * F0.x + T[1] * F0.y + T[2] * F0.z; out[1] = T[0] * F1.x + T[1] * F1.y + T[2] * F1.z;
    out[2] = T[3] * F0.x + T[4] * F0.y + T[5] * F0.z; out[3] = T[3] * F1.x + T[4] * F1.y + T[5] * F1.z;
}
// block (r, c) of a symmetric block matrix stored as its upper blocks: diag d[i], off-diagonal up[pair(r < c)]; transposed when r > c
EOLC_HD void get_block(const blk3 *diag, const blk3 *up, int nv, int r, int c, double *B) {
    if (r == c) { for (int q = 0; q < 9; ++q) B[q] = diag[r].m[q]; return; }
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    const int pair = nv == 3 ? (lo == 0 ? hi - 1 : 2) : (lo == 0 ? hi - 1 : lo == 1 ? hi + 1 : 5);   // 01 02 12 | 01 02 03 12 13 23
    const blk3 &U = up[pair];
    if (r < c) { for (int q = 0; q < 9; ++q) B[q] = U.m[q]; }
    else { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[3 * i + j] = U.m[3 * j + i]; }
}

// One EOL face: rec = (a, b, c, mask of EoL vertices).  Forces.cpp:399-497 with fillEOLInertia / fillEOLMembrane.
EOLC_HD void face_record(const int32_t *rec, const double *x, const double *X, const Params &p, double *out) {
    const int32_t n[3] = {rec[0], rec[1], rec[2]};
    const int mask = rec[3];
    v3 xs[3];
    for (int v = 0; v < 3; ++v) xs[v] = mk3(x[3 * (size_t)n[v]], x[3 * (size_t)n[v] + 1], x[3 * (size_t)n[v] + 2]);
    const double Xax = X[2 * (size_t)n[0]], Xay = X[2 * (size_t)n[0] + 1], Xbx = X[2 * (size_t)n[1]], Xby = X[2 * (size_t)n[1] + 1],
                 Xcx = X[2 * (size_t)n[2]], Xcy = X[2 * (size_t)n[2] + 1];
    FaceOut o;
    face_element(xs[0], xs[1], xs[2], Xax, Xay, Xbx, Xby, Xcx, Xcy, p.e, p.nu, p.rho, mk3(p.gx, p.gy, p.gz), p.dhh, o);
    v3 F0, F1;
    deform_grad(xs[0], xs[1], xs[2], Xax, Xay, Xbx, Xby, Xcx, Xcy, F0, F1);
    const double *fv[3] = {o.fa, o.fb, o.fc};
    const double md = o.t8 / 12.0, mo = o.t8 / 24.0;      // ComputeInertial.cpp:44-47: Mi_vw = m I
    const double FtF[4] = {dot(F0, F0), dot(F0, F1), dot(F1, F0), dot(F1, F1)};
    for (int P = 0; P < 3; ++P) {
        if (!((mask >> P) & 1)) continue;
        out[FACE_FX + 2 * P] = -(F0.x * fv[P][0] + F0.y * fv[P][1] + F0.z * fv[P][2]);
        out[FACE_FX + 2 * P + 1] = -(F1.x * fv[P][0] + F1.y * fv[P][1] + F1.z * fv[P][2]);
        for (int q = 0; q < 3; ++q) {
            double B[9];
            get_block(o.K, o.K + 3, 3, P, q, B);
            double *dst = out + FACE_Xx + 12 * (3 * P + q);
            neg_Ft_B(F0, F1, B, dst);
            const double m = P == q ? md : mo;
            dst[6] = -(m * F0.x); dst[7] = -(m * F0.y); dst[8] = -(m * F0.z); dst[9] = -(m * F1.x); dst[10] = -(m * F1.y); dst[11] = -(m * F1.z);
            if (q >= P && ((mask >> q) & 1)) {
                double *d2 = out + FACE_XX + 8 * (3 * P + q);
                Ft_B_F(F0, F1, B, d2);
                for (int k = 0; k < 4; ++k) d2[4 + k] = m * FtF[k];
            }
        }
    }
}

// bending force of one stencil, f_i = c (g0 x w0_i + g1 x w1_i) with W = c (1 - D) (ComputeBending.cpp:102, :126-258; notation of elements.cuh)
EOLC_HD void edge_force(v3 x0, v3 x1, v3 x2, v3 x3, double X0x, double X0y, double X1x, double X1y, double X2x, double X2y, double X3x,
                        double X3y, double beta, v3 *f) {
    const double ex = X1x - X0x, ey = X1y - X0y;
    const double t6 = beta * (ex * ex + ey * ey);
    const double den = stencil_area(X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y);
    const double c = 1.5 * t6 / den;
    const v3 e = x1 - x0, a = x2 - x0, b = x3 - x0;
    const v3 n0 = cross(e, a), n1 = cross(b, e);
    const double il0 = 1.0 / sqrt(dot(n0, n0)), il1 = 1.0 / sqrt(dot(n1, n1));
    const v3 u = il0 * n0, v = il1 * n1;
    const double D = dot(u, v);
    const v3 g0 = (c * il0) * (v - D * u), g1 = (c * il1) * (u - D * v);
    f[0] = cross(g0, x2 - x1) + cross(g1, x1 - x3);
    f[1] = cross(g0, x0 - x2) + cross(g1, b);
    f[2] = cross(g0, e);
    f[3] = cross(g1, x0 - x1);
}

// One EOL stencil: rec = (n0, n1, o0, o1, mask).  Forces.cpp:746-883 with fillEOLBending; F = (F(adjf0) + F(adjf1)) / 2 (:590-596),
// the faces taken as (n0, n1, o0) and (n0, n1, o1) — F does not depend on the vertex order beyond rounding.
EOLC_HD void edge_record(const int32_t *rec, const double *x, const double *X, const Params &p, double *out) {
    const int32_t n[4] = {rec[0], rec[1], rec[2], rec[3]};
    const int mask = rec[4];
    v3 xs[4];
    double Xs[8];
    for (int v = 0; v < 4; ++v) {
        xs[v] = mk3(x[3 * (size_t)n[v]], x[3 * (size_t)n[v] + 1], x[3 * (size_t)n[v] + 2]);
        Xs[2 * v] = X[2 * (size_t)n[v]]; Xs[2 * v + 1] = X[2 * (size_t)n[v] + 1];
    }
    EdgeOut o;
    edge_element(xs[0], xs[1], xs[2], xs[3], Xs[0], Xs[1], Xs[2], Xs[3], Xs[4], Xs[5], Xs[6], Xs[7], p.beta, p.dhh, o);
    v3 fb[4];
    edge_force(xs[0], xs[1], xs[2], xs[3], Xs[0], Xs[1], Xs[2], Xs[3], Xs[4], Xs[5], Xs[6], Xs[7], p.beta, fb);
    v3 A0, A1, B0, B1;
    deform_grad(xs[0], xs[1], xs[2], Xs[0], Xs[1], Xs[2], Xs[3], Xs[4], Xs[5], A0, A1);
    deform_grad(xs[0], xs[1], xs[3], Xs[0], Xs[1], Xs[2], Xs[3], Xs[6], Xs[7], B0, B1);
    const v3 F0 = 0.5 * (A0 + B0), F1 = 0.5 * (A1 + B1);
    for (int v = 0; v < 4; ++v) { out[EDGE_FX_LAG + 3 * v] = fb[v].x; out[EDGE_FX_LAG + 3 * v + 1] = fb[v].y; out[EDGE_FX_LAG + 3 * v + 2] = fb[v].z; }
    for (int P = 0; P < 4; ++P) {
        if (!((mask >> P) & 1)) continue;
        out[EDGE_FX + 2 * P] = -dot(F0, fb[P]);
        out[EDGE_FX + 2 * P + 1] = -dot(F1, fb[P]);
        for (int q = 0; q < 4; ++q) {
            double B[9];
            get_block(o.K, o.K + 4, 4, P, q, B);
            neg_Ft_B(F0, F1, B, out + EDGE_Xx + 6 * (4 * P + q));
            if (q >= P && ((mask >> q) & 1)) Ft_B_F(F0, F1, B, out + EDGE_XX + 4 * (4 * P + q));
        }
    }
}

// ---- gather ------------------------------------------------------------------------------------------------------------
// One target = one output entry: slot (value index; for f the row) | array << 62 (0 = MDK, 1 = M, 2 = f), and its sources
// [begin, begin + count) in `sources` (indices into the scene's scratch), summed left to right.  f rows below 3N are ADDED
// onto what the tiles kernel wrote (the bending force of the EOL stencils); everything else is set.
struct Target { uint32_t slot_lo, slot_hi, begin, count; };
static_assert(sizeof(Target) == 16, "Target");

EOLC_HD void gather_target(const Target &t, const uint32_t *sources, const double *scratch, uint32_t lag_dof, double *f, double *Mv, double *Kv) {
    const uint32_t arr = t.slot_hi >> 30;
    const unsigned long long slot = ((unsigned long long)(t.slot_hi & 0x3fffffffu) << 32) | t.slot_lo;
    double s = 0.0;
    for (uint32_t k = 0; k < t.count; ++k) s += scratch[sources[t.begin + k]];
    if (arr == 0) Kv[slot] = s;
    else if (arr == 1) Mv[slot] = s;
    else if (slot < lag_dof) f[slot] += s;
    else f[slot] = s;
}

// ---- plan (host) ---------------------------------------------------------------------------------------------------------
struct Plan {
    int32_t N = 0, n_eol = 0, dof = 0;
    int64_t nnzM = 0, nnzK = 0;
    std::vector<int32_t> outerM, innerM, outerK, innerK;   // scalar CSR (== CSC) arrays of the full matrices, Eigen layout
    std::vector<int64_t> dstM, dstK;                       // per node: value index of scalar row 3a
    std::vector<int32_t> extraM, extraK;                   // per node: Eulerian columns per scalar row of the node
    std::vector<int32_t> faces;                            // 4 per EOL face: a, b, c, mask
    std::vector<int32_t> edges;                            // 8 per EOL stencil: n0, n1, o0, o1, mask, 0, 0, 0
    std::vector<Target> targets;
    std::vector<uint32_t> sources;
    int64_t scratch_doubles = 0;
    int32_t n_faces() const { return (int32_t)(faces.size() / 4); }
    int32_t n_edges() const { return (int32_t)(edges.size() / 8); }
    std::string error;
};

// blkptr / nbr: the block pattern of the Lagrangian part (forces_plan.h Pattern); ie: 4 per interior stencil
inline bool build(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie, const int32_t *eol_index,
                  const std::vector<int64_t> &blkptrM, const std::vector<int32_t> &nbrM, const std::vector<int64_t> &blkptrK,
                  const std::vector<int32_t> &nbrK, Plan &P) {
    P = Plan();
    P.N = N;
    for (int32_t a = 0; a < N; ++a) P.n_eol = std::max(P.n_eol, eol_index[a] + 1);
    {
        std::vector<char> seen((size_t)P.n_eol, 0);
        for (int32_t a = 0; a < N; ++a)
            if (eol_index[a] >= 0) {
                if (seen[eol_index[a]]) { P.error = "EoL_index " + std::to_string(eol_index[a]) + " is used by two nodes"; return false; }
                seen[eol_index[a]] = 1;
            }
    }
    const int32_t L = 3 * N;
    P.dof = L + 2 * P.n_eol;
    auto is_eol = [&](int32_t a) { return eol_index[a] >= 0; };
    // EOL elements, ascending element order
    for (int32_t i = 0; i < F; ++i) {
        const int32_t *v = fn + 3 * (size_t)i;
        const int mask = (is_eol(v[0]) ? 1 : 0) | (is_eol(v[1]) ? 2 : 0) | (is_eol(v[2]) ? 4 : 0);
        if (mask) { P.faces.insert(P.faces.end(), v, v + 3); P.faces.push_back(mask); }
    }
    for (int32_t i = 0; i < Ei; ++i) {
        const int32_t *v = ie + 4 * (size_t)i;
        const int mask = (is_eol(v[0]) ? 1 : 0) | (is_eol(v[1]) ? 2 : 0) | (is_eol(v[2]) ? 4 : 0) | (is_eol(v[3]) ? 8 : 0);
        if (mask) { P.edges.insert(P.edges.end(), v, v + 4); P.edges.push_back(mask); P.edges.insert(P.edges.end(), 3, 0); }
    }
    P.scratch_doubles = (int64_t)P.n_faces() * FACE_REC + (int64_t)P.n_edges() * EDGE_REC;
    if (P.scratch_doubles >= ((int64_t)1 << 32)) { P.error = "too many EOL elements"; return false; }
    // pattern: Eulerian columns of the Lagrangian rows (per node, EoL indices) and the Eulerian rows (Lagrangian nodes + EoL indices)
    std::vector<std::vector<int32_t>> exM(N), exK(N), lagM(P.n_eol), lagK(P.n_eol), eeM(P.n_eol), eeK(P.n_eol);
    auto couple = [&](const int32_t *v, int nv, int mask, bool mass) {
        for (int Pv = 0; Pv < nv; ++Pv) {
            if (!((mask >> Pv) & 1)) continue;
            const int32_t k = eol_index[v[Pv]];
            for (int q = 0; q < nv; ++q) {
                exK[v[q]].push_back(k); lagK[k].push_back(v[q]);
                if (mass) { exM[v[q]].push_back(k); lagM[k].push_back(v[q]); }
                if ((mask >> q) & 1) { eeK[k].push_back(eol_index[v[q]]); if (mass) eeM[k].push_back(eol_index[v[q]]); }
            }
        }
    };
    for (int32_t i = 0; i < P.n_faces(); ++i) couple(P.faces.data() + 4 * (size_t)i, 3, P.faces[4 * (size_t)i + 3], true);
    for (int32_t i = 0; i < P.n_edges(); ++i) couple(P.edges.data() + 8 * (size_t)i, 4, P.edges[8 * (size_t)i + 4], false);
    auto uniq = [](std::vector<std::vector<int32_t>> &V) { for (auto &s : V) { std::sort(s.begin(), s.end()); s.erase(std::unique(s.begin(), s.end()), s.end()); } };
    uniq(exM); uniq(exK); uniq(lagM); uniq(lagK); uniq(eeM); uniq(eeK);
    auto arrays = [&](const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr, const std::vector<std::vector<int32_t>> &ex,
                      const std::vector<std::vector<int32_t>> &lag, const std::vector<std::vector<int32_t>> &ee, std::vector<int32_t> &outer,
                      std::vector<int32_t> &inner, std::vector<int64_t> &dst, std::vector<int32_t> &extra, int64_t &nnz) {
        int64_t total = 9 * blkptr[N];
        for (int32_t a = 0; a < N; ++a) total += 6 * (int64_t)ex[a].size();
        for (int32_t k = 0; k < P.n_eol; ++k) total += 2 * (3 * (int64_t)lag[k].size() + 2 * (int64_t)ee[k].size());
        nnz = total;
        if (total > (int64_t)INT32_MAX) return false;
        outer.assign((size_t)P.dof + 1, 0); inner.resize((size_t)total); dst.assign(N, 0); extra.assign(N, 0);
        int64_t at = 0;
        for (int32_t a = 0; a < N; ++a) {
            const int64_t b0 = blkptr[a];
            const int deg = (int)(blkptr[a + 1] - b0);
            dst[a] = at; extra[a] = 2 * (int32_t)ex[a].size();
            for (int j = 0; j < 3; ++j) {
                outer[3 * (size_t)a + j] = (int32_t)at;
                for (int p = 0; p < deg; ++p) for (int k = 0; k < 3; ++k) inner[at++] = 3 * nbr[b0 + p] + k;
                for (int32_t k : ex[a]) { inner[at++] = L + 2 * k; inner[at++] = L + 2 * k + 1; }
            }
        }
        for (int32_t k = 0; k < P.n_eol; ++k)
            for (int c = 0; c < 2; ++c) {
                outer[(size_t)L + 2 * k + c] = (int32_t)at;
                for (int32_t q : lag[k]) for (int j = 0; j < 3; ++j) inner[at++] = 3 * q + j;
                for (int32_t k2 : ee[k]) { inner[at++] = L + 2 * k2; inner[at++] = L + 2 * k2 + 1; }
            }
        outer[(size_t)P.dof] = (int32_t)at;
        return at == total;
    };
    if (!arrays(blkptrM, nbrM, exM, lagM, eeM, P.outerM, P.innerM, P.dstM, P.extraM, P.nnzM) ||
        !arrays(blkptrK, nbrK, exK, lagK, eeK, P.outerK, P.innerK, P.dstK, P.extraK, P.nnzK)) {
        P.error = "nnz exceeds int32 (Eigen StorageIndex is int)";
        return false;
    }
    // targets: (key = array << 62 | slot, source) pairs in the reference's insertion order, then grouped by key (stable)
    std::vector<std::pair<uint64_t, uint32_t>> pairs;
    bool ok = true;
    auto slot_of = [&](int which, int32_t R, int32_t C) -> uint64_t {
        const std::vector<int32_t> &o = which ? P.outerM : P.outerK, &in = which ? P.innerM : P.innerK;
        auto b = in.begin() + o[R], e = in.begin() + o[R + 1];
        auto it = std::lower_bound(b, e, C);
        if (it == e || *it != C) { ok = false; return 0; }
        return (uint64_t)(it - in.begin());
    };
    auto put = [&](int which /*0 = MDK, 1 = M*/, int32_t R, int32_t C, uint32_t src, bool mirror) {
        pairs.push_back({((uint64_t)which << 62) | slot_of(which, R, C), src});
        if (mirror) pairs.push_back({((uint64_t)which << 62) | slot_of(which, C, R), src});
    };
    auto put_f = [&](int32_t R, uint32_t src) { pairs.push_back({((uint64_t)2 << 62) | (uint64_t)R, src}); };
    for (int32_t i = 0; i < P.n_faces(); ++i) {
        const int32_t *v = P.faces.data() + 4 * (size_t)i;
        const int mask = v[3];
        const uint32_t base = (uint32_t)i * FACE_REC;
        for (int Pv = 0; Pv < 3; ++Pv) {
            if (!((mask >> Pv) & 1)) continue;
            const int32_t XP = L + 2 * eol_index[v[Pv]];
            for (int c = 0; c < 2; ++c) put_f(XP + c, base + FACE_FX + 2 * Pv + c);
            for (int q = 0; q < 3; ++q) {
                for (int c = 0; c < 2; ++c)
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t s = base + FACE_Xx + 12 * (3 * Pv + q) + 3 * c + j;
                        put(0, XP + c, 3 * v[q] + j, s, true);
                        put(1, XP + c, 3 * v[q] + j, s + 6, true);
                    }
                if (q >= Pv && ((mask >> q) & 1)) {
                    const int32_t XQ = L + 2 * eol_index[v[q]];
                    for (int c = 0; c < 2; ++c)
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const uint32_t s = base + FACE_XX + 8 * (3 * Pv + q) + 2 * c + c2;
                            put(0, XP + c, XQ + c2, s, q != Pv);
                            put(1, XP + c, XQ + c2, s + 4, q != Pv);
                        }
                }
            }
        }
    }
    const uint32_t ebase = (uint32_t)P.n_faces() * FACE_REC;
    for (int32_t i = 0; i < P.n_edges(); ++i) {
        const int32_t *v = P.edges.data() + 8 * (size_t)i;
        const int mask = v[4];
        const uint32_t base = ebase + (uint32_t)i * EDGE_REC;
        for (int q = 0; q < 4; ++q) for (int j = 0; j < 3; ++j) put_f(3 * v[q] + j, base + EDGE_FX_LAG + 3 * q + j);
        for (int Pv = 0; Pv < 4; ++Pv) {
            if (!((mask >> Pv) & 1)) continue;
            const int32_t XP = L + 2 * eol_index[v[Pv]];
            for (int c = 0; c < 2; ++c) put_f(XP + c, base + EDGE_FX + 2 * Pv + c);
            for (int q = 0; q < 4; ++q) {
                for (int c = 0; c < 2; ++c)
                    for (int j = 0; j < 3; ++j) put(0, XP + c, 3 * v[q] + j, base + EDGE_Xx + 6 * (4 * Pv + q) + 3 * c + j, true);
                if (q >= Pv && ((mask >> q) & 1)) {
                    const int32_t XQ = L + 2 * eol_index[v[q]];
                    for (int c = 0; c < 2; ++c)
                        for (int c2 = 0; c2 < 2; ++c2) put(0, XP + c, XQ + c2, base + EDGE_XX + 4 * (4 * Pv + q) + 2 * c + c2, q != Pv);
                }
            }
        }
    }
    if (!ok) { P.error = "internal: EOL entry outside the pattern"; return false; }
    // every Eulerian f row is a target, also the ones nothing contributes to (an EoL node without faces): they must be zeroed
    std::vector<char> has_f((size_t)2 * P.n_eol, 0);
    for (const auto &pr : pairs) if ((pr.first >> 62) == 2 && (pr.first & 0x3fffffffffffffffull) >= (uint64_t)L) has_f[(pr.first & 0x3fffffffffffffffull) - L] = 1;
    std::stable_sort(pairs.begin(), pairs.end(), [](const std::pair<uint64_t, uint32_t> &a, const std::pair<uint64_t, uint32_t> &b) { return a.first < b.first; });
    P.sources.reserve(pairs.size());
    for (size_t i = 0; i < pairs.size();) {
        size_t j = i;
        while (j < pairs.size() && pairs[j].first == pairs[i].first) ++j;
        Target t;
        t.slot_lo = (uint32_t)pairs[i].first; t.slot_hi = (uint32_t)(pairs[i].first >> 32);
        t.begin = (uint32_t)P.sources.size(); t.count = (uint32_t)(j - i);
        for (size_t k = i; k < j; ++k) P.sources.push_back(pairs[k].second);
        P.targets.push_back(t);
        i = j;
    }
    for (int32_t r = 0; r < 2 * P.n_eol; ++r)
        if (!has_f[r]) { Target t; t.slot_lo = (uint32_t)(L + r); t.slot_hi = 2u << 30; t.begin = 0; t.count = 0; P.targets.push_back(t); }
    return true;
}

}  // namespace eol
}  // namespace eolc
