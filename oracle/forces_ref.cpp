// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's Forces::fill (Lagrangian and EOL branches) over flat
// arrays, Eigen-free, single-threaded, linking the reference's own generated
// arithmetic (ComputeMembrane.cpp / ComputeBending.cpp / ComputeInertial.cpp,
// compiled UNMODIFIED from /root/reference/src into oracle/_ref/ by
// oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this.
//
// PARITY UNPINNED by the reference's own tests: the reference ships no tests or
// golden vectors and Eigen is absent, so Forces.cpp itself cannot be
// compiled here.  What IS pinned: the three Compute* kernels are the reference's
// own object code, and the known-answer vectors of SURVEY.md §8c are checked in
// tests/test_oracle.py.  The glue below follows the reference line by line:
//
//   poldec            src/Forces.cpp:33-52
//   faceBasedF        src/Forces.cpp:331-397 (frame, F, Fbar, Q) and :498-518 (non-EOL scatter)
//   fillxMI/fillxxMI  src/Forces.cpp:103-125
//   edgeBasedF        src/Forces.cpp:685-744 and :885-908 (non-EOL scatter)
//   fillxB/fillxxB    src/Forces.cpp:522-539
//   Forces::fill      src/Forces.cpp:912-930
//   EOL branch        deform_grad src/UtilEOL.cpp:13-28; fillEOLInertia / fillEOLMembrane src/Forces.cpp:177-329;
//                     face scatter :399-497 with fill{X,XX,Xx,xX}MI :127-175; fillEOLBending :580-683; edge scatter
//                     :746-883 with fill{X,XX,Xx,xX}B :541-578.  The inertial and membrane expansions are the same code
//                     with K = Mi / Km, and the bending one is the 4-vertex version of it, so one routine (expand_eol)
//                     restates all three; its body follows fillEOLBending block by block.
//   setFromTriplets   Eigen 3.3 SparseMatrix.h set_from_triplets (external, restated
//                     from its published algorithm: bucket by row in insertion order,
//                     collapse duplicates onto the first occurrence left-to-right,
//                     transposed copy -> column-major with sorted inner indices;
//                     explicit zeros are kept).
//
// Eigen small-vector arithmetic conventions restated here (Eigen 3.3.x, SSE2,
// EIGEN_DONT_ALIGN_STATICALLY — Forces.h:13): a 3-vector dot/squaredNorm is the
// linear-vectorised redux  (p0 + p1) + p2 ;  v / s is a true division per
// component; cross is the textbook formula; DX.inverse() (dynamic 2x2, Eigen uses
// PartialPivLU) is restated in closed form (adjugate / det) — an O(1 ulp)
// difference, inside the 1e-10 budget (SURVEY.md §8c rule 4).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>
#include <thread>
#include <utility>

// reference prototypes, exactly as in src/Compute*.h (C++ linkage)
void ComputeMembrane(const double *xa, const double *xb, const double *xc,
                     const double *Xa, const double *Xb, const double *Xc,
                     double e, double nu, const double *P, const double *Q,
                     double *W, double *f, double *K);
void ComputeBending(const double *x0, const double *x1, const double *x2, const double *x3,
                    const double *X0, const double *X1, const double *X2, const double *X3,
                    double beta, double *W, double *f, double *K);
void ComputeInertial(const double *xa, const double *xb, const double *xc,
                     const double *Xa, const double *Xb, const double *Xc,
                     const double *g, double rho, double *W, double *f, double *M);

namespace {

struct Trip { int32_t r, c; double v; };

inline double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline void cross3(double *d, const double *a, const double *b) {
    d[0] = a[1] * b[2] - a[2] * b[1];
    d[1] = a[2] * b[0] - a[0] * b[2];
    d[2] = a[0] * b[1] - a[1] * b[0];
}

// Forces.cpp:33-52 ; M and Q column-major 2x2
void poldec(const double *M, double *Q) {
    double m11 = M[0], m21 = M[1], m12 = M[2], m22 = M[3];
    double detM = m11 * m22 - m12 * m21;
    int sign = 1;
    if (detM < 0) sign = -1;
    else if (detM == 0) sign = 0;
    // MM << m22, -m21, -m12, m11  (row-major fill): MM(0,0)=m22 MM(0,1)=-m21 MM(1,0)=-m12 MM(1,1)=m11
    double q00 = m11 + sign * m22;
    double q01 = m12 + sign * (-m21);
    double q10 = m21 + sign * (-m12);
    double q11 = m22 + sign * m11;
    double clen = std::sqrt(q00 * q00 + q10 * q10);
    Q[0] = q00 / clen; Q[1] = q10 / clen; Q[2] = q01 / clen; Q[3] = q11 / clen;
}

struct Csc {
    int64_t nnz = 0;
    std::vector<int32_t> outer, inner;
    std::vector<double> vals;
};

// Eigen 3.3 set_from_triplets, restated.  Result: column-major compressed.
void set_from_triplets(int dof, const std::vector<Trip> &T, Csc &out) {
    // pass 1: count per row (trMat is row-major)
    std::vector<int64_t> start(dof + 1, 0);
    for (const Trip &t : T) start[t.r + 1]++;
    for (int i = 0; i < dof; ++i) start[i + 1] += start[i];
    // pass 2: insertBackUncompressed -> each row keeps insertion order
    std::vector<int32_t> tc(T.size());
    std::vector<double> tv(T.size());
    {
        std::vector<int64_t> fillp(start.begin(), start.end() - 1);
        for (const Trip &t : T) { int64_t p = fillp[t.r]++; tc[p] = t.c; tv[p] = t.v; }
    }
    // pass 3: collapseDuplicates: per row, first occurrence keeps the slot, later ones add into it
    std::vector<int64_t> wi(dof, -1);
    std::vector<int64_t> rstart(dof + 1, 0);
    int64_t count = 0;
    for (int r = 0; r < dof; ++r) {
        int64_t s = count;
        rstart[r] = s;
        for (int64_t k = start[r]; k < start[r + 1]; ++k) {
            int32_t c = tc[k];
            if (wi[c] >= s) {
                tv[wi[c]] = tv[wi[c]] + tv[k];   // dup_func = scalar_sum_op: old + new
            } else {
                tv[count] = tv[k]; tc[count] = c; wi[c] = count; ++count;
            }
        }
    }
    rstart[dof] = count;
    // pass 4: mat = trMat (transposed copy): column-major, rows ascending within a column
    out.nnz = count;
    out.outer.assign(dof + 1, 0);
    for (int64_t k = 0; k < count; ++k) out.outer[tc[k] + 1]++;
    for (int i = 0; i < dof; ++i) out.outer[i + 1] += out.outer[i];
    out.inner.resize(count); out.vals.resize(count);
    std::vector<int32_t> pos(out.outer.begin(), out.outer.end() - 1);
    for (int r = 0; r < dof; ++r)
        for (int64_t k = rstart[r]; k < rstart[r + 1]; ++k) {
            int32_t p = pos[tc[k]]++;
            out.inner[p] = r; out.vals[p] = tv[k];
        }
}

// ---- EOL branch -------------------------------------------------------------------------------------------------
// deform_grad, UtilEOL.cpp:13-28: F = Dx * DX.inverse() (3x2, returned column-major F[0..2] = col 0, F[3..5] = col 1).
// DX is a fixed-size Matrix2d there: Eigen's 2x2 inverse is adjugate * (1/det).
void deform_grad(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb, const double *Xc, double *F) {
    const double Dx[6] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2], xc[0] - xa[0], xc[1] - xa[1], xc[2] - xa[2]};
    const double D00 = Xb[0] - Xa[0], D01 = Xc[0] - Xa[0], D10 = Xb[1] - Xa[1], D11 = Xc[1] - Xa[1];
    const double invdet = 1.0 / (D00 * D11 - D10 * D01);
    const double i00 = D11 * invdet, i01 = -D01 * invdet, i10 = -D10 * invdet, i11 = D00 * invdet;
    for (int r = 0; r < 3; ++r) {
        F[r] = Dx[r] * i00 + Dx[3 + r] * i10;
        F[3 + r] = Dx[r] * i01 + Dx[3 + r] * i11;
    }
}

// The expanded element of fillEOLInertia / fillEOLMembrane (nv = 3, Forces.cpp:177-329) and fillEOLBending (nv = 4,
// :580-683): local dofs [x_v (3), X_v (2)] per vertex, vertex v at jv = 5 v, its Eulerian pair at jV = 5 v + 3.
// f: 3 nv, K(r, c): 3nv x 3nv accessor; F: the one deformation gradient every EoL vertex of the element uses (:188-191,
// :265-268, :593-596).  fe (5 nv) and Ke (5nv x 5nv, column-major, ld = 5 nv) are filled exactly where the reference
// fills them; everything else is NaN so that a read of an entry the reference leaves uninitialised shows up.
template <typename KAcc>
void expand_eol(int nv, const double *f, KAcc K, const double *F, const bool *eol, double *fe, double *Ke) {
    const int n = 5 * nv;
    const double nan = std::nan("");
    for (int i = 0; i < n; ++i) fe[i] = nan;
    for (int i = 0; i < n * n; ++i) Ke[i] = nan;
    auto KE = [&](int r, int c) -> double & { return Ke[c * n + r]; };
    auto Fm = [&](int r, int c) { return F[3 * c + r]; };   // 3x2
    // x parts: segments and the upper block triangle (:199-205)
    for (int v = 0; v < nv; ++v) {
        for (int j = 0; j < 3; ++j) fe[5 * v + j] = f[3 * v + j];
        for (int w = v; w < nv; ++w)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) KE(5 * v + j, 5 * w + k) = K(3 * v + j, 3 * w + k);
    }
    // Ft K_vw (2x3) and Ft K_vw F (2x2), products evaluated left to right like Fa.transpose() * Kiab * Fb
    auto FtK = [&](int v, int w, double *T /*2x3 row-major*/) {
        for (int a = 0; a < 2; ++a)
            for (int k = 0; k < 3; ++k)
                T[3 * a + k] = (Fm(0, a) * K(3 * v + 0, 3 * w + k) + Fm(1, a) * K(3 * v + 1, 3 * w + k)) + Fm(2, a) * K(3 * v + 2, 3 * w + k);
    };
    for (int v = 0; v < nv; ++v) {          // :207-218
        if (!eol[v]) continue;
        for (int a = 0; a < 2; ++a)
            fe[5 * v + 3 + a] = ((-Fm(0, a)) * f[3 * v] + (-Fm(1, a)) * f[3 * v + 1]) + (-Fm(2, a)) * f[3 * v + 2];
        double T[6];
        FtK(v, v, T);
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) KE(5 * v + 3 + a, 5 * v + 3 + b) = (T[3 * a] * Fm(0, b) + T[3 * a + 1] * Fm(1, b)) + T[3 * a + 2] * Fm(2, b);
    }
    for (int v = 0; v < nv; ++v)            // :220-228
        for (int w = v + 1; w < nv; ++w) {
            if (!(eol[v] && eol[w])) continue;
            double T[6];
            FtK(v, w, T);
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) KE(5 * v + 3 + a, 5 * w + 3 + b) = (T[3 * a] * Fm(0, b) + T[3 * a + 1] * Fm(1, b)) + T[3 * a + 2] * Fm(2, b);
        }
    for (int v = 0; v < nv; ++v) {          // :230-251
        if (!eol[v]) continue;
        for (int w = v; w < nv; ++w)        // X_v - x_w = -Ft K_vw
            for (int a = 0; a < 2; ++a)
                for (int k = 0; k < 3; ++k)
                    KE(5 * v + 3 + a, 5 * w + k) = ((-Fm(0, a)) * K(3 * v + 0, 3 * w + k) + (-Fm(1, a)) * K(3 * v + 1, 3 * w + k)) + (-Fm(2, a)) * K(3 * v + 2, 3 * w + k);
        for (int u = 0; u < v; ++u)         // x_u - X_v = -K_uv F
            for (int j = 0; j < 3; ++j)
                for (int b = 0; b < 2; ++b)
                    KE(5 * u + j, 5 * v + 3 + b) = ((-K(3 * u + j, 3 * v + 0)) * Fm(0, b) + (-K(3 * u + j, 3 * v + 1)) * Fm(1, b)) + (-K(3 * u + j, 3 * v + 2)) * Fm(2, b);
    }
}

struct Result {
    int dof = 0;
    std::vector<double> f;
    Csc M, MDK;
    double seconds_elements = 0, seconds_assembly = 0;
};

}  // namespace

extern "C" {

// mat = {density, e, nu, beta, dampingA, dampingB}  (src/Cloth.h:28-35)
// edge_stencil: 4 ints per mesh edge (n0, n1, opp(adjf0), opp(adjf1)); -1 in slot 2/3 = boundary.
// flags bit0: skip the assembly (setFromTriplets) — used only by the cpu_baseline timing split.
// eol_index: N ints, -1 = Lagrangian node, k >= 0 = Node::EoL_index of an EoL node (may be NULL: no EoL node);
// mesh.EoL_Count = 1 + the largest index.
void *oracle_forces_fill_eol(int N, int F, const int32_t *face_nodes, int E, const int32_t *edge_stencil,
                             const double *x, const double *X, const double *mat, const double *grav, double h,
                             const int32_t *eol_index, int flags) {
    auto t0 = std::chrono::steady_clock::now();
    Result *R = new Result;
    int eol_count = 0;
    if (eol_index) for (int a = 0; a < N; ++a) eol_count = std::max(eol_count, eol_index[a] + 1);
    const int dof = 3 * N + 2 * eol_count;       // Forces.cpp:914
    auto is_eol = [&](int a) { return eol_index && eol_index[a] >= 0; };
    auto Xdof = [&](int a) { return 3 * N + 2 * eol_index[a]; };   // :379-381
    // edge->adjf[k] of the EOL bending branch (:590-591): the face holding the edge's two nodes and the stencil's opposite node
    std::vector<std::pair<uint64_t, int>> face_key;
    if (eol_count) {
        face_key.reserve(F);
        for (int i = 0; i < F; ++i) {
            int v[3] = {face_nodes[3 * i], face_nodes[3 * i + 1], face_nodes[3 * i + 2]};
            std::sort(v, v + 3);
            face_key.push_back({((uint64_t)v[0] << 42) | ((uint64_t)v[1] << 21) | (uint64_t)v[2], i});
        }
        std::sort(face_key.begin(), face_key.end());
    }
    auto find_face = [&](int a, int b, int c) {
        int v[3] = {a, b, c};
        std::sort(v, v + 3);
        const uint64_t key = ((uint64_t)v[0] << 42) | ((uint64_t)v[1] << 21) | (uint64_t)v[2];
        auto it = std::lower_bound(face_key.begin(), face_key.end(), std::make_pair(key, -1));
        return (it != face_key.end() && it->first == key) ? it->second : -1;
    };
    R->dof = dof;
    R->f.assign(dof, 0.0);                       // Forces.cpp:914-915
    std::vector<Trip> M_, MDK_;                  // :916-917
    const double density = mat[0], e = mat[1], nu = mat[2], beta = mat[3], dampingB = mat[5];

    // The two element loops are lambdas over an index range so that the TIMING variant (flags bits 8..15 = worker threads, bench.py's
    // CPU legs) can run ranges on several threads; with one thread they run exactly as the reference's loops do.  fl != NULL: the
    // f contributions are logged instead of added, and replayed in element order afterwards (same sums, bit for bit).
    typedef std::vector<std::pair<int32_t, double>> FLog;
    // ---- faceBasedF, Forces.cpp:331-520 ----
    auto do_faces = [&](int i_begin, int i_end, std::vector<Trip> &M_, std::vector<Trip> &MDK_, FLog *fl) {
    auto fadd = [&](int at, double v) { if (fl) fl->push_back({at, v}); else R->f[at] += v; };
    for (int i = i_begin; i < i_end; ++i) {
        const int ia = face_nodes[3 * i], ib = face_nodes[3 * i + 1], ic = face_nodes[3 * i + 2];
        const double *xa = x + 3 * ia, *xb = x + 3 * ib, *xc = x + 3 * ic;
        const double *Xa = X + 2 * ia, *Xb = X + 2 * ib, *Xc = X + 2 * ic;
        double PP[6], QQ[4], Wi[1], Wm[1];
        double d1[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};   // Dxt col 0
        double d2[3] = {xc[0] - xa[0], xc[1] - xa[1], xc[2] - xa[2]};   // Dxt col 1
        double DX[4] = {Xb[0] - Xa[0], Xb[1] - Xa[1], Xc[0] - Xa[0], Xc[1] - Xa[1]};  // col-major
        double normm[3]; cross3(normm, d1, d2);                          // :361
        double l1 = std::sqrt(dot3(d1, d1));
        double Pxm[3] = {d1[0] / l1, d1[1] / l1, d1[2] / l1};            // :362
        double Pym[3]; cross3(Pym, normm, Pxm);                          // :363
        double l2 = std::sqrt(dot3(Pym, Pym));
        Pym[0] /= l2; Pym[1] /= l2; Pym[2] /= l2;                        // :364
        // DX.inverse(), closed form
        double det = DX[0] * DX[3] - DX[2] * DX[1];
        double inv[4] = {DX[3] / det, -DX[1] / det, -DX[2] / det, DX[0] / det};  // col-major
        // Fm = Dxt * inv  (3x2), col-major
        double Fm[6];
        for (int r = 0; r < 3; ++r) {
            Fm[r] = d1[r] * inv[0] + d2[r] * inv[1];
            Fm[3 + r] = d1[r] * inv[2] + d2[r] * inv[3];
        }
        // Fbarm = Pm * Fm (2x2), col-major
        double Fb[4] = {dot3(Pxm, Fm), dot3(Pym, Fm), dot3(Pxm, Fm + 3), dot3(Pym, Fm + 3)};
        poldec(Fb, QQ);                                                  // :369
        // PP column-major 2x3                                            // :371
        PP[0] = Pxm[0]; PP[1] = Pym[0]; PP[2] = Pxm[1]; PP[3] = Pym[1]; PP[4] = Pxm[2]; PP[5] = Pym[2];
        double fm[9], Km[81], fi[9], Mi[81];
        ComputeMembrane(xa, xb, xc, Xa, Xb, Xc, e, nu, PP, QQ, Wm, fm, Km);   // :389
        ComputeInertial(xa, xb, xc, Xa, Xb, Xc, grav, density, Wi, fi, Mi);   // :390
        // Kme(r,c) = Km[c*9+r] (column-major Map, :393)
        const int idx[3] = {3 * ia, 3 * ib, 3 * ic};
        const bool eolv[3] = {is_eol(ia), is_eol(ib), is_eol(ic)};
        if (eolv[0] || eolv[1] || eolv[2]) {                                  // :399
            const double dhh = dampingB * h * h;
            const int nodes[3] = {ia, ib, ic};
            double Fg[6], fme[15], fie[15], Kme[225], Mie[225];
            deform_grad(xa, xb, xc, Xa, Xb, Xc, Fg);                          // :188, :265
            expand_eol(3, fi, [&](int r, int c) { return Mi[c * 9 + r]; }, Fg, eolv, fie, Mie);   // fillEOLInertia :401
            expand_eol(3, fm, [&](int r, int c) { return Km[c * 9 + r]; }, Fg, eolv, fme, Kme);   // fillEOLMembrane :403
            auto KE = [&](int r, int c) { return Kme[c * 15 + r]; };
            auto ME = [&](int r, int c) { return Mie[c * 15 + r]; };
            for (int v = 0; v < 3; ++v) {                                     // :405-412
                for (int j = 0; j < 3; ++j) fadd(idx[v] + j, fme[5 * v + j] + fie[5 * v + j]);
                if (eolv[v]) for (int j = 0; j < 2; ++j) fadd(Xdof(nodes[v]) + j, fme[5 * v + 3 + j] + fie[5 * v + 3 + j]);
            }
            // one (nr x nc) block at local (lr, lc) -> global (gr, gc); mirror = the fillxx / fillXX / fillXx / fillxX forms
            auto put = [&](int lr, int lc, int nr, int nc, int gr, int gc, bool mirror) {
                for (int j = 0; j < nr; ++j)
                    for (int k = 0; k < nc; ++k) {
                        const double m = ME(lr + j, lc + k);
                        const double mdk = m + dhh * KE(lr + j, lc + k);
                        M_.push_back({gr + j, gc + k, m});
                        if (mirror) M_.push_back({gc + k, gr + j, m});
                        MDK_.push_back({gr + j, gc + k, mdk});
                        if (mirror) MDK_.push_back({gc + k, gr + j, mdk});
                    }
            };
            for (int v = 0; v < 3; ++v) put(5 * v, 5 * v, 3, 3, idx[v], idx[v], false);                  // fillxMI :414-421
            const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
            for (int p = 0; p < 3; ++p) put(5 * pr[p][0], 5 * pr[p][1], 3, 3, idx[pr[p][0]], idx[pr[p][1]], true);   // fillxxMI :423-429
            for (int v = 0; v < 3; ++v)                                                                  // fillXMI :431-444
                if (eolv[v]) put(5 * v + 3, 5 * v + 3, 2, 2, Xdof(nodes[v]), Xdof(nodes[v]), false);
            for (int p = 0; p < 3; ++p)                                                                  // fillXXMI :446-458
                if (eolv[pr[p][0]] && eolv[pr[p][1]])
                    put(5 * pr[p][0] + 3, 5 * pr[p][1] + 3, 2, 2, Xdof(nodes[pr[p][0]]), Xdof(nodes[pr[p][1]]), true);
            for (int v = 0; v < 3; ++v) {                                                                // :460-496
                if (!eolv[v]) continue;
                for (int w = v; w < 3; ++w) put(5 * v + 3, 5 * w, 2, 3, Xdof(nodes[v]), idx[w], true);   // fillXxMI
                for (int u = 0; u < v; ++u) put(5 * u, 5 * v + 3, 3, 2, idx[u], Xdof(nodes[v]), true);   // fillxXMI
            }
            continue;
        }
        for (int v = 0; v < 3; ++v)                                           // :500-502
            for (int j = 0; j < 3; ++j) fadd(idx[v] + j, fm[3 * v + j] + fi[3 * v + j]);
        const double dhh = dampingB * h * h;                                  // damping(1)*h*h, :105
        auto Kme = [&](int r, int c) { return Km[c * 9 + r]; };
        auto Mie = [&](int r, int c) { return Mi[c * 9 + r]; };
        // fillxMI  :103-112   (diag blocks a, b, c  :504-510)
        for (int v = 0; v < 3; ++v)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) {
                    double m = Mie(3 * v + j, 3 * v + k);
                    double mdk = m + dhh * Kme(3 * v + j, 3 * v + k);
                    M_.push_back({idx[v] + j, idx[v] + k, m});
                    MDK_.push_back({idx[v] + j, idx[v] + k, mdk});
                }
        // fillxxMI :114-125   (off-diag (a,b), (a,c), (b,c)  :512-517)
        const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
        for (int p = 0; p < 3; ++p) {
            int v0 = pr[p][0], v1 = pr[p][1];
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) {
                    double m = Mie(3 * v0 + j, 3 * v1 + k);
                    double mdk = m + dhh * Kme(3 * v0 + j, 3 * v1 + k);
                    M_.push_back({idx[v0] + j, idx[v1] + k, m});
                    M_.push_back({idx[v1] + k, idx[v0] + j, m});
                    MDK_.push_back({idx[v0] + j, idx[v1] + k, mdk});
                    MDK_.push_back({idx[v1] + k, idx[v0] + j, mdk});
                }
        }
    }
    };

    // ---- edgeBasedF, Forces.cpp:685-910 ----
    auto do_edges = [&](int e_begin, int e_end, std::vector<Trip> &MDK_, FLog *fl, bool &bad) {
    auto fadd = [&](int at, double v) { if (fl) fl->push_back({at, v}); else R->f[at] += v; };
    for (int ed = e_begin; ed < e_end; ++ed) {
        const int32_t *s = edge_stencil + 4 * ed;
        if (s[2] < 0 || s[3] < 0) continue;                                   // :688-690
        double Wb[1], fb[12], Kb[144];
        ComputeBending(x + 3 * s[0], x + 3 * s[1], x + 3 * s[2], x + 3 * s[3],
                       X + 2 * s[0], X + 2 * s[1], X + 2 * s[2], X + 2 * s[3], beta, Wb, fb, Kb);  // :741
        const double dhh = dampingB * h * h;
        auto Kbe = [&](int r, int c) { return Kb[c * 12 + r]; };              // :744
        const int idx[4] = {3 * s[0], 3 * s[1], 3 * s[2], 3 * s[3]};
        const bool eolv[4] = {is_eol(s[0]), is_eol(s[1]), is_eol(s[2]), is_eol(s[3])};
        if (eolv[0] || eolv[1] || eolv[2] || eolv[3]) {                       // :746
            // F = (deform_grad(adjf[0]) + deform_grad(adjf[1])) / 2, each in its face's own vertex order (:590-596)
            double F1[6], F2[6], Fg[6];
            const int f0 = find_face(s[0], s[1], s[2]), f1 = find_face(s[0], s[1], s[3]);
            if (f0 < 0 || f1 < 0) { bad = true; return; }
            const int32_t *a0 = face_nodes + 3 * f0, *a1 = face_nodes + 3 * f1;
            deform_grad(x + 3 * a0[0], x + 3 * a0[1], x + 3 * a0[2], X + 2 * a0[0], X + 2 * a0[1], X + 2 * a0[2], F1);
            deform_grad(x + 3 * a1[0], x + 3 * a1[1], x + 3 * a1[2], X + 2 * a1[0], X + 2 * a1[1], X + 2 * a1[2], F2);
            for (int q = 0; q < 6; ++q) Fg[q] = (F1[q] + F2[q]) / 2;
            double fbe[20], Kbx[400];
            expand_eol(4, fb, Kbe, Fg, eolv, fbe, Kbx);                       // fillEOLBending :748
            auto KX = [&](int r, int c) { return Kbx[c * 20 + r]; };
            for (int v = 0; v < 4; ++v) {                                     // :750-760
                for (int j = 0; j < 3; ++j) fadd(idx[v] + j, fbe[5 * v + j]);
                if (eolv[v]) for (int j = 0; j < 2; ++j) fadd(Xdof(s[v]) + j, fbe[5 * v + 3 + j]);
            }
            auto put = [&](int lr, int lc, int nr, int nc, int gr, int gc, bool mirror) {   // K?? = damping(1)*h*h*Kbe.block, fill?B
                for (int j = 0; j < nr; ++j)
                    for (int k = 0; k < nc; ++k) {
                        const double kv = dhh * KX(lr + j, lc + k);
                        MDK_.push_back({gr + j, gc + k, kv});
                        if (mirror) MDK_.push_back({gc + k, gr + j, kv});
                    }
            };
            const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
            for (int v = 0; v < 4; ++v) put(5 * v, 5 * v, 3, 3, idx[v], idx[v], false);                  // :762-770
            for (int p = 0; p < 6; ++p) put(5 * pr[p][0], 5 * pr[p][1], 3, 3, idx[pr[p][0]], idx[pr[p][1]], true);   // :772-783
            for (int v = 0; v < 4; ++v)                                                                  // fillXB :785-801
                if (eolv[v]) put(5 * v + 3, 5 * v + 3, 2, 2, Xdof(s[v]), Xdof(s[v]), false);
            for (int p = 0; p < 6; ++p)                                                                  // fillXXB :803-826
                if (eolv[pr[p][0]] && eolv[pr[p][1]])
                    put(5 * pr[p][0] + 3, 5 * pr[p][1] + 3, 2, 2, Xdof(s[pr[p][0]]), Xdof(s[pr[p][1]]), true);
            for (int v = 0; v < 4; ++v) {                                                                // :828-882
                if (!eolv[v]) continue;
                for (int w = v; w < 4; ++w) put(5 * v + 3, 5 * w, 2, 3, Xdof(s[v]), idx[w], true);       // fillXxB
                for (int u = 0; u < v; ++u) put(5 * u, 5 * v + 3, 3, 2, idx[u], Xdof(s[v]), true);       // fillxXB
            }
            continue;
        }
        for (int v = 0; v < 4; ++v)                                           // fillxB :886-893
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k)
                    MDK_.push_back({idx[v] + j, idx[v] + k, dhh * Kbe(3 * v + j, 3 * v + k)});
        const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};  // fillxxB :895-906
        for (int p = 0; p < 6; ++p) {
            int v0 = pr[p][0], v1 = pr[p][1];
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) {
                    double kv = dhh * Kbe(3 * v0 + j, 3 * v1 + k);
                    MDK_.push_back({idx[v0] + j, idx[v1] + k, kv});
                    MDK_.push_back({idx[v1] + k, idx[v0] + j, kv});
                }
        }
    }
    };
    const int nthreads = std::max(1, (flags >> 8) & 0xff);
    if (nthreads == 1) {
        bool bad = false;
        do_faces(0, F, M_, MDK_, nullptr);
        do_edges(0, E, MDK_, nullptr, bad);
        if (bad) { delete R; return nullptr; }
    } else {
        struct Part { std::vector<Trip> M, K; FLog f; bool bad = false; };
        std::vector<Part> pf(nthreads), pe(nthreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                // sized for the Lagrangian branch (81 + 81 triplets per face, 144 per stencil); EOL elements grow the vectors a little
                pf[t].M.reserve((size_t)81 * (F / nthreads + 1)); pf[t].K.reserve((size_t)81 * (F / nthreads + 1));
                pe[t].K.reserve((size_t)144 * (E / nthreads + 1));
                do_faces((int)((int64_t)F * t / nthreads), (int)((int64_t)F * (t + 1) / nthreads), pf[t].M, pf[t].K, &pf[t].f);
                do_edges((int)((int64_t)E * t / nthreads), (int)((int64_t)E * (t + 1) / nthreads), pe[t].K, &pe[t].f, pe[t].bad);
            });
        for (auto &t : th) t.join();
        for (int t = 0; t < nthreads; ++t) if (pe[t].bad) { delete R; return nullptr; }
        // the reference's order: every face, then every edge
        {
            size_t nm = 0, nk = 0;
            for (int t = 0; t < nthreads; ++t) { nm += pf[t].M.size(); nk += pf[t].K.size() + pe[t].K.size(); }
            M_.reserve(nm); MDK_.reserve(nk);
        }
        for (int pass = 0; pass < 2; ++pass)
            for (int t = 0; t < nthreads; ++t) {
                Part &P = pass ? pe[t] : pf[t];
                M_.insert(M_.end(), P.M.begin(), P.M.end());
                MDK_.insert(MDK_.end(), P.K.begin(), P.K.end());
                for (const auto &c : P.f) R->f[c.first] += c.second;
                Part().M.swap(P.M); Part().K.swap(P.K);
            }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (!(flags & 1)) {
        if (nthreads == 1) {
            set_from_triplets(dof, M_, R->M);                                 // :928
            set_from_triplets(dof, MDK_, R->MDK);                             // :929
        } else {                                                              // timing variant: the two matrices side by side
            std::thread tm([&]() { set_from_triplets(dof, M_, R->M); });
            set_from_triplets(dof, MDK_, R->MDK);
            tm.join();
        }
    }
    auto t2 = std::chrono::steady_clock::now();
    R->seconds_elements = std::chrono::duration<double>(t1 - t0).count();
    R->seconds_assembly = std::chrono::duration<double>(t2 - t1).count();
    return R;
}

void *oracle_forces_fill(int N, int F, const int32_t *face_nodes, int E, const int32_t *edge_stencil,
                         const double *x, const double *X, const double *mat, const double *grav, double h,
                         int flags) {
    return oracle_forces_fill_eol(N, F, face_nodes, E, edge_stencil, x, X, mat, grav, h, nullptr, flags);
}

int oracle_forces_dof(void *r) { return ((Result *)r)->dof; }
const double *oracle_forces_f(void *r) { return ((Result *)r)->f.data(); }
int64_t oracle_forces_nnz(void *r, int which) { return (which ? ((Result *)r)->MDK : ((Result *)r)->M).nnz; }
const int32_t *oracle_forces_outer(void *r, int which) { return (which ? ((Result *)r)->MDK : ((Result *)r)->M).outer.data(); }
const int32_t *oracle_forces_inner(void *r, int which) { return (which ? ((Result *)r)->MDK : ((Result *)r)->M).inner.data(); }
const double *oracle_forces_vals(void *r, int which) { return (which ? ((Result *)r)->MDK : ((Result *)r)->M).vals.data(); }
double oracle_forces_seconds(void *r, int which) { return which ? ((Result *)r)->seconds_assembly : ((Result *)r)->seconds_elements; }
void oracle_forces_free(void *r) { delete (Result *)r; }

// thin pass-throughs so tests can pin the reference kernels against SURVEY §8c known answers
void oracle_compute_membrane(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                             const double *Xc, double e, double nu, const double *P, const double *Q, double *W,
                             double *f, double *K) { ComputeMembrane(xa, xb, xc, Xa, Xb, Xc, e, nu, P, Q, W, f, K); }
void oracle_compute_bending(const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                            const double *X1, const double *X2, const double *X3, double beta, double *W, double *f,
                            double *K) { ComputeBending(x0, x1, x2, x3, X0, X1, X2, X3, beta, W, f, K); }
void oracle_compute_inertial(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                             const double *Xc, const double *g, double rho, double *W, double *f, double *M) {
    ComputeInertial(xa, xb, xc, Xa, Xb, Xc, g, rho, W, f, M);
}
// frame + polar decomposition of one face, for unit tests (Forces.cpp:357-372)
void oracle_face_frame(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                       const double *Xc, double *PP, double *QQ) {
    double d1[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};
    double d2[3] = {xc[0] - xa[0], xc[1] - xa[1], xc[2] - xa[2]};
    double DX[4] = {Xb[0] - Xa[0], Xb[1] - Xa[1], Xc[0] - Xa[0], Xc[1] - Xa[1]};
    double normm[3]; cross3(normm, d1, d2);
    double l1 = std::sqrt(dot3(d1, d1));
    double Pxm[3] = {d1[0] / l1, d1[1] / l1, d1[2] / l1};
    double Pym[3]; cross3(Pym, normm, Pxm);
    double l2 = std::sqrt(dot3(Pym, Pym));
    Pym[0] /= l2; Pym[1] /= l2; Pym[2] /= l2;
    double det = DX[0] * DX[3] - DX[2] * DX[1];
    double inv[4] = {DX[3] / det, -DX[1] / det, -DX[2] / det, DX[0] / det};
    double Fm[6];
    for (int r = 0; r < 3; ++r) { Fm[r] = d1[r] * inv[0] + d2[r] * inv[1]; Fm[3 + r] = d1[r] * inv[2] + d2[r] * inv[3]; }
    double Fb[4] = {dot3(Pxm, Fm), dot3(Pym, Fm), dot3(Pxm, Fm + 3), dot3(Pym, Fm + 3)};
    poldec(Fb, QQ);
    PP[0] = Pxm[0]; PP[1] = Pym[0]; PP[2] = Pxm[1]; PP[3] = Pym[1]; PP[4] = Pxm[2]; PP[5] = Pym[2];
}

}  // extern "C"
