// elements.cuh — per-element FP64 arithmetic of the Forces::fill hot path, hand-derived.
//
// Replaces (same results to rounding, far fewer operations, no spills):
//   faceBasedF frame + poldec       /root/reference/src/Forces.cpp:357-372, :33-52
//   ComputeMembrane                 /root/reference/src/ComputeMembrane.cpp:17-319   (~1040 flop -> ~250)
//   ComputeInertial                 /root/reference/src/ComputeInertial.cpp:15-135
//   ComputeBending (K only)         /root/reference/src/ComputeBending.cpp:30-737    (~6500 flop -> ~900)
//
// The generated reference code is the exact derivative of closed-form energies, so it can be
// re-derived in matrix form:
//
//  membrane:  R = Q^T P (2x3), F = Dx DX^-1 (3x2), E = R F - I, S = 2 mu E + lambda tr(E) I,
//             W = A (mu |E|^2 + lambda/2 tr(E)^2),   f_i = -A R^T S gradN_i,
//             K_ij = A ( 2 mu (gradN_i . gradN_j) R^T R + lambda (R^T gradN_i)(R^T gradN_j)^T )
//             with A = t7/2 (signed rest area), mu = e/(2(1+nu)), lambda = e nu/((1+nu)(1-2nu)).
//  inertial:  t8 = rho * t7 ; f_i = t8 g / 6 ; M_ii = t8/12 I ; M_ij = t8/24 I.
//  bending:   W = c (1 - u.v), c = 3/2 beta |X1-X0|^2 / (A0+A1), u = n0/|n0|, v = n1/|n1|,
//             n0 = (x1-x0) x (x2-x0), n1 = (x3-x0) x (x1-x0).  With D = u.v,
//             w0 = (x2-x1, x0-x2, x1-x0, 0), w1 = (x1-x3, x3-x0, 0, x0-x1)  (dn0/dx_i = [w0_i]x, dn1/dx_i = [w1_i]x),
//             U_i = u x w0_i, Y_i = v x w0_i, V_i = v x w1_i, Z_i = u x w1_i :
//             Hess(D)_ij = -1/|n0|^2 [ U_i Y_j^T + Y_i U_j^T - 3D U_i U_j^T + D((w0_i.w0_j) I - w0_j w0_i^T) ]
//                          -1/|n1|^2 [ V_i Z_j^T + Z_i V_j^T - 3D V_i V_j^T + D((w1_i.w1_j) I - w1_j w1_i^T) ]
//                          +1/(|n0||n1|) [ (w0_i.w1_j) I - w1_j w0_i^T - Y_i V_j^T - U_i (Z_j - D V_j)^T ]
//                          +1/(|n0||n1|) [ (w1_i.w0_j) I - w0_j w1_i^T - V_i Y_j^T - (Z_i - D V_i) U_j^T ]
//                          + k0_ij [g0]x + k1_ij [g1]x ,   g0 = (v - D u)/|n0|, g1 = (u - D v)/|n1|,
//             K = -c Hess(D).   Verified against the reference object code to <1e-15 of the block scale
//             (tests/test_element_math.py).
//
// The same source compiles for the device (nvcc) and for the host unit-test shim (g++), via EOLC_HD.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define EOLC_HD __host__ __device__ __forceinline__
#else
#define EOLC_HD inline
#endif

namespace eolc {

// 1/sqrt(x): one MUFU + Newton steps on the device (<= 2 ulp, far inside the 1e-10 budget) instead of sqrt + divide
EOLC_HD double rsq(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// ---- the reference's rest-area expressions, with the reference's roundings ------------------------------------------------------
// ComputeMembrane.cpp:43 / ComputeInertial.cpp:33 (t7, twice the signed rest area of a face) and ComputeBending.cpp:53 (A0 + A1 of
// a stencil) are generated as sums of products of ABSOLUTE material coordinates: a small difference of O(|X|^2) terms.  On a fine
// mesh the rounding of every single product shows in the result at eps |X|^2 / (2A) — 3e-10 relative on the 1024^2 sheet — and
// from there in every entry of M, of the membrane K and f, and of the bending K.  Agreement with the reference to 1e-10 therefore
// needs the SAME roundings: each product rounded on its own (no FMA contraction) and the terms added left to right, exactly as the
// generated code does (g++ without -mfma contracts nothing).  All other arithmetic is well conditioned and may contract.
EOLC_HD double ref_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
EOLC_HD double ref_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// Xax * Xby - Xax * Xcy - Xbx * Xay + Xcx * Xay + Xbx * Xcy - Xcx * Xby
EOLC_HD double rest_area2(double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy) {
    double s = ref_add(ref_mul(Xax, Xby), -ref_mul(Xax, Xcy));
    s = ref_add(s, -ref_mul(Xbx, Xay));
    s = ref_add(s, ref_mul(Xcx, Xay));
    s = ref_add(s, ref_mul(Xbx, Xcy));
    return ref_add(s, -ref_mul(Xcx, Xby));
}
// -X0x * X2y / 2 + X2x * X0y / 2 + X1x * X2y / 2 - X2x * X1y / 2 + X0x * X3y / 2 - X3x * X0y / 2 - X1x * X3y / 2 + X3x * X1y / 2
// (the halvings are exact, so this is half the left-to-right sum of the signed products)
EOLC_HD double stencil_area(double X0x, double X0y, double X1x, double X1y, double X2x, double X2y, double X3x, double X3y) {
    double s = ref_add(-ref_mul(X0x, X2y), ref_mul(X2x, X0y));
    s = ref_add(s, ref_mul(X1x, X2y));
    s = ref_add(s, -ref_mul(X2x, X1y));
    s = ref_add(s, ref_mul(X0x, X3y));
    s = ref_add(s, -ref_mul(X3x, X0y));
    s = ref_add(s, -ref_mul(X1x, X3y));
    s = ref_add(s, ref_mul(X3x, X1y));
    return 0.5 * s;
}

struct v3 { double x, y, z; };
EOLC_HD v3 mk3(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
EOLC_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
EOLC_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
EOLC_HD v3 operator*(double s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
EOLC_HD double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
EOLC_HD v3 cross(v3 a, v3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// 3x3 block, row-major: B[3*r + c]
struct blk3 { double m[9]; };
// B += a b^T
EOLC_HD void add_outer(blk3 &B, v3 a, v3 b) {
    B.m[0] += a.x * b.x; B.m[1] += a.x * b.y; B.m[2] += a.x * b.z;
    B.m[3] += a.y * b.x; B.m[4] += a.y * b.y; B.m[5] += a.y * b.z;
    B.m[6] += a.z * b.x; B.m[7] += a.z * b.y; B.m[8] += a.z * b.z;
}
EOLC_HD void add_diag(blk3 &B, double d) { B.m[0] += d; B.m[4] += d; B.m[8] += d; }
// upper triangle only (entries 0,1,2,4,5,8) of B += a b^T — for diagonal blocks, whose SUM of terms is symmetric
EOLC_HD void add_outer_up(blk3 &B, v3 a, v3 b) {
    B.m[0] += a.x * b.x; B.m[1] += a.x * b.y; B.m[2] += a.x * b.z;
    B.m[4] += a.y * b.y; B.m[5] += a.y * b.z;
    B.m[8] += a.z * b.z;
}
EOLC_HD void mirror_up(blk3 &B) { B.m[3] = B.m[1]; B.m[6] = B.m[2]; B.m[7] = B.m[5]; }
// B += s [g]x
EOLC_HD void add_skew(blk3 &B, double s, v3 g) {
    B.m[1] -= s * g.z; B.m[2] += s * g.y; B.m[3] += s * g.z; B.m[5] -= s * g.x; B.m[6] -= s * g.y; B.m[7] += s * g.x;
}

struct FaceOut {
    double fa[3], fb[3], fc[3];   // fm + fi per vertex (Forces.cpp:500-502)
    double t8;                    // rho * 2A(signed): M_ii = t8/12, M_ij = t8/24 (ComputeInertial.cpp:33,44-47)
    // MDK contribution blocks Mi + dhh*Km: order aa, bb, cc, ab, ac, bc  (Forces.cpp:504-517)
    blk3 K[6];
};

// One triangle: frame, polar decomposition, membrane f/K, inertial f/M.
EOLC_HD void face_element(v3 xa, v3 xb, v3 xc, double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy,
                          double e, double nu, double rho, v3 g, double dhh, FaceOut &o) {
    // frame (Forces.cpp:357-366)
    v3 d1 = xb - xa, d2 = xc - xa;
    v3 nrm = cross(d1, d2);
    double il1 = 1.0 / sqrt(dot(d1, d1));
    v3 Px = il1 * d1;
    v3 Py = cross(nrm, Px);
    double il2 = 1.0 / sqrt(dot(Py, Py));
    Py = il2 * Py;
    // rest-shape gradients (ComputeMembrane.cpp:43,52-70,101-110): rows of DX^-1
    double t7 = rest_area2(Xax, Xay, Xbx, Xby, Xcx, Xcy);
    double t17 = 1.0 / t7;
    double gb0 = t17 * (Xcy - Xay), gb1 = t17 * (Xax - Xcx);
    double gc0 = t17 * (Xay - Xby), gc1 = t17 * (Xbx - Xax);
    double ga0 = -gb0 - gc0, ga1 = -gb1 - gc1;
    // F = Dx DX^-1 (3x2) and Fbar = P F (Forces.cpp:367-368)
    v3 F0 = gb0 * d1 + gc0 * d2, F1 = gb1 * d1 + gc1 * d2;
    double m11 = dot(Px, F0), m21 = dot(Py, F0), m12 = dot(Px, F1), m22 = dot(Py, F1);
    // poldec (Forces.cpp:33-52)
    double detM = m11 * m22 - m12 * m21;
    double sg = detM < 0.0 ? -1.0 : (detM == 0.0 ? 0.0 : 1.0);
    double q00 = m11 + sg * m22, q01 = m12 - sg * m21, q10 = m21 - sg * m12, q11 = m22 + sg * m11;
    double icl = 1.0 / sqrt(q00 * q00 + q10 * q10);
    q00 *= icl; q01 *= icl; q10 *= icl; q11 *= icl;
    // R = Q^T P : rows r0, r1 (ComputeMembrane.cpp:47-69: t15,t29,t40 / t53,t57,t61)
    v3 r0 = q00 * Px + q10 * Py, r1 = q01 * Px + q11 * Py;
    // E = R F - I
    double E00 = dot(r0, F0) - 1.0, E10 = dot(r1, F0), E01 = dot(r0, F1), E11 = dot(r1, F1) - 1.0;
    double mu = e / (1.0 + nu) * 0.5;                              // t12
    double lam = e * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));         // t88*t93
    double A = 0.5 * t7;                                           // t8 of ComputeMembrane
    double tr = E00 + E11;
    double S00 = 2.0 * mu * E00 + lam * tr, S01 = 2.0 * mu * E01, S10 = 2.0 * mu * E10, S11 = 2.0 * mu * E11 + lam * tr;
    // f_i = -A R^T S gradN_i  + gravity  t8 g/6 (ComputeInertial.cpp:33-37)
    double t8 = rho * t7;
    v3 fg = (t8 / 6.0) * g;
    {
        v3 fa = (-A * (S00 * ga0 + S01 * ga1)) * r0 + (-A * (S10 * ga0 + S11 * ga1)) * r1;
        v3 fb = (-A * (S00 * gb0 + S01 * gb1)) * r0 + (-A * (S10 * gb0 + S11 * gb1)) * r1;
        v3 fc = (-A * (S00 * gc0 + S01 * gc1)) * r0 + (-A * (S10 * gc0 + S11 * gc1)) * r1;
        o.fa[0] = fa.x + fg.x; o.fa[1] = fa.y + fg.y; o.fa[2] = fa.z + fg.z;
        o.fb[0] = fb.x + fg.x; o.fb[1] = fb.y + fg.y; o.fb[2] = fb.z + fg.z;
        o.fc[0] = fc.x + fg.x; o.fc[1] = fc.y + fg.y; o.fc[2] = fc.z + fg.z;
    }
    o.t8 = t8;
    // K_ij = A(2mu (gi.gj) R^T R + lam q_i q_j^T),  MDK block = M_ij + dhh K_ij
    double RR[6] = {r0.x * r0.x + r1.x * r1.x, r0.x * r0.y + r1.x * r1.y, r0.x * r0.z + r1.x * r1.z,
                    r0.y * r0.y + r1.y * r1.y, r0.y * r0.z + r1.y * r1.z, r0.z * r0.z + r1.z * r1.z};
    v3 qa = ga0 * r0 + ga1 * r1, qb = gb0 * r0 + gb1 * r1, qc = gc0 * r0 + gc1 * r1;
    const double a2mu = dhh * A * 2.0 * mu, alam = dhh * A * lam;
    const double md = t8 / 12.0, mo = t8 / 24.0;
    auto block = [&](blk3 &B, double gg, v3 qi, v3 qj, double mass) {
        double s = a2mu * gg;
        v3 ql = alam * qi;
        B.m[0] = mass + (s * RR[0] + ql.x * qj.x); B.m[1] = s * RR[1] + ql.x * qj.y; B.m[2] = s * RR[2] + ql.x * qj.z;
        B.m[3] = s * RR[1] + ql.y * qj.x; B.m[4] = mass + (s * RR[3] + ql.y * qj.y); B.m[5] = s * RR[4] + ql.y * qj.z;
        B.m[6] = s * RR[2] + ql.z * qj.x; B.m[7] = s * RR[4] + ql.z * qj.y; B.m[8] = mass + (s * RR[5] + ql.z * qj.z);
    };
    block(o.K[0], ga0 * ga0 + ga1 * ga1, qa, qa, md); mirror_up(o.K[0]);   // diagonal blocks exactly symmetric,
    block(o.K[1], gb0 * gb0 + gb1 * gb1, qb, qb, md); mirror_up(o.K[1]);   // like the reference's copies
    block(o.K[2], gc0 * gc0 + gc1 * gc1, qc, qc, md); mirror_up(o.K[2]);   // (ComputeMembrane.cpp:182,202-203,...)
    block(o.K[3], ga0 * gb0 + ga1 * gb1, qa, qb, mo);
    block(o.K[4], ga0 * gc0 + ga1 * gc1, qa, qc, mo);
    block(o.K[5], gb0 * gc0 + gb1 * gc1, qb, qc, mo);
}

// rho * 2A(signed) of one face, for the mass matrix (ComputeInertial.cpp:33)
EOLC_HD double face_t8(double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy, double rho) {
    return rho * rest_area2(Xax, Xay, Xbx, Xby, Xcx, Xcy);
}

// One interior edge: the 10 upper 3x3 blocks of dhh * Kb, order 00,11,22,33,01,02,03,12,13,23
// (Forces.cpp:885-906).  The bending force is not produced: the non-EOL branch discards it.
struct EdgeOut { blk3 K[10]; };

EOLC_HD void edge_element(v3 x0, v3 x1, v3 x2, v3 x3, double X0x, double X0y, double X1x, double X1y, double X2x,
                          double X2y, double X3x, double X3y, double beta, double dhh, EdgeOut &o) {
    // c = 3/2 * t6 * t17 (ComputeBending.cpp:50-53,102)
    double ex = X1x - X0x, ey = X1y - X0y;
    double t6 = beta * (ex * ex + ey * ey);
    double den = stencil_area(X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y);
    double c = 1.5 * t6 / den;
    v3 e = x1 - x0, a = x2 - x0, b = x3 - x0;
    v3 n0 = cross(e, a), n1 = cross(b, e);
    double s0 = 1.0 / dot(n0, n0), s1 = 1.0 / dot(n1, n1);
    double il0 = sqrt(s0), il1 = sqrt(s1);
    v3 u = il0 * n0, v = il1 * n1;
    double D = dot(u, v);
    const double kk = -c * dhh;                  // K*dhh = kk * Hess(D)
    const double k0 = -kk * s0, k1 = -kk * s1, k01 = kk * il0 * il1;
    // w vectors
    v3 w00 = x2 - x1, w01 = x0 - x2, w02 = e;
    v3 w10 = x1 - x3, w11 = b, w13 = x0 - x1;
    // crosses
    v3 U0 = cross(u, w00), U1 = cross(u, w01), U2 = cross(u, w02);
    v3 Y0 = cross(v, w00), Y1 = cross(v, w01), Y2 = cross(v, w02);
    v3 V0 = cross(v, w10), V1 = cross(v, w11), V3 = cross(v, w13);
    v3 Z0 = cross(u, w10), Z1 = cross(u, w11), Z3 = cross(u, w13);
    // second-order vectors kk*g0, kk*g1
    v3 g0 = (kk * il0) * (v - D * u), g1 = (kk * il1) * (u - D * v);

    // --- triangle-0 term: k0 [ U_i T_j^T + Y_i U_j^T + D (w0i.w0j) I - D w0j w0i^T ],  T_j = Y_j - 3D U_j
    // --- triangle-1 term: k1 [ V_i T'_j^T + Z_i V_j^T + D (w1i.w1j) I - D w1j w1i^T ],  T'_j = Z_j - 3D V_j
    // --- cross terms:     k01 [ (w0i.w1j) I - w1j w0i^T - Y_i V_j^T - U_i ZZ_j^T ]  (+ transpose of (j,i)),  ZZ_j = Z_j - D V_j
    const double D3 = 3.0 * D;
    v3 T0 = Y0 - D3 * U0, T1 = Y1 - D3 * U1, T2 = Y2 - D3 * U2;
    v3 Tp0 = Z0 - D3 * V0, Tp1 = Z1 - D3 * V1, Tp3 = Z3 - D3 * V3;
    v3 ZZ0 = Z0 - D * V0, ZZ1 = Z1 - D * V1, ZZ3 = Z3 - D * V3;
    // pre-scaled left factors
    v3 kU0 = k0 * U0, kU1 = k0 * U1, kU2 = k0 * U2, kY0 = k0 * Y0, kY1 = k0 * Y1, kY2 = k0 * Y2;
    v3 kV0 = k1 * V0, kV1 = k1 * V1, kV3 = k1 * V3, kZ0 = k1 * Z0, kZ1 = k1 * Z1, kZ3 = k1 * Z3;
    const double k0D = k0 * D, k1D = k1 * D;
    v3 dw00 = k0D * w00, dw01 = k0D * w01, dw02 = k0D * w02;
    v3 dw10 = k1D * w10, dw11 = k1D * w11, dw13 = k1D * w13;
    v3 cY0 = k01 * Y0, cY1 = k01 * Y1, cY2 = k01 * Y2, cU0 = k01 * U0, cU1 = k01 * U1, cU2 = k01 * U2;
    v3 cw00 = k01 * w00, cw01 = k01 * w01, cw02 = k01 * w02;

#define EOLC_ZERO(B) { for (int q_ = 0; q_ < 9; ++q_) (B).m[q_] = 0.0; }
    // tri0(i,j): B += kU_i T_j^T + kY_i U_j^T + (dw0i.w0j) I - w0j dw0i^T
#define EOLC_TRI0(B, kUi, kYi, dwi, Tj, Uj, wj) { add_outer(B, kUi, Tj); add_outer(B, kYi, Uj); add_diag(B, dot(dwi, wj)); add_outer(B, -1.0 * (wj), dwi); }
#define EOLC_TRI1(B, kVi, kZi, dwi, Tpj, Vj, wj) { add_outer(B, kVi, Tpj); add_outer(B, kZi, Vj); add_diag(B, dot(dwi, wj)); add_outer(B, -1.0 * (wj), dwi); }
    // cross01(i,j): B += (cw0i.w1j) I - w1j cw0i^T - cY_i V_j^T - cU_i ZZ_j^T
#define EOLC_X01(B, cwi, cYi, cUi, w1j, Vj, ZZj) { add_diag(B, dot(cwi, w1j)); add_outer(B, -1.0 * (w1j), cwi); add_outer(B, -1.0 * (cYi), Vj); add_outer(B, -1.0 * (cUi), ZZj); }
    // cross10(i,j) = cross01(j,i)^T: B += (w1i.cw0j) I - cw0j w1i^T - V_i cY_j^T - ZZ_i cU_j^T
#define EOLC_X10(B, w1i, Vi, ZZi, cwj, cYj, cUj) { add_diag(B, dot(w1i, cwj)); add_outer(B, -1.0 * (cwj), w1i); add_outer(B, -1.0 * (Vi), cYj); add_outer(B, -1.0 * (ZZi), cUj); }

    // diagonal blocks: every bracket above is symmetric for i == j (X01 + X10 = X + X^T), so only the upper triangle
    // is accumulated and then mirrored — exactly symmetric like the reference's copied entries, and 1/3 fewer FMAs.
#define EOLC_TRI_UP(B, kAi, kBi, dwi, Tj, Aj, wj) { add_outer_up(B, kAi, Tj); add_outer_up(B, kBi, Aj); add_diag(B, dot(dwi, wj)); add_outer_up(B, -1.0 * (wj), dwi); }
    // X01(i,i) + X10(i,i) = 2 (cw.w1) I - (w1 cw^T + cw w1^T) - (cY V^T + V cY^T) - (cU ZZ^T + ZZ cU^T)
#define EOLC_XX_UP(B, cwi, cYi, cUi, w1i, Vi, ZZi) { add_diag(B, 2.0 * dot(cwi, w1i)); \
        add_outer_up(B, -1.0 * (w1i), cwi); add_outer_up(B, -1.0 * (cwi), w1i); add_outer_up(B, -1.0 * (cYi), Vi); add_outer_up(B, -1.0 * (Vi), cYi); \
        add_outer_up(B, -1.0 * (cUi), ZZi); add_outer_up(B, -1.0 * (ZZi), cUi); }
    blk3 B;
    // (0,0)
    EOLC_ZERO(B); EOLC_TRI_UP(B, kU0, kY0, dw00, T0, U0, w00); EOLC_TRI_UP(B, kV0, kZ0, dw10, Tp0, V0, w10);
    EOLC_XX_UP(B, cw00, cY0, cU0, w10, V0, ZZ0); mirror_up(B);
    o.K[0] = B;
    // (1,1)
    EOLC_ZERO(B); EOLC_TRI_UP(B, kU1, kY1, dw01, T1, U1, w01); EOLC_TRI_UP(B, kV1, kZ1, dw11, Tp1, V1, w11);
    EOLC_XX_UP(B, cw01, cY1, cU1, w11, V1, ZZ1); mirror_up(B);
    o.K[1] = B;
    // (2,2)
    EOLC_ZERO(B); EOLC_TRI_UP(B, kU2, kY2, dw02, T2, U2, w02); mirror_up(B);
    o.K[2] = B;
    // (3,3)
    EOLC_ZERO(B); EOLC_TRI_UP(B, kV3, kZ3, dw13, Tp3, V3, w13); mirror_up(B);
    o.K[3] = B;
#undef EOLC_TRI_UP
#undef EOLC_XX_UP
    // (0,1): second-order  -G0 + G1
    EOLC_ZERO(B); EOLC_TRI0(B, kU0, kY0, dw00, T1, U1, w01); EOLC_TRI1(B, kV0, kZ0, dw10, Tp1, V1, w11);
    EOLC_X01(B, cw00, cY0, cU0, w11, V1, ZZ1); EOLC_X10(B, w10, V0, ZZ0, cw01, cY1, cU1);
    add_skew(B, -1.0, g0); add_skew(B, 1.0, g1);
    o.K[4] = B;
    // (0,2): +G0
    EOLC_ZERO(B); EOLC_TRI0(B, kU0, kY0, dw00, T2, U2, w02); EOLC_X10(B, w10, V0, ZZ0, cw02, cY2, cU2);
    add_skew(B, 1.0, g0);
    o.K[5] = B;
    // (0,3): -G1
    EOLC_ZERO(B); EOLC_TRI1(B, kV0, kZ0, dw10, Tp3, V3, w13); EOLC_X01(B, cw00, cY0, cU0, w13, V3, ZZ3);
    add_skew(B, -1.0, g1);
    o.K[6] = B;
    // (1,2): -G0
    EOLC_ZERO(B); EOLC_TRI0(B, kU1, kY1, dw01, T2, U2, w02); EOLC_X10(B, w11, V1, ZZ1, cw02, cY2, cU2);
    add_skew(B, -1.0, g0);
    o.K[7] = B;
    // (1,3): +G1
    EOLC_ZERO(B); EOLC_TRI1(B, kV1, kZ1, dw11, Tp3, V3, w13); EOLC_X01(B, cw01, cY1, cU1, w13, V3, ZZ3);
    add_skew(B, 1.0, g1);
    o.K[8] = B;
    // (2,3): cross term only
    EOLC_ZERO(B); EOLC_X01(B, cw02, cY2, cU2, w13, V3, ZZ3);
    o.K[9] = B;
#undef EOLC_ZERO
#undef EOLC_TRI0
#undef EOLC_TRI1
#undef EOLC_X01
#undef EOLC_X10
}

// ------------------------------------------------------------------------------------------------
// Tile forms ("tiles" pipeline): each element is evaluated ONCE per tile and parked in shared memory in a compact layout
// (symmetric diagonal blocks as 6 doubles, off-diagonal blocks as 8 + 1 doubles, 16-byte aligned).
//
//  bending, reduced coordinates.  D depends on x only through e = x1-x0, a = x2-x0, b = x3-x0 (translation invariance),
//  n0 = e x a, n1 = b x e.  With H00, H01, H11 the blocks of the Hessian of D with respect to (n0, n1)
//       H00 = -1/|n0|^2 (q u^T + u q^T + D (I - u u^T)),  q = v - D u         (symmetric)
//       H11 = -1/|n1|^2 (r v^T + v r^T + D (I - v v^T)),  r = u - D v         (symmetric)
//       H01 =  1/(|n0||n1|) (I - u u^T - v v^T + D u v^T)
//  and dn0 = -[a]x de + [e]x da, dn1 = [b]x de - [e]x db, d2n0 = 2 de x da, d2n1 = 2 db x de, g0 = q/|n0|, g1 = r/|n1|:
//       H_aa = -[e]x H00 [e]x      H_ab = [e]x H01 [e]x      H_bb = -[e]x H11 [e]x
//       H_ea =  [a]x H00 [e]x - [b]x H01^T [e]x - [g0]x      H_eb = -[a]x H01 [e]x + [b]x H11 [e]x + [g1]x
//       H_ee = -[a]x H00 [a]x + S + S^T - [b]x H11 [b]x,     S = [a]x H01 [b]x
//  K_11 = H_ee, K_12 = H_ea, K_13 = H_eb, K_22 = H_aa, K_23 = H_ab, K_33 = H_bb and, because every row of K sums to zero,
//       K_0j = -(K_1j + K_2j + K_3j),  K_00 = -(K_01 + K_02 + K_03).
//  About 560 FP64 instructions instead of ~1000 for the outer-product form above; verified against the reference's
//  generated ComputeBending to < 1e-12 of the block scale (tests/test_element_math.py).
// ------------------------------------------------------------------------------------------------
struct sym3 { double xx, xy, xz, yy, yz, zz; };
// rows of M [w]x : M_r x w
EOLC_HD void mul_skew_r(const blk3 &M, v3 w, blk3 &R) {
    v3 r0 = cross(mk3(M.m[0], M.m[1], M.m[2]), w), r1 = cross(mk3(M.m[3], M.m[4], M.m[5]), w), r2 = cross(mk3(M.m[6], M.m[7], M.m[8]), w);
    R.m[0] = r0.x; R.m[1] = r0.y; R.m[2] = r0.z; R.m[3] = r1.x; R.m[4] = r1.y; R.m[5] = r1.z; R.m[6] = r2.x; R.m[7] = r2.y; R.m[8] = r2.z;
}
EOLC_HD void sym_mul_skew_r(const sym3 &S, v3 w, blk3 &R) {
    v3 r0 = cross(mk3(S.xx, S.xy, S.xz), w), r1 = cross(mk3(S.xy, S.yy, S.yz), w), r2 = cross(mk3(S.xz, S.yz, S.zz), w);
    R.m[0] = r0.x; R.m[1] = r0.y; R.m[2] = r0.z; R.m[3] = r1.x; R.m[4] = r1.y; R.m[5] = r1.z; R.m[6] = r2.x; R.m[7] = r2.y; R.m[8] = r2.z;
}
// B (+)= s * [w]x M   (columns: w x M_col), written so that every term contracts into one FMA
template <bool ACC>
EOLC_HD void skew_l(blk3 &B, double s, v3 w, const blk3 &M) {
    const v3 sw = s * w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double cx = M.m[c], cy = M.m[3 + c], cz = M.m[6 + c];
        double rx = ACC ? B.m[c] : 0.0, ry = ACC ? B.m[3 + c] : 0.0, rz = ACC ? B.m[6 + c] : 0.0;
        if (ACC) { rx += sw.y * cz; ry += sw.z * cx; rz += sw.x * cy; } else { rx = sw.y * cz; ry = sw.z * cx; rz = sw.x * cy; }
        rx -= sw.z * cy; ry -= sw.x * cz; rz -= sw.y * cx;
        B.m[c] = rx; B.m[3 + c] = ry; B.m[6 + c] = rz;
    }
}
EOLC_HD void add_skew_l(blk3 &B, double s, v3 w, const blk3 &M) { skew_l<true>(B, s, w, M); }
EOLC_HD void set_skew_l(blk3 &B, double s, v3 w, const blk3 &M) { skew_l<false>(B, s, w, M); }
// upper triangle of S (+)= s * [w]x M, for products known to be symmetric
template <bool ACC>
EOLC_HD void skew_l_sym(sym3 &S, double s, v3 w, const blk3 &M) {
    const v3 sw = s * w;
    if (ACC) {
        S.xx += sw.y * M.m[6]; S.xy += sw.y * M.m[7]; S.xz += sw.y * M.m[8]; S.yy += sw.z * M.m[1]; S.yz += sw.z * M.m[2]; S.zz += sw.x * M.m[5];
    } else {
        S.xx = sw.y * M.m[6]; S.xy = sw.y * M.m[7]; S.xz = sw.y * M.m[8]; S.yy = sw.z * M.m[1]; S.yz = sw.z * M.m[2]; S.zz = sw.x * M.m[5];
    }
    S.xx -= sw.z * M.m[3]; S.xy -= sw.z * M.m[4]; S.xz -= sw.z * M.m[5]; S.yy -= sw.x * M.m[7]; S.yz -= sw.x * M.m[8]; S.zz -= sw.y * M.m[2];
}
EOLC_HD void add_skew_l_sym(sym3 &S, double s, v3 w, const blk3 &M) { skew_l_sym<true>(S, s, w, M); }
EOLC_HD void set_skew_l_sym(sym3 &S, double s, v3 w, const blk3 &M) { skew_l_sym<false>(S, s, w, M); }

// emit.diag(i, sym3) for K_ii (i = 0..3), emit.off(k, blk3) for k = 0..5: K_01, K_02, K_03, K_12, K_13, K_23.
// All blocks are already multiplied by dhh = dampingB h^2 (Forces.cpp:885-906).
template <typename Emit>
EOLC_HD void edge_element_tile(v3 x0, v3 x1, v3 x2, v3 x3, double X0x, double X0y, double X1x, double X1y, double X2x, double X2y,
                               double X3x, double X3y, double beta, double dhh, Emit &emit) {
    // c = 3/2 t6 t17 (ComputeBending.cpp:50-53,102);  K dhh = kk Hess(D), kk = -c dhh
    const double ex = X1x - X0x, ey = X1y - X0y;
    const double t6 = beta * (ex * ex + ey * ey);
    const double den = stencil_area(X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y);
    const double kk = -(1.5 * t6 / den) * dhh;
    const v3 e = x1 - x0, a = x2 - x0, b = x3 - x0;
    const v3 n0 = cross(e, a), n1 = cross(b, e);
    const double il0 = rsq(dot(n0, n0)), il1 = rsq(dot(n1, n1));
    const v3 u = il0 * n0, v = il1 * n1;
    const double D = dot(u, v);
    const v3 q = v - D * u, r = u - D * v;
    const double c0 = -kk * il0 * il0, c1 = -kk * il1 * il1, c01 = kk * il0 * il1;
    sym3 H00, H11;
    {
        const v3 cu = c0 * u, cDu = (c0 * D) * u;
        H00.xx = 2.0 * (cu.x * q.x) - cDu.x * u.x + c0 * D; H00.xy = cu.x * q.y + cu.y * q.x - cDu.x * u.y;
        H00.xz = cu.x * q.z + cu.z * q.x - cDu.x * u.z;     H00.yy = 2.0 * (cu.y * q.y) - cDu.y * u.y + c0 * D;
        H00.yz = cu.y * q.z + cu.z * q.y - cDu.y * u.z;     H00.zz = 2.0 * (cu.z * q.z) - cDu.z * u.z + c0 * D;
        const v3 cv = c1 * v, cDv = (c1 * D) * v;
        H11.xx = 2.0 * (cv.x * r.x) - cDv.x * v.x + c1 * D; H11.xy = cv.x * r.y + cv.y * r.x - cDv.x * v.y;
        H11.xz = cv.x * r.z + cv.z * r.x - cDv.x * v.z;     H11.yy = 2.0 * (cv.y * r.y) - cDv.y * v.y + c1 * D;
        H11.yz = cv.y * r.z + cv.z * r.y - cDv.y * v.z;     H11.zz = 2.0 * (cv.z * r.z) - cDv.z * v.z + c1 * D;
    }
    blk3 H01;   // c01 (I - u u^T - v v^T + D u v^T) = c01 (I - v v^T) - (c01 u) r^T
    {
        const v3 cu = c01 * u, cv = c01 * v;
        H01.m[0] = c01 - cv.x * v.x - cu.x * r.x; H01.m[1] = -cv.x * v.y - cu.x * r.y; H01.m[2] = -cv.x * v.z - cu.x * r.z;
        H01.m[3] = -cv.y * v.x - cu.y * r.x; H01.m[4] = c01 - cv.y * v.y - cu.y * r.y; H01.m[5] = -cv.y * v.z - cu.y * r.z;
        H01.m[6] = -cv.z * v.x - cu.z * r.x; H01.m[7] = -cv.z * v.y - cu.z * r.y; H01.m[8] = c01 - cv.z * v.z - cu.z * r.z;
    }
    const v3 g0 = (kk * il0) * q, g1 = (kk * il1) * r;
    blk3 T, Hea, Heb, K02, K03;
    sym3 S, K00;
    // ---- H00: H_aa, first part of H_ea, first part of H_ee
    sym_mul_skew_r(H00, e, T);                                    // P0 = H00 [e]x
    set_skew_l_sym(S, -1.0, e, T);                                // H_aa = -[e]x P0
    emit.diag(2, S);
    set_skew_l(Hea, 1.0, a, T);                                   // [a]x P0
    K02.m[0] = S.xx; K02.m[1] = S.xy; K02.m[2] = S.xz; K02.m[3] = S.xy; K02.m[4] = S.yy; K02.m[5] = S.yz; K02.m[6] = S.xz; K02.m[7] = S.yz; K02.m[8] = S.zz;
    sym3 Hee;
    sym_mul_skew_r(H00, a, T);                                    // R0 = H00 [a]x
    set_skew_l_sym(Hee, -1.0, a, T);
    // ---- H01^T: rest of H_ea
    {
        blk3 H10;
        H10.m[0] = H01.m[0]; H10.m[1] = H01.m[3]; H10.m[2] = H01.m[6]; H10.m[3] = H01.m[1]; H10.m[4] = H01.m[4]; H10.m[5] = H01.m[7];
        H10.m[6] = H01.m[2]; H10.m[7] = H01.m[5]; H10.m[8] = H01.m[8];
        mul_skew_r(H10, e, T);                                    // P1 = H10 [e]x
    }
    add_skew_l(Hea, -1.0, b, T);
    add_skew(Hea, -1.0, g0);
    emit.off(3, Hea);                                             // K_12
    for (int k = 0; k < 9; ++k) K02.m[k] += Hea.m[k];
    // ---- H01: H_ab, first part of H_eb, cross part of H_ee
    mul_skew_r(H01, e, T);                                        // Q0 = H01 [e]x
    blk3 Hab;
    set_skew_l(Hab, 1.0, e, T);                                   // H_ab = [e]x Q0
    emit.off(5, Hab);                                             // K_23
    set_skew_l(Heb, -1.0, a, T);
    // K_02 = -(K_12 + K_22 + K_32) = -(H_ea + H_aa + H_ab^T)
    K02.m[0] = -K02.m[0] - Hab.m[0]; K02.m[1] = -K02.m[1] - Hab.m[3]; K02.m[2] = -K02.m[2] - Hab.m[6]; K02.m[3] = -K02.m[3] - Hab.m[1];
    K02.m[4] = -K02.m[4] - Hab.m[4]; K02.m[5] = -K02.m[5] - Hab.m[7]; K02.m[6] = -K02.m[6] - Hab.m[2]; K02.m[7] = -K02.m[7] - Hab.m[5];
    K02.m[8] = -K02.m[8] - Hab.m[8];
    emit.off(1, K02);
    K00.xx = -K02.m[0]; K00.xy = -K02.m[1]; K00.xz = -K02.m[2]; K00.yy = -K02.m[4]; K00.yz = -K02.m[5]; K00.zz = -K02.m[8];
    for (int k = 0; k < 9; ++k) K03.m[k] = Hab.m[k];
    {
        blk3 Sx;
        mul_skew_r(H01, b, T);                                    // R1 = H01 [b]x
        set_skew_l(Sx, 1.0, a, T);                                // S = [a]x R1 ; H_ee += S + S^T
        Hee.xx += 2.0 * Sx.m[0]; Hee.xy += Sx.m[1] + Sx.m[3]; Hee.xz += Sx.m[2] + Sx.m[6];
        Hee.yy += 2.0 * Sx.m[4]; Hee.yz += Sx.m[5] + Sx.m[7]; Hee.zz += 2.0 * Sx.m[8];
    }
    // ---- H11: H_bb, rest of H_eb, rest of H_ee
    sym_mul_skew_r(H11, e, T);                                    // Q1 = H11 [e]x
    set_skew_l_sym(S, -1.0, e, T);                                // H_bb = -[e]x Q1
    emit.diag(3, S);
    add_skew_l(Heb, 1.0, b, T);
    add_skew(Heb, 1.0, g1);
    emit.off(4, Heb);                                             // K_13
    // K_03 = -(K_13 + K_23 + K_33) = -(H_eb + H_ab + H_bb)
    K03.m[0] += S.xx; K03.m[1] += S.xy; K03.m[2] += S.xz; K03.m[3] += S.xy; K03.m[4] += S.yy; K03.m[5] += S.yz; K03.m[6] += S.xz; K03.m[7] += S.yz; K03.m[8] += S.zz;
    for (int k = 0; k < 9; ++k) K03.m[k] = -K03.m[k] - Heb.m[k];
    emit.off(2, K03);
    K00.xx -= K03.m[0]; K00.xy -= K03.m[1]; K00.xz -= K03.m[2]; K00.yy -= K03.m[4]; K00.yz -= K03.m[5]; K00.zz -= K03.m[8];
    sym_mul_skew_r(H11, b, T);                                    // R2 = H11 [b]x
    add_skew_l_sym(Hee, -1.0, b, T);
    emit.diag(1, Hee);                                            // K_11
    // K_01 = -(K_11 + K_21 + K_31) = -(H_ee + H_ea^T + H_eb^T)
    blk3 K01;
    K01.m[0] = -(Hee.xx + Hea.m[0] + Heb.m[0]); K01.m[1] = -(Hee.xy + Hea.m[3] + Heb.m[3]); K01.m[2] = -(Hee.xz + Hea.m[6] + Heb.m[6]);
    K01.m[3] = -(Hee.xy + Hea.m[1] + Heb.m[1]); K01.m[4] = -(Hee.yy + Hea.m[4] + Heb.m[4]); K01.m[5] = -(Hee.yz + Hea.m[7] + Heb.m[7]);
    K01.m[6] = -(Hee.xz + Hea.m[2] + Heb.m[2]); K01.m[7] = -(Hee.yz + Hea.m[5] + Heb.m[5]); K01.m[8] = -(Hee.zz + Hea.m[8] + Heb.m[8]);
    emit.off(0, K01);
    // K_00 = -(K_01 + K_02 + K_03): the sum is symmetric, only its upper triangle is formed
    K00.xx -= K01.m[0]; K00.xy -= K01.m[1]; K00.xz -= K01.m[2]; K00.yy -= K01.m[4]; K00.yz -= K01.m[5]; K00.zz -= K01.m[8];
    emit.diag(0, K00);
}

// One triangle for the tiles pipeline: emit.diag(i, sym3) i = 0..2 (K_aa, K_bb, K_cc incl. the mass diagonal),
// emit.off(k, blk3) k = 0..2 (K_ab, K_ac, K_bc), emit.force(i, v3) = fm + fi of vertex i, emit.mass(t8).
template <typename Emit>
EOLC_HD void face_element_tile(v3 xa, v3 xb, v3 xc, double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy, double mu,
                               double lam, double rho, v3 g, double dhh, Emit &emit) {
    const v3 d1 = xb - xa, d2 = xc - xa;
    const v3 nrm = cross(d1, d2);
    const v3 Px = rsq(dot(d1, d1)) * d1;
    v3 Py = cross(nrm, Px);
    Py = rsq(dot(Py, Py)) * Py;
    const double t7 = rest_area2(Xax, Xay, Xbx, Xby, Xcx, Xcy);
    const double t17 = 1.0 / t7;
    const double gb0 = t17 * (Xcy - Xay), gb1 = t17 * (Xax - Xcx);
    const double gc0 = t17 * (Xay - Xby), gc1 = t17 * (Xbx - Xax);
    const double ga0 = -gb0 - gc0, ga1 = -gb1 - gc1;
    const v3 F0 = gb0 * d1 + gc0 * d2, F1 = gb1 * d1 + gc1 * d2;
    const double m11 = dot(Px, F0), m21 = dot(Py, F0), m12 = dot(Px, F1), m22 = dot(Py, F1);
    const double detM = m11 * m22 - m12 * m21;
    const double sg = detM < 0.0 ? -1.0 : (detM == 0.0 ? 0.0 : 1.0);
    double q00 = m11 + sg * m22, q01 = m12 - sg * m21, q10 = m21 - sg * m12, q11 = m22 + sg * m11;
    const double icl = rsq(q00 * q00 + q10 * q10);
    q00 *= icl; q01 *= icl; q10 *= icl; q11 *= icl;
    const v3 r0 = q00 * Px + q10 * Py, r1 = q01 * Px + q11 * Py;
    const double E00 = dot(r0, F0) - 1.0, E10 = dot(r1, F0), E01 = dot(r0, F1), E11 = dot(r1, F1) - 1.0;
    const double A = 0.5 * t7;
    const double tr = E00 + E11;
    const double S00 = 2.0 * mu * E00 + lam * tr, S01 = 2.0 * mu * E01, S10 = 2.0 * mu * E10, S11 = 2.0 * mu * E11 + lam * tr;
    const double t8 = rho * t7;
    {
        const v3 fg = (t8 * (1.0 / 6.0)) * g;
        const v3 s0 = (-A) * r0, s1 = (-A) * r1;
        emit.force(0, (S00 * ga0 + S01 * ga1) * s0 + (S10 * ga0 + S11 * ga1) * s1 + fg);
        emit.force(1, (S00 * gb0 + S01 * gb1) * s0 + (S10 * gb0 + S11 * gb1) * s1 + fg);
        emit.force(2, (S00 * gc0 + S01 * gc1) * s0 + (S10 * gc0 + S11 * gc1) * s1 + fg);
    }
    emit.mass(t8);
    const double a2mu = dhh * A * 2.0 * mu, alam = dhh * A * lam;
    const double RR0 = a2mu * (r0.x * r0.x + r1.x * r1.x), RR1 = a2mu * (r0.x * r0.y + r1.x * r1.y), RR2 = a2mu * (r0.x * r0.z + r1.x * r1.z),
                 RR3 = a2mu * (r0.y * r0.y + r1.y * r1.y), RR4 = a2mu * (r0.y * r0.z + r1.y * r1.z), RR5 = a2mu * (r0.z * r0.z + r1.z * r1.z);
    const v3 qa = ga0 * r0 + ga1 * r1, qb = gb0 * r0 + gb1 * r1, qc = gc0 * r0 + gc1 * r1;
    const v3 la = alam * qa, lb = alam * qb, lc = alam * qc;
    const double md = t8 / 12.0, mo = 0.5 * md;
    {
        sym3 Sd;
        double s = ga0 * ga0 + ga1 * ga1;
        Sd.xx = md + (s * RR0 + la.x * qa.x); Sd.xy = s * RR1 + la.x * qa.y; Sd.xz = s * RR2 + la.x * qa.z;
        Sd.yy = md + (s * RR3 + la.y * qa.y); Sd.yz = s * RR4 + la.y * qa.z; Sd.zz = md + (s * RR5 + la.z * qa.z);
        emit.diag(0, Sd);
        s = gb0 * gb0 + gb1 * gb1;
        Sd.xx = md + (s * RR0 + lb.x * qb.x); Sd.xy = s * RR1 + lb.x * qb.y; Sd.xz = s * RR2 + lb.x * qb.z;
        Sd.yy = md + (s * RR3 + lb.y * qb.y); Sd.yz = s * RR4 + lb.y * qb.z; Sd.zz = md + (s * RR5 + lb.z * qb.z);
        emit.diag(1, Sd);
        s = gc0 * gc0 + gc1 * gc1;
        Sd.xx = md + (s * RR0 + lc.x * qc.x); Sd.xy = s * RR1 + lc.x * qc.y; Sd.xz = s * RR2 + lc.x * qc.z;
        Sd.yy = md + (s * RR3 + lc.y * qc.y); Sd.yz = s * RR4 + lc.y * qc.z; Sd.zz = md + (s * RR5 + lc.z * qc.z);
        emit.diag(2, Sd);
    }
    {
        blk3 B;
        auto off = [&](double s, v3 li, v3 qj) {
            B.m[0] = mo + (s * RR0 + li.x * qj.x); B.m[1] = s * RR1 + li.x * qj.y; B.m[2] = s * RR2 + li.x * qj.z;
            B.m[3] = s * RR1 + li.y * qj.x; B.m[4] = mo + (s * RR3 + li.y * qj.y); B.m[5] = s * RR4 + li.y * qj.z;
            B.m[6] = s * RR2 + li.z * qj.x; B.m[7] = s * RR4 + li.z * qj.y; B.m[8] = mo + (s * RR5 + li.z * qj.z);
        };
        off(ga0 * gb0 + ga1 * gb1, la, qb); emit.off(0, B);
        off(ga0 * gc0 + ga1 * gc1, la, qc); emit.off(1, B);
        off(gb0 * gc0 + gb1 * gc1, lb, qc); emit.off(2, B);
    }
}

// ------------------------------------------------------------------------------------------------
// Row forms (owner-computes assembly): the 3x3 blocks of ONE block-row of an element matrix, i.e. what one element
// contributes to the CSR rows of ONE of its nodes.  Same closed forms as above, grouped by the row vertex:
//
//  face, row v:   B_vj = M_vj + dhh A (2 mu (g_v.g_j) R^T R + lambda q_v q_j^T),  j = a,b,c   (+ f_v, t8)
//  edge, row i:   with the row's vectors U_i, Y_i, V_i, Z_i, w0_i, w1_i (zero where the vertex is not in that triangle)
//       a  = k0 (Y_i - 3D U_i) - k01 (Z_i - D V_i)    b  = k0 U_i - k01 V_i    c  = k0 D w0_i + k01 w1_i
//       a' = k1 (Z_i - 3D V_i) - k01 (Y_i - D U_i)    b' = k1 V_i - k01 U_i    c' = k1 D w1_i + k01 w0_i
//       B_ij = [j in t0] (a U_j^T + b Y_j^T - w0_j c^T + (c.w0_j) I) + [j in t1] (a' V_j^T + b' Z_j^T - w1_j c'^T + (c'.w1_j) I)
//              + s0_ij [kk g0]x + s1_ij [kk g1]x,      k0 = c dhh/|n0|^2, k1 = c dhh/|n1|^2, k01 = -c dhh/(|n0||n1|), kk = -c dhh
//  The code is uniform in v / i (selects, no branches), so a warp may mix rows.
struct FaceRowOut { blk3 K[3]; double f[3]; double md, mo; };   // md = t8/12, mo = t8/24 (ComputeInertial.cpp:44-47)

EOLC_HD v3 sel3(int i, v3 a, v3 b, v3 c) { return i == 0 ? a : (i == 1 ? b : c); }
EOLC_HD v3 sel4(int i, v3 a, v3 b, v3 c, v3 d) { return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d)); }
// Lame-type constants of ComputeMembrane.cpp:46,84-87 (evaluated once per fill on the host)
inline double membrane_mu(double e, double nu) { return e / (1.0 + nu) * 0.5; }
inline double membrane_lambda(double e, double nu) { return e * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)); }

EOLC_HD void face_row(int v, v3 xa, v3 xb, v3 xc, double Xax, double Xay, double Xbx, double Xby, double Xcx, double Xcy,
                      double mu, double lam, double rho, v3 g, double dhh, FaceRowOut &o) {
    v3 d1 = xb - xa, d2 = xc - xa;
    v3 nrm = cross(d1, d2);
    v3 Px = rsq(dot(d1, d1)) * d1;
    v3 Py = cross(nrm, Px);
    Py = rsq(dot(Py, Py)) * Py;
    double t7 = rest_area2(Xax, Xay, Xbx, Xby, Xcx, Xcy);
    double t17 = 1.0 / t7;
    double gb0 = t17 * (Xcy - Xay), gb1 = t17 * (Xax - Xcx);
    double gc0 = t17 * (Xay - Xby), gc1 = t17 * (Xbx - Xax);
    double ga0 = -gb0 - gc0, ga1 = -gb1 - gc1;
    v3 F0 = gb0 * d1 + gc0 * d2, F1 = gb1 * d1 + gc1 * d2;
    double m11 = dot(Px, F0), m21 = dot(Py, F0), m12 = dot(Px, F1), m22 = dot(Py, F1);
    double detM = m11 * m22 - m12 * m21;
    double sg = detM < 0.0 ? -1.0 : (detM == 0.0 ? 0.0 : 1.0);
    double q00 = m11 + sg * m22, q01 = m12 - sg * m21, q10 = m21 - sg * m12, q11 = m22 + sg * m11;
    double icl = rsq(q00 * q00 + q10 * q10);
    q00 *= icl; q01 *= icl; q10 *= icl; q11 *= icl;
    v3 r0 = q00 * Px + q10 * Py, r1 = q01 * Px + q11 * Py;
    double E00 = dot(r0, F0) - 1.0, E10 = dot(r1, F0), E01 = dot(r0, F1), E11 = dot(r1, F1) - 1.0;
    double A = 0.5 * t7;
    double tr = E00 + E11;
    double S00 = 2.0 * mu * E00 + lam * tr, S01 = 2.0 * mu * E01, S10 = 2.0 * mu * E10, S11 = 2.0 * mu * E11 + lam * tr;
    double t8 = rho * t7;
    const double gv0 = v == 0 ? ga0 : (v == 1 ? gb0 : gc0), gv1 = v == 0 ? ga1 : (v == 1 ? gb1 : gc1);
    {
        v3 fv = (-A * (S00 * gv0 + S01 * gv1)) * r0 + (-A * (S10 * gv0 + S11 * gv1)) * r1;
        double s6 = t8 * (1.0 / 6.0);
        o.f[0] = fv.x + s6 * g.x; o.f[1] = fv.y + s6 * g.y; o.f[2] = fv.z + s6 * g.z;
    }
    const double md = t8 / 12.0, mo = 0.5 * md;   // t8/24 == (t8/12)/2 exactly
    o.md = md; o.mo = mo;
    double RR[6] = {r0.x * r0.x + r1.x * r1.x, r0.x * r0.y + r1.x * r1.y, r0.x * r0.z + r1.x * r1.z,
                    r0.y * r0.y + r1.y * r1.y, r0.y * r0.z + r1.y * r1.z, r0.z * r0.z + r1.z * r1.z};
    v3 qa = ga0 * r0 + ga1 * r1, qb = gb0 * r0 + gb1 * r1, qc = gc0 * r0 + gc1 * r1;
    v3 qv = sel3(v, qa, qb, qc);
    const double a2mu = dhh * A * 2.0 * mu, alam = dhh * A * lam;
    // products q_v[p]*q_j[q] are formed first (commutative), so B_vj of this row and B_jv of row j are exact transposes
    auto block = [&](blk3 &B, double gj0, double gj1, v3 qj, double mass) {
        double s = a2mu * (gv0 * gj0 + gv1 * gj1);
        B.m[0] = mass + (s * RR[0] + alam * (qv.x * qj.x)); B.m[1] = s * RR[1] + alam * (qv.x * qj.y); B.m[2] = s * RR[2] + alam * (qv.x * qj.z);
        B.m[3] = s * RR[1] + alam * (qv.y * qj.x); B.m[4] = mass + (s * RR[3] + alam * (qv.y * qj.y)); B.m[5] = s * RR[4] + alam * (qv.y * qj.z);
        B.m[6] = s * RR[2] + alam * (qv.z * qj.x); B.m[7] = s * RR[4] + alam * (qv.z * qj.y); B.m[8] = mass + (s * RR[5] + alam * (qv.z * qj.z));
    };
    block(o.K[0], ga0, ga1, qa, v == 0 ? md : mo);
    block(o.K[1], gb0, gb1, qb, v == 1 ? md : mo);
    block(o.K[2], gc0, gc1, qc, v == 2 ? md : mo);
}

struct EdgeRowOut { blk3 K[4]; };

// Low-register formulation: the per-column vectors U_j, Y_j, V_j, Z_j are rebuilt from u, v and the positions right
// before block j is formed and die with it; `emit(j, B)` consumes each block immediately (the kernel parks it in
// shared memory), so only x, u, v, the six row vectors and g0/g1 stay live across the four blocks.
template <typename Emit>
EOLC_HD void edge_row_emit(int i, v3 x0, v3 x1, v3 x2, v3 x3, double X0x, double X0y, double X1x, double X1y, double X2x,
                           double X2y, double X3x, double X3y, double beta, double dhh, Emit emit) {
    double ex = X1x - X0x, ey = X1y - X0y;
    double t6 = beta * (ex * ex + ey * ey);
    double den = stencil_area(X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y);
    double c = 1.5 * t6 / den;
    v3 u, v;
    double D, k0, k1, k01, kg0, kg1;
    {
        v3 e = x1 - x0;
        v3 n0 = cross(e, x2 - x0), n1 = cross(x3 - x0, e);
        double il0 = rsq(dot(n0, n0)), il1 = rsq(dot(n1, n1));
        u = il0 * n0; v = il1 * n1;
        D = dot(u, v);
        const double kk = -c * dhh;
        kg0 = kk * il0; kg1 = kk * il1;
        k0 = -kg0 * il0; k1 = -kg1 * il1; k01 = kg0 * il1;
    }
    const v3 z = mk3(0.0, 0.0, 0.0);
    // the row's vectors: w0_i in (x2-x1, x0-x2, x1-x0, 0), w1_i in (x1-x3, x3-x0, 0, x0-x1)
    v3 ra, rb, rc, pa, pb, pc;
    {
        v3 w0i = sel4(i, x2 - x1, x0 - x2, x1 - x0, z), w1i = sel4(i, x1 - x3, x3 - x0, z, x0 - x1);
        v3 Ui = cross(u, w0i), Yi = cross(v, w0i), Vi = cross(v, w1i), Zi = cross(u, w1i);
        const double D3 = 3.0 * D;
        ra = k0 * (Yi - D3 * Ui) - k01 * (Zi - D * Vi);
        rb = k0 * Ui - k01 * Vi;
        rc = (k0 * D) * w0i + k01 * w1i;
        pa = k1 * (Zi - D3 * Vi) - k01 * (Yi - D * Ui);
        pb = k1 * Vi - k01 * Ui;
        pc = (k1 * D) * w1i + k01 * w0i;
    }
    v3 g0 = kg0 * (v - D * u), g1 = kg1 * (u - D * v);
    const double c0j0 = i == 1 ? 1.0 : (i == 2 ? -1.0 : 0.0), c0j1 = i == 0 ? -1.0 : (i == 2 ? 1.0 : 0.0), c0j2 = i == 0 ? 1.0 : (i == 1 ? -1.0 : 0.0);
    const double c1j0 = i == 1 ? -1.0 : (i == 3 ? 1.0 : 0.0), c1j1 = i == 0 ? 1.0 : (i == 3 ? -1.0 : 0.0), c1j3 = i == 0 ? -1.0 : (i == 1 ? 1.0 : 0.0);
#define EOLC_T0(B, wj) { v3 w_ = (wj); add_outer(B, ra, cross(u, w_)); add_outer(B, rb, cross(v, w_)); add_outer(B, -1.0 * w_, rc); add_diag(B, dot(rc, w_)); }
#define EOLC_T1(B, wj) { v3 w_ = (wj); add_outer(B, pa, cross(v, w_)); add_outer(B, pb, cross(u, w_)); add_outer(B, -1.0 * w_, pc); add_diag(B, dot(pc, w_)); }
    {
        blk3 B;
        for (int q = 0; q < 9; ++q) B.m[q] = 0.0;
        EOLC_T0(B, x2 - x1); EOLC_T1(B, x1 - x3); add_skew(B, c0j0, g0); add_skew(B, c1j0, g1);
        emit(0, B);
    }
    {
        blk3 B;
        for (int q = 0; q < 9; ++q) B.m[q] = 0.0;
        EOLC_T0(B, x0 - x2); EOLC_T1(B, x3 - x0); add_skew(B, c0j1, g0); add_skew(B, c1j1, g1);
        emit(1, B);
    }
    {
        blk3 B;
        for (int q = 0; q < 9; ++q) B.m[q] = 0.0;
        EOLC_T0(B, x1 - x0); add_skew(B, c0j2, g0);
        emit(2, B);
    }
    {
        blk3 B;
        for (int q = 0; q < 9; ++q) B.m[q] = 0.0;
        EOLC_T1(B, x0 - x1); add_skew(B, c1j3, g1);
        emit(3, B);
    }
#undef EOLC_T0
#undef EOLC_T1
}

EOLC_HD void edge_row(int i, v3 x0, v3 x1, v3 x2, v3 x3, double X0x, double X0y, double X1x, double X1y, double X2x,
                      double X2y, double X3x, double X3y, double beta, double dhh, EdgeRowOut &o) {
    edge_row_emit(i, x0, x1, x2, x3, X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y, beta, dhh,
                  [&](int j, const blk3 &B) { o.K[j] = B; });
}

}  // namespace eolc
