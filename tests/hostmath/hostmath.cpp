// Test shim: compiles the SAME element arithmetic the CUDA kernels use (eol_cloth_b200/csrc/elements.cuh)
// for the host, so the hand-derived formulas can be checked against the oracle without a GPU.
// Test infrastructure only — never linked into libeolc_b200.so.
#include "../../eol_cloth_b200/csrc/elements.cuh"
using namespace eolc;
extern "C" {
void hostmath_face(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                   const double *Xc, double e, double nu, double rho, const double *g, double dhh,
                   double *f9, double *t8, double *K6x9) {
    FaceOut o;
    face_element(mk3(xa[0], xa[1], xa[2]), mk3(xb[0], xb[1], xb[2]), mk3(xc[0], xc[1], xc[2]), Xa[0], Xa[1], Xb[0], Xb[1],
                 Xc[0], Xc[1], e, nu, rho, mk3(g[0], g[1], g[2]), dhh, o);
    for (int i = 0; i < 3; ++i) { f9[i] = o.fa[i]; f9[3 + i] = o.fb[i]; f9[6 + i] = o.fc[i]; }
    *t8 = o.t8;
    for (int b = 0; b < 6; ++b) for (int k = 0; k < 9; ++k) K6x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_edge(const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                   const double *X1, const double *X2, const double *X3, double beta, double dhh, double *K10x9) {
    EdgeOut o;
    edge_element(mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
                 X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, dhh, o);
    for (int b = 0; b < 10; ++b) for (int k = 0; k < 9; ++k) K10x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_face_row(int v, const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                       const double *Xc, double e, double nu, double rho, const double *g, double dhh,
                       double *f3, double *t8, double *K3x9) {
    FaceRowOut o;
    face_row(v, mk3(xa[0], xa[1], xa[2]), mk3(xb[0], xb[1], xb[2]), mk3(xc[0], xc[1], xc[2]), Xa[0], Xa[1], Xb[0], Xb[1],
             Xc[0], Xc[1], membrane_mu(e, nu), membrane_lambda(e, nu), rho, mk3(g[0], g[1], g[2]), dhh, o);
    for (int i = 0; i < 3; ++i) f3[i] = o.f[i];
    *t8 = o.md * 12.0;
    for (int b = 0; b < 3; ++b) for (int k = 0; k < 9; ++k) K3x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_edge_row(int i, const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                       const double *X1, const double *X2, const double *X3, double beta, double dhh, double *K4x9) {
    EdgeRowOut o;
    edge_row(i, mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
             X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, dhh, o);
    for (int b = 0; b < 4; ++b) for (int k = 0; k < 9; ++k) K4x9[9 * b + k] = o.K[b].m[k];
}
}
