// cd_math.cuh — strict (never-contracted) FP64 helpers for the collision narrow phase.
//
// Contact index sets must be bit-exact with the reference, so every product/sum below is an individually
// rounded IEEE operation (__dmul_rn/__dadd_rn/__dsub_rn on the device — these are never fused into FMA,
// independent of -fmad; plain operators on the host, built without FMA contraction), in the evaluation
// order of the reference's Eigen 3.3 expressions:
//   dot / squaredNorm   (p0 + p1) + p2                     (linear-vectorised redux, Packet2d + tail)
//   cross               (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
//   normalized()        v / sqrt(v.v) if v.v > 0 else v     (true division)
// and of raytri.cpp's DOT/CROSS/SUB macros (same orders).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define CD_HD __host__ __device__ __forceinline__
#else
#define CD_HD inline
#endif

namespace cdm {

CD_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CD_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CD_HD double sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
CD_HD double dv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
CD_HD double sq(double a) {
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}

struct V3 { double x, y, z; };
CD_HD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CD_HD V3 ld(const double *p) { return mk(p[0], p[1], p[2]); }
CD_HD void st(double *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
CD_HD V3 operator+(V3 a, V3 b) { return mk(add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)); }
CD_HD V3 operator-(V3 a, V3 b) { return mk(sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)); }
CD_HD V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
CD_HD V3 scale(double s, V3 a) { return mk(mul(s, a.x), mul(s, a.y), mul(s, a.z)); }
CD_HD V3 divs(V3 a, double s) { return mk(dv(a.x, s), dv(a.y, s), dv(a.z, s)); }
CD_HD double dot(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
CD_HD V3 cross(V3 a, V3 b) {
    return mk(sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x)));
}
CD_HD double norm(V3 a) { return sq(dot(a, a)); }
CD_HD V3 normalized(V3 a) {
    double z = dot(a, a);
    if (z > 0.0) return divs(a, sq(z));
    return a;
}

// check_AABB, /root/reference/src/boxTriCollision.cpp:471-486 (both boxes padded by 1e-3)
CD_HD bool check_aabb(const double *a1, const double *a2) {
    const double t = 1e-3;
    return add(a1[3], t) >= sub(a2[0], t) && sub(a1[0], t) <= add(a2[3], t) && add(a1[4], t) >= sub(a2[1], t) &&
           sub(a1[1], t) <= add(a2[4], t) && add(a1[5], t) >= sub(a2[2], t) && sub(a1[2], t) <= add(a2[5], t);
}
CD_HD bool check_aabb_point(V3 p, const double *a2) {
    double a1[6] = {p.x, p.y, p.z, p.x, p.y, p.z};
    return check_aabb(a1, a2);
}

// barycentric, boxTriCollision.cpp:488-505 (returns alpha, beta)
CD_HD void barycentric(double &alpha, double &beta, V3 a, V3 b, V3 c, V3 p) {
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    double denom = sub(mul(d00, d11), mul(d01, d01));
    beta = dv(sub(mul(d11, d20), mul(d01, d21)), denom);
    double gamma = dv(sub(mul(d00, d21), mul(d01, d20)), denom);
    alpha = sub(sub(1.0, beta), gamma);
}

// lineline, boxTriCollision.cpp:507-525
CD_HD void lineline(double &a, double &b, V3 A1, V3 A2, V3 B1, V3 B2) {
    V3 B2B1 = B2 - B1, A1B1 = A1 - B1, A2A1 = A2 - A1;
    V3 X = cross(A2A1, B2B1);
    double nA = dot(cross(B2B1, A1B1), X);
    double nB = dot(cross(A2A1, A1B1), X);
    double d = dot(cross(A2A1, B2B1), X);
    a = dv(nA, d);
    b = dv(nB, d);
}

// linepoint, boxTriCollision.cpp:527-536
CD_HD double linepoint(V3 A, V3 B, V3 P) {
    V3 AP = P - A, AB = B - A;
    return dv(dot(AP, AB), dot(AB, AB));
}

// intersect_triangle3_inc, /root/reference/src/raytri.cpp:260-316.  EPSILON 1e-6, ZERO = -EPSILON:
// `u < -ZERO` means u < +1e-6, i.e. the borders are inset, not inclusive.
CD_HD int raytri3_inc(V3 orig, V3 dir, V3 vert0, V3 vert1, V3 vert2, double &t) {
    const double EPS = 1e-6, ZERO = -1e-6;
    V3 edge1 = vert1 - vert0, edge2 = vert2 - vert0;
    V3 pvec = cross(dir, edge2);
    double det = dot(edge1, pvec);
    V3 tvec = orig - vert0;
    double inv_det = dv(1.0, det);
    V3 qvec = cross(tvec, edge1);
    double u, v;
    if (det > EPS) {
        u = dot(tvec, pvec);
        if (u < -ZERO || u > add(det, ZERO)) return 0;
        v = dot(dir, qvec);
        if (v < -ZERO || add(u, v) > add(det, ZERO)) return 0;
    } else if (det < -EPS) {
        u = dot(tvec, pvec);
        if (u > ZERO || u < sub(det, ZERO)) return 0;
        v = dot(dir, qvec);
        if (v > ZERO || add(u, v) < sub(det, ZERO)) return 0;
    } else {
        return 0;
    }
    t = mul(dot(edge2, qvec), inv_det);
    return 1;
}

// intersect_square, boxTriCollision.cpp:550-600
CD_HD int intersect_square(V3 x0, V3 dx, V3 xa, V3 xb, V3 xc, double &t) {
    V3 xd = xa + scale(2.0, xc - xa);
    V3 xe = xb + scale(2.0, xc - xb);
    if (raytri3_inc(x0, dx, xa, xb, xc, t)) return 1;
    if (raytri3_inc(x0, dx, xb, xd, xc, t)) return 1;
    if (raytri3_inc(x0, dx, xd, xe, xc, t)) return 1;
    if (raytri3_inc(x0, dx, xe, xa, xc, t)) return 1;
    t = -1.0;
    return 0;
}

}  // namespace cdm
