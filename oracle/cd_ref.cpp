// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's collision narrow phase over flat arrays, Eigen-free,
// single-threaded, linking the reference's own raytri.cpp (intersect_triangle3_inc, compiled
// UNMODIFIED from /root/reference/src into oracle/_ref/ by oracle/Makefile).
//
// PARITY PINNED TO THE REFERENCE'S OWN CODE: oracle/_ref/libbtc_ref.so is boxTriCollision.cpp + Collisions.cpp + raytri.cpp
// compiled UNMODIFIED from /root/reference/src against oracle/mini_eigen (oracle/Makefile, oracle/ref_cd_driver.cpp), and
// tests/test_cd_ref_pin.py requires this restatement to equal it in every field of every contact, bit for bit, on every case
// where the reference is defined (its `int` edge hash overflows above ~19 k nodes).  The reference ships no tests or golden
// vectors of its own.  What remains restated rather than run: the arithmetic inside Eigen's small-vector
// operators (mini_eigen header; the contact set is shown not to depend on the one ambiguous detail, the reduction order).
// This file exists because the GPU is also checked above the reference's size limit (256^2 ... 1024^2 sheets), where the
// 64-bit key below continues the order the reference intends.
//
// Follows, line by line:
//   createEdges        src/boxTriCollision.cpp:141-231   (DEVIATION: the sort key is int64; the reference's
//                      `int hash = kmin + (n+1)*kmax` overflows (UB) once (3F+1)*N > 2^31 — identical order
//                      wherever the reference is defined)
//   createFaceNormals  :233-247     createVertNormals :249-283     box tables :289-398   createBox :400-420
//   build_AABB_*       :425-468     check_AABB :471-486  barycentric :488-505  lineline :507-525
//   linepoint          :527-536     intersect_square :550-600
//   boxTriCollision    :617-1063    pointTriCollision :1067-1224
//   CD / CD2           src/Collisions.cpp:11-78
//
// Eigen 3.3 small-vector conventions restated (see oracle/forces_ref.cpp header): dot = (p0+p1)+p2,
// normalized() = v / sqrt(squaredNorm) if squaredNorm > 0 else v, true division, textbook cross,
// 4x4 products accumulate k = 0..3 left to right.  Compile with -ffp-contract=off (no FMA).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>
#include <memory>
#include "../include/eolc.h"

// From raytri.cpp (reference object code, C++ linkage as declared at boxTriCollision.cpp:80-83)
int intersect_triangle3_inc(const double *orig, const double *dir, const double *vert0, const double *vert1,
                            const double *vert2, double *t, double *u, double *v);

namespace {

struct V3 {
    double v[3];
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    const double *data() const { return v; }
};
inline V3 mk(double a, double b, double c) { V3 r; r.v[0] = a; r.v[1] = b; r.v[2] = c; return r; }
inline V3 operator-(const V3 &a, const V3 &b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator+(const V3 &a, const V3 &b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a) { return mk(-a[0], -a[1], -a[2]); }
inline V3 operator*(double s, const V3 &a) { return mk(s * a[0], s * a[1], s * a[2]); }
inline V3 operator/(const V3 &a, double s) { return mk(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const V3 &a, const V3 &b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline V3 cross(const V3 &a, const V3 &b) {
    return mk(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(const V3 &a) {
    double z = dot(a, a);
    if (z > 0.0) return a / std::sqrt(z);
    return a;
}
inline V3 col(const double *M, int i) { return mk(M[3 * i], M[3 * i + 1], M[3 * i + 2]); }

struct Edge {  // btc::Edge, boxTriCollision.h:30-41
    int verts[4] = {0, 0, 0, 0};
    int faces[2] = {0, 0};
    bool internal = false;
    double angle = 0;
    V3 normals[2] = {mk(0, 0, 0), mk(0, 0, 0)};
};

// ---- createEdges, boxTriCollision.cpp:141-231 -------------------------------------------------
void createEdges(std::vector<Edge> &edges, int nf, const int32_t *faces /*3 per face*/, const double *verts) {
    struct FaceEdge { int verts[3]; int face; int64_t hash; };
    int64_t n = 3 * (int64_t)nf;
    std::vector<FaceEdge> tmp;
    tmp.reserve(n);
    for (int k = 0; k < nf; ++k)
        for (int i = 0; i < 3; ++i) {
            FaceEdge fe;
            fe.verts[0] = (i + 0) % 3; fe.verts[1] = (i + 1) % 3; fe.verts[2] = (i + 2) % 3;
            fe.face = k;
            int a = faces[3 * k + fe.verts[0]], b = faces[3 * k + fe.verts[1]];
            int64_t kmin = std::min(a, b) + 1, kmax = std::max(a, b) + 1;
            fe.hash = kmin + (n + 1) * kmax;
            tmp.push_back(fe);
        }
    std::stable_sort(tmp.begin(), tmp.end(), [](const FaceEdge &a, const FaceEdge &b) { return a.hash < b.hash; });
    int64_t k = 0;
    while (k < n) {
        int f0 = tmp[k].face;
        int e0 = faces[3 * f0 + tmp[k].verts[0]], e1 = faces[3 * f0 + tmp[k].verts[1]];
        V3 xa0 = col(verts, faces[3 * f0]), xb0 = col(verts, faces[3 * f0 + 1]), xc0 = col(verts, faces[3 * f0 + 2]);
        V3 n0 = normalized(cross(xb0 - xa0, xc0 - xa0));
        Edge edge;
        if (k < n - 1 && tmp[k].hash == tmp[k + 1].hash) {
            int f1 = tmp[k + 1].face;
            V3 xa1 = col(verts, faces[3 * f1]), xb1 = col(verts, faces[3 * f1 + 1]), xc1 = col(verts, faces[3 * f1 + 2]);
            V3 n1 = normalized(cross(xb1 - xa1, xc1 - xa1));
            edge.verts[0] = e0; edge.verts[1] = e1;
            edge.verts[2] = faces[3 * f0 + tmp[k].verts[2]];
            edge.verts[3] = faces[3 * f1 + tmp[k + 1].verts[2]];
            edge.faces[0] = f0; edge.faces[1] = f1;
            edge.internal = true;
            edge.angle = std::acos(dot(n0, n1));
            edge.normals[0] = n0; edge.normals[1] = n1;
            k += 2;
        } else {
            edge.verts[0] = e0; edge.verts[1] = e1;
            edge.verts[2] = faces[3 * f0 + tmp[k].verts[2]];
            edge.verts[3] = -1;
            edge.faces[0] = f0; edge.faces[1] = -1;
            edge.internal = false;
            edge.angle = 1e9;
            edge.normals[0] = n0;
            k += 1;
        }
        edges.push_back(edge);
    }
}

// :233-247 ; normals out 3 per face
void createFaceNormals(std::vector<double> &normals, int nf, const int32_t *faces, const double *verts) {
    normals.assign(3 * (size_t)nf, 0.0);
    for (int k = 0; k < nf; ++k) {
        V3 xa = col(verts, faces[3 * k]), xb = col(verts, faces[3 * k + 1]), xc = col(verts, faces[3 * k + 2]);
        V3 dba = xb - xa, dac = xa - xc;
        V3 nn = normalized(cross(dba, -dac));
        normals[3 * k] = nn[0]; normals[3 * k + 1] = nn[1]; normals[3 * k + 2] = nn[2];
    }
}

// :249-283
void createVertNormals(std::vector<double> &normals, int nv, int nf, const int32_t *faces, const double *verts) {
    normals.assign(3 * (size_t)nv, 0.0);
    std::vector<double> angles(nv, 0.0);
    for (int k = 0; k < nf; ++k) {
        const int32_t *f = faces + 3 * k;
        V3 xa = col(verts, f[0]), xb = col(verts, f[1]), xc = col(verts, f[2]);
        V3 dba = xb - xa, dcb = xc - xb, dac = xa - xc;
        V3 nor = normalized(cross(dba, -dac));
        dba = normalized(dba); dcb = normalized(dcb); dac = normalized(dac);
        double angle1 = std::acos(dot(dba, -dac));
        double angle2 = std::acos(dot(dcb, -dba));
        double angle3 = std::acos(dot(dac, -dcb));
        for (int i = 0; i < 3; ++i) {
            normals[3 * f[0] + i] += angle1 * nor[i];
            normals[3 * f[1] + i] += angle2 * nor[i];
            normals[3 * f[2] + i] += angle3 * nor[i];
        }
        angles[f[0]] += angle1; angles[f[1]] += angle2; angles[f[2]] += angle3;
    }
    for (int k = 0; k < nv; ++k) {
        V3 nor = col(normals.data(), k) / angles[k];
        nor = normalized(nor);
        normals[3 * k] = nor[0]; normals[3 * k + 1] = nor[1]; normals[3 * k + 2] = nor[2];
    }
}

// ---- box tables, boxTriCollision.cpp:289-398 (data) ----------------------------------------------
const double kUnitVerts[14][3] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1},
                                  {1, 1, 1},    {-1, 0, 0},  {1, 0, 0},   {0, -1, 0}, {0, 1, 0},   {0, 0, -1}, {0, 0, 1}};
const int32_t kFaces1[24 * 3] = {0, 8,  2, 1, 8,  0, 3, 8,  1, 2, 8,  3, 4, 10, 0, 5, 10, 4, 1, 10, 5, 0, 10, 1,
                                 6, 9,  4, 7, 9,  6, 5, 9,  7, 4, 9,  5, 2, 11, 6, 3, 11, 2, 7, 11, 3, 6, 11, 7,
                                 1, 13, 3, 5, 13, 1, 7, 13, 5, 3, 13, 7, 0, 12, 4, 2, 12, 0, 6, 12, 2, 4, 12, 6};
const int kEdgeVerts1[12][4] = {{0, 1, 8, 10}, {2, 0, 8, 12}, {1, 3, 8, 13}, {3, 2, 8, 11}, {0, 4, 10, 12}, {5, 1, 10, 13},
                                {4, 5, 10, 9}, {6, 2, 11, 12}, {4, 6, 9, 12}, {3, 7, 11, 13}, {7, 5, 9, 13}, {6, 7, 9, 11}};
const int kVertEdges1[8][3] = {{0, 1, 4}, {0, 2, 5}, {1, 3, 7}, {2, 3, 9}, {4, 6, 8}, {5, 6, 10}, {7, 8, 11}, {9, 10, 11}};
const int kEdgeFaces1[12][2] = {{1, 7}, {0, 21}, {2, 16}, {3, 13}, {4, 20}, {6, 17}, {5, 11}, {12, 22}, {8, 23}, {14, 19}, {10, 18}, {9, 15}};

struct Box {
    double verts1[14 * 3];
    std::vector<double> faceNors1, vertNors1;
    Edge edges1[12];
};

// createBox :400-420 ; E1 column-major 4x4
void createBox(Box &B, const double *whd, const double *E1) {
    double S[4] = {0.5 * whd[0], 0.5 * whd[1], 0.5 * whd[2], 1.0};
    double E[16];  // E = E1 * S (S diagonal; products with the zero entries of S add exact zeros)
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            double acc = E1[0 * 4 + r] * (c == 0 ? S[0] : 0.0);
            acc = acc + E1[1 * 4 + r] * (c == 1 ? S[1] : 0.0);
            acc = acc + E1[2 * 4 + r] * (c == 2 ? S[2] : 0.0);
            acc = acc + E1[3 * 4 + r] * (c == 3 ? S[3] : 0.0);
            E[c * 4 + r] = acc;
        }
    for (int i = 0; i < 14; ++i)
        for (int r = 0; r < 3; ++r) {  // verts1 = E * verts1_ (homogeneous w = 1), k = 0..3 left to right
            double acc = E[0 * 4 + r] * kUnitVerts[i][0];
            acc = acc + E[1 * 4 + r] * kUnitVerts[i][1];
            acc = acc + E[2 * 4 + r] * kUnitVerts[i][2];
            acc = acc + E[3 * 4 + r] * 1.0;
            B.verts1[3 * i + r] = acc;
        }
    createFaceNormals(B.faceNors1, 24, kFaces1, B.verts1);
    createVertNormals(B.vertNors1, 14, 24, kFaces1, B.verts1);
    for (int k = 0; k < 12; ++k) {
        Edge &e = B.edges1[k];
        for (int i = 0; i < 4; ++i) e.verts[i] = kEdgeVerts1[k][i];
        e.faces[0] = kEdgeFaces1[k][0]; e.faces[1] = kEdgeFaces1[k][1];
        e.internal = false;
        e.normals[0] = col(B.faceNors1.data(), e.faces[0]);
        e.normals[1] = col(B.faceNors1.data(), e.faces[1]);
        e.angle = std::acos(dot(e.normals[0], e.normals[1]));
    }
}

struct AABB { double v[6]; };

void build_AABB_B(AABB &aabb, const double *V, int n) {  // :425-433
    for (int r = 0; r < 3; ++r) {
        double mn = V[r], mx = V[r];
        for (int i = 1; i < n; ++i) { mn = std::min(mn, V[3 * i + r]); mx = std::max(mx, V[3 * i + r]); }
        aabb.v[r] = mn; aabb.v[3 + r] = mx;
    }
}
AABB aabb_face(const double *V, const int32_t *f) {  // :436-452
    AABB a;
    for (int i = 0; i < 3; ++i) {
        double xa = V[3 * f[0] + i], xb = V[3 * f[1] + i], xc = V[3 * f[2] + i];
        a.v[i] = std::min(xc, std::min(xb, xa));
        a.v[i + 3] = std::max(xc, std::max(xb, xa));
    }
    return a;
}
AABB aabb_edge(const double *V, int e0, int e1) {  // :455-468
    AABB a;
    for (int i = 0; i < 3; ++i) {
        double x0 = V[3 * e0 + i], x1 = V[3 * e1 + i];
        a.v[i] = std::min(x1, x0); a.v[i + 3] = std::max(x1, x0);
    }
    return a;
}
bool check_AABB(const AABB &a1, const AABB &a2) {  // :471-486
    const double thresh = 1e-3;
    double min1[3], max1[3], min2[3], max2[3];
    for (int i = 0; i < 3; ++i) {
        min1[i] = a1.v[i] - thresh; max1[i] = a1.v[3 + i] + thresh;
        min2[i] = a2.v[i] - thresh; max2[i] = a2.v[3 + i] + thresh;
    }
    return max1[0] >= min2[0] && min1[0] <= max2[0] && max1[1] >= min2[1] && min1[1] <= max2[1] &&
           max1[2] >= min2[2] && min1[2] <= max2[2];
}
AABB aabb_point(const V3 &p) { AABB a; for (int i = 0; i < 3; ++i) { a.v[i] = p[i]; a.v[3 + i] = p[i]; } return a; }

void barycentric(double &alpha, double &beta, const V3 &a, const V3 &b, const V3 &c, const V3 &p) {  // :488-505
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    double denom = d00 * d11 - d01 * d01;
    beta = (d11 * d20 - d01 * d21) / denom;
    double gamma = (d00 * d21 - d01 * d20) / denom;
    alpha = 1.0 - beta - gamma;
}
void lineline(double &a, double &b, const V3 &A1, const V3 &A2, const V3 &B1, const V3 &B2) {  // :507-525
    V3 B2B1 = B2 - B1, A1B1 = A1 - B1, A2A1 = A2 - A1;
    V3 A2A1xB2B1 = cross(A2A1, B2B1);
    double nA = dot(cross(B2B1, A1B1), A2A1xB2B1);
    double nB = dot(cross(A2A1, A1B1), A2A1xB2B1);
    double d = dot(cross(A2A1, B2B1), A2A1xB2B1);
    a = nA / d; b = nB / d;
}
double linepoint(const V3 &A, const V3 &B, const V3 &P) {  // :527-536
    V3 AP = P - A, AB = B - A;
    double ab2 = dot(AB, AB), apab = dot(AP, AB);
    return apab / ab2;
}
int intersect_square(const V3 &x0, const V3 &dx, const V3 &xa, const V3 &xb, const V3 &xc, double &t) {  // :550-600
    double u1, u2;
    V3 xd = xa + 2.0 * (xc - xa);
    V3 xe = xb + 2.0 * (xc - xb);
    if (intersect_triangle3_inc(x0.data(), dx.data(), xa.data(), xb.data(), xc.data(), &t, &u1, &u2)) return 1;
    if (intersect_triangle3_inc(x0.data(), dx.data(), xb.data(), xd.data(), xc.data(), &t, &u1, &u2)) return 1;
    if (intersect_triangle3_inc(x0.data(), dx.data(), xd.data(), xe.data(), xc.data(), &t, &u1, &u2)) return 1;
    if (intersect_triangle3_inc(x0.data(), dx.data(), xe.data(), xa.data(), xc.data(), &t, &u1, &u2)) return 1;
    t = -1.0;
    return 0;
}

struct Collision {  // btc::Collision, ctor :103-120
    double dist = 0;
    V3 nor1 = mk(0, 0, 0), nor2 = mk(0, 0, 0), pos1 = mk(0, 0, 0), pos2 = mk(0, 0, 0), pos1_ = mk(0, 0, 0);
    int count1 = 0, count2 = 0;
    int verts1[3] = {0, 0, 0}, verts2[3] = {0, 0, 0};
    V3 weights1 = mk(0, 0, 0), weights2 = mk(0, 0, 0);
    int tri1 = -1, tri2 = -1;
    std::vector<int> edge1;
    int edge2 = -1;
    V3 edgeDir = mk(0, 0, 0);
};
typedef std::shared_ptr<Collision> CP;

void perturb(std::vector<double> &verts2, int N, const double *x, double threshold) {  // :648-659 / :1080-1090
    verts2.assign(x, x + 3 * (size_t)N);
    std::mt19937 gen;
    std::uniform_real_distribution<> dis(-1.0, 1.0);
    gen.seed(1);
    for (int i2 = 0; i2 < N; ++i2) {
        double r0 = dis(gen) * threshold * 1e-3;
        double r1 = dis(gen) * threshold * 1e-3;
        double r2 = dis(gen) * threshold * 1e-3;
        verts2[3 * i2] += r0; verts2[3 * i2 + 1] += r1; verts2[3 * i2 + 2] += r2;
    }
}

// boxTriCollision :617-1063 (EOL == false from both CD and CD2, Collisions.cpp:35,75)
void boxTriCollision(std::vector<CP> &collisions, double threshold, const double *whd1, const double *E1, int N,
                     const double *verts2_, int F, const int32_t *faces2, bool EOL, const std::vector<Edge> &edges2) {
    Box B;
    createBox(B, whd1, E1);
    const double *verts1 = B.verts1;
    std::vector<double> verts2;
    perturb(verts2, N, verts2_, threshold);
    std::vector<double> faceNors2;
    createFaceNormals(faceNors2, F, faces2, verts2.data());
    AABB aabbB1, aabbB2;
    build_AABB_B(aabbB1, verts1, 14);
    build_AABB_B(aabbB2, verts2.data(), N);
    AABB aabbF1[24];
    for (int k = 0; k < 24; ++k) aabbF1[k] = aabb_face(verts1, kFaces1 + 3 * k);

    if (!EOL) {  // (A) Vertex2-Triangle1 :675-764
        for (int i2 = 0; i2 < N; ++i2) {
            V3 x2 = col(verts2.data(), i2);
            if (!check_AABB(aabb_point(x2), aabbB1)) continue;
            CP cmin;
            bool inside = true;
            for (int j1 = 0; j1 < 24; ++j1) {
                V3 x1a = col(verts1, kFaces1[3 * j1]);
                V3 nor = col(B.faceNors1.data(), j1);
                V3 dx = x2 - x1a;
                if (dot(nor, dx) > 0.0) { inside = false; break; }
            }
            if (!inside) continue;
            for (int j1 = 0; j1 < 24; ++j1) {
                const int32_t *f1 = kFaces1 + 3 * j1;
                V3 x1a = col(verts1, f1[0]), x1b = col(verts1, f1[1]), x1c = col(verts1, f1[2]);
                V3 nor1 = col(B.faceNors1.data(), j1);
                V3 dx = x2 - x1a;
                double proj = dot(nor1, dx);
                if (proj > 0.0) continue;
                V3 x1 = x2 - proj * nor1;
                dx = x2 - x1;
                double dist = norm(dx);
                if (dist > 5.0 * threshold) continue;
                double u, v;
                barycentric(u, v, x1a, x1b, x1c, x1);
                double w = 1.0 - u - v;
                if (u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w) continue;
                // faceNors2.col(i2): indexed with the VERTEX id (reference quirk, :731). Out of range (i2 >= F) is
                // an out-of-bounds read in the reference; here: zero vector.
                V3 nor2 = i2 < F ? col(faceNors2.data(), i2) : mk(0, 0, 0);
                if (dot(nor2, nor1) < 0.0) nor2 = -nor2;
                CP c = std::make_shared<Collision>();
                c->dist = dist; c->nor1 = nor1; c->nor2 = nor2; c->pos1 = x1; c->pos2 = x2;
                c->count1 = 3; c->count2 = 1;
                c->verts1[0] = f1[0]; c->verts1[1] = f1[1]; c->verts1[2] = f1[2];
                c->verts2[0] = i2; c->verts2[1] = -1; c->verts2[2] = -1;
                c->weights1 = mk(u, v, w); c->weights2 = mk(1.0, 0.0, 0.0);
                c->tri1 = j1; c->tri2 = -1;
                if (!cmin) cmin = c;
                else if (c->dist < cmin->dist) cmin = c;
            }
            if (cmin) collisions.push_back(cmin);
        }
    }

    // (B) Vertex1-Triangle2 :771-845
    for (int i1 = 0; i1 < 8; ++i1) {
        V3 x1 = col(verts1, i1);
        if (!check_AABB(aabb_point(x1), aabbB2)) continue;
        CP cmin;
        V3 nor1 = col(B.vertNors1.data(), i1);
        for (int j2 = 0; j2 < F; ++j2) {
            const int32_t *f2 = faces2 + 3 * j2;
            V3 x2a = col(verts2.data(), f2[0]), x2b = col(verts2.data(), f2[1]), x2c = col(verts2.data(), f2[2]);
            V3 nor2 = col(faceNors2.data(), j2);
            if (dot(nor1, nor2) < 0.0) nor2 = -nor2;
            V3 dx = x1 - x2a;
            double proj = dot(dx, nor2);
            if (proj < 0.0) continue;
            V3 x2 = x1 - proj * nor2;
            dx = x2 - x1;
            double dist = norm(dx);
            if (dist > 5.0 * threshold) continue;
            double u, v;
            barycentric(u, v, x2a, x2b, x2c, x1);
            double w = 1.0 - u - v;
            if (u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w) continue;
            CP c = std::make_shared<Collision>();
            c->dist = dist; c->nor1 = nor1; c->nor2 = nor2; c->pos1 = x1; c->pos2 = x2;
            c->count1 = 1; c->count2 = 3;
            c->verts1[0] = i1; c->verts1[1] = -1; c->verts1[2] = -1;
            c->verts2[0] = f2[0]; c->verts2[1] = f2[1]; c->verts2[2] = f2[2];
            c->weights1 = mk(1.0, 0.0, 0.0); c->weights2 = mk(u, v, w);
            c->edge1.push_back(kVertEdges1[i1][0]); c->edge1.push_back(kVertEdges1[i1][1]); c->edge1.push_back(kVertEdges1[i1][2]);
            c->tri1 = -1; c->tri2 = j2;
            if (!cmin) cmin = c;
            else if (c->dist < cmin->dist) cmin = c;
        }
        if (cmin) collisions.push_back(cmin);
    }

    // (C) Edge2-Edge1 :849-1017
    for (int k2 = 0; k2 < (int)edges2.size(); ++k2) {
        const Edge &e2 = edges2[k2];
        V3 x2a = col(verts2.data(), e2.verts[0]), x2b = col(verts2.data(), e2.verts[1]);
        V3 dx2 = x2b - x2a;
        double len2 = norm(dx2);
        V3 nor2 = normalized(e2.normals[0] + e2.normals[1]);
        AABB aabbE2k = aabb_edge(verts2.data(), e2.verts[0], e2.verts[1]);
        for (int k1 = 0; k1 < 12; ++k1) {
            const Edge &e1 = B.edges1[k1];
            if (e1.angle < M_PI / 6.0) continue;
            bool aabbC = check_AABB(aabbF1[e1.faces[0]], aabbE2k);
            if (!aabbC) {
                bool aabbD = check_AABB(aabbF1[e1.faces[1]], aabbE2k);
                if (!aabbD) continue;
            }
            V3 x1a = col(verts1, e1.verts[0]), x1b = col(verts1, e1.verts[1]);
            V3 dx1 = x1b - x1a;
            double len1 = norm(dx1);
            V3 tan1 = dx1 / len1;
            double threshAng = 2.0 * M_PI / 180.0;
            double angle = std::acos(dot(tan1, nor2));
            if (std::fabs(angle) < threshAng || std::fabs(M_PI - angle) < threshAng) continue;
            angle = std::acos(dot(tan1, dx2) / len2);
            if (std::fabs(angle) < threshAng || std::fabs(M_PI - angle) < threshAng) continue;
            V3 nor = normalized(cross(dx1, dx2));
            V3 x1c = col(verts1, e1.verts[2]), x1d = col(verts1, e1.verts[3]);
            const V3 &n1c = e1.normals[0], &n1d = e1.normals[1];
            V3 nor1 = normalized(n1c + n1d);
            if (dot(nor, nor1) < 0.0) nor = -nor;
            double angleCD = std::acos(dot(n1c, n1d));
            double angleCN = std::acos(dot(n1c, nor));
            if (angleCD < 0.0) { angleCD = -angleCD; angleCN = -angleCN; }
            if (angleCN < -threshAng || angleCN - angleCD > threshAng) continue;
            double u2c, u2d;
            int i2c = intersect_square(x2a, dx2, x1a, x1b, x1c, u2c);
            int i2d = intersect_square(x2a, dx2, x1b, x1a, x1d, u2d);
            i2c = i2c && (0.0 <= u2c && u2c <= 1.0);
            i2d = i2d && (0.0 <= u2d && u2d <= 1.0);
            double u1, u2;
            lineline(u1, u2, x1a, x1b, x2a, x2b);
            double thresh1 = 1.0 * threshold / len1;
            double thresh2 = 1.0 * threshold / len2;
            if (u1 < -thresh1 || u1 > 1.0 + thresh1 || u2 < -thresh2 || u2 > 1.0 + thresh2) continue;
            V3 x1, x2, dx;
            if (i2c && i2d) {
                x1 = (1.0 - u1) * x1a + u1 * x1b;
                x2 = (1.0 - u2) * x2a + u2 * x2b;
            } else if (i2c && !i2d) {
                dx = x2a - x1a;
                if (dot(dx, n1c) < 0.0) u2 = std::max(0.0, std::min(u2c, u2));
                else u2 = std::max(u2c, std::min(1.0, u2));
                x2 = (1.0 - u2) * x2a + u2 * x2b;
                u1 = linepoint(x1a, x1b, x2);
                x1 = (1.0 - u1) * x1a + u1 * x1b;
            } else if (!i2c && i2d) {
                dx = x2a - x1b;
                if (dot(dx, n1d) < 0.0) u2 = std::max(0.0, std::min(u2d, u2));
                else u2 = std::max(u2d, std::min(1.0, u2));
                x2 = (1.0 - u2) * x2a + u2 * x2b;
                u1 = linepoint(x1a, x1b, x2);
                x1 = (1.0 - u1) * x1a + u1 * x1b;
            } else {
                continue;
            }
            if (u1 < -thresh1 || u1 > 1.0 + thresh1 || u2 < -thresh2 || u2 > 1.0 + thresh2) continue;
            dx = x2 - x1;
            double thresh = 2.0 * threshold;
            if (dot(dx, dx) > thresh * thresh) continue;
            CP c = std::make_shared<Collision>();
            c->dist = norm(dx);
            c->nor1 = nor1; c->nor2 = nor; c->pos1 = x1; c->pos2 = x2;
            c->count1 = 2; c->count2 = 2;
            c->verts1[0] = e1.verts[0]; c->verts1[1] = e1.verts[1]; c->verts1[2] = -1;
            c->verts2[0] = e2.verts[0]; c->verts2[1] = e2.verts[1]; c->verts2[2] = -1;
            c->weights1 = mk(1.0 - u1, u1, 0.0);
            c->weights2 = mk(1.0 - u2, u2, 0.0);
            c->edge1.push_back(k1);
            c->edge2 = k2;
            c->edgeDir = tan1;
            collisions.push_back(c);
        }
    }

    // (D) :1022-1052
    for (int i1 = 0; i1 < 8; ++i1) {
        V3 x1 = col(verts1, i1);
        int kmin = -1;
        double dmin = 1e9;
        for (int k = 0; k < (int)collisions.size(); ++k) {
            if (collisions[k]->count1 == 1 && collisions[k]->verts1[0] == i1) {
                V3 dx = collisions[k]->pos2 - x1;
                double d = dot(dx, dx);
                if (d < dmin) { kmin = k; dmin = d; }
            }
        }
        if (kmin != -1) {
            std::vector<int> dlist;
            for (int k = 0; k < (int)collisions.size(); ++k)
                if (collisions[k]->count1 == 1 && collisions[k]->verts1[0] == i1 && k != kmin) dlist.push_back(k);
            for (int kdel : dlist) {  // forward iteration with swap-with-back, literally as written
                collisions[kdel] = collisions.back();
                collisions.pop_back();
            }
        }
    }
    // (E) :1057-1060
    double snapDepth = 0.1 * threshold;
    for (auto &c : collisions) c->pos1_ = c->pos1 - snapDepth * c->nor1;
}

// pointTriCollision :1067-1224
void pointTriCollision(std::vector<CP> &collisions, double threshold, int P, const double *verts1, const double *norms1,
                       int N, const double *verts2_, int F, const int32_t *faces2, bool EOL) {
    std::vector<double> verts2;
    perturb(verts2, N, verts2_, threshold);
    std::vector<double> faceNors2;
    createFaceNormals(faceNors2, F, faces2, verts2.data());
    AABB aabbB2;
    build_AABB_B(aabbB2, verts2.data(), N);
    if (EOL) {
        for (int i2 = 0; i2 < N; ++i2) {
            V3 x2 = col(verts2.data(), i2);
            CP cmin;
            for (int i1 = 0; i1 < P; ++i1) {
                V3 x1 = col(verts1, i1);
                V3 dx = x2 - x1;
                double dist = norm(dx);
                if (dist < threshold) {
                    V3 nor1 = col(norms1, i1);
                    CP c = std::make_shared<Collision>();
                    c->dist = dist; c->nor1 = nor1; c->nor2 = nor1; c->pos1 = x1; c->pos2 = x2;
                    c->count1 = 3; c->count2 = 1;
                    c->verts1[0] = i1; c->verts1[1] = -1; c->verts1[2] = -1;
                    c->verts2[0] = i2; c->verts2[1] = -1; c->verts2[2] = -1;
                    c->weights1 = mk(1.0, 0.0, 0.0); c->weights2 = mk(1.0, 0.0, 0.0);
                    c->tri1 = -1; c->tri2 = -1;
                    if (!cmin) cmin = c;
                    else if (c->dist < cmin->dist) cmin = c;
                }
            }
            if (cmin) collisions.push_back(cmin);
        }
    }
    for (int i1 = 0; i1 < P; ++i1) {
        V3 x1 = col(verts1, i1);
        if (!check_AABB(aabb_point(x1), aabbB2)) continue;
        CP cmin;
        V3 nor1 = col(norms1, i1);
        for (int j2 = 0; j2 < F; ++j2) {
            const int32_t *f2 = faces2 + 3 * j2;
            V3 x2a = col(verts2.data(), f2[0]), x2b = col(verts2.data(), f2[1]), x2c = col(verts2.data(), f2[2]);
            V3 nor2 = col(faceNors2.data(), j2);
            if (dot(nor1, nor2) < 0.0) nor2 = -nor2;
            V3 dx = x1 - x2a;
            double proj = dot(dx, nor2);
            if (proj < 0.0) continue;
            V3 x2 = x1 - proj * nor2;
            dx = x2 - x1;
            double dist = norm(dx);
            if (dist > 5.0 * threshold) continue;
            double u, v;
            barycentric(u, v, x2a, x2b, x2c, x1);
            double w = 1.0 - u - v;
            if (u < 0.0 || 1.0 < u || v < 0.0 || 1.0 < v || w < 0.0 || 1.0 < w) continue;
            CP c = std::make_shared<Collision>();
            c->dist = dist; c->nor1 = nor1; c->nor2 = nor2; c->pos1 = x1; c->pos2 = x2;
            c->count1 = 1; c->count2 = 3;
            c->verts1[0] = i1; c->verts1[1] = -1; c->verts1[2] = -1;
            c->verts2[0] = f2[0]; c->verts2[1] = f2[1]; c->verts2[2] = f2[2];
            c->weights1 = mk(1.0, 0.0, 0.0); c->weights2 = mk(u, v, w);
            c->tri1 = -1; c->tri2 = j2;
            if (!cmin) cmin = c;
            else if (c->dist < cmin->dist) cmin = c;
        }
        if (cmin) collisions.push_back(cmin);
    }
    double snapDepth = 0.1 * threshold;
    for (auto &c : collisions) c->pos1_ = c->pos1 - snapDepth * c->nor1;
}

void to_pod(const Collision &c, eolc_contact &o) {
    std::memset(&o, 0, sizeof(o));
    o.dist = c.dist;
    for (int i = 0; i < 3; ++i) {
        o.nor1[i] = c.nor1[i]; o.nor2[i] = c.nor2[i]; o.pos1[i] = c.pos1[i]; o.pos2[i] = c.pos2[i]; o.pos1_[i] = c.pos1_[i];
        o.weights1[i] = c.weights1[i]; o.weights2[i] = c.weights2[i]; o.edgeDir[i] = c.edgeDir[i];
        o.verts1[i] = c.verts1[i]; o.verts2[i] = c.verts2[i];
        o.edge1[i] = i < (int)c.edge1.size() ? c.edge1[i] : -1;
    }
    o.count1 = c.count1; o.count2 = c.count2; o.tri1 = c.tri1; o.tri2 = c.tri2;
    o.n_edge1 = (int)c.edge1.size(); o.edge2 = c.edge2;
}

}  // namespace

extern "C" {

// CD  (Collisions.cpp:11-53): point_eol_flag = 1, remap = 1 ;  CD2 (:55-78): 0, 0
int oracle_cd(int N, int F, const int32_t *face_nodes, const double *x, double threshold, int n_points,
              const double *pxyz, const double *pnorms, int n_boxes, const double *box_whd, const double *box_E,
              int point_eol_flag, int remap, eolc_contact *out, int capacity, int *n_out) {
    std::vector<CP> cls;
    pointTriCollision(cls, threshold, n_points, pxyz, pnorms, N, x, F, face_nodes, point_eol_flag != 0);
    size_t c = cls.size();
    for (int b = 0; b < n_boxes; ++b) {
        std::vector<CP> clst;
        std::vector<Edge> edges2;
        createEdges(edges2, F, face_nodes, x);  // boxTriCollision.cpp:612-613: rebuilt per box, UNPERTURBED verts
        boxTriCollision(clst, threshold, box_whd + 3 * b, box_E + 16 * b, N, x, F, face_nodes, false, edges2);
        cls.insert(cls.end(), clst.begin(), clst.end());
        if (remap) {  // Collisions.cpp:39-48 ; Box::num_points = 8, num_edges = 12 (Box.cpp:76-77)
            for (; c < cls.size(); c++) {
                if (cls[c]->count1 == 1 && cls[c]->count2 == 3)
                    cls[c]->verts1[0] = n_points + (b * 8) + (b * 12) + cls[c]->verts1[0];
                for (size_t e = 0; e < cls[c]->edge1.size(); e++)
                    cls[c]->edge1[e] = n_points + (b * 8) + (b * 12) + (8 + cls[c]->edge1[e]);
            }
        }
    }
    *n_out = (int)cls.size();
    if ((int)cls.size() > capacity) return -3;
    for (size_t i = 0; i < cls.size(); ++i) to_pod(*cls[i], out[i]);
    return 0;
}

int oracle_cd_edges(int N, int F, const int32_t *face_nodes, const double *x, int32_t *out6E, double *normals6E) {
    (void)N;
    std::vector<Edge> edges;
    createEdges(edges, F, face_nodes, x);
    if (out6E)
        for (size_t k = 0; k < edges.size(); ++k) {
            for (int i = 0; i < 4; ++i) out6E[6 * k + i] = edges[k].verts[i];
            out6E[6 * k + 4] = edges[k].faces[0]; out6E[6 * k + 5] = edges[k].faces[1];
            if (normals6E)
                for (int i = 0; i < 3; ++i) { normals6E[6 * k + i] = edges[k].normals[0][i]; normals6E[6 * k + 3 + i] = edges[k].normals[1][i]; }
        }
    return (int)edges.size();
}

void oracle_perturbation(int n, double threshold, double *out3n) {  // out = dis(gen)*threshold*1e-3 stream
    std::vector<double> z(3 * (size_t)n, 0.0), v;
    perturb(v, n, z.data(), threshold);
    std::memcpy(out3n, v.data(), sizeof(double) * 3 * (size_t)n);
}
void oracle_rng_raw(int n, double *out) {  // dis(gen) itself, for the RNG known-answer test
    std::mt19937 gen; std::uniform_real_distribution<> dis(-1.0, 1.0); gen.seed(1);
    for (int i = 0; i < n; ++i) out[i] = dis(gen);
}
int oracle_raytri(const double *orig, const double *dir, const double *v0, const double *v1, const double *v2, double *tuv) {
    return intersect_triangle3_inc(orig, dir, v0, v1, v2, tuv, tuv + 1, tuv + 2);
}
void oracle_box(const double *whd, const double *E1, double *verts14x3, double *faceNors24x3, double *vertNors14x3, double *edgeAngles12) {
    Box B; createBox(B, whd, E1);
    std::memcpy(verts14x3, B.verts1, sizeof(B.verts1));
    std::memcpy(faceNors24x3, B.faceNors1.data(), sizeof(double) * 72);
    std::memcpy(vertNors14x3, B.vertNors1.data(), sizeof(double) * 42);
    for (int k = 0; k < 12; ++k) edgeAngles12[k] = B.edges1[k].angle;
}

}  // extern "C"
