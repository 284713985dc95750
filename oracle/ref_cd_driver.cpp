// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points around the REFERENCE'S OWN collision code.  oracle/Makefile compiles this file together with
//   /root/reference/src/boxTriCollision.cpp, /root/reference/src/Collisions.cpp, /root/reference/src/raytri.cpp
// UNMODIFIED, from where they lie, against oracle/mini_eigen (a functional stand-in for the Eigen subset they use;
// Eigen itself is not in this image) into oracle/_ref/libbtc_ref.so (and libbtc_ref_scalar.so with the other
// reduction order, see mini_eigen/Eigen/Dense).  So btc::createEdges (boxTriCollision.cpp:141-231), createBox
// (:400-420), boxTriCollision (:617-1063), pointTriCollision (:1067-1224) and CD / CD2 (Collisions.cpp:11-78) run
// here exactly as the reference wrote them; tests/test_cd_ref_pin.py holds oracle/cd_ref.cpp (the restatement the
// GPU is compared with at sizes where the reference's `int` edge hash overflows) to them bit for bit.
//
// What this file supplies itself: (1) flat arrays -> the ArcSim pointer mesh / Obstacles the reference entry points
// take, and btc::Collision -> the POD record of include/eolc.h; (2) link-time definitions of members that live in
// reference sources OUTSIDE the path (Obstacles::Obstacles, Box::Box and its virtuals, uuid_src) — CD / CD2 only read
// the data members those constructors set.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Collisions.h"  // reference header (pulls external/ArcSim/mesh.hpp, boxTriCollision.h, Obstacles.h)
#include "Box.h"
#include "Points.h"
#include "../include/eolc.h"

// ---- definitions the reference keeps in sources outside the path ---------------------------------------------------
int uuid_src = 0;                                             // external/ArcSim/mesh.cpp
Obstacles::Obstacles() : num_boxes(0) { points = std::make_shared<Points>(); }  // Obstacles.cpp:14-18
Box::Box(const std::shared_ptr<Shape> shape, std::string en)  // Box.cpp:74-78 sets num_points(8), num_edges(12)
    : num_points(8), num_edges(12), exportName(en), boxShape(shape) {}
Box::~Box() {}
void Box::step(const double) {}
int Box::getBrenderCount() const { return 0; }
std::vector<std::string> Box::getBrenderNames() const { return std::vector<std::string>(); }
void Box::exportBrender(std::vector<std::shared_ptr<std::ofstream> >) const {}

namespace btc {  // non-static globals / helpers of boxTriCollision.cpp (:388-420), read back for the box-table test
extern Eigen::Matrix<double, 4, 14> verts1;
extern Eigen::MatrixXd faceNors1;
extern Eigen::MatrixXd vertNors1;
void createBox(std::vector<std::shared_ptr<Edge> > &edges1, const Eigen::Vector3d &whd1, const Eigen::Matrix4d &E1);
}

namespace {

void to_pod(const btc::Collision &c, eolc_contact &o) {
    std::memset(&o, 0, sizeof(o));
    o.dist = c.dist;
    for (int i = 0; i < 3; ++i) {
        o.nor1[i] = c.nor1(i); o.nor2[i] = c.nor2(i); o.pos1[i] = c.pos1(i); o.pos2[i] = c.pos2(i); o.pos1_[i] = c.pos1_(i);
        o.weights1[i] = c.weights1(i); o.weights2[i] = c.weights2(i); o.edgeDir[i] = c.edgeDir(i);
        o.verts1[i] = c.verts1(i); o.verts2[i] = c.verts2(i);
        o.edge1[i] = i < (int)c.edge1.size() ? c.edge1[i] : -1;
    }
    o.count1 = c.count1; o.count2 = c.count2; o.tri1 = c.tri1; o.tri2 = c.tri2;
    o.n_edge1 = (int)c.edge1.size(); o.edge2 = c.edge2;
}

int emit(const std::vector<std::shared_ptr<btc::Collision> > &cls, eolc_contact *out, int capacity, int *n_out) {
    *n_out = (int)cls.size();
    if ((int)cls.size() > capacity) return -3;
    for (size_t i = 0; i < cls.size(); ++i) to_pod(*cls[i], out[i]);
    return 0;
}

// The pointer mesh CD / CD2 read (Collisions.cpp:21-27): nodes[i]->x, ->EoL, faces[k]->v[j]->node->index.
struct FlatMesh {
    Mesh mesh;
    std::vector<Node> nodes;
    std::vector<Vert> verts;
    std::vector<Face> faces;
    FlatMesh(int N, int F, const int32_t *face_nodes, const double *x, const int32_t *eol) : nodes(N), verts(N), faces(F) {
        for (int i = 0; i < N; ++i) {
            nodes[i].x = Vec3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
            nodes[i].index = i;
            nodes[i].EoL = eol ? eol[i] != 0 : false;
            nodes[i].verts.push_back(&verts[i]);
            verts[i].node = &nodes[i];
            verts[i].index = i;
            mesh.nodes.push_back(&nodes[i]);
            mesh.verts.push_back(&verts[i]);
        }
        for (int k = 0; k < F; ++k) {
            for (int j = 0; j < 3; ++j) faces[k].v[j] = &verts[face_nodes[3 * k + j]];
            faces[k].index = k;
            mesh.faces.push_back(&faces[k]);
        }
    }
};

Eigen::MatrixXd mat3xn(const double *p, int n) {
    Eigen::MatrixXd m(3, n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) m(j, i) = p[3 * i + j];
    return m;
}
Eigen::MatrixXi imat3xn(const int32_t *p, int n) {
    Eigen::MatrixXi m(3, n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) m(j, i) = p[3 * i + j];
    return m;
}
Eigen::Matrix4d mat4(const double *colmajor16) {
    Eigen::Matrix4d E;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) E(i, j) = colmajor16[4 * j + i];
    return E;
}

}  // namespace

extern "C" {

// 0 = vectorised reduction order (p0+p1)+p2, 1 = scalar unroller p0+(p1+p2)
int ref_redux_order(void) {
#ifdef MINI_EIGEN_SCALAR_REDUX
    return 1;
#else
    return 0;
#endif
}

// The reference's `int hash = kmin + (3F+1)*kmax` (boxTriCollision.cpp:167-169) is defined only while it fits an int.
int ref_btc_hash_defined(int N, int F) { return (3LL * F + 1) * (long long)N + N < 2147483647LL; }

// CD (which = 1, Collisions.cpp:11-53) or CD2 (which = 0, :55-78) of the reference, on a mesh rebuilt from flat arrays.
int ref_cd(int N, int F, const int32_t *face_nodes, const double *x, const int32_t *eol, double threshold, int n_points,
           const double *pxyz, const double *pnorms, int n_boxes, const double *box_whd, const double *box_E, int which,
           eolc_contact *out, int capacity, int *n_out) {
    FlatMesh fm(N, F, face_nodes, x, eol);
    auto obs = std::make_shared<Obstacles>();
    obs->cdthreshold = threshold;
    obs->num_boxes = n_boxes;
    obs->points->num_points = n_points;
    obs->points->pxyz = mat3xn(pxyz, n_points);
    obs->points->norms = mat3xn(pnorms, n_points);
    for (int b = 0; b < n_boxes; ++b) {
        auto box = std::make_shared<Box>(std::shared_ptr<Shape>(), "box");
        box->dim = Eigen::Vector3d(box_whd[3 * b], box_whd[3 * b + 1], box_whd[3 * b + 2]);
        box->E1 = mat4(box_E + 16 * b);
        obs->boxes.push_back(box);
    }
    std::vector<std::shared_ptr<btc::Collision> > cls;
    if (which) CD(fm.mesh, obs, cls); else CD2(fm.mesh, obs, cls);
    return emit(cls, out, capacity, n_out);
}

// btc::boxTriCollision (first overload, :603-615: builds the edge table itself) for one box.
int ref_btc_box(int N, int F, const int32_t *face_nodes, const double *x, double threshold, const double *whd,
                const double *E1, int EOL, eolc_contact *out, int capacity, int *n_out) {
    std::vector<std::shared_ptr<btc::Collision> > cls;
    Eigen::VectorXi isEOL;
    btc::boxTriCollision(cls, threshold, Eigen::Vector3d(whd[0], whd[1], whd[2]), mat4(E1), mat3xn(x, N), imat3xn(face_nodes, F), isEOL, EOL != 0);
    return emit(cls, out, capacity, n_out);
}

// btc::pointTriCollision (:1067-1224)
int ref_btc_points(int N, int F, const int32_t *face_nodes, const double *x, double threshold, int n_points,
                   const double *pxyz, const double *pnorms, int EOL, eolc_contact *out, int capacity, int *n_out) {
    std::vector<std::shared_ptr<btc::Collision> > cls;
    btc::pointTriCollision(cls, threshold, mat3xn(pxyz, n_points), mat3xn(pnorms, n_points), mat3xn(x, N), imat3xn(face_nodes, F), EOL != 0);
    return emit(cls, out, capacity, n_out);
}

// btc::createEdges (:141-231): per edge verts(4), faces(2) -> out6E; normals[0], normals[1] -> normals6E; internal, angle.
int ref_btc_edges(int N, int F, const int32_t *face_nodes, const double *x, int32_t *out6E, double *normals6E, int32_t *internal, double *angle) {
    std::vector<std::shared_ptr<btc::Edge> > edges;
    btc::createEdges(edges, imat3xn(face_nodes, F), mat3xn(x, N));
    for (size_t k = 0; k < edges.size(); ++k) {
        for (int i = 0; i < 4; ++i) out6E[6 * k + i] = edges[k]->verts(i);
        out6E[6 * k + 4] = edges[k]->faces(0); out6E[6 * k + 5] = edges[k]->faces(1);
        for (int i = 0; i < 3; ++i) { normals6E[6 * k + i] = edges[k]->normals[0](i); normals6E[6 * k + 3 + i] = edges[k]->normals[1](i); }
        if (internal) internal[k] = edges[k]->internal ? 1 : 0;
        if (angle) angle[k] = edges[k]->angle;
    }
    return (int)edges.size();
}

// btc::createBox (:400-420) and the globals it fills.
void ref_btc_boxtables(const double *whd, const double *E1, double *verts14x3, double *faceNors24x3, double *vertNors14x3, double *edgeAngles12) {
    std::vector<std::shared_ptr<btc::Edge> > edges1;
    btc::createBox(edges1, Eigen::Vector3d(whd[0], whd[1], whd[2]), mat4(E1));
    for (int k = 0; k < 14; ++k) for (int i = 0; i < 3; ++i) { verts14x3[3 * k + i] = btc::verts1(i, k); vertNors14x3[3 * k + i] = btc::vertNors1(i, k); }
    for (int k = 0; k < 24; ++k) for (int i = 0; i < 3; ++i) faceNors24x3[3 * k + i] = btc::faceNors1(i, k);
    for (int k = 0; k < 12; ++k) edgeAngles12[k] = edges1[k]->angle;
}

}  // extern "C"
