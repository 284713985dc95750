// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry point around the REFERENCE'S OWN Constraints::fill (Constraints.cpp:122-513), which calls CD2 itself (:423) and turns the
// contact list into the inequality rows (:424-468) and the fixed corners into equality rows (:470-497).  oracle/Makefile compiles this
// file together with /root/reference/src/Constraints.cpp, Collisions.cpp, boxTriCollision.cpp, raytri.cpp, Box.cpp, Rigid.cpp,
// Obstacles.cpp, Points.cpp, Shape.cpp, BrenderManager.cpp, conversions.cpp and ArcSim's mesh / geometry / util / vectors /
// transformation .cpp UNMODIFIED against oracle/mini_eigen into oracle/_ref/libconstraints_ref.so (same -D__Cloth__ / prelude /
// -fpermissive as libforces_ref.so, see mini_eigen/shim/forces_prelude.h).  Here Box::Box, Obstacles::Obstacles are the reference's own.
// tests/test_constraints_rows.py holds eolc_constraints_contact_rows / eolc_constraints_fixed_rows (and the device rows) to it.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Constraints.h"
#include "Obstacles.h"
#include "Box.h"
#include "Points.h"
#include "FixedList.h"
#include "external/ArcSim/geometry.hpp"
#include "external/ArcSim/util.hpp"

namespace {
struct Run {
    Mesh mesh;
    Material material;
    Constraints cons;
    ~Run() { delete_mesh(mesh); }
};
Eigen::Matrix4d mat4(const double *colmajor16) {
    Eigen::Matrix4d E;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) E(i, j) = colmajor16[4 * j + i];
    return E;
}
}  // namespace

extern "C" {

// corner_id: per node, >= 0 marks the node EoL with Node::cornerID = that obstacle POINT (must be < n_points: the reference then adds
// one inequality row -n and two equality rows for it before the contact rows, Constraints.cpp:144-211); NULL / -1 = Lagrangian.
// fixed_c: 4 x 6 doubles (FixedList::c1..c4), fixed_ci: the four node indices.  v: node velocities (3N).
void *ref_constraints_fill(int N, int F, const int32_t *face_nodes, const double *x, const double *X, const double *v,
                           const int32_t *corner_id, double threshold, int n_points, const double *pxyz, const double *pnorms,
                           int n_boxes, const double *box_whd, const double *box_E, const double *fixed_c, const int32_t *fixed_ci,
                           double h) {
    Run *R = new Run;
    std::memset(&R->material, 0, sizeof(R->material));
    // the mesh as Cloth::build makes it (Cloth.cpp:63-132), see ref_forces_driver.cpp
    for (int i = 0; i < N; ++i) {
        R->mesh.add(new Vert(Vec3(X[2 * i], X[2 * i + 1], 0.0), Vec3(0)));
        const Vec3 xi(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        R->mesh.add(new Node(xi, xi, Vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]), 0, 0, false));
        connect(R->mesh.verts.back(), R->mesh.nodes.back());
    }
    for (int k = 0; k < F; ++k)
        R->mesh.add(new Face(R->mesh.verts[face_nodes[3 * k]], R->mesh.verts[face_nodes[3 * k + 1]], R->mesh.verts[face_nodes[3 * k + 2]],
                             Mat3x3(1), Mat3x3(0), &R->material, 0));
    mark_nodes_to_preserve(R->mesh);
    compute_ms_data(R->mesh);          // also compute_ws_data: face->n, which the EoL rows read (:176-182)
    set_indices(R->mesh);
    int eol_count = 0;
    if (corner_id)
        for (int i = 0; i < N; ++i)
            if (corner_id[i] >= 0) {
                R->mesh.nodes[i]->EoL = true;
                R->mesh.nodes[i]->EoL_index = eol_count++;
                R->mesh.nodes[i]->cornerID = corner_id[i];
            }
    R->mesh.EoL_Count = eol_count;

    auto obs = std::make_shared<Obstacles>();           // Obstacles.cpp:14-18
    obs->cdthreshold = threshold;
    obs->num_boxes = n_boxes;
    obs->points->num_points = n_points;
    obs->points->pxyz.resize(3, n_points);
    obs->points->norms.resize(3, n_points);
    for (int p = 0; p < n_points; ++p) for (int j = 0; j < 3; ++j) { obs->points->pxyz(j, p) = pxyz[3 * p + j]; obs->points->norms(j, p) = pnorms[3 * p + j]; }
    for (int b = 0; b < n_boxes; ++b) {
        auto box = std::make_shared<Box>(std::shared_ptr<Shape>(), "box");     // Box.cpp:74-95
        box->dim = Eigen::Vector3d(box_whd[3 * b], box_whd[3 * b + 1], box_whd[3 * b + 2]);
        box->E1 = mat4(box_E + 16 * b);
        obs->boxes.push_back(box);
    }
    auto fs = std::make_shared<FixedList>();
    Eigen::VectorXd *c[4] = {&fs->c1, &fs->c2, &fs->c3, &fs->c4};
    int *ci[4] = {&fs->c1i, &fs->c2i, &fs->c3i, &fs->c4i};
    for (int k = 0; k < 4; ++k) {
        for (int j = 0; j < 6; ++j) (*c[k])(j) = fixed_c[6 * k + j];
        *ci[k] = fixed_ci[k];
    }
    R->cons.init(obs);                                   // Scene.cpp: constraints->init(obs) before the first fill
    R->cons.fill(R->mesh, obs, fs, h, false);            // the call of Cloth.cpp:359
    return R;
}
void ref_constraints_free(void *p) { delete static_cast<Run *>(p); }

static Eigen::SparseMatrix<double> &mat_of(void *p, int which) { Run *R = static_cast<Run *>(p); return which ? R->cons.Aeq : R->cons.Aineq; }
int ref_constraints_rows(void *p, int which) { return (int)mat_of(p, which).rows(); }
int ref_constraints_cols(void *p, int which) { return (int)mat_of(p, which).cols(); }
int64_t ref_constraints_nnz(void *p, int which) { return (int64_t)mat_of(p, which).nonZeros(); }
const int *ref_constraints_outer(void *p, int which) { return mat_of(p, which).outerIndexPtr(); }
const int *ref_constraints_inner(void *p, int which) { return mat_of(p, which).innerIndexPtr(); }
const double *ref_constraints_vals(void *p, int which) { return mat_of(p, which).valuePtr(); }
const double *ref_constraints_b(void *p, int which) { Run *R = static_cast<Run *>(p); return which ? R->cons.beq.data() : R->cons.bineq.data(); }
int ref_constraints_flags(void *p) { Run *R = static_cast<Run *>(p); return (R->cons.hasFixed ? 1 : 0) | (R->cons.hasCollisions ? 2 : 0); }

}  // extern "C"
