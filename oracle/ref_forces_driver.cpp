// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points around the REFERENCE'S OWN Forces::fill.  oracle/Makefile compiles this file together with
//   /root/reference/src/Forces.cpp, UtilEOL.cpp, conversions.cpp, ComputeMembrane.cpp, ComputeBending.cpp, ComputeInertial.cpp and
//   /root/reference/src/external/ArcSim/{mesh,geometry,util,vectors,transformation}.cpp
// UNMODIFIED, from where they lie, against oracle/mini_eigen (functional stand-in for the Eigen subset they use; Eigen itself is
// not in this image) into oracle/_ref/libforces_ref.so.  So faceBasedF / edgeBasedF with their EOL branches (Forces.cpp:331-520,
// 685-910, fillEOL* :177-329, :580-683), poldec (:33-52), deform_grad (UtilEOL.cpp:13-28), Forces::fill itself (:912-930), and on
// the mesh side Mesh::add(Face*) with its edge creation (external/ArcSim/mesh.cpp:356-378), connect, compute_ms_data /
// compute_ws_data (mesh.cpp:135-243, geometry.cpp:302-316) run here exactly as the reference wrote them.
// tests/test_forces_ref_pin.py holds oracle/forces_ref.cpp (the Eigen-free restatement the GPU is compared with at full size) and
// the library's edge-stencil / normals routines to them.
//
// Two things the build needs that are not in those sources (oracle/mini_eigen/shim/forces_prelude.h says why):
//   * src/Cloth.h cannot be parsed by g++ (`extern struct Material {`): its include guard is pre-defined and the prelude declares
//     struct Material field for field;
//   * vectors.cpp needs -fpermissive (an explicit instantiation without a definition).
// The same driver linked with adapter/Forces_fill_b200.cpp INSTEAD of Forces.cpp gives oracle/_ref/libadapter_forces.so: the
// reference-side drop-in body of Forces::fill, executed (reference Mesh / Forces types -> include/eolc_host.hpp -> C ABI -> GPU).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "Forces.h"                      // reference header: class Forces (pulls mesh.hpp, <Eigen/Dense>, <Eigen/Sparse>)
#include "external/ArcSim/geometry.hpp"  // compute_ms_data / compute_ws_data
#include "external/ArcSim/util.hpp"

// adapter builds only (adapter/Forces_fill_b200.cpp): drops what the adapter holds for a Forces object before the object dies
extern "C" void eolc_adapter_forces_release(const void *forces_this) __attribute__((weak));

namespace {

struct Run {
    Mesh mesh;
    Material material;
    Forces forces;
    double grav[3], h;
    ~Run() {
        if (eolc_adapter_forces_release) eolc_adapter_forces_release(&forces);
        delete_mesh(mesh);
    }
};

// The mesh the way Cloth::build makes it (Cloth.cpp:63-129): one Vert + one Node per grid point, connect(), faces through
// Mesh::add(Face*) — which creates mesh.edges in ArcSim's order — all faces pointing at the cloth's material, then
// mark_nodes_to_preserve + compute_ms_data (Cloth.cpp:131-132).  triangulateARC (io.cpp:244-271) is not used: for three verts it only
// picks which of them comes first, and the caller's face_nodes already fix that.
void build_mesh(Run &R, int N, int F, const int32_t *face_nodes, const double *x, const double *X, const int32_t *eol_index) {
    for (int i = 0; i < N; ++i) {
        R.mesh.add(new Vert(Vec3(X[2 * i], X[2 * i + 1], 0.0), Vec3(0)));
        const Vec3 xi(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        R.mesh.add(new Node(xi, xi, Vec3(0), 0, 0, false));
        connect(R.mesh.verts.back(), R.mesh.nodes.back());
    }
    for (int k = 0; k < F; ++k)
        R.mesh.add(new Face(R.mesh.verts[face_nodes[3 * k]], R.mesh.verts[face_nodes[3 * k + 1]], R.mesh.verts[face_nodes[3 * k + 2]],
                            Mat3x3(1), Mat3x3(0), &R.material, 0));
    mark_nodes_to_preserve(R.mesh);
    compute_ms_data(R.mesh);
    // EoL flags as the preprocessor leaves them: Node::EoL, Node::EoL_index, Mesh::EoL_Count (mesh.cpp set_indices, :393-410)
    int count = 0;
    if (eol_index)
        for (int i = 0; i < N; ++i)
            if (eol_index[i] >= 0) {
                R.mesh.nodes[i]->EoL = true;
                R.mesh.nodes[i]->EoL_index = eol_index[i];
                if (eol_index[i] + 1 > count) count = eol_index[i] + 1;
            }
    R.mesh.EoL_Count = count;
}

}  // namespace

extern "C" {
void ref_forces_run(void *p);

// The mesh, material, gravity and step of one run; no fill yet (so that bench.py can time Forces::fill alone, and several
// instances side by side).  mat6 = density, e, nu, beta, dampingA, dampingB.
void *ref_forces_new(int N, int F, const int32_t *face_nodes, const double *x, const double *X, const int32_t *eol_index,
                     const double *mat6, const double *grav3, double h) {
    Run *R = new Run;
    R->material.density = mat6[0]; R->material.e = mat6[1]; R->material.nu = mat6[2]; R->material.beta = mat6[3];
    R->material.dampingA = mat6[4]; R->material.dampingB = mat6[5];
    build_mesh(*R, N, F, face_nodes, x, X, eol_index);
    for (int j = 0; j < 3; ++j) R->grav[j] = grav3[j];
    R->h = h;
    return R;
}
// Forces::fill(mesh, mat, grav, h): the call of Cloth.cpp:365
void ref_forces_run(void *p) {
    Run *R = static_cast<Run *>(p);
    R->forces.fill(R->mesh, R->material, Eigen::Vector3d(R->grav[0], R->grav[1], R->grav[2]), R->h);
}
// Reference Forces::fill on a mesh rebuilt from flat arrays.
void *ref_forces_fill(int N, int F, const int32_t *face_nodes, const double *x, const double *X, const int32_t *eol_index,
                      const double *mat6, const double *grav3, double h) {
    void *R = ref_forces_new(N, F, face_nodes, x, X, eol_index, mat6, grav3, h);
    ref_forces_run(R);
    return R;
}
// The next step on the SAME Mesh and Forces objects: new world positions (and, if X is given, new material coordinates — what an
// EoL node's Eulerian update does, Cloth.cpp:401-407), then Forces::fill again.
void ref_forces_refill(void *p, const double *x, const double *X) {
    Run *R = static_cast<Run *>(p);
    for (size_t i = 0; i < R->mesh.nodes.size(); ++i) {
        R->mesh.nodes[i]->x = Vec3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        if (X) R->mesh.nodes[i]->verts[0]->u = Vec3(X[2 * i], X[2 * i + 1], 0.0);
    }
    R->forces.fill(R->mesh, R->material, Eigen::Vector3d(R->grav[0], R->grav[1], R->grav[2]), R->h);
}
// Mesh only (edge order, normals), no fill
void *ref_forces_mesh(int N, int F, const int32_t *face_nodes, const double *x, const double *X) {
    Run *R = new Run;
    std::memset(&R->material, 0, sizeof(R->material));
    build_mesh(*R, N, F, face_nodes, x, X, nullptr);
    return R;
}
void ref_forces_free(void *p) { delete static_cast<Run *>(p); }

// Seconds per Forces::fill over `steps` further fills on the same objects, every node moved a little before each one (what a time
// step does between two fills); the clock covers the moves' successor only: Forces::fill — for the adapter build that is flatten +
// host -> device + kernels + device -> the Eigen members.
double ref_forces_time_steps(void *p, int steps, double eps) {
    Run *R = static_cast<Run *>(p);
    double total = 0.0;
    for (int s = 0; s < steps; ++s) {
        for (size_t i = 0; i < R->mesh.nodes.size(); ++i) R->mesh.nodes[i]->x[2] += eps * (double)((i + (size_t)s) % 3 == 0 ? 1 : -1);
        const auto t0 = std::chrono::steady_clock::now();
        ref_forces_run(p);
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return steps > 0 ? total / steps : 0.0;
}

int ref_forces_dof(void *p) { return (int)static_cast<Run *>(p)->forces.f.size(); }
int ref_forces_eol_cutoff(void *p) { return static_cast<Run *>(p)->forces.EoL_cutoff; }
const double *ref_forces_f(void *p) { return static_cast<Run *>(p)->forces.f.data(); }
int64_t ref_forces_nnz(void *p, int which) { Run *R = static_cast<Run *>(p); return (int64_t)(which ? R->forces.MDK : R->forces.M).nonZeros(); }
const int *ref_forces_outer(void *p, int which) { Run *R = static_cast<Run *>(p); return (which ? R->forces.MDK : R->forces.M).outerIndexPtr(); }
const int *ref_forces_inner(void *p, int which) { Run *R = static_cast<Run *>(p); return (which ? R->forces.MDK : R->forces.M).innerIndexPtr(); }
const double *ref_forces_vals(void *p, int which) { Run *R = static_cast<Run *>(p); return (which ? R->forces.MDK : R->forces.M).valuePtr(); }

// mesh.edges as Mesh::add(Face*) created them; per edge the stencil edgeBasedF reads (Forces.cpp:688-697):
// n[0], n[1], the vertex of adjf[0] / adjf[1] opposite the edge (get_other_vert, mesh.hpp:276-280), -1 where there is no face
int ref_forces_n_edges(void *p) { return (int)static_cast<Run *>(p)->mesh.edges.size(); }
void ref_forces_edge_stencils(void *p, int32_t *out4E) {
    Run *R = static_cast<Run *>(p);
    for (size_t e = 0; e < R->mesh.edges.size(); ++e) {
        Edge *edge = R->mesh.edges[e];
        Vert *v0 = edge->n[0]->verts[0], *v1 = edge->n[1]->verts[0];
        out4E[4 * e] = edge->n[0]->index; out4E[4 * e + 1] = edge->n[1]->index;
        for (int s = 0; s < 2; ++s)
            out4E[4 * e + 2 + s] = edge->adjf[s] ? get_other_vert(edge->adjf[s], v0, v1)->node->index : -1;
    }
}
// face->n and node->n after compute_ws_data (mesh.cpp:135-143,150-151; what Cloth::updatePosNor and Constraints::fill read);
// x: new positions (3N) applied to Node::x first, as Cloth::step does before compute_ws_data (Cloth.cpp:394-410)
void ref_forces_normals(void *p, const double *x, double *face_n, double *node_n) {
    Run *R = static_cast<Run *>(p);
    if (x)
        for (size_t i = 0; i < R->mesh.nodes.size(); ++i) R->mesh.nodes[i]->x = Vec3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    compute_ws_data(R->mesh);
    for (size_t k = 0; k < R->mesh.faces.size(); ++k) for (int j = 0; j < 3; ++j) face_n[3 * k + j] = R->mesh.faces[k]->n[j];
    for (size_t i = 0; i < R->mesh.nodes.size(); ++i) for (int j = 0; j < 3; ++j) node_n[3 * i + j] = R->mesh.nodes[i]->n[j];
}

}  // extern "C"
