// eolc_host.hpp — C++ host side of the B200 hot path, above the C ABI (include/eolc.h).
//
// Header-only and Eigen-free, so it compiles wherever a C++11 compiler exists.  It mirrors the reference's own
// interface for this path — same names, same argument meaning, same error behaviour (message on stdout + abort(),
// like /root/reference/src/parseParams.cpp:31-39 and GeneralizedSolver.cpp:139-143):
//
//   eolc::host::Forces       members f, M, MDK, EoL_cutoff; method fill(mesh, mat, grav, h)
//                            == class Forces, /root/reference/src/Forces.h:26-45, body Forces.cpp:912-930
//   eolc::host::CD / CD2     append btc::Collision-shaped records to the caller's vector of shared_ptr
//                            == /root/reference/src/Collisions.h:9,11, bodies Collisions.cpp:11-78
//   eolc::host::flatten()    ArcSim pointer mesh -> flat arrays; duck-typed template, so the
//                            same code serves the reference's `Mesh` (external/ArcSim/mesh.hpp:166-190) and test meshes
//
// The adapter sources in adapter/ put the reference's exact signatures (Eigen types) on top of these classes.
// There is NO CPU fallback: without a CUDA device the first call prints the C-ABI error and aborts.
#ifndef EOLC_HOST_HPP_
#define EOLC_HOST_HPP_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "eolc.h"

namespace eolc {
namespace host {

// Reference convention: print and abort().  Tests define EOLC_HOST_THROW to get an exception instead.
inline void fail(const char *where, int rc) {
    std::string msg = std::string(where) + ": " + eolc_last_error() + " (status " + std::to_string(rc) + ")";
#ifdef EOLC_HOST_THROW
    throw std::runtime_error(msg);
#else
    std::printf("%s\n", msg.c_str());
    std::fflush(stdout);
    std::abort();
#endif
}
inline void check(int rc, const char *where) { if (rc != EOLC_OK) fail(where, rc); }

// One ctx per host thread and device (the reference has a single simulation thread, runner.cpp:58-63,121-129).
class Context {
public:
    explicit Context(int device = 0) { check(eolc_ctx_create(device, &h_), "eolc_ctx_create"); }
    ~Context() { eolc_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    eolc_ctx *handle() const { return h_; }
    // process-wide default (device from EOLC_DEVICE, default 0), created on first use by the calling thread
    static Context &instance() {
        static thread_local std::unique_ptr<Context> c;
        if (!c) { const char *d = std::getenv("EOLC_DEVICE"); c.reset(new Context(d ? std::atoi(d) : 0)); }
        return *c;
    }
private:
    eolc_ctx *h_ = nullptr;
};

// The part of the ArcSim mesh the hot path reads, flattened.
struct FlatMesh {
    int32_t N = 0, F = 0, E = 0;
    std::vector<double> x;             // 3N  Node::x
    std::vector<double> X;             // 2N  node->verts[0]->u[0..1]
    std::vector<int32_t> face_nodes;   // 3F  faces[k]->v[0..2]->node->index
    std::vector<int32_t> edge_stencil; // 4E  (n[0], n[1], opp(adjf[0]), opp(adjf[1])), -1 = no face
    std::vector<int32_t> eol_index;    // N   Node::EoL_index, -1 = Lagrangian
    int32_t EoL_Count = 0;
};

// Works on any mesh type with ArcSim's field names (mesh.hpp:57-190): nodes[i]->{x, verts, index, EoL, EoL_index},
// verts->{u, node}, faces[k]->v[3], edges[e]->{n[2], adjf[2]}.  `positions_only` refreshes x (and X) of an existing
// FlatMesh without touching the topology arrays — what a step without remeshing needs.
template <class MeshT>
void flatten(const MeshT &mesh, FlatMesh &out, bool positions_only = false) {
    const size_t N = mesh.nodes.size();
    out.N = (int32_t)N;
    out.x.resize(3 * N);
    out.X.resize(2 * N);
    for (size_t i = 0; i < N; ++i) {
        const auto *n = mesh.nodes[i];
        out.x[3 * i] = n->x[0]; out.x[3 * i + 1] = n->x[1]; out.x[3 * i + 2] = n->x[2];      // Forces.cpp:343-348
        out.X[2 * i] = n->verts[0]->u[0]; out.X[2 * i + 1] = n->verts[0]->u[1];              // Forces.cpp:349-355
    }
    if (positions_only) return;
    out.eol_index.assign(N, -1);
    out.EoL_Count = 0;
    for (size_t i = 0; i < N; ++i)
        if (mesh.nodes[i]->EoL) { out.eol_index[i] = mesh.nodes[i]->EoL_index; ++out.EoL_Count; }
    const size_t F = mesh.faces.size();
    out.F = (int32_t)F;
    out.face_nodes.resize(3 * F);
    for (size_t k = 0; k < F; ++k)
        for (int j = 0; j < 3; ++j) out.face_nodes[3 * k + j] = mesh.faces[k]->v[j]->node->index;   // Forces.cpp:376-378
    const size_t E = mesh.edges.size();
    out.E = (int32_t)E;
    out.edge_stencil.resize(4 * E);
    for (size_t e = 0; e < E; ++e) {
        const auto *ed = mesh.edges[e];
        int32_t *s = &out.edge_stencil[4 * e];
        s[0] = ed->n[0]->index; s[1] = ed->n[1]->index; s[2] = -1; s[3] = -1;
        for (int side = 0; side < 2; ++side) {
            const auto *f = ed->adjf[side];
            if (!f) continue;                                                                      // Forces.cpp:688-690
            for (int j = 0; j < 3; ++j) {                                                          // get_other_vert, mesh.hpp:276-280
                const auto *nd = f->v[j]->node;
                if (nd != ed->n[0] && nd != ed->n[1]) { s[2 + side] = nd->index; break; }
            }
        }
    }
}

// Column-major compressed sparse matrix laid out exactly like Eigen::SparseMatrix<double> (outer/inner/values).
// outer and inner point into the plan (valid until the next topology change); values are owned.
struct SparseCSC {
    int32_t rows = 0, cols = 0;
    int64_t nnz = 0;
    const int32_t *outer = nullptr;   // cols + 1
    const int32_t *inner = nullptr;   // nnz, ascending within a column
    std::vector<double> values;       // nnz
};

namespace detail {
inline uint64_t fnv1a(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
inline uint64_t topo_key(const FlatMesh &m) {
    uint64_t h = fnv1a(&m.N, sizeof m.N);
    h = fnv1a(m.face_nodes.data(), m.face_nodes.size() * sizeof(int32_t), h);
    h = fnv1a(m.edge_stencil.data(), m.edge_stencil.size() * sizeof(int32_t), h);
    return fnv1a(m.eol_index.data(), m.eol_index.size() * sizeof(int32_t), h);
}
}  // namespace detail

// class Forces (Forces.h:26-45).  The topology plan (CSR pattern + element->slot maps) is cached and rebuilt only when
// the flattened topology changes, i.e. after dynamic_remesh / set_indices (Scene.cpp:87-90) or the preprocessor.
class Forces {
public:
    explicit Forces(Context *ctx = nullptr) : EoL_cutoff(0), ctx_(ctx) {}
    ~Forces() { eolc_forces_plan_destroy(plan_); }
    Forces(const Forces &) = delete;
    Forces &operator=(const Forces &) = delete;

    std::vector<double> f;   // Eigen::VectorXd f
    SparseCSC M;             // Eigen::SparseMatrix<double> M
    SparseCSC MDK;           // Eigen::SparseMatrix<double> MDK
    int EoL_cutoff;

    // void Forces::fill(const Mesh&, const Material&, const Vector3d& grav, double h)   Forces.cpp:912-930
    void fill(const FlatMesh &mesh, const eolc_material &mat, const double grav[3], double h) {
        Context &c = ctx_ ? *ctx_ : Context::instance();
        const uint64_t key = detail::topo_key(mesh);
        if (!plan_ || key != key_) {
            eolc_forces_plan_destroy(plan_);
            plan_ = nullptr;
            check(eolc_forces_plan_create(c.handle(), mesh.N, mesh.F, mesh.face_nodes.data(), mesh.E, mesh.edge_stencil.data(),
                                          mesh.eol_index.empty() ? nullptr : mesh.eol_index.data(), mesh.X.data(), &plan_),
                  "eolc_forces_plan_create");
            key_ = key;
            bind(0, M);
            bind(1, MDK);
        }
        f.resize((size_t)M.rows);                       // f.resize(3N + 2 EoL_Count), Forces.cpp:914
        EoL_cutoff = 3 * mesh.N;                        // Forces.cpp:919
        check(eolc_forces_fill(plan_, mesh.x.data(), mesh.X.data(), &mat, grav, h, f.data(), M.values.data(), MDK.values.data()),
              "eolc_forces_fill");
    }
    const eolc_forces_plan *plan() const { return plan_; }

private:
    void bind(int which, SparseCSC &A) {
        int32_t dof = 0;
        check(eolc_forces_pattern(plan_, which, &dof, &A.nnz, &A.outer, &A.inner), "eolc_forces_pattern");
        A.rows = A.cols = dof;
        A.values.resize((size_t)A.nnz);
    }
    Context *ctx_;
    eolc_forces_plan *plan_ = nullptr;
    uint64_t key_ = 0;
};

// What CD / CD2 read from `Obstacles` (Obstacles.h:31-35, Points.h:22-24, Box.h:43-47).
struct ObstaclesFlat {
    double cdthreshold = 0.0;
    int32_t num_points = 0;
    std::vector<double> pxyz, norms;      // 3 x num_points, column-major (Points::pxyz / norms)
    int32_t num_boxes = 0;
    std::vector<double> box_dim;          // 3 per box (Box::dim)
    std::vector<double> box_E1;           // 16 per box, column-major 4x4 (Box::E1)
};

typedef eolc_contact Collision;           // POD mirror of btc::Collision (boxTriCollision.h:49-110)

namespace detail {
struct CdCache {
    eolc_cd_plan *plan = nullptr;
    uint64_t key = 0;
    std::vector<eolc_contact> buf;
    ~CdCache() { eolc_cd_plan_destroy(plan); }
};
inline void run_cd(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, int cd1, Context *ctx) {
    static thread_local CdCache cache;
    Context &c = ctx ? *ctx : Context::instance();
    uint64_t key = fnv1a(&mesh.N, sizeof mesh.N);
    key = fnv1a(mesh.face_nodes.data(), mesh.face_nodes.size() * sizeof(int32_t), key);
    key = fnv1a(&obs.cdthreshold, sizeof obs.cdthreshold, key);
    if (!cache.plan || cache.key != key) {
        eolc_cd_plan_destroy(cache.plan);
        cache.plan = nullptr;
        check(eolc_cd_plan_create(c.handle(), mesh.N, mesh.F, mesh.face_nodes.data(), obs.cdthreshold, &cache.plan), "eolc_cd_plan_create");
        cache.key = key;
    }
    if (cache.buf.size() < 1024) cache.buf.resize(1024);
    int32_t n = 0;
    for (;;) {
        int rc = eolc_cd_run(cache.plan, mesh.x.data(), obs.num_points, obs.pxyz.data(), obs.norms.data(), obs.num_boxes,
                             obs.box_dim.data(), obs.box_E1.data(), cd1, cd1, cache.buf.data(), (int32_t)cache.buf.size(), &n);
        if (rc == EOLC_ERR_CAPACITY) { cache.buf.resize((size_t)n); continue; }
        check(rc, "eolc_cd_run");
        break;
    }
    cls.reserve(cls.size() + (size_t)n);
    for (int32_t i = 0; i < n; ++i) cls.push_back(std::make_shared<Collision>(cache.buf[(size_t)i]));   // appended, caller clears (Scene.cpp:93)
}
}  // namespace detail

// void CD(const Mesh&, const shared_ptr<Obstacles>, vector<shared_ptr<btc::Collision>>& cls)    Collisions.cpp:11-53
inline void CD(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, Context *ctx = nullptr) {
    detail::run_cd(mesh, obs, cls, 1, ctx);
}
// void CD2(...)                                                                                  Collisions.cpp:55-78
inline void CD2(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, Context *ctx = nullptr) {
    detail::run_cd(mesh, obs, cls, 0, ctx);
}

}  // namespace host
}  // namespace eolc
#endif  // EOLC_HOST_HPP_
