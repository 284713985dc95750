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

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
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

// Page-locked host array (eolc_host_alloc): what the host entry points DMA from / into directly.  A pageable std::vector
// costs an extra pass through the library's staging buffer (1.5 GB per fill at 1024^2).  The small vector-like surface the
// host layer and the adapters need; contents are NOT preserved by a growing resize().
template <class T>
class PinnedArray {
public:
    PinnedArray() {}
    ~PinnedArray() { eolc_host_free(p_); }
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    void resize(size_t n) {
        if (n > cap_) {
            eolc_host_free(p_);
            p_ = nullptr; cap_ = 0;
            p_ = static_cast<T *>(eolc_host_alloc(n * sizeof(T)));
            if (!p_) fail("eolc_host_alloc", EOLC_ERR_CUDA);
            cap_ = n;
        }
        n_ = n;
    }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
private:
    T *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

// Process-wide source of version numbers: two FlatMesh objects never carry the same one, so a cache keyed on a version cannot
// mistake one mesh for another.
inline uint64_t next_version() { static std::atomic<uint64_t> v(0); return ++v; }

// The part of the ArcSim mesh the hot path reads, flattened.
struct FlatMesh {
    int32_t N = 0, F = 0, E = 0;
    PinnedArray<double> x;             // 3N  Node::x
    PinnedArray<double> X;             // 2N  node->verts[0]->u[0..1]
    std::vector<int32_t> face_nodes;   // 3F  faces[k]->v[0..2]->node->index
    std::vector<int32_t> edge_stencil; // 4E  (n[0], n[1], opp(adjf[0]), opp(adjf[1])), -1 = no face
    std::vector<int32_t> eol_index;    // N   Node::EoL_index, -1 = Lagrangian
    int32_t EoL_Count = 0;
    // Change counters maintained by flatten(): the consumers key their caches on them instead of re-hashing the arrays.
    uint64_t topology_version = 0;     // renewed when N, face_nodes, edge_stencil or eol_index changed (remesh / set_indices)
    uint64_t X_version = 0;            // renewed when any material coordinate changed (remesh, or EoL nodes moving in X)
    // for callers that fill the arrays themselves instead of through flatten(): declare what changed
    void touch_topology() { topology_version = next_version(); X_version = next_version(); }
    void touch_X() { X_version = next_version(); }
};

// Works on any mesh type with ArcSim's field names (mesh.hpp:57-190): nodes[i]->{x, verts, index, EoL, EoL_index},
// verts->{u, node}, faces[k]->v[3], edges[e]->{n[2], adjf[2]}.  `positions_only` refreshes x (and X) of an existing
// FlatMesh without touching the topology arrays — what a step without remeshing needs.
// Changes are detected while the values are copied (a compare per value written, no second pass, no hash): the version counters
// of `out` move only when something really changed, whatever the caller believes.  The loops run on up to 16 host threads
// (EOLC_HOST_THREADS overrides): the walk over a million-node pointer mesh is latency-bound pointer chasing (100 ms on one core),
// every node / face / edge is independent, and the result does not depend on the thread count.
namespace detail {
inline int host_threads() {
    const char *ev = std::getenv("EOLC_HOST_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    const int n = ev ? std::atoi(ev) : (int)(hw ? (hw < 16u ? hw : 16u) : 1u);
    return n < 1 ? 1 : n;
}
// body(lo, hi) -> bool "something changed", over contiguous slices of [0, n); returns the OR
template <class Body> inline bool par_any(size_t n, Body body) {
    const int nw = (int)std::min<size_t>((size_t)host_threads(), n / 16384 + 1);
    if (nw <= 1) return body((size_t)0, n);
    std::vector<char> changed((size_t)nw, 0);
    std::vector<std::thread> th;
    for (int w = 0; w < nw; ++w)
        th.emplace_back([&, w]() { changed[(size_t)w] = body(n * (size_t)w / (size_t)nw, n * (size_t)(w + 1) / (size_t)nw) ? 1 : 0; });
    for (auto &t : th) t.join();
    bool any = false;
    for (char c : changed) any = any || c;
    return any;
}
}  // namespace detail

// flatten() for an adapter that sees the mesh once per step: the full walk (23 ms for a 1 M-node mesh on 16 threads) is what notices a
// remesh; a run WITHOUT remeshing (`"remeshing": false` in the reference's settings) may promise that with EOLC_STATIC_TOPOLOGY=1 in
// the environment, and then only positions and material coordinates are refreshed (2 ms) as long as the element counts agree.
template <class MeshT> void flatten(const MeshT &mesh, FlatMesh &out, bool positions_only);
template <class MeshT> inline void flatten_step(const MeshT &mesh, FlatMesh &out) {
    static const bool static_topology = [] { const char *e = std::getenv("EOLC_STATIC_TOPOLOGY"); return e && std::atoi(e) != 0; }();
    const bool same_counts = out.topology_version != 0 && (size_t)out.N == mesh.nodes.size() && (size_t)out.F == mesh.faces.size() &&
                             (size_t)out.E == mesh.edges.size();
    flatten(mesh, out, static_topology && same_counts);
}

template <class MeshT> inline void flatten(const MeshT &mesh, FlatMesh &out) { flatten(mesh, out, false); }

// memcpy of a large array on the host threads (a single thread moves ~10 GB/s: 0.1 s for the 1 GB of MDK values at 1024^2)
template <class T> inline void par_copy(T *dst, const T *src, size_t n) {
    detail::par_any(n, [&](size_t lo, size_t hi) { std::memcpy(dst + lo, src + lo, (hi - lo) * sizeof(T)); return false; });
}

template <class MeshT>
void flatten(const MeshT &mesh, FlatMesh &out, bool positions_only) {
    const size_t N = mesh.nodes.size();
    bool topo_changed = out.N != (int32_t)N || out.x.size() != 3 * N, X_changed = topo_changed;
    out.N = (int32_t)N;
    out.x.resize(3 * N);
    out.X.resize(2 * N);
    X_changed |= detail::par_any(N, [&](size_t lo, size_t hi) {
        bool ch = false;
        for (size_t i = lo; i < hi; ++i) {
            const auto *n = mesh.nodes[i];
            out.x[3 * i] = n->x[0]; out.x[3 * i + 1] = n->x[1]; out.x[3 * i + 2] = n->x[2];      // Forces.cpp:343-348
            const double u0 = n->verts[0]->u[0], u1 = n->verts[0]->u[1];                         // Forces.cpp:349-355
            if (out.X[2 * i] != u0 || out.X[2 * i + 1] != u1) ch = true;
            out.X[2 * i] = u0; out.X[2 * i + 1] = u1;
        }
        return ch;
    });
    if (X_changed) out.X_version = next_version();
    if (positions_only) { if (topo_changed) out.topology_version = next_version(); return; }
    if (out.eol_index.size() != N) { out.eol_index.assign(N, -1); topo_changed = true; }
    int32_t eol_count = 0;
    for (size_t i = 0; i < N; ++i) {
        const int32_t v = mesh.nodes[i]->EoL ? (int32_t)mesh.nodes[i]->EoL_index : -1;
        if (mesh.nodes[i]->EoL) ++eol_count;
        if (out.eol_index[i] != v) { out.eol_index[i] = v; topo_changed = true; }
    }
    out.EoL_Count = eol_count;
    const size_t F = mesh.faces.size();
    if (out.face_nodes.size() != 3 * F) { out.face_nodes.assign(3 * F, -1); topo_changed = true; }
    out.F = (int32_t)F;
    topo_changed |= detail::par_any(F, [&](size_t lo, size_t hi) {
        bool ch = false;
        for (size_t k = lo; k < hi; ++k)
            for (int j = 0; j < 3; ++j) {                                                          // Forces.cpp:376-378
                const int32_t v = mesh.faces[k]->v[j]->node->index;
                if (out.face_nodes[3 * k + j] != v) { out.face_nodes[3 * k + j] = v; ch = true; }
            }
        return ch;
    });
    const size_t E = mesh.edges.size();
    if (out.edge_stencil.size() != 4 * E) { out.edge_stencil.assign(4 * E, -2); topo_changed = true; }
    out.E = (int32_t)E;
    topo_changed |= detail::par_any(E, [&](size_t lo, size_t hi) {
        bool ch = false;
        for (size_t e = lo; e < hi; ++e) {
            const auto *ed = mesh.edges[e];
            int32_t s[4] = {(int32_t)ed->n[0]->index, (int32_t)ed->n[1]->index, -1, -1};
            for (int side = 0; side < 2; ++side) {
                const auto *f = ed->adjf[side];
                if (!f) continue;                                                                  // Forces.cpp:688-690
                for (int j = 0; j < 3; ++j) {                                                      // get_other_vert, mesh.hpp:276-280
                    const auto *nd = f->v[j]->node;
                    if (nd != ed->n[0] && nd != ed->n[1]) { s[2 + side] = nd->index; break; }
                }
            }
            int32_t *d = &out.edge_stencil[4 * e];
            for (int q = 0; q < 4; ++q) if (d[q] != s[q]) { d[q] = s[q]; ch = true; }
        }
        return ch;
    });
    if (topo_changed) out.topology_version = next_version();
}

// Column-major compressed sparse matrix laid out exactly like Eigen::SparseMatrix<double> (outer/inner/values).
// outer and inner point into the plan (valid until the next topology change); values are owned.
struct SparseCSC {
    int32_t rows = 0, cols = 0;
    int64_t nnz = 0;
    const int32_t *outer = nullptr;   // cols + 1
    const int32_t *inner = nullptr;   // nnz, ascending within a column
    PinnedArray<double> values;       // nnz, page-locked: eolc_forces_fill copies device -> here without staging
};

// class Forces (Forces.h:26-45).  The topology plan (CSR pattern + element->slot maps) is cached and rebuilt only when
// the flattened topology changes, i.e. after dynamic_remesh / set_indices (Scene.cpp:87-90) or the preprocessor.
class Forces {
public:
    // The Context outlives the plan: it is looked up BEFORE the plan can exist, so a thread_local Forces constructed by an adapter is
    // destroyed before the thread's Context (reverse order of construction).
    explicit Forces(Context *ctx = nullptr) : EoL_cutoff(0), ctx_(ctx ? ctx : &Context::instance()) {}
    ~Forces() { eolc_forces_plan_destroy(plan_); }
    Forces(const Forces &) = delete;
    Forces &operator=(const Forces &) = delete;

    PinnedArray<double> f;   // Eigen::VectorXd f
    SparseCSC M;             // Eigen::SparseMatrix<double> M
    SparseCSC MDK;           // Eigen::SparseMatrix<double> MDK
    int EoL_cutoff;

    // void Forces::fill(const Mesh&, const Material&, const Vector3d& grav, double h)   Forces.cpp:912-930
    // M_updated tells the caller whether M.values was rewritten by the last fill (false: M is the matrix of the previous step;
    // it depends on X and the density only, ComputeInertial.cpp:33,44-47, so a step without remeshing leaves it alone).
    bool M_updated = true;
    // Caller-owned result arrays (dof, nnz(M), nnz(MDK) doubles) instead of this object's own f / M.values / MDK.values: what an
    // adapter passes to have the results written straight into the Eigen members of the reference's class Forces (their value
    // arrays page-locked in place with eolc_host_register) — no second 1 GB copy on the host.
    struct External { double *f, *M_vals, *MDK_vals; };
    // The plan for this topology (rebuilt only when mesh.topology_version moved) and the patterns M / MDK (rows, nnz, outer, inner)
    // that go with it; true if it was rebuilt, i.e. if arrays sized from an earlier pattern are stale.
    bool prepare(const FlatMesh &mesh) {
        Context &c = *ctx_;
        if (plan_ && mesh.topology_version == topo_version_) return false;
        eolc_forces_plan_destroy(plan_);
        plan_ = nullptr;
        check(eolc_forces_plan_create(c.handle(), mesh.N, mesh.F, mesh.face_nodes.data(), mesh.E, mesh.edge_stencil.data(),
                                      mesh.eol_index.empty() ? nullptr : mesh.eol_index.data(), mesh.X.data(), &plan_),
              "eolc_forces_plan_create");
        topo_version_ = mesh.topology_version;
        bind(0, M);
        bind(1, MDK);
        have_M_ = false;
        return true;
    }
    void fill(const FlatMesh &mesh, const eolc_material &mat, const double grav[3], double h, const External *ext = nullptr) {
        prepare(mesh);
        double *fo, *Mo, *Ko;
        if (ext) { fo = ext->f; Mo = ext->M_vals; Ko = ext->MDK_vals; }
        else {
            f.resize((size_t)M.rows);                   // f.resize(3N + 2 EoL_Count), Forces.cpp:914
            M.values.resize((size_t)M.nnz); MDK.values.resize((size_t)MDK.nnz);
            fo = f.data(); Mo = M.values.data(); Ko = MDK.values.data();
        }
        if (Mo != last_M_out_) have_M_ = false;         // another destination: it does not hold the previous step's M
        EoL_cutoff = 3 * mesh.N;                        // Forces.cpp:919
        const bool same_M = have_M_ && mesh.EoL_Count == 0 && mesh.X_version == X_version_ && mat.density == density_;
        check(eolc_forces_fill_ex(plan_, mesh.x.data(), mesh.X.data(), &mat, grav, h, fo, Mo, Ko, same_M ? EOLC_FILL_M_UNCHANGED : 0u),
              "eolc_forces_fill");
        M_updated = !same_M;
        have_M_ = true; X_version_ = mesh.X_version; density_ = mat.density; last_M_out_ = Mo;
    }
    const eolc_forces_plan *plan() const { return plan_; }

private:
    void bind(int which, SparseCSC &A) {
        int32_t dof = 0;
        check(eolc_forces_pattern(plan_, which, &dof, &A.nnz, &A.outer, &A.inner), "eolc_forces_pattern");
        A.rows = A.cols = dof;       // A.values is sized by fill() when the results go to this object's own arrays
    }
    Context *ctx_;
    eolc_forces_plan *plan_ = nullptr;
    uint64_t topo_version_ = 0, X_version_ = 0;
    double density_ = 0.0;
    bool have_M_ = false;
    const double *last_M_out_ = nullptr;
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
    uint64_t topo_version = 0;
    double threshold = 0.0;
    PinnedArray<eolc_contact> buf;       // page-locked: the contact list is DMA'd straight into it
    ~CdCache() { eolc_cd_plan_destroy(plan); }
};
// The contacts of one call as a view into the thread's page-locked record buffer (valid until the next call of the thread)
struct ContactView { const eolc_contact *data; int32_t size; };
inline ContactView run_cd_view(const FlatMesh &mesh, const ObstaclesFlat &obs, int cd1, Context *ctx) {
    Context &c = ctx ? *ctx : Context::instance();   // before the cache: the thread's Context must outlive the cached plan
    static thread_local CdCache cache;
    // the plan (btc edge table + perturbation stream) follows the mesh's topology counter and the threshold
    if (!cache.plan || cache.topo_version != mesh.topology_version || cache.threshold != obs.cdthreshold) {
        eolc_cd_plan_destroy(cache.plan);
        cache.plan = nullptr;
        check(eolc_cd_plan_create(c.handle(), mesh.N, mesh.F, mesh.face_nodes.data(), obs.cdthreshold, &cache.plan), "eolc_cd_plan_create");
        cache.topo_version = mesh.topology_version; cache.threshold = obs.cdthreshold;
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
    ContactView v = {cache.buf.data(), n};
    return v;
}
inline void run_cd(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, int cd1, Context *ctx) {
    const ContactView v = run_cd_view(mesh, obs, cd1, ctx);
    cls.reserve(cls.size() + (size_t)v.size);
    for (int32_t i = 0; i < v.size; ++i) cls.push_back(std::make_shared<Collision>(v.data[i]));   // appended, caller clears (Scene.cpp:93)
}
}  // namespace detail

// void CD(const Mesh&, const shared_ptr<Obstacles>, vector<shared_ptr<btc::Collision>>& cls)    Collisions.cpp:11-53
inline void CD(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, Context *ctx = nullptr) {
    detail::run_cd(mesh, obs, cls, 1, ctx);
}
// The same calls for a caller that builds its own record objects (the reference-side adapter makes btc::Collision objects): the
// contacts as a view into the calling thread's page-locked buffer, valid until the thread's next CD / CD2 call.
typedef detail::ContactView ContactView;
inline ContactView CD_view(const FlatMesh &mesh, const ObstaclesFlat &obs, Context *ctx = nullptr) { return detail::run_cd_view(mesh, obs, 1, ctx); }
inline ContactView CD2_view(const FlatMesh &mesh, const ObstaclesFlat &obs, Context *ctx = nullptr) { return detail::run_cd_view(mesh, obs, 0, ctx); }
// void CD2(...)                                                                                  Collisions.cpp:55-78
inline void CD2(const FlatMesh &mesh, const ObstaclesFlat &obs, std::vector<std::shared_ptr<Collision> > &cls, Context *ctx = nullptr) {
    detail::run_cd(mesh, obs, cls, 0, ctx);
}

}  // namespace host
}  // namespace eolc
#endif  // EOLC_HOST_HPP_
