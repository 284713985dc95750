// host_driver.cpp — C++ test driver for include/eolc_host.hpp (TEST INFRASTRUCTURE).
//
// Builds an ArcSim-shaped pointer mesh (same member names as /root/reference/src/external/ArcSim/mesh.hpp:57-190, edges created
// the way Mesh::add(Face) does, mesh.cpp:356-378), then goes through exactly what the reference-side adapters do:
//   eolc::host::flatten(mesh) -> eolc::host::Forces::fill(...) -> eolc::host::CD2(...)
// and dumps the results for tests/test_host_cpp.py to compare with the oracle.
//   usage: host_driver <in.bin> <out.bin> [n]     with n: the mesh is an n x n grid sheet; flag the interior nodes of its grid line
//                                                 j = n / 2 as EoL nodes (Node::EoL, EoL_index in grid order, mesh.EoL_Count)
//          host_driver bench <n> <steps>          adapter-level step time on an n x n regular2 sheet (bench.py's e2e.host_layer):
//                                                 flatten of the pointer mesh + Forces::fill through the host layer, one JSON line
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <unordered_map>
#include <vector>
#include "eolc_host.hpp"

struct Node; struct Face;
struct Vert { double u[3]; Node *node; };
struct Node { double x[3]; std::vector<Vert *> verts; int index; bool EoL; int EoL_index; };
struct Edge { Node *n[2]; Face *adjf[2]; };
struct Face { Vert *v[3]; };
struct Mesh {
    std::vector<Vert *> verts; std::vector<Node *> nodes; std::vector<Edge *> edges; std::vector<Face *> faces;
    int EoL_Count = 0;
};

template <class T> static void rd(FILE *f, T *p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { std::printf("short read\n"); std::abort(); } }
template <class T> static void wr(FILE *f, const T *p, size_t n) { fwrite(p, sizeof(T), n, f); }

static void build_pointer_mesh(Mesh &mesh, int32_t N, int32_t F, const double *x, const double *X, const int32_t *fn) {
    for (int i = 0; i < N; ++i) {
        Node *n = new Node; Vert *v = new Vert;
        for (int k = 0; k < 3; ++k) n->x[k] = x[3 * (size_t)i + k];
        v->u[0] = X[2 * (size_t)i]; v->u[1] = X[2 * (size_t)i + 1]; v->u[2] = 0; v->node = n;
        n->verts.push_back(v); n->index = i; n->EoL = false; n->EoL_index = -1;
        mesh.nodes.push_back(n); mesh.verts.push_back(v);
    }
    std::unordered_map<uint64_t, Edge *> edge_of;
    edge_of.reserve((size_t)2 * F);
    auto key_of = [](Node *a, Node *b) { uint32_t i = (uint32_t)a->index, j = (uint32_t)b->index; return ((uint64_t)(i < j ? i : j) << 32) | (i < j ? j : i); };
    for (int k = 0; k < F; ++k) {
        Face *f = new Face;
        for (int j = 0; j < 3; ++j) f->v[j] = mesh.verts[fn[3 * (size_t)k + j]];
        mesh.faces.push_back(f);
        for (int i = 0; i < 3; ++i) {   // add_edges_if_needed (mesh.cpp:356-363)
            Node *a = f->v[i]->node, *b = f->v[(i + 1) % 3]->node;
            const uint64_t key = key_of(a, b);
            if (!edge_of.count(key)) { Edge *e = new Edge; e->n[0] = a; e->n[1] = b; e->adjf[0] = e->adjf[1] = nullptr; edge_of[key] = e; mesh.edges.push_back(e); }
        }
        for (int i = 0; i < 3; ++i) {   // adjf[side], side = 0 iff the face runs n[0] -> n[1] (mesh.cpp:370-377)
            Node *v0 = f->v[(i + 1) % 3]->node, *v1 = f->v[(i + 2) % 3]->node;
            Edge *e = edge_of[key_of(v0, v1)];
            e->adjf[e->n[0] == v0 ? 0 : 1] = f;
        }
    }
}

// Adapter-level step time (what adapter/Forces_fill_b200.cpp does each step, minus the copy into Eigen, which needs Eigen):
// flatten of the ArcSim pointer mesh + eolc::host::Forces::fill.  Reported: the first call (plan build), a step after a remesh that
// changed X (full fill incl. M), a steady step (positions only: M unchanged), and the flatten alone.
static int bench_main(int n, int steps) {
    const int32_t N = n * n, F = 2 * (n - 1) * (n - 1);
    std::vector<double> x(3 * (size_t)N), X(2 * (size_t)N);
    std::vector<int32_t> fn(3 * (size_t)F);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const size_t k = (size_t)i * n + j;
            X[2 * k] = i / (n - 1.0); X[2 * k + 1] = j / (n - 1.0);
            x[3 * k] = X[2 * k]; x[3 * k + 1] = X[2 * k + 1];
            x[3 * k + 2] = 0.05 * std::sin(6.283185307179586 * X[2 * k]) * std::cos(6.283185307179586 * X[2 * k + 1]) + 1e-3 * std::sin(12.9898 * k);
        }
    size_t q = 0;
    for (int i = 0; i + 1 < n; ++i)
        for (int j = 0; j + 1 < n; ++j) {
            const int32_t k0 = i * n + j;
            fn[q++] = k0; fn[q++] = k0 + n; fn[q++] = k0 + n + 1;
            fn[q++] = k0; fn[q++] = k0 + n + 1; fn[q++] = k0 + 1;
        }
    Mesh mesh;
    build_pointer_mesh(mesh, N, F, x.data(), X.data(), fn.data());
    eolc_material mat = {0.05, 50.0, 0.01, 1.0e-5, 0.0, 1.0};
    const double grav[3] = {0.0, 0.0, -9.8};
    const double h = 0.5e-2;
    typedef std::chrono::steady_clock clk;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    eolc::host::FlatMesh flat;
    eolc::host::Forces forces;
    auto t0 = clk::now();
    eolc::host::flatten(mesh, flat);
    forces.fill(flat, mat, grav, h);
    const double first_ms = ms(t0, clk::now());
    double full_ms = 0, steady_ms = 0, flat_ms = 0, flat_pos_ms = 0;
    for (int s = 0; s < steps; ++s) {
        // a step after something moved in material space: M is recomputed and copied out
        mesh.nodes[(size_t)s % N]->verts[0]->u[0] += 1e-13;
        auto a = clk::now();
        eolc::host::flatten(mesh, flat);
        auto b = clk::now();
        forces.fill(flat, mat, grav, h);
        auto c = clk::now();
        if (!forces.M_updated) { std::printf("M should have been updated\n"); return 3; }
        flat_ms += ms(a, b); full_ms += ms(a, c);
        // a steady step: positions only
        for (size_t i = 0; i < (size_t)N; i += 97) mesh.nodes[i]->x[2] += 1e-9;
        a = clk::now();
        eolc::host::flatten(mesh, flat, true);
        b = clk::now();
        forces.fill(flat, mat, grav, h);
        c = clk::now();
        if (forces.M_updated) { std::printf("M should not have been updated\n"); return 3; }
        flat_pos_ms += ms(a, b); steady_ms += ms(a, c);
    }
    double chk = 0;
    for (size_t i = 0; i < forces.f.size(); i += 1001) chk += forces.f[i];
    std::printf("{\"n\": %d, \"steps\": %d, \"first_call_ms\": %.3f, \"step_full_ms\": %.3f, \"step_M_unchanged_ms\": %.3f, \"flatten_ms\": %.3f, "
                "\"flatten_positions_only_ms\": %.3f, \"elements\": %lld, \"checksum\": %.17g}\n",
                n, steps, first_ms, full_ms / steps, steady_ms / steps, flat_ms / steps, flat_pos_ms / steps,
                (long long)F + (long long)(3LL * (n - 1) * (n - 1) + 2 * (n - 1) - 4 * (n - 1)), chk);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && std::string(argv[1]) == "bench") return bench_main(std::atoi(argv[2]), std::atoi(argv[3]));
    if (argc < 3) { std::printf("usage: host_driver in.bin out.bin\n"); return 2; }
    FILE *in = fopen(argv[1], "rb");
    if (!in) { std::printf("cannot open %s\n", argv[1]); return 2; }
    int32_t N, F;
    rd(in, &N, 1); rd(in, &F, 1);
    std::vector<double> x(3 * (size_t)N), X(2 * (size_t)N), par(6 + 3 + 1 + 1 + 3 + 16);
    std::vector<int32_t> fn(3 * (size_t)F);
    rd(in, x.data(), x.size()); rd(in, X.data(), X.size()); rd(in, fn.data(), fn.size()); rd(in, par.data(), par.size());
    fclose(in);
    // ---- pointer mesh, ArcSim style
    Mesh mesh;
    build_pointer_mesh(mesh, N, F, x.data(), X.data(), fn.data());
    if (argc > 3) {
        const int n = std::atoi(argv[3]);
        for (int i = 1; i + 1 < n; ++i) { Node *nd = mesh.nodes[(size_t)i * n + n / 2]; nd->EoL = true; nd->EoL_index = mesh.EoL_Count++; }
    }
    // ---- the adapter path
    eolc::host::FlatMesh flat;
    eolc::host::flatten(mesh, flat);
    eolc_material mat = {par[0], par[1], par[2], par[3], par[4], par[5]};
    const double grav[3] = {par[6], par[7], par[8]};
    const double h = par[9];
    eolc::host::Forces forces;
    forces.fill(flat, mat, grav, h);
    // move the cloth (positions only) and fill again: the plan must be reused, the result must follow the new positions
    for (auto *n : mesh.nodes) n->x[2] += 0.0;
    eolc::host::flatten(mesh, flat, true);
    forces.fill(flat, mat, grav, h);
    // M depends on X and the density only: the second fill must have left it alone on a Lagrangian mesh (and M.values still holds
    // the first fill's matrix, which the test compares with the oracle); with EoL nodes M follows x and is recomputed
    const int32_t m_updated_second = forces.M_updated ? 1 : 0;
    // a changed material coordinate is noticed by flatten() and M comes back
    mesh.nodes[0]->verts[0]->u[0] += 0.0;
    eolc::host::flatten(mesh, flat);
    const uint64_t v0 = flat.X_version, t0v = flat.topology_version;
    mesh.nodes[0]->verts[0]->u[1] += 1e-300;      // denormal nudge: a different double, same physics
    mesh.nodes[0]->verts[0]->u[1] -= 1e-300;
    eolc::host::flatten(mesh, flat);
    if (flat.X_version != v0 || flat.topology_version != t0v) { std::printf("versions moved without a change\n"); return 3; }
    eolc::host::ObstaclesFlat obs;
    obs.cdthreshold = par[10];
    obs.num_boxes = 1;
    obs.box_dim.assign(par.begin() + 11, par.begin() + 14);
    obs.box_E1.assign(par.begin() + 14, par.begin() + 30);
    std::vector<std::shared_ptr<eolc::host::Collision> > cls;
    eolc::host::CD2(flat, obs, cls);
    // ---- dump
    FILE *out = fopen(argv[2], "wb");
    int32_t dof = forces.M.rows, E = flat.E, ncls = (int32_t)cls.size(), cutoff = forces.EoL_cutoff;
    int64_t nnzM = forces.M.nnz, nnzK = forces.MDK.nnz;
    wr(out, &dof, 1); wr(out, &E, 1); wr(out, &cutoff, 1); wr(out, &ncls, 1); wr(out, &nnzM, 1); wr(out, &nnzK, 1);
    wr(out, &m_updated_second, 1); wr(out, &m_updated_second, 1);   // (twice: keeps the arrays that follow 8-byte aligned)
    wr(out, flat.edge_stencil.data(), flat.edge_stencil.size());
    wr(out, forces.f.data(), forces.f.size());
    wr(out, forces.M.outer, (size_t)dof + 1); wr(out, forces.M.inner, (size_t)nnzM); wr(out, forces.M.values.data(), (size_t)nnzM);
    wr(out, forces.MDK.outer, (size_t)dof + 1); wr(out, forces.MDK.inner, (size_t)nnzK); wr(out, forces.MDK.values.data(), (size_t)nnzK);
    for (auto &c : cls) wr(out, c.get(), 1);
    fclose(out);
    std::printf("host_driver OK: dof=%d nnz(M)=%lld nnz(MDK)=%lld contacts=%d\n", dof, (long long)nnzM, (long long)nnzK, ncls);
    return 0;
}
