// host_driver.cpp — C++ test driver for include/eolc_host.hpp (TEST INFRASTRUCTURE).
//
// Builds an ArcSim-shaped pointer mesh (same member names as /root/reference/src/external/ArcSim/mesh.hpp:57-190, edges created
// the way Mesh::add(Face) does, mesh.cpp:356-378), then goes through exactly what the reference-side adapters do:
//   eolc::host::flatten(mesh) -> eolc::host::Forces::fill(...) -> eolc::host::CD2(...)
// and dumps the results for tests/test_host_cpp.py to compare with the oracle.
//   usage: host_driver <in.bin> <out.bin> [n]     with n: the mesh is an n x n grid sheet; flag the interior nodes of its grid line
//                                                 j = n / 2 as EoL nodes (Node::EoL, EoL_index in grid order, mesh.EoL_Count)
#include <cstdio>
#include <cstdlib>
#include <map>
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

int main(int argc, char **argv) {
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
    for (int i = 0; i < N; ++i) {
        Node *n = new Node; Vert *v = new Vert;
        for (int k = 0; k < 3; ++k) n->x[k] = x[3 * (size_t)i + k];
        v->u[0] = X[2 * (size_t)i]; v->u[1] = X[2 * (size_t)i + 1]; v->u[2] = 0; v->node = n;
        n->verts.push_back(v); n->index = i; n->EoL = false; n->EoL_index = -1;
        mesh.nodes.push_back(n); mesh.verts.push_back(v);
    }
    if (argc > 3) {
        const int n = std::atoi(argv[3]);
        for (int i = 1; i + 1 < n; ++i) { Node *nd = mesh.nodes[(size_t)i * n + n / 2]; nd->EoL = true; nd->EoL_index = mesh.EoL_Count++; }
    }
    std::map<std::pair<Node *, Node *>, Edge *> edge_of;
    for (int k = 0; k < F; ++k) {
        Face *f = new Face;
        for (int j = 0; j < 3; ++j) f->v[j] = mesh.verts[fn[3 * (size_t)k + j]];
        mesh.faces.push_back(f);
        for (int i = 0; i < 3; ++i) {   // add_edges_if_needed
            Node *a = f->v[i]->node, *b = f->v[(i + 1) % 3]->node;
            auto key = std::make_pair(std::min(a, b), std::max(a, b));
            if (!edge_of.count(key)) { Edge *e = new Edge; e->n[0] = a; e->n[1] = b; e->adjf[0] = e->adjf[1] = nullptr; edge_of[key] = e; mesh.edges.push_back(e); }
        }
        for (int i = 0; i < 3; ++i) {   // adjf[side], side = 0 iff the face runs n[0] -> n[1]
            Node *v0 = f->v[(i + 1) % 3]->node, *v1 = f->v[(i + 2) % 3]->node;
            Edge *e = edge_of[std::make_pair(std::min(v0, v1), std::max(v0, v1))];
            e->adjf[e->n[0] == v0 ? 0 : 1] = f;
        }
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
    wr(out, flat.edge_stencil.data(), flat.edge_stencil.size());
    wr(out, forces.f.data(), forces.f.size());
    wr(out, forces.M.outer, (size_t)dof + 1); wr(out, forces.M.inner, (size_t)nnzM); wr(out, forces.M.values.data(), (size_t)nnzM);
    wr(out, forces.MDK.outer, (size_t)dof + 1); wr(out, forces.MDK.inner, (size_t)nnzK); wr(out, forces.MDK.values.data(), (size_t)nnzK);
    for (auto &c : cls) wr(out, c.get(), 1);
    fclose(out);
    std::printf("host_driver OK: dof=%d nnz(M)=%lld nnz(MDK)=%lld contacts=%d\n", dof, (long long)nnzM, (long long)nnzK, ncls);
    return 0;
}
