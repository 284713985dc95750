// Forces_fill_b200.cpp — drop-in body of Forces::fill for the reference tree (sueda/eol-cloth), B200 backend.
//
// How to use: in the reference's src/Forces.cpp delete (or #if 0) the body of
//     void Forces::fill(const Mesh& mesh, const Material& mat, const Vector3d& grav, double h)      (Forces.cpp:912-930)
// and add this file to the build (CMake snippet in INTEGRATION.md).  Nothing else changes: the declaration in
// src/Forces.h:40, the members f / M / MDK / EoL_cutoff (Forces.h:34-38) and the call site src/Cloth.cpp:365 stay as
// they are, so Cloth::step, GeneralizedSolver::velocitySolve and the Mosek / Gurobi wrappers keep receiving
// Eigen::VectorXd / Eigen::SparseMatrix<double>.
//
// This file needs the reference's headers and Eigen 3.3 (neither is available in the development container of this
// repository, where the same code path is exercised through include/eolc_host.hpp by tests/cpp/host_driver.cpp).
#include "Forces.h"          // reference: class Forces, struct Material (via Cloth.h)
#include "eolc_host.hpp"     // this repository: include/

#include <algorithm>
#include <iostream>
#include <memory>
#include <vector>

// Where the results land.  Default: the library writes into the host layer's page-locked arrays and the values are copied into the
// Eigen members on the host threads (one pass over M / MDK).  With -DEOLC_ADAPTER_ZERO_COPY the members' own value arrays are
// page-locked IN PLACE (eolc_host_register) and handed to the library as its output buffers, so the device -> host copy of a fill
// lands in the members themselves (1024^2 sheet: 58 instead of ~85 ms per step).  That mode has a lifetime rule the reference's
// headers cannot enforce: a registered array must be unregistered before it is freed — call eolc_adapter_forces_release(this) from
// Forces::~Forces() (Forces.h:32), or keep the Forces object alive until the process ends (Cloth owns one for its whole life,
// Cloth.cpp:26-34).  A member reallocated behind the adapter's back (someone assigns to M / MDK) is noticed and registered again.
namespace {
// One plan cache per Forces object would be cleaner, but Forces.h must stay untouched: keyed by `this`.
struct Backend {
    eolc::host::Forces forces;
    eolc::host::FlatMesh flat;
    double *reg[3] = {nullptr, nullptr, nullptr};   // zero-copy mode: the registered value arrays of f / M / MDK
    bool zero_copy = true;          // false after a failed registration: results then go through the host layer's own arrays
    ~Backend() { release(); }
    void release() { for (double *&p : reg) { eolc_host_unregister(p); p = nullptr; } }
};
typedef std::vector<std::pair<const Forces *, std::unique_ptr<Backend> > > Table;
Table &table() {
    eolc::host::Context::instance();   // constructed before the table below, hence destroyed after the plans it holds
    static thread_local Table t;
    return t;
}
Backend &backend_of(const Forces *self) {
    Table &t = table();
    for (auto &e : t) if (e.first == self) return *e.second;
    t.emplace_back(self, std::unique_ptr<Backend>(new Backend));
    return *t.back().second;
}
// An Eigen::SparseMatrix<double> with the plan's pattern (the plan's arrays ARE Eigen's compressed column-major layout); values are
// written by the fill.
void shape_like(Eigen::SparseMatrix<double> &A, int dof, const eolc::host::SparseCSC &P) {
    A.resize(dof, dof);
    A.resizeNonZeros((Eigen::Index)P.nnz);
    std::copy(P.outer, P.outer + dof + 1, A.outerIndexPtr());
    std::copy(P.inner, P.inner + P.nnz, A.innerIndexPtr());
}
}  // namespace

// Drops what the adapter holds for one Forces object (device plan, page-locked buffers, registrations).  See the lifetime rule above.
extern "C" void eolc_adapter_forces_release(const void *forces_this) {
    Table &t = table();
    for (size_t i = 0; i < t.size(); ++i)
        if ((const void *)t[i].first == forces_this) { t.erase(t.begin() + (std::ptrdiff_t)i); return; }
}

void Forces::fill(const Mesh &mesh, const Material &mat, const Eigen::Vector3d &grav, double h) {
    Backend &B = backend_of(this);
    // Flatten the ArcSim pointer mesh into page-locked arrays.  flatten() compares while it copies and renews
    // B.flat's version counters only when an index / a material coordinate really changed; fill() rebuilds the device plan when
    // the topology version moved (dynamic_remesh / preprocess, Scene.cpp:83-90) and skips M when X and the density did not.
    eolc::host::flatten_step(mesh, B.flat);   // full walk, or positions only under EOLC_STATIC_TOPOLOGY=1
    // EoL nodes (flat.eol_index) switch the touched elements to the Eulerian-on-Lagrangian blocks (Forces.cpp:177-329, 399-497,
    // 580-683, 746-883) inside the library; dof = 3N + 2 (1 + largest EoL_index) must agree with mesh.EoL_Count.
    if ((int)mesh.EoL_Count != B.flat.EoL_Count) {
        std::cout << "Forces::fill (B200): mesh.EoL_Count does not match the nodes flagged EoL" << std::endl;
        abort();
    }
    eolc_material m = {mat.density, mat.e, mat.nu, mat.beta, mat.dampingA, mat.dampingB};
    const double g[3] = {grav(0), grav(1), grav(2)};
    const int dof = (int)mesh.nodes.size() * 3 + mesh.EoL_Count * 2;     // Forces.cpp:914
    EoL_cutoff = (int)mesh.nodes.size() * 3;                              // Forces.cpp:919
    // After a topology change (or if a member was resized behind our back) the members take the plan's pattern — the plan's arrays
    // ARE Eigen's compressed column-major layout (outerIndexPtr / innerIndexPtr / valuePtr), no triplet pass — once per remesh.
    const bool rebuilt = B.forces.prepare(B.flat);                        // eolc_forces_plan_create on a topology change
    const bool reshaped = rebuilt || f.size() != dof || M.rows() != dof || MDK.rows() != dof ||
                          M.nonZeros() != (Eigen::Index)B.forces.M.nnz || MDK.nonZeros() != (Eigen::Index)B.forces.MDK.nnz;
    if (reshaped) {
        B.release();
        f.resize(dof);
        shape_like(M, dof, B.forces.M);
        shape_like(MDK, dof, B.forces.MDK);
    }
#ifdef EOLC_ADAPTER_ZERO_COPY
    if (B.zero_copy && (reshaped || B.reg[0] != f.data() || B.reg[1] != M.valuePtr() || B.reg[2] != MDK.valuePtr())) {
        B.release();
        double *p[3] = {f.data(), M.valuePtr(), MDK.valuePtr()};
        const size_t n[3] = {(size_t)dof, (size_t)B.forces.M.nnz, (size_t)B.forces.MDK.nnz};
        for (int k = 0; k < 3 && B.zero_copy; ++k) {
            if (n[k] == 0) continue;
            if (eolc_host_register(p[k], n[k] * sizeof(double)) == EOLC_OK) B.reg[k] = p[k];
            else { B.release(); B.zero_copy = false; }                    // e.g. a locked-memory limit: fall back to the copying path
        }
    }
    if (B.zero_copy) {
        eolc::host::Forces::External out = {f.data(), M.valuePtr(), MDK.valuePtr()};
        B.forces.fill(B.flat, m, g, h, &out);             // eolc_forces_fill_ex; M is left alone when X and the density did not change
        return;
    }
#endif
    B.forces.fill(B.flat, m, g, h);                       // eolc_forces_fill_ex into the host layer's page-locked arrays
    std::copy(B.forces.f.data(), B.forces.f.data() + dof, f.data());
    // M depends on X and the density only (ComputeInertial.cpp:33,44-47): when neither changed the member still holds it
    if (B.forces.M_updated || reshaped) eolc::host::par_copy(M.valuePtr(), B.forces.M.values.data(), (size_t)B.forces.M.nnz);
    eolc::host::par_copy(MDK.valuePtr(), B.forces.MDK.values.data(), (size_t)B.forces.MDK.nnz);
}
