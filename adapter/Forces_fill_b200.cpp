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

#include <iostream>

namespace {
// One plan cache per Forces object would be cleaner, but Forces.h must stay untouched: keyed by `this`.
struct Backend {
    eolc::host::Forces forces;
    eolc::host::FlatMesh flat;
};
Backend &backend_of(const Forces *self) {
    eolc::host::Context::instance();   // constructed before the table below, hence destroyed after the plans it holds
    static thread_local std::vector<std::pair<const Forces *, std::unique_ptr<Backend> > > table;
    for (auto &e : table) if (e.first == self) return *e.second;
    table.emplace_back(self, std::unique_ptr<Backend>(new Backend));
    return *table.back().second;
}
}  // namespace

void Forces::fill(const Mesh &mesh, const Material &mat, const Eigen::Vector3d &grav, double h) {
    Backend &B = backend_of(this);
    // Flatten the ArcSim pointer mesh into page-locked arrays.  flatten() compares while it copies and renews
    // B.flat's version counters only when an index / a material coordinate really changed; fill() rebuilds the device plan when
    // the topology version moved (dynamic_remesh / preprocess, Scene.cpp:83-90) and skips M when X and the density did not.
    eolc::host::flatten(mesh, B.flat);
    // EoL nodes (flat.eol_index) switch the touched elements to the Eulerian-on-Lagrangian blocks (Forces.cpp:177-329, 399-497,
    // 580-683, 746-883) inside the library; dof = 3N + 2 (1 + largest EoL_index) must agree with mesh.EoL_Count.
    if ((int)mesh.EoL_Count != B.flat.EoL_Count) {
        std::cout << "Forces::fill (B200): mesh.EoL_Count does not match the nodes flagged EoL" << std::endl;
        abort();
    }
    eolc_material m = {mat.density, mat.e, mat.nu, mat.beta, mat.dampingA, mat.dampingB};
    const double g[3] = {grav(0), grav(1), grav(2)};
    B.forces.fill(B.flat, m, g, h);                       // eolc_forces_plan_create (on topology change) + eolc_forces_fill

    const int dof = (int)mesh.nodes.size() * 3 + mesh.EoL_Count * 2;     // Forces.cpp:914
    f = Eigen::Map<const Eigen::VectorXd>(B.forces.f.data(), dof);
    EoL_cutoff = B.forces.EoL_cutoff;                                     // Forces.cpp:919
    // The plan's pattern IS Eigen's compressed column-major layout (outerIndexPtr / innerIndexPtr / valuePtr), so the
    // matrices are assigned from a Map without any triplet pass.
    typedef Eigen::Map<const Eigen::SparseMatrix<double> > SpMap;
    // M depends on X and the density only (ComputeInertial.cpp:33,44-47): when neither changed the member still holds it
    if (B.forces.M_updated || M.rows() != dof || M.nonZeros() != (Eigen::Index)B.forces.M.nnz)
        M = SpMap(dof, dof, (Eigen::Index)B.forces.M.nnz, B.forces.M.outer, B.forces.M.inner, B.forces.M.values.data());
    MDK = SpMap(dof, dof, (Eigen::Index)B.forces.MDK.nnz, B.forces.MDK.outer, B.forces.MDK.inner, B.forces.MDK.values.data());
}
