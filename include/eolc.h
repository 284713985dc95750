/* eolc.h — C ABI of the B200-native EOL-Cloth hot path (libeolc_b200.so).
 *
 * This is the drop-in boundary under the reference's three C++ entry points
 *   Forces::fill   /root/reference/src/Forces.h:40     (body: src/Forces.cpp:912-930)
 *   CD             /root/reference/src/Collisions.h:9  (body: src/Collisions.cpp:11-53)
 *   CD2            /root/reference/src/Collisions.h:11 (body: src/Collisions.cpp:55-78)
 * whose call sites (src/Cloth.cpp:365, src/Scene.cpp:83, src/Constraints.cpp:423) stay
 * unchanged; the C++ adapter bodies that forward to these functions are shown in
 * INTEGRATION.md.  Plain pointers and sizes only; no C++/torch types cross this ABI.
 *
 * Conventions: every function returns 0 on success and a negative eolc_status on
 * error; eolc_last_error() returns the message of the last failure on this thread.
 * Arrays are caller-owned HOST pointers unless the function name ends in _dev.
 * A ctx binds one CUDA device; it must be used from one host thread at a time
 * (the reference has a single simulation thread, src/runner.cpp:58-63,121-129).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails
 * with EOLC_ERR_CUDA.
 *
 * Simplifying assumption (true for Cloth::build, src/Cloth.cpp:73-90): one Vert per
 * Node, so the material coordinate X = node->verts[0]->u is per node.
 * Both branches of Forces::fill are implemented: the Lagrangian one (mesh.EoL_Count == 0,
 * every BASELINE config) and, when eol_index names EoL nodes, the Eulerian-on-Lagrangian one
 * (src/Forces.cpp:177-329, 399-497, 580-683, 746-883): dof = 3N + 2 EoL_Count, the Eulerian
 * dofs of node a at 3N + 2 eol_index[a] (src/Forces.cpp:379-381).
 */
#ifndef EOLC_H_
#define EOLC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    EOLC_OK = 0,
    EOLC_ERR_ARG = -1,      /* bad argument (NULL, negative size, index out of range) */
    EOLC_ERR_CUDA = -2,     /* CUDA runtime failure / no device */
    EOLC_ERR_CAPACITY = -3, /* caller buffer too small (n_out still reports the needed size) */
    EOLC_ERR_UNSUPPORTED = -4
} eolc_status;

typedef struct eolc_ctx eolc_ctx;
typedef struct eolc_forces_plan eolc_forces_plan;
typedef struct eolc_cd_plan eolc_cd_plan;

/* mirrors `struct Material`, src/Cloth.h:28-35 (dampingA is never read by Forces.cpp) */
typedef struct {
    double density, e, nu, beta, dampingA, dampingB;
} eolc_material;

/* POD mirror of btc::Collision, src/boxTriCollision.h:49-110 (ctor boxTriCollision.cpp:103-120).
 * edge1 is a std::vector<int> of size 0, 1 or 3 in the reference -> edge1[3] + n_edge1. */
typedef struct {
    double dist;
    double nor1[3], nor2[3], pos1[3], pos2[3], pos1_[3];
    double weights1[3], weights2[3];
    double edgeDir[3];
    int32_t count1, count2;
    int32_t verts1[3], verts2[3];
    int32_t tri1, tri2;
    int32_t edge1[3];
    int32_t n_edge1;
    int32_t edge2;
    int32_t reserved;
} eolc_contact;

/* ---- context ---------------------------------------------------------------------- */
int eolc_ctx_create(int device, eolc_ctx **out);
void eolc_ctx_destroy(eolc_ctx *ctx);
const char *eolc_last_error(void);
int eolc_device_count(void);
/* the CUDA stream all work of this ctx is launched on (a cudaStream_t), for event timing */
void *eolc_ctx_stream(eolc_ctx *ctx);

/* ---- mesh flattening helper --------------------------------------------------------- */
/* Reproduces the edge list ArcSim builds while faces are added (Mesh::add(Face),
 * src/external/ArcSim/mesh.cpp:356-378): edge order = creation order, n[0]->n[1] = direction of
 * the creating face, adjf[0] = creating face.  Writes 4 ints per edge:
 * (n0, n1, opp(adjf[0]), opp(adjf[1])), -1 where the face is absent — exactly the stencil
 * edgeBasedF reads (src/Forces.cpp:688-697).  edge_stencil must hold 4*3F ints. */
int eolc_mesh_edge_stencils(int32_t N, int32_t F, const int32_t *face_nodes, int32_t *E_out, int32_t *edge_stencil);

/* ---- Forces::fill ------------------------------------------------------------------- */
/* Topology plan: CSR pattern of M and MDK + element->slot maps. Rebuild only after a remesh /
 * set_indices (src/Scene.cpp:87-90).  X_hint (2N, may be NULL) is used only to group nodes into
 * spatially compact tiles; results do not depend on it. eol_index (N, may be NULL): Node::EoL_index of the EoL nodes, -1 =
 * Lagrangian node; the indices must be distinct and EoL_Count = 1 + the largest (mesh.EoL_Count of the reference).
 * Limits (EOLC_ERR_UNSUPPORTED otherwise): at most 255 neighbours per node; the faces and bending stencils touching one node must fit
 * one tile (about 150 faces + 150 stencils, e.g. a fan of valence 150); bending stencils may not repeat a node.  The
 * device consumers below work on both kinds of plan (block structure for Lagrangian plans, Eigen's scalar arrays for EOL plans). */
int eolc_forces_plan_create(eolc_ctx *ctx, int32_t N, int32_t F, const int32_t *face_nodes, int32_t E,
                            const int32_t *edge_stencil, const int32_t *eol_index, const double *X_hint,
                            eolc_forces_plan **out);
void eolc_forces_plan_destroy(eolc_forces_plan *plan);
/* which: 0 = M, 1 = MDK.  outer (dof+1) / inner (nnz, ascending) are host arrays owned by the plan and are
 * identical to Eigen's outerIndexPtr()/innerIndexPtr() of the column-major matrices the reference builds with
 * setFromTriplets (src/Forces.cpp:928-929); both matrices are symmetric, so CSC == CSR. */
int eolc_forces_pattern(const eolc_forces_plan *plan, int which, int32_t *dof, int64_t *nnz, const int32_t **outer,
                        const int32_t **inner);
/* counts for the metric: faces + interior edges assembled per fill */
int eolc_forces_counts(const eolc_forces_plan *plan, int32_t *n_faces, int32_t *n_interior_edges);
/* One Forces::fill.  x: 3N (Node::x), X: 2N (verts[0]->u), outputs f (dof = 3N + 2 EoL_Count), M_vals (nnz(M)), MDK_vals (nnz(MDK)).
 * Host version copies x/X in and f/M/MDK out (pinned staging inside the plan). */
int eolc_forces_fill(eolc_forces_plan *plan, const double *x, const double *X, const eolc_material *mat,
                     const double grav[3], double h, double *f, double *M_vals, double *MDK_vals);
/* Flags of the *_ex entry points.
 * EOLC_FILL_M_UNCHANGED: the caller states that X and the density are the ones of the previous fill on this plan, so M, which
 *   depends on nothing else (src/ComputeInertial.cpp:33,44-47; SURVEY §8a row 5: "constant between remeshes when EOL is off"),
 *   is the same matrix: M_vals is left untouched (host entry: no device-to-host copy of M and no second upload of X, the plan
 *   keeps the previous host fill's copy on the device; device entries: the M rows are neither recomputed nor written).  f and MDK are produced as always and are bit-identical to a full fill.  Honoured for
 *   Lagrangian plans only (with EoL nodes M depends on x through F = deform_grad); otherwise M is recomputed.  M_vals must be
 *   the buffer of the previous fill or at least a valid one.
 * EOLC_FILL_EXACT_SYMMETRY (device entries; the host entries always do it): MDK symmetric bit for bit, like the reference's mirrored
 *   triplets (src/Forces.cpp:114-125, :531-539).  Pairs of nodes owned by one tile are summed once and mirrored anyway; the blocks of
 *   pairs across two tiles agree to rounding only, and this flag adds a pass that copies the lower node's block onto the higher
 *   node's transposed block (for plans with EoL nodes too: their Eulerian rows / columns are mirrored from one source already). */
#define EOLC_FILL_M_UNCHANGED 1u
#define EOLC_FILL_EXACT_SYMMETRY 2u
int eolc_forces_fill_ex(eolc_forces_plan *plan, const double *x, const double *X, const eolc_material *mat,
                        const double grav[3], double h, double *f, double *M_vals, double *MDK_vals, uint32_t flags);
/* Page-locked host memory for the buffers handed to the host entry points (x, X, f, M_vals, MDK_vals, contact lists): such
 * buffers are the target of the DMA itself, pageable ones cost a staging copy (1.5 GB per fill at 1024^2).  NULL on failure. */
void *eolc_host_alloc(size_t bytes);
void eolc_host_free(void *p);
/* The same for an array the CALLER owns and cannot allocate through eolc_host_alloc — e.g. valuePtr() of the Eigen::SparseMatrix
 * members of the reference's class Forces (src/Forces.h:35-36): eolc_host_register page-locks it in place (cudaHostRegister), after
 * which eolc_forces_fill[_ex] copies device -> that array directly; eolc_host_unregister before the array is freed or reallocated.
 * Registering is slow (about 0.1 s per GB): once per topology change, not per step. */
int eolc_host_register(void *p, size_t bytes);
int eolc_host_unregister(void *p);
/* Same, every pointer is a DEVICE pointer on the plan's device; asynchronous on eolc_ctx_stream(). */
int eolc_forces_fill_dev(eolc_forces_plan *plan, const double *x_dev, const double *X_dev, const eolc_material *mat,
                         const double grav[3], double h, double *f_dev, double *M_vals_dev, double *MDK_vals_dev);
/* Ensemble: n_scenes independent states sharing the plan's topology; scene s reads x_dev + s*3N, X_dev + s*2N and
 * writes f_dev + s*dof, M_vals_dev + s*nnz(M), MDK_vals_dev + s*nnz(MDK). */
int eolc_forces_fill_batched_dev(eolc_forces_plan *plan, int32_t n_scenes, const double *x_dev, const double *X_dev,
                                 const eolc_material *mat, const double grav[3], double h, double *f_dev,
                                 double *M_vals_dev, double *MDK_vals_dev);
int eolc_forces_fill_batched_dev_ex(eolc_forces_plan *plan, int32_t n_scenes, const double *x_dev, const double *X_dev,
                                    const eolc_material *mat, const double grav[3], double h, double *f_dev,
                                    double *M_vals_dev, double *MDK_vals_dev, uint32_t flags);
/* ---- consumer of the fill, on the device -------------------------------- */
/* Cloth::solve right-hand side, src/Cloth.cpp:345:  b = -(M v + h f).  M_vals_dev / f_dev: outputs of a fill with this plan;
 * v_dev, b_dev: dof doubles.  Asynchronous on eolc_ctx_stream(). */
int eolc_forces_rhs_dev(eolc_forces_plan *plan, const double *M_vals_dev, const double *f_dev, const double *v_dev, double h,
                        double *b_dev);
/* GeneralizedSolver::velocitySolve, collision-free branch without fixed points (src/GeneralizedSolver.cpp:120-126):
 * ConjugateGradient<SparseMatrix<double>, Lower|Upper> cg; cg.compute(MDK); v = cg.solve(-b) — Eigen's CG with its default
 * diagonal preconditioner, started from 0, stopped at ||MDK v + b|| <= tol ||b|| (Eigen's default tol is DBL_EPSILON, its default
 * iteration cap 2 dof).  fixed_dev (dof bytes, may be NULL): dofs with a non-zero flag keep velocity 0 and drop out of the system
 * (fixed points with zero prescribed velocity, the `fixedPoints` case the reference hands to a KKT / QP solve); the residual is
 * then that of the free dofs.  v_dev receives the solution; iters_out (may be NULL) the iterations issued (a multiple of the
 * host's check interval), rel_resid_out (may be NULL) the final relative residual.  Synchronises the stream. */
int eolc_solve_cg_dev(eolc_forces_plan *plan, const double *MDK_vals_dev, const double *b_dev, const unsigned char *fixed_dev,
                      double *v_dev, double tol, int32_t max_iter, int32_t *iters_out, double *rel_resid_out);
/* Position update of Cloth::step (src/Cloth.cpp:394-400): x += h v for the 3N Lagrangian dofs.  Asynchronous on the stream. */
int eolc_forces_integrate_dev(eolc_forces_plan *plan, const double *v_dev, double h, double *x_dev);
/* Its Eulerian part (src/Cloth.cpp:401-407): vert->u += h vert->v for the EoL nodes, X_dev (2N) updated in place from the Eulerian
 * entries of v_dev (dof).  No-op for a plan without EoL nodes. */
int eolc_forces_integrate_X_dev(eolc_forces_plan *plan, const double *v_dev, double h, double *X_dev);
/* ---- per-step derived mesh data ---------------------------------------------- */
/* World-space normals of compute_ws_data (src/external/ArcSim/mesh.cpp:135-140, 142-143): face_n (3F) = normalize(cross(x1 - x0,
 * x2 - x0)); node_n (3N) = normal<WS>(node), src/external/ArcSim/geometry.cpp:302-316 (sum over the node's faces, in ascending face
 * index = vert->adjf order of a mesh whose faces were added in index order, of cross(e1, e2) / (2 |e1|^2 |e2|^2), normalized; a
 * node without faces gets 0).  These are what Cloth::updatePosNor (src/Cloth.cpp:150-171) and Constraints::fill
 * (src/Constraints.cpp:181, 296-318) read as node->n / face->n.  Either output may be NULL.  The _dev form takes device pointers
 * and is asynchronous on eolc_ctx_stream(); the node curvature of compute_ws_data(Node*) is remesher state and not produced. */
int eolc_mesh_normals(eolc_forces_plan *plan, const double *x, double *face_n, double *node_n);
int eolc_mesh_normals_dev(eolc_forces_plan *plan, const double *x_dev, double *face_n_dev, double *node_n_dev);
/* number of kernels one fill launches (for bench accounting): 1, or 3 with EoL nodes */
int eolc_forces_launches_per_fill(const eolc_forces_plan *plan);

/* ---- CD / CD2 ----------------------------------------------------------------------- */
/* Builds the btc edge table in createEdges order (src/boxTriCollision.cpp:141-231, 64-bit sort key — see
 * DESIGN.md) and the perturbation table for (N, threshold) (std::mt19937 seed 1, :648-659). */
int eolc_cd_plan_create(eolc_ctx *ctx, int32_t N, int32_t F, const int32_t *face_nodes, double threshold,
                        eolc_cd_plan **out);
void eolc_cd_plan_destroy(eolc_cd_plan *plan);
int eolc_cd_edge_count(const eolc_cd_plan *plan);
/* host copy of the btc edge table: 4 verts + 2 faces per edge (6 ints), createEdges order */
int eolc_cd_edge_table(const eolc_cd_plan *plan, int32_t *out6E);
/* CD  = eolc_cd_run(..., point_eol_flag=1, remap_box_indices=1)   (src/Collisions.cpp:30,39-48)
 * CD2 = eolc_cd_run(..., point_eol_flag=0, remap_box_indices=0)   (src/Collisions.cpp:71)
 * pxyz / pnorms: 3 x n_points column-major (Points::pxyz / norms); box_whd: 3 per box (Box::dim);
 * box_E: 16 per box, column-major 4x4 (Box::E1).  Contacts are written in the reference's list order.
 * On EOLC_ERR_CAPACITY *n_out holds the required capacity. */
int eolc_cd_run(eolc_cd_plan *plan, const double *x, int32_t n_points, const double *pxyz, const double *pnorms,
                int32_t n_boxes, const double *box_whd, const double *box_E, int point_eol_flag, int remap_box_indices,
                eolc_contact *out, int32_t capacity, int32_t *n_out);
/* x_dev is a device pointer (3N); contacts still land in host memory (the list is small). */
int eolc_cd_run_dev(eolc_cd_plan *plan, const double *x_dev, int32_t n_points, const double *pxyz, const double *pnorms,
                    int32_t n_boxes, const double *box_whd, const double *box_E, int point_eol_flag,
                    int remap_box_indices, eolc_contact *out, int32_t capacity, int32_t *n_out);
/* Ensemble: n_scenes states (x_dev + s*3N) against the same obstacles.  Contacts of scene s are written to
 * out[scene_offset[s] .. scene_offset[s+1]) ; scene_offset has n_scenes+1 entries. */
int eolc_cd_run_batched_dev(eolc_cd_plan *plan, int32_t n_scenes, const double *x_dev, int32_t n_points,
                            const double *pxyz, const double *pnorms, int32_t n_boxes, const double *box_whd,
                            const double *box_E, int point_eol_flag, int remap_box_indices, eolc_contact *out,
                            int32_t capacity, int32_t *scene_offset);
/* Same run, but the records stay on the device (no device-to-host copy of the list; only the per-scene offsets come back):
 * for a consumer that lives on the device too — eolc_cd_contact_rows below (112 B of rows per contact instead of the 264 B
 * record), or any kernel reading eolc_cd_contacts_dev.  Step (D) needs no pass: section Bc emits at most one record per
 * corner by construction (a lexicographic argmin).  The device buffer is owned by the plan and valid until its next run.
 * The call returns with the offsets known and the records possibly still being written on the context's stream (work queued on
 * that stream afterwards — the row entry points below, the caller's own kernels — is ordered behind them).  From the plan's
 * second run on, the record pass is launched before the host has seen the totals, guarded by the existing buffers' capacities,
 * and repeated if they were too small (EOLC_CD_NO_SPECULATION=1 in the environment switches that off). */
int eolc_cd_run_batched_resident_dev(eolc_cd_plan *plan, int32_t n_scenes, const double *x_dev, int32_t n_points,
                                     const double *pxyz, const double *pnorms, int32_t n_boxes, const double *box_whd,
                                     const double *box_E, int point_eol_flag, int remap_box_indices, int32_t *scene_offset);
int eolc_cd_contacts_dev(const eolc_cd_plan *plan, const eolc_contact **contacts_dev, int32_t *count);
/* Limits of one run: n_scenes * max(n_boxes, n_points, 1) <= 65535 (a grid dimension); larger batches / point clouds must be
 * split by the caller (EOLC_ERR_ARG otherwise). */
/* Host-only diagnostic (no GPU needed).  Section C's three acos() (src/boxTriCollision.cpp:879-887, :907-915) only feed
 * threshold comparisons; the device decides them in cosine space against critical doubles that the host finds by bisection
 * WITH ITS OWN libm acos (the function a reference build on the same machine calls), checking every double in a window on
 * either side of each switch — so the contact set cannot differ from a libm-linked reference through acos rounding.
 * cuts14 = [cosParHi, cosParLo, cosWedge[0..11]]:
 *   |acos c| < 2 deg        <=>  cosParHi <= c <= 1
 *   |pi - acos c| < 2 deg   <=>  -1 <= c <= cosParLo
 *   acos c - acos(n1c.n1d of box edge k) > 2 deg  <=>  -1 <= c < cosWedge[k]   (and c <= 1)
 * Returns EOLC_ERR_UNSUPPORTED if libm's acos is not monotone inside a window (eolc_cd_run then refuses to run). */
int eolc_cd_angle_cuts(const double *box_whd, const double *box_E, double *cuts14);
/* ---- consumer of the contact list ------------------------------------------ */
/* Constraints::fill, contact part (src/Constraints.cpp:424-468): one inequality row per CD2 contact, in list order,
 *   (3,1) cloth vertex / box face : 3 entries  -nor1[k]               at column 3 verts2[0] + k
 *   (2,2) edge / edge             : 6 entries  -nor2[k] weights2[j]   at column 3 verts2[j] + k, j = 0, 1
 *   (1,3) box corner / cloth tri  : 9 entries  -nor2[k] weights2[j]   at column 3 verts2[j] + k, j = 0, 1, 2
 * i.e. exactly the triplets Aineq_ receives (row = running ineqsize, the reference's push order); bineq stays zero.  A contact
 * touching an EoL node is skipped and takes no row (node_eol: N flags, may be NULL = no EoL node).  Output in a fixed-width
 * layout: row i has row_nnz[i] entries cols[9 i ..], vals[9 i ..] (unused entries: col -1, val 0); arrays sized for n rows.
 * eolc_constraints_contact_rows works on a host contact list; eolc_cd_contact_rows on the contacts of the plan's LAST run, which are
 * still on the device (the rows are 112 B per contact instead of the 264 B record: for a host that only builds Aineq). */
int eolc_constraints_contact_rows(const eolc_contact *contacts, int32_t n, const uint8_t *node_eol, int32_t *n_rows,
                                  int32_t *row_nnz, int32_t *cols, double *vals);
int eolc_cd_contact_rows(eolc_cd_plan *plan, const uint8_t *node_eol, int32_t capacity_rows, int32_t *n_rows, int32_t *row_nnz,
                         int32_t *cols, double *vals);
/* The same rows in compressed (CSR) form, compacted on the device: row r holds entries row_ptr[r] .. row_ptr[r + 1) of cols / vals —
 * the reference's exact Aineq_ triplet sequence (row = running ineqsize), 40 B per (3,1) contact instead of 112 B fixed-width or
 * the 264 B record.  *n_rows and *nnz are always set; EOLC_ERR_CAPACITY (nothing copied) if they exceed capacity_rows /
 * capacity_nnz (row_ptr holds capacity_rows + 1 entries; nnz <= 9 * eolc_cd_last_count).  Page-locked arrays (eolc_host_alloc) are
 * the DMA targets themselves; pageable ones cost a staging copy. */
int eolc_cd_contact_rows_csr(eolc_cd_plan *plan, const uint8_t *node_eol, int32_t capacity_rows, int32_t capacity_nnz,
                             int32_t *n_rows, int32_t *nnz, int32_t *row_ptr, int32_t *cols, double *vals);
/* Constraints::fill, fixed-corner part (src/Constraints.cpp:114-119, :470-497): the equality rows of the four corner records of a
 * FixedList (src/FixedList.h:32-37).  corner k: c[6 k .. 6 k + 5] = (mask x, y, z, prescribed velocity x, y, z), node index ci[k];
 * a corner with c[6 k] == -1 is absent, a component with mask == 1.0 gets one row
 *     Aeq(row, 3 ci[k] + j) = mask_j ,   beq(row) = (1 - 0.01) v[3 ci[k] + j] + c[6 k + 3 + j]        (v: 3N node velocities)
 * in the reference's push order (corners 1..4, components x, y, z), rows numbered from eq_row0 (the reference's running eqsize).
 * Host-only (at most 12 rows).  Outputs hold 12 entries; *n_rows rows are written. */
int eolc_constraints_fixed_rows(const double *c /*24*/, const int32_t *ci /*4*/, const double *v, int32_t n_nodes, int32_t eq_row0,
                                int32_t *n_rows, int32_t *rows, int32_t *cols, double *vals, double *beq);
/* number of contacts of the plan's last run (all scenes) */
int eolc_cd_last_count(const eolc_cd_plan *plan);
/* counters of the last run: candidate pair tests executed on the device (A: N*24, B: 8*F, C: E*12 after culls) */
int eolc_cd_last_stats(const eolc_cd_plan *plan, int64_t *pair_tests, int32_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* EOLC_H_ */
