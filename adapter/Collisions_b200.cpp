// Collisions_b200.cpp — drop-in replacement of the reference's src/Collisions.cpp (CD, CD2), B200 backend.
//
// How to use: remove src/Collisions.cpp from the reference's source list and add this file (CMake snippet in
// INTEGRATION.md).  src/Collisions.h:9,11 and the call sites src/Scene.cpp:83,105 and src/Constraints.cpp:423 stay as they
// are; src/boxTriCollision.cpp stays in the build because Constraints / Preprocessor use btc::Collision.
//
// Needs the reference's headers and Eigen 3.3 (see Forces_fill_b200.cpp).
#include "Collisions.h"      // reference: CD, CD2, btc::Collision, Obstacles
#include "Box.h"
#include "Points.h"
#include "eolc_host.hpp"     // this repository: include/

using namespace std;
using namespace Eigen;

namespace {
void flatten_obstacles(const shared_ptr<Obstacles> &obs, eolc::host::ObstaclesFlat &o) {
    o.cdthreshold = obs->cdthreshold;
    o.num_points = obs->points->num_points;
    o.pxyz.assign(obs->points->pxyz.data(), obs->points->pxyz.data() + 3 * (size_t)o.num_points);       // 3 x P, column-major
    o.norms.assign(obs->points->norms.data(), obs->points->norms.data() + 3 * (size_t)o.num_points);
    o.num_boxes = obs->num_boxes;
    o.box_dim.clear(); o.box_E1.clear();
    for (int b = 0; b < obs->num_boxes; b++) {
        const Vector3d &d = obs->boxes[b]->dim;
        const Matrix4d &E = obs->boxes[b]->E1;
        o.box_dim.insert(o.box_dim.end(), d.data(), d.data() + 3);
        o.box_E1.insert(o.box_E1.end(), E.data(), E.data() + 16);                                        // column-major 4x4
    }
}

// eolc_contact (POD) -> btc::Collision (boxTriCollision.h:49-110)
shared_ptr<btc::Collision> to_btc(const eolc_contact &c) {
    auto r = make_shared<btc::Collision>();
    r->dist = c.dist;
    r->nor1 = Map<const Vector3d>(c.nor1); r->nor2 = Map<const Vector3d>(c.nor2);
    r->pos1 = Map<const Vector3d>(c.pos1); r->pos2 = Map<const Vector3d>(c.pos2); r->pos1_ = Map<const Vector3d>(c.pos1_);
    r->count1 = c.count1; r->count2 = c.count2;
    r->verts1 = Vector3i(c.verts1[0], c.verts1[1], c.verts1[2]); r->verts2 = Vector3i(c.verts2[0], c.verts2[1], c.verts2[2]);
    r->weights1 = Map<const Vector3d>(c.weights1); r->weights2 = Map<const Vector3d>(c.weights2);
    r->tri1 = c.tri1; r->tri2 = c.tri2;
    r->edge1.assign(c.edge1, c.edge1 + c.n_edge1);
    r->edge2 = c.edge2;
    r->edgeDir = Map<const Vector3d>(c.edgeDir);
    return r;
}

void run(const Mesh &mesh, const shared_ptr<Obstacles> &obs, vector<shared_ptr<btc::Collision> > &cls, bool cd1) {
    static thread_local eolc::host::FlatMesh flat;
    static thread_local eolc::host::ObstaclesFlat of;
    eolc::host::flatten_step(mesh, flat);                   // verts2 / faces2 of Collisions.cpp:13-27 (+ the stencils, unused here)
    flatten_obstacles(obs, of);
    // eolc_cd_plan_create (on a topology / threshold change) + eolc_cd_run; the records arrive in the thread's page-locked buffer and
    // are turned into btc::Collision objects straight from there (one allocation per contact, as the interface demands)
    const eolc::host::ContactView raw = cd1 ? eolc::host::CD_view(flat, of) : eolc::host::CD2_view(flat, of);
    cls.reserve(cls.size() + (size_t)raw.size);
    for (int32_t i = 0; i < raw.size; ++i) cls.push_back(to_btc(raw.data[i]));    // appended: the caller clears (Scene.cpp:93)
}
}  // namespace

void CD(const Mesh &mesh, const shared_ptr<Obstacles> obs, std::vector<std::shared_ptr<btc::Collision> > &cls) { run(mesh, obs, cls, true); }

void CD2(const Mesh &mesh, const shared_ptr<Obstacles> obs, std::vector<std::shared_ptr<btc::Collision> > &cls) { run(mesh, obs, cls, false); }
