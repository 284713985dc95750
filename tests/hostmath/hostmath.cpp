// Test shim: compiles the SAME element arithmetic the CUDA kernels use (eol_cloth_b200/csrc/elements.cuh)
// for the host, so the hand-derived formulas can be checked against the oracle without a GPU.
// Test infrastructure only — never linked into libeolc_b200.so.
#include "../../eol_cloth_b200/csrc/elements.cuh"
#include "../../eol_cloth_b200/csrc/forces_plan.h"
#include "../../eol_cloth_b200/csrc/tile_exec.cuh"
#include "../../eol_cloth_b200/csrc/forces_eol.h"
using namespace eolc;
extern "C" {
void hostmath_face(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                   const double *Xc, double e, double nu, double rho, const double *g, double dhh,
                   double *f9, double *t8, double *K6x9) {
    FaceOut o;
    face_element(mk3(xa[0], xa[1], xa[2]), mk3(xb[0], xb[1], xb[2]), mk3(xc[0], xc[1], xc[2]), Xa[0], Xa[1], Xb[0], Xb[1],
                 Xc[0], Xc[1], e, nu, rho, mk3(g[0], g[1], g[2]), dhh, o);
    for (int i = 0; i < 3; ++i) { f9[i] = o.fa[i]; f9[3 + i] = o.fb[i]; f9[6 + i] = o.fc[i]; }
    *t8 = o.t8;
    for (int b = 0; b < 6; ++b) for (int k = 0; k < 9; ++k) K6x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_edge(const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                   const double *X1, const double *X2, const double *X3, double beta, double dhh, double *K10x9) {
    EdgeOut o;
    edge_element(mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
                 X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, dhh, o);
    for (int b = 0; b < 10; ++b) for (int k = 0; k < 9; ++k) K10x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_face_row(int v, const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                       const double *Xc, double e, double nu, double rho, const double *g, double dhh,
                       double *f3, double *t8, double *K3x9) {
    FaceRowOut o;
    face_row(v, mk3(xa[0], xa[1], xa[2]), mk3(xb[0], xb[1], xb[2]), mk3(xc[0], xc[1], xc[2]), Xa[0], Xa[1], Xb[0], Xb[1],
             Xc[0], Xc[1], membrane_mu(e, nu), membrane_lambda(e, nu), rho, mk3(g[0], g[1], g[2]), dhh, o);
    for (int i = 0; i < 3; ++i) f3[i] = o.f[i];
    *t8 = o.md * 12.0;
    for (int b = 0; b < 3; ++b) for (int k = 0; k < 9; ++k) K3x9[9 * b + k] = o.K[b].m[k];
}
void hostmath_edge_row(int i, const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                       const double *X1, const double *X2, const double *X3, double beta, double dhh, double *K4x9) {
    EdgeRowOut o;
    edge_row(i, mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
             X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, dhh, o);
    for (int b = 0; b < 4; ++b) for (int k = 0; k < 9; ++k) K4x9[9 * b + k] = o.K[b].m[k];
}

void hostmath_edge_force(const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                         const double *X1, const double *X2, const double *X3, double beta, double *f12) {
    v3 f[4];
    eol::edge_force(mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
                    X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, f);
    for (int i = 0; i < 4; ++i) { f12[3 * i] = f[i].x; f12[3 * i + 1] = f[i].y; f12[3 * i + 2] = f[i].z; }
}

// tiles pipeline element forms
struct CollectEdge {
    double *K;   // 10 x 9: 00,11,22,33,01,02,03,12,13,23
    void diag(int i, const sym3 &S) { double *d = K + 9 * i; d[0] = S.xx; d[1] = S.xy; d[2] = S.xz; d[3] = S.xy; d[4] = S.yy; d[5] = S.yz; d[6] = S.xz; d[7] = S.yz; d[8] = S.zz; }
    void off(int k, const blk3 &B) { for (int q = 0; q < 9; ++q) K[9 * (4 + k) + q] = B.m[q]; }
};
void hostmath_edge_tile(const double *x0, const double *x1, const double *x2, const double *x3, const double *X0,
                        const double *X1, const double *X2, const double *X3, double beta, double dhh, double *K10x9) {
    CollectEdge c{K10x9};
    edge_element_tile(mk3(x0[0], x0[1], x0[2]), mk3(x1[0], x1[1], x1[2]), mk3(x2[0], x2[1], x2[2]), mk3(x3[0], x3[1], x3[2]),
                      X0[0], X0[1], X1[0], X1[1], X2[0], X2[1], X3[0], X3[1], beta, dhh, c);
}
struct CollectFace {
    double *K, *f, *t8;   // K: 6 x 9: aa,bb,cc,ab,ac,bc
    void diag(int i, const sym3 &S) { double *d = K + 9 * i; d[0] = S.xx; d[1] = S.xy; d[2] = S.xz; d[3] = S.xy; d[4] = S.yy; d[5] = S.yz; d[6] = S.xz; d[7] = S.yz; d[8] = S.zz; }
    void off(int k, const blk3 &B) { for (int q = 0; q < 9; ++q) K[9 * (3 + k) + q] = B.m[q]; }
    void force(int i, v3 v) { f[3 * i] = v.x; f[3 * i + 1] = v.y; f[3 * i + 2] = v.z; }
    void mass(double m) { *t8 = m; }
};
void hostmath_face_tile(const double *xa, const double *xb, const double *xc, const double *Xa, const double *Xb,
                        const double *Xc, double e, double nu, double rho, const double *g, double dhh,
                        double *f9, double *t8, double *K6x9) {
    CollectFace c{K6x9, f9, t8};
    face_element_tile(mk3(xa[0], xa[1], xa[2]), mk3(xb[0], xb[1], xb[2]), mk3(xc[0], xc[1], xc[2]), Xa[0], Xa[1], Xb[0], Xb[1],
                      Xc[0], Xc[1], membrane_mu(e, nu), membrane_lambda(e, nu), rho, mk3(g[0], g[1], g[2]), dhh, c);
}

// ---- host emulation of the "tiles" kernel: same plan builder (forces_plan.h), same per-thread phases (tile_exec.cuh),
// threads run one after the other, the barrier between the phases is the loop boundary.
struct HmPlan {
    int32_t N, F, Ei;
    std::vector<int32_t> fn, ie;
    Pattern pat;
    tiles::Plan tp;
    bool has_eol = false;
    eol::Plan ep;
};
// eol_index: N ints (-1 = Lagrangian) or NULL
void *hm_plan_create_eol(int32_t N, int32_t F, const int32_t *fn, int32_t E, const int32_t *es, const double *X_hint, int dedup, const int32_t *eol_index,
                         char *err, int errlen) {
    HmPlan *P = new HmPlan;
    P->N = N; P->F = F;
    P->fn.assign(fn, fn + 3 * (size_t)F);
    if (!extract_interior_edges(N, E, es, P->ie)) { snprintf(err, errlen, "bad stencil"); delete P; return nullptr; }
    P->Ei = (int32_t)(P->ie.size() / 4);
    build_pattern(N, F, P->fn.data(), P->Ei, P->ie.data(), P->pat);
    tiles::RowLayout rows, *prows = nullptr;
    if (eol_index) for (int32_t a = 0; a < N; ++a) P->has_eol |= eol_index[a] >= 0;
    if (P->has_eol) {
        if (!eol::build(N, F, P->fn.data(), P->Ei, P->ie.data(), eol_index, P->pat.blkptrM, P->pat.nbrM, P->pat.blkptrK, P->pat.nbrK, P->ep)) {
            snprintf(err, errlen, "%s", P->ep.error.c_str());
            delete P;
            return nullptr;
        }
        rows.dstM = P->ep.dstM.data(); rows.dstK = P->ep.dstK.data(); rows.extraM = P->ep.extraM.data(); rows.extraK = P->ep.extraK.data();
        prows = &rows;
    }
    if (!tiles::build(N, F, P->fn.data(), P->Ei, P->ie.data(), P->pat, X_hint, dedup != 0, P->tp, prows)) {
        snprintf(err, errlen, "%s", P->tp.error.c_str());
        delete P;
        return nullptr;
    }
    return P;
}
void *hm_plan_create(int32_t N, int32_t F, const int32_t *fn, int32_t E, const int32_t *es, const double *X_hint, int dedup, char *err, int errlen) {
    return hm_plan_create_eol(N, F, fn, E, es, X_hint, dedup, nullptr, err, errlen);
}
void hm_plan_destroy(void *p) { delete (HmPlan *)p; }
// info[17]: nnzM, nnzK, n_tiles, n_templates, elem_evals, geo bytes, template bytes, max scratch doubles, max loc, Ei, runs, groups, pull rows, max staging
void hm_plan_info(void *p, int64_t *info) {
    HmPlan *P = (HmPlan *)p;
    info[0] = 9 * P->pat.nblkM; info[1] = 9 * P->pat.nblkK; info[2] = P->tp.n_tiles; info[3] = P->tp.n_templates; info[4] = P->tp.elem_evals;
    info[5] = (int64_t)P->tp.geo.size() * 4; info[6] = (int64_t)P->tp.tmpl.size() * 4; info[7] = P->tp.max_scratch; info[8] = P->tp.max_loc;
    if (P->has_eol) { info[0] = P->ep.nnzM; info[1] = P->ep.nnzK; }
    info[15] = P->has_eol ? P->ep.dof : 3 * P->N;
    info[16] = tiles::mostly_full_tiles(P->tp) ? 1 : 0;
    info[9] = P->Ei; info[10] = P->tp.n_runs; info[11] = P->tp.n_groups; info[12] = P->tp.pull_rows; info[13] = P->tp.max_kstage; info[14] = P->tp.max_mstage;
}
void hm_plan_pattern(void *p, int which, int32_t *outer, int32_t *inner) {
    HmPlan *P = (HmPlan *)p;
    if (P->has_eol) {
        const std::vector<int32_t> &eo = which ? P->ep.outerK : P->ep.outerM, &ei = which ? P->ep.innerK : P->ep.innerM;
        memcpy(outer, eo.data(), eo.size() * 4);
        if (!ei.empty()) memcpy(inner, ei.data(), ei.size() * 4);
        return;
    }
    std::vector<int32_t> o, in;
    build_eigen_arrays(P->N, which ? P->pat.blkptrK : P->pat.blkptrM, which ? P->pat.nbrK : P->pat.nbrM, o, in);
    memcpy(outer, o.data(), o.size() * 4);
    if (!in.empty()) memcpy(inner, in.data(), in.size() * 4);
}
// Bank-conflict model of the phase-2 block pulls (developer statistic): for every O group, list row and half (s0 / s1) the
// number of shared-memory wavefronts one 128-bit load instruction takes = sum over quarter warps of the largest number of
// DISTINCT 16-byte addresses falling into one of the 8 bank groups; ideal = number of quarter warps with an active lane.
// out[0] = modelled wavefronts, out[1] = ideal, both summed over all tiles (per load instruction of a block, i.e. x4.5 in reality).
void hm_plan_pull_conflicts(void *p, double *out) {
    HmPlan *P = (HmPlan *)p;
    out[0] = out[1] = 0.0;
    for (int32_t t = 0; t < P->tp.n_tiles; ++t) {
        const uint32_t *geo = P->tp.geo.data() + (size_t)t * P->tp.max_geo16 * 4;
        const uint32_t *B = P->tp.tmpl.data() + (size_t)geo[0] * 4 + (size_t)(geo[2] & 0xffffu) * 4;
        const int nOwn = B[0] & 255, nG = (B[0] >> 8) & 255, n4 = (nOwn + 3) & ~3;
        const uint32_t *grp = B + tiles::B_HDR + 3 * n4;
        const uint32_t *pulls = grp + 4 * nG + 2 * tiles::GROUP * nG;
        for (int g = 0; g < nG; ++g) {
            const int kind = grp[4 * g] & 255, rows = ((grp[4 * g] >> 8) & 255) + ((grp[4 * g] >> 16) & 255);
            if (kind != tiles::KIND_O) continue;
            const uint32_t *pl = pulls + grp[4 * g + 1];
            for (int r = 0; r < rows; ++r)
                for (int half = 0; half < 2; ++half)
                    for (int q = 0; q < 4; ++q) {
                        int addr[8], na = 0, cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int l = 0; l < 8; ++l) {
                            const uint32_t e = pl[r * 32 + 8 * q + l];
                            const int off = half ? (int)(e >> 16) : (int)(e & 0xffffu);
                            bool dup = false;
                            for (int k = 0; k < na; ++k) dup |= addr[k] == off;
                            if (!dup) { addr[na++] = off; ++cnt[(off >> 1) & 7]; }
                        }
                        int mx = 0;
                        for (int k = 0; k < 8; ++k) mx = std::max(mx, cnt[k]);
                        out[0] += mx; out[1] += 1.0;
                    }
        }
    }
}
// FNV-1a over the plan's geometry blobs and templates: equal for equal plans (used to check that the build does not depend on the worker count)
uint64_t hm_plan_hash(void *p) {
    HmPlan *P = (HmPlan *)p;
    uint64_t h = 1469598103934665603ull;
    for (uint32_t w : P->tp.geo) { h ^= w; h *= 1099511628211ull; }
    for (uint32_t w : P->tp.tmpl) { h ^= w; h *= 1099511628211ull; }
    h ^= (uint64_t)P->tp.n_tiles; h *= 1099511628211ull;
    h ^= (uint64_t)P->tp.n_templates; h *= 1099511628211ull;
    return h;
}
// developer helper: copies template part B of tile t (u32 words) into out; returns the number of words
int hm_plan_dump(void *p, int t, uint32_t *out, int cap) {
    HmPlan *P = (HmPlan *)p;
    const uint32_t *geo = P->tp.geo.data() + (size_t)t * P->tp.max_geo16 * 4;
    const uint32_t *B = P->tp.tmpl.data() + (size_t)geo[0] * 4 + (size_t)(geo[2] & 0xffffu) * 4;
    const int n = (int)(geo[2] >> 16) * 4;
    if (n > cap) return -1;
    memcpy(out, B, (size_t)n * 4);
    return n;
}
struct HostBulk {   // host stand-in of the device's bulk copy: same alignment contract
    bool *ok;
    void operator()(double *dst, const double *src, uint32_t bytes) const {
        if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15)) *ok = false;
        memcpy(dst, src, bytes);
    }
};
// returns 0, or -1 if a bulk copy was issued with a misaligned address / size
int hm_plan_fill_ex(void *p, const double *x, const double *X, const double *mat6, const double *grav, double h, double *f, double *Mv, double *Kv, int skip_m);
int hm_plan_fill(void *p, const double *x, const double *X, const double *mat6, const double *grav, double h, double *f, double *Mv, double *Kv) {
    return hm_plan_fill_ex(p, x, X, mat6, grav, h, f, Mv, Kv, 0);
}
// skip_m: EOLC_FILL_M_UNCHANGED as the kernel runs it (off-diagonal mass groups skipped, M rows not copied out)
int hm_plan_fill_ex(void *p, const double *x, const double *X, const double *mat6, const double *grav, double h, double *f, double *Mv, double *Kv, int skip_m) {
    HmPlan *P = (HmPlan *)p;
    tiles::FillParams prm;
    prm.mu = membrane_mu(mat6[1], mat6[2]); prm.lam = membrane_lambda(mat6[1], mat6[2]); prm.rho = mat6[0]; prm.beta = mat6[3];
    prm.gx = grav[0]; prm.gy = grav[1]; prm.gz = grav[2]; prm.dhh = mat6[5] * h * h;
    std::vector<double> scr(P->tp.max_scratch + 16, 0.0), xs(3 * (size_t)P->tp.max_loc + 4), Xs(2 * (size_t)P->tp.max_loc + 4);
    // staging arrays: 16-byte aligned, shifted by the phase of the output pointers like the kernel does; persistent across tiles
    std::vector<double> kbuf(P->tp.max_kstage + 6, 1e300), mbuf(P->tp.max_mstage + 6, 1e300), fbuf(tiles::MAX_FSTAGE + 6, 1e300);
    auto align16 = [](double *q) { return (reinterpret_cast<uintptr_t>(q) & 15) ? q + 1 : q; };
    auto phase = [](const double *q) { return (int)((reinterpret_cast<uintptr_t>(q) >> 3) & 1); };
    bool ok = true;
    uint32_t tM_id = 0xffffffffu;
    for (int32_t t = 0; t < P->tp.n_tiles; ++t) {
        const uint32_t *geo = P->tp.geo.data() + (size_t)t * P->tp.max_geo16 * 4;
        const int nLoc = (geo[1] >> 8) & 255;
        const uint32_t *loc = geo + 4;
        for (int l = 0; l < nLoc; ++l) {
            for (int k = 0; k < 3; ++k) xs[3 * l + k] = x[3 * (size_t)loc[l] + k];
            for (int k = 0; k < 2; ++k) Xs[2 * l + k] = X[2 * (size_t)loc[l] + k];
        }
        for (auto &v : scr) v = 1e300;   // poison: a pull from an unparked slot shows up immediately
        for (int z = 0; z < tiles::ZPAD; ++z) scr[z] = 0.0;   // the zero block the padded pull entries point at
        tiles::TileView V;
        V.geo = geo; V.tmpl = P->tp.tmpl.data() + (size_t)geo[0] * 4; V.tmplB = V.tmpl + (size_t)(geo[2] & 0xffffu) * 4;
        V.xs = xs.data(); V.Xs = Xs.data(); V.scr = scr.data();
        V.kst = align16(kbuf.data()) + phase(Kv); V.mst = align16(mbuf.data()) + phase(Mv); V.fst = align16(fbuf.data()) + phase(f);
        for (auto &v : kbuf) v = 1e300;
        for (auto &v : fbuf) v = 1e300;
        const bool m_full = geo[0] != tM_id;   // like the kernel: the M staging keeps the explicit zeros of the resident template
        tM_id = geo[0];
        if (m_full) for (auto &v : mbuf) v = 1e300;
        for (int tid = 0; tid < tiles::NTHREADS; ++tid) tiles::phase1(tid, tiles::NTHREADS, V, prm);
        for (int tid = 0; tid < tiles::P2THREADS; ++tid) tiles::phase2(tid, tiles::P2THREADS, V, m_full, skip_m != 0);
        for (int tid = 0; tid < tiles::P2THREADS; ++tid) tiles::phase3(tid, tiles::P2THREADS, V);
        for (int tid = 0; tid < tiles::P2THREADS; ++tid) tiles::copy_out_runs(tid, tiles::P2THREADS, V, f, Mv, Kv, skip_m ? 5u : 7u, HostBulk{&ok});
    }
    if (P->has_eol) {   // the two EOL kernels of forces.cu, one "thread" after the other
        const eol::Plan &ep = P->ep;
        eol::Params ep_prm{mat6[1], mat6[2], mat6[0], mat6[3], grav[0], grav[1], grav[2], mat6[5] * h * h};
        std::vector<double> scratch((size_t)ep.scratch_doubles, 1e300);
        for (int32_t i = 0; i < ep.n_faces(); ++i) eol::face_record(ep.faces.data() + 4 * (size_t)i, x, X, ep_prm, scratch.data() + (size_t)i * eol::FACE_REC);
        for (int32_t i = 0; i < ep.n_edges(); ++i)
            eol::edge_record(ep.edges.data() + 8 * (size_t)i, x, X, ep_prm, scratch.data() + (size_t)ep.n_faces() * eol::FACE_REC + (size_t)i * eol::EDGE_REC);
        for (const eol::Target &t : ep.targets) eol::gather_target(t, ep.sources.data(), scratch.data(), 3u * (uint32_t)P->N, f, Mv, Kv);
    }
    return ok ? 0 : -1;
}
}
