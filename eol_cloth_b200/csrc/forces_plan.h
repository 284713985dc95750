// forces_plan.h — host-side topology plan of Forces::fill: the fixed CSR pattern of M / MDK and the "tiles" schedule.
// Pure C++ (no CUDA), so the SAME code builds the plan for the GPU kernel (forces.cu) and for the host emulation of the
// kernel used by the CPU tests (tests/hostmath/hostmath.cpp).
//
// Pattern (== what Eigen's setFromTriplets produces from the reference's triplets, /root/reference/src/Forces.cpp:103-125,
// 522-539, 928-929): node a has one 3x3 block per neighbour b in nbr(a) (ascending), nbrM(a) = {a} + nodes sharing a face,
// nbrK(a) = nbrM(a) + nodes sharing a bending stencil (Forces.cpp:692-697).  Column-major == row-major (symmetric pattern);
// the values of node a's three rows are stored at 9*blkptr[a] + j*3*deg(a) + 3*p + k.
//
// Tiles: the nodes are partitioned into spatially compact tiles of <= MAX_OWN nodes (recursive coordinate bisection of the
// material coordinates).  A tile evaluates every face / bending stencil that touches one of its nodes ONCE, parks the
// element blocks in shared memory and then every output block of its nodes pulls its contributions in a fixed order
// (faces ascending, then stencils ascending; not-transposed before transposed) — deterministic, no atomics.
// Tiles with identical local structure (all interior tiles of a regular sheet) share one template.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace eolc {

struct Pattern {
    int32_t N = 0;
    std::vector<int64_t> blkptrM, blkptrK;   // N+1
    std::vector<int32_t> nbrM, nbrK;         // neighbour (column) node per block, ascending within a node
    int64_t nblkM = 0, nblkK = 0;
};

inline int64_t find_block(const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr, int32_t a, int32_t b) {
    auto beg = nbr.begin() + blkptr[a], end = nbr.begin() + blkptr[a + 1];
    return std::lower_bound(beg, end, b) - nbr.begin();
}

// fn: 3F face nodes; ie: 4Ei interior-edge stencils
inline void build_pattern(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie, Pattern &P) {
    P.N = N;
    std::vector<int64_t> cntM(N + 1, 0), cntK(N + 1, 0);
    std::vector<char> used(N, 0);
    for (int64_t i = 0; i < 3 * (int64_t)F; ++i) { used[fn[i]] = 1; cntM[fn[i] + 1] += 2; cntK[fn[i] + 1] += 2; }
    for (int64_t i = 0; i < 4 * (int64_t)Ei; ++i) cntK[ie[i] + 1] += 3;
    for (int32_t a = 0; a < N; ++a) {
        if (used[a]) { cntM[a + 1] += 1; cntK[a + 1] += 1; } else { cntM[a + 1] = 0; cntK[a + 1] = 0; }   // isolated nodes own no block
    }
    for (int32_t a = 0; a < N; ++a) { cntM[a + 1] += cntM[a]; cntK[a + 1] += cntK[a]; }
    std::vector<int32_t> rawM(cntM[N]), rawK(cntK[N]);
    std::vector<int64_t> pM(cntM.begin(), cntM.end() - 1), pK(cntK.begin(), cntK.end() - 1);
    for (int32_t a = 0; a < N; ++a) if (used[a]) { rawM[pM[a]++] = a; rawK[pK[a]++] = a; }
    for (int32_t i = 0; i < F; ++i)
        for (int v = 0; v < 3; ++v) {
            int32_t a = fn[3 * (size_t)i + v];
            for (int w = 0; w < 3; ++w) if (w != v) { rawM[pM[a]++] = fn[3 * (size_t)i + w]; rawK[pK[a]++] = fn[3 * (size_t)i + w]; }
        }
    for (int32_t i = 0; i < Ei; ++i)
        for (int v = 0; v < 4; ++v) {
            int32_t a = ie[4 * (size_t)i + v];
            if (!used[a]) continue;   // cannot happen for a stencil built from faces; keeps the arrays consistent
            for (int w = 0; w < 4; ++w) if (w != v) rawK[pK[a]++] = ie[4 * (size_t)i + w];
        }
    auto compress = [&](std::vector<int64_t> &cnt, std::vector<int64_t> &fill, std::vector<int32_t> &raw, std::vector<int64_t> &blkptr,
                        std::vector<int32_t> &nbr) {
        blkptr.assign(N + 1, 0);
        nbr.clear();
        nbr.reserve(raw.size() / 2);
        for (int32_t a = 0; a < N; ++a) {
            auto b = raw.begin() + cnt[a], e = raw.begin() + fill[a];
            std::sort(b, e);
            auto u = std::unique(b, e);
            nbr.insert(nbr.end(), b, u);
            blkptr[a + 1] = (int64_t)nbr.size();
        }
    };
    compress(cntM, pM, rawM, P.blkptrM, P.nbrM);
    compress(cntK, pK, rawK, P.blkptrK, P.nbrK);
    P.nblkM = P.blkptrM[N];
    P.nblkK = P.blkptrK[N];
}

// interior bending stencils (both adjacent faces present, Forces.cpp:688-690) of a 4E stencil array; false on a bad index
inline bool extract_interior_edges(int32_t N, int32_t E, const int32_t *edge_stencil, std::vector<int32_t> &ie) {
    ie.clear();
    for (int32_t e = 0; e < E; ++e) {
        const int32_t *s = edge_stencil + 4 * (size_t)e;
        if (s[2] < 0 || s[3] < 0) continue;
        for (int v = 0; v < 4; ++v) if (s[v] < 0 || s[v] >= N) return false;
        ie.insert(ie.end(), s, s + 4);
    }
    return true;
}

// Eigen-style outer/inner arrays of the block pattern (outerIndexPtr / innerIndexPtr of the reference's matrices)
inline void build_eigen_arrays(int32_t N, const std::vector<int64_t> &blkptr, const std::vector<int32_t> &nbr, std::vector<int32_t> &outer,
                               std::vector<int32_t> &inner) {
    outer.assign(3 * (size_t)N + 1, 0);
    inner.resize(9 * (size_t)blkptr[N]);
    for (int32_t a = 0; a < N; ++a) {
        int64_t b0 = blkptr[a];
        int deg = (int)(blkptr[a + 1] - b0);
        for (int j = 0; j < 3; ++j) {
            int64_t rs = 9 * b0 + (int64_t)j * 3 * deg;
            outer[3 * (size_t)a + j] = (int32_t)rs;
            for (int p = 0; p < deg; ++p)
                for (int k = 0; k < 3; ++k) inner[rs + 3 * p + k] = 3 * nbr[b0 + p] + k;
        }
    }
    outer[3 * (size_t)N] = (int32_t)(9 * blkptr[N]);
}

// ---------------------------------------------------------------------------------------------------------------------
// tiles
// ---------------------------------------------------------------------------------------------------------------------
namespace tiles {

#ifndef EOLC_TILE_OWN
#define EOLC_TILE_OWN 32
#endif
#ifndef EOLC_TILE_CTAS
#define EOLC_TILE_CTAS 1
#endif
#ifndef EOLC_TILE_THREADS
#define EOLC_TILE_THREADS 256
#endif
constexpr int NTHREADS = 256;        // element slots per tile (phase 1 uses the first NTHREADS threads)
constexpr int CTA_THREADS = EOLC_TILE_THREADS;   // threads per CTA (>= NTHREADS); all of them work in phase 2 and in the prefetch
constexpr int CTAS_PER_SM = EOLC_TILE_CTAS;
constexpr int MAX_OWN = EOLC_TILE_OWN;   // nodes owned by a tile (<= 32: 5-bit fields)
constexpr int MAX_LOC = 128;         // distinct nodes referenced by a tile's elements (8-bit local ids)
constexpr int EDGE_STRIDE = 86;      // doubles parked per bending stencil: 4 diagonal blocks x 6, 6 off-diagonal x 10 (+2: bank spread)
constexpr int FACE_STRIDE = 62;      // doubles parked per face: 3 x (diag 6 + force 3 + pad) + 3 x (off 9 + pad), t8 in the first pad
constexpr int FACE_T8 = 39;          // offset of t8 (rho * 2A) inside a face slot
constexpr int ZPAD = 16;             // doubles at the start of the scratch that stay zero: pull lists are padded to an even length
                                     // with offset 0, so the pull loops take two contributions per trip without a tail test
constexpr int MAX_SCRATCH_DOUBLES = CTAS_PER_SM == 1 ? 20480 : 12288;   // 160 KB (96 KB with two CTAs per SM) of parked blocks per tile
constexpr int MAX_KSTAGE = 4096;     // doubles of MDK rows one tile stages in shared memory before the coalesced copy-out (32 KB)
constexpr int MAX_MSTAGE = 2304;     // doubles of M rows one tile stages (expanded blocks: m on the block diagonal, explicit zeros off it)
constexpr int COPY_CHUNK = 256;      // doubles per copy-out chunk of staged MDK rows
constexpr int MAX_COUNT = 63;        // PAIRS of contributions per loop of one output block (6-bit fields)

// smem offsets (in doubles) of the parked blocks inside an element slot
inline int edge_diag_off(int i) { return 6 * i; }
inline int edge_off_off(int lo, int hi) { static const int k[4][4] = {{-1, 0, 1, 2}, {0, -1, 3, 4}, {1, 3, -1, 5}, {2, 4, 5, -1}}; return 24 + 10 * k[lo][hi]; }
inline int face_diag_off(int v) { return 10 * v; }
inline int face_off_off(int lo, int hi) { static const int k[3][3] = {{-1, 0, 1}, {0, -1, 2}, {1, 2, -1}}; return 30 + 10 * k[lo][hi]; }

// 64-bit record of one phase-2 work item:
//   q0:16 first pull entry (even) | c0:6 | c1:6 | p:8 block position in the node's row | own:5 | p2:8 | own2:5 | has2:1 | flag:1
//   c0 / c1 count PAIRS of pull entries (lists are padded to an even length with offset 0 = the zero block).
//   D (diagonal MDK block + f): c0 = faces, c1 = stencils.
//   O (off-diagonal MDK block (own, p)): c0 = contributions parked in this orientation, c1 = parked transposed.  If the
//     column node is owned by the same tile (has2), the item also writes the mirrored block (own2, p2) = transpose, and the
//     mirrored pair has no item of its own.
//   M (mass block): c0 = faces, flag = diagonal block, has2 as for O.
inline uint64_t pack_rec(unsigned q0, unsigned c0, unsigned c1, unsigned p, unsigned own, unsigned p2, unsigned own2, unsigned has2, unsigned flag) {
    return (uint64_t)q0 | ((uint64_t)c0 << 16) | ((uint64_t)c1 << 22) | ((uint64_t)p << 28) | ((uint64_t)own << 36) | ((uint64_t)p2 << 41) |
           ((uint64_t)own2 << 49) | ((uint64_t)has2 << 54) | ((uint64_t)flag << 55);
}

// Template header (4 x u32) + items + row sizes (degK | degM << 8 per owned node) + staging offsets (MDK rows | mass scalars
// << 16 per owned node) + records + pull entries, every part padded to 16 bytes:
//   w0 = nE | nEpad << 8 | nF << 16     (stencil slots [0, nE), face slots [nEpad, nEpad + nF))
//   w1 = nD | nDpad << 16               (D records [0, nD), padded to a warp)
//   w2 = nO | nOpad << 16               (O records [nDpad, nDpad + nO))
//   w3 = nM | npull16 << 16             (M records [nDpad + nOpad, ... + nM); npull16 = pull entries / 8, rounded up)
// Geometry header (4 x u32) + (kbase, mbase) int64 pairs per owned node + local->global node table:
//   w0 = template offset (16-byte units)   w1 = nOwn | nLoc << 8   w2 = size of part A | size of part B << 16 (16-byte units)
//   w3 = number of copy-out chunks;   ... + the copy-out chunks (see below), offsets and lengths in doubles
struct Plan {
    int32_t n_tiles = 0, n_templates = 0;
    std::vector<uint32_t> geo;         // geometry blobs (u32 words), tile t at t * 4 * max_geo16 (fixed stride: no offset lookup)
    std::vector<uint32_t> tmpl;        // template blobs (u32 words)
    uint32_t max_geo16 = 0, max_tmplA16 = 0, max_tmplB16 = 0, max_loc = 0, max_scratch = 0, max_kstage = 0, max_mstage = 0;   // per-tile maxima
    int64_t elem_evals = 0;            // element evaluations per fill (>= F + Ei because of halo re-evaluation)
    std::string error;
};

struct Builder {
    int32_t N, F, Ei;
    const int32_t *fn, *ie;
    const Pattern &pat;
    std::vector<int32_t> nfp, nep;     // node -> incident faces / stencils (CSR), ascending element index, elem << 2 | pos
    std::vector<uint32_t> nfl, nel;
    Builder(int32_t N_, int32_t F_, const int32_t *fn_, int32_t Ei_, const int32_t *ie_, const Pattern &p) : N(N_), F(F_), Ei(Ei_), fn(fn_), ie(ie_), pat(p) {
        nfp.assign(N + 1, 0); nep.assign(N + 1, 0);
        for (int64_t i = 0; i < 3 * (int64_t)F; ++i) nfp[fn[i] + 1]++;
        for (int64_t i = 0; i < 4 * (int64_t)Ei; ++i) nep[ie[i] + 1]++;
        for (int32_t a = 0; a < N; ++a) { nfp[a + 1] += nfp[a]; nep[a + 1] += nep[a]; }
        nfl.resize(nfp[N]); nel.resize(nep[N]);
        std::vector<int32_t> pf(nfp.begin(), nfp.end() - 1), pe(nep.begin(), nep.end() - 1);
        for (int32_t i = 0; i < F; ++i) for (int v = 0; v < 3; ++v) nfl[pf[fn[3 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
        for (int32_t i = 0; i < Ei; ++i) for (int v = 0; v < 4; ++v) nel[pe[ie[4 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
        fstamp.assign(F, -1); estamp.assign(Ei, -1); lstamp.assign(N, -1); local.assign(N, 0);
        fslot.assign(F, 0); eslot.assign(Ei, 0);
    }
    // scratch for tile construction
    std::vector<int32_t> fstamp, estamp, lstamp, local, fslot, eslot;   // *slot: element -> slot of the tile being built
    int32_t stamp = 0;

    // element sets of a candidate tile; returns false if the tile exceeds the kernel's capacities
    bool fits(const int32_t *own, int n_own, std::vector<int32_t> &faces, std::vector<int32_t> &edges) {
        ++stamp;
        faces.clear(); edges.clear();
        for (int o = 0; o < n_own; ++o) {
            const int32_t a = own[o];
            for (int32_t k = nfp[a]; k < nfp[a + 1]; ++k) { int32_t f = nfl[k] >> 2; if (fstamp[f] != stamp) { fstamp[f] = stamp; faces.push_back(f); } }
            for (int32_t k = nep[a]; k < nep[a + 1]; ++k) { int32_t e = nel[k] >> 2; if (estamp[e] != stamp) { estamp[e] = stamp; edges.push_back(e); } }
        }
        const int nE = (int)edges.size(), nF = (int)faces.size();
        const int nEpad = (nE + 31) / 32 * 32;
        if (n_own > MAX_OWN || nEpad + nF > NTHREADS || nEpad > 255) return false;
        if (ZPAD + (int64_t)nEpad * EDGE_STRIDE + (int64_t)nF * FACE_STRIDE > MAX_SCRATCH_DOUBLES) return false;
        int64_t kst = 0, mst = 0;
        for (int o = 0; o < n_own; ++o) { kst += 9 * (pat.blkptrK[own[o] + 1] - pat.blkptrK[own[o]]); mst += 9 * (pat.blkptrM[own[o] + 1] - pat.blkptrM[own[o]]); }
        if (kst > MAX_KSTAGE || mst > MAX_MSTAGE) return false;
        int nloc = 0;
        auto touch = [&](int32_t g) { if (lstamp[g] != stamp) { lstamp[g] = stamp; ++nloc; } };
        for (int o = 0; o < n_own; ++o) touch(own[o]);
        for (int32_t f : faces) for (int j = 0; j < 3; ++j) touch(fn[3 * (size_t)f + j]);
        for (int32_t e : edges) for (int j = 0; j < 4; ++j) touch(ie[4 * (size_t)e + j]);
        return nloc <= MAX_LOC;
    }
};

// Recursive coordinate bisection of `idx[lo, hi)` into `leaves` parts of near-equal size; deterministic (ties by index).
inline void rcb(std::vector<int32_t> &idx, size_t lo, size_t hi, size_t leaves, const double *cx, const double *cy,
                std::vector<std::pair<size_t, size_t>> &out) {
    if (leaves <= 1 || hi - lo <= 1) { out.push_back({lo, hi}); return; }
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (size_t i = lo; i < hi; ++i) {
        x0 = std::min(x0, cx[idx[i]]); x1 = std::max(x1, cx[idx[i]]);
        y0 = std::min(y0, cy[idx[i]]); y1 = std::max(y1, cy[idx[i]]);
    }
    const bool ax = (x1 - x0) >= (y1 - y0);
    const double *c0 = ax ? cx : cy, *c1 = ax ? cy : cx;
    const size_t lleaves = leaves / 2;
    const size_t mid = lo + (size_t)(((hi - lo) * (uint64_t)lleaves + leaves / 2) / leaves);
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int32_t a, int32_t b) {
        if (c0[a] != c0[b]) return c0[a] < c0[b];
        if (c1[a] != c1[b]) return c1[a] < c1[b];
        return a < b;
    });
    rcb(idx, lo, mid, lleaves, cx, cy, out);
    rcb(idx, mid, hi, leaves - lleaves, cx, cy, out);
}

inline bool build(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie, const Pattern &pat, const double *X_hint, bool dedup,
                  Plan &P) {
    P = Plan();
    if (N == 0) return true;
    Builder B(N, F, fn, Ei, ie, pat);
    for (int32_t e = 0; e < Ei; ++e) {
        const int32_t *s = ie + 4 * (size_t)e;
        if (s[0] == s[1] || s[0] == s[2] || s[0] == s[3] || s[1] == s[2] || s[1] == s[3] || s[2] == s[3]) {
            P.error = "bending stencil " + std::to_string(e) + " repeats a node (non-manifold edge)";
            return false;
        }
    }
    // ---- coordinates for the bisection: material coordinates if given, else a breadth-first numbering
    std::vector<double> cx(N), cy(N, 0.0);
    if (X_hint) {
        for (int32_t a = 0; a < N; ++a) { cx[a] = X_hint[2 * (size_t)a]; cy[a] = X_hint[2 * (size_t)a + 1]; }
    } else {
        std::vector<int32_t> order(N, -1), queue;
        queue.reserve(N);
        int32_t next = 0;
        for (int32_t seed = 0; seed < N; ++seed) {
            if (order[seed] >= 0) continue;
            order[seed] = next++; queue.push_back(seed);
            for (size_t h = queue.size() - 1; h < queue.size(); ++h) {
                const int32_t a = queue[h];
                for (int64_t b = pat.blkptrM[a]; b < pat.blkptrM[a + 1]; ++b) {
                    const int32_t c = pat.nbrM[b];
                    if (order[c] < 0) { order[c] = next++; queue.push_back(c); }
                }
            }
        }
        for (int32_t a = 0; a < N; ++a) cx[a] = order[a];
    }
    std::vector<int32_t> idx(N);
    for (int32_t a = 0; a < N; ++a) idx[a] = a;
    std::vector<std::pair<size_t, size_t>> leaves;
    rcb(idx, 0, (size_t)N, ((size_t)N + MAX_OWN - 1) / MAX_OWN, cx.data(), cy.data(), leaves);
    // ---- split leaves that exceed the kernel's capacities
    std::vector<int32_t> faces, edges;
    {
        std::vector<std::pair<size_t, size_t>> ok;
        std::vector<std::pair<size_t, size_t>> work(leaves.rbegin(), leaves.rend());
        while (!work.empty()) {
            auto r = work.back(); work.pop_back();
            if (r.second == r.first) continue;
            if (B.fits(idx.data() + r.first, (int)std::min<size_t>(r.second - r.first, MAX_OWN + 1), faces, edges) && r.second - r.first <= (size_t)MAX_OWN) {
                ok.push_back(r);
                continue;
            }
            if (r.second - r.first == 1) {
                P.error = "node " + std::to_string(idx[r.first]) + " has too many incident faces / bending stencils for one tile";
                return false;
            }
            std::vector<std::pair<size_t, size_t>> two;
            rcb(idx, r.first, r.second, 2, cx.data(), cy.data(), two);
            work.push_back(two[1]); work.push_back(two[0]);
        }
        leaves.swap(ok);
    }
    P.n_tiles = (int32_t)leaves.size();
    std::vector<uint32_t> geo_off;     // variable-size blobs first, re-laid with a fixed stride at the end
    geo_off.reserve(leaves.size() + 1);
    std::unordered_map<uint64_t, std::vector<uint32_t>> seen;   // hash -> template offsets (16-byte units)
    std::vector<uint32_t> T;                                   // template under construction
    std::vector<uint64_t> recD, recO, recM;
    std::vector<uint16_t> pull;
    struct OTmp { int cnt; uint64_t rec; };
    std::vector<OTmp> otmp, mtmp;
    std::vector<uint16_t> listN, listT, listF;
    for (size_t t = 0; t < leaves.size(); ++t) {
        int32_t *own = idx.data() + leaves[t].first;
        const int n_own = (int)(leaves[t].second - leaves[t].first);
        std::sort(own, own + n_own);
        B.fits(own, n_own, faces, edges);
        std::sort(faces.begin(), faces.end());
        std::sort(edges.begin(), edges.end());
        const int nE = (int)edges.size(), nF = (int)faces.size(), nEpad = (nE + 31) / 32 * 32;
        const int fbase = ZPAD + nEpad * EDGE_STRIDE;
        P.elem_evals += nE + nF;
        P.max_scratch = std::max<uint32_t>(P.max_scratch, (uint32_t)(fbase + nF * FACE_STRIDE));
        // local node table: owned nodes first (ascending), then halo nodes by first appearance; element -> slot maps
        ++B.stamp;
        std::vector<uint32_t> loc;
        auto lid = [&](int32_t g) {
            if (B.lstamp[g] != B.stamp) { B.lstamp[g] = B.stamp; B.local[g] = (int32_t)loc.size(); loc.push_back((uint32_t)g); }
            return (uint32_t)B.local[g];
        };
        for (int o = 0; o < n_own; ++o) lid(own[o]);
        std::vector<uint32_t> items((size_t)nEpad + nF, 0u);
        for (int s = 0; s < nE; ++s) {
            const int32_t *v = ie + 4 * (size_t)edges[s];
            items[s] = lid(v[0]) | (lid(v[1]) << 8) | (lid(v[2]) << 16) | (lid(v[3]) << 24);
            B.eslot[edges[s]] = s;
        }
        for (int s = 0; s < nF; ++s) {
            const int32_t *v = fn + 3 * (size_t)faces[s];
            items[nEpad + s] = lid(v[0]) | (lid(v[1]) << 8) | (lid(v[2]) << 16);
            B.fslot[faces[s]] = s;
        }
        P.max_loc = std::max<uint32_t>(P.max_loc, (uint32_t)loc.size());
        // ---- phase-2 work items
        recD.clear(); otmp.clear(); mtmp.clear(); pull.clear();
        std::vector<uint32_t> degs((size_t)n_own, 0u), offs((size_t)n_own, 0u);   // row sizes / staging offsets per owned node
        uint32_t kstage = 0, mstage = 0;
        bool overflow = false;
        auto pad2 = [](std::vector<uint16_t> &l) { if (l.size() & 1) l.push_back(0); };
        auto owned_index = [&](int32_t g) { return (B.lstamp[g] == B.stamp && B.local[g] < n_own) ? B.local[g] : -1; };
        for (int o = 0; o < n_own; ++o) {
            const int32_t a = own[o];
            const int64_t b0 = pat.blkptrK[a], m0 = pat.blkptrM[a];
            const int deg = (int)(pat.blkptrK[a + 1] - b0), degM = (int)(pat.blkptrM[a + 1] - m0);
            if (deg > 255) { P.error = "node " + std::to_string(a) + " has more than 255 neighbours"; return false; }
            degs[o] = (uint32_t)deg | ((uint32_t)degM << 8);
            offs[o] = kstage | (mstage << 16);
            kstage += 9u * (uint32_t)deg; mstage += 9u * (uint32_t)degM;
            for (int p = 0; p < std::max(deg, 1); ++p) {
                const int32_t b = deg ? pat.nbrK[b0 + p] : a;
                const int ob = b == a ? -1 : owned_index(b);
                if (ob >= 0 && b < a) continue;      // the pair is assembled by the item of (b, a), which mirrors it into this row
                listN.clear(); listT.clear(); listF.clear();
                std::vector<uint16_t> listE;         // diagonal block: stencil entries, kept apart from the face entries
                // faces ascending, then stencils ascending
                for (int32_t k = B.nfp[a]; k < B.nfp[a + 1]; ++k) {
                    const int32_t f = B.nfl[k] >> 2;
                    const int va = B.nfl[k] & 3, sf = B.fslot[f];
                    const int32_t *v = fn + 3 * (size_t)f;
                    const int base = fbase + sf * FACE_STRIDE;
                    if (b == a) { listN.push_back((uint16_t)(base + face_diag_off(va))); listF.push_back((uint16_t)(base + FACE_T8)); continue; }
                    for (int vj = 0; vj < 3; ++vj)
                        if (v[vj] == b) {
                            (va < vj ? listN : listT).push_back((uint16_t)(base + face_off_off(std::min(va, vj), std::max(va, vj))));
                            listF.push_back((uint16_t)(base + FACE_T8));
                        }
                }
                for (int32_t k = B.nep[a]; k < B.nep[a + 1]; ++k) {
                    const int32_t e = B.nel[k] >> 2;
                    const int ia = B.nel[k] & 3, se = B.eslot[e];
                    const int32_t *v = ie + 4 * (size_t)e;
                    const int base = ZPAD + se * EDGE_STRIDE;
                    if (b == a) { listE.push_back((uint16_t)(base + edge_diag_off(ia))); continue; }
                    for (int ij = 0; ij < 4; ++ij)
                        if (v[ij] == b) (ia < ij ? listN : listT).push_back((uint16_t)(base + edge_off_off(std::min(ia, ij), std::max(ia, ij))));
                }
                const unsigned q0 = (unsigned)pull.size();
                unsigned p2 = 0, own2 = 0, has2 = 0;
                if (ob >= 0) { has2 = 1; own2 = (unsigned)ob; p2 = (unsigned)(find_block(pat.blkptrK, pat.nbrK, b, a) - pat.blkptrK[b]); }
                if (b == a) {
                    pad2(listN); pad2(listE);
                    if (listN.size() / 2 > (size_t)MAX_COUNT || listE.size() / 2 > (size_t)MAX_COUNT) overflow = true;
                    pull.insert(pull.end(), listN.begin(), listN.end());
                    pull.insert(pull.end(), listE.begin(), listE.end());
                    recD.push_back(pack_rec(q0, (unsigned)listN.size() / 2, (unsigned)listE.size() / 2, (unsigned)p, (unsigned)o, 0, 0, 0, 0));
                } else {
                    pad2(listN); pad2(listT);
                    if (listN.size() / 2 > (size_t)MAX_COUNT || listT.size() / 2 > (size_t)MAX_COUNT) overflow = true;
                    pull.insert(pull.end(), listN.begin(), listN.end());
                    pull.insert(pull.end(), listT.begin(), listT.end());
                    otmp.push_back({(int)(listN.size() + listT.size()),
                                    pack_rec(q0, (unsigned)listN.size() / 2, (unsigned)listT.size() / 2, (unsigned)p, (unsigned)o, p2, own2, has2, 0)});
                }
                if (!listF.empty()) {   // the pair shares a face -> mass block
                    pad2(listF);
                    const unsigned qm = (unsigned)pull.size();
                    if (listF.size() / 2 > (size_t)MAX_COUNT) overflow = true;
                    pull.insert(pull.end(), listF.begin(), listF.end());
                    const unsigned pM = (unsigned)(find_block(pat.blkptrM, pat.nbrM, a, b) - m0);
                    const unsigned pM2 = has2 ? (unsigned)(find_block(pat.blkptrM, pat.nbrM, b, a) - pat.blkptrM[b]) : 0u;
                    mtmp.push_back({(int)listF.size(), pack_rec(qm, (unsigned)listF.size() / 2, 0, pM, (unsigned)o, pM2, own2, has2, b == a ? 1u : 0u)});
                }
            }
        }
        if (getenv("EOLC_TILES_DEBUG_CONFLICT_FREE")) {
            // TIMING EXPERIMENT ONLY (wrong results): every item reads bank-conflict-free addresses, to bound what a
            // bank-aware ordering of the pull lists could gain
            auto fix = [&](uint64_t rec, int lane) {
                const unsigned q0 = (unsigned)(rec & 0xffffu), n = 2 * ((unsigned)((rec >> 16) & 63u) + (unsigned)((rec >> 22) & 63u));
                for (unsigned k = 0; k < n; ++k) if (pull[q0 + k]) pull[q0 + k] = (uint16_t)(ZPAD + 2 * (lane & 7) + 16 * (k & 15));
            };
            for (size_t k = 0; k < recD.size(); ++k) fix(recD[k], (int)k);
            for (size_t k = 0; k < otmp.size(); ++k) fix(otmp[k].rec, (int)k);
        }
        if (overflow || pull.size() > 65535) { P.error = "tile " + std::to_string(t) + ": contribution list overflow"; return false; }
        // heavier items first (stable): the lanes of a warp run similar trip counts
        std::stable_sort(otmp.begin(), otmp.end(), [](const OTmp &u, const OTmp &v) { return u.cnt > v.cnt; });
        std::stable_sort(mtmp.begin(), mtmp.end(), [](const OTmp &u, const OTmp &v) { return u.cnt > v.cnt; });
        const unsigned nD = (unsigned)recD.size(), nDpad = (nD + 31) / 32 * 32, nO = (unsigned)otmp.size(), nOpad = (nO + 31) / 32 * 32,
                       nM = (unsigned)mtmp.size();
        const unsigned npull16 = (unsigned)((pull.size() + 7) / 8);
        // ---- serialise the template: part A (phase 1: header + items), part B (phase 2: header + row sizes + staging offsets +
        //      records + pull entries); A is staged one tile ahead, B only while its own tile is being assembled
        T.clear();
        T.push_back((uint32_t)nE | ((uint32_t)nEpad << 8) | ((uint32_t)nF << 16));
        T.push_back(0); T.push_back(0); T.push_back(0);
        T.insert(T.end(), items.begin(), items.end());
        while (T.size() % 4) T.push_back(0);
        const uint32_t sizeA16 = (uint32_t)(T.size() / 4);
        T.push_back(nD | (nDpad << 16));
        T.push_back(nO | (nOpad << 16));
        T.push_back(nM | (npull16 << 16));
        T.push_back(0);
        T.insert(T.end(), degs.begin(), degs.end());
        while (T.size() % 4) T.push_back(0);
        T.insert(T.end(), offs.begin(), offs.end());
        while (T.size() % 4) T.push_back(0);
        P.max_kstage = std::max(P.max_kstage, kstage); P.max_mstage = std::max(P.max_mstage, mstage);
        auto push64 = [&](uint64_t r) { T.push_back((uint32_t)r); T.push_back((uint32_t)(r >> 32)); };
        for (unsigned k = 0; k < nDpad; ++k) push64(k < nD ? recD[k] : 0);
        for (unsigned k = 0; k < nOpad; ++k) push64(k < nO ? otmp[k].rec : 0);
        for (unsigned k = 0; k < nM; ++k) push64(mtmp[k].rec);
        while (T.size() % 4) T.push_back(0);
        pull.resize((size_t)npull16 * 8, 0);
        for (size_t k = 0; k < pull.size(); k += 2) T.push_back((uint32_t)pull[k] | ((uint32_t)pull[k + 1] << 16));
        while (T.size() % 4) T.push_back(0);
        const uint32_t sizeB16 = (uint32_t)(T.size() / 4) - sizeA16;
        if (sizeA16 > 0xffffu || sizeB16 > 0xffffu) { P.error = "tile " + std::to_string(t) + ": template too large"; return false; }
        P.max_tmplA16 = std::max(P.max_tmplA16, sizeA16); P.max_tmplB16 = std::max(P.max_tmplB16, sizeB16);
        // ---- deduplicate
        uint32_t toff = 0;
        bool found = false;
        uint64_t h = 1469598103934665603ull;
        if (dedup) {
            for (uint32_t w : T) { h ^= w; h *= 1099511628211ull; }
            auto it = seen.find(h);
            if (it != seen.end())
                for (uint32_t cand : it->second)
                    if ((size_t)cand * 4 + T.size() <= P.tmpl.size() && memcmp(P.tmpl.data() + (size_t)cand * 4, T.data(), T.size() * 4) == 0) {
                        toff = cand; found = true; break;
                    }
        }
        if (!found) {
            toff = (uint32_t)(P.tmpl.size() / 4);
            P.tmpl.insert(P.tmpl.end(), T.begin(), T.end());
            if (dedup) seen[h].push_back(toff);
            ++P.n_templates;
        }
        // ---- geometry blob
        geo_off.push_back((uint32_t)(P.geo.size() / 4));
        const size_t g0 = P.geo.size();
        P.geo.push_back(toff);
        P.geo.push_back((uint32_t)n_own | ((uint32_t)loc.size() << 8));
        P.geo.push_back(sizeA16 | (sizeB16 << 16));
        P.geo.push_back(0);
        for (int o = 0; o < n_own; ++o) {
            const int64_t kb = 9 * pat.blkptrK[own[o]], mb = 9 * pat.blkptrM[own[o]];
            P.geo.push_back((uint32_t)(uint64_t)kb); P.geo.push_back((uint32_t)((uint64_t)kb >> 32));
            P.geo.push_back((uint32_t)(uint64_t)mb); P.geo.push_back((uint32_t)((uint64_t)mb >> 32));
        }
        P.geo.insert(P.geo.end(), loc.begin(), loc.end());
        while (P.geo.size() % 4) P.geo.push_back(0);
        {
            // copy-out chunks: owned nodes with consecutive ids have adjacent rows in the global arrays as well, so the staged
            // rows form a few long runs; runs are cut into chunks of <= COPY_CHUNK doubles (one warp each).
            // chunk = u64 (destination offset | kind << 62; kind 0 = MDK values, 1 = M values, 2 = f), u32 staging offset, u32 length
            uint32_t nchunks = 0;
            for (int kind = 0; kind < 3; ++kind) {
                int o = 0;
                while (o < n_own) {
                    auto len_of = [&](int q) { return kind == 0 ? 9u * (degs[q] & 255u) : kind == 1 ? 9u * (degs[q] >> 8) : 3u; };
                    int o1 = o;
                    uint32_t len = len_of(o);
                    while (o1 + 1 < n_own && own[o1 + 1] == own[o1] + 1) { ++o1; len += len_of(o1); }
                    const uint64_t dst = kind == 0 ? (uint64_t)(9 * pat.blkptrK[own[o]]) : kind == 1 ? (uint64_t)(9 * pat.blkptrM[own[o]]) : (uint64_t)3 * (uint64_t)own[o];
                    const uint32_t src = kind == 0 ? (offs[o] & 0xffffu) : kind == 1 ? (offs[o] >> 16) : 3u * (uint32_t)o;
                    for (uint32_t c = 0; c < len; c += COPY_CHUNK) {
                        const uint64_t d = (dst + c) | ((uint64_t)kind << 62);
                        P.geo.push_back((uint32_t)d); P.geo.push_back((uint32_t)(d >> 32));
                        P.geo.push_back(src + c); P.geo.push_back(std::min<uint32_t>(COPY_CHUNK, len - c));
                        ++nchunks;
                    }
                    o = o1 + 1;
                }
            }
            P.geo[g0 + 3] = nchunks;
        }
        P.max_geo16 = std::max<uint32_t>(P.max_geo16, (uint32_t)((P.geo.size() - g0) / 4));
        if (P.geo.size() / 4 >= ((size_t)1 << 32) || P.tmpl.size() / 4 >= ((size_t)1 << 32)) { P.error = "plan too large"; return false; }
    }
    geo_off.push_back((uint32_t)(P.geo.size() / 4));
    {
        std::vector<uint32_t> fixed((size_t)P.n_tiles * P.max_geo16 * 4, 0u);
        for (int32_t t = 0; t < P.n_tiles; ++t)
            memcpy(fixed.data() + (size_t)t * P.max_geo16 * 4, P.geo.data() + (size_t)geo_off[t] * 4, (size_t)(geo_off[t + 1] - geo_off[t]) * 16);
        P.geo.swap(fixed);
    }
    return true;
}

}  // namespace tiles
}  // namespace eolc
