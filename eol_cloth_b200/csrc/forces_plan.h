// forces_plan.h — host-side topology plan of Forces::fill: the fixed CSR pattern of M / MDK and the "tiles" schedule.
// Pure C++ (no CUDA), so the SAME code builds the plan for the GPU kernel (forces.cu) and for the host emulation of the
// kernel used by the CPU tests (tests/hostmath/hostmath.cpp).
//
// Pattern (== what Eigen's setFromTriplets produces from the reference's triplets, /root/reference/src/Forces.cpp:103-125,
// 522-539, 928-929): node a has one 3x3 block per neighbour b in nbr(a) (ascending), nbrM(a) = {a} + nodes sharing a face,
// nbrK(a) = nbrM(a) + nodes sharing a bending stencil (Forces.cpp:692-697).  Column-major == row-major (symmetric pattern);
// the values of node a's three rows are stored at 9*blkptr[a] + j*3*deg(a) + 3*p + k.
//
// Tiles: the nodes are partitioned into spatially compact tiles of <= MAX_OWN nodes (snapped strips / recursive coordinate
// bisection of the material coordinates).  A tile evaluates every face / bending stencil that touches one of its nodes ONCE, parks the
// element blocks in shared memory and then every output block of its nodes pulls its contributions in a fixed order
// (faces ascending, then stencils ascending; not-transposed before transposed) — deterministic, no atomics.
// Tiles with identical local structure (all interior tiles of a regular sheet) share one template.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <chrono>
#include <cstdio>
#include <thread>
#include <unordered_map>
#include <vector>

namespace eolc {

// host threads the plan build may use: EOLC_PLAN_THREADS, else the hardware concurrency capped at 16
inline int n_workers_hw() {
    const char *ev = getenv("EOLC_PLAN_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1, ev ? atoi(ev) : (int)std::min<unsigned>(hw ? hw : 1u, 16u));
}
// body(lo, hi) over contiguous slices of [0, n) on the plan-build threads; results must not depend on the slicing
template <class Body> inline void par_for(size_t n, Body body) {
    const int nw = (int)std::min<size_t>((size_t)n_workers_hw(), std::max<size_t>(1, n / 4096));
    if (nw <= 1) { body((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int w = 0; w < nw; ++w) th.emplace_back([&, w]() { body(n * (size_t)w / (size_t)nw, n * (size_t)(w + 1) / (size_t)nw); });
    for (auto &t : th) t.join();
}

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

// node -> incident faces / stencils (CSR), ascending element index, elem << 2 | pos; read-only once built, shared by the workers
struct NodeCSR {
    std::vector<int32_t> nfp, nep;
    std::vector<uint32_t> nfl, nel;
    NodeCSR(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie) {
        // the face lists and the stencil lists are independent counting sorts: one thread each.  (Counting and filling on all the
        // plan-build threads with relaxed atomics + a sort of every node's list measured slower: 11.3 instead of 6.6-9.9 ms at 512^2.)
        auto faces = [&]() {
            nfp.assign(N + 1, 0);
            for (int64_t i = 0; i < 3 * (int64_t)F; ++i) nfp[fn[i] + 1]++;
            for (int32_t a = 0; a < N; ++a) nfp[a + 1] += nfp[a];
            nfl.resize(nfp[N]);
            std::vector<int32_t> pf(nfp.begin(), nfp.end() - 1);
            for (int32_t i = 0; i < F; ++i) for (int v = 0; v < 3; ++v) nfl[pf[fn[3 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
        };
        auto stencils = [&]() {
            nep.assign(N + 1, 0);
            for (int64_t i = 0; i < 4 * (int64_t)Ei; ++i) nep[ie[i] + 1]++;
            for (int32_t a = 0; a < N; ++a) nep[a + 1] += nep[a];
            nel.resize(nep[N]);
            std::vector<int32_t> pe(nep.begin(), nep.end() - 1);
            for (int32_t i = 0; i < Ei; ++i) for (int v = 0; v < 4; ++v) nel[pe[ie[4 * (size_t)i + v]]++] = ((uint32_t)i << 2) | v;
        };
        if ((int64_t)F + Ei > 65536 && n_workers_hw() > 1) {
            std::thread t(stencils);
            faces();
            t.join();
        } else { faces(); stencils(); }
    }
};

// fn: 3F face nodes; ie: 4Ei interior-edge stencils.  Per node: itself + the other vertices of its faces (M), + the other vertices
// of its bending stencils (MDK); sorted, unique.  Isolated nodes (no face) own no block.  Parallel over contiguous node ranges on the
// plan-build threads (each range gathers from the node -> element lists into its own buffer; the buffers are concatenated in
// range order, so the result does not depend on the thread count).
inline void build_pattern(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie, Pattern &P, const NodeCSR *csr_in = nullptr) {
    std::unique_ptr<NodeCSR> own_csr;
    if (!csr_in) { own_csr.reset(new NodeCSR(N, F, fn, Ei, ie)); csr_in = own_csr.get(); }
    const NodeCSR &c = *csr_in;
    P.N = N;
    P.blkptrM.assign((size_t)N + 1, 0); P.blkptrK.assign((size_t)N + 1, 0);
    const int nw = (int)std::min<size_t>((size_t)n_workers_hw(), std::max<size_t>(1, (size_t)N / 4096));
    std::vector<std::vector<int32_t>> partM((size_t)nw), partK((size_t)nw);
    auto work = [&](int w) {
        const size_t lo = (size_t)N * (size_t)w / (size_t)nw, hi = (size_t)N * (size_t)(w + 1) / (size_t)nw;
        std::vector<int32_t> &oM = partM[(size_t)w], &oK = partK[(size_t)w];
        oM.reserve((hi - lo) * 8); oK.reserve((hi - lo) * 14);
        // a node's neighbour set is small (7 / 13 on a regular sheet): kept sorted and unique by insertion as the elements are walked
        // (the first version collected, std::sort-ed and std::unique-d two vectors per node: a third of the plan build at 512^2)
        std::vector<int32_t> bm, bk;
        auto insert_sorted = [](std::vector<int32_t> &v, int32_t x) {
            size_t i = v.size();
            while (i > 0 && v[i - 1] > x) --i;
            if (i > 0 && v[i - 1] == x) return;
            v.insert(v.begin() + (std::ptrdiff_t)i, x);
        };
        for (size_t a = lo; a < hi; ++a) {
            if (c.nfp[a + 1] == c.nfp[a]) continue;          // isolated: no block
            bm.clear();
            bm.push_back((int32_t)a);
            for (int32_t k = c.nfp[a]; k < c.nfp[a + 1]; ++k) {
                const int32_t *v = fn + 3 * (size_t)(c.nfl[k] >> 2);
                for (int j = 0; j < 3; ++j) insert_sorted(bm, v[j]);
            }
            bk = bm;
            for (int32_t k = c.nep[a]; k < c.nep[a + 1]; ++k) {
                const int32_t *v = ie + 4 * (size_t)(c.nel[k] >> 2);
                for (int j = 0; j < 4; ++j) insert_sorted(bk, v[j]);
            }
            P.blkptrM[a + 1] = (int64_t)bm.size(); P.blkptrK[a + 1] = (int64_t)bk.size();
            oM.insert(oM.end(), bm.begin(), bm.end()); oK.insert(oK.end(), bk.begin(), bk.end());
        }
    };
    if (nw == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < nw; ++w) th.emplace_back(work, w);
        for (auto &t : th) t.join();
    }
    for (int32_t a = 0; a < N; ++a) { P.blkptrM[a + 1] += P.blkptrM[a]; P.blkptrK[a + 1] += P.blkptrK[a]; }
    P.nbrM.resize((size_t)P.blkptrM[N]); P.nbrK.resize((size_t)P.blkptrK[N]);
    {
        std::vector<size_t> offM((size_t)nw + 1, 0), offK((size_t)nw + 1, 0);
        for (int w = 0; w < nw; ++w) { offM[(size_t)w + 1] = offM[(size_t)w] + partM[(size_t)w].size(); offK[(size_t)w + 1] = offK[(size_t)w] + partK[(size_t)w].size(); }
        auto copy = [&](int w) {
            std::copy(partM[(size_t)w].begin(), partM[(size_t)w].end(), P.nbrM.begin() + (std::ptrdiff_t)offM[(size_t)w]);
            std::copy(partK[(size_t)w].begin(), partK[(size_t)w].end(), P.nbrK.begin() + (std::ptrdiff_t)offK[(size_t)w]);
        };
        if (nw == 1) copy(0);
        else {
            std::vector<std::thread> th;
            for (int w = 0; w < nw; ++w) th.emplace_back(copy, w);
            for (auto &t : th) t.join();
        }
    }
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
using eolc::NodeCSR;   // defined above: the pattern build reads it too

#ifndef EOLC_TILE_OWN
#define EOLC_TILE_OWN 32
#endif
#ifndef EOLC_NTHREADS
#define EOLC_NTHREADS 256
#endif
#ifndef EOLC_CTAS_PER_SM
#define EOLC_CTAS_PER_SM 1
#endif
constexpr int NTHREADS = EOLC_NTHREADS;   // compute threads per CTA: phase 1 evaluates one element per thread (stencils from thread 0 up, faces from
                                     // the last thread down), phases 2 and 3 run on the same threads
#ifndef EOLC_P2_WARPS
#define EOLC_P2_WARPS 12             // warps that run phase-2 groups: 12 = the 8 compute warps + the service warpgroup, which would idle
#endif                               // between the two barriers of a tile otherwise (8: compute warps only; -3 % with 12, experiments.md p12)
constexpr int P2THREADS = 32 * EOLC_P2_WARPS;
constexpr int CTA_THREADS = NTHREADS + 128;   // + one service warpgroup: stages the inputs of the tiles ahead, issues the bulk copy-out
constexpr int CTAS_PER_SM = EOLC_CTAS_PER_SM;   // 2: experiment (half-size CTAs, smaller tiles, phases of two CTAs interleave; experiments.md)
constexpr int MAX_OWN = EOLC_TILE_OWN;   // nodes owned by a tile (<= 63: 6-bit fields)
constexpr int MAX_LOC = 128;         // distinct nodes referenced by a tile's elements (8-bit local ids)
// Only OFF-DIAGONAL element blocks are parked: every element matrix has zero row sums (translation invariance; the mass part sums
// to t8/6 per row), so the diagonal block of a node is  2 M_aa - sum of the off-diagonal blocks of its row  (phase 3).
constexpr int EDGE_STRIDE = 62;      // doubles parked per bending stencil: 6 off-diagonal blocks x 10 (9 + pad) (+2: odd number of 16-byte units)
constexpr int FACE_STRIDE = 42;      // doubles parked per face: 3 off-diagonal blocks x 10, forces 3 x 4 (3 + pad), t8 in the first force pad
constexpr int FACE_T8 = 33;          // offset of t8 (rho * 2A) inside a face slot
constexpr int ZPAD = 32;             // doubles at the start of the scratch that stay zero: padded pull entries read a zero block there (any
                                     // even offset <= 14, chosen by bank)
#ifndef EOLC_MAX_SCRATCH_DOUBLES
#define EOLC_MAX_SCRATCH_DOUBLES 14336
#endif
constexpr int MAX_SCRATCH_DOUBLES = EOLC_MAX_SCRATCH_DOUBLES;   // 112 KB of parked blocks per tile
constexpr int MAX_KSTAGE = 130 * MAX_OWN;          // doubles of MDK rows one tile stages in shared memory before the bulk copy-out
constexpr int MAX_MSTAGE = 74 * MAX_OWN;           // doubles of M rows one tile stages (expanded blocks: m on the block diagonal, explicit zeros off it)
constexpr int MAX_FSTAGE = 4 * MAX_OWN + 4;        // doubles of f one tile stages
constexpr int MAX_ITER = 255;        // PAIRS of contributions per loop of one phase-2 group (8-bit fields)
constexpr int GROUP = 32;            // phase-2 records per group (one warp)
constexpr int B_HDR = 4 + P2THREADS / 32;   // words before the per-node tables of template part B: header + the warps' group ranges

// smem offsets (in doubles) of the parked blocks inside an element slot
inline int edge_off_off(int lo, int hi) { static const int k[4][4] = {{-1, 0, 1, 2}, {0, -1, 3, 4}, {1, 3, -1, 5}, {2, 4, 5, -1}}; return 10 * k[lo][hi]; }
inline int face_off_off(int lo, int hi) { static const int k[3][3] = {{-1, 0, 1}, {0, -1, 2}, {1, 2, -1}}; return 10 * k[lo][hi]; }
inline int face_force_off(int v) { return 30 + 4 * v; }

// Phase 2 works in GROUPS of 32 records of one kind (one warp per group, one record per lane); every lane of a group runs the
// group's trip counts (shorter lists are padded with offset 0 = the zero block), so the loops are warp-uniform and the pull
// entries are stored transposed: entry of (trip r, lane l) at pulls[base + 32 r + l] (conflict-free, two 16-bit offsets each).
//   kind D (f of an owned node): loop A = the forces of its faces (3 doubles each).  Its diagonal MDK block comes from phase 3.
//   kind O (off-diagonal MDK block (own, p)): loop A = contributions parked in this orientation, loop B = parked transposed.  If the
//          column node is owned by the same tile (has2) the record also writes the mirrored block (own2, p2) = transpose, and
//          the mirrored pair has no record of its own.
//   kind M (mass block): loop A = t8 of the faces shared by the pair; flag = diagonal block (t8/12, else t8/24); has2 as for O.
// 64-bit record: off1:16 | stride1:10 | off2:16 | stride2:10 | has2:1 | flag:1 | valid:1 — staging offsets (doubles) of the first row
// of the block and of its mirror, and the distance between the block's rows (3 x the node's row degree); kind D: off1 = staged f of
// the node, off2 / stride2 = its diagonal MDK block (has2 = the node has a row at all).
inline uint64_t pack_rec(unsigned off1, unsigned stride1, unsigned off2, unsigned stride2, unsigned has2, unsigned flag) {
    return (uint64_t)off1 | ((uint64_t)stride1 << 16) | ((uint64_t)off2 << 26) | ((uint64_t)stride2 << 42) | ((uint64_t)has2 << 52) |
           ((uint64_t)flag << 53) | (1ull << 54);
}
enum { KIND_D = 0, KIND_O = 1, KIND_M = 2 };

// Template = part A (phase 1) + part B (phase 2), u32 words, every section padded to 16 bytes:
//   A: [nE | nF << 16, 0, 0, 0] [items: nE stencils (4 local ids, 8 bits each) then nF faces (3 local ids)]
//      stencil slot s is evaluated by thread s % 256, face slot s by thread 255 - s % 256
//   B: [nOwn | nGroups << 8, word offset of the phase-3 items, number of phase-3 items, 0] [per phase-2 warp: first group | groups << 16]
//      [degs: degK | degM << 8 | position of the diagonal block in the MDK row << 16 | in the M row << 24
//      per owned node] [offsKM: staging offset of the node's MDK rows | M rows << 16]
//      [offsF: staging offset of the node's f] [groups: 4 words each: kind | nA << 8 | nB << 16, pull base, warp, 0] [records: 32 x u64 per
//      group] [pulls]
//      every group names the phase-2 warp that runs it (longest-processing-time-first assignment by estimated cost)
// Staging offsets follow the PARITY of the global destination (offset of a run's first double in its value array, mod 2), so that
// staged rows and global rows are 16-byte aligned together and whole runs leave with one bulk copy (cp.async.bulk).
// Geometry blob (per tile): [template offset (16-byte units), nOwn | nLoc << 8, size A | size B << 16 (16-byte units), nRuns]
//   [local -> global node table] [runs: 4 words each: destination offset (u64, doubles) | kind << 62 (0 = MDK, 1 = M, 2 = f),
//   staging offset, length (doubles)]
struct Plan {
    int32_t n_tiles = 0, n_templates = 0;
    std::vector<uint32_t> geo;         // geometry blobs (u32 words), tile t at t * 4 * max_geo16 (fixed stride: no offset lookup)
    std::vector<uint32_t> tmpl;        // template blobs (u32 words)
    uint32_t max_geo16 = 0, max_tmplA16 = 0, max_tmplB16 = 0, max_loc = 0, max_scratch = 0, max_kstage = 0, max_mstage = 0, max_fstage = 0;
    int64_t elem_evals = 0;            // element evaluations per fill (>= F + Ei because of halo re-evaluation)
    int64_t n_runs = 0, n_groups = 0, pull_rows = 0;   // statistics
    std::vector<uint16_t> tile_elems;                  // elements evaluated per tile (faces + stencils, saturated), tile order
    std::string error;
};

struct Builder {
    int32_t N, F, Ei;
    const int32_t *fn, *ie;
    const Pattern &pat;
    const std::vector<int32_t> &nfp, &nep;
    const std::vector<uint32_t> &nfl, &nel;
    Builder(int32_t N_, int32_t F_, const int32_t *fn_, int32_t Ei_, const int32_t *ie_, const Pattern &p, const NodeCSR &csr)
        : N(N_), F(F_), Ei(Ei_), fn(fn_), ie(ie_), pat(p), nfp(csr.nfp), nep(csr.nep), nfl(csr.nfl), nel(csr.nel) {
        // zero pages from calloc are mapped on first touch: a worker pays for the part of the mesh its tiles touch, not for 6 arrays of
        // mesh size (stamps start at 1, so 0 means "never seen")
        fstamp.alloc((size_t)F); estamp.alloc((size_t)Ei); lstamp.alloc((size_t)N); local.alloc((size_t)N);
        fslot.alloc((size_t)F); eslot.alloc((size_t)Ei);
    }
    struct Zeroed {
        int32_t *p = nullptr;
        Zeroed() {}
        Zeroed(const Zeroed &) = delete;
        Zeroed &operator=(const Zeroed &) = delete;
        ~Zeroed() { free(p); }
        void alloc(size_t n) { free(p); p = static_cast<int32_t *>(calloc(n ? n : 1, sizeof(int32_t))); if (!p) throw std::bad_alloc(); }
        int32_t &operator[](size_t i) { return p[i]; }
        int32_t operator[](size_t i) const { return p[i]; }
    };
    // scratch for tile construction
    Zeroed fstamp, estamp, lstamp, local, fslot, eslot;   // *slot: element -> slot of the tile being built
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
        if (n_own > MAX_OWN || nE > 0xffff || nF > 0xffff) return false;
        if (ZPAD + (int64_t)nE * EDGE_STRIDE + (int64_t)nF * FACE_STRIDE > MAX_SCRATCH_DOUBLES) return false;
        int64_t kst = 0, mst = 0;
        for (int o = 0; o < n_own; ++o) { kst += 9 * (pat.blkptrK[own[o] + 1] - pat.blkptrK[own[o]]) + 2; mst += 9 * (pat.blkptrM[own[o] + 1] - pat.blkptrM[own[o]]) + 2; }
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
// the same bisection with the two halves of the top `depth` levels on two threads (disjoint ranges of idx; leaves in the same order)
inline void rcb_par(std::vector<int32_t> &idx, size_t lo, size_t hi, size_t leaves, const double *cx, const double *cy,
                    std::vector<std::pair<size_t, size_t>> &out, int depth) {
    if (depth <= 0 || leaves < 256) { rcb(idx, lo, hi, leaves, cx, cy, out); return; }
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
    std::vector<std::pair<size_t, size_t>> right;
    std::thread th([&]() { rcb_par(idx, mid, hi, leaves - lleaves, cx, cy, right, depth - 1); });
    rcb_par(idx, lo, mid, lleaves, cx, cy, out, depth - 1);
    th.join();
    out.insert(out.end(), right.begin(), right.end());
}

// one phase-2 record under construction
struct RecTmp {
    uint64_t rec;
    int p;                         // position of the block in its row (sort key: the same direction on a structured mesh)
    std::vector<uint16_t> A, B;    // pull offsets of loop A / loop B (single entries, paired when serialised)
};

// ---------------------------------------------------------------------------------------------------------------------
// Bank-aware layout of one template (post-pass over the unique templates; sizes do not change).
// The block pulls of phase 2 are 128-bit loads, served a quarter warp at a time: the 8 lanes' 16-byte pieces are fetched in one
// wavefront if they fall into 8 different bank groups, i.e. if (block offset / 2) mod 8 differs; otherwise the quarter takes as many
// wavefronts as its fullest bank group holds distinct addresses.  Two freedoms do not change any result's meaning:
//   (1) the slot of an element (the bank group of its blocks depends on slot mod 8): local search over slot swaps,
//   (2) the order in which a record adds its contributions (still fixed by the plan, hence deterministic): per quarter warp and
//       list position, every lane greedily takes the remaining contribution in the least used bank group; exhausted lanes read
//       a zero block placed in the least used group.
// ---------------------------------------------------------------------------------------------------------------------
inline void optimize_template(uint32_t *T, uint32_t sizeA16, int iters) {
    uint32_t *A = T, *B = T + (size_t)sizeA16 * 4;
    const int nE = (int)(A[0] & 0xffffu), nF = (int)(A[0] >> 16);
    const int nOwn = (int)(B[0] & 255u), nG = (int)((B[0] >> 8) & 255u), n4 = (nOwn + 3) & ~3;
    uint32_t *grp = B + B_HDR + 3 * n4;
    uint32_t *pulls = grp + 4 * nG + 2 * GROUP * nG;
    const int fbase = ZPAD + nE * EDGE_STRIDE;
    struct Ref { int elem; int boff; };                   // elem: stencil e -> e, face f -> nE + f; -1: zero pad
    auto decode = [&](uint32_t off) {
        Ref r{-1, 0};
        if ((int)off >= fbase) { r.elem = nE + ((int)off - fbase) / FACE_STRIDE; r.boff = ((int)off - fbase) % FACE_STRIDE; }
        else if ((int)off >= ZPAD) { r.elem = ((int)off - ZPAD) / EDGE_STRIDE; r.boff = ((int)off - ZPAD) % EDGE_STRIDE; }
        return r;
    };
    std::vector<int> slot((size_t)nE + nF);               // element -> slot within its kind
    for (int e = 0; e < nE; ++e) slot[e] = e;
    for (int f = 0; f < nF; ++f) slot[nE + f] = f;
    auto offset_of = [&](const Ref &r) { return r.elem < nE ? ZPAD + slot[r.elem] * EDGE_STRIDE + r.boff : fbase + slot[r.elem] * FACE_STRIDE + r.boff; };
    // ---- quarter rows of the O groups: up to 8 references that one 128-bit load instruction serves together
    struct QRow { Ref ref[8]; int n; };
    std::vector<QRow> rows;
    std::vector<std::vector<int>> inc((size_t)nE + nF);   // element -> rows it appears in
    for (int g = 0; g < nG; ++g) {
        if ((grp[4 * g] & 255u) != (uint32_t)KIND_O) continue;
        const int nrows = (int)((grp[4 * g] >> 8) & 255u) + (int)((grp[4 * g] >> 16) & 255u);
        const uint32_t *pl = pulls + grp[4 * g + 1];
        for (int r = 0; r < nrows; ++r)
            for (int half = 0; half < 2; ++half)
                for (int q = 0; q < GROUP / 8; ++q) {
                    QRow R; R.n = 0;
                    for (int l = 0; l < 8; ++l) {
                        const uint32_t e = pl[r * GROUP + 8 * q + l];
                        const Ref rf = decode(half ? e >> 16 : e & 0xffffu);
                        if (rf.elem < 0) continue;
                        bool dup = false;
                        for (int k = 0; k < R.n; ++k) dup |= R.ref[k].elem == rf.elem && R.ref[k].boff == rf.boff;
                        if (!dup) R.ref[R.n++] = rf;
                    }
                    if (R.n > 1) {
                        for (int k = 0; k < R.n; ++k) inc[R.ref[k].elem].push_back((int)rows.size());
                        rows.push_back(R);
                    }
                }
    }
    for (auto &v : inc) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    auto row_cost = [&](const QRow &R) {                  // wavefronts of the quarter (x 16) + a tie breaker that favours even spreading
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < R.n; ++k) ++cnt[(offset_of(R.ref[k]) >> 1) & 7];
        int mx = 0, sq = 0;
        for (int c = 0; c < 8; ++c) { mx = std::max(mx, cnt[c]); sq += cnt[c] * cnt[c]; }
        return 16 * mx + sq;
    };
    if (iters > 0 && !rows.empty()) {
        std::vector<int> cost(rows.size());
        for (size_t r = 0; r < rows.size(); ++r) cost[r] = row_cost(rows[r]);
        uint64_t rng = 0x9e3779b97f4a7c15ull;
        auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
        std::vector<int> touched;
        auto eval_swap = [&](int e0, int e1) {            // cost change of swapping the slots of e0 and e1 (same kind)
            touched.clear();
            touched.insert(touched.end(), inc[e0].begin(), inc[e0].end());
            touched.insert(touched.end(), inc[e1].begin(), inc[e1].end());
            std::sort(touched.begin(), touched.end());
            touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
            std::swap(slot[e0], slot[e1]);
            int d = 0;
            for (int r : touched) d += row_cost(rows[r]) - cost[r];
            std::swap(slot[e0], slot[e1]);
            return d;
        };
        size_t cursor = 0;
        for (int it = 0; it < iters; ++it) {
            // next row with a conflict (round robin)
            size_t r = cursor, tried = 0;
            for (; tried < rows.size(); ++tried, r = (r + 1) % rows.size()) {
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
                for (int k = 0; k < rows[r].n; ++k) mx = std::max(mx, ++cnt[(offset_of(rows[r].ref[k]) >> 1) & 7]);
                if (mx > 1) break;
            }
            if (tried == rows.size()) break;              // conflict free
            cursor = (r + 1) % rows.size();
            const QRow &R = rows[r];
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < R.n; ++k) ++cnt[(offset_of(R.ref[k]) >> 1) & 7];
            int pick = -1;
            for (int k = 0, seen = 0; k < R.n; ++k)
                if (cnt[(offset_of(R.ref[k]) >> 1) & 7] > 1 && (next() % (uint32_t)(++seen)) == 0) pick = k;
            const int e0 = R.ref[pick].elem;
            const int lo = e0 < nE ? 0 : nE, n = e0 < nE ? nE : nF;
            int best = -1, bestd = 1;
            for (int c = 0; c < 12; ++c) {
                const int e1 = lo + (int)(next() % (uint32_t)n);
                if (e1 == e0 || (slot[e1] & 7) == (slot[e0] & 7)) continue;
                const int d = eval_swap(e0, e1);
                if (d < bestd || (d == bestd && d <= 0 && (next() & 1))) { bestd = d; best = e1; }
            }
            if (best >= 0 && bestd <= 0) {
                eval_swap(e0, best);                      // fills `touched`
                std::swap(slot[e0], slot[best]);
                for (int rr : touched) cost[rr] = row_cost(rows[rr]);
            }
        }
    }
    // ---- apply the slots: items (part A) and every pull entry (part B)
    {
        std::vector<uint32_t> items((size_t)nE + nF);
        for (int e = 0; e < nE; ++e) items[slot[e]] = A[4 + e];
        for (int f = 0; f < nF; ++f) items[nE + slot[nE + f]] = A[4 + nE + f];
        for (int i = 0; i < nE + nF; ++i) A[4 + i] = items[i];
    }
    for (int g = 0; g < nG; ++g) {
        const int kind = (int)(grp[4 * g] & 255u), nA = (int)((grp[4 * g] >> 8) & 255u), nB = (int)((grp[4 * g] >> 16) & 255u);
        uint32_t *pl = pulls + grp[4 * g + 1];
        for (int i = 0; i < (nA + nB) * GROUP; ++i) {
            const Ref r0 = decode(pl[i] & 0xffffu), r1 = decode(pl[i] >> 16);
            pl[i] = (uint32_t)(r0.elem < 0 ? 0 : offset_of(r0)) | ((uint32_t)(r1.elem < 0 ? 0 : offset_of(r1)) << 16);
        }
        if (kind != KIND_O) continue;
        // ---- (2) greedy order of every lane's contributions, per quarter warp and loop
        for (int loop = 0; loop < 2; ++loop) {
            const int n = loop ? nB : nA;
            uint32_t *base = pl + (loop ? nA * GROUP : 0);
            for (int q = 0; q < GROUP / 8; ++q) {
                std::vector<uint16_t> rem[8];
                for (int l = 0; l < 8; ++l)
                    for (int pos = 0; pos < 2 * n; ++pos) {
                        const uint32_t e = base[(pos >> 1) * GROUP + 8 * q + l];
                        const uint16_t off = (uint16_t)((pos & 1) ? e >> 16 : e & 0xffffu);
                        if (off >= ZPAD) rem[l].push_back(off);
                    }
                for (int pos = 0; pos < 2 * n; ++pos) {
                    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    uint16_t chosen[8];
                    bool pad[8];
                    for (int l = 0; l < 8; ++l) {
                        pad[l] = rem[l].empty();
                        if (pad[l]) continue;
                        size_t best = 0;
                        for (size_t c = 1; c < rem[l].size(); ++c)
                            if (cnt[(rem[l][c] >> 1) & 7] < cnt[(rem[l][best] >> 1) & 7]) best = c;
                        chosen[l] = rem[l][best];
                        rem[l].erase(rem[l].begin() + best);
                        ++cnt[(chosen[l] >> 1) & 7];
                    }
                    int zres = 0;
                    for (int r = 1; r < 8; ++r) if (cnt[r] < cnt[zres]) zres = r;
                    for (int l = 0; l < 8; ++l) {
                        const uint32_t off = pad[l] ? (uint32_t)(2 * zres) : chosen[l];
                        uint32_t &w = base[(pos >> 1) * GROUP + 8 * q + l];
                        w = (pos & 1) ? (w & 0xffffu) | (off << 16) : (w & 0xffff0000u) | off;
                    }
                }
            }
        }
    }
}

// True when at least 85 % of the tiles carry at least 95 % of the heaviest tile's elements (interior tiles of a structured sheet: 233
// elements, tiles on its boundary: 205).  The fill kernel then runs phase 3 on its service warps (forces.cu, assemble_tiles_kernel<true>):
// -1.7 % on the 1024^2 sheet, but +2.5 % on the 4096 x 64^2 ensemble, a third of whose tiles are light boundary tiles.
inline bool mostly_full_tiles(const Plan &P) {
    uint32_t mx = 0;
    for (uint16_t c : P.tile_elems) mx = std::max<uint32_t>(mx, c);
    size_t full = 0;
    for (uint16_t c : P.tile_elems) full += (uint32_t)c * 100u >= mx * 95u ? 1 : 0;
    return !P.tile_elems.empty() && full * 100 >= P.tile_elems.size() * 85;
}


// Where the rows of M / MDK go when they are not simply 9 * blkptr[node] apart: EOL meshes (forces_eol.h), whose Lagrangian rows
// carry `extra` Eulerian columns behind their 3x3 blocks.  dst = value index of the node's first scalar row.
struct RowLayout {
    const int64_t *dstM, *dstK;
    const int32_t *extraM, *extraK;
};

inline bool build(int32_t N, int32_t F, const int32_t *fn, int32_t Ei, const int32_t *ie, const Pattern &pat, const double *X_hint, bool dedup,
                  Plan &P, const RowLayout *rows = nullptr, const NodeCSR *csr_in = nullptr) {
    P = Plan();
    if (N == 0) return true;
    const bool timing = getenv("EOLC_PLAN_TIMING") != nullptr;
    auto tnow = []() { return std::chrono::steady_clock::now(); };
    auto tlast = tnow();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t = tnow();
        fprintf(stderr, "[plan] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tlast).count());
        tlast = t;
    };
    std::unique_ptr<NodeCSR> own_csr;
    if (!csr_in) { own_csr.reset(new NodeCSR(N, F, fn, Ei, ie)); csr_in = own_csr.get(); }
    const NodeCSR &csr = *csr_in;
    lap("node csr");
    Builder B(N, F, fn, Ei, ie, pat, csr);
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
    lap("builder + coordinates");
    std::vector<int32_t> idx(N);
    for (int32_t a = 0; a < N; ++a) idx[a] = a;
    std::vector<std::pair<size_t, size_t>> leaves;
    if ((MAX_OWN & (MAX_OWN - 1)) == 0 || !X_hint) {
        int depth = 0;
        while ((1 << depth) < n_workers_hw()) ++depth;
        rcb_par(idx, 0, (size_t)N, ((size_t)N + MAX_OWN - 1) / MAX_OWN, cx.data(), cy.data(), leaves, depth);
    } else {
        // Tile sizes that are not a power of two: bisection leaves ragged tiles on structured meshes.  Snapped strip tiling instead:
        // strips of ~sqrt(MAX_OWN) node columns along x, cut into chunks of <= MAX_OWN nodes along y; every cut moves to the nearest
        // place where the sort coordinate changes (structured meshes: whole columns / rows), if there is one close by.
        auto snapped_cuts = [&](size_t lo, size_t hi, double want, const double *c, std::vector<size_t> &cuts) {
            cuts.clear();
            cuts.push_back(lo);
            const size_t n = hi - lo;
            const size_t parts = std::max<size_t>(1, (size_t)((double)n / want + 0.999));
            for (size_t k = 1; k < parts; ++k) {
                size_t at = lo + (size_t)((double)n * (double)k / (double)parts + 0.5);
                const size_t reach = (size_t)(want / 3.0) + 1;
                size_t best = at;
                bool found = false;
                for (size_t d = 0; d <= reach && !found; ++d) {
                    if (at + d < hi && at + d > lo && c[idx[at + d]] != c[idx[at + d - 1]]) { best = at + d; found = true; }
                    else if (at >= lo + d + 1 && at - d > lo && c[idx[at - d]] != c[idx[at - d - 1]]) { best = at - d; found = true; }
                }
                if (best > cuts.back() && best < hi) cuts.push_back(best);
            }
            cuts.push_back(hi);
        };
        std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { if (cx[a] != cx[b]) return cx[a] < cx[b]; if (cy[a] != cy[b]) return cy[a] < cy[b]; return a < b; });
        // nodes per strip: (number of distinct-ish columns per strip) x (nodes per column); estimate the column population from the data
        size_t ncol = 1;
        for (int32_t i = 1; i < N; ++i) if (cx[idx[i]] != cx[idx[i - 1]]) ++ncol;
        const double per_col = (double)N / (double)ncol;
        int side = 1;
        while ((side + 1) * (side + 1) <= MAX_OWN) ++side;
        const double strip_nodes = ncol * 4 >= (size_t)N ? (double)side * std::max(1.0, std::sqrt((double)N)) : (double)side * per_col;
        std::vector<size_t> scuts, ccuts;
        snapped_cuts(0, (size_t)N, strip_nodes, cx.data(), scuts);
        for (size_t s = 0; s + 1 < scuts.size(); ++s) {
            const size_t lo = scuts[s], hi = scuts[s + 1];
            std::sort(idx.begin() + lo, idx.begin() + hi, [&](int32_t a, int32_t b) { if (cy[a] != cy[b]) return cy[a] < cy[b]; if (cx[a] != cx[b]) return cx[a] < cx[b]; return a < b; });
            snapped_cuts(lo, hi, (double)MAX_OWN, cy.data(), ccuts);
            for (size_t c = 0; c + 1 < ccuts.size(); ++c) leaves.push_back({ccuts[c], ccuts[c + 1]});
        }
    }
    // ---- experiment knob, OFF by default (EOLC_PLAN_TILING=hex): hexagonal lattice tiling of structured sheets.  On the triangular
    //      lattice of a regular sheet (neighbours (+-1, 0), (0, +-1), +-(1, 1)) the 29-node hexagon with rows of 5, 6, 7, 6, 5 nodes
    //      touches exactly 128 stencils and 78 faces — four full stencil warps, one per scheduler, instead of the 145 stencils (a
    //      fifth, half-empty warp next to a full one on scheduler 0) of the 8 x 4 rectangle the bisection finds; its translates by
    //      u = (5, 2), v = (-2, 5) (det 29) tile the plane.  Measured on the 1024^2 sheet: 0.837 ms against 0.758 ms for the rectangles
    //      (profiles/r01/experiments.md, h1): 12 % more tiles cost more in per-tile latency (barriers, the fixed latency of a phase-2
    //      group) than the balanced phase 1 saves.  Kept for the record and for meshes with the other aspect ratios.
    {
        const char *tl = getenv("EOLC_PLAN_TILING");
        std::vector<double> ux, uy;
        if (X_hint && MAX_OWN >= 29 && tl && strcmp(tl, "hex") == 0) {
            ux.assign(cx.begin(), cx.end()); uy.assign(cy.begin(), cy.end());
            std::sort(ux.begin(), ux.end()); ux.erase(std::unique(ux.begin(), ux.end()), ux.end());
            std::sort(uy.begin(), uy.end()); uy.erase(std::unique(uy.begin(), uy.end()), uy.end());
        }
        if (!ux.empty() && (uint64_t)ux.size() * (uint64_t)uy.size() == (uint64_t)N) {
            const int32_t nJ = (int32_t)uy.size();
            std::vector<int32_t> gi(N), gj(N);
            std::vector<char> hit((size_t)N, 0);
            bool grid = true;
            for (int32_t a = 0; a < N && grid; ++a) {
                gi[a] = (int32_t)(std::lower_bound(ux.begin(), ux.end(), cx[a]) - ux.begin());
                gj[a] = (int32_t)(std::lower_bound(uy.begin(), uy.end(), cy[a]) - uy.begin());
                char &h = hit[(size_t)gi[a] * nJ + gj[a]];
                if (h) grid = false;
                h = 1;
            }
            if (grid) {      // full tensor grid in the hint coordinates
                static const int HROW[5][2] = {{0, 5}, {0, 6}, {0, 7}, {1, 6}, {2, 5}};   // (first column, length) of the hexagon's rows
                const int U0 = 5, U1 = 2, V0 = -2, V1 = 5, DET = 29;
                std::vector<uint64_t> key((size_t)N);
                for (int32_t a = 0; a < N; ++a) {
                    const int i = gi[a], j = gj[a];
                    int64_t ta = 0, tb = 0;
                    for (int r = 0; r < 5; ++r)
                        for (int c = HROW[r][0]; c < HROW[r][0] + HROW[r][1]; ++c) {
                            const int64_t di = i - r, dj = j - c, p = di * V1 - dj * V0, q = U0 * dj - U1 * di;
                            if (p % DET == 0 && q % DET == 0) { ta = p / DET; tb = q / DET; r = 5; break; }
                        }
                    key[a] = ((uint64_t)(ta + (1 << 20)) << 32) | (uint64_t)(tb + (1 << 20));
                }
                std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
                leaves.clear();
                for (size_t lo = 0; lo < (size_t)N;) {
                    size_t hi = lo + 1;
                    while (hi < (size_t)N && key[idx[hi]] == key[idx[lo]]) ++hi;
                    leaves.push_back({lo, hi});
                    lo = hi;
                }
            }
        }
    }
    lap("bisection");
    // ---- split leaves that exceed the kernel's capacities.  The capacity check of every leaf (the same element gathering the tile
    //      builders do) runs on the plan-build threads, each with its own scratch; only leaves that fail it are bisected further,
    //      one after the other.  Per-worker scratch (Builder) is created once and reused by the tile builders below.
    std::vector<int32_t> faces, edges;
    const int n_workers = (int)std::min<size_t>((size_t)n_workers_hw(), std::max<size_t>(1, leaves.size() / 64));
    std::vector<std::unique_ptr<Builder>> wb((size_t)n_workers);
    {
        std::vector<char> pre_ok(leaves.size(), 0);
        auto check_range = [&](int w) {
            if (n_workers > 1 || !wb[0]) wb[(size_t)w].reset(new Builder(N, F, fn, Ei, ie, pat, csr));
            std::vector<int32_t> fcs, eds;
            for (size_t k = leaves.size() * (size_t)w / (size_t)n_workers; k < leaves.size() * (size_t)(w + 1) / (size_t)n_workers; ++k) {
                const auto r = leaves[k];
                pre_ok[k] = r.second > r.first && r.second - r.first <= (size_t)MAX_OWN &&
                            wb[(size_t)w]->fits(idx.data() + r.first, (int)(r.second - r.first), fcs, eds);
            }
        };
        if (n_workers == 1) check_range(0);
        else {
            std::vector<std::thread> th;
            for (int w = 0; w < n_workers; ++w) th.emplace_back(check_range, w);
            for (auto &t : th) t.join();
        }
        std::vector<std::pair<size_t, size_t>> ok;
        for (size_t k = 0; k < leaves.size(); ++k) {
            if (pre_ok[k]) { ok.push_back(leaves[k]); continue; }
            std::vector<std::pair<size_t, size_t>> work(1, leaves[k]);
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
        }
        leaves.swap(ok);
    }
    P.n_tiles = (int32_t)leaves.size();
    struct Run { uint64_t dst; uint32_t kind, src, len; };
    struct Grp { int kind, nA, nB, first, count; long cost; int warp; };
    const bool verify_dedup = getenv("EOLC_PLAN_VERIFY_DEDUP") != nullptr;
    const bool sort_own = getenv("EOLC_PLAN_SORT_OWN") != nullptr;   // developer knob (read once: getenv in a sort comparator costs a third of the build)
    // One worker builds the tiles [t0, t1) into its own partial plan Q (templates deduplicated within the range, geometry blobs
    // back to back with offsets in geo_off); the ranges are merged in tile order below, so the result does not depend on the
    // number of workers.
    auto build_range = [&](Builder &Bw, size_t t0, size_t t1, Plan &Q, std::vector<uint32_t> &geo_off, std::vector<std::pair<uint32_t, uint32_t>> &unique_tmpl,
                           std::vector<uint64_t> &unique_hash) -> bool {
        std::vector<int32_t> faces, edges;
        std::unordered_map<uint64_t, std::vector<uint32_t>> seen;   // hash -> template offsets (16-byte units)
        // Structural shortcut: the template of a tile is a function of its LOCAL structure only — the element -> local node lists, the
        // degrees and staging offsets of the owned nodes, and the neighbour rows of the owned nodes written in local ids (that fixes
        // every block position p, every mirrored position and which neighbours are owned).  On a structured mesh thousands of tiles
        // share one signature; only the first of them builds (sorts, groups, serialises) the template.  EOLC_PLAN_VERIFY_DEDUP=1
        // builds every template anyway and checks it against the one the shortcut would have reused.
        struct SigEntry { std::vector<uint32_t> words; uint32_t toff, sizeA16, sizeB16; int64_t n_groups, pull_rows; };
        std::unordered_map<uint64_t, std::vector<SigEntry>> sig_seen;
        std::vector<uint32_t> sigw;
        std::vector<uint32_t> T;                                   // template under construction
        std::vector<RecTmp> recs[3];
        std::vector<Run> runs;
        std::vector<Grp> groups;
            for (size_t t = t0; t < t1; ++t) {
            int32_t *own = idx.data() + leaves[t].first;
            const int n_own = (int)(leaves[t].second - leaves[t].first);
            std::sort(own, own + n_own);
            Bw.fits(own, n_own, faces, edges);
            std::sort(faces.begin(), faces.end());
            std::sort(edges.begin(), edges.end());
            const int nE = (int)edges.size(), nF = (int)faces.size();
            const int fbase = ZPAD + nE * EDGE_STRIDE;
            Q.elem_evals += nE + nF;
            Q.tile_elems.push_back((uint16_t)std::min(nE + nF, 0xffff));
            Q.max_scratch = std::max<uint32_t>(Q.max_scratch, (uint32_t)(fbase + nF * FACE_STRIDE));
            // local node table: owned nodes first (ascending), then halo nodes by first appearance; element -> slot maps
            ++Bw.stamp;
            std::vector<uint32_t> loc;
            auto lid = [&](int32_t g) {
                if (Bw.lstamp[g] != Bw.stamp) { Bw.lstamp[g] = Bw.stamp; Bw.local[g] = (int32_t)loc.size(); loc.push_back((uint32_t)g); }
                return (uint32_t)Bw.local[g];
            };
            for (int o = 0; o < n_own; ++o) lid(own[o]);
            std::vector<uint32_t> items((size_t)nE + nF, 0u);
            for (int s = 0; s < nE; ++s) {
                const int32_t *v = ie + 4 * (size_t)edges[s];
                items[s] = lid(v[0]) | (lid(v[1]) << 8) | (lid(v[2]) << 16) | (lid(v[3]) << 24);
                Bw.eslot[edges[s]] = s;
            }
            for (int s = 0; s < nF; ++s) {
                const int32_t *v = fn + 3 * (size_t)faces[s];
                items[nE + s] = lid(v[0]) | (lid(v[1]) << 8) | (lid(v[2]) << 16);
                Bw.fslot[faces[s]] = s;
            }
            Q.max_loc = std::max<uint32_t>(Q.max_loc, (uint32_t)loc.size());
            // ---- runs of owned nodes with consecutive ids (adjacent rows in the global arrays) and the staging offsets, which follow
            //      the parity of the run's destination so that one bulk copy moves the run
            std::vector<uint32_t> degs((size_t)n_own, 0u), offsKM((size_t)n_own, 0u), offsF((size_t)n_own, 0u);
            runs.clear();
            uint32_t stage_end[3] = {0, 0, 0};
            for (int kind = 0; kind < 3; ++kind) {
                uint32_t cur = 0;
                int o = 0;
                while (o < n_own) {
                    auto len_of = [&](int q) {
                        return kind == 0 ? 9u * (uint32_t)(pat.blkptrK[own[q] + 1] - pat.blkptrK[own[q]])
                             : kind == 1 ? 9u * (uint32_t)(pat.blkptrM[own[q] + 1] - pat.blkptrM[own[q]]) : 3u;
                    };
                    // destination of the node's first scalar row; with Eulerian columns (EOL meshes, forces_eol.h) a node's three
                    // scalar rows are `extra` entries apart and leave as three runs, everything else as before
                    const int32_t xtra = !rows ? 0 : kind == 0 ? rows->extraK[own[o]] : kind == 1 ? rows->extraM[own[o]] : 0;
                    const uint64_t dst = kind == 2 ? (uint64_t)3 * (uint64_t)own[o]
                                       : rows ? (uint64_t)(kind == 0 ? rows->dstK[own[o]] : rows->dstM[own[o]])
                                              : (uint64_t)(9 * (kind == 0 ? pat.blkptrK[own[o]] : pat.blkptrM[own[o]]));
                    cur = ((cur + 1u) & ~1u) + (uint32_t)(dst & 1u);      // even slot + the destination's parity
                    const uint32_t src = cur;
                    int o1 = o;
                    if (xtra) {        // extra is even, so the three rows keep the parity relation of the first
                        if (kind == 0) offsKM[o] |= cur; else offsKM[o] |= cur << 16;
                        const uint32_t row = len_of(o) / 3u;
                        for (uint32_t j = 0; j < 3u && row; ++j) runs.push_back({dst + (uint64_t)j * (row + (uint32_t)xtra), (uint32_t)kind, src + j * row, row});
                        cur += 3u * row;
                        o = o + 1;
                        continue;
                    }
                    for (;;) {
                        if (kind == 0) offsKM[o1] |= cur; else if (kind == 1) offsKM[o1] |= cur << 16; else offsF[o1] = cur;
                        cur += len_of(o1);
                        if (o1 + 1 < n_own && own[o1 + 1] == own[o1] + 1 &&
                            !(rows && kind != 2 && (kind == 0 ? rows->extraK[own[o1 + 1]] : rows->extraM[own[o1 + 1]]))) ++o1; else break;
                    }
                    if (cur - src) runs.push_back({dst, (uint32_t)kind, src, cur - src});
                    o = o1 + 1;
                }
                stage_end[kind] = cur;
            }
            if (stage_end[0] > 0xffffu || stage_end[1] > 0xffffu) { Q.error = "tile " + std::to_string(t) + ": staging overflow"; return false; }
            Q.max_kstage = std::max(Q.max_kstage, stage_end[0]); Q.max_mstage = std::max(Q.max_mstage, stage_end[1]); Q.max_fstage = std::max(Q.max_fstage, stage_end[2]);
            // ---- structural signature (see above)
            const SigEntry *reuse = nullptr;
            uint64_t sigh = 1469598103934665603ull;
            if (dedup) {
                sigw.clear();
                sigw.push_back((uint32_t)n_own); sigw.push_back((uint32_t)nE); sigw.push_back((uint32_t)nF);
                sigw.insert(sigw.end(), items.begin(), items.end());
                for (int o = 0; o < n_own; ++o) {
                    const int32_t a = own[o];
                    sigw.push_back((uint32_t)(pat.blkptrK[a + 1] - pat.blkptrK[a]) | ((uint32_t)(pat.blkptrM[a + 1] - pat.blkptrM[a]) << 16));
                    sigw.push_back(offsKM[o]); sigw.push_back(offsF[o]);
                    // every neighbour of an owned node shares an element with it, hence has a local id already
                    for (int64_t q = pat.blkptrK[a]; q < pat.blkptrK[a + 1]; ++q) sigw.push_back(lid(pat.nbrK[q]));
                    for (int64_t q = pat.blkptrM[a]; q < pat.blkptrM[a + 1]; ++q) sigw.push_back(lid(pat.nbrM[q]) | 0x80000000u);
                }
                for (uint32_t w : sigw) { sigh ^= w; sigh *= 1099511628211ull; }
                auto it = sig_seen.find(sigh);
                if (it != sig_seen.end())
                    for (const SigEntry &e : it->second)
                        if (e.words == sigw) { reuse = &e; break; }
            }
            uint32_t toff = 0, sizeA16 = 0, sizeB16 = 0;
            if (reuse && !verify_dedup) {
                toff = reuse->toff; sizeA16 = reuse->sizeA16; sizeB16 = reuse->sizeB16;
                Q.n_groups += reuse->n_groups; Q.pull_rows += reuse->pull_rows;
            } else {
            const int64_t groups_before = Q.n_groups, pulls_before = Q.pull_rows;
            // ---- phase-2 records
            for (auto &r : recs) r.clear();
            auto owned_index = [&](int32_t g) { return (Bw.lstamp[g] == Bw.stamp && Bw.local[g] < n_own) ? Bw.local[g] : -1; };
            for (int o = 0; o < n_own; ++o) {
                const int32_t a = own[o];
                const int64_t b0 = pat.blkptrK[a], m0 = pat.blkptrM[a];
                const int deg = (int)(pat.blkptrK[a + 1] - b0), degM = (int)(pat.blkptrM[a + 1] - m0);
                if (deg > 255) { Q.error = "node " + std::to_string(a) + " has more than 255 neighbours"; return false; }
                degs[o] = (uint32_t)deg | ((uint32_t)degM << 8);
                if (deg) degs[o] |= (uint32_t)(find_block(pat.blkptrK, pat.nbrK, a, a) - b0) << 16;
                if (degM) degs[o] |= (uint32_t)(find_block(pat.blkptrM, pat.nbrM, a, a) - m0) << 24;
                for (int p = 0; p < std::max(deg, 1); ++p) {
                    const int32_t b = deg ? pat.nbrK[b0 + p] : a;
                    const int ob = b == a ? -1 : owned_index(b);
                    if (ob >= 0 && b < a) continue;      // the pair is assembled by the record of (b, a), which mirrors it into this row
                    RecTmp R, Rm;
                    // faces ascending, then stencils ascending
                    for (int32_t k = Bw.nfp[a]; k < Bw.nfp[a + 1]; ++k) {
                        const int32_t f = Bw.nfl[k] >> 2;
                        const int va = Bw.nfl[k] & 3, sf = Bw.fslot[f];
                        const int32_t *v = fn + 3 * (size_t)f;
                        const int base = fbase + sf * FACE_STRIDE;
                        if (b == a) { R.A.push_back((uint16_t)(base + face_force_off(va))); Rm.A.push_back((uint16_t)(base + FACE_T8)); continue; }
                        for (int vj = 0; vj < 3; ++vj)
                            if (v[vj] == b) {
                                (va < vj ? R.A : R.B).push_back((uint16_t)(base + face_off_off(std::min(va, vj), std::max(va, vj))));
                                Rm.A.push_back((uint16_t)(base + FACE_T8));
                            }
                    }
                    for (int32_t k = Bw.nep[a]; k < Bw.nep[a + 1]; ++k) {
                        const int32_t e = Bw.nel[k] >> 2;
                        const int ia = Bw.nel[k] & 3, se = Bw.eslot[e];
                        const int32_t *v = ie + 4 * (size_t)e;
                        const int base = ZPAD + se * EDGE_STRIDE;
                        if (b == a) continue;                // diagonal block: phase 3
                        for (int ij = 0; ij < 4; ++ij)
                            if (v[ij] == b) (ia < ij ? R.A : R.B).push_back((uint16_t)(base + edge_off_off(std::min(ia, ij), std::max(ia, ij))));
                    }
                    const unsigned has2 = ob >= 0 ? 1u : 0u;
                    if (has2 && pat.blkptrK[b + 1] - pat.blkptrK[b] > 255) { Q.error = "node " + std::to_string(b) + " has more than 255 neighbours"; return false; }
                    if (b == a) {
                        R.rec = pack_rec(offsF[o], 0, (offsKM[o] & 0xffffu) + 3u * (unsigned)p, 3u * (unsigned)deg, deg ? 1u : 0u, 0);
                    } else {
                        unsigned off2 = 0, s2 = 0;
                        if (has2) {
                            const unsigned p2 = (unsigned)(find_block(pat.blkptrK, pat.nbrK, b, a) - pat.blkptrK[b]);
                            off2 = (offsKM[ob] & 0xffffu) + 3u * p2; s2 = 3u * (unsigned)(pat.blkptrK[b + 1] - pat.blkptrK[b]);
                        }
                        R.rec = pack_rec((offsKM[o] & 0xffffu) + 3u * (unsigned)p, 3u * (unsigned)deg, off2, s2, has2, 0);
                    }
                    R.p = p;
                    recs[b == a ? KIND_D : KIND_O].push_back(std::move(R));
                    if (!Rm.A.empty()) {   // the pair shares a face -> mass block
                        const unsigned pM = (unsigned)(find_block(pat.blkptrM, pat.nbrM, a, b) - m0);
                        unsigned off2 = 0, s2 = 0;
                        if (has2) {
                            const unsigned pM2 = (unsigned)(find_block(pat.blkptrM, pat.nbrM, b, a) - pat.blkptrM[b]);
                            off2 = (offsKM[ob] >> 16) + 3u * pM2; s2 = 3u * (unsigned)(pat.blkptrM[b + 1] - pat.blkptrM[b]);
                        }
                        Rm.rec = pack_rec((offsKM[o] >> 16) + 3u * pM, 3u * (unsigned)degM, off2, s2, has2, b == a ? 1u : 0u);
                        Rm.p = (int)pM;
                        recs[KIND_M].push_back(std::move(Rm));
                    }
                }
            }
            // ---- groups: records of one kind, sorted by trip counts (stable) so that the lanes of a group run similar lists
            auto pairs = [](const std::vector<uint16_t> &l) { return (int)(l.size() + 1) / 2; };
            groups.clear();
            for (int kind = 0; kind < 3; ++kind) {
                auto &rv = recs[kind];
                if (kind != KIND_D)
                    std::stable_sort(rv.begin(), rv.end(), [&](const RecTmp &u, const RecTmp &v) {
                        const int cu = pairs(u.A) + pairs(u.B), cv = pairs(v.A) + pairs(v.B);
                        if (cu != cv) return cu > cv;
                        if (pairs(u.A) != pairs(v.A)) return pairs(u.A) > pairs(v.A);
                        // same position in the row = same direction on a structured mesh: neighbouring lanes then pull the same block
                        // of neighbouring elements, whose slots fall into different bank groups
                        if (sort_own) return false;
                        return u.p < v.p;
                    });
                for (size_t g0 = 0; g0 < rv.size();) {
                    // (half-width groups for the long lists were tried: a group's time is set by the latency of its trips, not by its
                    // width, so splitting only lengthened every warp's chain: 0.85 -> 0.99 ms, profiles/r01/experiments.md)
                    const size_t width = GROUP;
                    Grp G{kind, 0, 0, (int)g0, (int)std::min(rv.size() - g0, width), 0, 0};
                    for (size_t k = g0; k < g0 + (size_t)G.count; ++k) { G.nA = std::max(G.nA, pairs(rv[k].A)); G.nB = std::max(G.nB, pairs(rv[k].B)); }
                    if (G.nA > MAX_ITER || G.nB > MAX_ITER) { Q.error = "tile " + std::to_string(t) + ": contribution list overflow"; return false; }
                    // rough clocks of one group on the kernel (latency bound: a fixed part + so much per trip), for the warp assignment below
                    // (measured on the 1024^2 sheet, B200: an O group takes ~1150 + 220 per trip, an M group ~600 + 50, a D group ~700 + 50;
                    // the fixed part is the latency chain record -> row table -> staged stores; nearly empty groups take about half)
                    G.cost = kind == KIND_D ? 700 + 50L * G.nA : kind == KIND_O ? 1150 + 220L * (G.nA + G.nB) : 600 + 50L * G.nA;
                    if (G.count <= 8) G.cost /= 2;
                    groups.push_back(G);
                    g0 += (size_t)G.count;
                }
            }
            std::stable_sort(groups.begin(), groups.end(), [](const Grp &u, const Grp &v) { return u.cost > v.cost; });
            if (groups.size() > 255) { Q.error = "tile " + std::to_string(t) + ": too many phase-2 groups"; return false; }
            {
                // longest-processing-time-first assignment of the groups to the phase-2 warps; then ordered by warp (stable), so that a
                // warp walks its groups in descending cost
                long load[P2THREADS / 32] = {0};
                for (Grp &G : groups) {
                    int w = 0;
                    for (int k = 1; k < P2THREADS / 32; ++k) if (load[k] < load[w]) w = k;
                    G.warp = w; load[w] += G.cost;
                }
                std::stable_sort(groups.begin(), groups.end(), [](const Grp &u, const Grp &v) { return u.warp < v.warp; });
            }
            // ---- serialise the template
            T.clear();
            T.push_back((uint32_t)nE | ((uint32_t)nF << 16));
            T.push_back(0); T.push_back(0); T.push_back(0);
            T.insert(T.end(), items.begin(), items.end());
            while (T.size() % 4) T.push_back(0);
            sizeA16 = (uint32_t)(T.size() / 4);
            T.push_back((uint32_t)n_own | ((uint32_t)groups.size() << 8));
            T.push_back(0); T.push_back(0); T.push_back(0);
            for (int w = 0; w < P2THREADS / 32; ++w) {          // groups [first, first + count) of warp w (the groups are ordered by warp)
                uint32_t first = 0, count = 0;
                for (size_t g = 0; g < groups.size(); ++g) if (groups[g].warp == w) { if (!count) first = (uint32_t)g; ++count; }
                T.push_back(first | (count << 16));
            }
            auto push_padded = [&](const std::vector<uint32_t> &v) { T.insert(T.end(), v.begin(), v.end()); while (T.size() % 4) T.push_back(0); };
            push_padded(degs); push_padded(offsKM); push_padded(offsF);
            {
                uint32_t pbase = 0;
                for (const Grp &G : groups) {
                    // bit 24: a mass group without a diagonal record — skipped as a whole when the caller says M is unchanged
                    uint32_t offdiag_only = G.kind == KIND_M ? 1u : 0u;
                    if (G.kind == KIND_M)
                        for (int l = 0; l < G.count; ++l) if ((recs[KIND_M][G.first + l].rec >> 53) & 1ull) offdiag_only = 0u;
                    T.push_back((uint32_t)G.kind | ((uint32_t)G.nA << 8) | ((uint32_t)G.nB << 16) | (offdiag_only << 24));
                    T.push_back(pbase); T.push_back((uint32_t)G.warp); T.push_back(0);
                    pbase += (uint32_t)(G.nA + G.nB) * GROUP;
                }
                Q.pull_rows += pbase / GROUP;
            }
            for (const Grp &G : groups) {
                const auto &rv = recs[G.kind];
                for (int l = 0; l < GROUP; ++l) {
                    const uint64_t r = l < G.count ? rv[G.first + l].rec : 0ull;
                    T.push_back((uint32_t)r); T.push_back((uint32_t)(r >> 32));
                }
            }
            for (const Grp &G : groups) {
                const auto &rv = recs[G.kind];
                for (int loop = 0; loop < 2; ++loop)
                    for (int r = 0; r < (loop ? G.nB : G.nA); ++r)
                        for (int l = 0; l < GROUP; ++l) {
                            uint32_t e = 0;
                            if (l < G.count) {
                                const std::vector<uint16_t> &L = loop ? rv[G.first + l].B : rv[G.first + l].A;
                                if ((size_t)2 * r < L.size()) e = L[2 * r];
                                if ((size_t)2 * r + 1 < L.size()) e |= (uint32_t)L[2 * r + 1] << 16;
                            }
                            T.push_back(e);
                        }
            }
            while (T.size() % 4) T.push_back(0);
            {
                // phase-3 items: one per entry (j, k), j <= k, of every owned node's diagonal MDK block, 3 words each:
                //   start of the strided row sum (staging offset of row j, column k of the node's first block) | deg << 16 | (j == k) << 24,
                //   destination | mirrored destination << 16,   staging offset of the node's M_aa
                const uint32_t p3_at = (uint32_t)(T.size() - (size_t)sizeA16 * 4);
                uint32_t n_items = 0;
                // entry-major order: neighbouring lanes work on the same entry of consecutive nodes, whose rows are 9 deg doubles apart
                // (an odd number for the usual odd degrees: conflict-free 64-bit accesses)
                for (uint32_t j = 0; j < 3; ++j)
                    for (uint32_t k = j; k < 3; ++k)
                        for (int o = 0; o < n_own; ++o) {
                            const uint32_t deg = degs[o] & 255u, pd = (degs[o] >> 16) & 255u, pdM = degs[o] >> 24;
                            if (!deg) continue;
                            const uint32_t base = offsKM[o] & 0xffffu, mo = (offsKM[o] >> 16) + 3 * pdM;
                            T.push_back((base + 3 * deg * j + k) | (deg << 16) | ((j == k ? 1u : 0u) << 24));
                            T.push_back((base + 3 * deg * j + 3 * pd + k) | ((base + 3 * deg * k + 3 * pd + j) << 16));
                            T.push_back(mo);
                            ++n_items;
                        }
                T[(size_t)sizeA16 * 4 + 1] = p3_at;
                T[(size_t)sizeA16 * 4 + 2] = n_items;
                while (T.size() % 4) T.push_back(0);
            }
            Q.n_groups += (int64_t)groups.size();
            sizeB16 = (uint32_t)(T.size() / 4) - sizeA16;
            if (sizeA16 > 0xffffu || sizeB16 > 0xffffu) { Q.error = "tile " + std::to_string(t) + ": template too large"; return false; }
            Q.max_tmplA16 = std::max(Q.max_tmplA16, sizeA16); Q.max_tmplB16 = std::max(Q.max_tmplB16, sizeB16);
            // ---- deduplicate
            bool found = false;
            uint64_t h = 1469598103934665603ull;
            if (dedup) {
                for (uint32_t w : T) { h ^= w; h *= 1099511628211ull; }
                auto it = seen.find(h);
                if (it != seen.end())
                    for (uint32_t cand : it->second)
                        if ((size_t)cand * 4 + T.size() <= Q.tmpl.size() && memcmp(Q.tmpl.data() + (size_t)cand * 4, T.data(), T.size() * 4) == 0) {
                            toff = cand; found = true; break;
                        }
            }
            if (!found) {
                toff = (uint32_t)(Q.tmpl.size() / 4);
                Q.tmpl.insert(Q.tmpl.end(), T.begin(), T.end());
                if (dedup) seen[h].push_back(toff);
                ++Q.n_templates;
                unique_tmpl.push_back({toff, sizeA16});
                unique_hash.push_back(h);          // of the template's words (dedup only): the merge does not hash them again
            }
            if (reuse) {   // EOLC_PLAN_VERIFY_DEDUP: the shortcut would have reused this template — it must be the one just built
                if (reuse->toff != toff || reuse->sizeA16 != sizeA16 || reuse->sizeB16 != sizeB16) {
                    Q.error = "tile " + std::to_string(t) + ": structural signature matched a different template"; return false;
                }
            } else if (dedup) {
                sig_seen[sigh].push_back(SigEntry{sigw, toff, sizeA16, sizeB16, Q.n_groups - groups_before, Q.pull_rows - pulls_before});
            }
            }
            // ---- geometry blob
            geo_off.push_back((uint32_t)(Q.geo.size() / 4));
            const size_t g0 = Q.geo.size();
            Q.geo.push_back(toff);
            Q.geo.push_back((uint32_t)n_own | ((uint32_t)loc.size() << 8));
            Q.geo.push_back(sizeA16 | (sizeB16 << 16));
            Q.geo.push_back((uint32_t)runs.size());
            Q.geo.insert(Q.geo.end(), loc.begin(), loc.end());
            while (Q.geo.size() % 4) Q.geo.push_back(0);
            for (const Run &r : runs) {
                const uint64_t d = r.dst | ((uint64_t)r.kind << 62);
                Q.geo.push_back((uint32_t)d); Q.geo.push_back((uint32_t)(d >> 32));
                Q.geo.push_back(r.src); Q.geo.push_back(r.len);
            }
            Q.n_runs += (int64_t)runs.size();
            Q.max_geo16 = std::max<uint32_t>(Q.max_geo16, (uint32_t)((Q.geo.size() - g0) / 4));
            if (Q.geo.size() / 4 >= ((size_t)1 << 32) || Q.tmpl.size() / 4 >= ((size_t)1 << 32)) { Q.error = "plan too large"; return false; }
        }
        geo_off.push_back((uint32_t)(Q.geo.size() / 4));
        return true;
    };
    lap("capacity check");
    std::vector<Plan> parts((size_t)n_workers);
    std::vector<std::vector<uint32_t>> part_geo_off((size_t)n_workers);
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> part_unique((size_t)n_workers);
    std::vector<std::vector<uint64_t>> part_hash((size_t)n_workers);
    std::vector<char> part_ok((size_t)n_workers, 1);
    auto range_of = [&](int w) { return std::make_pair(leaves.size() * (size_t)w / (size_t)n_workers, leaves.size() * (size_t)(w + 1) / (size_t)n_workers); };
    if (n_workers == 1) {
        part_ok[0] = build_range(B, 0, leaves.size(), parts[0], part_geo_off[0], part_unique[0], part_hash[0]) ? 1 : 0;
    } else {
        std::vector<std::thread> th;
        for (int w = 0; w < n_workers; ++w)
            th.emplace_back([&, w]() {
                // own scratch (stamps, slot maps), created for the capacity check above; the node CSR is shared
                const auto r = range_of(w);
                part_ok[w] = build_range(*wb[(size_t)w], r.first, r.second, parts[w], part_geo_off[w], part_unique[w], part_hash[w]) ? 1 : 0;
            });
        for (auto &t : th) t.join();
    }
    for (int w = 0; w < n_workers; ++w) if (!part_ok[w]) { P.error = parts[w].error; return false; }
    lap("tiles (workers)");
    // ---- merge in tile order: templates deduplicated across the ranges, geometry blobs re-laid with a fixed stride
    std::vector<std::pair<uint32_t, uint32_t>> unique_tmpl;   // (offset, size of part A) of every stored template, 16-byte units
    for (int w = 0; w < n_workers; ++w) {
        const Plan &Q = parts[w];
        P.max_geo16 = std::max(P.max_geo16, Q.max_geo16); P.max_tmplA16 = std::max(P.max_tmplA16, Q.max_tmplA16); P.max_tmplB16 = std::max(P.max_tmplB16, Q.max_tmplB16);
        P.max_loc = std::max(P.max_loc, Q.max_loc); P.max_scratch = std::max(P.max_scratch, Q.max_scratch);
        P.max_kstage = std::max(P.max_kstage, Q.max_kstage); P.max_mstage = std::max(P.max_mstage, Q.max_mstage); P.max_fstage = std::max(P.max_fstage, Q.max_fstage);
        P.tile_elems.insert(P.tile_elems.end(), Q.tile_elems.begin(), Q.tile_elems.end());
        P.elem_evals += Q.elem_evals; P.n_runs += Q.n_runs; P.n_groups += Q.n_groups; P.pull_rows += Q.pull_rows;
    }
    {
        std::unordered_map<uint64_t, std::vector<uint32_t>> seen;
        std::vector<uint32_t> fixed((size_t)P.n_tiles * P.max_geo16 * 4, 0u);
        {   // one allocation for the templates (an unstructured mesh brings one per tile: ~100 MB at 512^2, appended range by range)
            size_t total_words = 0;
            for (int w = 0; w < n_workers; ++w) total_words += parts[w].tmpl.size();
            P.tmpl.reserve(P.tmpl.size() + total_words);
        }
        size_t t = 0;
        for (int w = 0; w < n_workers; ++w) {
            const Plan &Q = parts[w];
            // local template offset -> global offset (first use in tile order decides the global order)
            std::unordered_map<uint32_t, uint32_t> remap;
            // local templates lie back to back in the order they were stored: offset -> (words, hash of the words from the worker)
            const std::vector<std::pair<uint32_t, uint32_t>> &lu = part_unique[w];
            std::unordered_map<uint32_t, size_t> local_index;
            local_index.reserve(lu.size() * 2);
            for (size_t q = 0; q < lu.size(); ++q) local_index.emplace(lu[q].first, q);
            auto tmpl_words = [&](uint32_t loff) {     // length of the local template starting at loff: up to the next one
                const size_t q = local_index.at(loff);
                const size_t end = q + 1 < lu.size() ? (size_t)lu[q + 1].first : Q.tmpl.size() / 4;
                return (end - loff) * 4;
            };
            const size_t ntw = part_geo_off[w].size() - 1;
            for (size_t k = 0; k < ntw; ++k, ++t) {
                const uint32_t *g = Q.geo.data() + (size_t)part_geo_off[w][k] * 4;
                const size_t glen = (size_t)(part_geo_off[w][k + 1] - part_geo_off[w][k]) * 4;
                uint32_t *dst = fixed.data() + t * P.max_geo16 * 4;
                memcpy(dst, g, glen * 4);
                const uint32_t loff = g[0];
                auto it = remap.find(loff);
                if (it == remap.end()) {
                    const size_t words = tmpl_words(loff);
                    const uint32_t *src = Q.tmpl.data() + (size_t)loff * 4;
                    const uint64_t h = dedup ? part_hash[w][local_index.at(loff)] : 0;      // FNV-1a of the words, computed by the worker
                    uint32_t goff = 0;
                    bool found = false;
                    if (dedup) {
                        auto si = seen.find(h);
                        if (si != seen.end())
                            for (uint32_t cand : si->second)
                                if ((size_t)cand * 4 + words <= P.tmpl.size() && memcmp(P.tmpl.data() + (size_t)cand * 4, src, words * 4) == 0) { goff = cand; found = true; break; }
                    }
                    if (!found) {
                        goff = (uint32_t)(P.tmpl.size() / 4);
                        P.tmpl.insert(P.tmpl.end(), src, src + words);
                        if (dedup) seen[h].push_back(goff);
                        ++P.n_templates;
                        unique_tmpl.push_back({goff, g[2] & 0xffffu});
                    }
                    it = remap.emplace(loff, goff).first;
                }
                dst[0] = it->second;
            }
            if (P.tmpl.size() / 4 >= ((size_t)1 << 32)) { P.error = "plan too large"; return false; }
        }
        P.geo.swap(fixed);
    }
    lap("merge");
    {
        // bank-aware post-pass over the stored templates; the slot search only where templates are shared (structured meshes)
        const char *ob = getenv("EOLC_PLAN_BANK_ITERS");
        const long budget = ob ? atol(ob) : 4000;
        const int iters = (int)std::max<long>(0, std::min<long>(budget, 400000 / (long)std::max<size_t>(1, unique_tmpl.size())));
        // templates are disjoint ranges of P.tmpl: one worker per slice of the list, same result for any worker count
        const int nw = std::max(1, std::min<int>(n_workers_hw(), (int)unique_tmpl.size()));
        // The search is a deterministic function of the template's words and the iteration budget, and a remeshed or re-planned mesh
        // brings mostly the same templates again: results are remembered process-wide (keyed by the words before the search).
        struct BankCache { std::mutex mu; std::unordered_map<uint64_t, std::vector<std::pair<std::vector<uint32_t>, std::vector<uint32_t>>>> map; size_t words = 0; };
        static BankCache cache;
        auto tmpl_len = [&](size_t k) {     // words of stored template k: up to the next stored template (they are appended in order)
            size_t end = P.tmpl.size();
            for (const auto &u : unique_tmpl) if ((size_t)u.first * 4 > (size_t)unique_tmpl[k].first * 4) end = std::min(end, (size_t)u.first * 4);
            return end - (size_t)unique_tmpl[k].first * 4;
        };
        const int eff_iters = iters < 20 ? 0 : iters;
        auto slice = [&](int w) {
            for (size_t k = (size_t)w; k < unique_tmpl.size(); k += (size_t)nw) {
                uint32_t *T = P.tmpl.data() + (size_t)unique_tmpl[k].first * 4;
                const size_t len = unique_tmpl.size() <= 64 ? tmpl_len(k) : 0;   // cache only where templates are shared (structured meshes)
                uint64_t h = 1469598103934665603ull ^ (uint64_t)eff_iters;
                std::vector<uint32_t> before;
                if (len) {
                    before.assign(T, T + len);
                    for (uint32_t x : before) { h ^= x; h *= 1099511628211ull; }
                    std::lock_guard<std::mutex> g(cache.mu);
                    auto it = cache.map.find(h);
                    if (it != cache.map.end()) {
                        bool hit = false;
                        for (const auto &e : it->second) if (e.first == before) { memcpy(T, e.second.data(), len * 4); hit = true; break; }
                        if (hit) continue;
                    }
                }
                optimize_template(T, unique_tmpl[k].second, eff_iters);
                if (len) {
                    std::lock_guard<std::mutex> g(cache.mu);
                    if (cache.words + 2 * len < ((size_t)64 << 20)) {      // at most 256 MB of remembered templates
                        cache.map[h].push_back({std::move(before), std::vector<uint32_t>(T, T + len)});
                        cache.words += 2 * len;
                    }
                }
            }
        };
        if (nw == 1) slice(0);
        else {
            std::vector<std::thread> th;
            for (int w = 0; w < nw; ++w) th.emplace_back(slice, w);
            for (auto &t : th) t.join();
        }
    }
    lap("bank post-pass");
    return true;
}

}  // namespace tiles
}  // namespace eolc
