// clip_thread.cu — K5t: one THREAD builds one Voronoi cell (the small-cell tier of the clip pass).
//
// Replaces, per cell, exactly what clip.cu replaces:
//   Cell::compute_voronoi_cell / cut_with_point          interface.rs:257-334
//   ExpandingSearch::{new, expand_all_*}                  celery.rs:882-1075
//   Polyhedron::{find_outgoing_edge, cut_with_plane}      polyhedron.rs:396-642
//   Polyhedron::{weighted_normal, compute_volume, compute_neighbors}   polyhedron.rs:776-881
//   VoronoiFace::compute_area                             interface.rs:408-410
//
// Why a second form of the same pass: the warp-per-cell kernel spends ~20 k warp-instructions
// on a cell whose serial work is ~60 k thread-instructions — a cut has ~5 crossings, a cell ~27 vertices, and
// 32 lanes plus the ballots / shuffles / list ranking that coordinate them are mostly overhead.  Here every
// lane runs the reference's own serial algorithm on its own cell; nothing is coordinated inside a cell.
//   * The half-edge mesh of each cell lives in SHARED MEMORY, lane-interleaved (element j of lane l at
//     [j][l]): whatever slots the 32 lanes touch, they hit 32 different banks.  One half-edge = one 32-bit word
//     {next, flip, target, face}; free half-edge and face slots are chained through the free slots themselves,
//     which IS pool.rs's LIFO free list (pool.rs:85-121), so slot numbers — and with them find_outgoing_edge's
//     "first edge in slot order", every face's starting edge, the face order, the fan anchors and the
//     summation orders — are the reference's, and volumes / areas come out bit-identical to the CPU oracle.
//   * Tables are sized for the common cell (44 vertices / 136 half-edges / 24 faces = 1764 B per thread:
//     128 cell-building threads fill the SM's 227 KB).  A cell that outgrows them, meets a vertex ON a plane
//     (the reference's Incident case, polyhedron.rs:555-565) or runs out of search table is handed back through
//     the failed-cell list and redone by the warp-per-cell kernel, like every other tier's leftovers.
//   * The search is NOT done by the thread that builds the cell.  First version (profiles/r02_thread_kernel.md):
//     every lane walked its own search table — three dependent global loads per grid cell with one warp per
//     scheduler to hide them, and lanes that need 1 or 100 table steps to find their next candidate in the same
//     loop: 4.8 of 32 lanes active, 142 ms per 10^6 cells.  Now the CTA is warp-specialised: 4 CONSUMER warps
//     (thread per cell) and PW PRODUCER warps.  A producer serves one consumer lane at a time with all 32 lanes:
//     32 search-table entries per step (coalesced), their grid cells' particle runs, |r|^2 against the consumer's
//     published threshold, survivors compacted IN TABLE ORDER into that lane's ring of candidate slots in shared
//     memory.  The published threshold may be stale — it only ever shrinks, so a stale one lets through a superset;
//     the consumer re-tests every candidate against its own threshold, and a candidate beyond the reference's
//     stopping entry has |r|^2 >= key > threshold: the tested sequence is exactly the reference's.
//   * Divergence is managed, not avoided: each consumer lane is a small state machine (take a candidate /
//     classify / cut / results) and the warp runs a phase when enough lanes wait for it.
//   * arithmetic is tess_math.cuh's, operation for operation the reference's.
#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "cube_tables.cuh"
#include "tess_math.cuh"

namespace tess {

namespace {

constexpr uint32_t TFULL = 0xffffffffu;

#ifndef TESS_T_CUT_MIN
#define TESS_T_CUT_MIN 24  // lanes with a cut pending before the cut phase runs (unless nobody can do anything else)
#endif
#ifndef TESS_T_TEST_MIN
#define TESS_T_TEST_MIN 20 // lanes with a candidate in hand before the classification phase runs
#endif
#ifndef TESS_T_SPINS
#define TESS_T_SPINS 2     // short sleeps a warp takes for starved lanes before it runs a phase below its threshold
#endif
#ifndef TESS_T_DONE_MIN
#define TESS_T_DONE_MIN 8  // finished cells waiting before the results phase runs
#endif

#ifndef TESS_T_PWARPS
#define TESS_T_PWARPS 8    // producer warps per CTA (each serves 128 / PW consumer lanes)
#endif
#ifndef TESS_T_PASSES
#define TESS_T_PASSES 4    // 32-particle passes per producer step
#endif
#ifndef TESS_T_TRIES
#define TESS_T_TRIES 2     // candidates a lane may take (and reject) per round
#endif

struct ThreadCfg {
    static constexpr int V = 44, E = 136, F = 24;
    static constexpr int WARPS = 4;                 // consumer warps: one thread per cell
    static constexpr int PWARPS = TESS_T_PWARPS;    // producer warps
    static constexpr int LPP = 32 * WARPS / PWARPS; // consumer lanes per producer warp
    static constexpr int QD = 8;                    // ring of candidate slots per consumer lane
    static constexpr uint32_t NONE = 0xFFu;
};
static_assert(ThreadCfg::LPP >= 1 && ThreadCfg::LPP <= 32 && ThreadCfg::LPP * ThreadCfg::PWARPS == 32 * ThreadCfg::WARPS, "producer warps must divide the consumer lanes");

// The tables of the 32 cells of one warp, lane-interleaved.
struct __align__(16) ThreadTables {
    double vx[ThreadCfg::V][32], vy[ThreadCfg::V][32], vz[ThreadCfg::V][32];
    uint32_t edge[ThreadCfg::E][32];   // {next, flip, target, face}; a free slot holds {next free slot, NONE, NONE, NONE}
    uint32_t fnbr[ThreadCfg::F][32];   // Face.point_index as the neighbour's sorted slot; WALL0 + k for container face k
    uint8_t vedge[ThreadCfg::V][32];   // one half-edge that starts at the vertex (the others: next(flip(e)) twice)
    uint8_t fstart[ThreadCfg::F][32];  // Face.starting_edge_index; free face slots are chained through it
};

// What producers and consumers exchange, per consumer thread c (0..127).  Ring items: a sorted slot (< 2^31), or
// Q_HALO | table index (the entry touches a plane this rank does not hold), or Q_END / Q_END_EXH (end of the walk: a key
// above the threshold / the end of a table that does not cover the grid).
struct __align__(16) ThreadShared {
    ThreadTables tab[ThreadCfg::WARPS];
    double thr[32 * ThreadCfg::WARPS];              // consumer -> producer: current threshold (4 max|v|^2, or the caller's radius); -2: stop now
    uint32_t q[ThreadCfg::QD][32 * ThreadCfg::WARPS];
    uint32_t head[32 * ThreadCfg::WARPS];           // items taken (consumer), items published (producer): free-running counters
    uint32_t tail[32 * ThreadCfg::WARPS];
    uint32_t cell[32 * ThreadCfg::WARPS];           // consumer -> producer: sorted slot of the cell under construction / C_IDLE / C_EXIT
};
static_assert(sizeof(ThreadShared) <= 232448, "tables and rings must fit the 227 KB of one SM");
#ifdef TESS_T_STATS  // development build: phase / participation statistics of the state machines (printed after the launch).
// Counters: 0 consumer rounds; 1/2 cut phases / lanes in them; 3/4 classify phases / lanes; 5/6 result phases / lanes; 7 idle
// sleeps; 11 ring items consumed; 13-16 producer steps / items / candidates looked at / ring room; 17/18 producer polls / with
// work; 19 producer step cycles; 21-23 consumer cycles in classify / cut / results; 24 consumer warp cycles.
__device__ unsigned long long g_tstats[32];
#define TCLK() clock64()
#define TSTAT(i, v) atomicAdd(&g_tstats[i], (unsigned long long)(v))
#define TSTAT_LEADER(i, v) do { if (lane == 0) atomicAdd(&g_tstats[i], (unsigned long long)(v)); } while (0)
#else
#define TSTAT(i, v) ((void)0)
#define TCLK() 0ll
#define TSTAT_LEADER(i, v) ((void)0)
#endif
constexpr uint32_t Q_END = 0xFFFFFFFFu, Q_END_EXH = 0xFFFFFFFEu, Q_HALO = 0x80000000u;
constexpr uint32_t C_EXIT = 0xFFFFFFFFu, C_IDLE = 0xFFFFFFFEu;

// release / acquire on shared-memory words (CTA scope)
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
#ifdef TESS_WARP_EMU
    *reinterpret_cast<volatile uint32_t*>(p) = v;
#else
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
#ifdef TESS_WARP_EMU
    return *reinterpret_cast<const volatile uint32_t*>(p);
#else
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
#endif
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
__device__ __forceinline__ void st_volatile_f64(double* p, double v) { *reinterpret_cast<volatile double*>(p) = v; }

constexpr uint32_t WALL0 = 0xFFFFFFF0u;

enum : int { S_NEW = 0, S_FETCH = 1, S_TEST = 2, S_CUT = 3, S_DONE = 4, S_EXIT = 5 };

__device__ __forceinline__ uint32_t ew_pack(uint32_t next, uint32_t flip, uint32_t tgt, uint32_t face) { return next | (flip << 8) | (tgt << 16) | (face << 24); }
__device__ __forceinline__ uint32_t ew_next(uint32_t w) { return w & 0xFFu; }
__device__ __forceinline__ uint32_t ew_flip(uint32_t w) { return (w >> 8) & 0xFFu; }
__device__ __forceinline__ uint32_t ew_tgt(uint32_t w) { return (w >> 16) & 0xFFu; }
__device__ __forceinline__ uint32_t ew_face(uint32_t w) { return w >> 24; }

// the mesh of one lane
struct TMesh {
    ThreadTables* t;
    int lane;
    unsigned long long vlive;
    uint32_t flive;
    uint32_t e_head, e_hwm, e_nfree;  // Pool<HalfEdge>: free-list head (NONE = empty), slots ever used, length of the free list
    uint32_t f_head, f_hwm;           // Pool<Face>

    __device__ __forceinline__ double x(uint32_t v) const { return t->vx[v][lane]; }
    __device__ __forceinline__ double y(uint32_t v) const { return t->vy[v][lane]; }
    __device__ __forceinline__ double z(uint32_t v) const { return t->vz[v][lane]; }
    __device__ __forceinline__ uint32_t ew(uint32_t e) const { return t->edge[e][lane]; }
    __device__ __forceinline__ void set_ew(uint32_t e, uint32_t w) { t->edge[e][lane] = w; }

    // Pool::add (pool.rs:85-110): most recently freed slot first, else append.  Capacity is checked by the caller.
    __device__ __forceinline__ uint32_t alloc_edge() {
        if (e_head != ThreadCfg::NONE) {
            const uint32_t s = e_head;
            e_head = ew_next(ew(s));
            --e_nfree;
            return s;
        }
        return e_hwm++;
    }
    // Pool::remove (pool.rs:113-121)
    __device__ __forceinline__ void free_edge(uint32_t s) {
        set_ew(s, e_head | 0xFFFFFF00u);
        e_head = s;
        ++e_nfree;
    }
    __device__ __forceinline__ int alloc_face() {
        uint32_t s;
        if (f_head != ThreadCfg::NONE) {
            s = f_head;
            f_head = t->fstart[s][lane];
        } else if (f_hwm < (uint32_t)ThreadCfg::F) {
            s = f_hwm++;
        } else {
            return -1;
        }
        flive |= 1u << s;
        return (int)s;
    }
    __device__ __forceinline__ void free_face(uint32_t s) {
        t->fstart[s][lane] = (uint8_t)f_head;
        f_head = s;
        flive &= ~(1u << s);
    }

    // Polyhedron::build_cube (polyhedron.rs:268-392) translated by -p (interface.rs:266)
    __device__ void build_cube(const double* box, double px, double py, double pz) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            // FDL FDR FUR FUL BDL BDR BUR BUL (polyhedron.rs:288-295); corner + (-p)
            const bool xh = (v == 1) | (v == 2) | (v == 5) | (v == 6);
            const bool yh = v >= 4;
            const bool zh = (v == 2) | (v == 3) | (v == 6) | (v == 7);
            t->vx[v][lane] = addd(xh ? box[3] : box[0], -px);
            t->vy[v][lane] = addd(yh ? box[4] : box[1], -py);
            t->vz[v][lane] = addd(zh ? box[5] : box[2], -pz);
            t->vedge[v][lane] = kCubeVout[3 * v];
        }
#pragma unroll
        for (int e = 0; e < 24; ++e) {
            const uint32_t c = kCubeEdges[e];  // {flip, target, next}
            t->edge[e][lane] = ew_pack(c & 0xFFu, (c >> 16) & 0xFFu, (c >> 8) & 0xFFu, (uint32_t)e >> 2);
        }
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            t->fstart[f][lane] = (uint8_t)(4 * f);  // FU RU BU LU UF DF (polyhedron.rs:319-379)
            t->fnbr[f][lane] = WALL0 + (uint32_t)f;
        }
        vlive = 0xFFull;
        flive = 0x3Fu;
        e_head = ThreadCfg::NONE;
        e_hwm = 24;
        e_nfree = 0;
        f_head = ThreadCfg::NONE;
        f_hwm = 6;
    }

    // max |v|^2 over the live vertices (left-associated dot, like Vector3::mag_sq) and the vertex that attains it
    __device__ double max_radius_sq(uint32_t& arg) const {
        double m = -1.0;
        uint32_t a = 0;
        for (unsigned long long b = vlive; b; b &= b - 1ull) {
            const uint32_t v = (uint32_t)__ffsll((long long)b) - 1u;
            const double X = x(v), Y = y(v), Z = z(v);
            const double r2 = dot3(X, Y, Z, X, Y, Z);
            if (r2 > m) {
                m = r2;
                a = v;
            }
        }
        arg = a;
        return m < 0.0 ? 0.0 : m;
    }
};

// 160 bits in five registers, set with a run-time index (half-edge slots that die in the current cut)
struct Bits160 {
    uint32_t w0, w1, w2, w3, w4;
    __device__ __forceinline__ void clear() { w0 = w1 = w2 = w3 = w4 = 0u; }
    __device__ __forceinline__ void set(uint32_t i) {
        const uint32_t b = 1u << (i & 31u), k = i >> 5;
        w0 |= k == 0u ? b : 0u;
        w1 |= k == 1u ? b : 0u;
        w2 |= k == 2u ? b : 0u;
        w3 |= k == 3u ? b : 0u;
        w4 |= k == 4u ? b : 0u;
    }
};
static_assert(ThreadCfg::E <= 160, "Bits160 holds one bit per half-edge slot");

constexpr int TCUT_OK = 1, TCUT_FAIL = -1;

// ---------------------------------------------------------------------------------------------
// Polyhedron::cut_with_plane (polyhedron.rs:438-642) for a plane that has vertices Outside and none Incident,
// on a mesh whose vertices are all 3-valent (true from the start cube on as long as no vertex was ever Incident).
// `in` / `out` are the Inside / Outside vertex sets.  Serial, in the reference's own order:
//   1. find_outgoing_edge (:413-432): among the half-edges Outside -> Inside (found by turning around each Outside
//      vertex) the one in the lowest slot; the walk starts on its flip.  The same sweep lists the half-edges with
//      both ends Outside — what clean_up (:645-730) frees — and the faces they belong to;
//   2. the walk (:475-623): per crossed face, follow the loop from the outgoing half-edge to the re-entering one,
//      create the intersection vertex (:567-572), the bridge (:582-587) and the NEXT crossing's cap half-edge
//      (:592-598) — slots come from the LIFO free list in exactly this order;
//   3. close the cap loop, free the redundant last cap half-edge, then the dead half-edges and the
//      dead faces in ascending slot order (the order clean_up visits them in).
// Returns TCUT_OK or TCUT_FAIL (tables too small / mesh not as assumed: the cell is handed back).
// ---------------------------------------------------------------------------------------------
__device__ int thread_cut(TMesh& M, const Plane& pl, uint32_t nbr_ref, unsigned long long in, unsigned long long out, uint32_t& n_new) {
    ThreadTables* t = M.t;
    const int lane = M.lane;
    // ---- 1. around the Outside vertices ----------------------------------------------------------
    uint32_t K = 0, rmin = 0xFFFFu;
    Bits160 dying;
    dying.clear();
    uint32_t dying_faces = 0;
    for (unsigned long long b = out; b; b &= b - 1ull) {
        const uint32_t v = (uint32_t)__ffsll((long long)b) - 1u;
        uint32_t e = t->vedge[v][lane];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t w = M.ew(e);
            const uint32_t tg = ew_tgt(w);
            if ((out >> tg) & 1ull) {
                dying.set(e);
                dying_faces |= 1u << ew_face(w);
            } else {
                ++K;
                rmin = e < rmin ? e : rmin;
            }
            e = ew_next(M.ew(ew_flip(w)));  // the next half-edge that starts at v
        }
    }
    if (K == 0u) return TCUT_FAIL;  // every vertex Outside: not a cell of this particle (rounding) — let the reference-shaped walk decide
    // ---- capacity: K vertices, 2K + 1 half-edges, one face ---------------------------------------
    {
        const int vfree = ThreadCfg::V - __popcll(M.vlive);
        const int efree = (int)M.e_nfree + (ThreadCfg::E - (int)M.e_hwm);
        if ((int)K > vfree || 2 * (int)K + 1 > efree) return TCUT_FAIL;
    }
    // first_outside_face_edge_index, then the cap face (:478-484)
    const uint32_t cap_first = M.alloc_edge();
    const int cap_face = M.alloc_face();
    if (cap_face < 0) return TCUT_FAIL;  // (nothing else has been written yet; the cell is abandoned anyway)
    t->fnbr[cap_face][lane] = nbr_ref;
    t->fstart[cap_face][lane] = (uint8_t)cap_first;
    // ---- 2. the walk ------------------------------------------------------------------------------
    const uint32_t o0 = ew_flip(M.ew(rmin));
    uint32_t o = o0, ck = cap_first, ck_prev = ThreadCfg::NONE, nv_prev = ThreadCfg::NONE;
    unsigned long long vfree_mask = ~M.vlive;
    unsigned long long newbits = 0ull;
    uint32_t crossed = 0;
    for (uint32_t i = 0; i < K; ++i) {
        const uint32_t wo = M.ew(o);
        const uint32_t f = ew_face(wo), fo = ew_flip(wo);
        uint32_t pv = ew_tgt(wo);   // previous_vertex_index (:491), Outside
        uint32_t r = ew_next(wo);   // :506
        uint32_t wr = M.ew(r);
        uint32_t cv = ew_tgt(wr);
        int guard = ThreadCfg::E;
        while (!((in >> cv) & 1ull)) {  // :529-544
            pv = cv;
            r = ew_next(wr);
            wr = M.ew(r);
            cv = ew_tgt(wr);
            if (--guard < 0) return TCUT_FAIL;
        }
        // Pool::add for the vertex: any free slot will do (vertex slots carry no order the results depend on)
        const uint32_t nv = (uint32_t)__ffsll((long long)vfree_mask) - 1u;
        vfree_mask &= vfree_mask - 1ull;
        newbits |= 1ull << nv;
        const Vec3 a = {M.x(pv), M.y(pv), M.z(pv)};
        const Vec3 b = {M.x(cv), M.y(cv), M.z(cv)};
        const Vec3 X = intersection(pl, a, b);  // :567-572 (a = outside end, b = inside end)
        t->vx[nv][lane] = X.x;
        t->vy[nv][lane] = X.y;
        t->vz[nv][lane] = X.z;
        t->vedge[nv][lane] = (uint8_t)r;  // the re-entering half-edge now starts at the new vertex
        const uint32_t br = M.alloc_edge();       // :582-587
        const uint32_t ck_next = M.alloc_edge();  // :592-598 (the last one is the redundant twin of cap_first)
        M.set_ew(br, ew_pack(r, ck, nv, f));
        M.set_ew(ck, ew_pack(ck_prev, br, nv_prev, (uint32_t)cap_face));  // the first crossing's next / target are patched below
        M.set_ew(o, ew_pack(br, fo, nv_prev, f));                           // :550 + :590
        t->fstart[f][lane] = (uint8_t)o;                                    // :578-580
        crossed |= 1u << f;
        o = ew_flip(wr);  // :603-607
        ck_prev = ck;
        ck = ck_next;
        nv_prev = nv;
    }
    if (o != o0) return TCUT_FAIL;  // the crossings do not close into one loop (rounding broke convexity)
    // ---- 3. close the loop, free what was cut off --------------------------------------------------
    {
        uint8_t* p0 = reinterpret_cast<uint8_t*>(&t->edge[o0][lane]);
        p0[2] = (uint8_t)nv_prev;  // first outgoing half-edge ends at the last intersection
        uint8_t* pc = reinterpret_cast<uint8_t*>(&t->edge[cap_first][lane]);
        pc[2] = (uint8_t)nv_prev;
        pc[0] = (uint8_t)ck_prev;  // last_paired_cap_edge
    }
    M.free_edge(ck);  // the redundant cap half-edge goes back first
#define TESS_T_FREE_WORD(W, BASE)                                  \
    for (uint32_t m_ = dying.W; m_; m_ &= m_ - 1u) M.free_edge((BASE) + (uint32_t)__ffs((int)m_) - 1u);
    TESS_T_FREE_WORD(w0, 0u)
    TESS_T_FREE_WORD(w1, 32u)
    TESS_T_FREE_WORD(w2, 64u)
    TESS_T_FREE_WORD(w3, 96u)
    TESS_T_FREE_WORD(w4, 128u)
#undef TESS_T_FREE_WORD
    // a face whose half-edges died and which the plane does not cross has lost all of them
    for (uint32_t m = dying_faces & ~crossed & M.flive; m; m &= m - 1u) M.free_face((uint32_t)__ffs((int)m) - 1u);
    M.vlive = (M.vlive & ~out) | newbits;
    n_new += K;
    return TCUT_OK;
}

// ---------------------------------------------------------------------------------------------
// Producer: ExpandingSearch::expand_all_* (celery.rs:971-1075) for the cells of LPP consumer lanes, 32 search-table
// entries per step.  Lane j (< LPP) keeps the cursor of consumer thread c0 + j in its registers.
// ---------------------------------------------------------------------------------------------
__device__ void producer_warp(const ClipParams& P, ThreadShared* S, const int pw, const int lane) {
    const GridSpec& G = P.grid;
    const int cpd = (int)G.cpd;
    const bool radius_mode = !(P.search_radius != P.search_radius);  // not NaN
    const int c = pw * ThreadCfg::LPP + (lane < ThreadCfg::LPP ? lane : 0);  // the consumer thread this lane keeps the cursor of
    const bool serving = lane < ThreadCfg::LPP;
    uint32_t cell_seen = C_IDLE, ti = 0, off = 0, tail = 0;
    bool walk_done = true;
    double px = 0, py = 0, pz = 0;
    int hx = 0, hy = 0, hz = 0;

    for (;;) {
        // ---- who needs candidates --------------------------------------------------------------------
        bool gone = true, want = false;
        uint32_t room = 0;
        if (serving) {
            const uint32_t cs = ld_acquire(&S->cell[c]);
            if (cs != C_EXIT) {
                gone = false;
                if (cs != cell_seen && cs != C_IDLE) {
                    // a new cell (ExpandingSearch::new, celery.rs:882-902: home cell of the position)
                    cell_seen = cs;
                    const double2* q = reinterpret_cast<const double2*>(P.sorted + cs);
                    const double2 a = __ldg(q);
                    px = a.x; py = a.y; pz = __ldg(reinterpret_cast<const double*>(q + 1));
                    hx = (int)axis_index(px, G.xmin, G.xmax, G.ix, G.cpd);
                    hy = (int)axis_index(py, G.ymin, G.ymax, G.iy, G.cpd);
                    hz = (int)axis_index(pz, G.zmin, G.zmax, G.iz, G.cpd);
                    ti = 0; off = 0;
                    walk_done = false;
                }
                if (!walk_done) {
                    room = (uint32_t)ThreadCfg::QD - (tail - ld_acquire(&S->head[c]));
                    want = room >= (uint32_t)ThreadCfg::QD / 2u;
                }
            }
        }
        if (__all_sync(TFULL, gone)) break;
        uint32_t need = __ballot_sync(TFULL, want);
        TSTAT_LEADER(17, 1); TSTAT_LEADER(18, need ? 1 : 0);
        if (!need) {
            __nanosleep(100);
            continue;
        }
        // ---- one step of the walk for each of them -----------------------------------------------------
        while (need) {
            const long long t_step0 = TCLK();
            const int j = __ffs((int)need) - 1;
            need &= need - 1u;
            const int cj = pw * ThreadCfg::LPP + j;
            const uint32_t ti0 = __shfl_sync(TFULL, ti, j), off0 = __shfl_sync(TFULL, off, j), tail0 = __shfl_sync(TFULL, tail, j);
            const uint32_t room0 = __shfl_sync(TFULL, room, j), self = __shfl_sync(TFULL, cell_seen, j);
            const double qx = __shfl_sync(TFULL, px, j), qy = __shfl_sync(TFULL, py, j), qz = __shfl_sync(TFULL, pz, j);
            const int h0 = __shfl_sync(TFULL, hx, j), h1 = __shfl_sync(TFULL, hy, j), h2 = __shfl_sync(TFULL, hz, j);
            const double thr = ld_volatile_f64(&S->thr[cj]);
            // lane i looks at table entry ti0 + i; the walk stops at the first entry whose key exceeds the threshold
            // (celery.rs:1036) or at the end of the table
            const uint32_t e_idx = ti0 + (uint32_t)lane;
            const bool in_table = e_idx < P.table_len;
            ShellEntry e;
            e.key = 0.0; e.di = e.dj = e.dk = e.pad = 0;
            if (in_table) e = P.table[e_idx];
            const uint32_t stopmask = __ballot_sync(TFULL, !in_table || e.key > thr);
            int s_lane = stopmask ? __ffs((int)stopmask) - 1 : 32;
            uint32_t cur = 0, end = 0;
            bool halo = false;
            if (lane < s_lane) {
                const int gx = h0 + e.di, gy = h1 + e.dj, gz = h2 + e.dk;
                if (!(gx < 0 || gx >= cpd || gy < 0 || gy >= cpd || gz < 0 || gz >= cpd)) {
                    if (gx < (int)G.local_lo || gx >= (int)G.local_hi) {
                        halo = true;  // a plane this rank does not hold
                    } else {
                        const uint32_t gc = ((uint32_t)(gx - (int)G.local_lo) * G.cpd + (uint32_t)gy) * G.cpd + (uint32_t)gz;
                        cur = __ldg(P.delim + gc);
                        end = __ldg(P.delim + gc + 1);
                    }
                }
                if (lane == 0) cur = (end - cur > off0) ? cur + off0 : end;  // the part of the first entry's run already handed over
            }
            // The items of this step in table order: per entry its halo marker or the particles of its run, then (if the walk
            // ends inside the step) the end marker.  Flattened: item p belongs to the entry whose prefix covers p, so the
            // 32 lanes load 32 particles at a time whatever the lengths of the runs.
            const bool ends = s_lane < 32;
            uint32_t n = halo ? 1u : end - cur;
            if (ends && lane == s_lane) n = 1u;
            uint32_t inc = n;  // inclusive scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(TFULL, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t total = __shfl_sync(TFULL, inc, 31);
            const uint32_t exc = inc - n;
            const uint32_t lt = (1u << lane) - 1u;
            uint32_t emitted = 0, p0 = 0, p_next = 0;
            bool full = false;
#pragma unroll 1
            for (int pass_no = 0; pass_no < TESS_T_PASSES && p0 < total && !full; ++pass_no) {
                const uint32_t p = p0 + (uint32_t)lane;
                // entry of item p: the number of entries whose inclusive prefix is <= p
                int en = 0;
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) {
                    const uint32_t v = __shfl_sync(TFULL, inc, en + st - 1);
                    if (v <= p) en += st;
                }
                const uint32_t e_exc = __shfl_sync(TFULL, exc, en), e_cur = __shfl_sync(TFULL, cur, en);
                const bool e_halo = __shfl_sync(TFULL, halo ? 1 : 0, en) != 0;
                const bool is_end = ends && en == s_lane;
                uint32_t item = e_cur + (p - e_exc);
                bool ok = p < total;
                if (ok) {
                    if (is_end) {
                        item = (ti0 + (uint32_t)s_lane < P.table_len || P.table_full) ? Q_END : Q_END_EXH;
                    } else if (e_halo) {
                        item = Q_HALO | (ti0 + (uint32_t)en);
                    } else {
                        // interface.rs:280-312: self, group, then (security mode) |r|^2 against the threshold
                        if (item == self) ok = false;  // interface.rs:283/301 (by index)
                        if (ok && P.target_group != -1)    // interface.rs:284/293 (-2: no particle carries the requested group)
                            if (!(P.target_group >= 0 && P.groups_sorted[item] == (uint64_t)P.target_group)) ok = false;
                        if (ok && !radius_mode) {
                            const double2* cq = reinterpret_cast<const double2*>(P.sorted + item);
                            const double2 a = __ldg(cq);
                            const double zz = __ldg(reinterpret_cast<const double*>(cq + 1));
                            const double rx = subd(a.x, qx), ry = subd(a.y, qy), rz = subd(zz, qz);  // interface.rs:322-326
                            if (dot3(rx, ry, rz, rx, ry, rz) >= thr) ok = false;  // cannot have a vertex Outside (header of clip.cu)
                        }
                    }
                }
                const uint32_t m = __ballot_sync(TFULL, ok);
                const uint32_t rank = emitted + (uint32_t)__popc(m & lt);
                if (ok && rank < room0) S->q[(tail0 + rank) % (uint32_t)ThreadCfg::QD][cj] = item;
                const uint32_t cnt = (uint32_t)__popc(m);
                if (emitted + cnt > room0) {
                    // the ring is full: the walk resumes after the last item that went in
                    uint32_t mm = m;
                    for (uint32_t r = emitted + 1u; r < room0; ++r) mm &= mm - 1u;
                    p_next = p0 + (uint32_t)__ffs((int)mm);
                    emitted = room0;
                    full = true;
                } else {
                    emitted += cnt;
                    p0 += 32u;
                    p_next = p0 < total ? p0 : total;
                    full = emitted == room0;
                }
            }
            // where the next step starts
            uint32_t n_ti, n_off;
            bool n_done = false;
            if (p_next >= total) {
                n_ti = ti0 + (ends ? (uint32_t)s_lane : 32u);
                n_off = 0;
                n_done = ends;  // the end marker was the last item
            } else {
                int en = 0;
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) {
                    const uint32_t v = __shfl_sync(TFULL, inc, en + st - 1);
                    if (v <= p_next) en += st;
                }
                const uint32_t e_exc = __shfl_sync(TFULL, exc, en);
                n_ti = ti0 + (uint32_t)en;
                n_off = (en == 0 ? off0 : 0u) + (p_next - e_exc);
            }
            const uint32_t n_tail = tail0 + emitted;
            TSTAT_LEADER(19, TCLK() - t_step0);
            TSTAT_LEADER(13, 1); TSTAT_LEADER(14, emitted); TSTAT_LEADER(15, total); TSTAT_LEADER(16, room0);
            __syncwarp();
            if (lane == j) {
                ti = n_ti; off = n_off; walk_done = n_done; tail = n_tail;
                st_release(&S->tail[cj], n_tail);  // the items written above (by all lanes, ordered by the __syncwarp) become visible with it
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The kernel: persistent CTAs, one per SM; every consumer lane pulls cells from the work counter.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * (ThreadCfg::WARPS + ThreadCfg::PWARPS), 1) clip_thread_kernel(const ClipParams P) {
#ifdef TESS_WARP_EMU  // tests/emu: this source run lane by lane on the CPU (test infrastructure only)
    unsigned char* smem_raw = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    ThreadShared* S = reinterpret_cast<ThreadShared*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 32 * ThreadCfg::WARPS) {
        S->head[threadIdx.x] = 0u;
        S->tail[threadIdx.x] = 0u;
        S->cell[threadIdx.x] = C_IDLE;
        S->thr[threadIdx.x] = -2.0;
    }
    __syncthreads();
    if (warp >= ThreadCfg::WARPS) {
        producer_warp(P, S, warp - ThreadCfg::WARPS, lane);
        return;
    }

    // ---- consumer ------------------------------------------------------------------------------------
    const int c = (int)threadIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    TMesh M;
    M.t = &S->tab[warp];
    M.lane = lane;
    M.vlive = 0ull;
    M.flive = 0u;
    M.e_head = M.f_head = ThreadCfg::NONE;
    M.e_hwm = M.e_nfree = M.f_hwm = 0u;
    const bool radius_mode = !(P.search_radius != P.search_radius);  // not NaN
    const double last_key = P.table_len ? P.table[P.table_len - 1].key : -1.0;

    // ---- per-lane state of the cell under construction ------------------------------------------
    int state = S_NEW;
    uint32_t work = 0, self_slot = 0xFFFFFFFFu, status = 0, head = 0;
    bool failed = false;
    double px = 0, py = 0, pz = 0;
    double stop_thr = 0.0;                   // 4 * max|v|^2, or the caller's radius
    uint32_t far_v = 0;                      // the vertex that attains max|v|^2
    // the staged ring item: taken from the ring (and its position requested) ahead of its use
    bool staged = false;
    uint32_t st_item = 0;
    double sx = 0, sy = 0, sz = 0;
    double rx = 0, ry = 0, rz = 0, r2 = 0;   // the candidate handed to the classification
    uint32_t cand_slot = 0;
    Plane pl = {0, 0, 0, 0};
    unsigned long long in = 0, out = 0;
    uint32_t c_nv = 0;

    int idle = 0;
    const long long t_k0 = TCLK();
    for (;;) {
        TSTAT_LEADER(0, 1);
        // ---- claim the next cells (one atomic per warp) and set up their start polyhedra -------------
        const uint32_t need = __ballot_sync(TFULL, state == S_NEW);
        if (need) {
            const int leader = __ffs((int)need) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(P.work_counter, (uint32_t)__popc(need));
            base = __shfl_sync(TFULL, base, leader);
            if (state == S_NEW) {
                work = base + (uint32_t)__popc(need & lt);
                if (work >= P.n_work) {
                    state = S_EXIT;
                    st_release(&S->cell[c], C_EXIT);
                } else {
                    // the cell's particle (Diagram::get_cell_at_index, interface.rs:193-207)
                    self_slot = P.work_slots ? P.work_slots[work] : P.slot_begin + work;
                    const double2* q = reinterpret_cast<const double2*>(P.sorted + self_slot);
                    const double2 a = __ldg(q), b = __ldg(q + 1);
                    px = a.x; py = a.y; pz = b.x;
                    M.build_cube(P.box, px, py, pz);
                    status = 0;
                    failed = false;
                    const double rmax2 = M.max_radius_sq(far_v);
                    // security mode compares table keys AND |r|^2 with 4*max|v|^2; reference-radius mode compares table
                    // keys with the caller's radius (celery.rs:1036) and rejects nothing
                    stop_thr = radius_mode ? P.search_radius : mul(4.0, rmax2);
                    st_volatile_f64(&S->thr[c], stop_thr);
                    st_release(&S->cell[c], self_slot);  // the producer starts this cell's walk
                    state = S_FETCH;
                }
            }
        }
        if (__all_sync(TFULL, state == S_EXIT)) {
            TSTAT_LEADER(24, TCLK() - t_k0);
            break;
        }

        // ---- stage the next ring item; a particle's position is requested now and used a phase later ---------
        if (!staged && (state == S_FETCH || state == S_CUT)) {
            if (ld_acquire(&S->tail[c]) != head) {
                st_item = S->q[head % (uint32_t)ThreadCfg::QD][c];
                ++head;
                st_release(&S->head[c], head);
                staged = true;
                if (st_item < Q_HALO) {
                    const double2* cq = reinterpret_cast<const double2*>(P.sorted + st_item);
                    const double2 a = __ldg(cq);
                    sx = a.x; sy = a.y;
                    sz = __ldg(reinterpret_cast<const double*>(cq + 1));
                }
            }
        }

        // ---- which phases run: a phase costs the same whatever the number of lanes in it, so each waits for enough
        //      lanes — unless nothing else can happen
        const uint32_t m_test = __ballot_sync(TFULL, state == S_FETCH && staged);
        const uint32_t m_cut = __ballot_sync(TFULL, state == S_CUT);
        const uint32_t m_done = __ballot_sync(TFULL, state == S_DONE);
        const uint32_t m_starved = __ballot_sync(TFULL, state == S_FETCH && !staged);
        const int n_test = __popc(m_test), n_cut = __popc(m_cut), n_done = __popc(m_done);
        bool run_test = n_test >= TESS_T_TEST_MIN, run_cut = n_cut >= TESS_T_CUT_MIN, run_done = n_done >= TESS_T_DONE_MIN;
        if (!run_test && !run_cut && !run_done) {
            if ((n_test | n_cut | n_done) == 0 || (m_starved && idle < TESS_T_SPINS)) {
                ++idle;
                TSTAT_LEADER(7, 1);
                __nanosleep(20);
                continue;
            }
            // the largest group goes (a results phase is the longest: it goes last)
            if (n_test >= n_cut && n_test > 0) run_test = true;
            else if (n_cut > 0) run_cut = true;
            else run_done = true;
        }
        idle = 0;

        // ---- take the staged candidates (interface.rs:280-312 in the producer's order) and classify every live vertex
        //      against each one's bisector plane (find_outgoing_edge's vertex scan, polyhedron.rs:399-405, and every later
        //      vector_location call of the walk)
        const long long t_ph0 = TCLK();
        if (run_test) {
            TSTAT_LEADER(3, 1); TSTAT_LEADER(4, n_test);
            if (state == S_FETCH && staged) {
                staged = false;
                TSTAT(11, 1);
                if (st_item >= Q_END_EXH) {
                    // the end of the walk; a table that ended before a key exceeded the threshold has to be widened
                    // (keys ascend: no key exceeded it iff the last one does not)
                    if (st_item == Q_END_EXH && !failed && !(last_key > stop_thr)) {
                        status |= ST_TABLE_EXHAUSTED;
                        failed = true;
                    }
                    state = S_DONE;
                } else if (failed) {
                    // a cell that was handed back only drains its ring
                } else if (st_item >= Q_HALO) {
                    // reached by the reference's walk iff its key is within the threshold as of now (keys ascend, thresholds shrink)
                    if (radius_mode || !(P.table[st_item & 0x7FFFFFFFu].key > stop_thr)) status |= ST_HALO_INSUFFICIENT;
                } else {
                    rx = subd(sx, px); ry = subd(sy, py); rz = subd(sz, pz);  // interface.rs:322-326: search point - position
                    r2 = dot3(rx, ry, rz, rx, ry, rz);
                    if (radius_mode || r2 < stop_thr) {  // else: cannot have a vertex Outside (header of clip.cu)
                        cand_slot = st_item;
                        state = S_TEST;
                    }
                }
            }
            if (state == S_TEST) {
                {
                    // Plane::halfway_from_origin_to (vector3.rs:223-225); mag_sq(rel) is r2
                    const double m = __dsqrt_rn(r2);
                    const double inv = __ddiv_rn(1.0, m);
                    pl.nx = mul(rx, inv); pl.ny = mul(ry, inv); pl.nz = mul(rz, inv);
                    pl.off = dot3(pl.nx, pl.ny, pl.nz, mul(rx, 0.5), mul(ry, 0.5), mul(rz, 0.5));
                }
                uint32_t in_lo = 0, in_hi = 0, out_lo = 0, out_hi = 0;
                const int top = 64 - __clzll((long long)M.vlive);  // slots above the highest live one are not read
#pragma unroll
                for (int j = 0; j < ThreadCfg::V; ++j) {
                    if ((j & 3) == 0 && j >= top) break;
                    const double sd = signed_distance(pl, M.t->vx[j][lane], M.t->vy[j][lane], M.t->vz[j][lane]);
                    if (j < 32) {
                        if (sd < -TESS_TOL) in_lo |= 1u << j;   // vector3.rs:173
                        if (sd > TESS_TOL) out_lo |= 1u << j;   // vector3.rs:171
                    } else {
                        if (sd < -TESS_TOL) in_hi |= 1u << (j - 32);
                        if (sd > TESS_TOL) out_hi |= 1u << (j - 32);
                    }
                }
                in = (((unsigned long long)in_hi << 32) | in_lo) & M.vlive;  // dead slots hold stale coordinates
                out = (((unsigned long long)out_hi << 32) | out_lo) & M.vlive;
                if (out == 0ull) {
                    state = S_FETCH;  // polyhedron.rs:408-410: no cut
                } else if ((M.vlive & ~in & ~out) != 0ull) {
                    // a vertex ON the plane: the reference destroys it and re-creates it as a copy (polyhedron.rs:555-565),
                    // after which vertices are no longer 3-valent — the warp-per-cell kernel's serial walk does that
                    status |= ST_TABLE_EXHAUSTED;
                    failed = true;
                    st_volatile_f64(&S->thr[c], -2.0);
                    state = S_FETCH;  // drains the ring up to the end marker
                } else {
                    state = S_CUT;
                }
            }
        }

        const long long t_ph1 = TCLK();
        TSTAT_LEADER(21, t_ph1 - t_ph0);
        // ---- cut ---------------------------------------------------------------------------------------
        if (run_cut) {
            TSTAT_LEADER(1, 1); TSTAT_LEADER(2, n_cut);
            if (state == S_CUT) {
                const int rc = thread_cut(M, pl, cand_slot, in, out, c_nv);
                if (rc != TCUT_OK) {
                    status |= ST_TABLE_EXHAUSTED;  // handed back: redone by the warp-per-cell kernel
                    failed = true;
                    st_volatile_f64(&S->thr[c], -2.0);  // every key exceeds it: the producer sends the end marker
                } else if (!radius_mode && ((out >> far_v) & 1ull)) {
                    // the farthest vertex only ever moves inwards: new vertices lie between an Outside and an Inside
                    // one, so max|v|^2 changes only when the vertex that attained it was cut off
                    stop_thr = mul(4.0, M.max_radius_sq(far_v));
                    st_volatile_f64(&S->thr[c], stop_thr);
                }
                state = S_FETCH;
            }
        }

        const long long t_ph2 = TCLK();
        TSTAT_LEADER(22, t_ph2 - t_ph1);
        // ---- results: weighted normals, areas, volume, neighbours ------------------------------------
        if (run_done) {
            TSTAT_LEADER(5, 1); TSTAT_LEADER(6, n_done);
            if (state == S_DONE) {
                const long long self_id = __double_as_longlong(__ldg(reinterpret_cast<const double*>(P.sorted + self_slot) + 3));
                const size_t row = P.row_of_slot ? P.row_of_slot[self_slot] : (size_t)(self_slot - P.row_base);
                const size_t srow = P.stage_by_work ? (size_t)work : row;
                uint32_t nf = 0;
                double vol = 0.0;
                if (!failed) {
                    nf = (uint32_t)__popc(M.flive);
                    uint32_t rank = 0;
                    for (uint32_t fm = M.flive; fm; fm &= fm - 1u) {
                        const uint32_t f = (uint32_t)__ffs((int)fm) - 1u;
                        // Polyhedron::weighted_normal (polyhedron.rs:776-808)
                        const uint32_t s = M.t->fstart[f][lane];
                        uint32_t w = M.ew(s);
                        const uint32_t av = ew_tgt(w);
                        const Vec3 A = {M.x(av), M.y(av), M.z(av)};
                        uint32_t e = ew_next(w);
                        w = M.ew(e);
                        uint32_t tv = ew_tgt(w);
                        Vec3 cu = sub(Vec3{M.x(tv), M.y(tv), M.z(tv)}, A);
                        e = ew_next(w);
                        Vec3 wn = {0.0, 0.0, 0.0};
                        int guard = 0;
                        while (e != s && guard++ < ThreadCfg::E) {
                            w = M.ew(e);
                            tv = ew_tgt(w);
                            const Vec3 prev = cu;
                            cu = sub(Vec3{M.x(tv), M.y(tv), M.z(tv)}, A);
                            wn = add(wn, cross(prev, cu));
                            e = ew_next(w);
                        }
                        // volume = volume + dot(...) face after face in ascending slot order (polyhedron.rs:843-850)
                        vol = addd(vol, dot(A, wn));
                        if (rank < P.fstride) {
                            const uint32_t nb = M.t->fnbr[f][lane];
                            long long id;
                            if (nb >= WALL0) id = -(long long)(nb - WALL0 + 1u);  // container faces: -1..-6
                            else id = __double_as_longlong(__ldg(reinterpret_cast<const double*>(P.sorted + nb) + 3));
                            P.st_nbr[srow * P.fstride + rank] = id;
                            if (P.st_area) P.st_area[srow * P.fstride + rank] = mul(0.5, __dsqrt_rn(dot(wn, wn)));  // interface.rs:408-410
                        }
                        ++rank;
                    }
                    if (nf > P.fstride) {  // cannot happen with F <= fstride; kept for a caller with a smaller staging stride
                        status |= ST_TABLE_EXHAUSTED;
                        failed = true;
                    }
                }
                if (failed && P.failed_slots) {
                    const uint32_t k = atomicAdd(P.n_failed, 1u);
                    if (k < P.failed_cap) P.failed_slots[k] = self_slot;
                    atomicAdd(P.n_failed + 4, 1u);  // "only ran out of table": the next tier is the warp-per-cell kernel, not the medium one
                }
                P.vol[row] = failed ? 0.0 : __ddiv_rn(vol, 6.0);  // polyhedron.rs:854
                P.nfaces[row] = failed ? 0u : nf;
                P.status[row] = status | (P.mark_large ? ST_LARGE_PATH : 0u);
                if (P.cell_id) P.cell_id[row] = self_id;
                state = S_NEW;
            }
        }
        TSTAT_LEADER(23, TCLK() - t_ph2);
    }
}

void launch_thread_cfg(const ClipParams& p, cudaStream_t s) {
    if (!p.n_work) return;
    const size_t smem = sizeof(ThreadShared);
    const int threads = 32 * (ThreadCfg::WARPS + ThreadCfg::PWARPS);
#ifdef TESS_WARP_EMU
    *p.work_counter = 0u;
    emu_launch_kernel([](const void* a) { clip_thread_kernel(*static_cast<const ClipParams*>(a)); }, &p, threads, smem);
#else
    int dev = 0, sms = 0;
    TESS_CUDA_CHECK(cudaGetDevice(&dev));
    TESS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // (set on every launch: the attribute is per device and this is one driver call next to a multi-millisecond kernel)
    TESS_CUDA_CHECK(cudaFuncSetAttribute(clip_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: one CTA per SM (its tables fill the SM's shared memory)
    const unsigned int want = (unsigned int)((p.n_work + ThreadCfg::WARPS * 32 - 1) / (ThreadCfg::WARPS * 32));
    const unsigned int grid = std::min<unsigned int>(want, (unsigned int)sms);
    TESS_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s));
#ifdef TESS_T_STATS
    {
        unsigned long long z[32] = {0};
        TESS_CUDA_CHECK(cudaMemcpyToSymbol(g_tstats, z, sizeof(z)));
    }
#endif
    clip_thread_kernel<<<grid, threads, smem, s>>>(p);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
#ifdef TESS_T_STATS
    {
        unsigned long long h[32];
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        TESS_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_tstats, sizeof(h)));
        const double nc = (double)p.n_work;
        std::fprintf(stderr, "[thread stats] cells %u | per consumer warp-cell-batch(32 cells): rounds %.1f cut phases %.1f (lanes %.1f) test phases %.1f (lanes %.1f) done phases %.1f (lanes %.1f)\n",
                     p.n_work, h[0] / nc * 32, h[1] / nc * 32, h[2] / (double)(h[1] ? h[1] : 1), h[3] / nc * 32, h[4] / (double)(h[3] ? h[3] : 1), h[5] / nc * 32, h[6] / (double)(h[5] ? h[5] : 1));
        std::fprintf(stderr, "[thread stats] lane-rounds per round: starved %.2f waiting-cut %.2f waiting-done %.2f exited %.2f | items consumed per cell %.1f\n",
                     h[7] / (double)h[0], h[8] / (double)h[0], h[9] / (double)h[0], h[10] / (double)h[0], h[11] / nc);
        std::fprintf(stderr, "[thread stats] consumer warp cycles per 32 cells: total %.0f test %.0f cut %.0f results %.0f | per phase: test %.0f cut %.0f results %.0f | producer cycles per step %.0f\n",
                     h[24] / nc * 32, h[21] / nc * 32, h[22] / nc * 32, h[23] / nc * 32, h[21] / (double)(h[3] ? h[3] : 1), h[22] / (double)(h[1] ? h[1] : 1), h[23] / (double)(h[5] ? h[5] : 1), h[19] / (double)(h[13] ? h[13] : 1));
        std::fprintf(stderr, "[thread stats] producer: steps per cell %.2f items/step %.2f particles+markers/step %.1f room/step %.2f | polls %.3g with work %.3g\n",
                     h[13] / nc, h[14] / (double)(h[13] ? h[13] : 1), h[15] / (double)(h[13] ? h[13] : 1), h[16] / (double)(h[13] ? h[13] : 1), (double)h[17], (double)h[18]);
    }
#endif
#endif
}

}  // namespace

// Preconditions (the host checks them, capi.cu): cells of the diagram's own particles (no query positions), no
// geometry output, no work counters.  fstride >= 24.
void launch_clip_thread(const ClipParams& p, cudaStream_t s) { launch_thread_cfg(p, s); }
uint32_t clip_thread_vmax() { return ThreadCfg::V; }
uint32_t clip_thread_fmax() { return ThreadCfg::F; }

}  // namespace tess
