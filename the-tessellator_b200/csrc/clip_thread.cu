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
//   * Tables are sized for the common cell (45 vertices / 140 half-edges / 24 faces = 1805 B per thread:
//     128 threads fill the SM's 227 KB).  A cell that outgrows them, meets a vertex ON a plane (the
//     reference's Incident case, polyhedron.rs:555-565) or runs out of search table is handed back through
//     the failed-cell list and redone by the warp-per-cell kernel, like every other tier's leftovers.
//   * Divergence is managed, not avoided: each lane is a small state machine (fetch a candidate / classify /
//     cut / results) and the warp runs a phase when enough lanes wait for it (cuts and results are batched,
//     fetch + classify run for whoever needs them).
//   * arithmetic is tess_math.cuh's, operation for operation the reference's.
#include <algorithm>

#include "common.cuh"
#include "cube_tables.cuh"
#include "tess_math.cuh"

namespace tess {

namespace {

constexpr uint32_t TFULL = 0xffffffffu;

#ifndef TESS_T_CUT_MIN
#define TESS_T_CUT_MIN 24  // lanes with a cut pending before the cut phase runs (unless nobody can do anything else)
#endif
#ifndef TESS_T_DONE_MIN
#define TESS_T_DONE_MIN 8  // finished cells waiting before the results phase runs
#endif

struct ThreadCfg {
    static constexpr int V = 45, E = 140, F = 24;
    static constexpr int WARPS = 4;
    static constexpr uint32_t NONE = 0xFFu;
};

// The tables of the 32 cells of one warp, lane-interleaved.
struct __align__(16) ThreadTables {
    double vx[ThreadCfg::V][32], vy[ThreadCfg::V][32], vz[ThreadCfg::V][32];
    uint32_t edge[ThreadCfg::E][32];   // {next, flip, target, face}; a free slot holds {next free slot, NONE, NONE, NONE}
    uint32_t fnbr[ThreadCfg::F][32];   // Face.point_index as the neighbour's sorted slot; WALL0 + k for container face k
    uint8_t vedge[ThreadCfg::V][32];   // one half-edge that starts at the vertex (the others: next(flip(e)) twice)
    uint8_t fstart[ThreadCfg::F][32];  // Face.starting_edge_index; free face slots are chained through it
};
static_assert(sizeof(ThreadTables) * ThreadCfg::WARPS <= 232448, "four warps of tables must fit the 227 KB of one SM");

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
// The kernel: persistent warps; every lane pulls cells from the work counter.
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(ThreadCfg::WARPS * 32, 1) clip_thread_kernel(const ClipParams P) {
#ifdef TESS_WARP_EMU  // tests/emu: this source run lane by lane on the CPU (test infrastructure only)
    unsigned char* smem_raw = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    TMesh M;
    M.t = reinterpret_cast<ThreadTables*>(smem_raw) + (threadIdx.x >> 5);
    M.lane = lane;
    M.vlive = 0ull;
    M.flive = 0u;
    M.e_head = M.f_head = ThreadCfg::NONE;
    M.e_hwm = M.e_nfree = M.f_hwm = 0u;

    const GridSpec& G = P.grid;
    const int cpd = (int)G.cpd;
    const bool radius_mode = !(P.search_radius != P.search_radius);  // not NaN
    unsigned long long t_vis = 0, t_test = 0, t_vc = 0, t_cuts = 0, t_nv = 0, t_tab = 0, t_faces = 0;

    // ---- per-lane state of the cell under construction ------------------------------------------
    int state = S_NEW;
    uint32_t work = 0, self_slot = 0xFFFFFFFFu, status = 0;
    bool failed = false;
    double px = 0, py = 0, pz = 0;
    int hx = 0, hy = 0, hz = 0;
    uint32_t ti = 0, cur = 0, end = 0;       // next search_order entry; particles of the current entry still to visit
    double stop_thr = 0.0;                   // 4 * max|v|^2, or the caller's radius
    uint32_t far_v = 0;                      // the vertex that attains max|v|^2
    double rx = 0, ry = 0, rz = 0, r2 = 0;   // the candidate handed to the classification
    uint32_t cand_slot = 0;
    Plane pl = {0, 0, 0, 0};
    unsigned long long in = 0, out = 0;
    uint32_t c_vis = 0, c_test = 0, c_vc = 0, c_cuts = 0, c_nv = 0, c_tab = 0;

    for (;;) {
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
                } else {
                    // the cell's particle (Diagram::get_cell_at_index, interface.rs:193-207)
                    self_slot = P.work_slots ? P.work_slots[work] : P.slot_begin + work;
                    const double2* q = reinterpret_cast<const double2*>(P.sorted + self_slot);
                    const double2 a = __ldg(q), b = __ldg(q + 1);
                    px = a.x; py = a.y; pz = b.x;
                    M.build_cube(P.box, px, py, pz);
                    // ExpandingSearch::new (celery.rs:882-902): home cell of the position
                    hx = (int)axis_index(px, G.xmin, G.xmax, G.ix, G.cpd);
                    hy = (int)axis_index(py, G.ymin, G.ymax, G.iy, G.cpd);
                    hz = (int)axis_index(pz, G.zmin, G.zmax, G.iz, G.cpd);
                    status = 0;
                    failed = false;
                    ti = 0; cur = 0; end = 0;
                    const double rmax2 = M.max_radius_sq(far_v);
                    // security mode compares table keys AND |r|^2 with 4*max|v|^2; reference-radius mode compares table
                    // keys with the caller's radius (celery.rs:1036) and rejects nothing
                    stop_thr = radius_mode ? P.search_radius : mul(4.0, rmax2);
                    c_vis = c_test = c_vc = c_cuts = c_nv = c_tab = 0;
                    state = S_FETCH;
                }
            }
        }
        if (__all_sync(TFULL, state == S_EXIT)) break;

        // ---- fetch: walk the search order until a candidate has to be tested (celery.rs:981-1014 /
        //      interface.rs:280-312) ------------------------------------------------------------------
        if (state == S_FETCH) {
            for (;;) {
                if (cur < end) {
                    const uint32_t slot = cur++;
                    if (COUNT) ++c_vis;
                    if (slot == self_slot) continue;  // interface.rs:283/301 (by index)
                    if (P.target_group != -1)         // interface.rs:284/293 (-2: no particle carries the requested group)
                        if (!(P.target_group >= 0 && P.groups_sorted[slot] == (uint64_t)P.target_group)) continue;
                    const double2* cq = reinterpret_cast<const double2*>(P.sorted + slot);
                    const double2 a = __ldg(cq);
                    const double zz = __ldg(reinterpret_cast<const double*>(cq + 1));
                    // interface.rs:322-326: search point - position
                    rx = subd(a.x, px); ry = subd(a.y, py); rz = subd(zz, pz);
                    r2 = dot3(rx, ry, rz, rx, ry, rz);
                    if (!radius_mode && r2 >= stop_thr) continue;  // cannot have a vertex Outside (header of clip.cu)
                    cand_slot = slot;
                    if (COUNT) ++c_test;
                    state = S_TEST;
                    break;
                }
                if (ti >= P.table_len) {
                    if (!P.table_full && !radius_mode) {
                        status |= ST_TABLE_EXHAUSTED;
                        failed = true;
                    }
                    state = S_DONE;
                    break;
                }
                const ShellEntry e = P.table[ti];
                if (e.key > stop_thr) {  // the walk stops at the first entry whose key exceeds the threshold (celery.rs:1036)
                    state = S_DONE;
                    break;
                }
                ++ti;
                if (COUNT) ++c_tab;
                const int gx = hx + e.di, gy = hy + e.dj, gz = hz + e.dk;
                if (gx < 0 || gx >= cpd || gy < 0 || gy >= cpd || gz < 0 || gz >= cpd) continue;
                if (gx < (int)G.local_lo || gx >= (int)G.local_hi) {
                    status |= ST_HALO_INSUFFICIENT;  // a plane this rank does not hold
                    continue;
                }
                const uint32_t c = ((uint32_t)(gx - (int)G.local_lo) * G.cpd + (uint32_t)gy) * G.cpd + (uint32_t)gz;
                cur = __ldg(P.delim + c);
                end = __ldg(P.delim + c + 1);
            }
        }

        // ---- classify every live vertex against the candidate's bisector plane (find_outgoing_edge's vertex scan,
        //      polyhedron.rs:399-405, and every later vector_location call of the walk) ------------------
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
            if (COUNT) c_vc += (uint32_t)__popcll(M.vlive);
            if (out == 0ull) {
                state = S_FETCH;  // polyhedron.rs:408-410: no cut
            } else if ((M.vlive & ~in & ~out) != 0ull) {
                // a vertex ON the plane: the reference destroys it and re-creates it as a copy (polyhedron.rs:555-565),
                // after which vertices are no longer 3-valent — the warp-per-cell kernel's serial walk does that
                status |= ST_TABLE_EXHAUSTED;
                failed = true;
                state = S_DONE;
            } else {
                state = S_CUT;
            }
        }

        // ---- cut: when enough lanes wait for it, or nobody can do anything else --------------------------
        {
            const uint32_t m_cut = __ballot_sync(TFULL, state == S_CUT);
            const uint32_t m_fetch = __ballot_sync(TFULL, state == S_FETCH);
            if (m_cut && (__popc(m_cut) >= TESS_T_CUT_MIN || m_fetch == 0u)) {
                if (state == S_CUT) {
                    const int rc = thread_cut(M, pl, cand_slot, in, out, c_nv);
                    if (rc != TCUT_OK) {
                        status |= ST_TABLE_EXHAUSTED;  // handed back: redone by the warp-per-cell kernel
                        failed = true;
                        state = S_DONE;
                    } else {
                        ++c_cuts;
                        // the farthest vertex only ever moves inwards: new vertices lie between an Outside and an Inside
                        // one, so max|v|^2 changes only when the vertex that attained it was cut off
                        if (!radius_mode && (COUNT || ((out >> far_v) & 1ull))) stop_thr = mul(4.0, M.max_radius_sq(far_v));
                        state = S_FETCH;
                    }
                }
            }
        }

        // ---- results: weighted normals, areas, volume, neighbours ------------------------------------
        {
            const uint32_t m_done = __ballot_sync(TFULL, state == S_DONE);
            const uint32_t m_busy = __ballot_sync(TFULL, state == S_FETCH || state == S_CUT);
            if (m_done && (__popc(m_done) >= TESS_T_DONE_MIN || m_busy == 0u)) {
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
                    if (COUNT && !failed) {  // only cells this pass finished are counted; the others are counted by the redo pass
                        t_vis += c_vis; t_test += c_test; t_vc += c_vc; t_cuts += c_cuts; t_nv += c_nv; t_tab += c_tab; t_faces += nf;
                    }
                    state = S_NEW;
                }
            }
        }
    }

    if (COUNT && P.counters) {
        atomicAdd(&P.counters[CNT_VISITED], t_vis);
        atomicAdd(&P.counters[CNT_TESTED], t_test);
        atomicAdd(&P.counters[CNT_VC], t_vc);
        atomicAdd(&P.counters[CNT_CUTS], t_cuts);
        atomicAdd(&P.counters[CNT_NV], t_nv);
        atomicAdd(&P.counters[CNT_TABLE], t_tab);
        atomicAdd(&P.counters[CNT_FACES], t_faces);
    }
}

template <bool COUNT>
void launch_thread_cfg(const ClipParams& p, cudaStream_t s) {
    if (!p.n_work) return;
    const size_t smem = sizeof(ThreadTables) * ThreadCfg::WARPS;
#ifdef TESS_WARP_EMU
    *p.work_counter = 0u;
    emu_launch_kernel([](const void* a) { clip_thread_kernel<COUNT>(*static_cast<const ClipParams*>(a)); }, &p, ThreadCfg::WARPS * 32, smem);
#else
    int dev = 0, sms = 0;
    TESS_CUDA_CHECK(cudaGetDevice(&dev));
    TESS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // (set on every launch: the attribute is per device and this is one driver call next to a multi-millisecond kernel)
    TESS_CUDA_CHECK(cudaFuncSetAttribute(clip_thread_kernel<COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: one CTA of four warps per SM (its tables fill the SM's shared memory)
    const unsigned int want = (unsigned int)((p.n_work + ThreadCfg::WARPS * 32 - 1) / (ThreadCfg::WARPS * 32));
    const unsigned int grid = std::min<unsigned int>(want, (unsigned int)sms);
    TESS_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s));
    clip_thread_kernel<COUNT><<<grid, ThreadCfg::WARPS * 32, smem, s>>>(p);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
#endif
}

}  // namespace

// Preconditions (the host checks them, capi.cu): cells of the diagram's own particles (no query positions), no
// geometry output.  fstride >= 24.
void launch_clip_thread(const ClipParams& p, cudaStream_t s) {
    if (p.counters) launch_thread_cfg<true>(p, s); else launch_thread_cfg<false>(p, s);
}
uint32_t clip_thread_vmax() { return ThreadCfg::V; }
uint32_t clip_thread_fmax() { return ThreadCfg::F; }

}  // namespace tess
