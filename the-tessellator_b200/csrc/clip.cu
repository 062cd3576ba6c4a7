// clip.cu — K5: one warp builds one Voronoi cell.
//
// Replaces, per cell:
//   Cell::compute_voronoi_cell / cut_with_point          interface.rs:257-334
//   ExpandingSearch::{new, expand_all_*}                  celery.rs:882-1075
//   Polyhedron::{find_outgoing_edge, cut_with_plane}      polyhedron.rs:396-642
//   Polyhedron::{weighted_normal, compute_volume, compute_neighbors}   polyhedron.rs:776-881
//   VoronoiFace::compute_area                             interface.rs:408-410
//
// Design:
//   * the polyhedron is a half-edge mesh held in SHARED MEMORY (fixed-capacity tables) instead of pool.rs's heap
//     pools; one half-edge = one 32-bit word {next, flip, target, face} of 8-bit slot ids (64-bit words of 16-bit ids
//     in the medium / large configurations).  Half-edge and face slots are handed out and recycled in pool.rs's LIFO
//     order (free stacks in shared memory), so slot numbers — and with them find_outgoing_edge's "first edge in slot
//     order", every face's starting edge, the face order, the fan anchors and the summation orders — are the
//     reference's: volumes, areas and the neighbour order come out bit-identical to the CPU oracle;
//   * candidates are staged 32 at a time: one lane per search-table entry reads the grid delimiters, a warp scan
//     flattens the per-cell ranges, then one lane per candidate loads its 32-byte particle record, evaluates |r|^2
//     and builds its bisector plane in parallel; a tile with enough planes waiting is screened lane-per-plane first;
//   * plane-side classification is one lane per vertex + warp ballots;
//   * the cut is lane-parallel when no vertex lies on the plane (cut_parallel: one lane per half-edge leaving an
//     Outside vertex, crossings linked into the reference's walk order by list ranking) and the reference-shaped
//     serial walk, warp-uniform, otherwise; the main pass is the instantiation WITHOUT the walk (SERIAL = false):
//     cells that need it are handed back and finished by the instantiation that has it;
//   * one warp per CTA in the small configuration: the tables sit at the CTA's shared-memory base, a constant;
//   * arithmetic is the reference's, operation for operation (tess_math.cuh, no FMA contraction), so vertex
//     coordinates — and therefore every Inside/Incident/Outside decision — are bit-identical to the CPU oracle's;
//   * NEW vs the reference: the shell walk stops at the first table entry whose key exceeds
//     4*max|v|^2 and skips candidates with |r|^2 >= 4*max|v|^2.  Such candidates cannot have a
//     vertex Outside (n.v - |r|/2 <= |v| - |r|/2 <= 0 < tol), i.e. find_outgoing_edge
//     (polyhedron.rs:399-410) would have returned None for them: results are unchanged.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "cube_tables.cuh"
#include "tess_math.cuh"

namespace tess {

namespace {

constexpr uint32_t FULL = 0xffffffffu;

// A warp-uniform region: code that all 32 lanes of the (converged) warp execute redundantly on the same
// shared-memory state — every lane reads the same words and writes the same values.  The markers compile
// to nothing here; the CPU warp emulator of the test suite (tests/emu) uses them to give each lane the
// state the first lane found and to check that all lanes leave the region with identical state.
#ifndef TESS_UNIFORM_BEGIN
#define TESS_UNIFORM_BEGIN(ptr, bytes) ((void)0)
#define TESS_UNIFORM_END() ((void)0)
#endif

// ---------------------------------------------------------------------------------------------
// Bit masks over table slots.  Small cells keep them in registers, large cells in shared memory.
// ---------------------------------------------------------------------------------------------
template <int NW>
struct RegMask {
    uint32_t w[NW];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] = 0u;
    }
    __device__ __forceinline__ uint32_t word(int k) const { return w[k]; }
    __device__ __forceinline__ void set_word(int k, uint32_t v) { w[k] = v; }
    __device__ __forceinline__ bool test(uint32_t i) const {
        uint32_t x = w[0];
#pragma unroll
        for (int k = 1; k < NW; ++k) x = ((i >> 5) == (uint32_t)k) ? w[k] : x;
        return (x >> (i & 31u)) & 1u;
    }
    __device__ __forceinline__ void set(uint32_t i) {
#pragma unroll
        for (int k = 0; k < NW; ++k) w[k] |= ((i >> 5) == (uint32_t)k) ? (1u << (i & 31u)) : 0u;
    }
    // lowest clear bit below `limit` (and set it), or -1
    __device__ __forceinline__ int alloc(int limit) {
        int slot = -1;
#pragma unroll
        for (int k = NW - 1; k >= 0; --k) {
            const uint32_t fr = ~w[k];
            if (fr) slot = 32 * k + __ffs(fr) - 1;
        }
        if (slot >= limit) slot = -1;
        if (slot >= 0) set((uint32_t)slot);
        return slot;
    }
};

// 64 slots: one 64-bit register pair, native 64-bit shifts instead of select chains
template <>
struct RegMask<2> {
    unsigned long long m;
    __device__ __forceinline__ void clear() { m = 0ull; }
    __device__ __forceinline__ uint32_t word(int k) const { return (uint32_t)(m >> (32 * k)); }
    __device__ __forceinline__ void set_word(int k, uint32_t v) { m = (m & ~(0xFFFFFFFFull << (32 * k))) | ((unsigned long long)v << (32 * k)); }
    __device__ __forceinline__ bool test(uint32_t i) const { return (m >> i) & 1ull; }
    __device__ __forceinline__ void set(uint32_t i) { m |= 1ull << i; }
    __device__ __forceinline__ int alloc(int limit) {
        const int slot = __ffsll((long long)~m) - 1;  // -1 when full
        if (slot < 0 || slot >= limit) return -1;
        m |= 1ull << slot;
        return slot;
    }
};

template <int NW>
struct SmemMask {
    uint32_t* w;
    __device__ __forceinline__ void clear() {
        for (int k = 0; k < NW; ++k) w[k] = 0u;
    }
    __device__ __forceinline__ uint32_t word(int k) const { return w[k]; }
    __device__ __forceinline__ void set_word(int k, uint32_t v) { w[k] = v; }
    __device__ __forceinline__ bool test(uint32_t i) const { return (w[i >> 5] >> (i & 31u)) & 1u; }
    __device__ __forceinline__ void set(uint32_t i) { w[i >> 5] |= 1u << (i & 31u); }
    __device__ __forceinline__ int alloc(int limit) {
        for (int k = 0; k < NW; ++k) {
            const uint32_t fr = ~w[k];
            if (fr) {
                const int slot = 32 * k + __ffs(fr) - 1;
                if (slot >= limit) return -1;
                w[k] |= 1u << (slot & 31);
                return slot;
            }
        }
        return -1;
    }
};

// ---------------------------------------------------------------------------------------------
// Configurations
// ---------------------------------------------------------------------------------------------
#ifndef TESS_SMALL_WARPS
#define TESS_SMALL_WARPS 1    // warps per CTA of the small configuration (A/B on 1M uniform at 32 warps per SM: 4 -> 18.94 ms, 2 -> 19.20, 1 -> 17.98)
#endif
#ifndef TESS_CLIP_MINBLOCKS
#define TESS_CLIP_MINBLOCKS (32 / TESS_SMALL_WARPS)  // resident CTAs per SM the small kernel is compiled for: 32 warps per SM, 64 registers (A/B on 1M uniform with 4-warp CTAs: 20 warps -> 20.75 ms, 24 -> 20.05, 28 -> 19.17, 32 -> 18.59, 36 -> 18.94, 40 -> 20.06)
#endif
// A staged tile with at least this many candidates waiting is screened candidate-parallel before the planes
// are offered one by one (0 disables the screen); see the comment at its use.
#ifndef TESS_PREFILTER_MIN
#define TESS_PREFILTER_MIN 6
#endif
struct SmallCfg {
    static constexpr int MINB = TESS_CLIP_MINBLOCKS;
    static constexpr int VMAX = 64, EMAX = 256, FMAX = 64;
    static constexpr int E_LIMIT = 255;  // slot 255 is the "none" marker of 8-bit ids
    static constexpr int WARPS = TESS_SMALL_WARPS;
    static constexpr bool REG = true;
    using Idx = uint8_t;
    using EdgeWord = uint32_t;
    static constexpr uint32_t NONE = 0xFFu;
    static constexpr int SHIFT = 8;
};
// Cells that outgrow the small tables at some point of their construction — mostly cells at the rim of a
// dense cluster, cut by hundreds of planes in arrival order before they shrink to 15-25 faces — fit this
// middle configuration, which keeps ~10 warps per SM resident where the large one keeps 2.
struct MediumCfg {
    static constexpr int MINB = 5;
    static constexpr int VMAX = 256, EMAX = 768, FMAX = 128;
    static constexpr int E_LIMIT = 768;
    static constexpr int WARPS = 2;
    static constexpr bool REG = false;
    using Idx = uint16_t;
    using EdgeWord = unsigned long long;
    static constexpr uint32_t NONE = 0xFFFFu;
    static constexpr int SHIFT = 16;
};
struct LargeCfg {
    static constexpr int MINB = 1;
    static constexpr int VMAX = 1024, EMAX = 3072, FMAX = 512;
    static constexpr int E_LIMIT = 3072;
    static constexpr int WARPS = 1;
    static constexpr bool REG = false;
    using Idx = uint16_t;
    using EdgeWord = unsigned long long;
    static constexpr uint32_t NONE = 0xFFFFu;
    static constexpr int SHIFT = 16;
};

template <class Cfg>
struct WarpSmem {
    double vx[Cfg::VMAX], vy[Cfg::VMAX], vz[Cfg::VMAX];
    long long fnbr[Cfg::FMAX];
    typename Cfg::EdgeWord edge[Cfg::EMAX];  // {next, flip, target, face}
    typename Cfg::Idx fstart[Cfg::FMAX];
    typename Cfg::Idx estack[Cfg::EMAX];     // free half-edge slots (LIFO, like pool.rs's free list)
    typename Cfg::Idx fstack[Cfg::FMAX];     // free face slots (LIFO)
    typename Cfg::EdgeWord xlist[32];        // crossings of the current cut: {outside end, inside end, new vertex, copy flag}
    // scratch of the warp-parallel cut (small configuration only)
    typename Cfg::Idx olist[32];             // outgoing half-edges of the cut, one per crossed face
    typename Cfg::Idx pred[32];              // crossing k -> the crossing that precedes it around the cut
    typename Cfg::Idx vfree[32];             // first free vertex slots
    uint8_t kof[Cfg::EMAX];                  // outgoing half-edge slot -> crossing index
    typename Cfg::Idx ovl[32];               // the (first 32) Outside vertices of the current plane
    uint32_t dlist[8];                       // the dying half-edge slots of the current cut, one byte each
    // While every vertex is 3-valent (no plane has ever passed through a vertex of this cell) vout[3v..3v+2]
    // are the half-edges that start at vertex v: a cut is then found from its Outside vertices, without
    // sweeping the half-edge table.
    typename Cfg::Idx vout[Cfg::REG ? 3 * Cfg::VMAX : 4];
    // candidate tile: bisector plane {n, offset} and id of the candidate each lane staged
    double4 cand_plane[32];
    long long cand_id[32];
    double cpos[4];                          // position of the cell's particle (kept here, not in registers)
    uint16_t flen[Cfg::FMAX];                // geometry output: loop length / loop offset per face slot
    uint16_t floff[Cfg::FMAX];
    // masks of the large configuration live here (1-word placeholders otherwise)
    uint32_t m_vlive[Cfg::REG ? 1 : Cfg::VMAX / 32], m_vbefore[Cfg::REG ? 1 : Cfg::VMAX / 32], m_inside[Cfg::REG ? 1 : Cfg::VMAX / 32],
        m_outside[Cfg::REG ? 1 : Cfg::VMAX / 32], m_removed[Cfg::REG ? 1 : Cfg::VMAX / 32];
    uint32_t m_flive[Cfg::REG ? 1 : Cfg::FMAX / 32], m_fkeep[Cfg::REG ? 1 : Cfg::FMAX / 32];
};

template <class Cfg, int NW>
using MaskT = typename std::conditional<Cfg::REG, RegMask<NW>, SmemMask<NW>>::type;

template <class Cfg>
struct Mesh {
    using Idx = typename Cfg::Idx;
    using EW = typename Cfg::EdgeWord;
    static constexpr int NWV = Cfg::VMAX / 32, NWE = Cfg::EMAX / 32, NWF = Cfg::FMAX / 32;
    static constexpr int SH = Cfg::SHIFT;
    static constexpr EW FM = (EW)Cfg::NONE;  // field mask

    WarpSmem<Cfg>* sm;
    MaskT<Cfg, NWV> vlive, vbefore, inside, outside, removed;
    MaskT<Cfg, NWF> flive, fkeep;
    // Half-edge slots: [0, e_hwm) have been handed out at least once; a free slot holds FREE_EDGE
    // and sits on the LIFO stack sm->estack[0, e_top) — the shape of pool.rs's Pool (free list +
    // append at the end) without a per-slot bitmask.
    int e_top, e_hwm;
    int f_top, f_hwm;  // same for faces: sm->fstack[0, f_top) + high-water mark (flive stays the iteration mask)
    // 32-slot words of the vertex / face tables that have ever been used (large cells sweep only these)
    int v_words, f_words;
    bool simple;         // all vertices 3-valent and vout[] maintained (true until the serial walk first cuts)
    uint32_t n_outside;  // Outside vertices of the last classification
    int lane;
    __device__ __forceinline__ int nwv() const { return Cfg::REG ? NWV : v_words; }
    __device__ __forceinline__ int nwf() const { return Cfg::REG ? NWF : f_words; }
    __device__ __forceinline__ void note_vertex(int slot) { if (!Cfg::REG && (slot >> 5) + 1 > v_words) v_words = (slot >> 5) + 1; }
    __device__ __forceinline__ void note_face(int slot) { if (!Cfg::REG && (slot >> 5) + 1 > f_words) f_words = (slot >> 5) + 1; }
    static constexpr EW FREE_EDGE = ~(EW)0;
    static __device__ __forceinline__ bool e_is_free(EW w) { return e_face(w) == Cfg::NONE; }
    __device__ __forceinline__ int alloc_edge() {
        if (e_top > 0) return (int)sm->estack[--e_top];
        if (e_hwm < Cfg::E_LIMIT) return e_hwm++;
        return -1;
    }
    // Pool::add for faces (pool.rs:85-110): most recently freed slot first, else append
    // uniform_slot: pass the slot read from shared memory through a warp reduction.  Every lane reads the same value
    // anyway; the reduction makes that visible to the compiler, which otherwise treats the face mask — and every
    // branch that depends on it — as lane-dependent and guards each later collective with a divergence check.
    __device__ __forceinline__ int alloc_face(bool uniform_slot = false) {
        int slot;
        if (f_top > 0) {
            slot = (int)sm->fstack[--f_top];
            if (uniform_slot) slot = (int)__reduce_max_sync(FULL, (uint32_t)slot);
        }
        else if (f_hwm < Cfg::FMAX) slot = f_hwm++;
        else return -1;
        flive.set((uint32_t)slot);
        note_face(slot);
        return slot;
    }
    // Pool::remove for every face of word q that lost all its half-edges, in ascending slot order
    __device__ __forceinline__ void retire_faces(int q, uint32_t dead_word) {
        if (dead_word == 0u) return;
        if ((dead_word >> lane) & 1u) sm->fstack[f_top + __popc(dead_word & ((1u << lane) - 1u))] = (Idx)(32 * q + lane);
        f_top += __popc(dead_word);
        __syncwarp();
    }
    // edge word of slot 32*p+lane, FREE_EDGE beyond the high-water mark
    __device__ __forceinline__ EW edge_of_pass(int p) const {
        const int e = 32 * p + lane;
        return e < e_hwm ? sm->edge[e] : FREE_EDGE;
    }
    __device__ __forceinline__ int edge_passes() const { return (e_hwm + 31) >> 5; }

    __device__ __forceinline__ void bind(WarpSmem<Cfg>* s, int lane_) {
        sm = s;
        lane = lane_;
        if constexpr (!Cfg::REG) {
            vlive.w = s->m_vlive; vbefore.w = s->m_vbefore; inside.w = s->m_inside; outside.w = s->m_outside; removed.w = s->m_removed;
            flive.w = s->m_flive; fkeep.w = s->m_fkeep;
        }
    }

    static __device__ __forceinline__ EW pack(uint32_t next, uint32_t flip, uint32_t tgt, uint32_t face) {
        return (EW)next | ((EW)flip << SH) | ((EW)tgt << (2 * SH)) | ((EW)face << (3 * SH));
    }
    static __device__ __forceinline__ uint32_t e_next(EW w) { return (uint32_t)(w & FM); }
    static __device__ __forceinline__ uint32_t e_flip(EW w) { return (uint32_t)((w >> SH) & FM); }
    static __device__ __forceinline__ uint32_t e_tgt(EW w) { return (uint32_t)((w >> (2 * SH)) & FM); }
    static __device__ __forceinline__ uint32_t e_face(EW w) { return (uint32_t)((w >> (3 * SH)) & FM); }
    // field: 0 next, 1 flip, 2 target, 3 face
    __device__ __forceinline__ void set_field(uint32_t e, int field, uint32_t v) { reinterpret_cast<Idx*>(&sm->edge[e])[field] = (Idx)v; }

    // Polyhedron::build_cube (polyhedron.rs:268-392) translated by -p (interface.rs:266).
    __device__ void build_cube(const double* box, double px, double py, double pz) {
        vlive.clear(); flive.clear();
        e_top = 0;
        e_hwm = 24;
        f_top = 0;
        f_hwm = 6;
        v_words = 1;
        f_words = 1;
        simple = Cfg::REG;
        if constexpr (Cfg::REG) {
            if (lane < 24) sm->vout[lane] = (Idx)kCubeVout[lane];
        }
        __syncwarp();
        if (lane < 8) {
            // FDL FDR FUR FUL BDL BDR BUR BUL (polyhedron.rs:288-295); corner + (-p)
            const bool xh = (lane == 1) | (lane == 2) | (lane == 5) | (lane == 6);
            const bool yh = lane >= 4;
            const bool zh = (lane == 2) | (lane == 3) | (lane == 6) | (lane == 7);
            sm->vx[lane] = addd(xh ? box[3] : box[0], -px);
            sm->vy[lane] = addd(yh ? box[4] : box[1], -py);
            sm->vz[lane] = addd(zh ? box[5] : box[2], -pz);
        }
        if (lane < 24) {
            // {flip, target, next} per half-edge, ids of polyhedron.rs:97-199 (DR belongs to face D)
            const uint32_t t = kCubeEdges[lane];
            sm->edge[lane] = pack(t & 0xFFu, (t >> 16) & 0xFFu, (t >> 8) & 0xFFu, (uint32_t)lane >> 2);
        }
        if (lane < 6) {
            sm->fstart[lane] = (Idx)(4 * lane);  // FU RU BU LU UF DF (polyhedron.rs:319-379)
            sm->fnbr[lane] = -(long long)(lane + 1);
        }
        vlive.set_word(0, 0xFFu);
        flive.set_word(0, 0x3Fu);
        __syncwarp();
    }

    // max |v|^2 over live vertices (left-associated dot, like Vector3::mag_sq)
    __device__ double max_radius_sq() const {
        double m = 0.0;
#pragma unroll
        for (int p = 0; p < NWV; ++p) {
            if (p >= nwv()) break;
            const uint32_t lw = vlive.word(p);
            if (lw == 0u) continue;
            if ((lw >> lane) & 1u) {
                const int v = 32 * p + lane;
                const double x = sm->vx[v], y = sm->vy[v], z = sm->vz[v];
                const double r2 = dot3(x, y, z, x, y, z);
                m = r2 > m ? r2 : m;
            }
        }
        // non-negative doubles order like their bit patterns
        const unsigned long long b = (unsigned long long)__double_as_longlong(m);
        const uint32_t hi = __reduce_max_sync(FULL, (uint32_t)(b >> 32));
        const uint32_t lo = __reduce_max_sync(FULL, ((uint32_t)(b >> 32) == hi) ? (uint32_t)b : 0u);
        return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
};

constexpr int CUT_FALLBACK = 3;
constexpr int CUT_NEEDS_SERIAL = -3;

// ---------------------------------------------------------------------------------------------
// Warp-parallel form of Polyhedron::cut_with_plane (polyhedron.rs:438-642) for the generic case:
// no vertex lies on the plane (none is Incident).  Every face crossed by the plane then has exactly
// one outgoing half-edge (Inside -> Outside, polyhedron.rs:491-503) and one re-entering half-edge
// (Outside -> Inside, :529-544); the reference visits these faces one after the other
// (`outgoing = flip(re-entering)`, :603-607).  Here one lane takes one crossed face:
//   1. one sweep over the half-edge table finds the outgoing half-edges, the half-edges that die
//      (both ends Outside) and the faces that keep at least one half-edge;
//   2. each lane follows its face loop to the re-entering half-edge (:529-544);
//   3. `flip(re-entering)` names the next crossed face: the lanes form the cyclic order of the
//      reference's walk; if they do not form ONE cycle through all outgoing half-edges (possible
//      only when rounding breaks convexity) the serial walk takes over;
//   4. each lane creates its intersection vertex (:567-572), its bridge half-edge (:582-587) and
//      the cap half-edge paired with it (:592-598), linked exactly as the serial walk links them.
// The start edge of the cap face is the one of the reference's first crossing (the outgoing
// half-edge whose flip has the lowest slot, :413-432), so anchors and summation orders — hence
// areas and volumes — are identical to the serial walk's.
// Returns 1 (cut), CUT_FALLBACK (let the serial walk decide) or -1 (table overflow).
// ---------------------------------------------------------------------------------------------
template <class Cfg, bool SERIAL>
__device__ int cut_parallel(Mesh<Cfg>& M, const Plane& pl, long long neighbor_id, uint32_t& cnt_nv, bool sweep_only) {
    using MeshT = Mesh<Cfg>;
    using EW = typename Cfg::EdgeWord;
    using Idx = typename Cfg::Idx;
    WarpSmem<Cfg>* sm = M.sm;
    const int lane = M.lane;
    const uint32_t lt = (1u << lane) - 1u;

    // ---- 1-3. the crossings: per crossed face the outgoing half-edge o (Inside -> Outside), the
    //      re-entering half-edge r (pv Outside -> cv Inside), and the lanes ks / pk that hold the next /
    //      previous crossing of the reference's walk --------------------------------------------------
    uint32_t K = 0;
    uint32_t deadbits = 0;              // sweep variant: bit p = my half-edge of pass p dies
    uint32_t keep_lo = 0, keep_hi = 0;  // sweep variant: faces my surviving half-edges belong to
    uint32_t my_dead = Cfg::NONE;       // adjacency variant: the dying half-edge this lane found
    uint32_t df_lo = 0, df_hi = 0;      //                    and the face it belongs to
    int n_pass = 0;
    bool act = false;
    uint32_t o = 0, f = 0, r = 0, pv = 0, cv = 0, fo = 0xFFFFu, ks = 0, pk = 0;
    bool adj = false;
    if constexpr (Cfg::REG) adj = M.simple && !sweep_only && 3u * M.n_outside <= 32u;
    if (adj) {
        // Every vertex is 3-valent and vout[] lists the half-edges leaving it: lane (i, j) takes
        // half-edge j of the i-th Outside vertex.  It either ends Outside too (it dies; so does its
        // flip, which the other vertex's lane finds) or ends Inside — then it is the re-entering
        // half-edge of its face, and its flip is the outgoing half-edge of the next crossed face.
        if constexpr (Cfg::REG) {
            uint32_t so = 0, fs = 0;
            if ((uint32_t)lane < 3u * M.n_outside) {
                const uint32_t vi = (uint32_t)lane / 3u;
                pv = sm->ovl[vi];
                const uint32_t e = sm->vout[3u * pv + ((uint32_t)lane - 3u * vi)];
                const EW w = sm->edge[e];
                const uint32_t t = MeshT::e_tgt(w);
                const uint32_t face = MeshT::e_face(w);
                if (M.outside.test(t)) {
                    my_dead = e;
                    if (face < 32u) df_lo = 1u << face; else df_hi = 1u << (face - 32u);
                } else {
                    act = true;
                    r = e;
                    cv = t;
                    f = face;
                    so = MeshT::e_flip(w);
                    fs = MeshT::e_face(sm->edge[so]);
                    sm->kof[f] = (uint8_t)lane;  // who holds the crossing of face f
                }
            }
            K = __popc(__ballot_sync(FULL, act));
            if (K == 0u) return CUT_FALLBACK;  // SURVEY D17, reported by the serial path
            __syncwarp();
            // successor: the crossing of the face my re-entering half-edge's flip lies in
            if (act) ks = sm->kof[fs] & 31u;
            // one shuffle: an inactive lane answers with a face id nobody has
            const uint32_t s_f = __shfl_sync(FULL, act ? f : 0xFFFFu, (int)ks);
            const bool ok1 = !act || s_f == fs;
            if (act && ok1) sm->pred[ks] = (Idx)lane;
            __syncwarp();
            // predecessor; succ is a bijection of the crossings iff every crossing is its predecessor's successor
            if (act) pk = (uint32_t)sm->pred[lane] & 31u;
            // one shuffle: {successor (0xFF from an inactive lane), flip of the re-entering half-edge, the half-edge}
            const uint32_t from_pred = __shfl_sync(FULL, (act ? ks : 0xFFu) | (so << 8) | (r << 16), (int)pk);
            o = (from_pred >> 8) & 0xFFu;
            fo = act ? (from_pred >> 16) & 0xFFu : 0xFFFFu;
            // (a lane whose successor was wrong wrote no predecessor entry: whatever the others read there, the vote fails)
            if (!__all_sync(FULL, ok1 && (!act || (from_pred & 0xFFu) == (uint32_t)lane))) return CUT_FALLBACK;
        }
    } else {
        // 1. sweep the half-edge table for the outgoing half-edges and for what dies
        n_pass = M.edge_passes();
        if (n_pass > 32) return CUT_FALLBACK;  // deadbits holds one bit per pass
        for (int p = 0; p < n_pass; ++p) {
            const EW w = M.edge_of_pass(p);
            bool is_out = false, dead = false, keep = false;
            uint32_t face = 0;
            if (!MeshT::e_is_free(w)) {
                const uint32_t t = MeshT::e_tgt(w);
                const uint32_t s = MeshT::e_tgt(sm->edge[MeshT::e_flip(w)]);
                const bool tout = M.outside.test(t), sout = M.outside.test(s);
                is_out = tout && !sout;  // source Inside (nothing is Incident)
                dead = tout && sout;
                keep = !dead;
                face = MeshT::e_face(w);
            }
            const uint32_t om = __ballot_sync(FULL, is_out);
            if (is_out) {
                const uint32_t k = K + __popc(om & lt);
                if (k < 32u) {
                    sm->olist[k] = (Idx)(32 * p + lane);
                    sm->kof[32 * p + lane] = (uint8_t)k;
                }
            }
            K += __popc(om);
            deadbits |= (dead ? 1u : 0u) << p;
            if (keep) {
                if constexpr (Cfg::REG) {
                    if (face < 32u) keep_lo |= 1u << face; else keep_hi |= 1u << (face - 32u);
                } else {
                    keep_lo = 1u;  // large cells: the face mask is rebuilt below, only if something died
                }
            }
        }
        if (K == 0u || K > 32u) return CUT_FALLBACK;  // K == 0: SURVEY D17, reported by the serial path
        __syncwarp();
        // 2. one lane per crossed face: walk from the outgoing to the re-entering half-edge
        act = (uint32_t)lane < K;
        uint32_t so = 0;
        bool ok = true;
        if (act) {
            o = sm->olist[lane];
            const EW wo = sm->edge[o];
            f = MeshT::e_face(wo);
            fo = MeshT::e_flip(wo);
            pv = MeshT::e_tgt(wo);  // previous_vertex_index (:491), Outside
            r = MeshT::e_next(wo);  // :506
            EW wc = sm->edge[r];
            cv = MeshT::e_tgt(wc);
            int budget = Cfg::EMAX;
            while (!M.inside.test(cv) && --budget > 0) {  // :529-544
                pv = cv;
                r = MeshT::e_next(wc);
                wc = sm->edge[r];
                cv = MeshT::e_tgt(wc);
            }
            ok = budget > 0;
            so = MeshT::e_flip(wc);  // the next crossed face's outgoing half-edge (:603-607)
        }
        // 3. cyclic order of the crossings (flip is a bijection, so succ is a permutation of them)
        if (act) {
            ks = sm->kof[so];
            ok = ok && ks < K && sm->olist[ks] == (Idx)so;
            if (ok) sm->pred[ks] = (Idx)lane;
        }
        if (!__all_sync(FULL, ok)) return CUT_FALLBACK;
        __syncwarp();
        if (act) pk = (uint32_t)sm->pred[lane];
    }
    // the reference's first crossing: outgoing half-edge whose flip has the lowest slot (:413-432)
    const uint32_t k0 = __reduce_min_sync(FULL, (fo << 8) | (uint32_t)lane) & 0xFFu;
    // position of my crossing in the reference's walk order (k0 is 0): list ranking by pointer
    // jumping — after ceil(log2 K) rounds every lane of k0's cycle knows its distance to k0
    uint32_t wi = 0;
    {
        const bool term = !act || (uint32_t)lane == k0;
        uint32_t nxt = term ? k0 : ks, d = term ? 0u : 1u;
#pragma unroll
        for (int step = 0; step < 5; ++step) {
            if ((1u << step) >= K) break;
            const uint32_t d2 = __shfl_sync(FULL, d, (int)nxt);
            const uint32_t n2 = __shfl_sync(FULL, nxt, (int)nxt);
            d += d2;
            nxt = n2;
        }
        // succ is a permutation of the crossings: everyone reaches k0 <=> one cycle
        if (__any_sync(FULL, act && nxt != k0)) return CUT_FALLBACK;  // not a convex cut
        wi = ((uint32_t)lane == k0) ? 0u : K - d;
    }
    // ---- capacity: K vertices, 2K half-edges, one face ------------------------------------------
    {
        int vfree_n = 32 * (MeshT::NWV - M.nwv());
#pragma unroll
        for (int p = 0; p < MeshT::NWV; ++p) {
            if (p >= M.nwv()) break;
            vfree_n += 32 - __popc(M.vlive.word(p));
        }
        const int efree_n = M.e_top + (Cfg::E_LIMIT - M.e_hwm);
        if ((int)K > vfree_n || 2 * (int)K + 1 > efree_n) return -1;
    }
    const int cap_face = M.alloc_face(/*uniform_slot=*/!SERIAL);
    if (cap_face < 0) return -1;
    // first free vertex slots, in ascending order (what repeated lowest-free-bit allocation yields)
    {
        uint32_t base = 0;
#pragma unroll
        for (int p = 0; p < MeshT::NWV; ++p) {
            if (base >= 32u) break;
            const uint32_t fr = ~M.vlive.word(p);
            if ((fr >> lane) & 1u) {
                const uint32_t rk = base + __popc(fr & lt);
                if (rk < 32u) sm->vfree[rk] = (Idx)(32 * p + lane);
            }
            base += __popc(fr);
        }
    }
    __syncwarp();
    // ---- 4. new vertex, bridge and cap half-edge of every crossing ------------------------------
    // Half-edge slots are handed out in the order of the reference's walk (Pool::add, pool.rs:85-110):
    // cap_first (:478), then per crossing the bridge (:582) and the NEXT crossing's cap edge (:592);
    // the cap edge created by the last crossing is redundant and goes straight back.
    auto pop_slot = [&](int j) -> uint32_t { return j < M.e_top ? (uint32_t)sm->estack[M.e_top - 1 - j] : (uint32_t)(M.e_hwm + (j - M.e_top)); };
    uint32_t nv = 0, ck = 0, br = 0;
    if (act) {
        nv = sm->vfree[wi];
        ck = pop_slot(2 * (int)wi);
        br = pop_slot(2 * (int)wi + 1);
    }
    uint32_t nv_pred, ck_pred;  // previous_intersection (:550, :600) and its cap edge
    if constexpr (Cfg::REG) {
        const uint32_t both = __shfl_sync(FULL, nv | (ck << 8), (int)pk);
        nv_pred = both & 0xFFu;
        ck_pred = both >> 8;
    } else {
        nv_pred = __shfl_sync(FULL, nv, (int)pk);
        ck_pred = __shfl_sync(FULL, ck, (int)pk);
    }
    const uint32_t ck0 = __shfl_sync(FULL, ck, (int)k0);
    const uint32_t br_succ = __shfl_sync(FULL, br, (int)ks);
    uint32_t nb_lo = 0, nb_hi = 0;
    if (act) {
        const Vec3 a = {sm->vx[pv], sm->vy[pv], sm->vz[pv]};
        const Vec3 b = {sm->vx[cv], sm->vy[cv], sm->vz[cv]};
        const Vec3 x = intersection(pl, a, b);  // :567-572 (a = outside end, b = inside end)
        sm->vx[nv] = x.x;
        sm->vy[nv] = x.y;
        sm->vz[nv] = x.z;
        sm->edge[br] = MeshT::pack(r, ck, nv, f);                     // bridge (:582-587)
        sm->edge[ck] = MeshT::pack(ck_pred, br, nv_pred, (uint32_t)cap_face);  // cap edge (:592-598 / D6)
        sm->edge[o] = MeshT::pack(br, fo, nv_pred, f);                // :550 + :590
        sm->fstart[f] = (Idx)o;                                       // :578-580
        if constexpr (Cfg::REG) {
            if (nv < 32u) nb_lo = 1u << nv; else nb_hi = 1u << (nv - 32u);
            // the three half-edges leaving the new vertex: the re-entering one, the cap edge, and the
            // bridge of the next crossed face
            sm->vout[3u * nv] = (Idx)r;
            sm->vout[3u * nv + 1u] = (Idx)ck;
            sm->vout[3u * nv + 2u] = (Idx)br_succ;
        }
    }
    if (lane == 0) {
        sm->fnbr[cap_face] = neighbor_id;  // Face.point_index (:479-482)
        sm->fstart[cap_face] = (Idx)ck0;
    }
    // allocator state after 2K+1 pops; the redundant cap edge is freed first (it lands where it was
    // when it came off the stack)
    {
        const uint32_t red = pop_slot(2 * (int)K);
        const int pops = 2 * (int)K + 1;
        const int from_stack = pops < M.e_top ? pops : M.e_top;
        M.e_hwm += pops - from_stack;
        M.e_top -= from_stack;
        __syncwarp();  // every lane has read its slots off the stack (pop_slot) before lane 0 puts one back (racecheck: WAR)
        if (lane == 0) {
            sm->estack[M.e_top] = (Idx)red;
            sm->edge[red] = MeshT::FREE_EDGE;
        }
        M.e_top += 1;
    }
    __syncwarp();
    // ---- retire: Outside vertices, dead half-edges (ascending slot order), faces without edges ----
    if constexpr (Cfg::REG) {
        nb_lo = __reduce_or_sync(FULL, nb_lo);
        nb_hi = __reduce_or_sync(FULL, nb_hi);
        M.vlive.set_word(0, (M.vlive.word(0) & ~M.outside.word(0)) | nb_lo);
        if (MeshT::NWV > 1) M.vlive.set_word(1, (M.vlive.word(1) & ~M.outside.word(1)) | nb_hi);
    } else {
        for (int p = lane; p < M.nwv(); p += 32) sm->m_vlive[p] &= ~sm->m_outside[p];
        __syncwarp();
        if (act) atomicOr(&sm->m_vlive[nv >> 5], 1u << (nv & 31u));
        M.note_vertex((int)__reduce_max_sync(FULL, act ? nv : 0u));
        __syncwarp();
    }
    if (adj) {
        if constexpr (Cfg::REG) {
            const uint32_t dm = __ballot_sync(FULL, my_dead != Cfg::NONE);
            if (dm) {
                // Pool::remove in ascending slot order, as the sweep would do it
                const uint32_t nd = __popc(dm);
                uint8_t* dl = reinterpret_cast<uint8_t*>(sm->dlist);
                dl[lane] = 0xFFu;  // padding: no slot is below it
                __syncwarp();
                if (my_dead != Cfg::NONE) dl[__popc(dm & lt)] = (uint8_t)my_dead;
                __syncwarp();
                // rank = how many dying slots are below mine: four byte-compares per word
                const uint32_t y4 = my_dead * 0x01010101u;
                const uint32_t nw = (nd + 3u) >> 2;
                uint32_t below = 0;
#pragma unroll
                for (uint32_t i = 0; i < 8u; ++i) {
                    if (i >= nw) break;
                    below |= (__vcmpltu4(sm->dlist[i], y4) & 0x80808080u) >> i;
                }
                const uint32_t rank = __popc(below);
                if (my_dead != Cfg::NONE) {
                    sm->estack[M.e_top + rank] = (Idx)my_dead;
                    sm->edge[my_dead] = MeshT::FREE_EDGE;
                }
                M.e_top += nd;
                // faces of dying half-edges die unless the plane crosses them (then they hold an outgoing half-edge)
                df_lo = __reduce_or_sync(FULL, df_lo);
                df_hi = __reduce_or_sync(FULL, df_hi);
                const uint32_t cr_lo = __reduce_or_sync(FULL, (act && f < 32u) ? 1u << f : 0u);
                const uint32_t cr_hi = __reduce_or_sync(FULL, (act && f >= 32u) ? 1u << (f - 32u) : 0u);
                const uint32_t d0 = df_lo & ~cr_lo & M.flive.word(0);
                M.flive.set_word(0, M.flive.word(0) & ~d0);
                M.retire_faces(0, d0);
                if (MeshT::NWF > 1) {
                    const uint32_t d1 = df_hi & ~cr_hi & M.flive.word(1);
                    M.flive.set_word(1, M.flive.word(1) & ~d1);
                    M.retire_faces(1, d1);
                }
            }
        }
    } else if (__any_sync(FULL, deadbits != 0u)) {
        const int old_top = M.e_top;
        for (int p = 0; p < n_pass; ++p) {
            const bool dead = (deadbits >> p) & 1u;
            const uint32_t dm = __ballot_sync(FULL, dead);
            if (dead) sm->estack[M.e_top + __popc(dm & lt)] = (Idx)(32 * p + lane);
            M.e_top += __popc(dm);
        }
        __syncwarp();
        for (int i = old_top + lane; i < M.e_top; i += 32) sm->edge[sm->estack[i]] = MeshT::FREE_EDGE;
        // a face can only lose all its half-edges when some half-edge died
        if constexpr (Cfg::REG) {
            keep_lo = __reduce_or_sync(FULL, keep_lo) | (cap_face < 32 ? 1u << cap_face : 0u);
            keep_hi = __reduce_or_sync(FULL, keep_hi) | (cap_face >= 32 ? 1u << (cap_face - 32) : 0u);
            const uint32_t o0 = M.flive.word(0), o1 = MeshT::NWF > 1 ? M.flive.word(1) : 0u;
            M.flive.set_word(0, o0 & keep_lo);
            M.retire_faces(0, o0 & ~keep_lo);
            if (MeshT::NWF > 1) {
                M.flive.set_word(1, o1 & keep_hi);
                M.retire_faces(1, o1 & ~keep_hi);
            }
        } else {
            // faces that still own a half-edge: one more sweep (the retired slots are marked free by now)
            for (int q = lane; q < M.nwf(); q += 32) sm->m_fkeep[q] = 0u;
            __syncwarp();
            const int np2 = M.edge_passes();
            for (int p = 0; p < np2; ++p) {
                const EW w = M.edge_of_pass(p);
                if (!MeshT::e_is_free(w)) {
                    const uint32_t face = MeshT::e_face(w);
                    atomicOr(&sm->m_fkeep[face >> 5], 1u << (face & 31u));
                }
            }
            __syncwarp();
            for (int q = 0; q < M.nwf(); ++q) {
                const uint32_t o = sm->m_flive[q], k = sm->m_fkeep[q];
                __syncwarp();
                if (lane == 0) sm->m_flive[q] = o & k;
                M.retire_faces(q, o & ~k);
            }
            __syncwarp();
        }
    }
    cnt_nv += K;
    __syncwarp();
    return 1;
}

// ---------------------------------------------------------------------------------------------
// One plane against the mesh: Polyhedron::cut_with_plane (polyhedron.rs:438-642).
// Returns 0 = no cut, 1 = cut, 2 = skipped (D17), <0 = capacity overflow / inconsistency.
// ---------------------------------------------------------------------------------------------
// SERIAL = false compiles the reference-shaped serial walk out: a plane that needs it (a vertex on the plane, a cut
// the lane-parallel form declines) returns CUT_NEEDS_SERIAL and the cell is redone by an instantiation that has it.
// Without the walk — whose loops run on values read from shared memory — and with alloc_face's uniform slot the
// compiler can prove the warp converged at every collective of the kernel and drops its divergence guards.
template <class Cfg, bool SERIAL>
__device__ int cut_with_plane(Mesh<Cfg>& M, const Plane& pl, long long neighbor_id, uint32_t& status, uint32_t& cnt_vc, uint32_t& cnt_nv, bool serial_only,
                              bool sweep_only) {
    using MeshT = Mesh<Cfg>;
    using EW = typename Cfg::EdgeWord;
    WarpSmem<Cfg>* sm = M.sm;
    const int lane = M.lane;

    // ---- classify every live vertex (find_outgoing_edge's vertex scan, polyhedron.rs:399-405,
    //      and every later vector_location call of the walk) --------------------------------------
    uint32_t any_out = 0, any_incident = 0;
    uint32_t nlive = 0, n_out = 0;
    if constexpr (Cfg::REG) __syncwarp();  // the previous plane's readers of ovl[] are done (racecheck: the full-mask votes order execution, not memory)
#pragma unroll
    for (int p = 0; p < MeshT::NWV; ++p) {
        if (p >= M.nwv()) break;
        const uint32_t lw = M.vlive.word(p);
        M.vbefore.set_word(p, lw);
        if (lw == 0u) {  // an empty word of slots (the upper half of a small cell, most of the time)
            M.inside.set_word(p, 0u);
            M.outside.set_word(p, 0u);
            continue;
        }
        const int v = 32 * p + lane;
        const bool live = (lw >> lane) & 1u;
        double sd = 0.0;
        if (live) sd = signed_distance(pl, sm->vx[v], sm->vy[v], sm->vz[v]);
        const uint32_t in = __ballot_sync(FULL, live && sd < -TESS_TOL);  // vector3.rs:173
        const uint32_t out = __ballot_sync(FULL, live && sd > TESS_TOL);  // vector3.rs:171
        M.inside.set_word(p, in);
        M.outside.set_word(p, out);
        if ((out >> lane) & 1u) {  // list of the Outside vertices (for the sweep-free cut)
            const uint32_t rk = n_out + __popc(out & ((1u << lane) - 1u));
            if (rk < 32u) sm->ovl[rk] = (typename Cfg::Idx)v;
        }
        n_out += __popc(out);
        any_out |= out;
        any_incident |= lw & ~in & ~out;
        nlive += __popc(lw);
    }
    M.n_outside = n_out;
    cnt_vc += nlive;
    if (!any_out) return 0;  // polyhedron.rs:408-410

    // ---- fast path: no vertex on the plane -> the whole cut in lane-parallel form ----------------
    if (!any_incident && !serial_only) {
        __syncwarp();
        const int rc = cut_parallel<Cfg, SERIAL>(M, pl, neighbor_id, cnt_nv, sweep_only);
        if (rc != CUT_FALLBACK) return rc;
    }
    if constexpr (!SERIAL) return CUT_NEEDS_SERIAL;

    // ---- first edge (slot order) with target Inside whose flip's target is Outside; the walk
    //      starts on the flip (polyhedron.rs:413-432) ---------------------------------------------
    int first = -1;
    const int n_pass = M.edge_passes();
    for (int p = 0; p < n_pass && first < 0; ++p) {
        const EW w = M.edge_of_pass(p);
        bool hit = false;
        uint32_t fl = 0;
        if (!MeshT::e_is_free(w) && M.inside.test(MeshT::e_tgt(w))) {
            fl = MeshT::e_flip(w);
            hit = M.outside.test(MeshT::e_tgt(sm->edge[fl]));
        }
        const uint32_t hm = __ballot_sync(FULL, hit);
        if (hm) first = (int)__shfl_sync(FULL, fl, __ffs(hm) - 1);
    }
    if (first < 0) {  // the reference silently skips this plane
        status |= ST_DEGENERATE_SKIP;
        return 2;
    }

    // ---- the walk (polyhedron.rs:475-623), warp-uniform ---------------------------------------
    M.simple = false;  // vertices created here are not entered into vout[] (and may have valence > 3)
    TESS_UNIFORM_BEGIN(sm, sizeof(*sm));
    const int cap_first = M.alloc_edge();
    const int cap_face = M.alloc_face();
    if (cap_first < 0 || cap_face < 0) {
        TESS_UNIFORM_END();
        return -1;
    }
    sm->fnbr[cap_face] = neighbor_id;  // Face.point_index (polyhedron.rs:479-482)
    sm->fstart[cap_face] = (typename Cfg::Idx)cap_first;
    sm->edge[cap_first] = MeshT::pack(Cfg::NONE, Cfg::NONE, Cfg::NONE, (uint32_t)cap_face);

    M.removed.clear();  // vertices_to_destroy (polyhedron.rs:487)
    uint32_t out_e = (uint32_t)first;
    uint32_t prev_int = Cfg::NONE;  // previous_intersection
    uint32_t cap_cur = (uint32_t)cap_first;  // outside_face_edge_index (:483)
    uint32_t nbridge = 0;
    int budget = 2 * Cfg::EMAX;  // a consistent mesh cannot need more steps

    // The walk touches topology only.  Each crossing is recorded in sm->xlist and its vertex
    // (Plane::intersection, or the copy of an Incident vertex) is computed afterwards, one lane per
    // crossing: coordinates of NEW vertices are never read while walking.
    auto emit_vertices = [&](uint32_t count) {
        TESS_UNIFORM_END();
        __syncwarp();
        if ((uint32_t)lane < count) {
            const EW rec = sm->xlist[lane];
            const uint32_t pv = MeshT::e_next(rec), cv = MeshT::e_flip(rec), nv = MeshT::e_tgt(rec);
            const Vec3 a = {sm->vx[pv], sm->vy[pv], sm->vz[pv]};
            Vec3 x = a;  // previous vertex Incident: plain copy (polyhedron.rs:555-565)
            if (MeshT::e_face(rec) == 0u) {
                const Vec3 b = {sm->vx[cv], sm->vy[cv], sm->vz[cv]};
                x = intersection(pl, a, b);  // :567-572 (a = outside end, b = inside end)
            }
            sm->vx[nv] = x.x;
            sm->vy[nv] = x.y;
            sm->vz[nv] = x.z;
        }
        __syncwarp();
        TESS_UNIFORM_BEGIN(sm, sizeof(*sm));
    };

    do {
        const EW w_out = sm->edge[out_e];
        uint32_t pv = MeshT::e_tgt(w_out);  // previous_vertex_index (:491)
        M.removed.set(pv);
        uint32_t cur_e = MeshT::e_next(w_out);  // :506
        EW w_cur = sm->edge[cur_e];
        uint32_t cv = MeshT::e_tgt(w_cur);
        bool need = M.outside.test(pv);  // :519
        while (!M.inside.test(cv)) {     // :529-544
            need = true;
            M.removed.set(cv);
            pv = cv;
            cur_e = MeshT::e_next(w_cur);
            w_cur = sm->edge[cur_e];
            cv = MeshT::e_tgt(w_cur);
            if (--budget < 0) {
                status |= ST_INCONSISTENT;
                TESS_UNIFORM_END();
                return -2;
            }
        }
        if (need) {  // :552-601
            const int nv = M.vlive.alloc(Cfg::VMAX);
            const int br = M.alloc_edge();
            if (nv < 0 || br < 0) {
                TESS_UNIFORM_END();
                return -1;
            }
            M.note_vertex(nv);
            sm->xlist[nbridge & 31u] = MeshT::pack(pv, cv, (uint32_t)nv, M.outside.test(pv) ? 0u : 1u);
            const uint32_t f = MeshT::e_face(w_out);
            sm->fstart[f] = (typename Cfg::Idx)out_e;  // :578-580
            // bridge (:582-587), paired with the current cap edge (:589); then, as the reference does,
            // the cap edge of the NEXT crossing (:592-598): target = this intersection, next = current cap edge
            sm->edge[br] = MeshT::pack(cur_e, cap_cur, (uint32_t)nv, f);
            M.set_field(cap_cur, 1, (uint32_t)br);
            sm->edge[out_e] = MeshT::pack((uint32_t)br, MeshT::e_flip(w_out), prev_int, f);  // :550 + :590
            const int nc = M.alloc_edge();
            if (nc < 0) {
                TESS_UNIFORM_END();
                return -1;
            }
            sm->edge[nc] = MeshT::pack(cap_cur, Cfg::NONE, (uint32_t)nv, (uint32_t)cap_face);
            cap_cur = (uint32_t)nc;
            prev_int = (uint32_t)nv;
            ++nbridge;
            if ((nbridge & 31u) == 0u) emit_vertices(32u);
        } else {
            M.set_field(out_e, 2, prev_int);  // :550
        }
        out_e = MeshT::e_flip(w_cur);  // :603-607
        if (--budget < 0) {
            status |= ST_INCONSISTENT;
            TESS_UNIFORM_END();
            return -2;
        }
    } while (out_e != (uint32_t)first);  // :620-622
    if (nbridge & 31u) emit_vertices(nbridge & 31u);
    cnt_nv += nbridge;

    // close the loop: first outgoing edge and first cap edge end at the last intersection;
    // the cap edge created by the last crossing is the redundant twin of cap_first and is freed first
    {
        const uint32_t redundant = cap_cur;
        const uint32_t last_paired = MeshT::e_next(sm->edge[redundant]);
        M.set_field((uint32_t)first, 2, prev_int);
        M.set_field((uint32_t)cap_first, 2, prev_int);
        M.set_field((uint32_t)cap_first, 0, last_paired);
        sm->edge[redundant] = MeshT::FREE_EDGE;
        sm->estack[M.e_top] = (typename Cfg::Idx)redundant;
        M.e_top += 1;
    }
    TESS_UNIFORM_END();
    __syncwarp();

    // ---- retire what was cut off (clean_up_vertices / clean_up_edges / mark_sweep,
    //      polyhedron.rs:645-730): vertices reached by the walk plus everything connected to them
    //      through non-Inside vertices; half-edges with both ends removed; faces left without edges.
    uint32_t missing = 0;
#pragma unroll
    for (int p = 0; p < MeshT::NWV; ++p) {
        if (p >= M.nwv()) break;
        missing |= (M.vbefore.word(p) & ~M.inside.word(p) & ~M.removed.word(p));
    }
    if (missing) {
        // interior of the cut-off region: grow `removed` along edges between non-Inside vertices
        bool changed = true;
        int it = 0;
        while (changed && it++ < Cfg::VMAX) {
            changed = false;
            const int np = M.edge_passes();
            for (int p = 0; p < np; ++p) {
                uint32_t add_v = Cfg::NONE;
                const EW w = M.edge_of_pass(p);
                if (!MeshT::e_is_free(w)) {
                    const uint32_t t = MeshT::e_tgt(w), fl = MeshT::e_flip(w);
                    if (t != Cfg::NONE && fl != Cfg::NONE) {
                        const uint32_t s = MeshT::e_tgt(sm->edge[fl]);
                        if (s != Cfg::NONE && M.removed.test(s) && !M.removed.test(t) && M.vbefore.test(t) && !M.inside.test(t)) add_v = t;
                    }
                }
                uint32_t am = __ballot_sync(FULL, add_v != Cfg::NONE);
                TESS_UNIFORM_BEGIN(sm, sizeof(*sm));
                while (am) {  // rare: serialise
                    const int l = __ffs(am) - 1;
                    am &= am - 1;
                    const uint32_t t = __shfl_sync(FULL, add_v, l);
                    if (!M.removed.test(t)) {
                        M.removed.set(t);
                        changed = true;
                    }
                }
                TESS_UNIFORM_END();
            }
        }
    }
#pragma unroll
    for (int p = 0; p < MeshT::NWV; ++p) {
        if (p >= M.nwv()) break;
        M.vlive.set_word(p, M.vlive.word(p) & ~M.removed.word(p));
    }

    M.fkeep.clear();
    if constexpr (!Cfg::REG) __syncwarp();
    bool inconsistent = false;
    const int old_top = M.e_top;
    const int n_pass2 = M.edge_passes();
    for (int p = 0; p < n_pass2; ++p) {
        bool dead = false;
        uint32_t face = 0;
        const EW w = M.edge_of_pass(p);
        const bool live = !MeshT::e_is_free(w);
        if (live) {
            const uint32_t t = MeshT::e_tgt(w);
            const uint32_t s = MeshT::e_tgt(sm->edge[MeshT::e_flip(w)]);
            const bool tr = M.removed.test(t), sr = M.removed.test(s);
            dead = tr && sr;
            inconsistent |= (tr != sr);
            face = MeshT::e_face(w);
        }
        // retire dead half-edges: mark the slot free and push it on the free stack (ascending slot order)
        const uint32_t dm = __ballot_sync(FULL, dead);
        if (dead) {
            sm->estack[M.e_top + __popc(dm & ((1u << lane) - 1u))] = (typename Cfg::Idx)(32 * p + lane);
        }
        M.e_top += __popc(dm);
        const bool keep = live && !dead;  // this half-edge keeps its face alive
        if constexpr (Cfg::REG) {
#pragma unroll
            for (int q = 0; q < MeshT::NWF; ++q) {
                const uint32_t bits = (keep && (face >> 5) == (uint32_t)q) ? (1u << (face & 31u)) : 0u;
                M.fkeep.set_word(q, M.fkeep.word(q) | __reduce_or_sync(FULL, bits));
            }
        } else {
            if (keep) atomicOr(&sm->m_fkeep[face >> 5], 1u << (face & 31u));
        }
    }
    __syncwarp();
    // only now overwrite the retired slots (their words were still needed as flips above)
    for (int i = old_top + lane; i < M.e_top; i += 32) sm->edge[sm->estack[i]] = MeshT::FREE_EDGE;
    if (__any_sync(FULL, inconsistent)) status |= ST_INCONSISTENT;
#pragma unroll
    for (int q = 0; q < MeshT::NWF; ++q) {
        if (q >= M.nwf()) break;
        const uint32_t o = M.flive.word(q), k = M.fkeep.word(q);
        if constexpr (!Cfg::REG) __syncwarp();
        M.flive.set_word(q, o & k);
        M.retire_faces(q, o & ~k);
    }
    __syncwarp();
    return 1;
}

// ---------------------------------------------------------------------------------------------
// The kernel: persistent warps pull cells from a work counter.
// ---------------------------------------------------------------------------------------------
// COUNT: also accumulate the work counters (a separate instantiation keeps them out of the timed kernel).
template <class Cfg, bool COUNT, bool SERIAL = true>
__global__ void __launch_bounds__(Cfg::WARPS * 32, Cfg::MINB) clip_kernel(const ClipParams P) {
    using MeshT = Mesh<Cfg>;
    using EW = typename Cfg::EdgeWord;
#ifdef TESS_WARP_EMU  // tests/emu: this source run lane by lane on the CPU (test infrastructure only)
    unsigned char* smem_raw = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    // one warp per CTA (small and large configurations): the tables sit at the CTA's shared-memory base, a constant —
    // no pointer to keep in registers or to re-derive from the thread index
    WarpSmem<Cfg>* sm = reinterpret_cast<WarpSmem<Cfg>*>(smem_raw) + (Cfg::WARPS == 1 ? 0u : (threadIdx.x >> 5));
    const int lane = Cfg::WARPS == 1 ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    MeshT M;
    M.bind(sm, lane);

    const GridSpec& G = P.grid;
    const int cpd = (int)G.cpd;
    const bool radius_mode = !(P.search_radius != P.search_radius);  // not NaN
    unsigned long long t_vis = 0, t_test = 0, t_vc = 0, t_cuts = 0, t_nv = 0, t_tab = 0, t_deg = 0, t_faces = 0;  // totals of finished cells

    // the next cell is claimed while the current one is built: the atomic's round trip is off the critical path
    uint32_t claimed = 0;
    if (lane == 0) claimed = atomicAdd(P.work_counter, 1u);
    for (;;) {
        const uint32_t work = __shfl_sync(FULL, claimed, 0);
        if (work >= P.n_work) break;
        if (lane == 0) claimed = atomicAdd(P.work_counter, 1u);
        uint32_t c_vis = 0, c_test = 0, c_vc = 0, c_cuts = 0, c_nv = 0, c_tab = 0, c_deg = 0;  // this cell

        // ---- the cell's particle (Diagram::get_cell_at_index, interface.rs:193-207) -----------
        uint32_t self_slot = 0xFFFFFFFFu;
        uint32_t item;  // what the work lists hold: the cell's sorted slot, or the index of its query position
        int hx, hy, hz;
        {
            double px, py, pz;
            item = P.work_slots ? P.work_slots[work] : P.slot_begin + work;
            if (P.query_xyz) {  // get_cell_at_particle (interface.rs:218-231): no self exclusion
                px = P.query_xyz[3 * (size_t)item];
                py = P.query_xyz[3 * (size_t)item + 1];
                pz = P.query_xyz[3 * (size_t)item + 2];
            } else {
                self_slot = item;
                const double2* q = reinterpret_cast<const double2*>(P.sorted + self_slot);
                const double2 a = __ldg(q), b = __ldg(q + 1);
                px = a.x; py = a.y; pz = b.x;
            }
            if (lane == 0) { sm->cpos[0] = px; sm->cpos[1] = py; sm->cpos[2] = pz; }
            M.build_cube(P.box, px, py, pz);
            // ExpandingSearch::new (celery.rs:882-902): home cell of the position
            hx = (int)axis_index(px, G.xmin, G.xmax, G.ix, G.cpd);
            hy = (int)axis_index(py, G.ymin, G.ymax, G.iy, G.cpd);
            hz = (int)axis_index(pz, G.zmin, G.zmax, G.iz, G.cpd);
        }
        uint32_t status = 0;
        const double rmax2 = M.max_radius_sq();


        // one threshold: security mode compares table keys AND |r|^2 with 4*max|v|^2; reference-radius
        // mode compares table keys with the caller's radius (celery.rs:1036) and rejects nothing
        double stop_thr = radius_mode ? P.search_radius : mul(4.0, rmax2);
#define rej_ok(r2_) (radius_mode || (r2_) < stop_thr)
        bool done = false;
        bool failed = false;
        bool stale_thr = false;  // the threshold lags behind the last cuts (non-counting instantiation only)

        for (uint32_t t0 = 0; !done; t0 += 32) {
            // ---- 32 search_order entries, one per lane (celery.rs:981-1014) -------------------
            const uint32_t ti = t0 + lane;
            const bool tvalid = ti < P.table_len;
            double key = 0.0;
            uint32_t d0 = 0, cnt = 0;
            bool marker = false;
            if (tvalid) {
                const ShellEntry e = P.table[ti];
                key = e.key;
                const int gx = hx + e.di, gy = hy + e.dj, gz = hz + e.dk;
                if (gx >= 0 && gx < cpd && gy >= 0 && gy < cpd && gz >= 0 && gz < cpd) {
                    if (gx < (int)G.local_lo || gx >= (int)G.local_hi) {
                        marker = true;  // a plane this rank does not hold
                        cnt = 1;
                    } else {
                        const uint32_t c = ((uint32_t)(gx - (int)G.local_lo) * G.cpd + (uint32_t)gy) * G.cpd + (uint32_t)gz;
                        d0 = __ldg(P.delim + c);
                        cnt = __ldg(P.delim + c + 1) - d0;
                    }
                }
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += u;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            const uint32_t excl = incl - cnt;

            for (uint32_t base = 0; base < total && !done; base += 32) {
                // ---- one lane per candidate of the flattened ranges --------------------------
                const uint32_t q = base + lane;
                const bool has = q < total;
                int L = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t v = __shfl_sync(FULL, incl, L + step - 1);
                    if (v <= q) L += step;
                }
                L = has ? L : 31;
                const uint32_t src_d0 = __shfl_sync(FULL, d0, L);
                const uint32_t src_excl = __shfl_sync(FULL, excl, L);
                const double src_key = __shfl_sync(FULL, key, L);
                const bool src_marker = __shfl_sync(FULL, (int)marker, L) != 0;
                const uint32_t slot = src_d0 + (q - src_excl);
                double rx = 0, ry = 0, rz = 0, r2 = 0;
                long long cid = 0;
                bool adm = false;
                if (has && !src_marker) {
                    const double2* cq = reinterpret_cast<const double2*>(P.sorted + slot);
                    const double2 a = __ldg(cq), b = __ldg(cq + 1);
                    // interface.rs:322-326: search point - position
                    rx = subd(a.x, sm->cpos[0]); ry = subd(a.y, sm->cpos[1]); rz = subd(b.x, sm->cpos[2]);
                    cid = __double_as_longlong(b.y);
                    r2 = dot3(rx, ry, rz, rx, ry, rz);
                    adm = slot != self_slot;  // interface.rs:283/301 (by index)
                    if (adm && P.target_group != -1)  // interface.rs:284/293 (-2: no particle carries the requested group)
                        adm = P.target_group >= 0 && P.groups_sorted[slot] == (uint64_t)P.target_group;
                }
                uint32_t uncounted = __ballot_sync(FULL, has && !src_marker);  // C_vis bookkeeping
                // bisector plane of every candidate that can still matter (Plane::halfway_from_origin_to)
                const bool cand = (has && src_marker) || (adm && rej_ok(r2));
                __syncwarp();  // the previous tile's planes have all been read
                if (cand && !src_marker) {
                    const Plane mypl = halfway_from_origin_to(Vec3{rx, ry, rz});
                    sm->cand_plane[lane] = make_double4(mypl.nx, mypl.ny, mypl.nz, mypl.off);
                    sm->cand_id[lane] = cid;
                }
                __syncwarp();
                uint32_t pending = __ballot_sync(FULL, cand && !(src_key > stop_thr));
                const uint32_t marker_lanes = __ballot_sync(FULL, src_marker);
                // Screen (not in the reference, results unchanged): with many candidates waiting, every lane
                // first tests ITS plane against all live vertices.  A plane that keeps every vertex of the
                // current polytope inside by a margin far above any rounding error (1e-9 of the cell's radius;
                // later vertices are convex combinations of these, each off by a few ulp) can never have a
                // vertex Outside: find_outgoing_edge (polyhedron.rs:399-410) would return None for it whenever
                // it is offered, so the one-plane-at-a-time classification is skipped for it (it still counts
                // as tested).  Dense tiles of clustered inputs lose most of their non-cutting planes here.
                uint32_t nocut = 0;
                auto screen = [&]() {
                    const double mar = mul(1e-9, __dsqrt_rn(stop_thr));
                    bool clear_of_all = false;
                    if (((pending >> lane) & 1u) && !src_marker) {
                        clear_of_all = true;
                        const double4 q = sm->cand_plane[lane];
                        const double thr = subd(q.w, mar);
#pragma unroll
                        for (int p = 0; p < MeshT::NWV; ++p) {
                            if (p >= M.nwv()) break;
                            for (uint32_t m = M.vlive.word(p); m; m &= m - 1u) {
                                const int v = 32 * p + __ffs(m) - 1;
                                const double d = __fma_rn(q.x, sm->vx[v], __fma_rn(q.y, sm->vy[v], mul(q.z, sm->vz[v])));
                                clear_of_all = clear_of_all && (d < thr);  // false for NaN planes too: those take the normal path
                            }
                        }
                    }
                    nocut |= __ballot_sync(FULL, clear_of_all);
                    // without work counters the screened-out planes simply leave the queue; with them each is
                    // still counted when its turn comes (tested, all vertices classified), as the oracle counts it
                    if (!COUNT) pending &= ~nocut;
                };
                // (not the first tile of a cell: it meets the whole container, which nearly every plane cuts, and a
                // screen half-way through it catches 1-2 planes for the price of one more pass over the vertices)
                if (TESS_PREFILTER_MIN > 0 && !radius_mode && (t0 | base) != 0u && __popc(pending) >= TESS_PREFILTER_MIN) screen();
                while (pending) {
                    const int l = __ffs(pending) - 1;
                    pending &= pending - 1;
                    if ((marker_lanes >> l) & 1u) {
                        if (stale_thr) {  // a plane this rank does not hold: decide with the exact threshold
                            stop_thr = mul(4.0, M.max_radius_sq());
                            stale_thr = false;
                            pending &= __ballot_sync(FULL, cand && !(src_key > stop_thr) && (src_marker || rej_ok(r2)));
                            if (__shfl_sync(FULL, (int)(src_key > stop_thr), l)) continue;
                        }
                        status |= ST_HALO_INSUFFICIENT;
                        continue;
                    }
                    if (COUNT && ((nocut >> l) & 1u)) {  // screened out above: tested, every vertex classified Inside, no cut
                        c_test += 1;
#pragma unroll
                        for (int p = 0; p < MeshT::NWV; ++p) {
                            if (p >= M.nwv()) break;
                            c_vc += __popc(M.vlive.word(p));
                        }
                        continue;
                    }
                    const double4 pq = sm->cand_plane[l];  // broadcast read
                    const Plane pl = {pq.x, pq.y, pq.z, pq.w};
                    const long long nid = sm->cand_id[l];
                    c_test += 1;
                    const int rc = cut_with_plane<Cfg, SERIAL>(M, pl, nid, status, c_vc, c_nv, (P.flags & 1u) != 0u, (P.flags & 2u) != 0u);
                    if (rc < 0) {
                        if (rc == -1) status |= ST_CAPACITY_OVERFLOW;
                        // queued like a cell that ran out of table: redone by the small configuration WITH the serial walk
                        if (rc == CUT_NEEDS_SERIAL) status |= ST_TABLE_EXHAUSTED;
                        failed = true;
                        done = true;
                        break;
                    }
                    if (rc == 2) c_deg += 1;
                    if (rc == 1) {
                        c_cuts += 1;
                        if (!radius_mode) {
                            if (!COUNT) {
                                // The threshold only ever rejects planes that cannot cut (header comment): without
                                // work counters it may lag behind.  It is brought up to date once per tile, not after
                                // every cut; the few planes it would have dropped meanwhile are classified (no vertex
                                // is Outside) or screened out.
                                stale_thr = true;
                                continue;
                            }
                            // candidates up to this lane were offered under the old threshold
                            const uint32_t upto = uncounted & ((2u << l) - 1u);
                            c_vis += __popc(upto & __ballot_sync(FULL, !(src_key > stop_thr)));
                            uncounted &= ~upto;
                            stop_thr = mul(4.0, M.max_radius_sq());
                            pending &= __ballot_sync(FULL, cand && !(src_key > stop_thr) && (src_marker || rej_ok(r2)));
                        }
                    }
                }
                if (stale_thr && !failed) {
                    stop_thr = mul(4.0, M.max_radius_sq());
                    stale_thr = false;
                }
                if (COUNT) c_vis += __popc(uncounted & __ballot_sync(FULL, !(src_key > stop_thr)));
                // the walk stops at the first entry whose key exceeds the threshold (celery.rs:1036)
                if (__ballot_sync(FULL, has && src_key > stop_thr)) done = true;
            }
            if (!done) {
                const uint32_t stopm = __ballot_sync(FULL, tvalid && key > stop_thr);
                if (stopm) {
                    done = true;
                    if (COUNT) c_tab += __ffs(stopm) - 1;
                } else if (COUNT) {
                    c_tab += __popc(__ballot_sync(FULL, tvalid));
                }
                if (!done && t0 + 32 >= P.table_len) {
                    if (!P.table_full) status |= ST_TABLE_EXHAUSTED;  // (both modes: a table that ends before a key exceeds the threshold has to be widened)
                    done = true;
                }
            } else if (COUNT && !failed) {
                c_tab += __popc(__ballot_sync(FULL, tvalid && !(key > stop_thr)));
            }
        }

        // ---- results: weighted normals, areas, volume, neighbours ------------------------------
        uint32_t nf = 0;
#pragma unroll
        for (int q = 0; q < MeshT::NWF; ++q) {
            if (q >= M.nwf()) break;
            nf += __popc(M.flive.word(q));
        }
        double vol_part = 0.0;
        uint32_t rank_base = 0;
        // output row of this cell (recomputed here rather than kept live through the cuts)
        size_t row = (size_t)(item - P.row_base);
        long long self_id = (long long)item;
        if (!P.query_xyz) {
            self_id = __double_as_longlong(__ldg(reinterpret_cast<const double*>(P.sorted + self_slot) + 3));
            row = P.row_of_slot ? P.row_of_slot[self_slot] : (size_t)(self_slot - P.row_base);
        }
        const size_t srow = P.stage_by_work ? (size_t)work : row;
        if (!failed) {
#pragma unroll
            for (int q = 0; q < MeshT::NWF; ++q) {
                if (q >= M.nwf()) break;
                const uint32_t lw = M.flive.word(q);
                if (lw == 0u) continue;
                double contrib = 0.0;
                if ((lw >> lane) & 1u) {
                    const int f = 32 * q + lane;
                    // Polyhedron::weighted_normal (polyhedron.rs:776-808)
                    const uint32_t s = sm->fstart[f];
                    EW w = sm->edge[s];
                    const uint32_t av = MeshT::e_tgt(w);
                    const Vec3 A = {sm->vx[av], sm->vy[av], sm->vz[av]};
                    uint32_t e = MeshT::e_next(w);
                    w = sm->edge[e];
                    uint32_t tv = MeshT::e_tgt(w);
                    Vec3 cur = sub(Vec3{sm->vx[tv], sm->vy[tv], sm->vz[tv]}, A);
                    e = MeshT::e_next(w);
                    Vec3 wn = {0.0, 0.0, 0.0};
                    int guard = 0;
                    while (e != s && guard++ < Cfg::EMAX) {
                        w = sm->edge[e];
                        tv = MeshT::e_tgt(w);
                        const Vec3 prev = cur;
                        cur = sub(Vec3{sm->vx[tv], sm->vy[tv], sm->vz[tv]}, A);
                        wn = add(wn, cross(prev, cur));
                        e = MeshT::e_next(w);
                    }
                    if (P.gv_xyz) sm->flen[f] = (uint16_t)(guard + 2);      // vertices of this face's loop
                    const double area = mul(0.5, __dsqrt_rn(dot(wn, wn)));  // interface.rs:408-410
                    contrib = dot(A, wn);                                    // polyhedron.rs:849
                    const uint32_t rank = rank_base + __popc(lw & ((1u << lane) - 1u));
                    if (rank < P.fstride) {
                        P.st_nbr[srow * P.fstride + rank] = sm->fnbr[f];
                        if (P.st_area) P.st_area[srow * P.fstride + rank] = area;
                    }
                }
                rank_base += __popc(lw);
                // volume = volume + dot(...) face after face in ascending slot order (polyhedron.rs:843-850):
                // the same summation order as the reference, so the sum is reproduced bit for bit
                for (uint32_t m = lw; m; m &= m - 1u) vol_part = addd(vol_part, __shfl_sync(FULL, contrib, __ffs(m) - 1));
            }
            if (nf > P.fstride) status |= ST_CAPACITY_OVERFLOW;
        }
        // volume = (sum over faces) / 6 (polyhedron.rs:854)
        // ---- geometry (TESS_OUT_VERTICES): Cell::compute_vertices (interface.rs:368-370) and
        //      VoronoiFace::compute_vertices (interface.rs:403-405 -> polyhedron.rs:897-919) ----------
        if (P.gv_xyz && !failed && (status & (ST_CAPACITY_OVERFLOW | ST_INCONSISTENT)) == 0) {
            __syncwarp();
            // rank of a vertex slot in the cell's vertex list (ascending slot order)
            uint32_t nv = 0;
            if constexpr (!Cfg::REG) {
                for (int p0 = 0; p0 < M.nwv(); p0 += 32) {  // exclusive prefix of the per-word populations
                    const int p = p0 + lane;
                    uint32_t c = p < M.nwv() ? __popc(sm->m_vlive[p]) : 0u, inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t u = __shfl_up_sync(FULL, inc, o);
                        if (lane >= o) inc += u;
                    }
                    if (p < M.nwv()) sm->m_removed[p] = nv + inc - c;
                    nv += __shfl_sync(FULL, inc, 31);
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int p = 0; p < MeshT::NWV; ++p) nv += __popc(M.vlive.word(p));
            }
            auto vrank = [&](uint32_t v) -> uint32_t {
                if constexpr (Cfg::REG) {
                    uint32_t r = __popc(M.vlive.word(0) & (v < 32u ? (1u << v) - 1u : 0xFFFFFFFFu));
                    if (MeshT::NWV > 1 && v >= 32u) r += __popc(M.vlive.word(1) & ((1u << (v - 32u)) - 1u));
                    return r;
                } else {
                    return sm->m_removed[v >> 5] + __popc(sm->m_vlive[v >> 5] & ((1u << (v & 31u)) - 1u));
                }
            };
            // loop offsets of the faces in slot order
            uint32_t nl = 0;
#pragma unroll
            for (int q = 0; q < MeshT::NWF; ++q) {
                if (q >= M.nwf()) break;
                const uint32_t lw = M.flive.word(q);
                const uint32_t c = ((lw >> lane) & 1u) ? sm->flen[32 * q + lane] : 0u;
                uint32_t inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += u;
                }
                if ((lw >> lane) & 1u) sm->floff[32 * q + lane] = (uint16_t)(nl + inc - c);
                nl += __shfl_sync(FULL, inc, 31);
            }
            unsigned long long vb = 0, lb = 0;
            if (lane == 0) {
                vb = atomicAdd(&P.g_cursor[0], (unsigned long long)nv);
                lb = atomicAdd(&P.g_cursor[1], (unsigned long long)nl);
                P.nverts[row] = nv;
                P.nloops[row] = nl;
                P.vbase[row] = vb;
                P.lbase[row] = lb;
            }
            vb = __shfl_sync(FULL, vb, 0);
            lb = __shfl_sync(FULL, lb, 0);
            if (vb + nv <= P.gv_cap && lb + nl <= P.gl_cap) {  // otherwise the host sees the cursors and reports
#pragma unroll
                for (int p = 0; p < MeshT::NWV; ++p) {
                    if (p >= M.nwv()) break;
                    const uint32_t lw = M.vlive.word(p);
                    if ((lw >> lane) & 1u) {
                        const uint32_t v = 32 * p + lane;
                        double* o = P.gv_xyz + (vb + vrank(v)) * 3;
                        o[0] = sm->vx[v]; o[1] = sm->vy[v]; o[2] = sm->vz[v];
                    }
                }
                uint32_t frank_base = 0;
#pragma unroll
                for (int q = 0; q < MeshT::NWF; ++q) {
                    if (q >= M.nwf()) break;
                    const uint32_t lw = M.flive.word(q);
                    if ((lw >> lane) & 1u) {
                        const int f = 32 * q + lane;
                        // compute_face_vertices (polyhedron.rs:897-919): targets from the starting edge on
                        uint32_t* o = P.gl_idx + lb + sm->floff[f];
                        const uint32_t s0 = sm->fstart[f];
                        uint32_t e = s0;
                        int guard = 0;
                        do {
                            const EW w = sm->edge[e];
                            *o++ = vrank(MeshT::e_tgt(w));
                            e = MeshT::e_next(w);
                        } while (e != s0 && ++guard < Cfg::EMAX);
                        const uint32_t rank = frank_base + __popc(lw & ((1u << lane) - 1u));
                        if (rank < P.fstride) P.st_flen[srow * P.fstride + rank] = sm->flen[f];
                    }
                    frank_base += __popc(lw);
                }
            }
        }
        if (lane == 0) {
            // cells this configuration cannot finish are queued for the large-cell / larger-table pass
            const bool bad = (status & (ST_CAPACITY_OVERFLOW | ST_INCONSISTENT | ST_TABLE_EXHAUSTED)) != 0;
            if (bad && P.failed_slots) {
                const uint32_t k = atomicAdd(P.n_failed, 1u);
                if (k < P.failed_cap) P.failed_slots[k] = item;
                // cells that only ran out of search table (a wider table in the same configuration will do)
                if ((status & (ST_CAPACITY_OVERFLOW | ST_INCONSISTENT)) == 0) atomicAdd(P.n_failed + 4, 1u);
            }
            const bool empty = (status & (ST_CAPACITY_OVERFLOW | ST_INCONSISTENT)) != 0;
            P.vol[row] = empty ? 0.0 : __ddiv_rn(vol_part, 6.0);
            P.nfaces[row] = empty ? 0u : nf;
            if (P.gv_xyz && empty) { P.nverts[row] = 0u; P.nloops[row] = 0u; }
            P.status[row] = status | (P.mark_large ? ST_LARGE_PATH : 0u);
            if (P.cell_id) P.cell_id[row] = self_id;
        }
        if (COUNT && (status & (ST_CAPACITY_OVERFLOW | ST_INCONSISTENT | ST_TABLE_EXHAUSTED)) == 0) {
            // only cells this pass finished are counted; the others are counted by the redo pass
            t_vis += c_vis; t_test += c_test; t_vc += c_vc; t_cuts += c_cuts; t_nv += c_nv; t_tab += c_tab; t_deg += c_deg; t_faces += nf;
        }
        __syncwarp();
    }

    if (COUNT && P.counters && lane == 0) {
        atomicAdd(&P.counters[CNT_VISITED], t_vis);
        atomicAdd(&P.counters[CNT_TESTED], t_test);
        atomicAdd(&P.counters[CNT_VC], t_vc);
        atomicAdd(&P.counters[CNT_CUTS], t_cuts);
        atomicAdd(&P.counters[CNT_NV], t_nv);
        atomicAdd(&P.counters[CNT_TABLE], t_tab);
        atomicAdd(&P.counters[CNT_DEGEN], t_deg);
        atomicAdd(&P.counters[CNT_FACES], t_faces);
    }
}

template <class Cfg, bool COUNT, bool SERIAL = true>
void launch_cfg(const ClipParams& p, cudaStream_t s) {
    if (!p.n_work) return;
    const size_t smem = sizeof(WarpSmem<Cfg>) * Cfg::WARPS;
#ifdef TESS_WARP_EMU
    *p.work_counter = 0u;
    emu_launch_kernel([](const void* a) { clip_kernel<Cfg, COUNT, SERIAL>(*static_cast<const ClipParams*>(a)); }, &p, Cfg::WARPS * 32, smem);
#else
    // Per device: the dynamic shared-memory opt-in, the SM count and the occupancy belong to the device the launch goes
    // to, and tess_compute_* may run on several host threads at once (tess.h).
    struct DevCfg {
        bool configured = false;
        int per_sm = 0, sms = 0;
    };
    static DevCfg cfgs[64];
    static std::mutex cfg_mutex;
    int dev = 0;
    TESS_CUDA_CHECK(cudaGetDevice(&dev));
    int sms_cached = 0, per_sm_cached = 0;
    {
        std::lock_guard<std::mutex> lock(cfg_mutex);
        DevCfg local;
        DevCfg& c = (dev >= 0 && dev < 64) ? cfgs[dev] : local;
        if (!c.configured) {
            TESS_CUDA_CHECK(cudaFuncSetAttribute(clip_kernel<Cfg, COUNT, SERIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            TESS_CUDA_CHECK(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
            TESS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, clip_kernel<Cfg, COUNT, SERIAL>, Cfg::WARPS * 32, smem));
            c.configured = true;
        }
        sms_cached = c.sms;
        per_sm_cached = c.per_sm;
    }
    int sms = sms_cached, per_sm = per_sm_cached;
    if (per_sm < 1) per_sm = 1;
    // persistent grid: a whole number of resident waves (148 SMs x resident CTAs)
    const unsigned int want = (unsigned int)((p.n_work + Cfg::WARPS - 1) / Cfg::WARPS);
    const unsigned int grid = std::min<unsigned int>(want, (unsigned int)(sms * per_sm));
    TESS_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s));
    clip_kernel<Cfg, COUNT, SERIAL><<<grid, Cfg::WARPS * 32, smem, s>>>(p);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
#endif
}

}  // namespace

void launch_clip(const ClipParams& p, int tier, cudaStream_t s) {
    const bool count = p.counters != nullptr;
    if (tier == CLIP_LARGE) {
        if (count) launch_cfg<LargeCfg, true>(p, s); else launch_cfg<LargeCfg, false>(p, s);
    } else if (tier == CLIP_MEDIUM) {
        if (count) launch_cfg<MediumCfg, true>(p, s); else launch_cfg<MediumCfg, false>(p, s);
    } else if (tier == CLIP_THREAD) {
        launch_clip_thread(p, s);
    } else if (tier == CLIP_SMALL_FAST) {
        if (count) launch_cfg<SmallCfg, true, false>(p, s); else launch_cfg<SmallCfg, false, false>(p, s);
    } else {
        if (count) launch_cfg<SmallCfg, true>(p, s); else launch_cfg<SmallCfg, false>(p, s);
    }
}
uint32_t clip_medium_fmax() { return MediumCfg::FMAX; }
uint32_t clip_medium_vmax() { return MediumCfg::VMAX; }
uint32_t clip_small_fmax() { return SmallCfg::FMAX; }
uint32_t clip_small_vmax() { return SmallCfg::VMAX; }
uint32_t clip_large_fmax() { return LargeCfg::FMAX; }
uint32_t clip_large_vmax() { return LargeCfg::VMAX; }

}  // namespace tess
