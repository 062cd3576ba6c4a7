// warp_emu.hpp — TEST INFRASTRUCTURE, not part of the product.
//
// A lane-by-lane CPU interpreter of the CUDA warp execution model, just large enough to run
// the-tessellator_b200/csrc/clip.cu UNCHANGED on a machine without a GPU: every lane of a warp is a
// fiber (own stack, hand-written x86-64 context switch), and every warp-level primitive
// (__shfl_sync, __ballot_sync, __reduce_*_sync, __all/__any_sync, __syncwarp) is a rendez-vous of the
// 32 fibers.  The `-m "not gpu"` suite uses it to check the kernel source itself — not a restatement —
// against the CPU oracle, and to check properties the GPU cannot show:
//   * every collective is reached by all 32 lanes from the SAME source line (convergence);
//   * results do not depend on the order in which lanes run between two collectives (the lanes are
//     run 0..31 and 31..0: a shared-memory exchange that lacks its __syncwarp shows up as a mismatch).
// Nothing under the-tessellator_b200/ links or loads this; the product has no CPU path.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <cmath>
#include <thread>
#include <vector>

namespace emu {

struct Warp;
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    unsigned tid = 0, bid = 0;
    unsigned lane = 0;
    uint32_t ncoll = 0;
    Warp* warp = nullptr;
    // warp-uniform region (code every lane runs redundantly on the same shared state, in lock step on the GPU)
    unsigned char* uni_ptr = nullptr;
    size_t uni_len = 0;
    uint32_t uni_span = 0;
};
struct Warp {
    uint64_t buf[2][32];
    uint32_t tag[2][32];
    uint32_t seq[2][32];
    Fiber f[32];
    unsigned long long deposits = 0;  // lanes that have entered a collective so far (32 per collective)
    std::vector<unsigned char> uni_pre, uni_post;
    uint32_t pre_span = 0, post_span = 0;
};
struct Block {
    std::vector<Warp*> warps;
    unsigned char* smem = nullptr;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    void (*entry)(const void*) = nullptr;
    const void* arg = nullptr;
    unsigned bid = 0, nblocks = 0, nthreads = 0;
    unsigned long long collectives = 0;
    unsigned long long* line_hist = nullptr;  // optional: executions of each collective by source line (65536 entries)
    unsigned bar_count = 0, bar_gen = 0;      // __syncthreads
};
extern thread_local Block* tl_block;

extern "C" void emu_switch(void** save_sp, void* load_sp);

static constexpr size_t kStack = 256 * 1024;

[[noreturn]] inline void die(const char* what, int line_a, int line_b) {
    fprintf(stderr, "[warp_emu] %s (source lines %d / %d)\n", what, line_a, line_b);
    abort();
}

inline void yield_to_scheduler() {
    Block* b = tl_block;
    emu_switch(&b->cur->sp, b->sched_sp);
}

// Warp-uniform regions.  On the GPU the 32 lanes of a converged warp execute such a region in lock step:
// all read the same shared-memory words, compute the same values, write the same words.  Here lanes run one
// after the other between two collectives, so each lane must start the span from the state the FIRST lane
// found, and all lanes must leave it with the same state (checked).
inline void uniform_span_start() {
    Fiber* f = tl_block->cur;
    if (!f->uni_ptr) return;
    Warp* w = f->warp;
    const uint32_t id = ++f->uni_span;
    if (w->pre_span != id) {
        w->uni_pre.assign(f->uni_ptr, f->uni_ptr + f->uni_len);
        w->pre_span = id;
    } else {
        memcpy(f->uni_ptr, w->uni_pre.data(), f->uni_len);
    }
}
inline void uniform_span_end(int line) {
    Fiber* f = tl_block->cur;
    if (!f->uni_ptr) return;
    Warp* w = f->warp;
    const uint32_t id = f->uni_span;
    if (w->post_span != id) {
        w->uni_post.assign(f->uni_ptr, f->uni_ptr + f->uni_len);
        w->post_span = id;
    } else if (memcmp(f->uni_ptr, w->uni_post.data(), f->uni_len) != 0) {
        die("lanes leave a warp-uniform region with different shared-memory contents", line, -1);
    }
}
inline void uniform_begin(void* p, size_t n, int line) {
    Fiber* f = tl_block->cur;
    if (f->uni_ptr) die("nested warp-uniform region", line, -1);
    f->uni_ptr = static_cast<unsigned char*>(p);
    f->uni_len = n;
    uniform_span_start();
}
inline void uniform_end(int line) {
    Fiber* f = tl_block->cur;
    if (!f->uni_ptr) die("warp-uniform region closed twice", line, -1);
    uniform_span_end(line);
    f->uni_ptr = nullptr;
}

// deposit, wait for the other 31 lanes, return the buffer they deposited into
inline const uint64_t* rendezvous(uint64_t v, uint32_t tag) {
    Block* b = tl_block;
    Fiber* f = b->cur;
    Warp* w = f->warp;
    const int k = f->ncoll & 1u;
    w->buf[k][f->lane] = v;
    w->tag[k][f->lane] = tag;
    w->seq[k][f->lane] = f->ncoll;
    const uint32_t my_seq = f->ncoll++;
    uniform_span_end((int)(tag & 0xFFFFFu));
    // wait for all 32 lanes: lanes of a warp may be one scheduling round apart after a block barrier
    w->deposits++;
    unsigned long long spins = 0;
    do {
        yield_to_scheduler();
        if (++spins > (1ull << 22)) die("a warp collective never completed (some lane does not reach it)", (int)(tag & 0xFFFFFu), -1);
    } while (w->deposits < 32ull * ((unsigned long long)my_seq + 1ull));
    uniform_span_start();
    // every lane checks its neighbour: a chain of equalities makes all 32 equal
    const unsigned nb = (f->lane + 1u) & 31u;
    if (w->seq[k][nb] != my_seq) die("lanes disagree on the number of collectives executed", (int)(tag & 0xFFFFFu), (int)(w->tag[k][nb] & 0xFFFFFu));
    if (w->tag[k][nb] != tag) die("lanes meet in different collectives (divergent call sites)", (int)(tag & 0xFFFFFu), (int)(w->tag[k][nb] & 0xFFFFFu));
    if (f->lane == 0) {
        b->collectives++;
        if (b->line_hist) b->line_hist[tag & 0xFFFFu]++;
    }
    return w->buf[k];
}

inline unsigned lane_id() { return tl_block->cur->lane; }

template <class T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

enum : uint32_t { OP_SHFL = 1u << 20, OP_UP = 2u << 20, OP_BALLOT = 3u << 20, OP_ALL = 4u << 20, OP_ANY = 5u << 20, OP_RMAX = 6u << 20, OP_RMIN = 7u << 20, OP_ROR = 8u << 20, OP_SYNC = 9u << 20, OP_DOWN = 10u << 20, OP_XOR = 11u << 20, OP_MATCH = 12u << 20, OP_RADD = 13u << 20 };

inline void need_full(unsigned mask, int line) {
    if (mask != 0xffffffffu) die("only full-warp masks are modelled", line, -1);
}

template <class T>
inline T shfl(unsigned mask, T v, int src, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(to_bits(v), OP_SHFL | (uint32_t)line);
    return from_bits<T>(b[(unsigned)src & 31u]);
}
template <class T>
inline T shfl_up(unsigned mask, T v, unsigned delta, int line) {
    need_full(mask, line);
    const unsigned l = lane_id();
    const uint64_t* b = rendezvous(to_bits(v), OP_UP | (uint32_t)line);
    return l >= delta ? from_bits<T>(b[l - delta]) : v;
}
template <class T>
inline T shfl_down(unsigned mask, T v, unsigned delta, int line) {
    need_full(mask, line);
    const unsigned l = lane_id();
    const uint64_t* b = rendezvous(to_bits(v), OP_DOWN | (uint32_t)line);
    return l + delta < 32u ? from_bits<T>(b[l + delta]) : v;
}
template <class T>
inline T shfl_xor(unsigned mask, T v, unsigned x, int line) {
    need_full(mask, line);
    const unsigned l = lane_id();
    const uint64_t* b = rendezvous(to_bits(v), OP_XOR | (uint32_t)line);
    return from_bits<T>(b[(l ^ x) & 31u]);
}
inline unsigned ballot(unsigned mask, bool p, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(p ? 1u : 0u, OP_BALLOT | (uint32_t)line);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(b[l] & 1u) << l;
    return r;
}
inline bool all_(unsigned mask, bool p, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(p ? 1u : 0u, OP_ALL | (uint32_t)line);
    for (int l = 0; l < 32; ++l)
        if (!b[l]) return false;
    return true;
}
inline bool any_(unsigned mask, bool p, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(p ? 1u : 0u, OP_ANY | (uint32_t)line);
    for (int l = 0; l < 32; ++l)
        if (b[l]) return true;
    return false;
}
inline unsigned reduce_max(unsigned mask, unsigned v, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(v, OP_RMAX | (uint32_t)line);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r = (unsigned)b[l] > r ? (unsigned)b[l] : r;
    return r;
}
inline unsigned reduce_min(unsigned mask, unsigned v, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(v, OP_RMIN | (uint32_t)line);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l) r = (unsigned)b[l] < r ? (unsigned)b[l] : r;
    return r;
}
inline unsigned reduce_or(unsigned mask, unsigned v, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(v, OP_ROR | (uint32_t)line);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)b[l];
    return r;
}
inline unsigned reduce_add(unsigned mask, unsigned v, int line) {
    need_full(mask, line);
    const uint64_t* b = rendezvous(v, OP_RADD | (uint32_t)line);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r += (unsigned)b[l];
    return r;
}
inline unsigned match_any(unsigned mask, uint64_t v, int line) {
    need_full(mask, line);
    const unsigned me = lane_id();
    const uint64_t* b = rendezvous(v, OP_MATCH | (uint32_t)line);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (b[l] == b[me] ? 1u : 0u) << l;
    return r;
}
inline void syncwarp(int line) { rendezvous(0, OP_SYNC | (uint32_t)line); }

// __syncthreads(): all threads of the block
inline void syncthreads(int line) {
    Block* b = tl_block;
    const unsigned gen = b->bar_gen;
    if (++b->bar_count == b->nthreads) {
        b->bar_count = 0;
        b->bar_gen++;
        return;
    }
    unsigned long long spins = 0;
    while (b->bar_gen == gen) {
        yield_to_scheduler();
        if (++spins > (1ull << 22)) die("__syncthreads never completed (a thread left the kernel or skipped the barrier)", line, -1);
    }
}

// ---- running a grid -------------------------------------------------------------------------
void fiber_trampoline();

struct LaunchStats {
    unsigned long long collectives = 0;
};
// when set (single host thread only), every launch adds its per-source-line collective counts here
extern unsigned long long* g_line_hist;
// how TESS_LAUNCH runs a grid: host threads, lane order
extern unsigned g_os_threads;
extern bool g_reverse;

// Runs `entry(arg)` as a grid of `nblocks` CTAs of `nthreads` threads; CTAs are spread over
// `os_threads` host threads.  `reverse`: lanes of a warp are resumed 31..0 instead of 0..31.
LaunchStats launch(void (*entry)(const void*), const void* arg, unsigned nblocks, unsigned nthreads, size_t smem_bytes, unsigned os_threads, bool reverse);

inline unsigned char* dynamic_smem() { return tl_block->smem; }

struct Dim3 {
    unsigned x, y, z;
};
inline Dim3 thread_idx() { return Dim3{tl_block->cur->tid, 0, 0}; }
inline Dim3 block_idx() { return Dim3{tl_block->cur->bid, 0, 0}; }
inline Dim3 block_dim() { return Dim3{tl_block->nthreads, 1, 1}; }
inline Dim3 grid_dim() { return Dim3{tl_block->nblocks, 1, 1}; }

}  // namespace emu
