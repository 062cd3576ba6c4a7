// warp_emu.cpp — TEST INFRASTRUCTURE (see warp_emu.hpp): fibers, scheduler, grid launch.
#include "warp_emu.hpp"

#include <atomic>
#include <functional>

#if !defined(__x86_64__)
#error "the warp emulator's context switch is written for x86-64"
#endif

// void emu_switch(void** save_sp, void* load_sp): save the callee-saved registers and the stack
// pointer of the running context, continue the other one.
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

thread_local Block* tl_block = nullptr;
unsigned long long* g_line_hist = nullptr;

void fiber_trampoline() {
    Block* b = tl_block;
    Fiber* f = b->cur;
    b->entry(b->arg);
    f->done = true;
    emu_switch(&f->sp, b->sched_sp);
    abort();  // a finished fiber is never resumed
}

static void prepare(Fiber& f) {
    if (!f.stack) {
        void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) {
            perror("[warp_emu] mmap");
            abort();
        }
        f.stack = static_cast<char*>(p);
    }
    uintptr_t top = reinterpret_cast<uintptr_t>(f.stack + kStack) & ~uintptr_t(15);
    void** s = reinterpret_cast<void**>(top);
    *--s = nullptr;                                        // return address slot of the trampoline (never used)
    *--s = reinterpret_cast<void*>(&fiber_trampoline);     // popped by emu_switch's ret
    for (int i = 0; i < 6; ++i) *--s = nullptr;            // rbp rbx r12 r13 r14 r15
    f.sp = s;
    f.done = false;
    f.ncoll = 0;
    f.uni_ptr = nullptr;
    f.uni_span = 0;
}

static unsigned long long run_block(Block& b, bool reverse) {
    tl_block = &b;
    const unsigned nw = (b.nthreads + 31u) / 32u;
    for (unsigned w = 0; w < nw; ++w) {
        Warp* W = b.warps[w];
        for (unsigned l = 0; l < 32; ++l) {
            Fiber& f = W->f[l];
            f.lane = l;
            f.tid = 32u * w + l;
            f.bid = b.bid;
            f.warp = W;
            prepare(f);
        }
        W->pre_span = W->post_span = 0;
        W->deposits = 0;
    }
    b.bar_count = 0;
    b.bar_gen = 0;
    b.collectives = 0;
    bool alive = true;
    while (alive) {
        alive = false;
        for (unsigned w = 0; w < nw; ++w) {
            Warp* W = b.warps[w];
            for (unsigned i = 0; i < 32; ++i) {
                Fiber& f = W->f[reverse ? 31u - i : i];
                if (f.done) continue;
                b.cur = &f;
                emu_switch(&b.sched_sp, f.sp);
                if (!f.done) alive = true;
            }
            // (lanes of a warp may leave one round apart after a block barrier; a lane that leaves while the others
            // wait in a collective is caught by the collective's own watchdog)
        }
    }
    tl_block = nullptr;
    return b.collectives;
}

LaunchStats launch(void (*entry)(const void*), const void* arg, unsigned nblocks, unsigned nthreads, size_t smem_bytes, unsigned os_threads, bool reverse) {
    if (nthreads % 32u) die("block size must be a multiple of 32", -1, -1);
    if (os_threads < 1) os_threads = 1;
    if (os_threads > nblocks) os_threads = nblocks;
    std::atomic<unsigned> next{0};
    std::atomic<unsigned long long> coll{0};
    auto worker = [&]() {
        Block b;
        const unsigned nw = nthreads / 32u;
        for (unsigned w = 0; w < nw; ++w) b.warps.push_back(new Warp());
        std::vector<unsigned char> smem(smem_bytes + 64);
        b.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 63) & ~uintptr_t(63));
        b.entry = entry;
        b.arg = arg;
        b.nblocks = nblocks;
        b.nthreads = nthreads;
        b.line_hist = os_threads == 1 ? g_line_hist : nullptr;
        for (;;) {
            const unsigned id = next.fetch_add(1);
            if (id >= nblocks) break;
            b.bid = id;
            memset(b.smem, 0xCD, smem_bytes);  // shared memory starts undefined
            coll += run_block(b, reverse);
        }
        for (Warp* W : b.warps) {
            for (Fiber& f : W->f)
                if (f.stack) munmap(f.stack, kStack);
            delete W;
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < os_threads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    LaunchStats st;
    st.collectives = coll.load();
    return st;
}

unsigned g_os_threads = 1;
bool g_reverse = false;

}  // namespace emu

// TESS_LAUNCH of the emulated build (common.cuh)
void emu_launch_generic(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    if (!grid) return;
    emu::launch([](const void* a) { (*static_cast<const std::function<void()>*>(a))(); }, &body, grid, block, smem, emu::g_os_threads, emu::g_reverse);
}
