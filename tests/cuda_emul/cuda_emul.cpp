// cuda_emul.cpp -- TEST INFRASTRUCTURE: the fiber runtime behind cuda_emul.h (see there).
#include "cuda_emul.h"

#include <stdlib.h>

// ---- fibers: one per CUDA thread of the CTA, switched cooperatively inside this OS thread -------------------------
#if defined(__x86_64__)
// emul_switch(&save_sp, load_sp): push the callee-saved registers, park the stack pointer, adopt the other one
extern "C" void emul_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl emul_switch\n"
    ".type emul_switch,@function\n"
    "emul_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size emul_switch,.-emul_switch\n");
#else
#include <ucontext.h>
#endif

emul::Cta emul::g_cta;
emul::Fiber *emul::g_cur = nullptr;
emul_uint3 emul::g_block_dim, emul::g_grid_dim;

namespace {
constexpr size_t kStackBytes = 128 * 1024;
struct Job {
    dim3 grid;
    void (*fn)(const void *);
    const void *args;
};
Job g_job;
emul::Fiber g_fibers[emul::kMaxThreads];
unsigned g_nfibers = 0;
emul::Barrier g_cta_end;
void *g_main_sp = nullptr;
#if !defined(__x86_64__)
ucontext_t g_main_ctx, g_ctx[emul::kMaxThreads];
#endif

void switch_to(emul::Fiber *next)
{
    emul::Fiber *me = emul::g_cur;
    emul::g_cur = next;
#if defined(__x86_64__)
    emul_switch(me ? &me->sp : &g_main_sp, next ? next->sp : g_main_sp);
#else
    swapcontext(me ? &g_ctx[me - g_fibers] : &g_main_ctx, next ? &g_ctx[next - g_fibers] : &g_main_ctx);
#endif
}

// Scheduling order.  0: round robin (deterministic).  Any other seed: every switch goes to a pseudo-randomly chosen
// live fiber -- warps then overtake each other in ever different orders, so shared-memory traffic that lacks its
// barrier shows up as a wrong result under some seed instead of passing by the luck of a fixed order.
uint64_t g_sched_state = 0;

emul::Fiber *next_live(emul::Fiber *from)
{
    unsigned start = (unsigned)(from - g_fibers);
    if (g_sched_state) {
        g_sched_state ^= g_sched_state << 13;                // xorshift64
        g_sched_state ^= g_sched_state >> 7;
        g_sched_state ^= g_sched_state << 17;
        start = (unsigned)(g_sched_state % g_nfibers);
    }
    for (unsigned k = 1; k <= g_nfibers; ++k) {
        emul::Fiber *f = &g_fibers[(start + k) % g_nfibers];
        if (!f->done && f != from) return f;
    }
    return nullptr;
}

// body of every fiber: all CTAs of the grid, one after the other, in dispatch order (x fastest)
void fiber_main()
{
    emul::Fiber *me = emul::g_cur;
    for (unsigned bz = 0; bz < g_job.grid.z; ++bz)
        for (unsigned by = 0; by < g_job.grid.y; ++by)
            for (unsigned bx = 0; bx < g_job.grid.x; ++bx) {
                me->bid = emul_uint3{bx, by, bz};
                g_job.fn(g_job.args);
                emul::wait(g_cta_end);                       // the next CTA reuses the shared memory
            }
    me->done = true;
    emul::Fiber *n = next_live(me);
    switch_to(n);                                            // nullptr: everybody is done, back to launch()
    __builtin_unreachable();
}
#if defined(__x86_64__)
extern "C" void emul_fiber_entry() { fiber_main(); }
#else
void fiber_entry_uc() { fiber_main(); }
#endif
}  // namespace

extern "C" void emul_set_schedule(uint64_t seed) { g_sched_state = seed; }

void emul::yield()
{
    emul::Fiber *n = next_live(g_cur);
    if (n) switch_to(n);
}

void emul::launch(dim3 grid, unsigned threads, void (*fn)(const void *), const void *args)
{
    if (threads == 0 || threads > (unsigned)kMaxThreads || threads % 32u != 0) abort();   // whole warps only, as the kernels here launch
    g_block_dim = emul_uint3{threads, 1, 1};
    g_grid_dim = emul_uint3{grid.x, grid.y, grid.z};
    g_job = Job{grid, fn, args};
    g_nfibers = threads;
    g_cta.bar = Barrier{threads, 0, 0};
    g_cta_end = Barrier{threads, 0, 0};
    g_cta.or_acc = 0;
    for (unsigned w = 0; w < threads / 32; ++w) g_cta.warps[w].bar = Barrier{32, 0, 0};
    static char *stacks = static_cast<char *>(malloc((size_t)kMaxThreads * kStackBytes + 64));   // kept for the life of the process
    char *base = stacks + (64 - (reinterpret_cast<uintptr_t>(stacks) & 63));
    for (unsigned t = 0; t < threads; ++t) {
        Fiber &f = g_fibers[t];
        f.tid = emul_uint3{t, 0, 0};
        f.bid = emul_uint3{0, 0, 0};
        f.done = false;
        char *top = base + (size_t)(t + 1) * kStackBytes;    // 64-byte aligned
#if defined(__x86_64__)
        // what emul_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into the entry with rsp = 8 mod 16 as after a call
        void **sp = reinterpret_cast<void **>(top) - 2;      // the slot of the return address is 16-byte aligned
        sp[0] = reinterpret_cast<void *>(&emul_fiber_entry);
        sp -= 6;
        for (int i = 0; i < 6; ++i) sp[i] = nullptr;
        f.sp = sp;
#else
        getcontext(&g_ctx[t]);
        g_ctx[t].uc_stack.ss_sp = top - kStackBytes;
        g_ctx[t].uc_stack.ss_size = kStackBytes;
        g_ctx[t].uc_link = nullptr;
        makecontext(&g_ctx[t], fiber_entry_uc, 0);
#endif
    }
    g_cur = nullptr;
    switch_to(&g_fibers[0]);                                 // returns when the last fiber has finished
}

