// codec_emul.cpp -- TEST INFRASTRUCTURE: the device serializer's kernels (chunk_codec.cuh, unmodified text) executed
// on the CPU through cuda_emul.h, behind a C entry point with the shape of mb_serialize_chunks.
#define MBC_EMULATE 1
#include "cuda_emul.h"

#include <stdlib.h>

#include "chunk_codec.cuh"

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

emul::Fiber *next_live(emul::Fiber *from)
{
    const unsigned start = (unsigned)(from - g_fibers);
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

void emul::yield()
{
    emul::Fiber *n = next_live(g_cur);
    if (n) switch_to(n);
}

void emul::launch(dim3 grid, unsigned threads, void (*fn)(const void *), const void *args)
{
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

static int g_keep_mode = 0;
bool mbc_emul_keep_aggregate(uint32_t tile)
{
    if (g_keep_mode == 0) return false;
    if (g_keep_mode == 2) return true;                       // nobody ever publishes an inclusive prefix except tile 0
    return ((tile * 2654435761u) >> 28) < 11u;               // ~70 % of the tiles
}

static void run_rle1(const void *a) { mbc::rle_tiles<1>(*static_cast<const mbc::Batch *>(a)); }
static void run_rle2(const void *a) { mbc::rle_tiles<2>(*static_cast<const mbc::Batch *>(a)); }
static void run_raw(const void *a) { mbc::raw_fixup(*static_cast<const mbc::Batch *>(a)); }

extern "C" int emul_serialize_chunks(const uint8_t *payloads, size_t n, uint32_t n_chunks, size_t payload_stride, uint8_t *out,
                                     size_t out_stride, void *info, int keep_mode, int spans, unsigned grid_y)
{
    if (!payloads || !out || !info || n == 0 || n >= 0xffffffffull || payload_stride < n || out_stride < n + 1) return 1;
    if (spans < 1 || spans > mbc::kMaxSpans || grid_y == 0) return 1;
    g_keep_mode = keep_mode;
    mbc::Batch b;
    b.payloads = payloads;
    b.payload_stride = payload_stride;
    b.out = out;
    b.out_stride = out_stride;
    b.meta = static_cast<mbc::Meta *>(info);
    b.n = n;
    const size_t tile = (size_t)mbc::tile_bytes(spans);
    b.ntiles = (uint32_t)((n + tile - 1) / tile);
    b.n_chunks = n_chunks;
    b.status = static_cast<mbc::TileStatus *>(calloc((size_t)n_chunks * b.ntiles + n_chunks, sizeof(mbc::TileStatus)));
    if (!b.status) return 2;
    b.dead = reinterpret_cast<uint32_t *>(b.status + (size_t)n_chunks * b.ntiles);
    static_assert(mbc::smem_bytes(mbc::kMaxSpans) <= (int)sizeof(emul::g_cta.dyn_smem), "dynamic shared memory of the emulator");
    // grid_y < tiles exercises the three-dimensional grid (and its padding blocks) of very large payloads
    const unsigned gy = b.ntiles < grid_y ? b.ntiles : grid_y, gz = (b.ntiles + gy - 1) / gy;
    emul::launch(dim3(n_chunks, gy, gz), mbc::kThreads, spans == 1 ? run_rle1 : run_rle2, &b);
    b.ntiles = (uint32_t)((n + mbc::kTile - 1) / mbc::kTile);
    emul::launch(dim3((b.ntiles + mbc::kRawPieces - 1) / mbc::kRawPieces, n_chunks, 1), mbc::kThreads, run_raw, &b);
    free(b.status);
    return 0;
}
