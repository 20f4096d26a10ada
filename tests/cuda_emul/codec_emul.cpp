// codec_emul.cpp -- TEST INFRASTRUCTURE: the device serializer's kernels (chunk_codec.cuh, unmodified text) executed
// on the CPU through cuda_emul.h, behind a C entry point with the shape of mb_serialize_chunks.
#define MBC_EMULATE 1
#include "cuda_emul.h"

#include <stdlib.h>

#include "chunk_codec.cuh"

thread_local emul_uint3 threadIdx, blockIdx;
emul_uint3 blockDim, gridDim;
emul::Cta emul::g_cta;

static int g_keep_mode = 0;
bool mbc_emul_keep_aggregate(uint32_t tile)
{
    if (g_keep_mode == 0) return false;
    if (g_keep_mode == 2) return true;                       // nobody ever publishes an inclusive prefix except tile 0
    return ((tile * 2654435761u) >> 28) < 11u;               // ~70 % of the tiles
}

namespace {
struct Job {
    dim3 grid;
    unsigned t;
    void (*fn)(const void *);
    const void *args;
};
void *worker(void *arg)
{
    const Job *j = static_cast<const Job *>(arg);
    // the thread id was stored behind the job by launch()
    const unsigned t = *reinterpret_cast<const unsigned *>(reinterpret_cast<const char *>(arg) + sizeof(Job));
    threadIdx = emul_uint3{t, 0, 0};
    for (unsigned by = 0; by < j->grid.y; ++by)
        for (unsigned bx = 0; bx < j->grid.x; ++bx) {
            blockIdx = emul_uint3{bx, by, 0};
            j->fn(j->args);
            pthread_barrier_wait(&emul::g_cta.bar);          // the next CTA reuses the shared memory
        }
    return nullptr;
}
}  // namespace

void emul::launch(dim3 grid, unsigned threads, void (*fn)(const void *), const void *args)
{
    blockDim = emul_uint3{threads, 1, 1};
    gridDim = emul_uint3{grid.x, grid.y, grid.z};
    pthread_barrier_init(&g_cta.bar, nullptr, threads);
    g_cta.or_acc = 0;
    for (unsigned w = 0; w < threads / 32; ++w) pthread_barrier_init(&g_cta.warps[w].bar, nullptr, 32);
    struct Slot { Job j; unsigned t; };
    std::vector<Slot> slots(threads);
    std::vector<pthread_t> th(threads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (unsigned t = 0; t < threads; ++t) {
        slots[t].j = Job{grid, t, fn, args};
        slots[t].t = t;
        pthread_create(&th[t], &attr, worker, &slots[t]);
    }
    for (unsigned t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&attr);
    pthread_barrier_destroy(&g_cta.bar);
    for (unsigned w = 0; w < threads / 32; ++w) pthread_barrier_destroy(&g_cta.warps[w].bar);
}

static void run_rle(const void *a) { mbc::rle_tiles(*static_cast<const mbc::Batch *>(a)); }
static void run_raw(const void *a) { mbc::raw_fixup(*static_cast<const mbc::Batch *>(a)); }

extern "C" int emul_serialize_chunks(const uint8_t *payloads, size_t n, uint32_t n_chunks, size_t payload_stride, uint8_t *out,
                                     size_t out_stride, void *info, int keep_mode)
{
    if (!payloads || !out || !info || n == 0 || n >= 0xffffffffull || payload_stride < n || out_stride < n + 1) return 1;
    g_keep_mode = keep_mode;
    mbc::Batch b;
    b.payloads = payloads;
    b.payload_stride = payload_stride;
    b.out = out;
    b.out_stride = out_stride;
    b.meta = static_cast<mbc::Meta *>(info);
    b.n = n;
    b.ntiles = (uint32_t)((n + mbc::kTile - 1) / mbc::kTile);
    b.n_chunks = n_chunks;
    b.status = static_cast<mbc::TileStatus *>(calloc((size_t)n_chunks * b.ntiles + n_chunks, sizeof(mbc::TileStatus)));
    if (!b.status) return 2;
    b.dead = reinterpret_cast<uint32_t *>(b.status + (size_t)n_chunks * b.ntiles);
    static_assert(mbc::kSmemBytes <= (int)sizeof(emul::g_cta.dyn_smem), "dynamic shared memory of the emulator");
    emul::launch(dim3(b.ntiles * n_chunks, 1, 1), mbc::kThreads, run_rle, &b);
    emul::launch(dim3(b.ntiles, n_chunks, 1), mbc::kThreads, run_raw, &b);
    free(b.status);
    return 0;
}
