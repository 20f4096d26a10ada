// codec_emul.cpp -- TEST INFRASTRUCTURE: the device serializer's kernels (chunk_codec.cuh, unmodified text) executed
// on the CPU through cuda_emul.h, behind a C entry point with the shape of mb_serialize_chunks.
#define MB_EMULATE 1
#include "cuda_emul.h"

#include <stdlib.h>

#include "chunk_codec.cuh"

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
