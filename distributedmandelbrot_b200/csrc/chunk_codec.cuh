// chunk_codec.cuh -- DataChunk serialisation on the device, HBM-bound byte work.
//
// Replaces DataChunk.Serialize (DistributedMandelbrot/DataChunk.cs:173-206), which serialises the
// 16 MiB payload TWICE on one CPU thread (RawSerializer, DataChunkSerializer.cs:29-49, and
// RLESerializer, :51-100, each into a SizeCountStream) to pick the shorter and then a third time
// to write it.  Output bytes are identical to the reference's:
//     byte 0 = codec: 0x00 raw -> the n payload bytes ; 0x01 RLE -> (uint32 LE runLength, uint8
//     value) records, one per maximal run.  Raw wins ties (`<` with Raw first, DataChunk.cs:188).
//
// ONE pass over HBM for every chunk that ends up RLE (rle_tiles), one CTA per 16 KiB tile (two 8 KiB pieces):
//   1. the tile is loaded once (4 x 128-bit per thread), kept in shared memory, and its run starts are found
//      with SWAR byte compares (one 32-bit mask per 32-byte span);
//   2. a decoupled look-back over the chunk's tiles carries (number of run starts, position of the last
//      run start): a tile publishes its aggregate as soon as it has counted and sums its predecessors'
//      aggregates back to the nearest published inclusive prefix, 128 predecessors per trip to L2; tiles that
//      have nothing to emit only relay;
//   3. the in-tile offsets of the run starts are appended to a list in shared memory, two per 32-bit store;
//   4. records are emitted by the tile in which a run ENDS -- there both its length (this start minus the
//      previous start, which the scan carries across tiles however long the run was) and its value (the
//      byte before the new start) are known, so no second pass and no look-ahead are needed; the chunk's
//      last run is closed by the last tile.  The output is assembled 16 bytes at a time: a thread reads the
//      five list entries behind its four records, builds the 20-byte record stream in registers with byte
//      permutes and stores the aligned 16 bytes it owns straight to global memory (no staging, no byte
//      stores except at the two ragged ends of the tile's output range).
// A tile stops emitting once the runs before it already make RLE impossible (5 * runs >= n), and later tiles of
// that chunk are told so.  Only a chunk that ends up Raw takes a second kernel (raw_fixup), which copies the
// payload behind the header byte.  Traffic: n read + serialised bytes written for RLE chunks (the algorithmic
// minimum; measured 1.07 GB read / 0.95 GB written for 64 run-dense chunks of 16 MiB).
// What bounds it (ncu, 64 run-dense chunks): instruction issue (76 % of the slots) and the shared-memory data
// pipe (73 %) together, at 6 CTAs per SM; blocks run chunk-minor so that all chunks' look-back fronts advance
// side by side.
//
// The file also compiles as plain C++ under tests/cuda_emul (MB_EMULATE): a CPU execution of the very same
// kernel text, one fiber per CUDA thread, which lets the index arithmetic be checked against the restated
// serializer without a GPU (tests/test_codec_emulated.py).  The product never builds that way.
#pragma once

#ifdef MB_EMULATE
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace mbc {

constexpr int kThreads = 256;
constexpr int kHalf = 16;                              // one 128-bit load
constexpr int kBytesPerThread = 2 * kHalf;             // a thread's span: two loads, one 32-bit mask
constexpr int kTile = kThreads * kBytesPerThread;      // 8192 payload bytes: one span per thread (also raw_fixup's CTA)
constexpr int kWarps = kThreads / 32;
// rle_tiles<S>: a CTA owns S consecutive 8 KiB pieces (a thread: span h at piece h, h = 0 .. S-1), i.e. a tile of
// S * 8 KiB.  Dynamic shared memory: [15] the byte before the tile | [16, 16 + tile) the tile | uint16 list of the
// in-tile offsets of the run starts (+ the end-of-chunk sentinel on the last tile)
constexpr int kMaxSpans = 2;                           // the packed scan carries 16 bits per piece
constexpr int kCopyOff = 16;
__host__ __device__ constexpr int tile_bytes(int spans) { return spans * kTile; }
__host__ __device__ constexpr int list_off(int spans) { return kCopyOff + tile_bytes(spans); }
__host__ __device__ constexpr int list_cap(int) { return kTile; }           // entries the list holds (+ the sentinel): one piece's worst case
__host__ __device__ constexpr int smem_bytes(int spans) { return list_off(spans) + 2 * (list_cap(spans) + 8); }

// per-chunk result record (mb_chunk_info in include/mandelbrot_b200.h)
struct Meta {
    unsigned long long out_len;   // total serialised bytes, header included
    uint32_t runs;
    uint32_t codec;               // 0 raw, 1 RLE
    uint32_t entry_type;          // IndexEntry.Type (DataStorage.cs:41-49): 0 Regular, 1 Never, 2 Immediate
    uint32_t first_value;
    uint32_t pad[2];
};
static_assert(sizeof(Meta) == 32, "mb_chunk_info layout");

// look-back status of one tile: two self-validating 64-bit words (flag in the top two bits)
struct TileStatus {
    unsigned long long count;     // flag | run starts (aggregate of this tile, or inclusive prefix)
    unsigned long long last;      // flag | 1 + position of the last run start (0 = none)
};
constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagInclusive = 2ull << 62, kFlagMask = 3ull << 62;

// A batch of chunks is processed by one set of launches (rle_tiles: blockIdx.x = chunk, raw_fixup: blockIdx.y).
struct Batch {
    const uint8_t *payloads;      // chunk c at payloads + c * payload_stride
    size_t payload_stride;
    uint8_t *out;                 // chunk c at out + c * out_stride (16-byte aligned stride)
    size_t out_stride;
    TileStatus *status;           // per chunk: status + c * ntiles   (zeroed before every call)
    uint32_t *dead;               // dead[c] != 0: some tile of chunk c has seen 5 * runs >= n, RLE is ruled out (zeroed likewise)
    Meta *meta;                   // meta[c]
    size_t n;                     // payload bytes per chunk, < 2^32
    uint32_t ntiles;              // tiles per chunk (of S * 8 KiB for rle_tiles<S>, of 8 KiB for raw_fixup)
    uint32_t n_chunks;
};

// bit 7 of every byte of `cur` that differs from the byte before it (`sh` = the same word one byte earlier)
__device__ __forceinline__ uint32_t differ4(uint32_t cur, uint32_t sh)
{
    const uint32_t x = cur ^ sh;
    return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

// appends the four flags of `h` (bits 7, 15, 23, 31) as a nibble below the nibbles already in `mask`: the
// multiply gathers them in bits 28..31 (the ten partial products land on ten different bits, no carries)
__device__ __forceinline__ uint32_t push4(uint32_t mask, uint32_t h)
{
    return __funnelshift_l(h * 0x00204081u, mask, 4);
}

// index of the highest set bit of m != 0 (one FLO)
__device__ __forceinline__ uint32_t top_bit(uint32_t m)
{
#ifdef MB_EMULATE
    return 31u - (uint32_t)__clz(m);
#else
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
#endif
}

// 32-bit mask: bit j set iff byte j of this thread's 32-byte span starts a run (p[i] != p[i-1];
// i = 0 always starts one).  Fast path for spans that are whole and 16-byte aligned: two 128-bit
// loads, the byte before the span comes from the neighbouring lane.  `prev` = byte before the span.
__device__ __forceinline__ uint32_t span_mask_fast(const uint8_t *__restrict__ p, size_t base, uint4 &v0, uint4 &v1, uint32_t &prev)
{
    v0 = *reinterpret_cast<const uint4 *>(p + base);
    v1 = *reinterpret_cast<const uint4 *>(p + base + kHalf);
    prev = __shfl_up_sync(0xffffffffu, v1.w >> 24, 1);
    if ((threadIdx.x & 31) == 0) prev = base ? (uint32_t)p[base - 1] : 0u;
    // nibbles are pushed from the last word to the first, so that word 0 ends up in bits 0..3
    uint32_t mask = 0;
    mask = push4(mask, differ4(v1.w, __byte_perm(v1.z, v1.w, 0x6543)));
    mask = push4(mask, differ4(v1.z, __byte_perm(v1.y, v1.z, 0x6543)));
    mask = push4(mask, differ4(v1.y, __byte_perm(v1.x, v1.y, 0x6543)));
    mask = push4(mask, differ4(v1.x, __byte_perm(v0.w, v1.x, 0x6543)));
    mask = push4(mask, differ4(v0.w, __byte_perm(v0.z, v0.w, 0x6543)));
    mask = push4(mask, differ4(v0.z, __byte_perm(v0.y, v0.z, 0x6543)));
    mask = push4(mask, differ4(v0.y, __byte_perm(v0.x, v0.y, 0x6543)));
    mask = push4(mask, differ4(v0.x, __byte_perm(prev, v0.x, 0x6540)));
    if (base == 0) mask |= 1u;
    return mask;
}

// the general path (ragged end of the payload, unaligned payload pointer): byte by byte, kept a real
// loop so that the compiler branches around it instead of predicating it into the fast path
__device__ __forceinline__ uint32_t span_mask_slow(const uint8_t *__restrict__ p, size_t base, size_t n, uint4 &v0, uint4 &v1, uint32_t &prev)
{
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t mask = 0;
    prev = 0;
    if (base < n) {
        prev = base ? (uint32_t)p[base - 1] : 0u;
        uint32_t pb = prev;
#pragma unroll 1
        for (int j = 0; j < kBytesPerThread && base + j < n; ++j) {
            const uint32_t c = p[base + j];
            if (c != pb) mask |= 1u << j;
            pb = c;
            w[j >> 2] |= c << (8 * (j & 3));
        }
        if (base == 0) mask |= 1u;
    }
    v0 = make_uint4(w[0], w[1], w[2], w[3]);
    v1 = make_uint4(w[4], w[5], w[6], w[7]);
    return mask;
}

// status words travel through L2, device scope (every word is self-validating: no ordering between them is needed)
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
#ifdef MB_EMULATE
    return *reinterpret_cast<const volatile unsigned long long *>(p);
#else
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
#ifdef MB_EMULATE
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" : : "l"(p), "l"(v) : "memory");
#endif
}

// one look-back status: flag 0 while the two words are unpublished or belong to different publications
__device__ __forceinline__ void ld_pair(const TileStatus *st, uint32_t &flag, uint32_t &count, uint32_t &last)
{
    const unsigned long long sc = ld_status(&st->count), sl = ld_status(&st->last);
    const uint32_t fc = (uint32_t)(sc >> 62), fl = (uint32_t)(sl >> 62);
    flag = fc == fl ? fc : 0u;
    count = (uint32_t)sc;                                      // run counts and positions are below 2^32
    last = (uint32_t)sl;
}

// Decoupled look-back, executed by one whole warp of tile `tile` >= 1: sums the predecessors' aggregates back
// to the nearest published inclusive prefix.  Returns (1 + position of the last run start before this tile) << 32 |
// (run starts before it) in every lane.  A tile publishes its aggregate before it waits for anybody, and the
// blocks of a chunk are dispatched in increasing tile order, so every predecessor a tile waits for is running or
// done (the usual decoupled-look-back assumption).  One round trip to L2 inspects kLook * 32 predecessors: the
// front of inclusive prefixes advances by that many tiles per round trip, which is what bounds a single chunk.
constexpr int kLook = 4;
constexpr uint32_t kRelay = 16;                                // idle tiles that publish an inclusive prefix: one in 16 (< 32 * kLook)
__device__ __noinline__ unsigned long long look_back(const TileStatus *status, uint32_t tile, uint32_t lane)
{
    uint32_t rb = 0, lb = 0;
    int look = (int)tile - 1;                                  // lane l inspects tiles look - l - 32 k, k = 0 .. kLook - 1
    for (;;) {
        uint32_t flag[kLook], cv[kLook], lv[kLook];            // 0 = not published yet (or a torn pair), 1 aggregate, 2 inclusive
#pragma unroll
        for (int k = 0; k < kLook; ++k) {                      // all loads of the round are in flight together
            const int t = look - (int)lane - 32 * k;
            flag[k] = 2u;                                      // "before tile 0": empty inclusive prefix
            cv[k] = 0u;
            lv[k] = 0u;
            if (t >= 0) ld_pair(&status[t], flag[k], cv[k], lv[k]);
        }
        uint32_t c = 0, l = 0;
        bool found = false;                                    // an inclusive prefix among the nearer statuses (warp-uniform)
#pragma unroll
        for (int k = 0; k < kLook; ++k) {
            if (!found) {                                      // statuses beyond the nearest inclusive prefix are not needed
                const int t = look - (int)lane - 32 * k;
                while (flag[k] == 0u) ld_pair(&status[t], flag[k], cv[k], lv[k]);
                const unsigned incl = __ballot_sync(0xffffffffu, flag[k] == 2u);
                // use the window up to and including the nearest inclusive prefix
                const uint32_t stop = incl ? (uint32_t)__ffs(incl) - 1u : 31u;
                if (lane <= stop) {
                    c += cv[k];
                    l = l > lv[k] ? l : lv[k];
                }
                found = incl != 0;
            }
        }
        rb += __reduce_add_sync(0xffffffffu, c);
        const uint32_t lm = __reduce_max_sync(0xffffffffu, l);
        lb = lb > lm ? lb : lm;
        if (found) break;
        look -= 32 * kLook;
    }
    return ((unsigned long long)lb << 32) | rb;                // a call, not inlined: its registers stay out of the kernel's budget
}

// byte s of the record stream of a tile (records of the runs that end in it, back to back): the general form,
// used for the few bytes at the two ends of the tile's output range that do not fill an aligned 16-byte word
__device__ __forceinline__ uint32_t stream_byte(uint32_t s, uint32_t skip, uint32_t before, const uint16_t *s_pos, const uint8_t *cp)
{
    const uint32_t r = s / 5u, f = s - 5u * r, e = r + skip;
    const uint32_t pos = s_pos[e];
    const uint32_t pm = e ? (uint32_t)s_pos[e - 1] : before;
    const uint32_t len = pos - pm;                                     // BitConverter.GetBytes(runLength): little-endian
    return f < 4u ? (len >> (8u * f)) & 0xffu : (uint32_t)cp[pos];    // runValue (DataChunkSerializer.cs:84,98)
}

// The single pass.  grid = (n_chunks, tiles [, more tiles]): blockIdx.x = chunk runs fastest, so the blocks in
// flight at any time belong to all chunks of the batch and as many look-back fronts advance side by side, while the
// tiles of one chunk are still dispatched in increasing order.  kThreads threads, smem_bytes(S) of dynamic shared memory.
template <int S>
__global__ void __launch_bounds__(kThreads, S == 1 ? 8 : 6) rle_tiles(const Batch b)
{
    static_assert(S >= 1 && S <= kMaxSpans, "spans per thread");
    constexpr uint32_t kTileS = (uint32_t)tile_bytes(S);
#ifdef MB_EMULATE
    uint8_t *smem = EMUL_DYNAMIC_SMEM;
#else
    extern __shared__ __align__(16) uint8_t smem[];
#endif
    __shared__ uint32_t w_cnt[kWarps], w_last[kWarps];
    __shared__ uint32_t s_runs_before, s_last_before, s_ruled_out;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t chunk = blockIdx.x, tile = blockIdx.z * gridDim.y + blockIdx.y;
    if (tile >= b.ntiles) return;                                      // padding of a three-dimensional grid, behind every real tile
    const uint8_t *__restrict__ p = b.payloads + chunk * b.payload_stride;
    uint8_t *__restrict__ out = b.out + chunk * b.out_stride;
    TileStatus *status = b.status + (size_t)chunk * b.ntiles;
    const uint32_t n = (uint32_t)b.n;
    const uint32_t tile_base = tile * kTileS;                          // < n
    const uint32_t span = tid * (uint32_t)kBytesPerThread;             // offset of my 32 bytes inside each 8 KiB piece
    const bool is_last_tile = tile + 1u == b.ntiles;
    // a tile is `regular` when all of it lies inside the payload and the payload is 16-byte aligned
    const bool regular = (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && n - tile_base >= kTileS;

    // RLE already ruled out by an earlier tile?  (a hint, read once per CTA so that all threads agree -- and first,
    // so that its trip to L2 overlaps the tile's; a stale 0 only costs the work, and whoever set it saw
    // 5 * runs >= n, so the chunk ends up Raw whatever is skipped)
    uint32_t ruled_out_hint = 0;
    if (tid == 0) ruled_out_hint = *reinterpret_cast<const volatile uint32_t *>(b.dead + chunk);

    // ---- 1. the tile, once: run-start masks of my spans, the bytes themselves go to shared memory ---------
    uint32_t mask[S];
#pragma unroll
    for (int h = 0; h < S; ++h) {
        uint4 v0, v1;
        uint32_t prev;
        const size_t base = (size_t)tile_base + (uint32_t)(h * kTile) + span;
        mask[h] = regular ? span_mask_fast(p, base, v0, v1, prev) : span_mask_slow(p, base, n, v0, v1, prev);
        *reinterpret_cast<uint4 *>(smem + kCopyOff + h * kTile + span) = v0;
        *reinterpret_cast<uint4 *>(smem + kCopyOff + h * kTile + span + kHalf) = v1;
        if (h == 0 && tid == 0) smem[kCopyOff - 1] = (uint8_t)prev;
    }
    // counts of the pieces packed 16 bits each (a piece holds at most 8192 starts), scanned together
    uint32_t cnt = 0, mylast = 0;                                      // mylast: 1 + in-tile offset of my last start
#pragma unroll
    for (int h = 0; h < S; ++h) {
        cnt |= (uint32_t)__popc(mask[h]) << (16 * h);
        if (mask[h]) mylast = (uint32_t)(h * kTile) + span + 32u - (uint32_t)__clz(mask[h]);
    }

    // CTA-wide exclusive scan of the counts, totals, and the last start of the tile
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t c = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += c;
    }
    const uint32_t wl = __reduce_max_sync(0xffffffffu, mylast);
    if (lane == 31) w_cnt[warp] = inc;
    if (lane == 0) w_last[warp] = wl;
    if (tid == 0) s_ruled_out = ruled_out_hint;
    __syncthreads();
    // the eight warp totals are scanned by the first eight lanes of every warp
    uint32_t wsum = lane < (uint32_t)kWarps ? w_cnt[lane] : 0u;
#pragma unroll
    for (int o = 1; o < kWarps; o <<= 1) {
        const uint32_t c = __shfl_up_sync(0xffffffffu, wsum, o);
        if (lane >= (uint32_t)o) wsum += c;
    }
    const uint32_t tot_packed = __shfl_sync(0xffffffffu, wsum, kWarps - 1);
    const uint32_t excl = __shfl_sync(0xffffffffu, wsum, warp) - w_cnt[warp] + inc - cnt;   // per piece: starts before my span
    const uint32_t totlast = __reduce_max_sync(0xffffffffu, lane < (uint32_t)kWarps ? w_last[lane] : 0u);
    uint32_t tot = 0, rank[S];                                         // rank[h]: starts of this tile before my span of piece h
#pragma unroll
    for (int h = 0; h < S; ++h) {
        rank[h] = tot + ((excl >> (16 * h)) & 0xffffu);
        tot += (tot_packed >> (16 * h)) & 0xffffu;
    }
    const unsigned long long agg_last = totlast ? (unsigned long long)tile_base + totlast : 0ull;   // 1 + position in the chunk

    // ---- 2. publish the aggregate at once (tile 0: its inclusive prefix) ---------------------------------
    if (tid == 0) {
        const unsigned long long flag = tile ? kFlagAggregate : kFlagInclusive;
        st_status(&status[tile].count, flag | tot);
        st_status(&status[tile].last, flag | agg_last);
    }
    // a tile without a run start closes no run (unless it is the last one): it only passes the prefix on; so does
    // a tile that found RLE ruled out.  Such a tile needs no prefix itself: its aggregate is published, and only
    // every kRelay-th of them also looks back and publishes an inclusive prefix, which keeps everybody's look-back
    // within one window without making every idle tile wait for a trip to L2.
    const bool emits = (tot != 0 && s_ruled_out == 0) || is_last_tile;
    if (!emits && (tile & (kRelay - 1u)) != kRelay - 1u) return;      // uniform

    // The list holds list_cap(S) = 8192 entries.  A tile of two pieces with more run starts than that (runs
    // shorter than two bytes on average) is emitted piece by piece, in S rounds; every other tile in one.
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(smem + list_off(S));
    const uint8_t *cp = smem + kCopyOff - 1;                           // cp[o] = the byte before in-tile offset o
    const uint32_t n_rounds = (S > 1 && emits && tot > (uint32_t)list_cap(S)) ? (uint32_t)S : 1u;
    uint32_t done = 0;                                                 // entries of the rounds before this one
    uint32_t runs_before = 0, before = 0;
#pragma unroll 1
    for (uint32_t r = 0; r < n_rounds; ++r) {
        const bool closes_chunk = is_last_tile && r + 1u == n_rounds;
        uint32_t tot_r = tot;                                          // entries of this round
        if (n_rounds > 1) tot_r = (tot_packed >> (16u * r)) & 0xffffu;

        // ---- 3. the list of run-start offsets (does not need the prefix: predecessors get time to publish) ----
        if (emits) {
#pragma unroll
            for (int h = 0; h < S; ++h) {
                if (n_rounds > 1 && (uint32_t)h != r) continue;
                // filled from the top (the highest set bit is one FLO away), two entries per 32-bit store: halves the
                // scattered shared-memory stores, which is what this kernel has the least of to spare
                uint32_t e = rank[h] - done + (uint32_t)__popc(mask[h]);   // one past my top entry
                const uint32_t off = (uint32_t)(h * kTile) + span, off2 = off * 0x10001u;
                uint32_t m = mask[h];
                if (m) {
                    if (e & 1u) {                                      // my top entry is the low half of a word whose high half is the next span's
                        const uint32_t j = top_bit(m);
                        s_pos[--e] = (uint16_t)(off + j);
                        m ^= 1u << j;
                    }
                    uint32_t *qw = reinterpret_cast<uint32_t *>(s_pos + e);
#pragma unroll 1
                    while (m & (m - 1u)) {                             // a short divergent loop: runs cluster
                        const uint32_t j1 = top_bit(m);
                        m ^= 1u << j1;
                        const uint32_t j0 = top_bit(m);
                        m ^= 1u << j0;
                        *--qw = off2 + ((j1 << 16) | j0);
                    }
                    if (m) reinterpret_cast<uint16_t *>(qw)[-1] = (uint16_t)(off + top_bit(m));   // my lowest entry, the high half of its word
                }
            }
            // the end of the chunk closes its last run like one more start would
            if (closes_chunk && tid == 0) s_pos[tot_r] = (uint16_t)(n - tile_base);
        }

        // ---- look-back by warp 0 (first round), inclusive prefix published for the successors ----------------
        if (r == 0 && warp == 0) {
            uint32_t rb = 0, lb = 0;
            if (tile) {
                const unsigned long long lr = look_back(status, tile, lane);
                rb = (uint32_t)lr;
                lb = (uint32_t)(lr >> 32);
                if (lane == 0) {
#ifdef MB_EMULATE
                    // the emulator runs the CTAs one after the other, so every predecessor would be inclusive:
                    // leave most tiles at their aggregate to exercise the summation over several windows
                    if (!mbc_emul_keep_aggregate(tile))
#endif
                    {
                        st_status(&status[tile].count, kFlagInclusive | ((unsigned long long)rb + tot));
                        st_status(&status[tile].last, kFlagInclusive | (agg_last ? agg_last : (unsigned long long)lb));
                    }
                }
            }
            if (lane == 0) { s_runs_before = rb; s_last_before = lb; }
        }
        if (!emits) return;                                            // uniform: nobody is left behind at the barrier
        __syncthreads();
        if (r == 0) {
            runs_before = s_runs_before;
            before = s_last_before - 1u - tile_base;                   // in-tile offset of the start before entry 0 (wraps below 0)
            // the last tile knows the run count: pick the codec, write header + length + classification
            if (is_last_tile && tid == 0) {
                Meta *meta = b.meta + chunk;
                const uint32_t runs = runs_before + tot;
                // DataChunk.cs:181-204: the first serializer (Raw) is kept unless RLE is strictly shorter
                const bool rle = (unsigned long long)runs * 5ull < (unsigned long long)n;
                meta->runs = runs;
                meta->codec = rle ? 1u : 0u;
                meta->out_len = 1ull + (rle ? (unsigned long long)runs * 5ull : (unsigned long long)n);
                // one run of 0s = IsNeverChunk, one run of 1s = IsImmediateChunk (DataChunk.cs:82,87);
                // Never is tested first (DataStorage.cs:74-77) but the two cannot both hold
                const uint32_t first = p[0];
                meta->first_value = first;
                meta->entry_type = (runs == 1 && first == 0) ? 1u : (runs == 1 && first == 1) ? 2u : 0u;
                out[0] = rle ? 0x01 : 0x00;        // Serializer.SerializeCode, DataChunkSerializer.cs:20-23
            }
        } else {
            before = s_last_before;                                    // left by the round before (in-tile offset already)
        }

        // ---- 4. records of the runs that END at this round's starts ------------------------------------------
        // List entry e (e = 0 .. tot_r-1, plus the sentinel e = tot_r when the chunk ends here) is the start with
        // global index g = runs_r + e and closes run g - 1: length = this start minus the previous one, value = the
        // byte before this start.  Record k lives at out[1 + 5k .. 1 + 5k + 5).  Once the runs before make RLE
        // impossible (5 * runs >= n keeps Raw, and runs only grow) nothing is emitted any more.
        const uint32_t runs_r = runs_before + done;
        if ((unsigned long long)runs_r * 5ull >= (unsigned long long)n) {
            if (tid == 0) *reinterpret_cast<volatile uint32_t *>(b.dead + chunk) = 1u;
            return;
        }
        const uint32_t skip = runs_r ? 0u : 1u;                        // start 0 of the chunk closes nothing
        const uint32_t first_rec = runs_r ? runs_r - 1u : 0u;          // first record this round writes
        const uint32_t n_ent = tot_r + (closes_chunk ? 1u : 0u);
        const uint32_t n_rec = n_ent - skip;                           // tile 0 holds start 0, so n_ent >= skip there
        const unsigned long long out_lo = 1ull + 5ull * first_rec;     // <= n: the round's record stream starts here
        uint32_t slen = 5u * n_rec;                                    // ... and is this long;
        if ((unsigned long long)slen > (unsigned long long)n + 1ull - out_lo)
            slen = (uint32_t)((unsigned long long)n + 1ull - out_lo);  // bytes beyond n + 1 can never be part of an RLE result
        uint8_t *dst = out + out_lo;
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
        // aligned words that lie completely inside the stream: word q covers stream bytes [16q - shift, 16q - shift + 16)
        const uint32_t q_first = shift ? 1u : 0u, q_end = (shift + slen) >> 4;
        for (uint32_t q = q_first + tid; q < q_end; q += kThreads) {
            const uint32_t s0 = 16u * q - shift;
            const uint32_t ra = s0 / 5u, phase = s0 - 5u * ra;         // first record of the word, and where the word starts in it
            const uint32_t ea = ra + skip;
            const uint16_t *e = s_pos + ea;
            uint32_t pm = e[-1];                                       // (e = 0: the two bytes before the list, in bounds, unused)
            if (ea == 0) pm = before;
            const uint32_t p0 = e[0], p1 = e[1], p2 = e[2], p3 = e[3]; // a whole word always touches four records
            const uint32_t l0 = p0 - pm, l1 = p1 - p0, l2 = p2 - p1, l3 = p3 - p2;
            const uint32_t c0 = cp[p0], c1 = cp[p1], c2 = cp[p2], c3 = cp[p3];
            // the 20 bytes of the four records: l0 | c0 l1[0..2] | l1[3] c1 l2[0..1] | l2[2..3] c2 l3[0] | l3[1..3] c3
            const uint32_t w0 = l0;
            const uint32_t w1 = __byte_perm(c0, l1, 0x6540);
            const uint32_t w2 = __byte_perm(__byte_perm(l1, c1, 0x0043), l2, 0x5410);
            const uint32_t w3 = __byte_perm(__byte_perm(l2, c2, 0x0432), l3, 0x4210);
            const uint32_t w4 = __byte_perm(l3, c3, 0x4321);
            const uint32_t sel = 0x3210u + 0x1111u * phase;            // bytes phase .. phase + 3 of a word pair
            uint4 w;
            w.x = __byte_perm(w0, w1, sel);
            w.y = __byte_perm(w1, w2, sel);
            w.z = __byte_perm(w2, w3, sel);
            w.w = __byte_perm(w3, w4, sel);
            *reinterpret_cast<uint4 *>(dst + s0) = w;
        }
        // the bytes before the first and after the last whole word (< 16 each; < 32 when there is no whole word)
        if (warp == kWarps - 1) {
            const bool any = q_end > q_first;
            const uint32_t head_end = any ? 16u * q_first - shift : slen;
            const uint32_t tail_lo = any ? 16u * q_end - shift : slen;
            if (lane < head_end) dst[lane] = (uint8_t)stream_byte(lane, skip, before, s_pos, cp);
            if (tail_lo + lane < slen) dst[tail_lo + lane] = (uint8_t)stream_byte(tail_lo + lane, skip, before, s_pos, cp);
        }
        if (r + 1u < n_rounds) {                                       // (never taken with one piece per tile)
            // the next round overwrites the list: leave it the last start of this one (or what this round had)
            __syncthreads();
            if (tid == 0) s_last_before = tot_r ? (uint32_t)s_pos[tot_r - 1u] : before;
            done += tot_r;
            __syncthreads();
        }
    }
}

// Second kernel, does work only for chunks that ended up Raw: RawSerializer
// (DataChunkSerializer.cs:34-37), out[1 + i] = p[i].  The destination is 1 byte off alignment, so
// aligned 16-byte words are built from two neighbouring source words.  grid = (pieces / kRawPieces, n_chunks):
// a CTA owns kRawPieces consecutive 8 KiB pieces, so that the launch costs little when no chunk is Raw (the
// usual case: a block that finds its chunk RLE exits at once, and there are few blocks).
constexpr int kRawPieces = 8;
__global__ void __launch_bounds__(kThreads) raw_fixup(const Batch b)
{
    const Meta *__restrict__ meta = b.meta + blockIdx.y;
    if (meta->codec != 0) return;
    const uint8_t *__restrict__ p = b.payloads + blockIdx.y * b.payload_stride;
    uint8_t *__restrict__ out = b.out + blockIdx.y * b.out_stride;
    const size_t n = b.n;
#pragma unroll 1
    for (int piece = 0; piece < kRawPieces; ++piece) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            // 16-byte pieces, consecutive lanes on consecutive pieces (coalesced both ways)
            const size_t base = ((size_t)blockIdx.x * kRawPieces + piece) * kTile + ((size_t)half * kThreads + threadIdx.x) * kHalf;
            if (base >= n) return;
            const bool fast = base + kHalf <= n && ((reinterpret_cast<uintptr_t>(p + base) & 15u) == 0) &&
                              ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
            if (fast) {
                const uint4 v = *reinterpret_cast<const uint4 *>(p + base);
                const uint32_t before = base ? (uint32_t)p[base - 1] : 0u;         // header byte (0x00 = raw) for the very first word
                uint4 w;
                w.x = __byte_perm(before, v.x, 0x6540);       // {before.b0, v.x.b0, v.x.b1, v.x.b2}
                w.y = __byte_perm(v.x, v.y, 0x6543);          // {v.x.b3, v.y.b0, v.y.b1, v.y.b2}
                w.z = __byte_perm(v.y, v.z, 0x6543);
                w.w = __byte_perm(v.z, v.w, 0x6543);
                *reinterpret_cast<uint4 *>(out + base) = w;   // covers out[base .. base+15] = {p[base-1], p[base .. base+14]}
                if (base + kHalf == n) out[n] = (uint8_t)(v.w >> 24);              // last payload byte
            } else {
                // a word-building neighbour to the left leaves out[base] (= p[base-1]) to this thread
                if (base) out[base] = p[base - 1];
                for (int j = 0; j < kHalf && base + j < n; ++j) out[1 + base + j] = p[base + j];
            }
        }
    }
}

}  // namespace mbc
