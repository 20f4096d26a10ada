// DEV-ONLY (A/B): the round-2 segment kernel, kept until the tile kernel of chunk_codec.cuh is measured against it.
// chunk_codec.cuh -- DataChunk serialisation on the device, HBM-bound byte work.
//
// Replaces DataChunk.Serialize (DistributedMandelbrot/DataChunk.cs:173-206), which serialises the
// 16 MiB payload TWICE on one CPU thread (RawSerializer, DataChunkSerializer.cs:29-49, and
// RLESerializer, :51-100, each into a SizeCountStream) to pick the shorter and then a third time
// to write it.  Output bytes are identical to the reference's:
//     byte 0 = codec: 0x00 raw -> the n payload bytes ; 0x01 RLE -> (uint32 LE runLength, uint8
//     value) records, one per maximal run.  Raw wins ties (`<` with Raw first, DataChunk.cs:188).
//
// ONE pass over HBM for every chunk that ends up RLE (rle_single_pass): run starts found with SIMD
// byte compares, a decoupled look-back scan across the chunk's segments carrying (number of run
// starts, position of the last run start), complete 5-byte records assembled in shared memory and
// written with 16-byte stores.  A record is emitted
// by the CTA in whose tile the run ENDS -- there both its length (this start minus the previous
// start, which the scan carries across tiles however long the run was) and its value (the byte
// before the new start) are known, so no second pass and no look-ahead are needed; the chunk's last
// run is closed by the last tile.  Records go straight to their final place; a tile stops emitting
// once the runs before it already make RLE impossible (5 * runs >= n).  Only a chunk that ends up
// Raw takes a second kernel (raw_fixup), which copies the payload behind the header byte.
// Traffic: n read + serialised bytes written for RLE chunks (the algorithmic minimum).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mbc_seg {

constexpr int kThreads = 256;
constexpr int kHalf = 16;                              // one 128-bit load
constexpr int kBytesPerThread = 2 * kHalf;             // two loads in flight per thread: 32 KiB per SM at 4 CTAs/SM
constexpr int kTile = kThreads * kBytesPerThread;      // 8192 payload bytes per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kRound = 4096;                           // records assembled per round (a denser tile takes two rounds)
constexpr int kStageRecBytes = 5 * kRound + 32;        // staging of one round's records (+ alignment slack)
constexpr int kStageBytes = kStageRecBytes + 2 * (kTile + 8);   // ... followed by the tile's run-start offsets (uint16)

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

// A batch of chunks is processed by one set of launches: blockIdx.y = chunk.
struct Batch {
    const uint8_t *payloads;      // chunk c at payloads + c * payload_stride
    size_t payload_stride;
    uint8_t *out;                 // chunk c at out + c * out_stride (16-byte aligned stride)
    size_t out_stride;
    TileStatus *status;           // per chunk: status + c * nsegs   (zeroed before every call)
    Meta *meta;                   // meta[c]
    size_t n;                     // payload bytes per chunk
    uint32_t ntiles;              // 8 KiB tiles per chunk
    uint32_t seg_tiles;           // tiles per segment (one CTA of rle_single_pass)
    uint32_t nsegs;               // segments per chunk
};

// 4-bit mask of the non-zero bytes of x (x = a ^ b: the bytes that differ), branch-free SWAR:
// bit 7 of every non-zero byte, then the four bits gathered with one multiply
__device__ __forceinline__ uint32_t nonzero4(uint32_t x)
{
    const uint32_t h = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    return (((h >> 7) * 0x01020408u) >> 24) & 0xfu;
}

// 16-bit mask of the run starts inside 16 payload bytes v, `prev` being the byte before them
__device__ __forceinline__ uint32_t starts16(const uint4 &v, uint32_t prev)
{
    // the same 16 bytes shifted up by one byte, `prev` entering at the bottom
    const uint32_t s0 = __byte_perm(prev, v.x, 0x6540), s1 = __byte_perm(v.x, v.y, 0x6543);
    const uint32_t s2 = __byte_perm(v.y, v.z, 0x6543), s3 = __byte_perm(v.z, v.w, 0x6543);
    return nonzero4(v.x ^ s0) | (nonzero4(v.y ^ s1) << 4) | (nonzero4(v.z ^ s2) << 8) | (nonzero4(v.w ^ s3) << 12);
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
    uint32_t mask = starts16(v0, prev) | (starts16(v1, v0.w >> 24) << kHalf);
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

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// CTA-wide (sum, max) of a per-thread pair; every thread receives the totals
__device__ __forceinline__ void block_sum_max(uint32_t c, uint32_t l, uint32_t *w_cnt, uint32_t *w_last, uint32_t &tot, uint32_t &totlast)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        const uint32_t l2 = __shfl_xor_sync(0xffffffffu, l, o);
        l = l > l2 ? l : l2;
    }
    __syncthreads();                       // the arrays may still be read by the previous user
    if (lane == 0) { w_cnt[warp] = c; w_last[warp] = l; }
    __syncthreads();
    tot = 0;
    totlast = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        tot += w_cnt[w];
        totlast = totlast > w_last[w] ? totlast : w_last[w];
    }
}

// The single pass.  grid = (segments, n_chunks), kThreads threads, kStageBytes of dynamic shared memory.
// A CTA owns a SEGMENT of b.seg_tiles consecutive 8 KiB tiles:
//   phase 1  streams the segment once from HBM and counts its run starts (+ position of the last one);
//   look-back  one decoupled look-back per SEGMENT (a few hundred per chunk, not thousands of tiles:
//            the chain is short and several chunks are in flight at once);
//   phase 2  walks the segment's tiles again -- now L2 hits, the segment was just read -- and emits the
//            records, carrying the running (runs, last start) in shared memory from tile to tile.
// Segment ids follow blockIdx.x: blocks are dispatched in increasing linear index, so every
// predecessor a segment waits for is running or done (the usual decoupled-look-back assumption).
__global__ void __launch_bounds__(kThreads, 4) rle_single_pass(const Batch b)
{
    extern __shared__ __align__(16) uint8_t stage[];
    __shared__ uint32_t w_cnt[kWarps], w_last[kWarps];
    __shared__ uint32_t s_runs_before, s_last_before;

    const uint8_t *__restrict__ p = b.payloads + blockIdx.y * b.payload_stride;
    uint8_t *__restrict__ out = b.out + blockIdx.y * b.out_stride;
    TileStatus *status = b.status + (size_t)blockIdx.y * b.nsegs;
    const size_t n = b.n;
    const uint32_t seg = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile0 = seg * b.seg_tiles;
    const uint32_t ntl = b.ntiles - tile0 < b.seg_tiles ? b.ntiles - tile0 : b.seg_tiles;

    // a tile is `regular` when all of it lies inside the payload and the payload is 16-byte aligned
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;

    // ---- phase 1: run starts of the whole segment --------------------------------------------------
    uint32_t seg_cnt = 0, seg_last = 0;
    uint32_t busy = 0;                                   // bit k: my span of tile k holds a run start
    // regular tiles first, unrolled so that several 128-bit loads per thread are in flight (this loop is
    // the HBM stream of the kernel); a ragged last tile / unaligned payload takes the general path
    uint32_t nreg = 0;
    if (aligned) {
        const size_t whole = n / kTile;                  // tiles that lie completely inside the payload
        nreg = whole > tile0 ? (uint32_t)(whole - tile0 < ntl ? whole - tile0 : ntl) : 0u;
    }
#pragma unroll 4
    for (uint32_t k = 0; k < nreg; ++k) {
        uint4 v0, v1;
        uint32_t pv;
        const size_t base = (size_t)(tile0 + k) * kTile + (size_t)threadIdx.x * kBytesPerThread;
        const uint32_t m = span_mask_fast(p, base, v0, v1, pv);
        if (m) {
            seg_cnt += __popc(m);
            seg_last = (uint32_t)base + (31u - __clz(m)) + 1u;                // positions grow with k
            busy |= 1u << k;
        }
    }
    for (uint32_t k = nreg; k < ntl; ++k) {
        uint4 v0, v1;
        uint32_t pv;
        const size_t base = (size_t)(tile0 + k) * kTile + (size_t)threadIdx.x * kBytesPerThread;
        const uint32_t m = span_mask_slow(p, base, n, v0, v1, pv);
        if (m) {
            seg_cnt += __popc(m);
            seg_last = (uint32_t)base + (31u - __clz(m)) + 1u;
            busy |= 1u << k;
        }
    }
    uint32_t seg_tot, seg_totlast;
    block_sum_max(seg_cnt, seg_last, w_cnt, w_last, seg_tot, seg_totlast);

    // ---- decoupled look-back over segments: (runs before this segment, last start before it) -------
    if (warp == 0) {
        unsigned long long runs_before = 0, last_before = 0;
        if (seg == 0) {
            if (lane == 0) {
                st_status(&status[0].count, kFlagInclusive | seg_tot);
                st_status(&status[0].last, kFlagInclusive | seg_totlast);
            }
        } else {
            if (lane == 0) {
                st_status(&status[seg].count, kFlagAggregate | seg_tot);
                st_status(&status[seg].last, kFlagAggregate | seg_totlast);
            }
            int look = (int)seg - 1;                               // lane k inspects segment look - k
            for (;;) {
                const int t = look - (int)lane;
                unsigned long long sc = kFlagInclusive, sl = kFlagInclusive;     // "before segment 0": empty inclusive prefix
                if (t >= 0) {
                    do {
                        sc = ld_status(&status[t].count);
                        sl = ld_status(&status[t].last);
                    } while ((sc & kFlagMask) == 0 || (sc & kFlagMask) != (sl & kFlagMask));
                }
                const bool inclusive = (sc & kFlagMask) == kFlagInclusive;
                const unsigned incl_mask = __ballot_sync(0xffffffffu, inclusive);
                // use the window up to and including the nearest inclusive prefix
                const int stop = incl_mask ? __ffs(incl_mask) - 1 : 31;
                unsigned long long c = (int)lane <= stop ? (sc & ~kFlagMask) : 0ull;
                unsigned long long l = (int)lane <= stop ? (sl & ~kFlagMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    c += __shfl_down_sync(0xffffffffu, c, o);
                    const unsigned long long l2 = __shfl_down_sync(0xffffffffu, l, o);
                    l = l > l2 ? l : l2;
                }
                c = __shfl_sync(0xffffffffu, c, 0);
                l = __shfl_sync(0xffffffffu, l, 0);
                runs_before += c;
                last_before = last_before > l ? last_before : l;
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0) {
                const unsigned long long li = seg_totlast ? (unsigned long long)seg_totlast : last_before;
                st_status(&status[seg].count, kFlagInclusive | (runs_before + seg_tot));
                st_status(&status[seg].last, kFlagInclusive | li);
            }
        }
        if (lane == 0) { s_runs_before = (uint32_t)runs_before; s_last_before = (uint32_t)last_before; }
    }
    __syncthreads();
    const bool last_seg = (seg + 1 == b.nsegs);
    uint32_t runs_before = s_runs_before, last_before = s_last_before;        // carried in registers, identical in every thread
    const uint32_t runs_total = runs_before + seg_tot;                        // the chunk's run count, on the last segment

    // ---- phase 2: records of the runs that END in each tile of the segment -------------------------
    // run (g - 1) for every start with global index g >= 1, plus the chunk's last run on its last tile.
    // Record r lives at out[1 + 5r .. 1 + 5r + 5).  Tiles from which on RLE is already ruled out
    // (5 * runs >= n keeps Raw, and runs only grow) are skipped, and so are tiles without a run start.
    for (uint32_t k = 0; k < ntl; ++k) {
        if ((unsigned long long)runs_before * 5ull >= (unsigned long long)n) break;
        const uint32_t tile = tile0 + k;
        const bool is_last_tile = (tile + 1 == b.ntiles);
        // also the barrier that releases the staging buffer and w_* of the previous tile
        if (!__syncthreads_or((busy >> k) & 1u) && !is_last_tile) continue;
        uint4 v0, v1;
        uint32_t prev;
        const size_t base = (size_t)tile * kTile + (size_t)threadIdx.x * kBytesPerThread;
        const bool regular = aligned && (size_t)(tile + 1) * kTile <= n;
        uint32_t mask = regular ? span_mask_fast(p, base, v0, v1, prev) : span_mask_slow(p, base, n, v0, v1, prev);   // L2 hits
        const uint32_t cnt = __popc(mask);
        const uint32_t mylast = cnt ? (uint32_t)base + (31u - __clz(mask)) + 1u : 0u;
        // CTA-wide scan of (count, last start): inclusive in the warp, then across warps
        uint32_t inc = cnt, lastinc = mylast;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t c = __shfl_up_sync(0xffffffffu, inc, o), l = __shfl_up_sync(0xffffffffu, lastinc, o);
            if (lane >= (uint32_t)o) { inc += c; lastinc = lastinc > l ? lastinc : l; }
        }
        if (lane == 31) { w_cnt[warp] = inc; w_last[warp] = lastinc; }
        __syncthreads();
        uint32_t wbase = 0, tot = 0, totlast = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = w_cnt[w], l = w_last[w];
            if ((uint32_t)w < warp) wbase += c;
            tot += c;
            totlast = totlast > l ? totlast : l;
        }
        const uint32_t rank = wbase + inc - cnt;                       // starts of this tile before my span

        // a) every thread appends the in-tile offsets of its run starts to a list (a short divergent loop:
        //    runs cluster, a lane may hold ten starts while its neighbours hold none)
        uint16_t *s_pos = reinterpret_cast<uint16_t *>(stage + kStageRecBytes);
        {
            uint32_t e = rank;
            const uint32_t off = threadIdx.x * kBytesPerThread;
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                s_pos[e++] = (uint16_t)(off + j);
            }
        }
        __syncthreads();
        // b) the list is turned into records by ALL threads, one list entry each (balanced, no divergence):
        //    entry e is the start with global index g = runs_before + e and closes run g - 1, whose length
        //    is this start minus the previous one and whose value is the byte before this start
        const size_t tile_base = (size_t)tile * kTile;
        const uint32_t skip = runs_before ? 0u : 1u;                               // start 0 of the chunk closes nothing
        const uint32_t first_rec = runs_before ? runs_before - 1u : 0u;           // first record this tile writes
        const uint32_t n_rec = (tot > skip ? tot - skip : 0u) + (is_last_tile ? 1u : 0u);
        for (uint32_t r0 = 0; r0 < n_rec; r0 += kRound) {
            const uint32_t r1 = r0 + kRound < n_rec ? r0 + kRound : n_rec;        // records [r0, r1) of this tile in this round
            const size_t out_lo = 1 + ((size_t)first_rec + r0) * 5;               // their byte range in the output
            const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(out + out_lo) & 15u);   // stage congruent to out mod 16
            for (uint32_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
                uint32_t len, val;
                const uint32_t e = r + skip;                                       // list entry that closes record r
                if (e < tot) {
                    const uint32_t pos = (uint32_t)tile_base + s_pos[e];
                    const uint32_t before = e ? (uint32_t)tile_base + s_pos[e - 1] : last_before - 1u;
                    len = pos - before;                                            // BitConverter.GetBytes(runLength): little-endian
                    val = p[pos - 1];                                              // runValue (DataChunkSerializer.cs:84,98): an L1/L2 hit
                } else {
                    // the chunk's last run: from the last start anywhere to n
                    const uint32_t last_start = (totlast ? totlast : last_before) - 1u;
                    len = (uint32_t)n - last_start;
                    val = p[n - 1];
                }
                uint8_t *rec = stage + shift + (size_t)(r - r0) * 5;
                rec[0] = (uint8_t)len;
                rec[1] = (uint8_t)(len >> 8);
                rec[2] = (uint8_t)(len >> 16);
                rec[3] = (uint8_t)(len >> 24);
                rec[4] = (uint8_t)val;
            }
            __syncthreads();
            // copy out: bytes beyond n can never be part of an RLE result (5 * runs < n) and are dropped
            size_t nbytes = (size_t)(r1 - r0) * 5;
            if (out_lo + nbytes > n + 1) nbytes = out_lo < n + 1 ? n + 1 - out_lo : 0;
            uint8_t *dst = out + out_lo;
            const uint8_t *src = stage + shift;
            const size_t head = nbytes < ((16u - shift) & 15u) ? nbytes : ((16u - shift) & 15u);
            for (size_t q = threadIdx.x; q < head; q += kThreads) dst[q] = src[q];                 // up to the first 16-byte boundary
            const size_t body = (nbytes - head) / 16;
            for (size_t q = threadIdx.x; q < body; q += kThreads)
                reinterpret_cast<uint4 *>(dst + head)[q] = reinterpret_cast<const uint4 *>(src + head)[q];
            for (size_t q = head + body * 16 + threadIdx.x; q < nbytes; q += kThreads) dst[q] = src[q];
            if (r1 < n_rec) __syncthreads();                                       // the next round reuses the staging area
        }
        runs_before += tot;                                                        // carry the prefix to the next tile
        if (totlast) last_before = totlast;
    }

    // the last segment knows the run count: pick the codec, write header + length + classification
    if (last_seg && threadIdx.x == 0) {
        Meta *meta = b.meta + blockIdx.y;
        const uint32_t runs = runs_total;
        // DataChunk.cs:181-204: the first serializer (Raw) is kept unless RLE is strictly shorter
        const bool rle = (unsigned long long)runs * 5ull < (unsigned long long)n;
        meta->runs = runs;
        meta->codec = rle ? 1u : 0u;
        meta->out_len = 1ull + (rle ? (unsigned long long)runs * 5ull : (unsigned long long)n);
        // one run of 0s = IsNeverChunk, one run of 1s = IsImmediateChunk (DataChunk.cs:82,87);
        // Never is tested first (DataStorage.cs:74-77) but the two cannot both hold
        const uint32_t v0 = p[0];
        meta->first_value = v0;
        meta->entry_type = (runs == 1 && v0 == 0) ? 1u : (runs == 1 && v0 == 1) ? 2u : 0u;
        out[0] = rle ? 0x01 : 0x00;        // Serializer.SerializeCode, DataChunkSerializer.cs:20-23
    }
}

// Second kernel, does work only for chunks that ended up Raw: RawSerializer
// (DataChunkSerializer.cs:34-37), out[1 + i] = p[i].  The destination is 1 byte off alignment, so
// aligned 16-byte words are built from two neighbouring source words.
__global__ void __launch_bounds__(kThreads) raw_fixup(const Batch b)
{
    const Meta *__restrict__ meta = b.meta + blockIdx.y;
    if (meta->codec != 0) return;
    const uint8_t *__restrict__ p = b.payloads + blockIdx.y * b.payload_stride;
    uint8_t *__restrict__ out = b.out + blockIdx.y * b.out_stride;
    const size_t n = b.n;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        // 16-byte pieces, consecutive lanes on consecutive pieces (coalesced both ways)
        const size_t base = (size_t)blockIdx.x * kTile + ((size_t)half * kThreads + threadIdx.x) * kHalf;
        if (base >= n) return;
        const bool fast = base + kHalf <= n && ((reinterpret_cast<uintptr_t>(p + base) & 15u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
        if (fast) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + base);
            const uint32_t before = base ? (uint32_t)p[base - 1] : 0u;             // header byte (0x00 = raw) for the very first word
            uint4 w;
            w.x = __byte_perm(before, v.x, 0x6540);       // {before.b0, v.x.b0, v.x.b1, v.x.b2}
            w.y = __byte_perm(v.x, v.y, 0x6543);          // {v.x.b3, v.y.b0, v.y.b1, v.y.b2}
            w.z = __byte_perm(v.y, v.z, 0x6543);
            w.w = __byte_perm(v.z, v.w, 0x6543);
            *reinterpret_cast<uint4 *>(out + base) = w;   // covers out[base .. base+15] = {p[base-1], p[base .. base+14]}
            if (base + kHalf == n) out[n] = (uint8_t)(v.w >> 24);                  // last payload byte
        } else {
            // a word-building neighbour to the left leaves out[base] (= p[base-1]) to this thread
            if (base) out[base] = p[base - 1];
            for (int j = 0; j < kHalf && base + j < n; ++j) out[1 + base + j] = p[base + j];
        }
    }
}

}  // namespace mbc_seg
