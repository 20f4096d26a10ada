// chunk_codec.cuh -- DataChunk serialisation on the device, HBM-bound byte work.
//
// Replaces DataChunk.Serialize (DistributedMandelbrot/DataChunk.cs:173-206), which serialises the
// 16 MiB payload TWICE on one CPU thread (RawSerializer, DataChunkSerializer.cs:29-49, and
// RLESerializer, :51-100, each into a SizeCountStream) to pick the shorter and then a third time
// to write it.  Output bytes are identical to the reference's:
//     byte 0 = codec: 0x00 raw -> the n payload bytes ; 0x01 RLE -> (uint32 LE runLength, uint8
//     value) records, one per maximal run.  Raw wins ties (`<` with Raw first, DataChunk.cs:188).
//
// Four small kernels; every pass is a coalesced 128-bit stream over the payload:
//   1 count_run_starts   per-CTA number of run starts (p[i] != p[i-1], i = 0 always starts a run)
//   2 scan_block_counts  exclusive scan of the CTA counts, picks the codec, writes header + length
//   3 scatter_runs       RLE: every run start -> starts[r], value byte into its record
//                        raw: the payload copied behind the header byte (byte-shifted words)
//   4 write_run_lengths  RLE: runLength = starts[r+1] - starts[r] into the records
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mbc {

constexpr int kThreads = 256;
constexpr int kBytesPerThread = 16;
constexpr int kSegment = kThreads * kBytesPerThread;   // 4096 payload bytes per CTA

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

// A batch of chunks is processed by one set of launches: blockIdx.y = chunk.
struct Batch {
    const uint8_t *payloads;      // chunk c at payloads + c * payload_stride
    size_t payload_stride;
    uint8_t *out;                 // chunk c at out + c * out_stride (16-byte aligned stride)
    size_t out_stride;
    uint32_t *counts;             // per chunk: counts + c * counts_stride
    size_t counts_stride;
    uint32_t *starts;             // per chunk: starts + c * starts_stride
    size_t starts_stride;
    Meta *meta;                   // meta[c]
    size_t n;                     // payload bytes per chunk
    uint32_t nblocks;             // CTAs per chunk
};

// 4-bit mask of the bytes that differ between two words (SIMD byte compare, no per-byte loop)
__device__ __forceinline__ uint32_t differ4(uint32_t a, uint32_t b)
{
    const uint32_t ne = __vcmpne4(a, b) & 0x01010101u;          // bit 0 of each byte = "differs"
    return ((ne * 0x01020408u) >> 24) & 0xfu;                   // gather bits 0, 8, 16, 24 -> 0..3
}

// 16-bit mask: bit j set iff byte j of this thread's 16-byte span starts a run
// (p[i] != p[i-1]; i = 0 always starts one).  `v` receives the 16 payload bytes (zero-padded).
__device__ __forceinline__ uint32_t start_mask(const uint8_t *__restrict__ p, size_t base, size_t n, uint4 &v)
{
    v = make_uint4(0, 0, 0, 0);
    if (base >= n) return 0;
    const uint32_t prev = base ? (uint32_t)p[base - 1] : 0u;
    const bool full = base + kBytesPerThread <= n && ((reinterpret_cast<uintptr_t>(p + base) & 15u) == 0);
    uint32_t mask;
    if (full) {
        v = *reinterpret_cast<const uint4 *>(p + base);
        // the same 16 bytes shifted up by one byte, `prev` entering at the bottom
        const uint32_t s0 = __byte_perm(prev, v.x, 0x6540), s1 = __byte_perm(v.x, v.y, 0x6543);
        const uint32_t s2 = __byte_perm(v.y, v.z, 0x6543), s3 = __byte_perm(v.z, v.w, 0x6543);
        mask = differ4(v.x, s0) | (differ4(v.y, s1) << 4) | (differ4(v.z, s2) << 8) | (differ4(v.w, s3) << 12);
    } else {
        uint8_t bytes[kBytesPerThread];
        mask = 0;
        uint8_t pb = (uint8_t)prev;
#pragma unroll
        for (int j = 0; j < kBytesPerThread; ++j) {
            bytes[j] = (base + j < n) ? p[base + j] : 0;
            if (base + j < n && bytes[j] != pb) mask |= 1u << j;
            pb = bytes[j];
        }
        v.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | ((uint32_t)bytes[3] << 24);
        v.y = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | ((uint32_t)bytes[7] << 24);
        v.z = bytes[8] | (bytes[9] << 8) | (bytes[10] << 16) | ((uint32_t)bytes[11] << 24);
        v.w = bytes[12] | (bytes[13] << 8) | (bytes[14] << 16) | ((uint32_t)bytes[15] << 24);
    }
    if (base == 0) mask |= 1u;
    return mask;
}

__device__ __forceinline__ uint32_t byte_of(const uint4 &v, int j)
{
    const uint32_t w = j < 8 ? (j < 4 ? v.x : v.y) : (j < 12 ? v.z : v.w);
    return (w >> (8 * (j & 3))) & 0xffu;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[kThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if ((uint32_t)w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kThreads) count_run_starts(const Batch b)
{
    const uint8_t *__restrict__ p = b.payloads + blockIdx.y * b.payload_stride;
    uint32_t *__restrict__ block_counts = b.counts + blockIdx.y * b.counts_stride;
    const size_t n = b.n;
    uint4 v;
    const size_t base = (size_t)blockIdx.x * kSegment + (size_t)threadIdx.x * kBytesPerThread;
    const uint32_t c = __popc(start_mask(p, base, n, v));
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// one CTA per chunk; in-place exclusive scan of block_counts[0..nblocks)
__global__ void __launch_bounds__(kThreads) scan_block_counts(const Batch b)
{
    uint32_t *__restrict__ block_counts = b.counts + blockIdx.x * b.counts_stride;
    Meta *__restrict__ meta = b.meta + blockIdx.x;
    uint8_t *__restrict__ out = b.out + blockIdx.x * b.out_stride;
    const uint8_t *__restrict__ p = b.payloads + blockIdx.x * b.payload_stride;
    const uint32_t nblocks = b.nblocks;
    const size_t n = b.n;
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t first = 0; first < nblocks; first += kThreads) {
        const uint32_t i = first + threadIdx.x;
        const uint32_t v = i < nblocks ? block_counts[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        const uint32_t carry = carry_s;
        if (i < nblocks) block_counts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t runs = carry_s;
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

__global__ void __launch_bounds__(kThreads) scatter_runs(const Batch b)
{
    const uint8_t *__restrict__ p = b.payloads + blockIdx.y * b.payload_stride;
    const uint32_t *__restrict__ block_offsets = b.counts + blockIdx.y * b.counts_stride;
    const Meta *__restrict__ meta = b.meta + blockIdx.y;
    uint32_t *__restrict__ starts = b.starts + blockIdx.y * b.starts_stride;
    uint8_t *__restrict__ out = b.out + blockIdx.y * b.out_stride;
    const size_t n = b.n;
    uint4 v;
    const size_t base = (size_t)blockIdx.x * kSegment + (size_t)threadIdx.x * kBytesPerThread;
    if (meta->codec == 0) {
        // RawSerializer (DataChunkSerializer.cs:34-37): out[1 + i] = p[i].  The destination is
        // 1 byte off alignment, so build aligned 32-bit words from two neighbouring source words.
        if (base >= n) return;
        const bool fast = base + kBytesPerThread <= n && ((reinterpret_cast<uintptr_t>(p + base) & 15u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
        if (fast) {
            v = *reinterpret_cast<const uint4 *>(p + base);
            const uint32_t before = base ? (uint32_t)p[base - 1] : (uint32_t)out[0];   // header byte for the very first word
            uint4 w;
            w.x = __byte_perm(before, v.x, 0x6540);       // {before.b0, v.x.b0, v.x.b1, v.x.b2}
            w.y = __byte_perm(v.x, v.y, 0x6543);          // {v.x.b3, v.y.b0, v.y.b1, v.y.b2}
            w.z = __byte_perm(v.y, v.z, 0x6543);
            w.w = __byte_perm(v.z, v.w, 0x6543);
            *reinterpret_cast<uint4 *>(out + base) = w;   // covers out[base .. base+15] = {p[base-1], p[base .. base+14]}
            if (base + kBytesPerThread == n) out[n] = (uint8_t)(v.w >> 24);           // last payload byte
        } else {
            // a word-building neighbour to the left leaves out[base] (= p[base-1]) to this thread
            if (base) out[base] = p[base - 1];
            for (int j = 0; j < kBytesPerThread && base + j < n; ++j) out[1 + base + j] = p[base + j];
        }
        return;
    }
    uint32_t mask = start_mask(p, base, n, v);
    uint32_t total;
    uint32_t r = block_offsets[blockIdx.x] + block_exclusive_scan(__popc(mask), &total);
    while (mask) {                                        // one trip per run start, not per byte
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        starts[r] = (uint32_t)(base + j);
        out[1 + (size_t)r * 5 + 4] = (uint8_t)byte_of(v, j);   // runValue (DataChunkSerializer.cs:84,98)
        ++r;
    }
}

__global__ void __launch_bounds__(kThreads) write_run_lengths(const Batch b)
{
    const Meta *__restrict__ meta = b.meta + blockIdx.y;
    const uint32_t *__restrict__ starts = b.starts + blockIdx.y * b.starts_stride;
    uint8_t *__restrict__ out = b.out + blockIdx.y * b.out_stride;
    const size_t n = b.n;
    if (meta->codec == 0) return;
    const uint32_t runs = meta->runs;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < runs; r += (size_t)gridDim.x * blockDim.x) {
        const uint32_t len = (r + 1 < runs ? starts[r + 1] : (uint32_t)n) - starts[r];
        uint8_t *rec = out + 1 + r * 5;                   // BitConverter.GetBytes(runLength): little-endian
        rec[0] = (uint8_t)len;
        rec[1] = (uint8_t)(len >> 8);
        rec[2] = (uint8_t)(len >> 16);
        rec[3] = (uint8_t)(len >> 24);
    }
}

}  // namespace mbc
