// cuda_emul.h -- TEST INFRASTRUCTURE: runs the text of a CUDA kernel on the CPU.  Every CUDA thread of a CTA is a
// fiber (its own stack, cooperative switching inside ONE OS thread); CTAs run one after the other in dispatch order
// (blockIdx.x fastest).  Just enough of the CUDA surface for the kernels under distributedmandelbrot_b200/csrc/
// (escape_kernels.cuh, escape_refill.cuh, chunk_codec.cuh): thread/block indices, __syncthreads, the warp shuffles /
// votes / reductions they use, atomicAdd, the integer intrinsics, the *_rn floating-point intrinsics (plain IEEE
// operations: this code is compiled with -ffp-contract=off, fma() is the correctly rounded one), small vector types.
// It exists so that the kernels' logic -- the deferred escape test with its replay, the lane refill, the
// serializer's index arithmetic -- can be checked against the CPU oracle / the restated reference serializer in a
// container without a GPU (tests/test_escape_emulated.py, tests/test_codec_emulated.py).  Never part of the product
// build; nothing here is timed.
//
// Semantics kept: the threads of a CTA make progress independently and meet at barriers (a fiber that waits at a
// barrier yields to the next one); __shared__ variables are one per CTA (CTAs are sequential, so a function-level
// static does it); warp collectives exchange through a per-warp slot array between two warp barriers.  Not
// modelled: the memory model (a single OS thread is sequentially consistent), divergence inside a warp collective
// (the kernel never does that), more than one CTA in flight (a look-back never has to wait: see the keep-aggregate
// hook of the kernel for how its summation over several predecessors is still exercised).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

struct emul_uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct alignas(16) int4 { int32_t x, y, z, w; };
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { return int4{x, y, z, w}; }
struct alignas(8) uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) alignas(x)

namespace emul {

constexpr int kMaxThreads = 1024;

struct Barrier {
    unsigned count, arrived, generation;
};
struct Warp {
    Barrier bar;
    uint64_t slot[32];
};
struct Fiber {
    void *sp;                     // saved stack pointer (x86-64) / context (elsewhere) while the fiber is not running
    emul_uint3 tid, bid;
    bool done;
};
struct Cta {
    Barrier bar;
    int or_acc;
    Warp warps[kMaxThreads / 32];
    alignas(16) uint8_t dyn_smem[228 * 1024];
};
extern Cta g_cta;
extern Fiber *g_cur;              // the running fiber
extern emul_uint3 g_block_dim, g_grid_dim;

void yield();                     // run the other fibers of the CTA once round, then come back

inline void wait(Barrier &b)
{
    const unsigned g = b.generation;
    if (++b.arrived == b.count) {
        b.arrived = 0;
        ++b.generation;
    } else {
        while (b.generation == g) yield();
    }
}

inline Warp &my_warp() { return g_cta.warps[g_cur->tid.x >> 5]; }
inline unsigned my_lane() { return g_cur->tid.x & 31u; }

template <class T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "warp collectives move at most 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// every lane deposits its value, reads what `pick` tells it to, and leaves only when everybody has read
template <class T, class F>
inline T exchange(T v, F pick)
{
    Warp &w = my_warp();
    w.slot[my_lane()] = to_bits(v);
    wait(w.bar);
    const T r = pick(w.slot);
    wait(w.bar);
    return r;
}

// kernel launch: `fn` is the kernel body bound to its arguments
void launch(dim3 grid, unsigned threads, void (*fn)(const void *), const void *args);

}  // namespace emul

#define threadIdx (emul::g_cur->tid)
#define blockIdx (emul::g_cur->bid)
#define blockDim (emul::g_block_dim)
#define gridDim (emul::g_grid_dim)
#define EMUL_DYNAMIC_SMEM emul::g_cta.dyn_smem

inline void __syncthreads() { emul::wait(emul::g_cta.bar); }
inline int __syncthreads_or(int p)
{
    if (p) emul::g_cta.or_acc = 1;
    emul::wait(emul::g_cta.bar);
    const int r = emul::g_cta.or_acc;
    emul::wait(emul::g_cta.bar);
    if (threadIdx.x == 0) emul::g_cta.or_acc = 0;
    emul::wait(emul::g_cta.bar);
    return r;
}

template <class T>
inline T __shfl_sync(unsigned, T v, int src)
{
    return emul::exchange(v, [&](const uint64_t *s) { return emul::from_bits<T>(s[src & 31]); });
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const unsigned l = emul::my_lane();
    return emul::exchange(v, [&](const uint64_t *s) { return l >= d ? emul::from_bits<T>(s[l - d]) : v; });
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
    const unsigned l = emul::my_lane();
    return emul::exchange(v, [&](const uint64_t *s) { return l + d < 32 ? emul::from_bits<T>(s[l + d]) : v; });
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m)
{
    const unsigned l = emul::my_lane();
    return emul::exchange(v, [&](const uint64_t *s) { return emul::from_bits<T>(s[(l ^ (unsigned)m) & 31]); });
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    return emul::exchange<unsigned>(pred ? 1u : 0u, [&](const uint64_t *s) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i) r |= (s[i] ? 1u : 0u) << i;
        return r;
    });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    return emul::exchange<unsigned>(v, [&](const uint64_t *s) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i) r += (unsigned)s[i];
        return r;
    });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    return emul::exchange<unsigned>(v, [&](const uint64_t *s) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i) r = (unsigned)s[i] > r ? (unsigned)s[i] : r;
        return r;
    });
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
// PRMT, default mode: selector nibble i picks byte (nibble & 7) of {x = bytes 0..3, y = bytes 4..7}; bit 3 of a
// nibble replicates the sign of that byte
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const uint64_t pool = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(pool >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)      // wrap mode
{
    sh &= 31u;
    return (unsigned)(((((uint64_t)hi << 32) | lo) << sh) >> 32);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31u;
    return (unsigned)((((uint64_t)hi << 32) | lo) >> sh);
}
inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh)     // clamp mode
{
    sh = sh > 32u ? 32u : sh;
    return sh == 32u ? lo : __funnelshift_l(lo, hi, sh);
}
inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh)
{
    sh = sh > 32u ? 32u : sh;
    return sh == 32u ? hi : __funnelshift_r(lo, hi, sh);
}

inline int __all_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) == 0xffffffffu; }
inline int __any_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) != 0u; }
inline void __syncwarp(unsigned = 0xffffffffu) { emul::wait(emul::my_warp().bar); }

// one OS thread: a plain read-modify-write is atomic
template <class T>
inline T atomicAdd(T *p, T v)
{
    const T old = *p;
    *p = old + v;
    return old;
}

// the *_rn intrinsics: one IEEE operation each, round to nearest even (this TU is built with -ffp-contract=off, so
// the compiler never fuses a * b + c on its own; fma / fmaf are the correctly rounded library functions)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __double2float_rn(double a) { return (float)a; }
inline int __double2hiint(double a) { return (int)(emul::to_bits(a) >> 32); }
inline long long __double_as_longlong(double a) { return (long long)emul::to_bits(a); }
inline int __float_as_int(float a) { return (int)(uint32_t)emul::to_bits(a); }

// hook of chunk_codec.cuh: which tiles stay at their published aggregate (see the kernel)
bool mbc_emul_keep_aggregate(uint32_t tile);
