// mandelbrot_capi.cu -- implementation of include/mandelbrot_b200.h (the C-ABI boundary).
//
// Host side of the hot path: argument checking in the spirit of the DataChunk ctor
// (DistributedMandelbrot/DataChunk.cs:99-106), chunk geometry (.py:75-78 / DataChunk.cs:32-33,
// 59-66), linspace parameters (.py:24-32), kernel launches, and the synchronous
// process_workload equivalents (.py:70-100).  No CPU compute fallback exists in this file:
// every compute entry point either launches a CUDA kernel or fails.
#include <cuda_runtime.h>

#include <time.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "chunk_codec.cuh"
#include "escape_kernels.cuh"
#include "escape_refill.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_device = -1;
thread_local uint32_t g_tile_flags = 0;     // mb_set_period_shortcut: flags of the one-call entry points' tiles
thread_local unsigned long long *g_executed = nullptr;   // mb_set_executed_counter (shortcut launches only)
thread_local double g_first_claim = 0.0, g_last_claim = 0.0;   // mb_frame_last_stats
thread_local const mb_tile *g_tiles_dev_hint = nullptr;        // launch_tiles: descriptors already on the device (mb_process_workloads_u8)

double now_monotonic()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#ifdef MB_TUNING_VARIANTS
int g_variant = 0;
#endif

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define MB_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(MB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

constexpr int kMaxDevices = 64;

// per-device context for the synchronous process_workload path
struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    cudaStream_t stream = nullptr;      // private non-blocking stream for the host-buffer entry points
    cudaStream_t lanes[4] = {nullptr, nullptr, nullptr, nullptr};   // concurrent launch lanes of the dispatch loop
    void *scratch = nullptr;            // device output buffer of the host-buffer entry points
    size_t scratch_bytes = 0;
    unsigned long long *counters = nullptr;   // patch counters of the lane-refill kernel, one per launch in flight
    uint32_t next_counter = 0;
    mb_tile *desc = nullptr;            // descriptors of one mb_process_workloads_u8 call, uploaded once
    size_t desc_tiles = 0;
    void *stage[16] = {};               // per-slot device staging of mb_frame_render_u8_host
    size_t stage_bytes[16] = {};
    std::mutex mu;
};

// events that are destroyed on every exit path
struct EventSet {
    cudaEvent_t ev[20];
    uint32_t n = 0;
    int add(uint32_t count)
    {
        for (uint32_t k = 0; k < count && n < 20; ++k) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[n], cudaEventDisableTiming);
            if (e != cudaSuccess) return (int)e;
            ++n;
        }
        return 0;
    }
    ~EventSet() { for (uint32_t k = 0; k < n; ++k) cudaEventDestroy(ev[k]); }
};
DeviceCtx g_ctx[kMaxDevices];

int current_ctx(DeviceCtx **out)
{
    if (g_device < 0 || g_device >= kMaxDevices || !g_ctx[g_device].ready)
        return fail(MB_ERR_STATE, "mb_init(device) has not been called on this thread");
    MB_CUDA(cudaSetDevice(g_device));
    *out = &g_ctx[g_device];
    return MB_OK;
}

int ensure_scratch(DeviceCtx *c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return MB_OK;
    if (c->scratch) {
        MB_CUDA(cudaFree(c->scratch));
        c->scratch = nullptr;
        c->scratch_bytes = 0;
    }
    MB_CUDA(cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return MB_OK;
}

int check_tiles(const mb_tile *tiles, uint32_t n_tiles, uint32_t *max_w, uint32_t *max_h)
{
    if (!tiles || n_tiles == 0) return fail(MB_ERR_INVALID, "no tiles");
    uint32_t mw = 0, mh = 0;
    for (uint32_t k = 0; k < n_tiles; ++k) {
        const mb_tile &t = tiles[k];
        if (t.definition == 0) return fail(MB_ERR_INVALID, "tile %u: definition == 0", k);
        if (t.w == 0 || t.h == 0) return fail(MB_ERR_INVALID, "tile %u: empty rectangle", k);
        if ((uint64_t)t.x0 + t.w > t.definition || (uint64_t)t.y0 + t.h > t.definition)
            return fail(MB_ERR_INVALID, "tile %u: rectangle exceeds the chunk grid", k);
        if (t.mrd > 0x7fffffffu) return fail(MB_ERR_INVALID, "tile %u: mrd exceeds int32 (.py:39)", k);
        if (t.out_pitch < t.w) return fail(MB_ERR_INVALID, "tile %u: out_pitch < w", k);
        if (t.flags & ~(uint32_t)(MB_TILE_PERIOD_SHORTCUT | MB_TILE_LANE_REFILL)) return fail(MB_ERR_INVALID, "tile %u: unknown flag bits 0x%x", k, t.flags);
        if (t.flags != tiles[0].flags) return fail(MB_ERR_INVALID, "tile %u: every tile of one launch must carry the same flags", k);
        mw = t.w > mw ? t.w : mw;
        mh = t.h > mh ? t.h : mh;
    }
    *max_w = mw;
    *max_h = mh;
    return MB_OK;
}

// Launch shape.  CTA = 2 x 2 warps of 8x4 pixels (16 x 8 pixels, 128 threads): small CTAs
// release their SM slots soon after their own pixels finish, which matters because escape
// times differ by orders of magnitude between neighbouring CTAs at the set boundary.
// 34 registers/thread, 12 CTAs/SM = 48 resident warps/SM; the FP64 pipe saturates with ~3 warps
// per sub-partition, so occupancy beyond that buys nothing (profiles/r01_tuning_sweep.txt).
constexpr int kWX = 2, kWY = 2, kMinCtas = 12, kHead = 32;
constexpr uint32_t kCtaW = kWX * mbk::kWarpW, kCtaH = kWY * mbk::kWarpH;
constexpr uint32_t kMaxGridY = 65535;
constexpr int kRefillCtas = 8;             // resident CTAs per SM of the lane-refill kernel (<= 64 registers; 8 warps per sub-partition saturate the FP64 pipe)
constexpr uint32_t kCounterRing = 256;

template <typename OutT>
int launch_tiles(const mb_tile *tiles, uint32_t n_tiles, OutT *out_dev, int dtype, cudaStream_t st)
{
    uint32_t mw = 0, mh = 0;
    int rc = check_tiles(tiles, n_tiles, &mw, &mh);
    if (rc) return rc;
    if (!out_dev) return fail(MB_ERR_INVALID, "out_dev is NULL");
    if (dtype != MB_DTYPE_F64 && dtype != MB_DTYPE_F32) return fail(MB_ERR_INVALID, "unknown dtype %d", dtype);
    DeviceCtx *c = nullptr;
    if ((rc = current_ctx(&c))) return rc;

    // 88 bytes per tile is the whole input of the hot path (the reference ships 2 x 128 MiB of
    // float64 per chunk, .py:82-88).  Small batches ride in the kernel parameter block; larger
    // ones in a stream-ordered device allocation (the pool keeps its memory, see mb_init).
    mbk::TileArgs args;
    std::memset(&args, 0, sizeof(args));
    args.executed = g_executed;          // only read by the shortcut instantiations
    mb_tile *tiles_dev = nullptr;
    const mb_tile *hint = g_tiles_dev_hint;          // the caller uploaded these descriptors itself (one copy for many launches)
    g_tiles_dev_hint = nullptr;
    const bool owned = (hint == nullptr);
    if (n_tiles <= (uint32_t)mbk::kInlineTiles) {
        args.n_inline = n_tiles;
        std::memcpy(args.inl, tiles, sizeof(mb_tile) * n_tiles);
    } else if (hint) {
        tiles_dev = const_cast<mb_tile *>(hint);
    } else {
        const size_t bytes = sizeof(mb_tile) * (size_t)n_tiles;
        MB_CUDA(cudaMallocAsync((void **)&tiles_dev, bytes, st));
        const cudaError_t ce = cudaMemcpyAsync(tiles_dev, tiles, bytes, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) {
            cudaFreeAsync(tiles_dev, st);
            return fail(MB_ERR_CUDA, "descriptor upload failed: %s", cudaGetErrorString(ce));
        }
    }

#ifdef MB_TUNING_VARIANTS
    // development-only sweep of launch shapes (scripts/tune_variants.py); never built by default
    if (g_variant != 0 && dtype == MB_DTYPE_F64) {
        const int v = g_variant;
        auto go = [&](auto head, auto wx, auto wy, auto minc) -> int {
            constexpr int H = decltype(head)::value, WX = decltype(wx)::value, WY = decltype(wy)::value, MC = decltype(minc)::value;
            const uint32_t cw = WX * mbk::kWarpW, ch = WY * mbk::kWarpH;
            const uint32_t per_tile = ((mw + cw - 1) / cw) * ((mh + ch - 1) / ch);
            for (uint32_t first = 0; first < n_tiles; first += kMaxGridY) {
                const uint32_t cnt = (n_tiles - first < kMaxGridY) ? n_tiles - first : kMaxGridY;
                args.dev = tiles_dev ? tiles_dev + first : nullptr;
                mbk::escape_tiles_kernel<double, OutT, H, WX, WY, MC><<<dim3(per_tile, cnt, 1), WX * WY * 32, 0, st>>>(args, out_dev);
                MB_CUDA(cudaGetLastError());
            }
            return MB_OK;
        };
        using std::integral_constant;
#define MB_V(id, H, WX, WY, MC) if (v == id) { rc = go(integral_constant<int, H>{}, integral_constant<int, WX>{}, integral_constant<int, WY>{}, integral_constant<int, MC>{}); }
        MB_V(1, 16, 2, 2, 12) MB_V(2, 64, 2, 2, 12) MB_V(3, 32, 1, 2, 24) MB_V(4, 32, 2, 4, 6) MB_V(5, 32, 4, 1, 12)
        MB_V(6, 32, 1, 1, 32) MB_V(7, 32, 2, 2, 8) MB_V(8, 0, 2, 2, 12) MB_V(9, 32, 4, 2, 6)
#undef MB_V
        if (tiles_dev && owned) MB_CUDA(cudaFreeAsync(tiles_dev, st));
        return rc;
    }
#endif
    if (tiles[0].flags & MB_TILE_LANE_REFILL) {
        // persistent warps: patches handed out by one atomic counter per launch (escape_refill.cuh)
        cudaError_t e = cudaSuccess;
        if (!c->counters) {
            e = cudaMalloc((void **)&c->counters, sizeof(unsigned long long) * kCounterRing);
            if (e != cudaSuccess) c->counters = nullptr;
        }
        if (e == cudaSuccess) {
            args.counter = c->counters + (c->next_counter++ % kCounterRing);
            e = cudaMemsetAsync(args.counter, 0, sizeof(unsigned long long), st);
        }
        if (e != cudaSuccess) {
            rc = fail(MB_ERR_CUDA, "patch counter: %s", cudaGetErrorString(e));
        } else {
            args.dev = tiles_dev;
            args.n_tiles = n_tiles;
            args.pcols = (mw + mbk::kWarpW - 1) / mbk::kWarpW;
            args.prows = (mh + mbk::kWarpH - 1) / mbk::kWarpH;
            const unsigned long long slots = (unsigned long long)args.pcols * args.prows * n_tiles;
            unsigned long long ctas = (slots + 3) / 4;
            const unsigned long long cap = (unsigned long long)c->sm_count * kRefillCtas;
            if (ctas > cap) ctas = cap;
            const bool period = (tiles[0].flags & MB_TILE_PERIOD_SHORTCUT) != 0;
#ifdef MB_TUNING_VARIANTS
            // development-only: MB_REFILL_CTAS=10 selects the 51-register build of the refill kernel
            static const int tune_ctas = getenv("MB_REFILL_CTAS") ? atoi(getenv("MB_REFILL_CTAS")) : 0;
            if (tune_ctas == 10 && dtype == MB_DTYPE_F64) {
                unsigned long long ctas10 = (slots + 3) / 4;
                if (ctas10 > (unsigned long long)c->sm_count * 10) ctas10 = (unsigned long long)c->sm_count * 10;
                if (period) mbk::escape_refill_kernel<double, OutT, 10, true><<<dim3((unsigned)ctas10), 128, 0, st>>>(args, out_dev);
                else mbk::escape_refill_kernel<double, OutT, 10, false><<<dim3((unsigned)ctas10), 128, 0, st>>>(args, out_dev);
                e = cudaGetLastError();
                if (e != cudaSuccess) rc = fail(MB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
                if (tiles_dev && owned) cudaFreeAsync(tiles_dev, st);
                return rc;
            }
#endif
            const dim3 grid((unsigned)ctas, 1, 1), block(128, 1, 1);
            if (dtype == MB_DTYPE_F64) {
                if (period) mbk::escape_refill_kernel<double, OutT, kRefillCtas, true><<<grid, block, 0, st>>>(args, out_dev);
                else mbk::escape_refill_kernel<double, OutT, kRefillCtas, false><<<grid, block, 0, st>>>(args, out_dev);
            } else {
                if (period) mbk::escape_refill_kernel<float, OutT, kRefillCtas, true><<<grid, block, 0, st>>>(args, out_dev);
                else mbk::escape_refill_kernel<float, OutT, kRefillCtas, false><<<grid, block, 0, st>>>(args, out_dev);
            }
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = fail(MB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
        }
        if (tiles_dev && owned) cudaFreeAsync(tiles_dev, st);
        return rc;
    }
    const uint64_t ctas64 = (uint64_t)((mw + kCtaW - 1) / kCtaW) * ((mh + kCtaH - 1) / kCtaH);
    cudaError_t err = cudaSuccess;
    if (ctas64 > 0x7fffffffull) {
        rc = fail(MB_ERR_INVALID, "tile too large for one launch (%llu CTAs)", (unsigned long long)ctas64);
    } else {
        const uint32_t ctas_per_tile = (uint32_t)ctas64;
        for (uint32_t first = 0; first < n_tiles && err == cudaSuccess; first += kMaxGridY) {
            const uint32_t cnt = (n_tiles - first < kMaxGridY) ? n_tiles - first : kMaxGridY;
            const dim3 grid(ctas_per_tile, cnt, 1), block(kWX * kWY * 32, 1, 1);
            args.dev = tiles_dev ? tiles_dev + first : nullptr;
            const bool period = (tiles[0].flags & MB_TILE_PERIOD_SHORTCUT) != 0;   // exact shortcut, opt-in
            if (dtype == MB_DTYPE_F64) {
                if (period) mbk::escape_tiles_kernel<double, OutT, kHead, kWX, kWY, kMinCtas, true><<<grid, block, 0, st>>>(args, out_dev);
                else mbk::escape_tiles_kernel<double, OutT, kHead, kWX, kWY, kMinCtas><<<grid, block, 0, st>>>(args, out_dev);
            } else {
                if (period) mbk::escape_tiles_kernel<float, OutT, kHead, kWX, kWY, kMinCtas, true><<<grid, block, 0, st>>>(args, out_dev);
                else mbk::escape_tiles_kernel<float, OutT, kHead, kWX, kWY, kMinCtas><<<grid, block, 0, st>>>(args, out_dev);
            }
            err = cudaGetLastError();
        }
        if (err != cudaSuccess) rc = fail(MB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(err));
    }
    if (tiles_dev && owned) cudaFreeAsync(tiles_dev, st);      // also on the error paths: never leak the descriptor batch
    return rc;
}

// strict binary64 host arithmetic: `volatile` pins each rounding even if someone adds -mfma
inline double f_mul(double a, double b) { volatile double r = a * b; return r; }
inline double f_add(double a, double b) { volatile double r = a + b; return r; }
inline double f_sub(double a, double b) { volatile double r = a - b; return r; }
inline double f_div(double a, double b) { volatile double r = a / b; return r; }

int grid_for(size_t n, int sm_count)
{
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" {

#ifdef MB_TUNING_VARIANTS
void mb_internal_set_variant(int v) { g_variant = v; }
#endif

// shared with tile_wsq.cpp so that queue errors are readable through mb_last_error()
int mb_internal_fail(int code, const char *msg) { return fail(code, "%s", msg); }

int mb_abi_version(void) { return MB_ABI_VERSION; }

const char *mb_last_error(void) { return g_err; }

int mb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(MB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return n;
}

int mb_init(int device)
{
    if (device < 0 || device >= kMaxDevices) return fail(MB_ERR_INVALID, "device %d out of range", device);
    int n = mb_device_count();
    if (n < 0) return n;
    if (device >= n) return fail(MB_ERR_CUDA, "device %d requested but only %d CUDA device(s) present", device, n);
    MB_CUDA(cudaSetDevice(device));
    DeviceCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) {
        cudaDeviceProp prop;
        MB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(MB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        c.sm_count = prop.multiProcessorCount;
        MB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        for (auto &ln : c.lanes) MB_CUDA(cudaStreamCreateWithFlags(&ln, cudaStreamNonBlocking));
        // descriptor batches use cudaMallocAsync: keep the pool's memory across synchronisations
        // instead of returning it to the OS (default threshold 0 re-allocates on every call)
        cudaMemPool_t pool;
        MB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        MB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        c.ready = true;
    }
    g_device = device;
    return MB_OK;
}

void mb_shutdown(void)
{
    for (int d = 0; d < kMaxDevices; ++d) {
        DeviceCtx &c = g_ctx[d];
        std::lock_guard<std::mutex> lk(c.mu);
        if (!c.ready) continue;
        cudaSetDevice(d);
        if (c.scratch) cudaFree(c.scratch);
        if (c.counters) cudaFree(c.counters);
        c.counters = nullptr;
        if (c.desc) cudaFree(c.desc);
        c.desc = nullptr;
        c.desc_tiles = 0;
        for (int k = 0; k < 16; ++k) { if (c.stage[k]) cudaFree(c.stage[k]); c.stage[k] = nullptr; c.stage_bytes[k] = 0; }
        if (c.stream) cudaStreamDestroy(c.stream);
        for (auto &ln : c.lanes) { if (ln) cudaStreamDestroy(ln); ln = nullptr; }
        c.scratch = nullptr;
        c.scratch_bytes = 0;
        c.stream = nullptr;
        c.ready = false;
    }
    g_device = -1;
}

int mb_set_period_shortcut(int enabled)
{
    const int was = (g_tile_flags & MB_TILE_PERIOD_SHORTCUT) ? 1 : 0;
    g_tile_flags = enabled ? (g_tile_flags | MB_TILE_PERIOD_SHORTCUT) : (g_tile_flags & ~(uint32_t)MB_TILE_PERIOD_SHORTCUT);
    return was;
}

int mb_set_tile_flags(uint32_t flags)
{
    const int was = (int)g_tile_flags;
    if (flags & ~(uint32_t)(MB_TILE_PERIOD_SHORTCUT | MB_TILE_LANE_REFILL)) return fail(MB_ERR_INVALID, "unknown flag bits 0x%x", flags);
    g_tile_flags = flags;
    return was;
}

int mb_set_executed_counter(uint64_t *counter_dev)
{
    g_executed = reinterpret_cast<unsigned long long *>(counter_dev);
    return MB_OK;
}

int mb_sm_count(void)
{
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    return rc ? rc : c->sm_count;
}

int mb_chunk_geometry(uint32_t level, uint32_t index_real, uint32_t index_imag, uint32_t definition,
                      uint32_t mrd, mb_tile *out)
{
    if (!out) return fail(MB_ERR_INVALID, "out is NULL");
    // DataChunk.cs:99-106
    if (level == 0) return fail(MB_ERR_INVALID, "Level must be positive");
    if (index_real >= level) return fail(MB_ERR_INVALID, "Real index must be lesser than level");
    if (index_imag >= level) return fail(MB_ERR_INVALID, "Imag index must be lesser than level");
    if (definition == 0) return fail(MB_ERR_INVALID, "definition must be >= 1");
    if (mrd > 0x7fffffffu) return fail(MB_ERR_INVALID, "mrd exceeds int32 (.py:39)");
    // .py:75-78
    const double range = f_div(f_sub(MB_MAX_AXIS, MB_MIN_AXIS), (double)level);
    const double start_r = f_add(MB_MIN_AXIS, f_mul(range, (double)index_real));
    const double start_i = f_add(MB_MIN_AXIS, f_mul(range, (double)index_imag));
    std::memset(out, 0, sizeof(*out));
    out->start_r = start_r;
    out->start_i = start_i;
    if (definition > 1) {
        // .py:24-32 -> numpy.linspace: stop = start + _range ; step = (stop - start) / (num - 1)
        out->stop_r = f_add(start_r, range);
        out->stop_i = f_add(start_i, range);
        out->step_r = f_div(f_sub(out->stop_r, start_r), (double)(definition - 1));
        out->step_i = f_div(f_sub(out->stop_i, start_i), (double)(definition - 1));
    } else {
        // numpy.linspace(num=1) is [start]: the forced-endpoint slot (k == definition-1 == 0) holds start
        out->stop_r = start_r;
        out->stop_i = start_i;
        out->step_r = 0.0;
        out->step_i = 0.0;
    }
    out->definition = definition;
    out->mrd = mrd;
    out->x0 = out->y0 = 0;
    out->w = out->h = definition;
    out->out_offset = 0;
    out->out_pitch = definition;
    out->flags = 0;
    return MB_OK;
}

double mb_tile_coord(const mb_tile *t, int axis, uint32_t k)
{
    const double start = axis ? t->start_i : t->start_r;
    const double step = axis ? t->step_i : t->step_r;
    const double stop = axis ? t->stop_i : t->stop_r;
    if (k == t->definition - 1u) return stop;
    return f_add(f_mul((double)k, step), start);
}

int mb_render_counts(const mb_tile *tiles, uint32_t n_tiles, int32_t *out_dev, int dtype, void *stream)
{
    return launch_tiles<int32_t>(tiles, n_tiles, out_dev, dtype, (cudaStream_t)stream);
}

int mb_calc_mb_value_dev(const double *r_dev, const double *i_dev, size_t n, uint32_t mrd, int dtype,
                         int32_t *out_dev, void *stream)
{
    if (!r_dev || !i_dev || !out_dev) return fail(MB_ERR_INVALID, "NULL buffer");
    if (mrd > 0x7fffffffu) return fail(MB_ERR_INVALID, "mrd exceeds int32 (.py:39)");
    if (dtype != MB_DTYPE_F64 && dtype != MB_DTYPE_F32) return fail(MB_ERR_INVALID, "unknown dtype %d", dtype);
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (n == 0) return MB_OK;
    const size_t blocks = (n + 127) / 128;
    if (blocks > 0x7fffffffull) return fail(MB_ERR_INVALID, "too many elements for one launch");
    if (dtype == MB_DTYPE_F64)
        mbk::escape_points_kernel<double, kHead><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(r_dev, i_dev, n, (int)mrd, out_dev);
    else
        mbk::escape_points_kernel<float, kHead><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(r_dev, i_dev, n, (int)mrd, out_dev);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int mb_calc_mb_value(const double *r_host, const double *i_host, size_t n, uint32_t mrd, int dtype, int32_t *out_host)
{
    if (!r_host || !i_host || !out_host) return fail(MB_ERR_INVALID, "NULL buffer");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (n == 0) return MB_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    // layout of the scratch buffer: r[n] | i[n] | out[n]
    if ((rc = ensure_scratch(c, n * (2 * sizeof(double) + sizeof(int32_t))))) return rc;
    double *r_dev = static_cast<double *>(c->scratch);
    double *i_dev = r_dev + n;
    int32_t *o_dev = reinterpret_cast<int32_t *>(i_dev + n);
    MB_CUDA(cudaMemcpyAsync(r_dev, r_host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    MB_CUDA(cudaMemcpyAsync(i_dev, i_host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = mb_calc_mb_value_dev(r_dev, i_dev, n, mrd, dtype, o_dev, c->stream))) return rc;
    MB_CUDA(cudaMemcpyAsync(out_host, o_dev, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

int mb_render_u8(const mb_tile *tiles, uint32_t n_tiles, uint8_t *out_dev, int dtype, void *stream)
{
    for (uint32_t k = 0; tiles && k < n_tiles; ++k)
        if (tiles[k].mrd == 0) return fail(MB_ERR_INVALID, "tile %u: mrd == 0 divides by zero in the uint8 scaling (.py:96)", k);
    return launch_tiles<uint8_t>(tiles, n_tiles, out_dev, dtype, (cudaStream_t)stream);
}

int mb_counts_to_u8(const int32_t *counts_dev, uint8_t *out_dev, size_t n, uint32_t mrd, void *stream)
{
    if (!counts_dev || !out_dev) return fail(MB_ERR_INVALID, "NULL buffer");
    if (mrd == 0) return fail(MB_ERR_INVALID, "mrd == 0 divides by zero in the uint8 scaling (.py:96)");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (n == 0) return MB_OK;
    mbk::counts_to_u8_kernel<<<grid_for((n + 3) / 4, c->sm_count), 256, 0, (cudaStream_t)stream>>>(counts_dev, out_dev, n, mrd);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int mb_useful_iterations(const int32_t *counts_dev, size_t n, uint32_t mrd, uint64_t *sum_dev, void *stream)
{
    if (!counts_dev || !sum_dev) return fail(MB_ERR_INVALID, "NULL buffer");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (n == 0) return MB_OK;
    mbk::useful_iterations_kernel<<<grid_for(n, c->sm_count), 256, 0, (cudaStream_t)stream>>>(
        counts_dev, n, mrd, reinterpret_cast<unsigned long long *>(sum_dev));
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int mb_classify_u8(const uint8_t *payload_dev, size_t n, uint32_t *flags_dev, void *stream)
{
    if (!payload_dev || !flags_dev) return fail(MB_ERR_INVALID, "NULL buffer");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    const uint32_t ones[2] = {1u, 1u};
    MB_CUDA(cudaMemcpyAsync(flags_dev, ones, sizeof(ones), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (n == 0) return MB_OK;
    mbk::classify_u8_kernel<<<grid_for(n, c->sm_count), 256, 0, (cudaStream_t)stream>>>(payload_dev, n, flags_dev);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int mb_host_alloc(size_t bytes, void **host_ptr)
{
    if (!host_ptr || bytes == 0) return fail(MB_ERR_INVALID, "bad argument");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    MB_CUDA(cudaHostAlloc(host_ptr, bytes, cudaHostAllocPortable));
    return MB_OK;
}

int mb_host_free(void *host_ptr)
{
    if (!host_ptr) return MB_OK;
    MB_CUDA(cudaFreeHost(host_ptr));
    return MB_OK;
}

int mb_process_workload_u8(uint32_t level, uint32_t mrd, uint32_t index_real, uint32_t index_imag,
                           uint32_t definition, uint8_t *host_out)
{
    if (!host_out) return fail(MB_ERR_INVALID, "host_out is NULL");
    mb_tile t;
    int rc = mb_chunk_geometry(level, index_real, index_imag, definition, mrd, &t);
    if (rc) return rc;
    t.flags = g_tile_flags;
    DeviceCtx *c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t n = (size_t)definition * definition;
    if ((rc = ensure_scratch(c, n))) return rc;
    if ((rc = mb_render_u8(&t, 1, static_cast<uint8_t *>(c->scratch), MB_DTYPE_F64, c->stream))) return rc;
    MB_CUDA(cudaMemcpyAsync(host_out, c->scratch, n, cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

int mb_process_workload_counts(uint32_t level, uint32_t mrd, uint32_t index_real, uint32_t index_imag,
                               uint32_t definition, int dtype, int32_t *host_out)
{
    if (!host_out) return fail(MB_ERR_INVALID, "host_out is NULL");
    mb_tile t;
    int rc = mb_chunk_geometry(level, index_real, index_imag, definition, mrd, &t);
    if (rc) return rc;
    t.flags = g_tile_flags;
    DeviceCtx *c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t bytes = (size_t)definition * definition * sizeof(int32_t);
    if ((rc = ensure_scratch(c, bytes))) return rc;
    if ((rc = mb_render_counts(&t, 1, static_cast<int32_t *>(c->scratch), dtype, c->stream))) return rc;
    MB_CUDA(cudaMemcpyAsync(host_out, c->scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    return MB_OK;
}

// process_workload for a LIST of workloads (the loop of the reference worker, .py:178-184 +
// :70-100, with the per-chunk host work removed): geometry in C, several small chunks per launch
// (about `launch_bytes` of payload each), launches round-robin on the internal lanes, every launch
// followed by its device->host copy straight to the chunk's place in host_out.  Returns when all
// payloads have landed.  host_out should be page-locked (mb_host_alloc) for the copies to overlap.
int mb_process_workloads_u8(const uint32_t *workloads, uint32_t n_workloads, uint32_t definition, uint8_t *host_out,
                            size_t launch_bytes)
{
    if (!workloads || !host_out) return fail(MB_ERR_INVALID, "NULL buffer");
    if (n_workloads == 0) return MB_OK;
    if (definition == 0) return fail(MB_ERR_INVALID, "definition must be >= 1");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t n = (size_t)definition * definition;
    if (launch_bytes == 0) launch_bytes = (size_t)4 << 20;
    uint32_t group = (uint32_t)(launch_bytes / n);
    if (group == 0) group = 1;
    if (group > 1024) group = 1024;
    constexpr uint32_t kSlots = 8, kLanes = 4;     // 8 staging buffers: a frame of 8 launches is enqueued without a host-side wait
    for (uint32_t k = 0; k < kSlots; ++k) {
        if (c->stage_bytes[k] >= n * group) continue;
        if (c->stage[k]) { MB_CUDA(cudaFree(c->stage[k])); c->stage[k] = nullptr; c->stage_bytes[k] = 0; }
        MB_CUDA(cudaMalloc(&c->stage[k], n * group));
        c->stage_bytes[k] = n * group;
    }
    EventSet evs;
    if (evs.add(kSlots + 1)) return fail(MB_ERR_CUDA, "cudaEventCreate failed");
    // geometry of every workload on the host (C, strict binary64), ONE descriptor upload for the whole call
    std::vector<mb_tile> tiles(n_workloads);
    for (uint32_t k = 0; k < n_workloads; ++k) {
        const uint32_t *w = workloads + 4 * (size_t)k;                        // level, mrd, indexReal, indexImag (DistributerWorkload.cs:9-29)
        if (w[1] == 0) return fail(MB_ERR_INVALID, "workload %u: mrd == 0 divides by zero in the uint8 scaling (.py:96)", k);
        if ((rc = mb_chunk_geometry(w[0], w[2], w[3], definition, w[1], &tiles[k]))) return rc;
        tiles[k].out_offset = (uint64_t)(k % group) * n;                      // inside its launch's staging buffer
        tiles[k].flags = g_tile_flags;
    }
    if (c->desc_tiles < n_workloads) {
        if (c->desc) { MB_CUDA(cudaFree(c->desc)); c->desc = nullptr; c->desc_tiles = 0; }
        MB_CUDA(cudaMalloc((void **)&c->desc, sizeof(mb_tile) * (size_t)n_workloads));
        c->desc_tiles = n_workloads;
    }
    MB_CUDA(cudaMemcpyAsync(c->desc, tiles.data(), sizeof(mb_tile) * (size_t)n_workloads, cudaMemcpyHostToDevice, c->stream));
    MB_CUDA(cudaEventRecord(evs.ev[kSlots], c->stream));
    for (uint32_t k = 0; k < kLanes; ++k) MB_CUDA(cudaStreamWaitEvent(c->lanes[k], evs.ev[kSlots], 0));
    uint32_t launches = 0;
    int result = MB_OK;
    for (uint32_t first = 0; first < n_workloads && result == MB_OK; first += group, ++launches) {
        const uint32_t cnt = n_workloads - first < group ? n_workloads - first : group;
        const uint32_t slot = launches % kSlots;
        if (launches >= kSlots) {                                             // the slot's previous copy must have left the buffer
            cudaError_t e = cudaEventSynchronize(evs.ev[slot]);
            if (e != cudaSuccess) { result = fail(MB_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e)); break; }
        }
        cudaStream_t lane = c->lanes[launches % kLanes];
        g_tiles_dev_hint = c->desc + first;
        if ((result = mb_render_u8(tiles.data() + first, cnt, static_cast<uint8_t *>(c->stage[slot]), MB_DTYPE_F64, lane))) break;
        cudaError_t e = cudaMemcpyAsync(host_out + (size_t)first * n, c->stage[slot], (size_t)cnt * n, cudaMemcpyDeviceToHost, lane);
        if (e == cudaSuccess) e = cudaEventRecord(evs.ev[slot], lane);
        if (e != cudaSuccess) result = fail(MB_ERR_CUDA, "payload copy: %s", cudaGetErrorString(e));
    }
    g_tiles_dev_hint = nullptr;
    for (uint32_t k = 0; k < kLanes; ++k) {
        cudaError_t e = cudaStreamSynchronize(c->lanes[k]);
        if (e != cudaSuccess && !result) result = fail(MB_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    }
    return result;
}

// ---- peer-shared frame buffer (8e, alternative to the gather) ------------------------------------

int mb_peer_alloc(size_t bytes, void **dev_ptr, void *handle_out)
{
    if (!dev_ptr || !handle_out || bytes == 0) return fail(MB_ERR_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MB_PEER_HANDLE_BYTES, "handle size");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    void *p = nullptr;
    MB_CUDA(cudaMalloc(&p, bytes));                      // plain cudaMalloc: IPC needs a whole allocation
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(MB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return MB_OK;
}

int mb_peer_open(const void *handle, void **dev_ptr)
{
    if (!handle || !dev_ptr) return fail(MB_ERR_INVALID, "NULL argument");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    MB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MB_OK;
}

int mb_peer_close(void *dev_ptr)
{
    if (!dev_ptr) return MB_OK;
    MB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return MB_OK;
}

int mb_peer_free(void *dev_ptr)
{
    if (!dev_ptr) return MB_OK;
    MB_CUDA(cudaFree(dev_ptr));
    return MB_OK;
}

// ---- in-box dispatch loop (8e) ----------------------------------------------------------------

int mb_frame_render_stolen(mb_wsq *q, uint32_t rank, uint32_t n_ranks, const mb_tile *tiles, uint32_t n_tiles,
                           uint32_t bands_per_claim, uint32_t depth, int32_t *out_dev, int dtype, void *stream,
                           uint32_t *claimed, uint32_t max_claims, uint32_t *n_claims, uint32_t *n_stolen)
{
    if (!q || !tiles || !out_dev || !claimed || !n_claims) return fail(MB_ERR_INVALID, "NULL argument");
    if (n_ranks == 0 || rank >= n_ranks) return fail(MB_ERR_INVALID, "rank %u out of range", rank);
    if (bands_per_claim == 0) bands_per_claim = 1;
    if (depth == 0) depth = 1;
    if (depth > 16) depth = 16;
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);          // the lanes belong to the device context: one dispatch loop at a time
    cudaStream_t st = (cudaStream_t)stream;
    // Claims go round-robin onto kLanes internal streams so that the drain of one launch (a few
    // long-running CTAs at the set boundary) overlaps the head of the next; in one stream every
    // launch boundary would idle most of the GPU.  Tiles never overlap, so concurrency is safe.
    // The lanes start after everything already enqueued on `stream`, and `stream` waits for the
    // lanes before this function returns, so callers keep plain stream semantics.
    constexpr uint32_t kLanes = 4;
    EventSet evs;                                    // [0] = fence, [1 .. depth] = ring; destroyed on every exit path
    if (evs.add(1 + depth)) return fail(MB_ERR_CUDA, "cudaEventCreate failed");
    cudaEvent_t fence = evs.ev[0], *ring = &evs.ev[1];
    uint32_t claims = 0, stolen_total = 0;
    int result = MB_OK;
    cudaError_t e = cudaEventRecord(fence, st);
    for (uint32_t k = 0; k < kLanes && e == cudaSuccess; ++k) e = cudaStreamWaitEvent(c->lanes[k], fence, 0);
    if (e != cudaSuccess) result = fail(MB_ERR_CUDA, "lane fence: %s", cudaGetErrorString(e));
    g_first_claim = g_last_claim = now_monotonic();
    while (result == MB_OK) {
        const int64_t remaining = mb_wsq_remaining(q);
        uint32_t want = (uint32_t)(remaining / ((int64_t)n_ranks * 8));
        if (want > bands_per_claim) want = bands_per_claim;
        if (want == 0) want = 1;
        uint32_t first = 0, got = 0, stolen = 0;
        const int pr = mb_wsq_pop(q, rank, want, &first, &got, &stolen);
        if (pr < 0) { result = pr; break; }
        if (pr == 0) break;
        if (claims == 0) g_first_claim = now_monotonic();
        g_last_claim = now_monotonic();
        if ((uint64_t)first + got > n_tiles) { result = fail(MB_ERR_STATE, "queue handed out tile %u of %u", first + got - 1, n_tiles); break; }
        if (claims >= max_claims) { result = fail(MB_ERR_INVALID, "claimed[] too small (%u claims)", max_claims); break; }
        if (claims >= depth) {                              // at most `depth` launches in flight
            e = cudaEventSynchronize(ring[claims % depth]);
            if (e != cudaSuccess) { result = fail(MB_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e)); break; }
        }
        cudaStream_t lane = c->lanes[claims % kLanes];
        if ((result = mb_render_counts(tiles + first, got, out_dev, dtype, lane))) break;
        e = cudaEventRecord(ring[claims % depth], lane);
        if (e != cudaSuccess) { result = fail(MB_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e)); break; }
        claimed[2 * claims] = first;
        claimed[2 * claims + 1] = got;
        ++claims;
        stolen_total += stolen;
    }
    // every exit path: join the lanes back into the caller's stream
    for (uint32_t k = 0; k < kLanes; ++k)
        if (cudaEventRecord(fence, c->lanes[k]) == cudaSuccess) cudaStreamWaitEvent(st, fence, 0);
    *n_claims = claims;
    if (n_stolen) *n_stolen = stolen_total;
    return result;
}

int mb_frame_last_stats(double *first_claim_s, double *last_claim_s)
{
    if (first_claim_s) *first_claim_s = g_first_claim;
    if (last_claim_s) *last_claim_s = g_last_claim;
    return MB_OK;
}

// ---- band-granular rendering into a shared pinned host frame (8e, end-to-end form) ------------------

int mb_internal_hostframe_region(mb_hostframe *f, void **base, size_t *bytes, int **pinned_flag);   // host_frame.cpp

int mb_hostframe_pin(mb_hostframe *f)
{
    void *base = nullptr;
    size_t bytes = 0;
    int *pinned = nullptr;
    if (mb_internal_hostframe_region(f, &base, &bytes, &pinned)) return fail(MB_ERR_INVALID, "NULL frame");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    if (*pinned) return MB_OK;
    // page-lock this process's mapping of the segment: device->host copies into it are plain DMA
    MB_CUDA(cudaHostRegister(base, bytes, cudaHostRegisterPortable));
    *pinned = 1;
    return MB_OK;
}

int mb_hostframe_unpin(mb_hostframe *f)
{
    void *base = nullptr;
    size_t bytes = 0;
    int *pinned = nullptr;
    if (mb_internal_hostframe_region(f, &base, &bytes, &pinned)) return fail(MB_ERR_INVALID, "NULL frame");
    if (!*pinned) return MB_OK;
    MB_CUDA(cudaHostUnregister(base));
    *pinned = 0;
    return MB_OK;
}

}  // extern "C"

namespace {

constexpr uint32_t kMaxClaim = 64;

// one claim of consecutive band tiles: render into the slot's device staging buffer, then copy
// every band to its final offset of the host frame (adjacent bands merge into one DMA)
int claim_to_host(const mb_tile *tiles, uint32_t first, uint32_t got, uint8_t *stage, uint8_t *host, uint64_t host_bytes,
                  int dtype, cudaStream_t lane)
{
    mb_tile tmp[kMaxClaim];
    size_t off = 0;
    for (uint32_t k = 0; k < got; ++k) {
        tmp[k] = tiles[first + k];
        const size_t sz = (size_t)tmp[k].w * tmp[k].h;
        if (tmp[k].out_pitch != tmp[k].w) return fail(MB_ERR_INVALID, "tile %u: host-frame bands must be contiguous (out_pitch == w)", first + k);
        if (tmp[k].out_offset + sz > host_bytes) return fail(MB_ERR_INVALID, "tile %u: band ends beyond the host frame", first + k);
        tmp[k].out_offset = off;
        off += sz;
    }
    int rc = mb_render_u8(tmp, got, stage, dtype, lane);
    if (rc) return rc;
    for (uint32_t k = 0; k < got;) {
        uint32_t m = k + 1;
        size_t bytes = (size_t)tmp[k].w * tmp[k].h;
        while (m < got && tiles[first + m].out_offset == tiles[first + k].out_offset + bytes) { bytes += (size_t)tmp[m].w * tmp[m].h; ++m; }
        MB_CUDA(cudaMemcpyAsync(host + tiles[first + k].out_offset, stage + tmp[k].out_offset, bytes, cudaMemcpyDeviceToHost, lane));
        k = m;
    }
    return MB_OK;
}

// shared driver: `next` yields (first, got) claims until it returns 0
template <typename Next>
int render_to_host(DeviceCtx *c, const mb_tile *tiles, uint32_t n_tiles, uint32_t bands_per_claim, uint32_t depth,
                   mb_hostframe *frame, int dtype, uint32_t *claimed, uint32_t max_claims, uint32_t *n_claims,
                   uint32_t *completed, uint32_t max_completed, uint32_t *n_completed, Next next)
{
    constexpr uint32_t kLanes = 4;
    if (bands_per_claim == 0) bands_per_claim = 1;
    if (bands_per_claim > kMaxClaim) bands_per_claim = kMaxClaim;
    if (depth < 2) depth = 2;
    if (depth > 16) depth = 16;
    size_t max_band = 0;
    for (uint32_t k = 0; k < n_tiles; ++k) {
        const size_t sz = (size_t)tiles[k].w * tiles[k].h;
        max_band = sz > max_band ? sz : max_band;
    }
    const size_t need = max_band * bands_per_claim;
    for (uint32_t k = 0; k < depth; ++k) {
        if (c->stage_bytes[k] >= need) continue;
        if (c->stage[k]) { MB_CUDA(cudaFree(c->stage[k])); c->stage[k] = nullptr; c->stage_bytes[k] = 0; }
        MB_CUDA(cudaMalloc(&c->stage[k], need));
        c->stage_bytes[k] = need;
    }
    uint8_t *host = mb_hostframe_data(frame);
    const uint64_t host_bytes = mb_hostframe_bytes(frame);
    EventSet evs;
    if (evs.add(depth)) return fail(MB_ERR_CUDA, "cudaEventCreate failed");
    uint32_t slot_first[16] = {}, slot_got[16] = {};
    uint32_t claims = 0, retired = 0, done = 0;
    int result = MB_OK;
    auto retire = [&](uint32_t k) -> int {               // claim k has landed in host memory: publish it
        cudaError_t e = cudaEventSynchronize(evs.ev[k % depth]);
        if (e != cudaSuccess) return fail(MB_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
        return mb_hostframe_mark_done(frame, slot_first[k % depth], slot_got[k % depth], completed, max_completed, &done);
    };
    g_first_claim = g_last_claim = now_monotonic();
    for (;;) {
        uint32_t first = 0, got = 0;
        const int pr = next(bands_per_claim, &first, &got);
        if (pr < 0) { result = pr; break; }
        if (pr == 0) break;
        if (claims == 0) g_first_claim = now_monotonic();
        g_last_claim = now_monotonic();
        if ((uint64_t)first + got > n_tiles || got > kMaxClaim) { result = fail(MB_ERR_STATE, "bad claim [%u, %u)", first, first + got); break; }
        if (claimed && claims >= max_claims) { result = fail(MB_ERR_INVALID, "claimed[] too small (%u claims)", max_claims); break; }
        if (claims >= depth) { if ((result = retire(retired))) break; ++retired; }     // frees this slot's staging buffer
        cudaStream_t lane = c->lanes[claims % kLanes];
        if ((result = claim_to_host(tiles, first, got, static_cast<uint8_t *>(c->stage[claims % depth]), host, host_bytes, dtype, lane))) break;
        cudaError_t e = cudaEventRecord(evs.ev[claims % depth], lane);
        if (e != cudaSuccess) { result = fail(MB_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e)); break; }
        slot_first[claims % depth] = first;
        slot_got[claims % depth] = got;
        if (claimed) { claimed[2 * claims] = first; claimed[2 * claims + 1] = got; }
        ++claims;
    }
    for (; retired < claims; ++retired) {                 // drain in submission order, also after an error
        const int r = retire(retired);
        if (r && !result) result = r;
    }
    if (result) for (uint32_t k = 0; k < kLanes; ++k) cudaStreamSynchronize(c->lanes[k]);
    if (n_claims) *n_claims = claims;
    if (n_completed) *n_completed = done;
    return result;
}

}  // namespace

extern "C" {

int mb_frame_render_u8_host(mb_wsq *q, uint32_t rank, uint32_t n_ranks, const mb_tile *tiles, uint32_t n_tiles,
                            uint32_t bands_per_claim, uint32_t depth, mb_hostframe *frame, int dtype, uint32_t *claimed,
                            uint32_t max_claims, uint32_t *n_claims, uint32_t *n_stolen, uint32_t *completed,
                            uint32_t max_completed, uint32_t *n_completed)
{
    if (!q || !tiles || !frame) return fail(MB_ERR_INVALID, "NULL argument");
    if (n_ranks == 0 || rank >= n_ranks) return fail(MB_ERR_INVALID, "rank %u out of range", rank);
    if (mb_hostframe_units(frame) < n_tiles) return fail(MB_ERR_INVALID, "the host frame has %u units, the tile list %u", mb_hostframe_units(frame), n_tiles);
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    uint32_t stolen_total = 0;
    rc = render_to_host(c, tiles, n_tiles, bands_per_claim, depth, frame, dtype, claimed, max_claims, n_claims, completed,
                        max_completed, n_completed, [&](uint32_t cap, uint32_t *first, uint32_t *got) -> int {
                            // guided self-scheduling: full claims while plenty is left, single bands at the end
                            uint32_t want = (uint32_t)(mb_wsq_remaining(q) / ((int64_t)n_ranks * 8));
                            if (want > cap) want = cap;
                            if (want == 0) want = 1;
                            uint32_t stolen = 0;
                            const int pr = mb_wsq_pop(q, rank, want, first, got, &stolen);
                            stolen_total += stolen;
                            return pr;
                        });
    if (n_stolen) *n_stolen = stolen_total;
    return rc;
}

int mb_frame_render_u8_host_units(const uint32_t *units, uint32_t n_units, const mb_tile *tiles, uint32_t n_tiles,
                                  uint32_t depth, mb_hostframe *frame, int dtype, uint32_t *completed,
                                  uint32_t max_completed, uint32_t *n_completed)
{
    if (!tiles || !frame || (!units && n_units)) return fail(MB_ERR_INVALID, "NULL argument");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    uint32_t pos = 0;
    return render_to_host(c, tiles, n_tiles, 1, depth, frame, dtype, nullptr, 0, nullptr, completed, max_completed, n_completed,
                          [&](uint32_t, uint32_t *first, uint32_t *got) -> int {
                              if (pos >= n_units) return 0;
                              *first = units[pos++];
                              *got = 1;
                              return 1;
                          });
}

// ---- serialisation (f2) --------------------------------------------------------------------

static size_t codec_tiles(size_t n) { return (n + mbc::kTile - 1) / mbc::kTile; }
static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t mb_serialize_out_stride(size_t n) { return round_up(n + 1, 256); }

size_t mb_serialize_workspace_bytes(size_t n, uint32_t n_chunks)
{
    // per chunk: one 16-byte look-back status per 8 KiB tile, and a 4-byte "RLE ruled out" hint
    return 256 + round_up((size_t)n_chunks * (sizeof(mbc::TileStatus) * codec_tiles(n) + sizeof(uint32_t)), 256);
}

int mb_serialize_chunks(const uint8_t *payloads_dev, size_t n, uint32_t n_chunks, size_t payload_stride, uint8_t *out_dev,
                        size_t out_stride, void *workspace_dev, size_t workspace_bytes, mb_chunk_info *info_dev, void *stream)
{
    if (!payloads_dev || !out_dev || !workspace_dev || !info_dev) return fail(MB_ERR_INVALID, "NULL buffer");
    if (n == 0 || n >= 0xffffffffull) return fail(MB_ERR_INVALID, "payload size must be in [1, 2^32)");
    if (n_chunks == 0 || n_chunks > 65535) return fail(MB_ERR_INVALID, "n_chunks must be in [1, 65535]");
    if (payload_stride < n) return fail(MB_ERR_INVALID, "payload_stride < n");
    if (out_stride < n + 1) return fail(MB_ERR_INVALID, "out_stride must be >= n + 1");
    if (workspace_bytes < mb_serialize_workspace_bytes(n, n_chunks)) return fail(MB_ERR_INVALID, "workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 15u) return fail(MB_ERR_INVALID, "workspace must be 16-byte aligned");
    static_assert(sizeof(mb_chunk_info) == sizeof(mbc::Meta), "mb_chunk_info mirrors mbc::Meta");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    mbc::Batch b;
    b.payloads = payloads_dev;
    b.payload_stride = payload_stride;
    b.out = out_dev;
    b.out_stride = out_stride;
    b.status = reinterpret_cast<mbc::TileStatus *>(static_cast<uint8_t *>(workspace_dev) + 256);
    b.meta = reinterpret_cast<mbc::Meta *>(info_dev);
    b.n = n;
    constexpr int kSpans = 2;                                      // 16 KiB tiles: measured best (profiles/r02_codec_bench.json)
    const size_t tile = (size_t)mbc::tile_bytes(kSpans);
    b.ntiles = (uint32_t)((n + tile - 1) / tile);
    b.n_chunks = n_chunks;
    b.dead = reinterpret_cast<uint32_t *>(b.status + (size_t)n_chunks * b.ntiles);
    // the look-back flags must start out invalid, the hints clear
    MB_CUDA(cudaMemsetAsync(b.status, 0, (size_t)n_chunks * (b.ntiles * sizeof(mbc::TileStatus) + sizeof(uint32_t)), st));
    // blockIdx.x = chunk (fastest), tiles along y and, beyond 65535 of them, z
    const uint32_t gy = b.ntiles < 65535u ? b.ntiles : 65535u, gz = (b.ntiles + gy - 1) / gy;
    mbc::rle_tiles<kSpans><<<dim3(n_chunks, gy, gz), mbc::kThreads, mbc::smem_bytes(kSpans), st>>>(b);   // everything, for chunks that end up RLE
    b.ntiles = (uint32_t)codec_tiles(n);
    mbc::raw_fixup<<<dim3((b.ntiles + mbc::kRawPieces - 1) / mbc::kRawPieces, n_chunks, 1), mbc::kThreads, 0, st>>>(b);   // exits at once unless the chunk ended up Raw
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

// scratch layout of the host-pointer forms: info | payload[n] | out[n+1] (256B aligned) | workspace
static int serialize_scratch(DeviceCtx *c, size_t n, mb_chunk_info **info, uint8_t **payload, uint8_t **out, uint8_t **ws, size_t *ws_bytes)
{
    const size_t a = round_up(n, 256), b = mb_serialize_out_stride(n);
    *ws_bytes = mb_serialize_workspace_bytes(n, 1);
    int rc = ensure_scratch(c, 256 + a + b + *ws_bytes);
    if (rc) return rc;
    *info = static_cast<mb_chunk_info *>(c->scratch);
    *payload = static_cast<uint8_t *>(c->scratch) + 256;
    *out = *payload + a;
    *ws = *out + b;
    return MB_OK;
}

int mb_serialize_chunk_host(const uint8_t *payload_host, size_t n, uint8_t *out_host, size_t out_capacity, uint64_t *out_len)
{
    if (!payload_host || !out_host || !out_len) return fail(MB_ERR_INVALID, "NULL buffer");
    if (n == 0 || n >= 0xffffffffull) return fail(MB_ERR_INVALID, "payload size must be in [1, 2^32)");
    if (out_capacity < n + 1) return fail(MB_ERR_INVALID, "out_capacity must be >= n + 1");
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    mb_chunk_info *info;
    uint8_t *p, *o, *ws;
    size_t wsb;
    if ((rc = serialize_scratch(c, n, &info, &p, &o, &ws, &wsb))) return rc;
    MB_CUDA(cudaMemcpyAsync(p, payload_host, n, cudaMemcpyHostToDevice, c->stream));
    if ((rc = mb_serialize_chunks(p, n, 1, n, o, mb_serialize_out_stride(n), ws, wsb, info, c->stream))) return rc;
    mb_chunk_info h;
    MB_CUDA(cudaMemcpyAsync(&h, info, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    MB_CUDA(cudaMemcpyAsync(out_host, o, h.out_len, cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    *out_len = h.out_len;
    return MB_OK;
}

int mb_process_workload_serialized(uint32_t level, uint32_t mrd, uint32_t index_real, uint32_t index_imag,
                                   uint32_t definition, uint8_t *out_host, size_t out_capacity, uint64_t *out_len,
                                   int32_t *entry_type)
{
    if (!out_host || !out_len || !entry_type) return fail(MB_ERR_INVALID, "NULL buffer");
    mb_tile t;
    int rc = mb_chunk_geometry(level, index_real, index_imag, definition, mrd, &t);
    if (rc) return rc;
    t.flags = g_tile_flags;
    const size_t n = (size_t)definition * definition;
    if (out_capacity < n + 1) return fail(MB_ERR_INVALID, "out_capacity must be >= definition^2 + 1");
    DeviceCtx *c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    mb_chunk_info *info;
    uint8_t *p, *o, *ws;
    size_t wsb;
    if ((rc = serialize_scratch(c, n, &info, &p, &o, &ws, &wsb))) return rc;
    if ((rc = mb_render_u8(&t, 1, p, MB_DTYPE_F64, c->stream))) return rc;
    if ((rc = mb_serialize_chunks(p, n, 1, n, o, mb_serialize_out_stride(n), ws, wsb, info, c->stream))) return rc;
    mb_chunk_info h;
    MB_CUDA(cudaMemcpyAsync(&h, info, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    // Never/Immediate chunks have no file (DataStorage.cs:424-425) but the 6 bytes are cheap to return anyway
    MB_CUDA(cudaMemcpyAsync(out_host, o, h.out_len, cudaMemcpyDeviceToHost, c->stream));
    MB_CUDA(cudaStreamSynchronize(c->stream));
    *out_len = h.out_len;
    *entry_type = (int32_t)h.entry_type;   // classification falls out of the run count (chunk_codec.cuh)
    return MB_OK;
}

int mb_measure_pipe_rate(int dtype, int mode, int iters, double *tera_ops_per_s, float *elapsed_ms)
{
    if (!tera_ops_per_s) return fail(MB_ERR_INVALID, "NULL result pointer");
    if (iters <= 0) return fail(MB_ERR_INVALID, "iters must be positive");
    if (dtype != MB_DTYPE_F64 && dtype != MB_DTYPE_F32) return fail(MB_ERR_INVALID, "unknown dtype %d", dtype);
    if (mode != 0 && mode != 1) return fail(MB_ERR_INVALID, "unknown mode %d", mode);
    DeviceCtx *c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((rc = ensure_scratch(c, 64))) return rc;
    constexpr int kChains = 8, kThreads = 256, kCtasPerSm = 8;
    const int grid = c->sm_count * kCtasPerSm;
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    float ms = 0.f;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
        MB_CUDA(cudaEventRecord(e0, c->stream));
        if (dtype == MB_DTYPE_F64) {
            if (mode == 0) mbk::fma_peak_kernel<double, kChains><<<grid, kThreads, 0, c->stream>>>((double *)c->scratch, iters, 1.0);
            else mbk::mandel_mix_kernel<double><<<grid, kThreads, 0, c->stream>>>((double *)c->scratch, iters, 1.0);
        } else {
            if (mode == 0) mbk::fma_peak_kernel<float, kChains><<<grid, kThreads, 0, c->stream>>>((float *)c->scratch, iters, 1.0f);
            else mbk::mandel_mix_kernel<float><<<grid, kThreads, 0, c->stream>>>((float *)c->scratch, iters, 1.0f);
        }
        MB_CUDA(cudaGetLastError());
        MB_CUDA(cudaEventRecord(e1, c->stream));
        MB_CUDA(cudaEventSynchronize(e1));
        MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double per_thread = mode == 0 ? (double)kChains * 8.0 * iters : 7.0 * 8.0 * iters;
    const double ops = per_thread * (double)kThreads * (double)grid;
    *tera_ops_per_s = ops / ((double)ms * 1e-3) / 1e12;
    if (elapsed_ms) *elapsed_ms = ms;
    return MB_OK;
}

int mb_measure_fma_peak(int dtype, int iters, double *tera_fma_per_s, float *elapsed_ms)
{
    return mb_measure_pipe_rate(dtype, 0, iters, tera_fma_per_s, elapsed_ms);
}

}  // extern "C"
