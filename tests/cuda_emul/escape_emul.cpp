// escape_emul.cpp -- TEST INFRASTRUCTURE: the escape-time kernels of the hot path (escape_kernels.cuh,
// escape_refill.cuh, unmodified text, same template arguments as csrc/mandelbrot_capi.cu launches them with)
// executed on the CPU through cuda_emul.h, behind C entry points shaped like mb_render_counts / mb_render_u8 /
// mb_calc_mb_value.  Built with -ffp-contract=off: every *_rn intrinsic is one IEEE operation here as on the device.
#define MB_EMULATE 1
#include "cuda_emul.h"

#include "escape_refill.cuh"      // includes escape_kernels.cuh

namespace {
// csrc/mandelbrot_capi.cu: kWX, kWY, kMinCtas, kHead, kRefillCtas
constexpr int kWX = 2, kWY = 2, kMinCtas = 12, kHead = 32, kRefillCtas = 8;

template <typename OutT>
struct Launch {
    mbk::TileArgs args;
    OutT *out;
};

template <typename Real, typename OutT, bool PERIOD>
void run_tiles(const void *a)
{
    const auto *l = static_cast<const Launch<OutT> *>(a);
    mbk::escape_tiles_kernel<Real, OutT, kHead, kWX, kWY, kMinCtas, PERIOD>(l->args, l->out);
}
template <typename Real, typename OutT, bool PERIOD>
void run_refill(const void *a)
{
    const auto *l = static_cast<const Launch<OutT> *>(a);
    mbk::escape_refill_kernel<Real, OutT, kRefillCtas, PERIOD>(l->args, l->out);
}

template <typename OutT>
int render(const mb_tile *tiles, uint32_t n_tiles, OutT *out, int dtype, unsigned long long *executed)
{
    uint32_t mw = 0, mh = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        mw = tiles[t].w > mw ? tiles[t].w : mw;
        mh = tiles[t].h > mh ? tiles[t].h : mh;
    }
    const bool refill = (tiles[0].flags & MB_TILE_LANE_REFILL) != 0, period = (tiles[0].flags & MB_TILE_PERIOD_SHORTCUT) != 0;
    Launch<OutT> l;
    memset(&l, 0, sizeof l);
    l.out = out;
    l.args.executed = executed;
    l.args.dev = tiles;                                       // descriptors are always read through args.dev here ...
    if (n_tiles <= (uint32_t)mbk::kInlineTiles) {             // ... except where the product inlines them, as it does
        l.args.n_inline = n_tiles;
        memcpy(l.args.inl, tiles, sizeof(mb_tile) * n_tiles);
    }
    unsigned long long counter = 0;
    dim3 grid;
    unsigned threads;
    void (*fn)(const void *);
    const bool f64 = dtype == MB_DTYPE_F64;
    if (refill) {
        l.args.counter = &counter;
        l.args.n_tiles = n_tiles;
        l.args.pcols = (mw + mbk::kWarpW - 1) / mbk::kWarpW;
        l.args.prows = (mh + mbk::kWarpH - 1) / mbk::kWarpH;
        const unsigned long long slots = (unsigned long long)l.args.pcols * l.args.prows * n_tiles;
        unsigned long long ctas = (slots + 3) / 4;
        if (ctas > 3) ctas = 3;                               // CTAs run one after the other here: the first drains the counter
        grid = dim3((unsigned)ctas, 1, 1);
        threads = 128;
        fn = f64 ? (period ? run_refill<double, OutT, true> : run_refill<double, OutT, false>)
                 : (period ? run_refill<float, OutT, true> : run_refill<float, OutT, false>);
    } else {
        const uint32_t cw = kWX * mbk::kWarpW, ch = kWY * mbk::kWarpH;
        grid = dim3(((mw + cw - 1) / cw) * ((mh + ch - 1) / ch), n_tiles, 1);
        threads = kWX * kWY * 32;
        fn = f64 ? (period ? run_tiles<double, OutT, true> : run_tiles<double, OutT, false>)
                 : (period ? run_tiles<float, OutT, true> : run_tiles<float, OutT, false>);
    }
    emul::launch(grid, threads, fn, &l);
    return 0;
}

struct Points {
    const double *r, *i;
    size_t n;
    int mrd;
    int32_t *out;
};
template <typename Real>
void run_points(const void *a)
{
    const auto *p = static_cast<const Points *>(a);
    mbk::escape_points_kernel<Real, kHead>(p->r, p->i, p->n, p->mrd, p->out);
}
struct U8Args {
    const int32_t *counts;
    uint8_t *out;
    size_t n;
    uint32_t mrd;
};
void run_counts_to_u8(const void *a)
{
    const auto *p = static_cast<const U8Args *>(a);
    mbk::counts_to_u8_kernel(p->counts, p->out, p->n, p->mrd);
}
struct SumArgs {
    const int32_t *counts;
    size_t n;
    uint32_t mrd;
    unsigned long long *sum;
};
void run_useful(const void *a)
{
    const auto *p = static_cast<const SumArgs *>(a);
    mbk::useful_iterations_kernel(p->counts, p->n, p->mrd, p->sum);
}
struct ClassArgs {
    const uint8_t *p;
    size_t n;
    uint32_t *flags;
};
void run_classify(const void *a)
{
    const auto *p = static_cast<const ClassArgs *>(a);
    mbk::classify_u8_kernel(p->p, p->n, p->flags);
}
}  // namespace

// the small HBM-bound helpers (mb_counts_to_u8, mb_useful_iterations, mb_classify_u8), on a few CTAs
extern "C" int emul_counts_to_u8(const int32_t *counts, uint8_t *out, size_t n, uint32_t mrd)
{
    U8Args a{counts, out, n, mrd};
    emul::launch(dim3(3, 1, 1), 256, run_counts_to_u8, &a);
    return 0;
}
extern "C" unsigned long long emul_useful_iterations(const int32_t *counts, size_t n, uint32_t mrd)
{
    unsigned long long sum = 0;
    SumArgs a{counts, n, mrd, &sum};
    emul::launch(dim3(3, 1, 1), 256, run_useful, &a);
    return sum;
}
extern "C" int emul_classify_u8(const uint8_t *p, size_t n, uint32_t *flags)
{
    flags[0] = flags[1] = 1;                                  // DataChunk.cs:82,87: the kernel only clears
    ClassArgs a{p, n, flags};
    emul::launch(dim3(3, 1, 1), 256, run_classify, &a);
    return 0;
}

// the device's coordinate generator (in-kernel gen_arrays, .py:24-32) for every sample of one axis of a tile
extern "C" void emul_tile_axis(const mb_tile *t, int axis, double *out)
{
    const uint32_t d1 = t->definition - 1u;
    for (uint32_t k = 0; k < t->definition; ++k)
        out[k] = axis == 0 ? mbk::linspace_at(k, d1, t->start_r, t->step_r, t->stop_r) : mbk::linspace_at(k, d1, t->start_i, t->step_i, t->stop_i);
}

extern "C" int emul_render_counts(const mb_tile *tiles, uint32_t n_tiles, int32_t *out, int dtype, unsigned long long *executed)
{
    if (!tiles || !n_tiles || !out) return 1;
    return render<int32_t>(tiles, n_tiles, out, dtype, executed);
}
extern "C" int emul_render_u8(const mb_tile *tiles, uint32_t n_tiles, uint8_t *out, int dtype, unsigned long long *executed)
{
    if (!tiles || !n_tiles || !out) return 1;
    return render<uint8_t>(tiles, n_tiles, out, dtype, executed);
}
extern "C" int emul_calc_mb_value(const double *r, const double *i, size_t n, int mrd, int32_t *out, int dtype)
{
    if (!r || !i || !out) return 1;
    Points p{r, i, n, mrd, out};
    emul::launch(dim3((unsigned)((n + 127) / 128), 1, 1), 128, dtype == MB_DTYPE_F64 ? run_points<double> : run_points<float>, &p);
    return 0;
}
