/*
 * mandel_oracle_simd.c -- data-parallel build of the CPU oracle (TEST INFRASTRUCTURE ONLY; the
 * rules of mandel_oracle.h apply: never imported, linked or called by the product).
 *
 * Same restatement of DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py:39-68 as
 * mbo_escape_f64 / mbo_escape_f32 in mandel_oracle.c, evaluated on 8 (binary64) or 16 (binary32)
 * neighbouring pixels of a row at once with GCC vector extensions.  Every vector operation is the
 * lane-wise IEEE-754 operation, individually rounded (this file is also built -ffp-contract=off,
 * and GCC never contracts across statements without it being allowed), so every lane computes
 * exactly the scalar sequence; lanes that have escaped keep iterating but are masked out.
 *
 * Why it exists: the scalar oracle needs ~2.7 h on 8 cores for the 65536^2 / mrd 10000 frame.
 * This one renders WHOLE native 4096^2 chunks of cfg3/cfg4/cfg5 fast enough to pin every
 * full-size constant in tests/ and bench.py to the oracle instead of to the CUDA build
 * (oracle/gen_fullsize_golden.py -> tests/golden/fullsize_*.json).  It is itself pinned
 * bit-for-bit against the scalar oracle (tests/test_oracle_golden.py::test_simd_oracle_*).
 *
 * target_clones: one binary carries AVX-512, AVX2 and baseline x86-64 bodies and picks at load
 * time, so the .so built in the authoring container runs on whatever host CPU the GPU box has.
 */
#include "mandel_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#ifdef __FAST_MATH__
#error "the oracle must not be built with -ffast-math"
#endif

typedef double v8d __attribute__((vector_size(64)));
typedef long long v8l __attribute__((vector_size(64)));
typedef float v16f __attribute__((vector_size(64)));
typedef int v16i __attribute__((vector_size(64)));

#if defined(__x86_64__)
#define MBO_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define MBO_CLONES
#endif

/* .py:44-68 on 2 x 8 pixels (two independent vectors hide the latency of the dependent chain).
 * res lanes: first n with sqr_mag >= 4, else 0.  `n_valid` lanes of the 16 are real pixels. */
MBO_CLONES
static void escape16_f64(const double *cr, double ci, int32_t mrd, int32_t *res)
{
    v8d c0a, c0b, c1;
    for (int k = 0; k < 8; ++k) { c0a[k] = cr[k]; c0b[k] = cr[8 + k]; c1[k] = ci; }
    v8d z0a = c0a, z1a = c1, z0b = c0b, z1b = c1;                 /* z = c  (.py:44-45) */
    v8l ra = {0, 0, 0, 0, 0, 0, 0, 0}, rb = ra;
    v8l live_a = {-1, -1, -1, -1, -1, -1, -1, -1}, live_b = live_a;
    const v8d two = {2, 2, 2, 2, 2, 2, 2, 2}, four = {4, 4, 4, 4, 4, 4, 4, 4};
    for (int32_t n = 1; n < mrd; ++n) {                           /* .py:47 */
        v8d aa = z0a * z0a, ba = z1a * z1a;                       /* .py:50-53 */
        v8d ab = z0b * z0b, bb = z1b * z1b;
        v8d sq0a = aa - ba, sq0b = ab - bb;
        v8d ta = two * z0a, tb = two * z0b;
        v8d sq1a = ta * z1a, sq1b = tb * z1b;
        z0a = sq0a + c0a; z1a = sq1a + c1;                        /* .py:56-59 */
        z0b = sq0b + c0b; z1b = sq1b + c1;
        v8d ma = z0a * z0a + z1a * z1a;                           /* .py:62 (three roundings) */
        v8d mb = z0b * z0b + z1b * z1b;
        v8l ea = (v8l)(ma >= four) & live_a;                      /* .py:65: ordered compare, false for NaN */
        v8l eb = (v8l)(mb >= four) & live_b;
        const v8l nn = {n, n, n, n, n, n, n, n};
        ra |= nn & ea;
        rb |= nn & eb;
        live_a &= ~ea;
        live_b &= ~eb;
        v8l any = live_a | live_b;
        long long acc = 0;
        for (int k = 0; k < 8; ++k) acc |= any[k];
        if (!acc) break;
    }
    for (int k = 0; k < 8; ++k) { res[k] = (int32_t)ra[k]; res[8 + k] = (int32_t)rb[k]; }
}

MBO_CLONES
static void escape32_f32(const float *cr, float ci, int32_t mrd, int32_t *res)
{
    v16f c0a, c0b, c1;
    for (int k = 0; k < 16; ++k) { c0a[k] = cr[k]; c0b[k] = cr[16 + k]; c1[k] = ci; }
    v16f z0a = c0a, z1a = c1, z0b = c0b, z1b = c1;
    v16i ra, rb, live_a, live_b;
    v16f two, four;
    for (int k = 0; k < 16; ++k) { ra[k] = 0; rb[k] = 0; live_a[k] = -1; live_b[k] = -1; two[k] = 2.0f; four[k] = 4.0f; }
    for (int32_t n = 1; n < mrd; ++n) {
        v16f aa = z0a * z0a, ba = z1a * z1a;
        v16f ab = z0b * z0b, bb = z1b * z1b;
        v16f sq0a = aa - ba, sq0b = ab - bb;
        v16f ta = two * z0a, tb = two * z0b;
        v16f sq1a = ta * z1a, sq1b = tb * z1b;
        z0a = sq0a + c0a; z1a = sq1a + c1;
        z0b = sq0b + c0b; z1b = sq1b + c1;
        v16f ma = z0a * z0a + z1a * z1a;
        v16f mb = z0b * z0b + z1b * z1b;
        v16i ea = (v16i)(ma >= four) & live_a;
        v16i eb = (v16i)(mb >= four) & live_b;
        v16i nn;
        for (int k = 0; k < 16; ++k) nn[k] = n;
        ra |= nn & ea;
        rb |= nn & eb;
        live_a &= ~ea;
        live_b &= ~eb;
        v16i any = live_a | live_b;
        int acc = 0;
        for (int k = 0; k < 16; ++k) acc |= any[k];
        if (!acc) break;
    }
    for (int k = 0; k < 16; ++k) { res[k] = ra[k]; res[16 + k] = rb[k]; }
}

/* one row of w pixels; ragged ends are padded with the row's last coordinate and discarded */
static void simd_row_f64(const double *r, double ci, uint32_t w, int32_t mrd, int32_t *o)
{
    double cbuf[16];
    int32_t rbuf[16];
    for (uint32_t col = 0; col < w; col += 16) {
        const uint32_t m = w - col < 16 ? w - col : 16;
        if (m == 16) {
            escape16_f64(r + col, ci, mrd, o + col);
        } else {
            for (uint32_t k = 0; k < 16; ++k) cbuf[k] = r[col + (k < m ? k : m - 1)];
            escape16_f64(cbuf, ci, mrd, rbuf);
            memcpy(o + col, rbuf, sizeof(int32_t) * m);
        }
    }
}

static void simd_row_f32(const double *r, double ci, uint32_t w, int32_t mrd, int32_t *o)
{
    float cbuf[32];
    int32_t rbuf[32];
    const float cif = (float)ci;                /* Appendix A.5: binary64 coordinate rounded once */
    for (uint32_t col = 0; col < w; col += 32) {
        const uint32_t m = w - col < 32 ? w - col : 32;
        for (uint32_t k = 0; k < 32; ++k) cbuf[k] = (float)r[col + (k < m ? k : m - 1)];
        escape32_f32(cbuf, cif, mrd, rbuf);
        memcpy(o + col, rbuf, sizeof(int32_t) * m);
    }
}

typedef struct {
    const double *r, *im;
    int32_t *out;
    size_t pitch;
    uint32_t x0, y0, w, h;
    int32_t mrd;
    int dtype;
    atomic_uint next_row;
} simd_job;

static void *simd_worker(void *arg)
{
    simd_job *j = (simd_job *)arg;
    for (;;) {
        const uint32_t row = atomic_fetch_add(&j->next_row, 1u);
        if (row >= j->h) break;
        int32_t *o = j->out + (size_t)row * j->pitch;
        if (j->dtype == 0) simd_row_f64(j->r + j->x0, j->im[j->y0 + row], j->w, j->mrd, o);
        else simd_row_f32(j->r + j->x0, j->im[j->y0 + row], j->w, j->mrd, o);
    }
    return NULL;
}

/* same contract as mbo_render_rect (mandel_oracle.c) */
int mbo_simd_render_rect(double start_r, double start_i, double range, uint32_t d, int32_t mrd,
                         uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                         int32_t *out, size_t pitch, int dtype, int threads)
{
    if (d == 0 || (uint64_t)x0 + w > d || (uint64_t)y0 + h > d) return -1;
    if (dtype != 0 && dtype != 1) return -2;
    if (w == 0 || h == 0) return 0;
    double *r = (double *)malloc(sizeof(double) * d);
    double *im = (double *)malloc(sizeof(double) * d);
    if (!r || !im) { free(r); free(im); return -3; }
    volatile double stop_r = start_r + range;       /* .py:24-32: stop = start + _range */
    volatile double stop_i = start_i + range;
    mbo_linspace(start_r, stop_r, d, r);
    mbo_linspace(start_i, stop_i, d, im);
    simd_job job = {r, im, out, pitch, x0, y0, w, h, mrd, dtype, 0};
    int nt = threads > 0 ? threads : mbo_max_threads();
    if ((uint32_t)nt > h) nt = (int)h;
    pthread_t *th = nt > 1 ? (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt) : NULL;
    int started = 0;
    if (th)
        for (; started < nt - 1; ++started)
            if (pthread_create(&th[started], NULL, simd_worker, &job) != 0) break;
    simd_worker(&job);
    for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
    free(th);
    free(r);
    free(im);
    return 0;
}

/* which body the loader picks on this CPU: 512 / 256 / 128 (bits per vector register) */
int mbo_simd_width(void)
{
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return 512;
    if (__builtin_cpu_supports("avx2")) return 256;
#endif
    return 128;
}

/* whole-frame driver, same contract as mbo_render_frame: one dynamic queue over every
 * (chunk, row) item; used for the optional "vectorised CPU" baseline row of bench.py */
typedef struct {
    uint32_t level, ir0, ii0, blocks, d;
    int32_t mrd;
    int dtype;
    int32_t *out;
    atomic_ullong next, useful;
} simd_frame_job;

static void *simd_frame_worker(void *arg)
{
    simd_frame_job *j = (simd_frame_job *)arg;
    const uint32_t d = j->d;
    const uint64_t items = (uint64_t)j->blocks * j->blocks * d;
    double *r = (double *)malloc(sizeof(double) * d);
    double *im = (double *)malloc(sizeof(double) * d);
    int32_t *rowbuf = (int32_t *)malloc(sizeof(int32_t) * d);
    if (!r || !im || !rowbuf) { free(r); free(im); free(rowbuf); return NULL; }
    uint64_t have_chunk = UINT64_MAX, local = 0;
    const uint64_t never = j->mrd > 0 ? (uint64_t)(j->mrd - 1) : 0;
    for (;;) {
        const uint64_t it = atomic_fetch_add(&j->next, 1ull);
        if (it >= items) break;
        const uint64_t chunk = it / d;
        const uint32_t row = (uint32_t)(it % d);
        if (chunk != have_chunk) {
            const uint32_t bi = (uint32_t)(chunk / j->blocks), br = (uint32_t)(chunk % j->blocks);
            double range, sr, si;
            mbo_chunk_geometry(j->level, j->ir0 + br, j->ii0 + bi, &range, &sr, &si);
            volatile double stop_r = sr + range, stop_i = si + range;
            mbo_linspace(sr, stop_r, d, r);
            mbo_linspace(si, stop_i, d, im);
            have_chunk = chunk;
        }
        int32_t *o = j->out ? j->out + (size_t)it * d : rowbuf;
        if (j->dtype == 0) simd_row_f64(r, im[row], d, j->mrd, o);
        else simd_row_f32(r, im[row], d, j->mrd, o);
        for (uint32_t col = 0; col < d; ++col) local += o[col] ? (uint64_t)o[col] : never;
    }
    atomic_fetch_add(&j->useful, local);
    free(r); free(im); free(rowbuf);
    return NULL;
}

int mbo_simd_render_frame(uint32_t level, uint32_t ir0, uint32_t ii0, uint32_t blocks, uint32_t d,
                          int32_t mrd, int dtype, int threads, int32_t *out, uint64_t *useful)
{
    if (level == 0 || d == 0 || blocks == 0) return -1;
    if (dtype != 0 && dtype != 1) return -2;
    simd_frame_job job = {level, ir0, ii0, blocks, d, mrd, dtype, out, 0, 0};
    int nt = threads > 0 ? threads : mbo_max_threads();
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    if (th)
        for (; started < nt - 1; ++started)
            if (pthread_create(&th[started], NULL, simd_frame_worker, &job) != 0) break;
    simd_frame_worker(&job);
    for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
    free(th);
    if (useful) *useful = (uint64_t)atomic_load(&job.useful);
    return 0;
}
