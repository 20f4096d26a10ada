/*
 * mandel_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see mandel_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -lpthread (oracle/Makefile).
 * -ffp-contract=off is load-bearing: an FMA-contracted build changes 59 of 262,144
 * counts on one 512^2 boundary tile (SURVEY.md F2 / Appendix C last row).
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/.  ".py" = DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py.
 */
#include "mandel_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <unistd.h>

#if defined(__FMA__) || defined(__FP_FAST_FMA)
/* harmless with -ffp-contract=off, but make sure nobody builds this -march=native -ffast-math */
#endif
#ifdef __FAST_MATH__
#error "the oracle must not be built with -ffast-math"
#endif

/* .py:7-8 MIN_AXIS/MAX_AXIS ; DataChunk.cs:14-15 minAxis/maxAxis */
static const double MBO_MIN_AXIS = -2.0;
static const double MBO_MAX_AXIS = 2.0;

/* .py:75-78: chunk_range = (MAX_AXIS - MIN_AXIS) / level ; start = MIN_AXIS + (chunk_range * index)
 * DataChunk.cs:32-33 GetLevelChunkRange ; :59-66 StartValueReal/Imag.  Two roundings per start. */
void mbo_chunk_geometry(uint32_t level, uint32_t index_real, uint32_t index_imag,
                        double *range, double *start_r, double *start_i)
{
    volatile double rg = (MBO_MAX_AXIS - MBO_MIN_AXIS) / (double)level;
    volatile double pr = rg * (double)index_real;
    volatile double pi = rg * (double)index_imag;
    *range = rg;
    *start_r = MBO_MIN_AXIS + pr;
    *start_i = MBO_MIN_AXIS + pi;
}

/* numpy.linspace as called at .py:24-32 (endpoint=True, num=d):
 *   delta = stop - start ; step = delta / (d-1) ; y[k] = k*step + start ; y[d-1] = stop
 * (numpy/_core/function_base.py: `y = arange(0,num) ; y = y*step ; y += start ; y[-1] = stop`;
 *  the step==0 branch cannot trigger for range > 0).  d == 1 -> [start] (div = 0 path). */
void mbo_linspace(double start, double stop, uint32_t d, double *out)
{
    if (d == 0) return;
    if (d == 1) { out[0] = start; return; }
    volatile double delta = stop - start;
    volatile double step = delta / (double)(d - 1);
    for (uint32_t k = 0; k < d; ++k) {
        volatile double p = (double)k * step;
        out[k] = p + start;
    }
    out[d - 1] = stop;
}

/* .py:39-68 calc_mb_value, written the way the reference writes it.
 *   c = (r,i); z = (r,i)                       :44-45
 *   for n in range(1, mrd):                    :47
 *     z = (z0*z0 - z1*z1, 2*z0*z1)             :50-53   (2*z0*z1 parses as (2*z0)*z1)
 *     z = (z0 + c0, z1 + c1)                   :56-59
 *     sqr_mag = z0*z0 + z1*z1                  :62
 *     if sqr_mag >= 4: return n                :65-66
 *   return 0                                   :68
 */
int32_t mbo_escape_f64(double r, double i, int32_t mrd)
{
    const double c0 = r, c1 = i;
    double z0 = r, z1 = i;
    for (int32_t n = 1; n < mrd; ++n) {
        double a = z0 * z0;
        double b = z1 * z1;
        double sq0 = a - b;
        double t = 2.0 * z0;
        double sq1 = t * z1;
        z0 = sq0 + c0;
        z1 = sq1 + c1;
        double m0 = z0 * z0;
        double m1 = z1 * z1;
        double sqr_mag = m0 + m1;
        if (sqr_mag >= 4.0) return n;
    }
    return 0;
}

/* binary32 variant: same operation order, float arithmetic, threshold 4.0f (Appendix A.5). */
int32_t mbo_escape_f32(float r, float i, int32_t mrd)
{
    const float c0 = r, c1 = i;
    float z0 = r, z1 = i;
    for (int32_t n = 1; n < mrd; ++n) {
        float a = z0 * z0;
        float b = z1 * z1;
        float sq0 = a - b;
        float t = 2.0f * z0;
        float sq1 = t * z1;
        z0 = sq0 + c0;
        z1 = sq1 + c1;
        float m0 = z0 * z0;
        float m1 = z1 * z1;
        float sqr_mag = m0 + m1;
        if (sqr_mag >= 4.0f) return n;
    }
    return 0;
}

/* .py:96-98: out = (out.astype(float64) * 256) / mrd ; ceil(out).astype(uint8).
 * numpy's float64 -> uint8 cast of the value 256.0 wraps to 0 on x86-64 (goes through a
 * wider integer then truncates); reproduce that with an explicit int64 -> uint8 truncation. */
uint8_t mbo_count_to_u8(int32_t n, int32_t mrd)
{
    volatile double v = (double)n * 256.0;
    volatile double q = v / (double)mrd;
    double c = ceil(q);
    return (uint8_t)(int64_t)c;
}

void mbo_counts_to_u8(const int32_t *counts, uint8_t *out, size_t n, int32_t mrd)
{
    for (size_t k = 0; k < n; ++k) out[k] = mbo_count_to_u8(counts[k], mrd);
}

int mbo_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

/* dynamic row queue shared by the worker threads of one mbo_render_rect call */
typedef struct {
    const double *r, *im;
    int32_t *out;
    size_t pitch;
    uint32_t x0, y0, w, h;
    int32_t mrd;
    int dtype;
    atomic_uint next_row;
} mbo_job;

static void *mbo_worker(void *arg)
{
    mbo_job *j = (mbo_job *)arg;
    for (;;) {
        uint32_t row = atomic_fetch_add(&j->next_row, 1u);
        if (row >= j->h) break;
        int32_t *o = j->out + (size_t)row * j->pitch;
        const double ci = j->im[j->y0 + row];
        if (j->dtype == 0) {
            for (uint32_t col = 0; col < j->w; ++col)
                o[col] = mbo_escape_f64(j->r[j->x0 + col], ci, j->mrd);
        } else {
            /* Appendix A.5: each binary64 coordinate rounded once to binary32 */
            const float cif = (float)ci;
            for (uint32_t col = 0; col < j->w; ++col)
                o[col] = mbo_escape_f32((float)j->r[j->x0 + col], cif, j->mrd);
        }
    }
    return NULL;
}

/* .py:19-37 gen_arrays + .py:92 kernel call, restricted to a sub-rectangle of the chunk.
 * The coordinates always come from the CHUNK-level linspace (never re-linspaced per tile):
 * pixel p = row*d + col -> (r[col], i[row])   (np.tile(r, d) / np.repeat(i, d), .py:34-35). */
int mbo_render_rect(double start_r, double start_i, double range, uint32_t d, int32_t mrd,
                    uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                    int32_t *out, size_t pitch, int dtype, int threads)
{
    if (d == 0 || (uint64_t)x0 + w > d || (uint64_t)y0 + h > d) return -1;
    if (dtype != 0 && dtype != 1) return -2;
    double *r = (double *)malloc(sizeof(double) * d);
    double *im = (double *)malloc(sizeof(double) * d);
    if (!r || !im) { free(r); free(im); return -3; }
    /* .py:24-32: stop = start + _range */
    volatile double stop_r = start_r + range;
    volatile double stop_i = start_i + range;
    mbo_linspace(start_r, stop_r, d, r);
    mbo_linspace(start_i, stop_i, d, im);
    mbo_job job = {r, im, out, pitch, x0, y0, w, h, mrd, dtype, 0};
    int nt = threads > 0 ? threads : mbo_max_threads();
    if ((uint32_t)nt > h) nt = (int)h;
    if (nt <= 1) {
        mbo_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
        int started = 0;
        if (th)
            for (; started < nt - 1; ++started)
                if (pthread_create(&th[started], NULL, mbo_worker, &job) != 0) break;
        mbo_worker(&job); /* the calling thread works too */
        for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
        free(th);
    }
    free(r);
    free(im);
    return 0;
}

/* .py:70-100 process_workload with `definition` lifted to a parameter (the reference fixes 4096, .py:80) */
int mbo_process_workload(uint32_t level, int32_t mrd, uint32_t index_real, uint32_t index_imag,
                         uint32_t d, uint8_t *out_u8, int32_t *out_counts, int threads)
{
    if (level == 0 || d == 0) return -1;
    double range, sr, si;
    mbo_chunk_geometry(level, index_real, index_imag, &range, &sr, &si);
    size_t n = (size_t)d * d;
    int32_t *counts = out_counts ? out_counts : (int32_t *)malloc(sizeof(int32_t) * n);
    if (!counts) return -3;
    int rc = mbo_render_rect(sr, si, range, d, mrd, 0, 0, d, d, counts, d, 0, threads);
    if (rc == 0 && out_u8) mbo_counts_to_u8(counts, out_u8, n, mrd);
    if (!out_counts) free(counts);
    return rc;
}

/* ---- whole-frame driver for the CPU baseline: ONE dynamic queue over every (chunk, row) of a
 * blocks x blocks frame, so that all host cores stay busy (per-chunk calls leave cores idle at
 * small definitions).  Same per-pixel code path as mbo_render_rect. ---------------------------- */
typedef struct {
    uint32_t level, ir0, ii0, blocks, d;
    int32_t mrd;
    int dtype;
    int32_t *out;                 /* may be NULL: only the useful-iteration sum is wanted */
    atomic_ullong next;           /* next (chunk * d + row) item */
    atomic_ullong useful;
} mbo_frame_job;

static void *mbo_frame_worker(void *arg)
{
    mbo_frame_job *j = (mbo_frame_job *)arg;
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
            /* chunk order as in the product's frame layout: chunk = bi * blocks + br */
            const uint32_t bi = (uint32_t)(chunk / j->blocks), br = (uint32_t)(chunk % j->blocks);
            double range, sr, si;
            mbo_chunk_geometry(j->level, j->ir0 + br, j->ii0 + bi, &range, &sr, &si);
            volatile double stop_r = sr + range, stop_i = si + range;
            mbo_linspace(sr, stop_r, d, r);
            mbo_linspace(si, stop_i, d, im);
            have_chunk = chunk;
        }
        int32_t *o = j->out ? j->out + (size_t)it * d : rowbuf;
        const double ci = im[row];
        if (j->dtype == 0) {
            for (uint32_t col = 0; col < d; ++col) o[col] = mbo_escape_f64(r[col], ci, j->mrd);
        } else {
            const float cif = (float)ci;
            for (uint32_t col = 0; col < d; ++col) o[col] = mbo_escape_f32((float)r[col], cif, j->mrd);
        }
        for (uint32_t col = 0; col < d; ++col) local += o[col] ? (uint64_t)o[col] : never;
    }
    atomic_fetch_add(&j->useful, local);
    free(r); free(im); free(rowbuf);
    return NULL;
}

int mbo_render_frame(uint32_t level, uint32_t ir0, uint32_t ii0, uint32_t blocks, uint32_t d,
                     int32_t mrd, int dtype, int threads, int32_t *out, uint64_t *useful)
{
    if (level == 0 || d == 0 || blocks == 0) return -1;
    if (dtype != 0 && dtype != 1) return -2;
    mbo_frame_job job = {level, ir0, ii0, blocks, d, mrd, dtype, out, 0, 0};
    int nt = threads > 0 ? threads : mbo_max_threads();
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    if (th)
        for (; started < nt - 1; ++started)
            if (pthread_create(&th[started], NULL, mbo_frame_worker, &job) != 0) break;
    mbo_frame_worker(&job);
    for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
    free(th);
    if (useful) *useful = (uint64_t)atomic_load(&job.useful);
    return 0;
}

/* Empirical check of the property the CUDA kernel's deferred escape test relies on
 * (csrc/escape_kernels.cuh): once the computed |z|^2 reaches 4, it stays >= 4 and finite for the
 * next `extra` updates.  Counts the points that violate it.  Points with | |c|^2 - 4 | <= band are
 * skipped (the kernel never defers those).  Strict arithmetic as everywhere in this file. */
uint64_t mbo_sticky_violations(const double *r, const double *im, size_t n, int32_t mrd, int extra, double band)
{
    uint64_t bad = 0;
    for (size_t k = 0; k < n; ++k) {
        const double c0 = r[k], c1 = im[k];
        volatile double q0 = c0 * c0, q1 = c1 * c1;
        volatile double q = (q0 + q1) - 4.0;
        if (q <= band && q >= -band) continue;
        double z0 = c0, z1 = c1;
        int left = -1;                      /* updates still to verify after the first escape */
        for (int32_t it = 1; it < mrd; ++it) {
            double a = z0 * z0, b = z1 * z1;
            double sq0 = a - b;
            double t = 2.0 * z0;
            double sq1 = t * z1;
            z0 = sq0 + c0;
            z1 = sq1 + c1;
            double m0 = z0 * z0, m1 = z1 * z1;
            double m = m0 + m1;
            if (left < 0) {
                if (m >= 4.0) left = extra;
            } else {
                if (!(m >= 4.0) || !(m < 1.7976931348623157e308)) { ++bad; break; }
            }
            if (left == 0) break;
            if (left > 0) --left;
        }
    }
    return bad;
}

/* binary32 counterpart: the kernel's block-end test there is `!(m < 4)` (overflow to inf/NaN counts
 * as escaped), blocks of 16, band 2^-9 on |c|^2. */
uint64_t mbo_sticky_violations_f32(const double *r, const double *im, size_t n, int32_t mrd, int extra, float band)
{
    uint64_t bad = 0;
    for (size_t k = 0; k < n; ++k) {
        const float c0 = (float)r[k], c1 = (float)im[k];
        volatile float q0 = c0 * c0, q1 = c1 * c1;
        volatile float q = (q0 + q1) - 4.0f;
        if (q <= band && q >= -band) continue;
        float z0 = c0, z1 = c1;
        int left = -1;
        for (int32_t it = 1; it < mrd; ++it) {
            float a = z0 * z0, b = z1 * z1;
            float sq0 = a - b;
            float t = 2.0f * z0;
            float sq1 = t * z1;
            z0 = sq0 + c0;
            z1 = sq1 + c1;
            float m0 = z0 * z0, m1 = z1 * z1;
            float m = m0 + m1;
            if (left < 0) {
                if (m >= 4.0f) left = extra;
            } else if (m < 4.0f) { ++bad; break; }
            if (left == 0) break;
            if (left > 0) --left;
        }
    }
    return bad;
}

uint64_t mbo_useful_iterations(const int32_t *counts, size_t n, int32_t mrd)
{
    uint64_t s = 0;
    const uint64_t never = mrd > 0 ? (uint64_t)(mrd - 1) : 0;
    for (size_t k = 0; k < n; ++k) s += counts[k] ? (uint64_t)counts[k] : never;
    return s;
}

uint64_t mbo_fnv1a_words(const int32_t *counts, size_t n)
{
    uint64_t hsh = 0xcbf29ce484222325ULL;
    for (size_t k = 0; k < n; ++k) {
        hsh ^= (uint64_t)(uint32_t)counts[k];
        hsh *= 0x100000001b3ULL;
    }
    return hsh;
}
