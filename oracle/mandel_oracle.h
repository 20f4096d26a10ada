/*
 * mandel_oracle.h -- CPU oracle for the DistributedMandelbrot escape-time hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in distributedmandelbrot_b200/ (the product) may
 * include, link, import or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
 * (or as the timed CPU baseline), never as the thing shipped.
 *
 * It is a plain-C restatement of the reference worker
 *   DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py:7-8,19-37,39-68,70-100
 * cross-checked against the same geometry formulas in
 *   DistributedMandelbrot/DataChunk.cs:14-15,20,27,32-33,59-66.
 *
 * Evaluation rule: IEEE-754 binary64, round-to-nearest-even, every * + - /
 * individually rounded, NO contraction (build with -ffp-contract=off, no -ffast-math),
 * operations in the order Python parses them.
 *
 * PARITY PINNING: the reference has no tests, fixtures or golden vectors of its own
 * (SURVEY.md F3).  The oracle is pinned instead against outputs of the reference's own
 * function text (.py:19-37, :39-68, :96-98) executed in the authoring container under
 * numba's CPU target / numpy -- see oracle/gen_golden.py and tests/golden/.
 */
#ifndef MANDEL_ORACLE_H
#define MANDEL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* .py:75-78 == DataChunk.cs:32-33,59-66 */
void mbo_chunk_geometry(uint32_t level, uint32_t index_real, uint32_t index_imag,
                        double *range, double *start_r, double *start_i);

/* numpy.linspace(start, stop, num=d, endpoint=True) as called at .py:24-32 */
void mbo_linspace(double start, double stop, uint32_t d, double *out);

/* .py:39-68, literally (10 arithmetic ops + compare per iteration) */
int32_t mbo_escape_f64(double r, double i, int32_t mrd);

/* binary32 variant (no upstream counterpart) */
int32_t mbo_escape_f32(float r, float i, int32_t mrd);

/* .py:96-98 for one element, via the same float64 path numpy takes */
uint8_t mbo_count_to_u8(int32_t n, int32_t mrd);

/*
 * Render the sub-rectangle [x0,x0+w) x [y0,y0+h) of the chunk whose linspace is
 * (start_r, start_r+range, d) x (start_i, start_i+range, d).  Pixel (row, col) -> c =
 * (r[col], i[row]) (.py:34-35).  out[(row-y0)*pitch + (col-x0)].  dtype: 0 = f64, 1 = f32.
 * threads <= 0 -> all online cores (pthreads).
 */
int mbo_render_rect(double start_r, double start_i, double range, uint32_t d, int32_t mrd,
                    uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                    int32_t *out, size_t pitch, int dtype, int threads);

/* whole process_workload (.py:70-100) with `definition` as a parameter: uint8[d*d] */
int mbo_process_workload(uint32_t level, int32_t mrd, uint32_t index_real, uint32_t index_imag,
                         uint32_t d, uint8_t *out_u8, int32_t *out_counts /* may be NULL */,
                         int threads);

/*
 * A blocks x blocks frame of chunks (ir0+br, ii0+bi) at `level`, definition d, threaded over all
 * (chunk, row) items with one dynamic queue.  out (may be NULL) is chunk-contiguous
 * [bi*blocks+br][row][col]; *useful receives the sum of (n ? n : mrd-1).  Used by the CPU baseline.
 */
int mbo_render_frame(uint32_t level, uint32_t ir0, uint32_t ii0, uint32_t blocks, uint32_t d,
                     int32_t mrd, int dtype, int threads, int32_t *out, uint64_t *useful);

/* number of points whose computed |z|^2 drops below 4 (or overflows) within `extra` updates after
 * first reaching 4 -- the property behind the kernel's deferred escape test; expected 0 */
uint64_t mbo_sticky_violations(const double *r, const double *im, size_t n, int32_t mrd, int extra, double band);

uint64_t mbo_sticky_violations_f32(const double *r, const double *im, size_t n, int32_t mrd, int extra, float band);

void mbo_counts_to_u8(const int32_t *counts, uint8_t *out, size_t n, int32_t mrd);

/* sum over pixels of (n ? n : mrd-1): the unit of the Gpixel-iterations/s metric */
uint64_t mbo_useful_iterations(const int32_t *counts, size_t n, int32_t mrd);

/* 64-bit FNV-1a variant over counts taken as uint32 words */
uint64_t mbo_fnv1a_words(const int32_t *counts, size_t n);

int mbo_max_threads(void);

/*
 * Data-parallel build of the same restatement (mandel_oracle_simd.c): 8 / 16 pixels of a row per
 * vector, lane-wise IEEE operations, no contraction -- bit-identical to mbo_render_rect /
 * mbo_render_frame by construction and by test (tests/test_oracle_golden.py).  Fast enough to
 * render whole native chunks of the BASELINE frames (oracle/gen_fullsize_golden.py).
 */
int mbo_simd_render_rect(double start_r, double start_i, double range, uint32_t d, int32_t mrd,
                         uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                         int32_t *out, size_t pitch, int dtype, int threads);
int mbo_simd_render_frame(uint32_t level, uint32_t ir0, uint32_t ii0, uint32_t blocks, uint32_t d,
                          int32_t mrd, int dtype, int threads, int32_t *out, uint64_t *useful);
/* vector width (bits) of the body the loader selected for this CPU: 512, 256 or 128 */
int mbo_simd_width(void);

#ifdef __cplusplus
}
#endif
#endif
