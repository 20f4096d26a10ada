// threshold_points.c -- TEST INFRASTRUCTURE (generator of tests/golden/threshold_points.json).
// The kernels decide "|z|^2 >= 4" with an INTEGER compare of the high word of m (csrc/escape_kernels.cuh, Ops::escaped /
// Defer::tripped), once per block of 8 untested updates in the steady state.  This program searches, with the strict
// arithmetic of .py:44-68, for coordinates whose FIRST escaping update n is a multiple of 8 and whose m there has the
// high word 0x40100000 exactly (4 <= m < 4 + 2^-18): inputs on which that compare is decided by the last unit of
// the high word, at a block end of both kernel forms (the patch kernel's blocks end at 40, 48, ...; the refill
// kernel's at 8, 16, ...).  About one random boundary point in 4e7 qualifies.
// binary32 (argument 7 = f32): blocks of 16 there and the block-end test is `!(m < 4)`, so the points wanted are those
// whose first escaping update is a multiple of 16 with m == 4.0f exactly.
//     gcc -O2 -ffp-contract=off -fopenmp -o /tmp/threshold_points oracle/threshold_points.c
//     /tmp/threshold_points -0.79 -0.71 0.06 0.14 150000000 300 f64     # r0 r1 i0 i1 points-per-thread mrd type -> "cr ci n m" lines (hex floats)
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <omp.h>
static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static inline double unit(uint64_t *s) { return (rng(s) >> 11) * (1.0 / 9007199254740992.0); }

static void search_f64(double r0, double r1, double i0, double i1, long per, int mrd, uint64_t seed)
{
    uint64_t s = seed;
    for (long it = 0; it < per; ++it) {
        double cr = r0 + (r1 - r0) * unit(&s), ci = i0 + (i1 - i0) * unit(&s);
        double x = cr, y = ci;
        for (int n = 1; n < mrd; ++n) {
            double a = (x * x) - (y * y);
            double b = ((2 * x) * y);
            x = a + cr; y = b + ci;
            double m = (x * x) + (y * y);
            if (m >= 4) {
                uint64_t bits; memcpy(&bits, &m, 8);
                if ((uint32_t)(bits >> 32) == 0x40100000u && n >= 8 && n % 8 == 0) {
                    #pragma omp critical
                    { printf("%a %a %d %.17g\n", cr, ci, n, m); fflush(stdout); }
                }
                break;
            }
        }
    }
}

static void search_f32(double r0, double r1, double i0, double i1, long per, int mrd, uint64_t seed)
{
    uint64_t s = seed;
    for (long it = 0; it < per; ++it) {
        float cr = (float)(r0 + (r1 - r0) * unit(&s)), ci = (float)(i0 + (i1 - i0) * unit(&s));
        float x = cr, y = ci;
        for (int n = 1; n < mrd; ++n) {
            float a = (x * x) - (y * y);
            float b = ((2 * x) * y);
            x = a + cr; y = b + ci;
            float m = (x * x) + (y * y);
            if (m >= 4) {
                if (m == 4.0f && n >= 16 && n % 16 == 0) {
                    #pragma omp critical
                    { printf("%a %a %d %.9g\n", (double)cr, (double)ci, n, (double)m); fflush(stdout); }
                }
                break;
            }
        }
    }
}

int main(int argc, char **argv) {
    if (argc < 8) { fprintf(stderr, "usage: %s r0 r1 i0 i1 points-per-thread mrd f64|f32\n", argv[0]); return 2; }
    const double r0 = atof(argv[1]), r1 = atof(argv[2]), i0 = atof(argv[3]), i1 = atof(argv[4]);
    const long per = atol(argv[5]);
    const int mrd = atoi(argv[6]), f32 = strcmp(argv[7], "f32") == 0;
    #pragma omp parallel
    {
        const uint64_t seed = 0x9E3779B97F4A7C15ull ^ (uint64_t)(omp_get_thread_num() + 1) * 0xD1B54A32D192ED03ull;
        if (f32) search_f32(r0, r1, i0, i1, per, mrd, seed);
        else search_f64(r0, r1, i0, i1, per, mrd, seed);
    }
    return 0;
}
