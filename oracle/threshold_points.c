// threshold_points.c -- TEST INFRASTRUCTURE (generator of tests/golden/threshold_points.json).
// The kernels decide "|z|^2 >= 4" with an INTEGER compare of the high word of m (csrc/escape_kernels.cuh, Ops::escaped /
// Defer::tripped), once per block of 8 untested updates in the steady state.  This program searches, with the strict
// arithmetic of .py:44-68, for coordinates whose FIRST escaping update n is a multiple of 8 and whose m there has the
// high word 0x40100000 exactly (4 <= m < 4 + 2^-18): inputs on which that compare is decided by the last unit of
// the high word, at a block end of both kernel forms (the patch kernel's blocks end at 40, 48, ...; the refill
// kernel's at 8, 16, ...).  About one random boundary point in 4e7 qualifies.
//     gcc -O2 -ffp-contract=off -fopenmp -o /tmp/threshold_points oracle/threshold_points.c
//     /tmp/threshold_points -0.79 -0.71 0.06 0.14 150000000 300      # r0 r1 i0 i1 points-per-thread mrd -> "cr ci n m" lines (hex floats)
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <omp.h>
static inline uint64_t rng(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
int main(int argc, char **argv) {
    double r0 = atof(argv[1]), r1 = atof(argv[2]), i0 = atof(argv[3]), i1 = atof(argv[4]);
    long per = atol(argv[5]);
    int mrd = atoi(argv[6]);
    #pragma omp parallel
    {
        uint64_t s = 0x9E3779B97F4A7C15ull ^ (uint64_t)(omp_get_thread_num() + 1) * 0xD1B54A32D192ED03ull;
        for (long it = 0; it < per; ++it) {
            double cr = r0 + (r1 - r0) * ((rng(&s) >> 11) * (1.0 / 9007199254740992.0));
            double ci = i0 + (i1 - i0) * ((rng(&s) >> 11) * (1.0 / 9007199254740992.0));
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
    return 0;
}
