/* div6_check.c -- test infrastructure (not shipped): x / 6 of the cubic B-spline weights (src/interp.jl:820-835, `/6`) is computed
 * on the GPU as  q = x * RN(1/6);  r = fma(-6, q, x);  result = fma(r, RN(1/6), q)  (grid.jl_b200/csrc/common.cuh: div6_exact).
 * Why that is the correctly rounded quotient for EVERY finite x away from the subnormals: x is a multiple of 4 or 8 ulp(q), so
 * x / 6 is a multiple of ulp(q) / 3 and at least ulp(q) / 6 away from every rounding midpoint; r is exact (a multiple of 2 ulp(q)
 * not above 6 ulp(q)); and q + r * RN(1/6) differs from x / 6 by at most 2^-53 ulp(q).  This program checks it against the
 * hardware division on structured and random numerators.  usage: div6_check [millions of random samples]; prints mismatches. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double div6(double x) {
    const double z = 1.0 / 6.0;
    const double q = x * z;
    const double r = fma(-6.0, q, x);
    return fma(r, z, q);
}
static long check(double x, long* n) {
    volatile double six = 6.0;  /* a real division, not a compiler's reciprocal trick */
    const double want = x / six, got = div6(x);
    (*n)++;
    return memcmp(&want, &got, 8) != 0;
}
int main(int argc, char** argv) {
    const long millions = argc > 1 ? atol(argv[1]) : 50;
    long bad = 0, n = 0;
    /* structured: every power of two in the weights' range and its neighbours, multiples of 3 and 6 around them, [0, 1] lattice */
    for (int e = -200; e <= 60; e++) {
        const double p = ldexp(1.0, e);
        double x = p;
        for (int k = 0; k < 64; k++) { bad += check(x, &n); bad += check(3 * x, &n); bad += check(6 * x, &n); bad += check(x / 3, &n); x = nextafter(x, INFINITY); }
        x = p;
        for (int k = 0; k < 64; k++) { bad += check(x, &n); bad += check(3 * x, &n); x = nextafter(x, 0.0); }
    }
    for (int k = 0; k <= 1 << 20; k++) { const double t = (double)k / (1 << 20); bad += check(t * t * t, &n); bad += check((t * t) * (1 - t), &n); }
    bad += check(0.0, &n);
    /* random mantissas over the exponents the weights can take (t^3 and dx^2 dx with t, dx in [2^-52, 1]) and beyond */
    for (long i = 0; i < millions * 1000000L; i++) {
        const uint64_t m = (rnd() >> 12) | ((uint64_t)(1023 - 170 + (int)(rnd() % 200)) << 52);
        double x;
        memcpy(&x, &m, 8);
        bad += check(x, &n);
    }
    printf("%ld samples, %ld mismatches\n", n, bad);
    return bad != 0;
}
