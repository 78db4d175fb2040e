/* check_exact_div.c -- test infrastructure (not shipped): checks that the FMA-refined reciprocal
 * division used by the CUDA prefilter (grid.jl_b200/csrc/prefilter.cu, exact_div) returns the same
 * bits as the hardware division a / d for the pivots of the five prefilter systems
 * (src/interp.jl:955-1014) and for random divisors, over ~5e8 random numerators each for double and
 * float.  Build and run:  gcc -O2 -march=native -ffp-contract=off oracle/check_exact_div.c -lm && ./a.out
 * Expected output: zero mismatches for the two-iteration form (the one the kernel uses). */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double div1(double a, double b, double y) { double q0 = a * y; double r = fma(-b, q0, a); return fma(r, y, q0); }
static inline double div2(double a, double b, double y) { double q0 = a * y; double r = fma(-b, q0, a); double q1 = fma(r, y, q0); double r1 = fma(-b, q1, a); return fma(r1, y, q1); }
static inline float div1f(float a, float b, float y) { float q0 = a * y; float r = fmaf(-b, q0, a); return fmaf(r, y, q0); }
static inline float div2f(float a, float b, float y) { float q0 = a * y; float r = fmaf(-b, q0, a); float q1 = fmaf(r, y, q0); float r1 = fmaf(-b, q1, a); return fmaf(r1, y, q1); }
int main() {
    // divisors: LU pivots of tridiag(1/8,3/4,1/8), tridiag(1/6,4/6,1/6) variants + random ones
    double divs[4096]; int nd = 0;
    { double d = 0.75; for (int i = 0; i < 60; i++) { divs[nd++] = d; double f = 0.125 / d; d = 0.75 - f * 0.125; } }
    { double d = 9.0/8; for (int i = 0; i < 60; i++) { divs[nd++] = d; double f = 0.125 / d; d = 0.75 - f * (i==0 ? -0.25 : 0.125); } }
    { double d = 7.0/8; for (int i = 0; i < 60; i++) { divs[nd++] = d; double f = 0.125 / d; d = 0.75 - f * 0.125; } }
    { double d = 4.0/6; for (int i = 0; i < 60; i++) { divs[nd++] = d; double f = (1.0/6) / d; d = 4.0/6 - f * (1.0/6); } }
    divs[nd++] = 1.0; divs[nd++] = -2.0; divs[nd++] = 6.0; divs[nd++] = 3.0; divs[nd++] = 1.0/6;
    for (int i = 0; i < 1000; i++) { uint64_t b = (rnd() >> 12) | 0x3ff0000000000000ull; double d; memcpy(&d, &b, 8); divs[nd++] = d; }
    long bad1 = 0, bad2 = 0, tot = 0;
    for (int k = 0; k < nd; k++) {
        double b = divs[k], y = 1.0 / b;
        for (int i = 0; i < 400000; i++) {
            uint64_t m = (rnd() >> 12) | ((uint64_t)(1023 - 30 + (rnd() % 60)) << 52) | ((rnd() & 1) << 63);
            double a; memcpy(&a, &m, 8);
            double q = a / b;
            if (div1(a, b, y) != q) bad1++;
            if (div2(a, b, y) != q) bad2++;
            tot++;
        }
    }
    printf("f64: %ld samples, one-iteration mismatches %ld, two-iteration mismatches %ld\n", tot, bad1, bad2);
    bad1 = bad2 = tot = 0;
    for (int k = 0; k < nd; k++) {
        float b = (float)divs[k], y = 1.0f / b;
        for (int i = 0; i < 400000; i++) {
            uint32_t m = (uint32_t)(rnd() >> 41) | ((uint32_t)(127 - 20 + (rnd() % 40)) << 23) | ((uint32_t)(rnd() & 1) << 31);
            float a; memcpy(&a, &m, 4);
            float q = a / b;
            if (div1f(a, b, y) != q) bad1++;
            if (div2f(a, b, y) != q) bad2++;
            tot++;
        }
    }
    printf("f32: %ld samples, one-iteration mismatches %ld, two-iteration mismatches %ld\n", tot, bad1, bad2);
    return 0;
}
