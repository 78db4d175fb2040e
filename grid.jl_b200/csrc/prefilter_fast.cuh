// prefilter_fast.cuh -- the arithmetic of the GB_FAST prefilter (interp_invert!, src/interp.jl:909-1014), shared by
// the kernels of prefilter_fast.cu and by the host-side algorithm check tests/native/fast_prefilter_check.cu.
//
// The reference solves M c = f along every 1-D line with a sequential tridiagonal LU (+ Woodbury corners).  That
// order leaves grids with FEW LONG lines (an 8192 x 8192 image: 8192 lines) with 8192 threads on a 300k-thread
// machine.  The B-spline systems are strongly diagonally dominant: M is Toeplitz tridiag(a, b, a) except for its
// first / last rows, and its LU factors converge to the stationary pair  y[i] = f[i] - l y[i-1],
// x[i] = y[i]/d - l x[i+1]  with l = a/d, d = (b + sqrt(b^2 - 4a^2))/2, |l| = 0.172 (quadratic) / 0.268 (cubic):
// whatever happens K rows away is damped by |l|^K (K = 32: < 6e-19).  So
//   * INTERIOR rows (>= K from both ends) of a line can be solved in independent SEGMENTS: a segment starts its
//     forward recurrence K rows early from a zero state and runs its backward recurrence from a zero state K rows
//     late -- no communication between segments, each element read once per segment (+ the 2K overlap);
//   * the K rows at each END, where the special boundary rows / Woodbury corners of the reference's matrices act
//     (src/interp.jl:955-1014), are a small dense product with the exact inverse of the matrix's leading / trailing
//     block, computed once on the host in long double from the matrix AS THE REFERENCE BUILDS IT (entries rounded
//     to T);
//   * BCperiodic has no ends: the line is a ring and the overlaps wrap around.
// The result agrees with the exact solve to rounding (a few ulp of max|f|; tests: <= 4 ulp f64, 1e-5 relative f32).
#pragma once
#include <cmath>
#include <vector>

#include "common.cuh"

#ifndef GB_FAST_FN  // the host-side algorithm check compiles the same functions for the CPU
#define GB_FAST_FN __device__ __forceinline__
#endif

namespace gb {

constexpr int FAST_K = 32;              // overlap rows = dense end rows
constexpr int FAST_COLS = 2 * FAST_K;   // columns of an end block
constexpr int FAST_M = 3 * FAST_K;      // leading / trailing block inverted on the host
constexpr i64 FAST_MIN_N = 8 * FAST_K;  // shorter lines take the exact kernels

template <typename T> struct FastSys {
    T l, rd;             // y = f - l*y_prev;  x = y*rd - l*x_next
    int periodic;        // ring: no dense ends
    const double* wtop;  // device [FAST_K][FAST_COLS]: x[i] = sum_j wtop[i][j] f[j]
    const double* wbot;  // device [FAST_K][FAST_COLS]: x[n-K+i] = sum_j wbot[i][j] f[n-COLS+j]
};

// entry (i, j) of the reference's system matrix (_interp_invert_matrix, src/interp.jl:955-1014), entries rounded to T
// exactly where the reference rounds them (convert(T, 1/8) etc.); fam = BC family
template <typename T> inline long double fast_matrix_entry(int fam, int order, i64 n, i64 i, i64 j) {
    const i64 k = j - i;
    if (order == GB_QUADRATIC) {
        const T off = (T)(1.0 / 8), dia = (T)(3.0 / 4);
        if (fam == BCF_NIL) {  // rows 1 / n: [9/8, -1/4, 1/8] and mirrored
            if (i == 0) return j == 0 ? (long double)(T)(9.0 / 8) : j == 1 ? (long double)(T)(-1.0 / 4) : j == 2 ? (long double)(T)(1.0 / 8) : 0.0L;
            if (i == n - 1) return j == n - 1 ? (long double)(T)(9.0 / 8) : j == n - 2 ? (long double)(T)(-1.0 / 4) : j == n - 3 ? (long double)(T)(1.0 / 8) : 0.0L;
        } else if (fam == BCF_REFLECT) {
            if ((i == 0 && j == 0) || (i == n - 1 && j == n - 1)) return (long double)(T)((double)dia + 1.0 / 8);
        } else if (fam == BCF_PERIODIC) {
            if ((i == 0 && j == n - 1) || (i == n - 1 && j == 0)) return (long double)off;
        } else {  // nearest / fill
            if ((i == 0 && j == 1) || (i == n - 1 && j == n - 2)) return (long double)(T)((double)off + 1.0 / 8);
        }
        return k == 0 ? (long double)dia : (k == 1 || k == -1) ? (long double)off : 0.0L;
    }
    // cubic (nil / nan / na): rows [1,-2,1], [0,1,0], ..., [0,1,0], [1,-2,1]
    const T off = (T)(1.0 / 6), dia = (T)(4.0 / 6);
    if (i == 0) return j == 0 ? 1.0L : j == 1 ? -2.0L : j == 2 ? 1.0L : 0.0L;
    if (i == n - 1) return j == n - 1 ? 1.0L : j == n - 2 ? -2.0L : j == n - 3 ? 1.0L : 0.0L;
    if (i == 1 || i == n - 2) return k == 0 ? 1.0L : 0.0L;
    return k == 0 ? (long double)dia : (k == 1 || k == -1) ? (long double)off : 0.0L;
}

// Host set-up: stationary recurrence constants and the two end blocks (row-major [K][COLS], double).
template <typename T> struct FastHost {
    T l, rd;
    int periodic;
    std::vector<double> wtop, wbot;
    void setup(int fam, int order, i64 n) {
        const long double a = order == GB_QUADRATIC ? (long double)(T)(1.0 / 8) : (long double)(T)(1.0 / 6);
        const long double b = order == GB_QUADRATIC ? (long double)(T)(3.0 / 4) : (long double)(T)(4.0 / 6);
        const long double d = (b + sqrtl(b * b - 4 * a * a)) / 2;
        l = (T)(a / d);
        rd = (T)(1.0L / d);
        periodic = fam == BCF_PERIODIC ? 1 : 0;
        wtop.assign((size_t)FAST_K * FAST_COLS, 0.0);
        wbot.assign((size_t)FAST_K * FAST_COLS, 0.0);
        if (periodic) return;
        const int m = FAST_M;
        for (int end = 0; end < 2; end++) {
            // leading (end 0) / trailing (end 1) m x m block of the matrix, inverted by Gauss-Jordan with partial pivoting
            std::vector<long double> A((size_t)m * 2 * m, 0.0L);
            const i64 o = end ? n - m : 0;
            for (int i = 0; i < m; i++) {
                for (int j = 0; j < m; j++) A[(size_t)i * 2 * m + j] = fast_matrix_entry<T>(fam, order, n, o + i, o + j);
                A[(size_t)i * 2 * m + m + i] = 1.0L;
            }
            for (int c = 0; c < m; c++) {
                int p = c;
                for (int r = c + 1; r < m; r++)
                    if (fabsl(A[(size_t)r * 2 * m + c]) > fabsl(A[(size_t)p * 2 * m + c])) p = r;
                if (p != c)
                    for (int j = 0; j < 2 * m; j++) std::swap(A[(size_t)p * 2 * m + j], A[(size_t)c * 2 * m + j]);
                const long double piv = 1.0L / A[(size_t)c * 2 * m + c];
                for (int j = 0; j < 2 * m; j++) A[(size_t)c * 2 * m + j] *= piv;
                for (int r = 0; r < m; r++) {
                    if (r == c) continue;
                    const long double f = A[(size_t)r * 2 * m + c];
                    if (f == 0.0L) continue;
                    for (int j = 0; j < 2 * m; j++) A[(size_t)r * 2 * m + j] -= f * A[(size_t)c * 2 * m + j];
                }
            }
            for (int i = 0; i < FAST_K; i++)
                for (int j = 0; j < FAST_COLS; j++) {
                    if (!end) wtop[(size_t)i * FAST_COLS + j] = (double)A[(size_t)i * 2 * m + m + j];
                    else wbot[(size_t)i * FAST_COLS + j] = (double)A[(size_t)(m - FAST_K + i) * 2 * m + m + (m - FAST_COLS + j)];
                }
        }
    }
};

// ---- one segment of one line ------------------------------------------------------------------------------------------
// IO: load(row0, v[E]) reads rows row0 .. row0+E-1 of the line (rows outside [0, n): the ring's rows for periodic lines,
// zeros otherwise), store(row0, v[E]) writes the rows that exist.  Rows [r0, r1) are produced, r0 a multiple of E.
// The window holds the (scaled) forward solution of boxes b .. b+KB; box b+KB is appended, then the backward
// recurrence runs from the window's end (zero state, >= KB*E = FAST_K rows beyond box b) down to box b.
template <typename T, int E, int KB, typename IO> GB_FAST_FN void fast_segment(IO& io, i64 r0, i64 r1, T l, T rd) {
    constexpr int W = (1 + KB) * E;
    T w[W];
    T y = (T)0;
#pragma unroll
    for (int k = 0; k < KB; k++) {  // warm-up: the K rows before the segment
        T v[E];
        io.load(r0 - (i64)(KB - k) * E, v);
#pragma unroll
        for (int e = 0; e < E; e++) y = fma(-l, y, v[e]);
    }
#pragma unroll
    for (int k = 0; k < KB; k++) {
        T v[E];
        io.load(r0 + (i64)k * E, v);
#pragma unroll
        for (int e = 0; e < E; e++) {
            y = fma(-l, y, v[e]);
            w[k * E + e] = y * rd;
        }
    }
    for (i64 r = r0; r < r1; r += E) {
        T v[E];
        io.load(r + (i64)KB * E, v);
#pragma unroll
        for (int e = 0; e < E; e++) {
            y = fma(-l, y, v[e]);
            w[KB * E + e] = y * rd;
        }
        T x = (T)0;
#pragma unroll
        for (int j = W - 1; j >= E; j--) x = fma(-l, x, w[j]);
#pragma unroll
        for (int j = E - 1; j >= 0; j--) {
            x = fma(-l, x, w[j]);
            v[j] = x;
        }
        io.store(r, v);
#pragma unroll
        for (int j = 0; j < KB * E; j++) w[j] = w[j + E];
    }
}

// The K rows at one end of one line: out[i] = sum_j W[i][j] f[j] (double accumulation), 8 rows at a time.  (The kernels spread
// the same dot products over more threads -- fast_ends_*_kernel; this form serves the host-side arithmetic check.)
// get(j) returns f[j] of the end's COLS-wide window, put(i, v) stores row i of the end.
template <typename T, typename GET, typename PUT> GB_FAST_FN void fast_end_rows(const double* __restrict__ W, GET get, PUT put) {
    for (int i0 = 0; i0 < FAST_K; i0 += 8) {
        double acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.0;
        for (int j = 0; j < FAST_COLS; j++) {
            const double f = (double)get(j);
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = fma(W[(i0 + k) * FAST_COLS + j], f, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) put(i0 + k, (T)acc[k]);
    }
}

}  // namespace gb
