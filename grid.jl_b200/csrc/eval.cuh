// eval.cuh -- point evaluation of InterpGrid / CoordInterpGrid on sm_100a.
//
// One thread evaluates one query: it computes, per dimension, the tap coordinates under the
// boundary condition (branch structure of src/interp.jl:594-783, resolved at compile time per
// (order, BC family)), the B-spline value / gradient / Hessian weights in registers
// (src/interp.jl:785-847), and contracts the l^N coefficient neighbourhood.  Two contraction schemes
// share the taps and weights:
//   EXACT  the reference's arithmetic: coefficient of a tap = ((w_N*w_{N-1})*...)*w_1
//          (src/interp.jl:1130-1183), value = sequential sum over taps, dim 1 fastest, of
//          c==0 ? 0 : c*A[idx] (src/interp.jl:501-505); gradient / Hessian passes rebuild the
//          products with g/h weights (src/interp.jl:443-484).  All passes share ONE sweep over
//          the taps (each tap is loaded once); per-pass accumulation order is unchanged, so the
//          results are bit-identical to the scalar reference algorithm.
//   FAST   separable contraction, dimension by dimension, with FMA.
// and two schedules feed them (identical results):
//   plain  taps gathered from global memory in the caller's order (eval_kernel)
//   brick  the batch is counting-sorted by grid tile and evaluated tile by tile out of shared
//          memory (tile_hist / tile_scatter / eval_brick_kernel / eval_deferred_kernel /
//          unpermute_kernel, second half of this file)
// plus the separable tensor-product path (outer_table_kernel / outer_eval_kernel).
// Compiled with -fmad=false: plain * and + never fuse.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gb {

template <typename T> __device__ __forceinline__ T qnan();
template <> __device__ __forceinline__ float qnan<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double qnan<double>() { return __longlong_as_double(0x7ff8000000000000ll); }

__device__ __forceinline__ i64 rint_ll(double x) { return __double2ll_rn(x); }  // round(Int,x): ties to even
__device__ __forceinline__ i64 rint_ll(float x) { return __float2ll_rn(x); }
__device__ __forceinline__ i64 floor_ll(double x) { return __double2ll_rd(x); }
__device__ __forceinline__ i64 floor_ll(float x) { return __float2ll_rd(x); }
__device__ __forceinline__ i64 trunc_ll(double x) { return __double2ll_rz(x); }
__device__ __forceinline__ i64 trunc_ll(float x) { return __float2ll_rz(x); }

// Julia's floored mod for a positive 32-bit modulus (src/boundaryconditions.jl:23,25).
__device__ __forceinline__ int fmod_floor(i64 a, int m) {
    if ((i64)(int)a == a) {
        int r = (int)a % m;
        return r < 0 ? r + m : r;
    }
    i64 r = a % (i64)m;
    return (int)(r < 0 ? r + m : r);
}

// wrap(BC, pos, len) -> 1-based index (src/boundaryconditions.jl:22-26)
template <int BCF> __device__ __forceinline__ i64 wrap1(i64 pos, int len) {
    if constexpr (BCF == BCF_REFLECT) {
        int r = fmod_floor(pos - 1, 2 * len);
        return r < len ? r + 1 : 2 * len - r;
    } else if constexpr (BCF == BCF_PERIODIC) {
        return fmod_floor(pos - 1, len) + 1;
    } else if constexpr (BCF == BCF_CLAMP) {
        return pos < 1 ? 1 : (pos > len ? len : pos);
    } else {
        return pos;
    }
}

template <typename T> __device__ __forceinline__ bool representable(T x) {
    return fabs((double)x) < 4611686018427387904.0;  // false for NaN / Inf / |x| >= 2^62 (D4)
}

// ---- 1-D tap coordinates ---------------------------------------------------------------------
// Returns validity (src/interp.jl:585-591 with D2/D4); c[] receives 0-based coordinates.  A
// coordinate may lie outside [0,len) only for a tap the reference would not dereference or that
// the D1 guard skips.
template <typename T, int ORDER, int BCF> __device__ __forceinline__ bool dim_taps(T x, int len, int (&c)[ORDER + 1], T& dx, bool& iswrap) {
    constexpr bool F32 = sizeof(T) == 4;
    bool valid = representable(x);
    iswrap = false;  // only consulted for InterpLinear (see the kernel)
    const double xd = (double)x;
    if constexpr (ORDER == GB_NEAREST) {  // :620-624
        if constexpr (BCF == BCF_NIL) {
            double xm = F32 ? xd - 0.5 : (double)(x - (T)0.5);
            valid = valid && xd >= 0.5 && xm <= (double)len;
        }
        i64 ix = wrap1<BCF>(valid ? rint_ll(x) : 1, len);
        if constexpr (BCF == BCF_NIL) ix = ix < 1 ? 1 : (ix > len ? len : ix);  // D3
        c[0] = (int)ix - 1;
        dx = (T)0;
    } else if constexpr (ORDER == GB_LINEAR) {
        if constexpr (BCF == BCF_NIL) valid = valid && x >= (T)1 && xd <= (double)len;
        const T xs = valid ? x : (T)1;
        if constexpr (BCF == BCF_REFLECT) {  // :642-663
            i64 fx = floor_ll(xs);
            int ix = fmod_floor(fx - 1, 2 * len);
            dx = xs - (T)fx;
            int ixp;
            if (ix < len) {
                ix += 1;
                ixp = ix + 1;
                if (ixp > len) ixp = len;
            } else {
                ix = 2 * len - ix - 1;
                dx = (T)1 - dx;
                ixp = ix + 1;
                if (ix == 0) ix = 1;
            }
            c[0] = ix - 1;
            c[1] = (int)wrap1<BCF_REFLECT>(ixp, len) - 1;
            iswrap = (ix == ixp && dx > (T)0);
        } else {  // :632-641
            i64 ifx = floor_ll(xs);
            dx = xs - (T)ifx;
            ifx -= (xs < (T)ifx) ? 1 : 0;
            i64 ix = wrap1<BCF>(ifx, len);
            iswrap = (ix == len && dx > (T)0) || (ix == 1 && ifx < 1);
            c[0] = (int)ix - 1;
            c[1] = iswrap ? (int)ix - 1 : (int)wrap1<BCF>(ix + 1, len) - 1;
        }
    } else if constexpr (ORDER == GB_QUADRATIC) {
        if constexpr (BCF == BCF_NIL) valid = valid && x >= (T)1 && xd <= (double)len;
        const T xs = valid ? x : (T)2;
        const double xsd = (double)xs;
        if constexpr (BCF == BCF_NIL) {  // :689-702
            double xp = xsd + 0.5;
            i64 ix;
            if (xsd > 1.5 && xp < (double)len) ix = rint_ll(xs);
            else if (xsd <= 1.5) ix = 2;
            else ix = len - 1;
            dx = xs - (T)ix;
            c[0] = (int)ix - 2;
            c[1] = (int)ix - 1;
            c[2] = (int)ix;
        } else if constexpr (BCF == BCF_CLAMP) {  // :704-736
            double xp = xsd + 0.5;
            if (xsd > 1.5 && xp < (double)len) {
                i64 ix = rint_ll(xs);
                c[0] = (int)ix - 2;
                c[1] = (int)ix - 1;
                c[2] = (int)ix;
                dx = xs - (T)ix;
            } else if (xsd <= 1.5) {
                c[0] = 1; c[1] = 0; c[2] = 1;
                dx = (xs < (T)1) ? (T)0 : xs - (T)1;
            } else {
                c[0] = len - 2; c[1] = len - 1; c[2] = len - 2;
                dx = (xsd > (double)len) ? (T)0 : xs - (T)len;
            }
        } else if constexpr (BCF == BCF_REFLECT) {  // :737-757
            i64 r0 = rint_ll(xs);
            dx = xs - (T)r0;
            int ix = fmod_floor(r0 - 1, 2 * len);
            if (ix < len) ix += 1;
            else {
                dx = -dx;
                ix = 2 * len - ix;
            }
            c[1] = ix - 1;
            if (ix == 1 || ix == len) {
                c[0] = (int)wrap1<BCF_REFLECT>(ix - 1, len) - 1;
                c[2] = (int)wrap1<BCF_REFLECT>(ix + 1, len) - 1;
            } else {
                c[0] = ix - 2;
                c[2] = ix;
            }
        } else {  // generic :672-686 (BCperiodic)
            i64 r0 = rint_ll(xs);
            dx = xs - (T)r0;
            int ix = (int)wrap1<BCF>(r0, len);
            c[1] = ix - 1;
            if (ix == 1 || ix == len) {
                c[0] = (int)wrap1<BCF>(ix - 1, len) - 1;
                c[2] = (int)wrap1<BCF>(ix + 1, len) - 1;
            } else {
                c[0] = ix - 2;
                c[2] = ix;
            }
        }
    } else {  // GB_CUBIC :767-783 (BCnil/nan/na only; wrap is the identity)
        valid = valid && xd >= 2.0 && xd <= (double)(len - 1);  // D2 (reference: 1 <= x)
        const T xs = valid ? x : (T)2;
        i64 ix = trunc_ll(xs);
        dx = xs - (T)ix;
        c[0] = (int)ix - 2;
        c[1] = (int)ix - 1;
        c[2] = (int)ix;
        c[3] = (int)ix + 1;
    }
    return valid;
}

// ---- 1-D weights (src/interp.jl:785-847).  Julia's promotion rules for Float32 are spelled out:
// Float64 literals (0.5, 0.75, 2/3, 1/2, 3/2, 1.0) promote the expression to Float64 and the
// store rounds to T; Int literals keep T.  x^2 = x*x, x^3 = (x*x)*x.
// FW (GB_FAST only): the two cubic divisions by 6 become multiplications by the rounded reciprocal.
template <int ORDER, bool G, bool H, bool FW = false>
__device__ __forceinline__ void weights(double dx, double (&w)[ORDER + 1], double (&g)[ORDER + 1], double (&h)[ORDER + 1]) {
    if constexpr (ORDER == GB_NEAREST) {
        w[0] = 1.0;
        if (G) g[0] = 0.0;
    } else if constexpr (ORDER == GB_LINEAR) {
        w[0] = 1.0 - dx; w[1] = dx;
        if (G) { g[0] = -1.0; g[1] = 1.0; }
    } else if constexpr (ORDER == GB_QUADRATIC) {
        double a = dx - 0.5, b = dx + 0.5;
        w[0] = (a * a) / 2; w[1] = 0.75 - dx * dx; w[2] = (b * b) / 2;
        if (G) { g[0] = a; g[1] = -2 * dx; g[2] = b; }
        if (H) { h[0] = 1.0; h[1] = -2.0; h[2] = 1.0; }
    } else {
        double t = 1 - dx, u = dx - 1, d2 = dx * dx;
        w[0] = FW ? ((t * t) * t) * (1.0 / 6.0) : div6_exact((t * t) * t);
        w[1] = 2.0 / 3.0 - (d2 * (2 - dx)) / 2;
        w[2] = 2.0 / 3.0 - ((u * u) * (1 + dx)) / 2;
        w[3] = FW ? (d2 * dx) * (1.0 / 6.0) : div6_exact(d2 * dx);
        if (G) {
            g[0] = (-(t * t)) / 2;
            g[1] = (3 * d2) / 2 - 2 * dx;
            g[2] = 0.5 + dx * (1 - 1.5 * dx);
            g[3] = d2 / 2;
        }
        if (H) { h[0] = 1 - dx; h[1] = 3 * dx - 2; h[2] = 1 - 3 * dx; h[3] = dx; }
    }
}
template <int ORDER, bool G, bool H, bool FW = false>
__device__ __forceinline__ void weights(float dx, float (&w)[ORDER + 1], float (&g)[ORDER + 1], float (&h)[ORDER + 1]) {
    if constexpr (ORDER == GB_NEAREST) {
        w[0] = 1.0f;
        if (G) g[0] = 0.0f;
    } else if constexpr (ORDER == GB_LINEAR) {
        w[0] = (float)(1.0 - (double)dx); w[1] = dx;
        if (G) { g[0] = -1.0f; g[1] = 1.0f; }
    } else if constexpr (ORDER == GB_QUADRATIC) {
        double a = (double)dx - 0.5, b = (double)dx + 0.5;
        float d2 = dx * dx;
        w[0] = (float)((a * a) / 2); w[1] = (float)(0.75 - (double)d2); w[2] = (float)((b * b) / 2);
        if (G) { g[0] = (float)a; g[1] = -2.0f * dx; g[2] = (float)b; }
        if (H) { h[0] = 1.0f; h[1] = -2.0f; h[2] = 1.0f; }
    } else {
        float t = 1.0f - dx, u = dx - 1.0f, d2 = dx * dx;
        w[0] = ((t * t) * t) / 6.0f;
        float p = (d2 * (2.0f - dx)) / 2.0f;
        w[1] = (float)(2.0 / 3.0 - (double)p);
        float q = ((u * u) * (1.0f + dx)) / 2.0f;
        w[2] = (float)(2.0 / 3.0 - (double)q);
        w[3] = (d2 * dx) / 6.0f;
        if (G) {
            double d = (double)dx;
            g[0] = (-(t * t)) / 2.0f;
            g[1] = (3.0f * d2) / 2.0f - 2.0f * dx;
            g[2] = (float)(0.5 + d * (1 - 1.5 * d));
            g[3] = d2 / 2.0f;
        }
        if (H) { h[0] = 1.0f - dx; h[1] = 3.0f * dx - 2.0f; h[2] = 1.0f - 3.0f * dx; h[3] = dx; }
    }
}

// ---- pass tables ---------------------------------------------------------------------------
// pass 0: value; passes 1..N: gradient component d; then Hessian entries (i<=j), i-major.
// sel(k,d) picks the weight set used in dimension d for pass k: 0 value, 1 gradient, 2 Hessian
// (set_gradient_coordinate / set_hessian_coordinate, src/interp.jl:443-484).
__host__ __device__ constexpr int num_passes(int N, int outmode) { return outmode == 0 ? 1 : outmode == 1 ? 1 + N : 1 + N + N * (N + 1) / 2; }
__host__ __device__ constexpr int hess_i(int N, int k) {  // k-th pair (i<=j) -> i
    int i = 0, rem = k;
    while (rem >= N - i) { rem -= N - i; i++; }
    return i;
}
__host__ __device__ constexpr int hess_j(int N, int k) {
    int i = 0, rem = k;
    while (rem >= N - i) { rem -= N - i; i++; }
    return i + rem;
}
__host__ __device__ constexpr int sel(int N, int k, int d) {
    if (k == 0) return 0;
    if (k <= N) return d == k - 1 ? 1 : 0;
    int i = hess_i(N, k - 1 - N), j = hess_j(N, k - 1 - N);
    if (i == j) return d == i ? 2 : 0;
    return (d == i || d == j) ? 1 : 0;
}


// ---- tap sources ---------------------------------------------------------------------------------
// The contractions below walk the l^N neighbourhood through a `cursor` (element offset) that a tap
// source advances per dimension:
//   GlobalTaps  the coefficient array in global memory; per-dimension offsets of the (possibly
//               wrapped, possibly repeated) taps; GUARD = D1 (a tap past the array contributes c*0)
//   BrickTaps   the CTA's shared-memory brick (eval_brick_kernel); taps of every dimension are
//               consecutive cells and the brick strides are compile-time constants, so every tap is
//               one LDS at base + immediate
template <typename T> __device__ __forceinline__ bool nonfinite(T a);
template <> __device__ __forceinline__ bool nonfinite<double>(double a) { return (__double2hiint(a) & 0x7ff00000) == 0x7ff00000; }
template <> __device__ __forceinline__ bool nonfinite<float>(float a) { return (__float_as_int(a) & 0x7f800000) == 0x7f800000; }

template <typename T, int N, int L, bool GUARD> struct GlobalTaps {
    const T* __restrict__ A;
    u32 total;
    u32 off[N][L];
    template <int D> __device__ __forceinline__ u32 step(u32 cur, int i) const { return cur + off[D][i]; }
    __device__ __forceinline__ T load(u32 j) const {
        if constexpr (GUARD) return j < total ? A[j] : (T)0;
        else return A[j];
    }
};

template <typename T, int N, int L> struct BrickTaps {
    const T* b;  // brick + offset of the query's first tap
    template <int D> __device__ __forceinline__ u32 step(u32 cur, int i) const { return cur + (u32)(i * brick_stride(N, L, D, (int)sizeof(T))); }
    __device__ __forceinline__ T load(u32 j) const { return b[j]; }
};

// ---- EXACT contraction -------------------------------------------------------------------------
// acc_k += c_k * a per tap, tap order dim-1-fastest, c_k = ((w_N*w_{N-1})*...)*w_1 per pass k.
// The reference's `c == 0 ? 0 : c*a` (src/interp.jl:504) differs from the plain product only when
// a is NaN/Inf (0*a is then NaN instead of 0); adding the resulting +-0 never changes an
// accumulator that started at +0.  So the hot loop forms the product unconditionally, with no
// branch between the tap loads.  A non-finite tap always leaves the VALUE accumulator non-finite
// (every tap enters pass 0, c is finite, and Inf/NaN survive finite additions), so the caller
// re-evaluates exactly those queries with the literal select (eval_one_slow).  Same bits, a third
// of the instructions.  Taps are loaded a plane (L x L) at a time so the loads are in flight together.
template <typename T, int N, int L, int NP, typename SRC, int D> struct ExactRec {
    static __device__ __forceinline__ void run(const SRC& src, u32 cur, const T (&w)[3][N][L], const T (&pp)[NP], T (&acc)[NP]) {
        if constexpr (D <= 1) {
            constexpr int L1 = (D == 1) ? L : 1;
            T a[L1][L];
#pragma unroll
            for (int i1 = 0; i1 < L1; i1++) {
                const u32 row = (D == 1) ? src.template step<D>(cur, i1) : cur;
#pragma unroll
                for (int i0 = 0; i0 < L; i0++) a[i1][i0] = src.load(src.template step<0>(row, i0));
            }
#pragma unroll
            for (int i1 = 0; i1 < L1; i1++) {
                T p1[NP];
#pragma unroll
                for (int k = 0; k < NP; k++) {
                    if constexpr (D == 1) {
                        const T s = w[sel(N, k, 1)][1][i1];
                        p1[k] = (N == 2) ? s : pp[k] * s;
                    } else {
                        p1[k] = pp[k];  // N == 1: no outer factor
                    }
                }
#pragma unroll
                for (int i0 = 0; i0 < L; i0++) {
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        const T s = w[sel(N, k, 0)][0][i0];
                        const T c = (N == 1) ? s : p1[k] * s;
                        acc[k] = acc[k] + c * a[i1][i0];
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < L; i++) {
                T p[NP];
#pragma unroll
                for (int k = 0; k < NP; k++) {
                    const T s = w[sel(N, k, D)][D][i];
                    p[k] = (D == N - 1) ? s : pp[k] * s;
                }
                ExactRec<T, N, L, NP, SRC, D - 1>::run(src, src.template step<D>(cur, i), w, p, acc);
            }
        }
    }
};

// ---- FAST contraction: separable, FMA --------------------------------------------------------
// Q = value, gradient components 0..D and Hessian entries (i<=j<=D) of the partial contraction
// over dimensions 0..D.  Non-finite taps are handled like in EXACT (value non-finite -> eval_one_slow).
template <typename T, int N, int OUTMODE> struct FastQ {
    T v;
    T g[OUTMODE >= 1 ? N : 1];
    T h[OUTMODE >= 2 ? N * N : 1];  // [i*N+j], i<=j
};
template <typename T, int N, int OUTMODE, int D> __device__ __forceinline__ void fastq_zero(FastQ<T, N, OUTMODE>& q) {
    q.v = (T)0;
    if constexpr (OUTMODE >= 1) {
#pragma unroll
        for (int k = 0; k <= D; k++) q.g[k] = (T)0;
    }
    if constexpr (OUTMODE >= 2) {
#pragma unroll
        for (int a = 0; a <= D; a++)
#pragma unroll
            for (int b = a; b <= D; b++) q.h[a * N + b] = (T)0;
    }
}
// dimension 0 over L register-resident samples
template <typename T, int N, int L, int OUTMODE>
__device__ __forceinline__ FastQ<T, N, OUTMODE> fast_row(const T (&a)[L], const T (&w)[3][N][L]) {
    FastQ<T, N, OUTMODE> out;
    fastq_zero<T, N, OUTMODE, 0>(out);
#pragma unroll
    for (int i = 0; i < L; i++) {
        out.v = fma(w[0][0][i], a[i], out.v);
        if constexpr (OUTMODE >= 1) out.g[0] = fma(w[1][0][i], a[i], out.g[0]);
        if constexpr (OUTMODE >= 2) out.h[0] = fma(w[2][0][i], a[i], out.h[0]);
    }
    return out;
}
// fold a lower-dimensional partial result `in` into `out` along dimension D at tap i
template <typename T, int N, int L, int OUTMODE, int D>
__device__ __forceinline__ void fast_fold(FastQ<T, N, OUTMODE>& out, const FastQ<T, N, OUTMODE>& in, const T (&w)[3][N][L], int i) {
    const T wv = w[0][D][i];
    out.v = fma(wv, in.v, out.v);
    if constexpr (OUTMODE >= 1) {
        const T wg = w[1][D][i];
#pragma unroll
        for (int k = 0; k < D; k++) out.g[k] = fma(wv, in.g[k], out.g[k]);
        out.g[D] = fma(wg, in.v, out.g[D]);
        if constexpr (OUTMODE >= 2) {
            const T wh = w[2][D][i];
#pragma unroll
            for (int a = 0; a < D; a++) {
#pragma unroll
                for (int b = a; b < D; b++) out.h[a * N + b] = fma(wv, in.h[a * N + b], out.h[a * N + b]);
                out.h[a * N + D] = fma(wg, in.g[a], out.h[a * N + D]);
            }
            out.h[D * N + D] = fma(wh, in.v, out.h[D * N + D]);
        }
    }
}
template <typename T, int N, int L, int OUTMODE, typename SRC, int D> struct FastRec {
    static __device__ __forceinline__ FastQ<T, N, OUTMODE> run(const SRC& src, u32 cur, const T (&w)[3][N][L]) {
        if constexpr (D == 0) {  // N == 1
            T a[L];
#pragma unroll
            for (int i = 0; i < L; i++) a[i] = src.load(src.template step<0>(cur, i));
            return fast_row<T, N, L, OUTMODE>(a, w);
        } else if constexpr (D == 1) {  // one plane: L*L loads issued together
            T a[L][L];
#pragma unroll
            for (int i1 = 0; i1 < L; i1++) {
                const u32 row = src.template step<1>(cur, i1);
#pragma unroll
                for (int i0 = 0; i0 < L; i0++) a[i1][i0] = src.load(src.template step<0>(row, i0));
            }
            FastQ<T, N, OUTMODE> out;
            fastq_zero<T, N, OUTMODE, 1>(out);
#pragma unroll
            for (int i1 = 0; i1 < L; i1++) {
                const FastQ<T, N, OUTMODE> in = fast_row<T, N, L, OUTMODE>(a[i1], w);
                fast_fold<T, N, L, OUTMODE, 1>(out, in, w, i1);
            }
            return out;
        } else {
            FastQ<T, N, OUTMODE> out;
            fastq_zero<T, N, OUTMODE, D>(out);
#pragma unroll
            for (int i = 0; i < L; i++) {
                const FastQ<T, N, OUTMODE> in = FastRec<T, N, L, OUTMODE, SRC, D - 1>::run(src, src.template step<D>(cur, i), w);
                fast_fold<T, N, L, OUTMODE, D>(out, in, w, i);
            }
            return out;
        }
    }
};

// both schemes behind one call: acc[0] value, acc[1..N] gradient, then Hessian entries (i<=j), i-major
template <typename T, int N, int L, int OUTMODE, bool FAST, typename SRC>
__device__ __forceinline__ void contract(const SRC& src, const T (&w)[3][N][L], T (&acc)[num_passes(N, OUTMODE)]) {
    constexpr int NP = num_passes(N, OUTMODE);
    if constexpr (!FAST) {
        T pp[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) { acc[k] = (T)0; pp[k] = (T)1; }
        ExactRec<T, N, L, NP, SRC, N - 1>::run(src, 0u, w, pp, acc);
    } else {
        const FastQ<T, N, OUTMODE> r = FastRec<T, N, L, OUTMODE, SRC, N - 1>::run(src, 0u, w);
        acc[0] = r.v;
        if constexpr (OUTMODE >= 1) {
#pragma unroll
            for (int d = 0; d < N; d++) acc[1 + d] = r.g[d];
        }
        if constexpr (OUTMODE >= 2) {
#pragma unroll
            for (int k = 0; k < N * (N + 1) / 2; k++) acc[1 + N + k] = r.h[hess_i(N, k) * N + hess_j(N, k)];
        }
    }
}

// ---- coordinate fetch: user coordinate -> index coordinate in T --------------------------------
// (coordlookup src/coord.jl:30-32, setx src/interp.jl:75-139), each operation in Julia's type.
// load_raw reads the stored coordinate (Float32 values are carried exactly in a double).
template <int N> __device__ __forceinline__ double load_raw(const EvalParams<N>& p, int d, i64 q) {
    if (p.xdtype == GB_F64) return __ldg((const double*)p.xp[d] + q * p.xqs);
    return (double)__ldg((const float*)p.xp[d] + q * p.xqs);
}
template <typename T, int N> __device__ __forceinline__ T index_coord(const EvalParams<N>& p, int d, double raw) {
    if (p.xdtype == GB_F64) {
        const double xq = raw;
        if (!p.has_aff) return (T)(p.shift ? xq + 1.0 : xq);
        const int kind = (int)p.aff[d][0];
        if (kind == 0) {
            double u = (p.aff[d][3] * xq - p.aff[d][1]) / p.aff[d][2] + 1.0;
            if (p.shift) u = u + 1.0;
            return (T)u;
        } else if (kind == 1) {
            double u = (xq - p.aff[d][1]) / p.aff[d][2] + 1.0;
            if (p.shift) u = u + 1.0;
            return (T)u;
        } else {
            double u = (xq - p.aff[d][1]) + 1.0;
            if (p.shift) u = u + 1.0;
            return (T)u;
        }
    } else {
        const float xq = (float)raw;
        if (!p.has_aff) return (T)(p.shift ? xq + 1.0f : xq);
        const int kind = (int)p.aff[d][0];
        if (kind == 0) {
            double u = (p.aff[d][3] * (double)xq - p.aff[d][1]) / p.aff[d][2] + 1.0;
            if (p.shift) u = u + 1.0;
            return (T)u;
        } else if (kind == 1) {
            float t = (xq - (float)p.aff[d][1]) / (float)p.aff[d][2];
            double u = (double)t + 1.0;
            if (p.shift) u = u + 1.0;
            return (T)u;
        } else {
            float t = (xq - (float)p.aff[d][1]) + 1.0f;
            if (p.shift) t = t + 1.0f;
            return (T)t;
        }
    }
}
template <typename T, int N> __device__ __forceinline__ T scale_grad(const EvalParams<N>& p, int d, T g) {  // coordiscale :54-56
    if (!p.has_aff) return g;
    const int kind = (int)p.aff[d][0];
    if (kind == 0) return (T)(((double)g * p.aff[d][3]) / p.aff[d][2]);
    if (kind == 1) return g / (T)p.aff[d][2];
    return g;
}

// Element offset contributed by coordinate c of dimension d.  Linear (column-major) layout: c * stride[d].
// Blocked layout (p.blocked; 2-D to 4-D grids too large for L2 whose neighbourhoods are too light for the brick
// path): the array is stored as 4^N-element blocks, Morton order inside a block (bit 0 of every coordinate,
// then bit 1), blocks in column-major order.  The DRAM granule (64 B) then holds a 4x4 (f32) / 4x2 (f64) patch of
// a 2-D grid or a 4x2x2 / 2x2x2 patch of a 3-D grid instead of 16 / 8 elements of one row, so the l^N
// neighbourhood of a random query costs ~(1 + (l-1)/4)(1 + (l-1)/2)... granules instead of l^(N-1) or more:
// the plain path on such grids is bound by exactly that traffic (profiles/README.md).  The offset stays a
// SUM of per-dimension terms, so the contractions are untouched; results are bit-identical.
template <int N> __device__ __forceinline__ u32 tap_offset(const EvalParams<N>& p, int d, int c) {
    if constexpr (N >= 2 && N <= 4) {
        if (p.blocked) return (u32)(c >> 2) * p.bstride[d] + (u32)((((c >> 1) & 1) << (N + d)) | ((c & 1) << d));
    }
    return (u32)c * (u32)p.stride[d];
}
// column-major element index -> index in the resident array (rare paths only)
template <int N> __device__ __forceinline__ i64 resident_index(const EvalParams<N>& p, i64 j) {
    if constexpr (N >= 2 && N <= 4) {
        if (p.blocked) {
            u32 o = 0;
#pragma unroll
            for (int d = 0; d < N; d++) {
                o += tap_offset<N>(p, d, (int)(j % p.len[d]));
                j /= p.len[d];
            }
            return (i64)o;
        }
    }
    return j;
}
// Blocked layout and the D1 taps (one past a dimension's end: the reference's offset arithmetic carries them into
// the next column / plane): such queries take eval_one_slow, which addresses through resident_index.
template <int N, int L> __device__ __forceinline__ bool blocked_carry(const EvalParams<N>& p, const int (&c)[N][L]) {
    bool carry = false;
    if constexpr (N >= 2 && N <= 4) {
        if (p.blocked) {
#pragma unroll
            for (int d = 0; d < N; d++) {
#pragma unroll
                for (int i = 0; i < L; i++) carry = carry || c[d][i] >= p.len[d];
            }
        }
    }
    return carry;
}

// taps of one query: validity + 0-based tap coordinates, including the reference's offset-table
// addressing for InterpLinear.
// set_position (src/interp.jl:406-428): unless SOME dimension reported a wrap, the reference
// addresses taps through the precomputed offset table (ix + 0, ix + 1, ...) and ignores coord1d.
// The two differ only for InterpLinear exactly on a grid line at the upper edge (ix == len,
// dx == 0): the second tap is then ix+1 (its value weight is 0, its gradient weight is not).
// Reproduced for bit-parity; the D1 guard keeps the load in bounds.
template <typename T, int N, int ORDER, int BCF>
__device__ __forceinline__ bool query_taps(const EvalParams<N>& p, const T (&x)[N], int (&c)[N][ORDER + 1], T (&dx)[N]) {
    bool valid = true, anywrap = false;
#pragma unroll
    for (int d = 0; d < N; d++) {
        bool iswrap;
        valid = dim_taps<T, ORDER, BCF>(x[d], p.len[d], c[d], dx[d], iswrap) && valid;
        anywrap = anywrap || iswrap;
    }
    if constexpr (ORDER == GB_LINEAR && BCF != BCF_NIL) {
        if (!anywrap) {
#pragma unroll
            for (int d = 0; d < N; d++) c[d][1] = c[d][0] + 1;
        }
    }
    return valid;
}
template <typename T, int N> __device__ __forceinline__ void load_index_coords(const EvalParams<N>& p, i64 q, T (&x)[N]) {
#pragma unroll
    for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, load_raw<N>(p, d, q));
}

// outputs of one query (getindex / valgrad / valgradhess layouts, src/interp.jl:142-270).
// `pos` addresses the outputs, `q` is the query's index in the caller's batch (BCnil reporting).
// Plain mode: pos == q and val / grad / hess are the caller's arrays.  Tiled mode (p.aos != 0): pos is
// the query's sorted slot and the outputs of a slot form one record of p.aos words at p.val
// (value, gradient, full Hessian) -- written coalesced here, brought back to the caller's order and
// arrays by unpermute_kernel.
template <int N> __host__ __device__ constexpr int out_words(int outmode) { return 1 + (outmode >= 1 ? N : 0) + (outmode >= 2 ? N * N : 0); }
template <typename T, int N, int OUTMODE>
__device__ __forceinline__ void store_outputs(const EvalParams<N>& p, i64 pos, i64 q, bool valid, T (&acc)[num_passes(N, OUTMODE)]) {
    constexpr int NP = num_passes(N, OUTMODE);
    if (!valid) {
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = qnan<T>();
        if (p.first_invalid) atomicMin(p.first_invalid, (u64)(q + p.q0));
    }
    if (p.aos) {
        constexpr int W = out_words<N>(OUTMODE);
        alignas(16) T o[W];
        o[0] = acc[0];
        if constexpr (OUTMODE >= 1) {
#pragma unroll
            for (int d = 0; d < N; d++) o[1 + d] = valid ? scale_grad<T, N>(p, d, acc[1 + d]) : acc[1 + d];
        }
        if constexpr (OUTMODE >= 2) {
#pragma unroll
            for (int k = 0; k < N * (N + 1) / 2; k++) {
                const int i = hess_i(N, k), j = hess_j(N, k);
                o[1 + N + j * N + i] = acc[1 + N + k];
                o[1 + N + i * N + j] = acc[1 + N + k];
            }
        }
        T* dst = (T*)p.val + pos * W;
        if constexpr (W * sizeof(T) == 32) {
            reinterpret_cast<float4*>(dst)[0] = reinterpret_cast<const float4*>(o)[0];
            reinterpret_cast<float4*>(dst)[1] = reinterpret_cast<const float4*>(o)[1];
        } else if constexpr (W * sizeof(T) == 16) {
            reinterpret_cast<float4*>(dst)[0] = reinterpret_cast<const float4*>(o)[0];
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) dst[k] = o[k];
        }
        return;
    }
    if (p.val) ((T*)p.val)[pos] = acc[0];
    if constexpr (OUTMODE >= 1) {
        if (p.grad) {
            T* g = (T*)p.grad + pos * N;
#pragma unroll
            for (int d = 0; d < N; d++) g[d] = valid ? scale_grad<T, N>(p, d, acc[1 + d]) : acc[1 + d];
        }
    }
    if constexpr (OUTMODE >= 2) {
        T* h = (T*)p.hess + pos * (N * N);
#pragma unroll
        for (int k = 0; k < N * (N + 1) / 2; k++) {
            const int i = hess_i(N, k), j = hess_j(N, k);
            h[j * N + i] = acc[1 + N + k];
            h[i * N + j] = acc[1 + N + k];
        }
    }
}

// ---- rare path: a query that touched a NaN/Inf coefficient ----------------------------------------
// Literal restatement of interp (src/interp.jl:501-505) with the `c == 0 ? 0 : c*a` select, rolled
// loops, global memory only, one pass per output.  Executed for (almost) no query; kept out of line.
template <typename T, int N, int ORDER, int BCF>
__device__ __noinline__ void eval_one_slow(const EvalParams<N>& p, const T* __restrict__ A, const T (&x)[N], const i64 pos, const i64 q, const int outmode) {
    constexpr int L = ORDER + 1;
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    int c[N][L];
    T dx[N];
    T w[3][N][L];
    const bool valid = query_taps<T, N, ORDER, BCF>(p, x, c, dx);
#pragma unroll
    for (int d = 0; d < N; d++) {
#pragma unroll
        for (int i = 0; i < L; i++) { w[1][d][i] = (T)0; w[2][d][i] = (T)0; }
        weights<ORDER, true, (ORDER >= GB_QUADRATIC)>(dx[d], w[0][d], w[1][d], w[2][d]);
    }
    const int npass = num_passes(N, outmode);
#pragma unroll 1
    for (int k = 0; k < npass; k++) {
        int s[N], hi = 0, hj = 0;
        for (int d = 0; d < N; d++) s[d] = 0;
        if (k >= 1 && k <= N) s[k - 1] = 1;
        else if (k > N) {
            hi = hess_i(N, k - 1 - N);
            hj = hess_j(N, k - 1 - N);
            if (hi == hj) s[hi] = 2;
            else { s[hi] = 1; s[hj] = 1; }
        }
        int idx[N];
        T pr[N];
        i64 off[N];
        for (int d = 0; d < N; d++) idx[d] = 0;
        for (int d = N - 1; d >= 0; d--) {
            const T sw = w[s[d]][d][0];
            pr[d] = (d == N - 1) ? sw : pr[d + 1] * sw;
            off[d] = (d == N - 1 ? 0 : off[d + 1]) + (i64)c[d][0] * p.stride[d];
        }
        T acc = (T)0;
#pragma unroll 1
        while (true) {
            const T cf = pr[0];
            const i64 j = off[0];
            T term = (T)0;
            if (cf != (T)0 && (!GUARD || (u64)j < (u64)p.total)) term = cf * A[resident_index<N>(p, j)];
            acc = acc + term;
            int d = 0;
            for (; d < N; d++) {
                if (++idx[d] < L) break;
                idx[d] = 0;
            }
            if (d == N) break;
            for (int e = d; e >= 0; e--) {
                const T sw = w[s[e]][e][idx[e]];
                pr[e] = (e == N - 1) ? sw : pr[e + 1] * sw;
                off[e] = (e == N - 1 ? 0 : off[e + 1]) + (i64)c[e][idx[e]] * p.stride[e];
            }
        }
        if (!valid) acc = qnan<T>();
        if (p.aos) {  // sorted-slot records (see store_outputs)
            T* o = (T*)p.val + pos * p.aos;
            if (k == 0) o[0] = acc;
            else if (k <= N) o[k] = valid ? scale_grad<T, N>(p, k - 1, acc) : acc;
            else { o[1 + N + hj * N + hi] = acc; o[1 + N + hi * N + hj] = acc; }
        } else if (k == 0) {
            if (p.val) ((T*)p.val)[pos] = acc;
        } else if (k <= N) {
            if (p.grad) ((T*)p.grad)[pos * N + (k - 1)] = valid ? scale_grad<T, N>(p, k - 1, acc) : acc;
        } else {
            T* h = (T*)p.hess + pos * (N * N);
            h[hj * N + hi] = acc;
            h[hi * N + hj] = acc;
        }
    }
    if (!valid && p.first_invalid) atomicMin(p.first_invalid, (u64)(q + p.q0));
}

// ---- one query, taps gathered from global memory ---------------------------------------------------
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__device__ __forceinline__ void eval_one(const EvalParams<N>& p, const T (&x)[N], const i64 pos, const i64 q) {
    constexpr int L = ORDER + 1;
    constexpr int NP = num_passes(N, OUTMODE);
    // D1 guard: taps that can fall outside the array (upper-edge tap of Linear on the offset path,
    // 4th tap of Cubic at x == len-1).
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    int c[N][L];
    T dx[N];
    T w[3][N][L];
    const bool valid = query_taps<T, N, ORDER, BCF>(p, x, c, dx);
#pragma unroll
    for (int d = 0; d < N; d++) weights<ORDER, (OUTMODE >= 1), (OUTMODE >= 2), FAST>(dx[d], w[0][d], w[1][d], w[2][d]);
    GlobalTaps<T, N, L, GUARD> src;
    src.A = (const T*)p.coefs;
    src.total = p.blocked ? 0xffffffffu : (u32)p.total;
    if constexpr (GUARD) {
        if (blocked_carry<N, L>(p, c)) {
            eval_one_slow<T, N, ORDER, BCF>(p, src.A, x, pos, q, OUTMODE);
            return;
        }
    }
#pragma unroll
    for (int d = 0; d < N; d++) {
#pragma unroll
        for (int i = 0; i < L; i++) src.off[d][i] = tap_offset<N>(p, d, c[d][i]);
    }
    if (p.nchan == 1) {  // (kept apart from the channel loop: the loop costs the 3-D cubic kernels their registers)
        T acc[NP];
        contract<T, N, L, OUTMODE, FAST>(src, w, acc);
        if (nonfinite(acc[0])) {  // a NaN/Inf coefficient may have been touched: literal zero-weight semantics
            eval_one_slow<T, N, ORDER, BCF>(p, src.A, x, pos, q, OUTMODE);
            return;
        }
        store_outputs<T, N, OUTMODE>(p, pos, q, valid, acc);
        return;
    }
    // Multi-valued grids (README.md:151-209: set_position once, interp(ic, A, index) per channel): the
    // taps and weights above serve every channel; channel ch of a query lands at output pos*nchan + ch.
#pragma unroll 1
    for (int ch = 0; ch < p.nchan; ch++) {
        T acc[NP];
        contract<T, N, L, OUTMODE, FAST>(src, w, acc);
        const i64 opos = pos * p.nchan + ch;
        if (nonfinite(acc[0])) {
            eval_one_slow<T, N, ORDER, BCF>(p, src.A, x, opos, q, OUTMODE);
        } else {
            store_outputs<T, N, OUTMODE>(p, opos, q, valid, acc);
        }
        src.A += p.cstride;
    }
}
// ---- plain evaluation kernel -----------------------------------------------------------------------
// One thread per query in the caller's order, CTAs stride over blocks of EVAL_BLOCK queries; taps
// are gathered from global memory through L1/L2.  The next query's coordinates are in flight
// while the current one is evaluated.
constexpr int EVAL_BLOCK = 256;

// `only_if` (may be NULL): device flag of the tiled launcher -- the kernel runs only when *only_if != 0 (the batch was found
// coherent, coherence_probe_kernel) and exits at once otherwise.
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(EVAL_BLOCK, 2) eval_kernel(const EvalParams<N> p, const u32* __restrict__ only_if) {
    if (only_if && *only_if == 0) return;
    const i64 step = (i64)gridDim.x * EVAL_BLOCK;
    i64 q = (i64)blockIdx.x * EVAL_BLOCK + threadIdx.x;
    double raw_n[N];
    if (q < p.nq) {
#pragma unroll
        for (int d = 0; d < N; d++) raw_n[d] = load_raw<N>(p, d, q);
    }
    while (q < p.nq) {
        T x[N];
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, raw_n[d]);
        const i64 qc = q;
        q += step;
        if (q < p.nq) {
#pragma unroll
            for (int d = 0; d < N; d++) raw_n[d] = load_raw<N>(p, d, q);
        }
        eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, x, qc, qc);
    }
}

// ====================================================================================================
// Tiled ("brick") evaluation: the batch is sorted by grid tile, then evaluated tile by tile out of
// shared memory.
//   coherence_probe  looks at a sample of the batch in the order it ARRIVED: spatially coherent batches (lattices in raster
//                  order, points sorted by cell / tile / Morton code) skip everything below -- the kernels exit at
//                  once -- and are evaluated in place by the plain kernel, whose gathers then hit L1 / L2; decided on
//                  the device, the host never waits (nwork[2])
//   otherwise      a counting sort:
//   tile_hist      key of every query (tile of its lowest tap per dimension) + per-CTA histograms in
//                  shared memory (no global atomics), written bin-major to bh[bin][cta]
//   tile_offsets   (eval_tile.cu) per-bin totals, scan, work-item list (a tile's queries split into
//                  items of <= EVAL_CHUNK), per-(bin,cta) write cursors
//   tile_scatter   each CTA replays its slice and writes one RECORD per query at its sorted slot:
//                  the N index coordinates (already in T, affine map and setx shift applied) and the
//                  query's original index -- (N+1) words of T, a full 32-byte sector for f64 N=3
//   eval_brick     CTAs pull work items from an atomic counter.  The tile's coefficient brick
//                  (tile + L-1 halo) is staged into one of two shared-memory buffers while the previous item is
//                  being evaluated -- by ONE cp.async.bulk.tensor (TMA) per brick, completion on an mbarrier,
//                  cells past the array zero-filled by the TMA unit; grids whose row pitch is not a multiple of
//                  16 bytes fall back to element-wise cp.async.  Records are read coalesced (the first
//                  record of the next item is prefetched across the item boundary); every tap is an
//                  LDS at a compile-time offset from the query's first tap; outputs go to the
//                  query's sorted slot.
//   eval_deferred  a query whose taps are not L consecutive cells inside the brick (wrapped /
//                  repeated taps at the grid boundary, taps past the last sample, invalid location)
//                  or that met a NaN/Inf coefficient is not evaluated by eval_brick: its slot goes
//                  to a list and this kernel runs the global-memory path (eval_one) on the list, so
//                  no boundary condition needs special bricks and results are bit-identical to plain.
//   unpermute      sorted slots -> the caller's order and arrays
constexpr int BRICK_BLOCK = 256;
#ifndef GB_OUT_BY_QUERY_DEFAULT
#define GB_OUT_BY_QUERY_DEFAULT 1
#endif
constexpr u32 EVAL_CHUNK = 2048;  // queries per work item
constexpr int SORT_BLOCK = 256;
constexpr int SORT_UNROLL = 4;    // queries in flight per thread in the two sort passes
constexpr int PASS1_UNROLL = 4;   // records in flight per thread in the brick kernel's locate pass
constexpr i64 SORT_ITER = (i64)SORT_BLOCK * SORT_UNROLL;  // queries a sort CTA handles per iteration; per-CTA slices are multiples of it

template <typename T> __device__ __forceinline__ T bits_to(u32 q);
template <> __device__ __forceinline__ double bits_to<double>(u32 q) { return __longlong_as_double((long long)q); }
template <> __device__ __forceinline__ float bits_to<float>(u32 q) { return __uint_as_float(q); }
__device__ __forceinline__ u32 bits_of(double v) { return (u32)__double_as_longlong(v); }
__device__ __forceinline__ u32 bits_of(float v) { return __float_as_uint(v); }

template <typename T, int N> __device__ __forceinline__ void store_record(T* __restrict__ rec, u32 slot, const T (&x)[N], u32 q) {
    T* r = rec + (size_t)slot * (N + 1);
    if constexpr (sizeof(T) == 8 && N == 3) {
        reinterpret_cast<double2*>(r)[0] = make_double2(x[0], x[1]);
        reinterpret_cast<double2*>(r)[1] = make_double2(x[2], bits_to<double>(q));
    } else if constexpr (sizeof(T) == 8 && N == 1) {
        reinterpret_cast<double2*>(r)[0] = make_double2(x[0], bits_to<double>(q));
    } else if constexpr (sizeof(T) == 4 && N == 3) {
        reinterpret_cast<float4*>(r)[0] = make_float4(x[0], x[1], x[2], bits_to<float>(q));
    } else if constexpr (sizeof(T) == 4 && N == 1) {
        reinterpret_cast<float2*>(r)[0] = make_float2(x[0], bits_to<float>(q));
    } else {
#pragma unroll
        for (int d = 0; d < N; d++) r[d] = x[d];
        r[N] = bits_to<T>(q);
    }
}
template <typename T, int N> __device__ __forceinline__ void load_record(const T* __restrict__ rec, u32 slot, T (&x)[N], u32& q) {
    const T* r = rec + (size_t)slot * (N + 1);
    if constexpr (sizeof(T) == 8 && N == 3) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(r)), b = __ldg(reinterpret_cast<const double2*>(r) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = b.x; q = bits_of(b.y);
    } else if constexpr (sizeof(T) == 8 && N == 1) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(r));
        x[0] = a.x; q = bits_of(a.y);
    } else if constexpr (sizeof(T) == 4 && N == 3) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(r));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; q = bits_of(a.w);
    } else if constexpr (sizeof(T) == 4 && N == 1) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(r));
        x[0] = a.x; q = bits_of(a.y);
    } else {
#pragma unroll
        for (int d = 0; d < N; d++) x[d] = __ldg(r + d);
        q = bits_of(__ldg(r + N));
    }
}
// sort key of a query: the tile of its lowest tap per dimension (a locality hint only -- queries whose taps
// leave the brick are deferred); invalid locations go to the extra last bin.  first[] receives the coordinate of the
// query's FIRST tap per dimension (clamped into the array), what the coherence probe builds its bounding boxes from.
template <typename T, int N, int ORDER, int BCF>
__device__ __forceinline__ u32 tile_key(const EvalParams<N>& p, const T (&x)[N], const int (&shift)[N], int ntot, int (&first)[N], bool& valid) {
    int c[N][ORDER + 1];
    T dx[N];
    valid = query_taps<T, N, ORDER, BCF>(p, x, c, dx);
    u32 key = 0, mul = 1;
#pragma unroll
    for (int d = 0; d < N; d++) {
        int lo = c[d][0];
#pragma unroll
        for (int i = 1; i <= ORDER; i++) lo = c[d][i] < lo ? c[d][i] : lo;
        lo = lo < 0 ? 0 : (lo >= p.len[d] ? p.len[d] - 1 : lo);
        key += (u32)(lo >> shift[d]) * mul;
        mul *= (u32)p.ntile[d];
        // (taken from the second tap where there is one: at a wrapping boundary the first tap sits at the far end of the array,
        // which would tear the chunk's box apart; such a query is deferred by brick_locate whatever its chunk's brick is)
        const int f = ORDER >= 2 ? c[d][1] - 1 : c[d][0];
        first[d] = f < 0 ? 0 : (f >= p.len[d] ? p.len[d] - 1 : f);
    }
    return valid ? key : (u32)ntot;
}
template <typename T, int N, int ORDER, int BCF>
__device__ __forceinline__ u32 tile_key(const EvalParams<N>& p, const T (&x)[N], const int (&shift)[N], int ntot) {
    int first[N];
    bool valid;
    return tile_key<T, N, ORDER, BCF>(p, x, shift, ntot, first, valid);
}

// ---- coherent batches ------------------------------------------------------------------------------------------------
// Binning pays for batches whose consecutive queries are far apart (uniformly random: every warp of the plain path drags
// 32 x 16 tap rows through L2).  A batch that arrives spatially coherent -- a lattice walked in raster order, points sorted by
// cell / tile / Morton code, an image warp -- is served faster by the plain path, whose tap rows then hit L1 / L2
// (measured on C2, profiles/README.md: 0.62-1.0 ms against 0.9-1.15 ms through the bins).  The probe looks at a sample
// of the batch AS IT ARRIVED -- per sort CTA four groups of 256 consecutive queries, i.e. 32 leaves of 32 consecutive
// queries -- and calls a leaf coherent when the bounding box of its queries' first taps holds at most `vmax` cells (four brick
// tiles).  Nearly all leaves coherent: nwork[2] = 1, the binning kernels exit at once and the plain kernel evaluates the
// batch in place; else nwork[2] = 0 and the plain kernel exits.  Decided on the device: the host never waits.
// nwork[3] coherent leaves, nwork[4] leaves, nwork[5] CTAs done, nwork[6] queries sampled (all zeroed by the launcher).
template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(SORT_BLOCK) coherence_probe_kernel(const EvalParams<N> p, u32* __restrict__ nwork, i64 per_cta, u32 vmax, u32 need_percent,
                                                                     u32* __restrict__ shist) {
    __shared__ u32 s_last;
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    const u32 lane = threadIdx.x & 31;
    int ntot = 1, shift[N];
#pragma unroll
    for (int d = 0; d < N; d++) { ntot *= p.ntile[d]; shift[d] = __ffs(p.tile[d]) - 1; }
    u32 good = 0, seen = 0, sampled = 0;
    if (q0 < q1) {
        const i64 stride = ((q1 - q0) / SORT_UNROLL) & ~(i64)(SORT_BLOCK - 1);  // (groups start on multiples of 256 past q0)
        double raw[SORT_UNROLL][N];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = q0 + u * stride + threadIdx.x;
            if (q < q1 && (u == 0 || stride > 0)) {
#pragma unroll
                for (int d = 0; d < N; d++) raw[u][d] = load_raw<N>(p, d, q);
            }
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = q0 + u * stride + threadIdx.x;
            const bool act = q < q1 && (u == 0 || stride > 0);
            int first[N];
            bool ok = false;
#pragma unroll
            for (int d = 0; d < N; d++) first[d] = 0;
            if (act) {
                T x[N];
#pragma unroll
                for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, raw[u][d]);
                const u32 key = tile_key<T, N, ORDER, BCF>(p, x, shift, ntot, first, ok);
                if (shist) atomicAdd(shist + key, 1u);  // the sample's tile histogram: sizes the bins of the single-pass scatter
            }
            sampled += (u32)__popc(__ballot_sync(0xffffffffu, act));
            if (!__any_sync(0xffffffffu, act)) continue;
            double vol = 1.0;
#pragma unroll
            for (int d = 0; d < N; d++) {
                const int mn = __reduce_min_sync(0xffffffffu, ok ? first[d] : 0x7fffffff);
                const int mx = __reduce_max_sync(0xffffffffu, ok ? first[d] : -1);
                vol *= mx >= mn ? (double)(mx - mn + 1) : 1.0;
            }
            seen++;
            good += vol <= (double)vmax ? 1u : 0u;
        }
    }
    if (lane == 0 && seen) { atomicAdd(nwork + 3, good); atomicAdd(nwork + 4, seen); atomicAdd(nwork + 6, sampled); }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(nwork + 5, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {  // every CTA has reported
        const u32 g = atomicAdd(nwork + 3, 0u), n = atomicAdd(nwork + 4, 0u);
        nwork[2] = (n > 0 && (unsigned long long)g * 100ull >= (unsigned long long)n * need_percent) ? 1u : 0u;
    }
}

template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(SORT_BLOCK) tile_hist_kernel(const EvalParams<N> p, u32* __restrict__ bh, u32 nbins, i64 per_cta) {
    if (p.nwork[2] != 0) return;  // coherent batch (coherence_probe_kernel): the plain kernel takes it
    extern __shared__ u32 s_hist[];
    for (u32 b = threadIdx.x; b < nbins; b += SORT_BLOCK) s_hist[b] = 0;
    __syncthreads();
    int ntot = 1;
#pragma unroll
    for (int d = 0; d < N; d++) ntot *= p.ntile[d];
    int shift[N];  // tile extents are powers of two (choose_tile)
#pragma unroll
    for (int d = 0; d < N; d++) shift[d] = __ffs(p.tile[d]) - 1;
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    for (i64 qb = q0 + threadIdx.x; qb < q1; qb += SORT_ITER) {
        double raw[SORT_UNROLL][N];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
#pragma unroll
                for (int d = 0; d < N; d++) raw[u][d] = load_raw<N>(p, d, q);
            }
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
                T x[N];
#pragma unroll
                for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, raw[u][d]);
                atomicAdd(s_hist + tile_key<T, N, ORDER, BCF>(p, x, shift, ntot), 1u);
            }
        }
    }
    __syncthreads();
    for (u32 b = threadIdx.x; b < nbins; b += SORT_BLOCK) bh[(size_t)b * gridDim.x + blockIdx.x] = s_hist[b];
}

// work item {brick corner dim 1, first slot, one past the last, corner dims 2, 3 | no-brick flag} (written by tile_scan_kernel)
template <int N> __device__ __forceinline__ bool item_corner(const uint4 it, int (&lo)[N]) {
    lo[0] = (int)it.x;
    if constexpr (N == 2) lo[1] = (int)(it.w & 0x7fffffffu);
    if constexpr (N >= 3) { lo[1] = (int)(it.w & 0x7fffu); lo[2] = (int)((it.w >> 15) & 0x7fffu); }
    return (it.w & 0x80000000u) == 0;  // has a brick
}

// (the key is recomputed rather than carried over from tile_hist: this kernel waits on memory, the arithmetic is free)
template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(SORT_BLOCK) tile_scatter_kernel(const EvalParams<N> p, const u32* __restrict__ boff, T* __restrict__ rec,
                                                                  u32* __restrict__ inv, u32 nbins, i64 per_cta) {
    if (p.nwork[2] != 0) return;  // coherent batch: the plain kernel takes it
    extern __shared__ u32 s_cur[];
    for (u32 b = threadIdx.x; b < nbins; b += SORT_BLOCK) s_cur[b] = boff[(size_t)b * gridDim.x + blockIdx.x];
    __syncthreads();
    int ntot = 1, shift[N];
#pragma unroll
    for (int d = 0; d < N; d++) { ntot *= p.ntile[d]; shift[d] = __ffs(p.tile[d]) - 1; }
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    for (i64 qb = q0 + threadIdx.x; qb < q1; qb += SORT_ITER) {
        double raw[SORT_UNROLL][N];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
#pragma unroll
                for (int d = 0; d < N; d++) raw[u][d] = load_raw<N>(p, d, q);
            }
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
                T x[N];
#pragma unroll
                for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, raw[u][d]);
                const u32 slot = atomicAdd(s_cur + tile_key<T, N, ORDER, BCF>(p, x, shift, ntot), 1u);
                store_record<T, N>(rec, slot, x, (u32)q);
                if (inv) inv[q] = slot;
            }
        }
    }
}

// ---- the two sort passes for grids with more tiles than a shared-memory histogram holds (> sort_smem_bins(), e.g.
// 512^3 f64: 7.1e4 tiles): one histogram / cursor array in global memory, updated with warp-aggregated atomics
// (lanes of a warp that hit the same tile share one atomic, so a clustered batch does not serialise on one
// address).  The order of the records inside a tile then depends on scheduling; results do not (every query is
// evaluated on its own and returned to its own position).  The counting pass marks the run starts like tile_hist.
template <typename T, int N, int ORDER, int BCF, bool SCATTER>
__global__ void __launch_bounds__(SORT_BLOCK) tile_sort_global_kernel(const EvalParams<N> p, u32* __restrict__ bins, T* __restrict__ rec, u32* __restrict__ inv,
                                                                      i64 per_cta) {
    if (p.nwork[2] != 0) return;  // coherent batch: the plain kernel takes it
    int ntot = 1, shift[N];
#pragma unroll
    for (int d = 0; d < N; d++) { ntot *= p.ntile[d]; shift[d] = __ffs(p.tile[d]) - 1; }
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    const u32 lane = threadIdx.x & 31;
    for (i64 qb0 = q0; qb0 < q1; qb0 += SORT_ITER) {  // (uniform trip count: the warp votes below need every lane)
        const i64 qb = qb0 + threadIdx.x;
        double raw[SORT_UNROLL][N];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
#pragma unroll
                for (int d = 0; d < N; d++) raw[u][d] = load_raw<N>(p, d, q);
            }
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            const bool act = q < q1;
            T x[N];
            u32 key = 0xffffffffu;
            if (act) {
#pragma unroll
                for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, raw[u][d]);
                key = tile_key<T, N, ORDER, BCF>(p, x, shift, ntot);
            }
            const u32 peers = __match_any_sync(0xffffffffu, key);
            const u32 leader = __ffs(peers) - 1;
            u32 base = 0;
            if (act && lane == leader) base = atomicAdd(bins + key, (u32)__popc(peers));
            if constexpr (SCATTER) {
                base = __shfl_sync(0xffffffffu, base, leader);
                if (act) {
                    const u32 slot = base + __popc(peers & ((1u << lane) - 1));
                    store_record<T, N>(rec, slot, x, (u32)q);
                    if (inv) inv[q] = slot;
                }
            }
        }
    }
}

// ---- single-pass scatter into bins sized from the probe's sample -------------------------------------------------------
// The counting sort reads the batch twice (histogram, scatter).  The probe already holds a ~3 % sample of the tile histogram:
// tile_capacity (eval_tile.cu) gives bin b room for scale * (h_b + 4 sqrt(h_b + 1) + 4) records -- four standard
// deviations of the sampling noise above the estimate -- and one pass scatters the records with one global cursor per bin
// (warp-aggregated atomics).  A query that finds its bin full goes to an overflow list and is evaluated by the plain path
// (a few in 10^7 for smooth distributions; correct for any).  The bins are not packed any more: work items are built from
// (bin offset, records actually written) after the scatter.
// MEASURED (C2, profiles/r2_time_single_pass.txt): the histogram pass goes (0.120 + 0.038 -> 0.046 ms) but the scatter slows from
// 0.309 to 0.39 ms -- one L2 atomic per query instead of a shared-memory one, and more resident warps make it slower, not
// faster (0.41-0.49 ms at 32 warps/SM: the L2 atomic units are the limit) -- 1.266 vs 1.288 ms per step for 1.8x the record
// scratch.  Within run-to-run noise, so it is OPT-IN (GB_SORT_PASSES=1); the two-pass counting sort stays the default.
template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(SORT_BLOCK) tile_scatter_capped_kernel(const EvalParams<N> p, u32* __restrict__ cursor, const u32* __restrict__ offs,
                                                                         const u32* __restrict__ cap, T* __restrict__ rec, u32* __restrict__ over, u32* __restrict__ nover,
                                                                         i64 per_cta) {
    if (p.nwork[2] != 0) return;  // coherent batch: the plain kernel takes it
    int ntot = 1, shift[N];
#pragma unroll
    for (int d = 0; d < N; d++) { ntot *= p.ntile[d]; shift[d] = __ffs(p.tile[d]) - 1; }
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    const u32 lane = threadIdx.x & 31;
    for (i64 qb0 = q0; qb0 < q1; qb0 += SORT_ITER) {  // (uniform trip count: the warp votes below need every lane)
        const i64 qb = qb0 + threadIdx.x;
        double raw[SORT_UNROLL][N];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            if (q < q1) {
#pragma unroll
                for (int d = 0; d < N; d++) raw[u][d] = load_raw<N>(p, d, q);
            }
        }
        // all the cursor atomics of the iteration in flight before the first result is needed (their round trip is the pass's latency)
        T x[SORT_UNROLL][N];
        u32 key[SORT_UNROLL], peers[SORT_UNROLL], base[SORT_UNROLL];
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            key[u] = 0xffffffffu;
            if (q < q1) {
#pragma unroll
                for (int d = 0; d < N; d++) x[u][d] = index_coord<T, N>(p, d, raw[u][d]);
                key[u] = tile_key<T, N, ORDER, BCF>(p, x[u], shift, ntot);
            }
            peers[u] = __match_any_sync(0xffffffffu, key[u]);
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            base[u] = 0;
            if (key[u] != 0xffffffffu && lane == (u32)__ffs(peers[u]) - 1) base[u] = atomicAdd(cursor + key[u], (u32)__popc(peers[u]));
        }
#pragma unroll
        for (int u = 0; u < SORT_UNROLL; u++) {
            const i64 q = qb + u * SORT_BLOCK;
            const u32 b = __shfl_sync(0xffffffffu, base[u], __ffs(peers[u]) - 1);
            if (q < q1) {
                const u32 slot = b + __popc(peers[u] & ((1u << lane) - 1));
                if (slot < __ldg(cap + key[u])) store_record<T, N>(rec, __ldg(offs + key[u]) + slot, x[u], (u32)q);
                else over[atomicAdd(nover, 1u)] = (u32)q;  // bin full: the plain path takes the query
            }
        }
    }
}
// the queries that found their bin full: plain evaluation from the caller's coordinates, output record at the query's index
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(EVAL_BLOCK, 2) eval_overflow_kernel(const EvalParams<N> p, const u32* __restrict__ over, const u32* __restrict__ nover) {
    if (p.nwork[2] != 0) return;
    const u32 n = *nover;
    for (u32 i = blockIdx.x * EVAL_BLOCK + threadIdx.x; i < n; i += gridDim.x * EVAL_BLOCK) {
        const u32 q = over[i];
        T x[N];
        load_index_coords<T, N>(p, (i64)q, x);
        eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, x, (i64)q, (i64)q);
    }
}

// ---- brick staging --------------------------------------------------------------------------------------------------
// TMA: one cp.async.bulk.tensor per brick, issued by one thread; the box is the whole brick (its contiguous extent
// padded to 16 bytes, brick_extent), cells past the array are zero-filled by the TMA unit, completion is counted in
// bytes on an mbarrier.  Fallback for grids whose row pitch is not a multiple of 16 bytes (TMA's stride rule):
// Ampere-style element-wise cp.async with zero-fill.
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    const u32 a = (u32)__cvta_generic_to_shared(bar);
    u32 done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
template <int N> __device__ __forceinline__ void tma_load_brick(const CUtensorMap* tmap, void* dst, u64* bar, const int (&lo)[N]) {
    const u32 d = (u32)__cvta_generic_to_shared(dst), b = (u32)__cvta_generic_to_shared(bar);
    if constexpr (N == 2) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d), "l"(tmap), "r"(b),
                     "r"(lo[0]), "r"(lo[1])
                     : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d), "l"(tmap),
                     "r"(b), "r"(lo[0]), "r"(lo[1]), "r"(lo[2])
                     : "memory");
    }
}
template <int BYTES> __device__ __forceinline__ void cp_async_zfill(u32 saddr, const void* g, bool pred) {
    const int sz = pred ? BYTES : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(saddr), "l"(g), "n"(BYTES), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory"); }

// clip[d]: cells of the brick along d that exist (the brick may stick out of the array)
template <typename T, int N, int L> __device__ __forceinline__ void brick_clip(const EvalParams<N>& p, const int (&lo)[N], int (&clip)[N]) {
#pragma unroll
    for (int d = 0; d < N; d++) {
        const int full = p.tile[d] + L - 1, left = p.len[d] - lo[d];
        clip[d] = full < left ? full : left;
    }
}
// element-wise staging (vol = padded brick volume; padding cells are zero-filled like cells past the array)
template <typename T, int N, int L> __device__ __forceinline__ void brick_fill(const EvalParams<N>& p, const T* __restrict__ A, T* brick, const int (&lo)[N], int vol) {
    int clip[N];
    brick_clip<T, N, L>(p, lo, clip);
    const u32 sbase = (u32)__cvta_generic_to_shared(brick);
#pragma unroll 4
    for (int e = threadIdx.x; e < vol; e += BRICK_BLOCK) {
        int r = e;
        u32 g = 0;
        bool inside = true;
#pragma unroll
        for (int d = 0; d < N - 1; d++) {
            const int F = brick_extent(N, L, d, (int)sizeof(T));  // a constant after unrolling
            const int rd = r % F;
            r /= F;
            inside = inside && rd < clip[d];
            g += (u32)(lo[d] + rd) * (u32)p.stride[d];
        }
        inside = inside && r < clip[N - 1];
        g += (u32)(lo[N - 1] + r) * (u32)p.stride[N - 1];
        cp_async_zfill<(int)sizeof(T)>(sbase + (u32)e * (u32)sizeof(T), inside ? (const void*)(A + g) : (const void*)A, inside);
    }
}
// stage the brick of `tile` into `dst` (either mechanism); every thread calls
template <typename T, int N, int L>
__device__ __forceinline__ void stage_brick(const EvalParams<N>& p, const CUtensorMap* tmap, const T* __restrict__ A, T* dst, u64* bar, const int (&lo)[N], int vol) {
    if (p.use_tma) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (u32)vol * (u32)sizeof(T));
            tma_load_brick<N>(tmap, dst, bar, lo);
        }
    } else {
        brick_fill<T, N, L>(p, A, dst, lo, vol);
    }
}

// Deferral: append sorted slot j to the list of queries the brick kernel does not evaluate itself
// (warp-aggregated: one atomic per group of lanes deferring together).  defer[0] = count.
__device__ __forceinline__ void defer_slot(u32* __restrict__ defer, u32 j) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(defer, (u32)__popc(m));
    base = __shfl_sync(m, base, leader);
    defer[1 + base + __popc(m & ((1u << lane) - 1u))] = j;
}

// Brick offset of a query's first tap, or BRICK_DEFER if the brick path cannot take the query (taps
// that are not L consecutive cells inside the brick, invalid location); dx = fractional offsets.
constexpr u32 BRICK_DEFER = 0xffffu;
template <typename T, int N, int ORDER, int BCF>
__device__ __forceinline__ u32 brick_locate(const EvalParams<N>& p, const T (&x)[N], const int (&lo)[N], const int (&clip)[N], const bool have, T (&dx)[N]) {
    constexpr int L = ORDER + 1;
    int c[N][L];
    const bool valid = query_taps<T, N, ORDER, BCF>(p, x, c, dx);
    bool ok = valid && have;
    u32 boff = 0;
#pragma unroll
    for (int d = 0; d < N; d++) {
        const int rel = c[d][0] - lo[d];
        ok = ok && rel >= 0 && rel + L <= clip[d];
#pragma unroll
        for (int i = 1; i < L; i++) ok = ok && c[d][i] == c[d][0] + i;
        boff += (u32)rel * (u32)brick_stride(N, L, d, (int)sizeof(T));
    }
    return ok ? boff : BRICK_DEFER;
}
// one located query out of the brick: weights, contraction, outputs to slot j
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__device__ __forceinline__ void brick_eval_one(const EvalParams<N>& p, const T (&dx)[N], const i64 q, const u32 j, const T* brick, const u32 boff,
                                               u32* __restrict__ defer) {
    constexpr int L = ORDER + 1;
    constexpr int NP = num_passes(N, OUTMODE);
    T w[3][N][L];
#pragma unroll
    for (int d = 0; d < N; d++) weights<ORDER, (OUTMODE >= 1), (OUTMODE >= 2), FAST>(dx[d], w[0][d], w[1][d], w[2][d]);
    BrickTaps<T, N, L> src;
    src.b = brick + boff;
    T acc[NP];
    contract<T, N, L, OUTMODE, FAST>(src, w, acc);
    if (nonfinite(acc[0])) {  // a NaN/Inf coefficient may have been touched (see ExactRec): literal path, later
        defer_slot(defer, j);
        return;
    }
    store_outputs<T, N, OUTMODE>(p, p.by_query ? q : (i64)j, q, true, acc);
}

// the deferred queries: plain evaluation (eval_one) of the slots listed
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(EVAL_BLOCK, 2) eval_deferred_kernel(const EvalParams<N> p, const T* __restrict__ rec, const u32* __restrict__ defer) {
    if (p.nwork[2] != 0) return;  // coherent batch: the plain kernel took it
    const u32 n = defer[0];
    for (u32 i = blockIdx.x * EVAL_BLOCK + threadIdx.x; i < n; i += gridDim.x * EVAL_BLOCK) {
        T x[N];
        u32 q;
        const u32 j = defer[1 + i];
        load_record<T, N>(rec, j, x, q);
        eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, x, p.by_query ? (i64)q : (i64)j, (i64)q);
    }
}

// Shared-memory bank classes.  A warp's 64-bit shared loads are served per half-warp, 16 lanes x 8
// bytes = the 32 banks once.  Within an item the queries sit at random brick offsets, so a tap load
// (same compile-time offset from every lane's first tap) would hit random banks -- measured 2.9
// wavefronts per half-warp instead of 1, and the LSU pipe at 93 %.  Therefore lane t only evaluates
// queries whose first-tap offset is congruent to t modulo 16: every tap load of a half-warp then
// touches 16 different bank pairs.  Per item the records are bucketed by class (two short passes
// over the item with shared-memory counters); class c is served by the 16 threads t = c (mod 16).
// Classes are never equally full: each thread runs T = ceil(n/256) rounds, and a thread whose class
// runs dry pulls from the overflow of fuller classes (those few loads may conflict).
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(BRICK_BLOCK, 2)
    eval_brick_kernel(const EvalParams<N> p, const T* __restrict__ rec, u32* __restrict__ defer, const __grid_constant__ CUtensorMap tmap) {
    constexpr int L = ORDER + 1;
    // Bank-class binning pays where the shared-memory pipe is the limiter (FAST: 64 loads per ~270
    // FP64 operations); EXACT is bound by the FP64 pipe (816 dependent multiply/add per query) and
    // keeps the direct schedule, which has fewer barriers per item.
    constexpr bool BINNED = FAST;
    if (p.nwork[2] != 0) return;  // coherent batch (coherence_probe_kernel): the plain kernel takes it
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) u64 s_bar[2];
    __shared__ u32 s_it[2];
    __shared__ u32 s_cnt[2][16], s_cur[2][16], s_opull[2], s_off[17], s_ooff[17];  // counters: one set per item parity
    __shared__ unsigned short s_boff[BINNED ? EVAL_CHUNK : 1], s_idx[BINNED ? EVAL_CHUNK : 1];  // bank-class lists (FAST only)
    T* bricks = reinterpret_cast<T*>(smem_raw);
    const T* __restrict__ A = (const T*)p.coefs;
    const int vol = brick_stride(N, L, N - 1, (int)sizeof(T)) * (p.tile[N - 1] + L - 1);
    const int volp = (vol + 31) & ~31;  // buffers stay 128-byte aligned (TMA destination)
    const u32 nwork = p.nwork[0];
    u32* ctr = const_cast<u32*>(p.nwork) + 1;
    const u32* __restrict__ work = p.work;
    const u32 tid = threadIdx.x;
    const bool tma = p.use_tma != 0;
    // items differ in cost: CTAs pull them from a shared counter, one item ahead of the one being
    // staged (thread 0 keeps the next ticket in a register, so the atomic's latency is hidden)
    u32 pend = 0;
    if (tid == 0) {
        s_it[0] = atomicAdd(ctr, 1u);
        pend = atomicAdd(ctr, 1u);
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 16) { s_cnt[0][tid] = 0; s_cur[0][tid] = 0; }
    if (tid == 16) s_opull[0] = 0;
    __syncthreads();
    u32 cur = s_it[0];
    u32 beg_c = 0, end_c = 0;
    int lo_c[N];
    bool have_c = false;
#pragma unroll
    for (int d = 0; d < N; d++) lo_c[d] = 0;
    if (cur < nwork) {
        const uint4 it = reinterpret_cast<const uint4*>(work)[cur];
        beg_c = it.y; end_c = it.z;
        have_c = item_corner<N>(it, lo_c);
        if (have_c) stage_brick<T, N, L>(p, &tmap, A, bricks, &s_bar[0], lo_c, vol);
    }
    cp_async_commit();
    u32 phase[2] = {0, 0};  // mbarrier parity of each buffer's next completion (TMA)
    T xn[N];  // direct schedule: the prefetched query
    u32 qn = 0;
#pragma unroll
    for (int d = 0; d < N; d++) xn[d] = (T)0;
    for (int k = 0; cur < nwork; k++) {
        const int par = k & 1;
        if (tid == 0) {
            s_it[par ^ 1] = pend;
            pend = atomicAdd(ctr, 1u);
        }
        __syncthreads();  // ticket visible; every thread is done with item k-1, whose buffer is refilled now
        if (tid < 16) { s_cnt[par ^ 1][tid] = 0; s_cur[par ^ 1][tid] = 0; }  // the next item's counters (idle since item k-1)
        if (tid == 16) s_opull[par ^ 1] = 0;
        const u32 nxt = s_it[par ^ 1];
        u32 beg_n = 0, end_n = 0;
        int lo_n[N];
        bool have_n = false;
#pragma unroll
        for (int d = 0; d < N; d++) lo_n[d] = 0;
        if (nxt < nwork) {
            const uint4 it = reinterpret_cast<const uint4*>(work)[nxt];
            beg_n = it.y; end_n = it.z;
            have_n = item_corner<N>(it, lo_n);
            if (have_n) stage_brick<T, N, L>(p, &tmap, A, bricks + (size_t)(par ^ 1) * volp, &s_bar[par ^ 1], lo_n, vol);
        }
        cp_async_commit();
        int lo[N], clip[N];
        const bool have = have_c;  // items without a brick (invalid locations, chunks that do not fit one): every query is deferred
#pragma unroll
        for (int d = 0; d < N; d++) lo[d] = lo_c[d];
        brick_clip<T, N, L>(p, lo, clip);
        const u32 n = end_c - beg_c;
        // the brick of item k has landed: this thread's cp.async groups / the buffer's mbarrier phase
        auto brick_ready = [&]() {
            if (tma) {
                if (have) { mbar_wait(&s_bar[par], phase[par]); phase[par] ^= 1u; }
            } else {
                cp_async_wait<1>();
            }
        };
        if constexpr (BINNED) {
            // ---- pass 1 over the item's queries: locate, count the bank classes (the brick is not needed yet)
            for (u32 r0 = tid; r0 < n; r0 += BRICK_BLOCK * PASS1_UNROLL) {
                T x[PASS1_UNROLL][N];
                u32 q;
#pragma unroll
                for (int u = 0; u < PASS1_UNROLL; u++)  // the loads of PASS1_UNROLL records in flight together
                    if (r0 + u * BRICK_BLOCK < n) load_record<T, N>(rec, beg_c + r0 + u * BRICK_BLOCK, x[u], q);
#pragma unroll
                for (int u = 0; u < PASS1_UNROLL; u++) {
                    const u32 r = r0 + u * BRICK_BLOCK;
                    if (r < n) {
                        T dxu[N];
                        const u32 boff = brick_locate<T, N, ORDER, BCF>(p, x[u], lo, clip, have, dxu);
                        s_boff[r] = (unsigned short)boff;
                        if (boff == BRICK_DEFER) defer_slot(defer, beg_c + r);  // global-memory path, later
                        else atomicAdd(&s_cnt[par][boff & 15u], 1u);
                    }
                }
            }
            __syncthreads();
            if (tid < 32) {  // exclusive scans of the class counts and of the class overflows
                const u32 cnt = tid < 16 ? s_cnt[par][tid] : 0u;
                u32 incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)tid >= o) incl += t;
                }
                const u32 n_ok = __shfl_sync(0xffffffffu, incl, 15);
                const u32 cap = 16u * ((n_ok + BRICK_BLOCK - 1) / BRICK_BLOCK);
                const u32 over = cnt > cap ? cnt - cap : 0u;
                u32 oincl = over;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 t = __shfl_up_sync(0xffffffffu, oincl, o);
                    if ((int)tid >= o) oincl += t;
                }
                if (tid <= 16) { s_off[tid] = incl - cnt; s_ooff[tid] = oincl - over; }
            }
            __syncthreads();
            // ---- pass 2: class-sorted index list
            for (u32 r = tid; r < n; r += BRICK_BLOCK) {
                const u32 b = s_boff[r];
                if (b != BRICK_DEFER) {
                    const u32 cl = b & 15u;
                    s_idx[s_off[cl] + atomicAdd(&s_cur[par][cl], 1u)] = (unsigned short)r;
                }
            }
            brick_ready();
            __syncthreads();     // everybody's copies have landed; the index list is complete
            // ---- evaluation: T rounds, thread t serves class t mod 16
            const T* brick = bricks + (size_t)par * volp;
            const u32 cls = tid & 15u, rank = tid >> 4;
            const u32 n_ok = s_off[16], rounds = (n_ok + BRICK_BLOCK - 1) / BRICK_BLOCK, cap = 16u * rounds;
            const u32 mine = min(s_cnt[par][cls], cap), base = s_off[cls], n_over = s_ooff[16];
            // next_record: index of this thread's record for round i (own class, else overflow of a fuller class)
            auto next_record = [&](u32 i, u32& r) -> bool {
                if (i >= rounds) return false;
                const u32 e = rank + 16u * i;
                if (e < mine) {
                    r = s_idx[base + e];
                    return true;
                }
                const u32 o = atomicAdd(&s_opull[par], 1u);  // my class ran dry
                if (o >= n_over) return false;
                u32 cl = 0;
                while (o >= s_ooff[cl + 1]) cl++;
                r = s_idx[s_off[cl] + cap + (o - s_ooff[cl])];
                return true;
            };
            u32 rn = 0;
            bool more = next_record(0, rn);
            if (more) load_record<T, N>(rec, beg_c + rn, xn, qn);
            for (u32 i = 0; more; i++) {
                T x[N];
#pragma unroll
                for (int d = 0; d < N; d++) x[d] = xn[d];
                const u32 q = qn, r = rn;
                more = next_record(i + 1, rn);
                if (more) load_record<T, N>(rec, beg_c + rn, xn, qn);  // in flight during this round's evaluation
                int cc[N][L];
                T dx[N];
                query_taps<T, N, ORDER, BCF>(p, x, cc, dx);
                brick_eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, dx, (i64)q, beg_c + r, brick, (u32)s_boff[r], defer);
            }
        } else {
            // ---- direct: thread t takes queries t, t+256, ... of the item; the next one is in flight
            // while the current one is evaluated (across the item boundary too)
            brick_ready();
            __syncthreads();     // ... and everybody else's copies
            const T* brick = bricks + (size_t)par * volp;
            const bool next_has = nxt < nwork && beg_n + tid < end_n;
            u32 j = beg_c + tid;
            if (k == 0 && j < end_c) load_record<T, N>(rec, j, xn, qn);
            if (j >= end_c) {
                if (next_has) load_record<T, N>(rec, beg_n + tid, xn, qn);
            } else {
                while (true) {
                    T x[N];
#pragma unroll
                    for (int d = 0; d < N; d++) x[d] = xn[d];
                    const u32 q = qn, jc = j;
                    j += BRICK_BLOCK;
                    const bool more = j < end_c;
                    if (more) load_record<T, N>(rec, j, xn, qn);
                    else if (next_has) load_record<T, N>(rec, beg_n + tid, xn, qn);
                    T dx[N];
                    const u32 boff = brick_locate<T, N, ORDER, BCF>(p, x, lo, clip, have, dx);
                    if (boff == BRICK_DEFER) defer_slot(defer, jc);  // global-memory path, later
                    else brick_eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, dx, (i64)q, jc, brick, boff, defer);
                    if (!more) break;
                }
            }
        }
        cur = nxt; have_c = have_n; beg_c = beg_n; end_c = end_n;
#pragma unroll
        for (int d = 0; d < N; d++) lo_c[d] = lo_n[d];
    }
    cp_async_wait<0>();
}

// caller's order <- sorted slots: thread q gathers its slot's output record (one 32-byte sector for
// f64 valgrad in 3-D) and writes the caller's val / grad / hess arrays coalesced.
template <typename T, int N, int OUTMODE>
__global__ void __launch_bounds__(256) unpermute_kernel(const T* __restrict__ osorted, const u32* __restrict__ inv, T* __restrict__ val, T* __restrict__ grad,
                                                        T* __restrict__ hess, i64 nq, const u32* __restrict__ nwork) {
    if (nwork[2] != 0) return;  // coherent batch: the plain kernel wrote the outputs in place
    constexpr int W = out_words<N>(OUTMODE);
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (i64)gridDim.x * blockDim.x) {
        const T* src = osorted + (size_t)(inv ? inv[q] : (u32)q) * W;  // (inv == NULL: the records are in query order already)
        alignas(16) T o[W];
        if constexpr (W * sizeof(T) == 32) {
            reinterpret_cast<float4*>(o)[0] = __ldcs(reinterpret_cast<const float4*>(src));
            reinterpret_cast<float4*>(o)[1] = __ldcs(reinterpret_cast<const float4*>(src) + 1);
        } else if constexpr (W * sizeof(T) == 16) {
            reinterpret_cast<float4*>(o)[0] = __ldcs(reinterpret_cast<const float4*>(src));
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) o[k] = __ldcs(src + k);
        }
        if (val) val[q] = o[0];
        if constexpr (OUTMODE >= 1) {
            if (grad) {
#pragma unroll
                for (int d = 0; d < N; d++) grad[q * N + d] = o[1 + d];
            }
        }
        if constexpr (OUTMODE >= 2) {
            if (hess) {
#pragma unroll
                for (int k = 0; k < N * N; k++) hess[q * (N * N) + k] = o[1 + N + k];
            }
        }
    }
}

// eval_tile.cu
struct TileGeo { int N, tile[MAXD], ntile[MAXD]; };  // for the scan's work items (tile index -> brick corner)
int tile_offsets(const u32* bh, u32 nbins, u32 nctas, u32* hist, u32* offs, u32* boff, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s);
int tile_offsets_global(const u32* hist, u32 nbins, u32* offs, u32* cursor, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s);
bool choose_tile(int N, int L, size_t esz, const int* len, i64 nq, bool forced, int* tile, int* ntile);
u32 sort_smem_bins();  // 12000: per-CTA histograms live in shared memory (<= 48 KB); more tiles: tile_sort_global_kernel
int tile_capacity(const u32* shist, u32 nbins, const u32* nwork, i64 nq, u32* offs, u32* cap, u32* cursor, cudaStream_t s);
int tile_items(const u32* cursor, const u32* cap, const u32* offs, u32 nbins, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s);
size_t tile_capacity_bound(i64 nq, i64 nsample, u32 nbins);  // records the bins of tile_capacity can add up to
bool sort_single_pass();
u32 coherent_need_percent(bool hinted);  // share (percent) of the probed leaves that must be coherent for the plain path to take the batch; > 100: never
int make_brick_tensor_map(CUtensorMap* m, int dtype, int N, const void* base, const int* len, const i64* stride, const int* box);
// Larger batches are evaluated in slices: bounds the record scratch, and keeps the randomly written regions -- the
// slice's records and its outputs -- from outgrowing L2 by too much: the two permutations of a 4e7-query batch
// cost 3.13 ms in one piece, 4 x 0.55 ms in four (256^3 f64: 6.11 -> 5.41 ms).  Slices must stay large against the
// number of tiles, though (every slice stages every occupied brick again: 512^3, 7.1e4 tiles, 4e7 queries: 6.96 ms
// in one piece, 8.69 ms in four), and slicing a 1e7 batch loses.  Hence: equal slices of at most
// clamp(1024 queries per tile, 2^24, 2^26) queries; GB_TILED_BATCH overrides the limit.
inline i64 tiled_max_batch(i64 ntiles = -1) {
    static const i64 env = [] {
        const char* e = getenv("GB_TILED_BATCH");
        const long long x = e ? atoll(e) : 0;
        return (i64)(x >= 1024 ? x : 0);
    }();
    if (env) return env;
    if (ntiles < 0) return (i64)1 << 26;
    return std::max<i64>((i64)1 << 24, std::min<i64>((i64)1 << 26, ntiles * 1024));
}

// plain launch geometry, shared by launch_one and the coherent branch of launch_tiled
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST> int plain_grid(i64 nq, int* grid) {
    static int blocks_plain = 0;  // benign race: every thread computes the same value
    if (blocks_plain == 0) {
        int b = 0;
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, eval_kernel<T, N, ORDER, BCF, OUTMODE, FAST>, EVAL_BLOCK, 0));
        blocks_plain = b > 0 ? b : 1;
    }
    const i64 need = (nq + EVAL_BLOCK - 1) / EVAL_BLOCK;
    // one resident wave for small batches (measured: 1e6 queries 43 vs 52 us), a few waves otherwise; CTAs stride over the blocks
    const i64 cap = (i64)num_sms() * blocks_plain * (nq < (1ll << 22) ? 1 : 4);
    *grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    return GB_OK;
}

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
int launch_tiled(const EvalParams<N>& p_in, int flags, cudaStream_t s) {
    constexpr int L = ORDER + 1;
    auto kern = eval_brick_kernel<T, N, ORDER, BCF, OUTMODE, FAST>;
    EvalParams<N> p = p_in;
    u32 nbins = 1;
    for (int d = 0; d < N; d++) nbins *= (u32)p.ntile[d];
    nbins += 1;  // + the bin of invalid queries
    int box[N];
    for (int d = 0; d < N; d++) box[d] = d < N - 1 ? brick_extent(N, L, d, (int)sizeof(T)) : p.tile[d] + L - 1;
    size_t vol = 1;
    for (int d = 0; d < N; d++) vol *= (size_t)box[d];
    const size_t smem = 2 * ((vol + 31) & ~(size_t)31) * sizeof(T);
    const u32 nctas = (u32)std::max<i64>(1, std::min<i64>((i64)num_sms() * 2, (p.nq + 4095) / 4096));
    const i64 per_cta = ((p.nq + nctas - 1) / nctas + SORT_BLOCK - 1) / SORT_BLOCK * SORT_BLOCK;
    const size_t max_items = (size_t)nbins + (size_t)(p.nq / EVAL_CHUNK) + 2;
    const size_t pitch = ((size_t)p.nq + 15) & ~(size_t)15;
    // one scratch allocation: records | sorted outputs | inv | deferred-slot list (`keys`) | work | bh | boff | hist | offs | nwork
    constexpr int W = out_words<N>(OUTMODE);
    const bool global_bins = nbins > sort_smem_bins();
    const size_t nbc = global_bins ? 16 : (size_t)nbins * nctas;
    // Where do the brick kernel's output records go?  To the query's sorted slot (coalesced writes, then un-permuted by a random
    // 32-byte gather) or straight to record q (random full-sector writes hidden under the FP64-bound evaluation, then a
    // streaming unpack).  GB_OUT_BY_QUERY selects (measured, profiles/README.md).
    static const int by_query_env = [] { const char* e = getenv("GB_OUT_BY_QUERY"); return e ? atoi(e) : GB_OUT_BY_QUERY_DEFAULT; }();
    const bool packed = p_in.aos != 0;  // GB_OUT_PACKED: the caller's buffer IS the array of output records, in query order
    const int by_query = packed ? 1 : by_query_env;
    // One pass over the batch instead of two (tile_scatter_capped_kernel): bins sized from the probe's sample.
    const bool single = sort_single_pass() && !global_bins && by_query != 0;
    size_t nrec = pitch;
    if (single) {  // the probe's sample size, as the kernel will count it
        i64 ns = 0;
        for (u32 c = 0; c < nctas; c++) {
            const i64 q0 = (i64)c * per_cta, q1 = std::min<i64>(q0 + per_cta, p.nq);
            if (q0 >= q1) continue;
            const i64 stride = ((q1 - q0) / SORT_UNROLL) & ~(i64)(SORT_BLOCK - 1);
            for (int u = 0; u < SORT_UNROLL; u++)
                if (u == 0 || stride > 0) ns += std::max<i64>(0, std::min<i64>(SORT_BLOCK, q1 - (q0 + u * stride)));
        }
        nrec = (tile_capacity_bound(p.nq, ns, nbins) + 15) & ~(size_t)15;
        if (nrec >= (size_t)0xfffffff0u) return fail(GB_ERR_ARG, "tiled evaluation: batch slice too large");
    }
    const size_t rec_bytes = (nrec * (N + 1) * sizeof(T) + 31) & ~(size_t)31;
    const size_t out_bytes = (pitch * W * sizeof(T) + 31) & ~(size_t)31;
    const size_t words = (pitch + 16) + 4 * max_items + 2 * nbc + 2 * (size_t)(nbins + 4) + 8;
    AsyncBuf guard(s);  // released on every exit path
    GB_CUDA(guard.alloc(rec_bytes + out_bytes + (pitch + words) * sizeof(u32)));
    unsigned char* scratch = (unsigned char*)guard.p;
    T* rec = reinterpret_cast<T*>(scratch);
    T* osorted = reinterpret_cast<T*>(scratch + rec_bytes);
    u32* inv = reinterpret_cast<u32*>(scratch + rec_bytes + out_bytes);
    u32* keys = inv + pitch;
    u32* work = keys + pitch + 16;  // (16-byte aligned: items are uint4, see tile_item)
    u32* bh = work + 4 * max_items;
    u32* boff = bh + nbc;
    u32* hist = boff + nbc;
    u32* offs = hist + (nbins + 4);
    u32* nwork = offs + (nbins + 4);  // [0] items, [1] ticket counter, [2] coherent flag, [3..5] the probe's counters, [6] queries sampled, [7] overflow count
    TileGeo geo;
    geo.N = N;
    for (int d = 0; d < MAXD; d++) { geo.tile[d] = d < N ? p.tile[d] : 1; geo.ntile[d] = d < N ? p.ntile[d] : 1; }
    const size_t hsmem = (size_t)nbins * sizeof(u32);
    // brick staging: TMA when the array obeys its stride rule (every stride but the first a multiple of 16 bytes)
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    static const bool tma_off = [] { const char* e = getenv("GB_NO_TMA"); return e && *e && *e != '0'; }();
    p.use_tma = 0;
    if (!tma_off && make_brick_tensor_map(&tmap, dtype_of<T>::value, N, p.coefs, p.len, p.stride, box) == GB_OK) p.use_tma = 1;
    p.work = work;
    p.nwork = nwork;
    p.by_query = by_query;
    u32* over = inv;  // (single pass: the overflow list lives where the inverse permutation would)
    if (by_query) inv = nullptr;
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device, cheap
    int bt = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt, kern, BRICK_BLOCK, smem));
    if (bt < 1) return fail(GB_ERR_CUDA, "tiled evaluation: brick does not fit in shared memory");
    // Is the batch coherent as it arrived?  Then the plain kernel at the end takes it and everything in between exits at once.
    // (GB_TILED alone forces the bins: no probe.)
    const u32 need = coherent_need_percent((flags & GB_SORTED) != 0);
    const bool probe = (!(flags & GB_TILED) || (flags & GB_SORTED)) && need > 0 && need <= 100;  // (GB_TILED | GB_SORTED: forced onto this path, probe on)
    prof_mark(0, s);
    GB_CUDA(cudaMemsetAsync(nwork, 0, 8 * sizeof(u32), s));
    if (single) GB_CUDA(cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(u32), s));
    if (probe || single) {  // (single pass without the probe's verdict: a share no batch reaches)
        u32 vmax = 4;
        for (int d = 0; d < N; d++) vmax *= (u32)p.tile[d];
        coherence_probe_kernel<T, N, ORDER, BCF><<<nctas, SORT_BLOCK, 0, s>>>(p, nwork, per_cta, vmax, probe ? need : 101u, single ? hist : nullptr);
        count_launch();
    }
    int rc;
    if (single) {  // bh: bin capacities, boff: write cursors
        prof_mark(1, s);
        rc = tile_capacity(hist, nbins, nwork, p.nq, offs, bh, boff, s);
        if (rc) return rc;
        prof_mark(2, s);
        // (no per-CTA state: more, smaller CTAs than the two-pass sort -- GB_SCATTER_CTAS per SM -- to cover the cursor round trips)
        static const int sc_per_sm = [] { const char* e = getenv("GB_SCATTER_CTAS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 32 ? v : 16; }();
        const u32 sc_ctas = (u32)std::max<i64>(1, std::min<i64>((i64)num_sms() * sc_per_sm, (p.nq + SORT_ITER - 1) / SORT_ITER));
        const i64 sc_per = ((p.nq + sc_ctas - 1) / sc_ctas + SORT_BLOCK - 1) / SORT_BLOCK * SORT_BLOCK;
        tile_scatter_capped_kernel<T, N, ORDER, BCF><<<sc_ctas, SORT_BLOCK, 0, s>>>(p, boff, offs, bh, rec, over, nwork + 7, sc_per);
        count_launch();
        rc = tile_items(boff, bh, offs, nbins, work, nwork, EVAL_CHUNK, geo, s);
        if (rc) return rc;
    } else if (global_bins) {  // hist doubles as the write cursors after the scan
        GB_CUDA(cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(u32), s));
        tile_sort_global_kernel<T, N, ORDER, BCF, false><<<nctas, SORT_BLOCK, 0, s>>>(p, hist, rec, inv, per_cta);
        count_launch();
        prof_mark(1, s);
        rc = tile_offsets_global(hist, nbins, offs, hist, work, nwork, EVAL_CHUNK, geo, s);
        if (rc) return rc;
        prof_mark(2, s);
        tile_sort_global_kernel<T, N, ORDER, BCF, true><<<nctas, SORT_BLOCK, 0, s>>>(p, hist, rec, inv, per_cta);
        count_launch();
    } else {
        tile_hist_kernel<T, N, ORDER, BCF><<<nctas, SORT_BLOCK, hsmem, s>>>(p, bh, nbins, per_cta);
        count_launch();
        prof_mark(1, s);
        rc = tile_offsets(bh, nbins, nctas, hist, offs, boff, work, nwork, EVAL_CHUNK, geo, s);
        if (rc) return rc;
        prof_mark(2, s);
        // (write cursors shared by a 4-CTA cluster through distributed shared memory -- four times longer runs per tile -- were
        // built and measured SLOWER: 0.36 vs 0.31 ms, the remote atomics' latency; profiles/README.md)
        tile_scatter_kernel<T, N, ORDER, BCF><<<nctas, SORT_BLOCK, hsmem, s>>>(p, boff, rec, inv, nbins, per_cta);
        count_launch();
    }
    GB_CUDA(cudaGetLastError());
    const i64 grid = std::min<i64>((i64)max_items, (i64)num_sms() * bt);
    // the deferred-slot list (count first)
    u32* defer = keys;
    GB_CUDA(cudaMemsetAsync(defer, 0, sizeof(u32), s));
    EvalParams<N> pp = p;  // the plain kernel's view: the caller's order and arrays (or packed records)
    pp.aos = p_in.aos;
    p.aos = W;  // both brick-path kernels write one output record per slot / query ...
    p.val = packed ? p_in.val : (void*)osorted;
    p.grad = nullptr;
    p.hess = nullptr;
    prof_mark(3, s);
    kern<<<(int)grid, BRICK_BLOCK, smem, s>>>(p, rec, defer, tmap);
    prof_mark(4, s);
    eval_deferred_kernel<T, N, ORDER, BCF, OUTMODE, FAST><<<num_sms() * 2, EVAL_BLOCK, 0, s>>>(p, rec, defer);
    if (single) {  // the queries that found their bin full
        eval_overflow_kernel<T, N, ORDER, BCF, OUTMODE, FAST><<<num_sms() * 2, EVAL_BLOCK, 0, s>>>(p, over, nwork + 7);
        count_launch();
    }
    prof_mark(5, s);
    // ... and the records go back to the caller's order and arrays
    const int ugrid = (int)std::min<i64>((p.nq + 255) / 256, (i64)num_sms() * 32);
    if (!packed) unpermute_kernel<T, N, OUTMODE><<<ugrid, 256, 0, s>>>(osorted, inv, (T*)p_in.val, (T*)p_in.grad, (T*)p_in.hess, p.nq, nwork);
    count_launch(packed ? 2 : 3);
    if (probe) {  // coherent batch: evaluated in place by the plain kernel (exits at once otherwise)
        int pgrid = 1;
        rc = plain_grid<T, N, ORDER, BCF, OUTMODE, FAST>(p.nq, &pgrid);
        if (rc) return rc;
        eval_kernel<T, N, ORDER, BCF, OUTMODE, FAST><<<pgrid, EVAL_BLOCK, 0, s>>>(pp, nwork + 2);
        count_launch();
    }
    prof_mode(nwork, s);
    prof_mark(6, s);
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
int launch_one(const EvalParams<N>& p_in, int flags, cudaStream_t s) {
    constexpr int L = ORDER + 1;
    constexpr bool BRICK = brick_compiled(N, ORDER);
    auto kern = eval_kernel<T, N, ORDER, BCF, OUTMODE, FAST>;
    EvalParams<N> p = p_in;
    const size_t grid_bytes = (size_t)p.total * sizeof(T);
    const bool forced = (flags & GB_TILED) != 0;
    bool tiled = forced;
    // Heuristic (measured, profiles/README.md): binning costs two random 32-byte-class permutations of the
    // batch (records out, results back), which only pays when a query's neighbourhood is heavy -- 3-D cubic
    // f64 (64 taps, 512 B): tiled 1.4 ms vs plain 3 ms per 1e7; 2-D quadratic f32 (36 B) and 4-D linear f64
    // (128 B): plain is 3-5x faster -- and the grid is too large to be served out of L2.
    size_t tap_bytes = sizeof(T);
    for (int d = 0; d < N; d++) tap_bytes *= (size_t)L;
    if (!(flags & (GB_TILED | GB_PLAIN))) tiled = N >= 2 && tap_bytes >= 256 && p.nq >= (1ll << 20) && grid_bytes >= (48u << 20);
    if (p.nchan > 1 || p.blocked || !BRICK) tiled = false;  // multi-valued / blocked grids, light neighbourhoods: plain path
    if (N == 3 && (p.len[1] > 0x7fff || p.len[2] > 0x7fff)) tiled = false;  // (work items pack the brick corner: make_item)
    if (tiled) tiled = choose_tile(N, L, sizeof(T), p.len, std::min<i64>(p.nq, tiled_max_batch()), forced, p.tile, p.ntile);
    p.use_tma = 0;  // (p.aos stays what the caller set: the words of a GB_OUT_PACKED record, 0 for separate arrays)
    p.by_query = 0;
    if (!tiled) {
        for (int d = 0; d < N; d++) { p.tile[d] = 1; p.ntile[d] = 1; }
        int grid = 1;
        const int rc = plain_grid<T, N, ORDER, BCF, OUTMODE, FAST>(p.nq, &grid);
        if (rc) return rc;
        kern<<<grid, EVAL_BLOCK, 0, s>>>(p, nullptr);
        count_launch();
        GB_CUDA(cudaGetLastError());
        return GB_OK;
    }
    if constexpr (BRICK) {
        // equal slices of at most tiled_max_batch() queries, one after the other on the stream
        i64 ntiles = 1;
        for (int d = 0; d < N; d++) ntiles *= p.ntile[d];
        const i64 limit = tiled_max_batch(ntiles), nslices = (p_in.nq + limit - 1) / limit;
        const i64 slice = std::min<i64>(limit, (((p_in.nq + nslices - 1) / nslices) + 255) & ~(i64)255);
        const size_t xs = p.xdtype == GB_F64 ? 8 : 4;
        for (i64 q0 = 0; q0 < p_in.nq; q0 += slice) {
            EvalParams<N> sp = p;
            sp.nq = std::min<i64>(slice, p_in.nq - q0);
            sp.q0 = p_in.q0 + q0;
            for (int d = 0; d < N; d++) sp.xp[d] = (const char*)p.xp[d] + (size_t)q0 * (size_t)p.xqs * xs;
            if (p.val) sp.val = (T*)p.val + q0 * (p.aos ? p.aos : 1);
            if (p.grad) sp.grad = (T*)p.grad + q0 * N;
            if (p.hess) sp.hess = (T*)p.hess + q0 * N * N;
            const int rc = launch_tiled<T, N, ORDER, BCF, OUTMODE, FAST>(sp, flags, s);
            if (rc) return rc;
        }
    }
    return GB_OK;
}

// ====================================================================================================
// Tensor-product getindex(G, xs, ys, ...) (src/interp.jl:273-306): out[i1,i2,...] = G[xs[i1], ys[i2], ...].
// Separable set-up: the tap coordinates, the wrap flag and the value weights of an axis point depend on
// that axis only, so they are computed once per axis point (outer_table_kernel: O(sum n_d) coordinate
// work instead of O(prod n_d)) and the main kernel only combines table rows and contracts -- in the
// reference's order, so the result is bit-identical to evaluating the expanded query list.
template <int N> struct OuterTables {
    i64 lens[N];
    const int* tap[N];           // [n_d][L] 0-based tap coordinates
    const void* w[N];            // [n_d][L] value weights (T)
    const unsigned char* flg[N];  // bit 0 valid, bit 1 iswrap
};
template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(256) outer_table_kernel(const EvalParams<N> p, int d, i64 n, int* __restrict__ tap, T* __restrict__ w, unsigned char* __restrict__ flg) {
    constexpr int L = ORDER + 1;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T x = index_coord<T, N>(p, d, load_raw<N>(p, d, i));
        int c[L];
        T dx;
        bool iswrap;
        const bool valid = dim_taps<T, ORDER, BCF>(x, p.len[d], c, dx, iswrap);
        T wv[L], g[L], h[L];
        weights<ORDER, false, false, false>(dx, wv, g, h);
#pragma unroll
        for (int k = 0; k < L; k++) { tap[i * L + k] = c[k]; w[i * L + k] = wv[k]; }
        flg[i] = (unsigned char)((valid ? 1 : 0) | (iswrap ? 2 : 0));
    }
}
template <typename T, int N, int ORDER, int BCF, bool FAST>
__global__ void __launch_bounds__(256, 2) outer_eval_kernel(const EvalParams<N> p, const OuterTables<N> t, T* __restrict__ out, i64 nq) {
    constexpr int L = ORDER + 1;
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (i64)gridDim.x * blockDim.x) {
        i64 rem = q, idx[N];
        bool valid = true, anywrap = false;
        int c[N][L];
        T w[3][N][L];
#pragma unroll
        for (int d = 0; d < N; d++) {
            idx[d] = rem % t.lens[d];
            rem /= t.lens[d];
            const unsigned char f = __ldg(t.flg[d] + idx[d]);
            valid = valid && (f & 1);
            anywrap = anywrap || (f & 2);
#pragma unroll
            for (int k = 0; k < L; k++) {
                c[d][k] = __ldg(t.tap[d] + idx[d] * L + k);
                w[0][d][k] = __ldg((const T*)t.w[d] + idx[d] * L + k);
            }
        }
        if constexpr (ORDER == GB_LINEAR && BCF != BCF_NIL) {  // the offset-table rule of query_taps
            if (!anywrap) {
#pragma unroll
                for (int d = 0; d < N; d++) c[d][1] = c[d][0] + 1;
            }
        }
        GlobalTaps<T, N, L, GUARD> src;
        src.A = (const T*)p.coefs;
        src.total = p.blocked ? 0xffffffffu : (u32)p.total;
        bool slow = false;
        if constexpr (GUARD) slow = blocked_carry<N, L>(p, c);
        T acc[1];
        if (!slow) {
#pragma unroll
            for (int d = 0; d < N; d++) {
#pragma unroll
                for (int k = 0; k < L; k++) src.off[d][k] = tap_offset<N>(p, d, c[d][k]);
            }
            contract<T, N, L, 0, FAST>(src, w, acc);
            slow = nonfinite(acc[0]);
        }
        if (slow) {  // literal zero-weight semantics, from the coordinates
            T x[N];
#pragma unroll
            for (int d = 0; d < N; d++) x[d] = index_coord<T, N>(p, d, load_raw<N>(p, d, idx[d]));
            EvalParams<N> ps = p;
            ps.val = out;
            ps.grad = nullptr;
            ps.hess = nullptr;
            eval_one_slow<T, N, ORDER, BCF>(ps, src.A, x, q, q, 0);
            continue;
        }
        if (!valid) {
            acc[0] = qnan<T>();
            if (p.first_invalid) atomicMin(p.first_invalid, (u64)q);
        }
        out[q] = acc[0];
    }
}
template <typename T, int N, int ORDER, int BCF>
int launch_outer_bc(const EvalParams<N>& p, const i64* lens, int flags, void* out, cudaStream_t s) {
    constexpr int L = ORDER + 1;
    OuterTables<N> t;
    i64 rows = 0, nq = 1;
    for (int d = 0; d < N; d++) { t.lens[d] = lens[d]; rows += lens[d]; nq *= lens[d]; }
    if (nq == 0) return GB_OK;
    // one scratch block: taps (int) | weights (T) | flags
    const size_t tap_b = ((size_t)rows * L * sizeof(int) + 15) & ~(size_t)15, w_b = ((size_t)rows * L * sizeof(T) + 15) & ~(size_t)15;
    AsyncBuf scratch(s);
    GB_CUDA(scratch.alloc(tap_b + w_b + (size_t)rows + 16));
    unsigned char* blk = (unsigned char*)scratch.p;
    int* tap = reinterpret_cast<int*>(blk);
    T* w = reinterpret_cast<T*>(blk + tap_b);
    unsigned char* flg = blk + tap_b + w_b;
    i64 r0 = 0;
    for (int d = 0; d < N; d++) {
        const int g = (int)std::max<i64>(1, std::min<i64>((lens[d] + 255) / 256, (i64)num_sms() * 8));
        outer_table_kernel<T, N, ORDER, BCF><<<g, 256, 0, s>>>(p, d, lens[d], tap + r0 * L, w + r0 * L, flg + r0);
        t.tap[d] = tap + r0 * L;
        t.w[d] = w + r0 * L;
        t.flg[d] = flg + r0;
        r0 += lens[d];
    }
    const int grid = (int)std::max<i64>(1, std::min<i64>((nq + 255) / 256, (i64)num_sms() * 16));
    if (flags & GB_FAST) outer_eval_kernel<T, N, ORDER, BCF, true><<<grid, 256, 0, s>>>(p, t, (T*)out, nq);
    else outer_eval_kernel<T, N, ORDER, BCF, false><<<grid, 256, 0, s>>>(p, t, (T*)out, nq);
    count_launch(N + 1);
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
template <typename T, int N, int ORDER> int launch_outer_order(const EvalParams<N>& p, int bcf, const i64* lens, int flags, void* out, cudaStream_t s) {
    if constexpr (ORDER == GB_CUBIC) {
        if (bcf != BCF_NIL) return fail(GB_ERR_UNSUPPORTED, "InterpCubic supports only BCnil/BCnan/BCna");
        return launch_outer_bc<T, N, ORDER, BCF_NIL>(p, lens, flags, out, s);
    } else {
        switch (bcf) {
        case BCF_NIL: return launch_outer_bc<T, N, ORDER, BCF_NIL>(p, lens, flags, out, s);
        case BCF_REFLECT: return launch_outer_bc<T, N, ORDER, BCF_REFLECT>(p, lens, flags, out, s);
        case BCF_PERIODIC: return launch_outer_bc<T, N, ORDER, BCF_PERIODIC>(p, lens, flags, out, s);
        case BCF_CLAMP: return launch_outer_bc<T, N, ORDER, BCF_CLAMP>(p, lens, flags, out, s);
        }
        return fail(GB_ERR_ARG, "bad boundary condition");
    }
}

template <typename T, int N, int ORDER, int BCF>
int launch_mode(const EvalParams<N>& p, int outmode, int flags, cudaStream_t s) {
    const bool fast = (flags & GB_FAST) != 0;
    switch (outmode) {
    case 0: return fast ? launch_one<T, N, ORDER, BCF, 0, true>(p, flags, s) : launch_one<T, N, ORDER, BCF, 0, false>(p, flags, s);
    case 1: return fast ? launch_one<T, N, ORDER, BCF, 1, true>(p, flags, s) : launch_one<T, N, ORDER, BCF, 1, false>(p, flags, s);
    case 2:
        if constexpr (ORDER >= GB_QUADRATIC)
            return fast ? launch_one<T, N, ORDER, BCF, 2, true>(p, flags, s) : launch_one<T, N, ORDER, BCF, 2, false>(p, flags, s);
        else
            return fail(GB_ERR_UNSUPPORTED, "valgradhess is not defined for InterpNearest/InterpLinear (no interp_hcoefs_1d method)");
    }
    return fail(GB_ERR_ARG, "bad output mode");
}

template <typename T, int N, int ORDER> int launch_order(const EvalParams<N>& p, int bcf, int outmode, int flags, cudaStream_t s) {
    if constexpr (ORDER == GB_CUBIC) {
        if (bcf != BCF_NIL) return fail(GB_ERR_UNSUPPORTED, "InterpCubic supports only BCnil/BCnan/BCna");
        return launch_mode<T, N, ORDER, BCF_NIL>(p, outmode, flags, s);
    } else {
        switch (bcf) {
        case BCF_NIL: return launch_mode<T, N, ORDER, BCF_NIL>(p, outmode, flags, s);
        case BCF_REFLECT: return launch_mode<T, N, ORDER, BCF_REFLECT>(p, outmode, flags, s);
        case BCF_PERIODIC: return launch_mode<T, N, ORDER, BCF_PERIODIC>(p, outmode, flags, s);
        case BCF_CLAMP: return launch_mode<T, N, ORDER, BCF_CLAMP>(p, outmode, flags, s);
        }
        return fail(GB_ERR_ARG, "bad boundary condition");
    }
}

}  // namespace gb
