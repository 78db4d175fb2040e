// eval.cuh -- point evaluation of InterpGrid / CoordInterpGrid on sm_100a.
//
// One thread evaluates one query: it computes, per dimension, the tap coordinates under the
// boundary condition (branch structure of src/interp.jl:594-783, resolved at compile time per
// (order, BC family)), the B-spline value / gradient / Hessian weights in registers
// (src/interp.jl:785-847), and gathers the l^N coefficient neighbourhood through the read-only
// path.  Two contraction schemes share the taps and weights:
//   EXACT  the reference's arithmetic: coefficient of a tap = ((w_N*w_{N-1})*...)*w_1
//          (src/interp.jl:1130-1183), value = sequential sum over taps, dim 1 fastest, of
//          c==0 ? 0 : c*A[idx] (src/interp.jl:501-505); gradient / Hessian passes rebuild the
//          products with g/h weights (src/interp.jl:443-484).  All passes share ONE sweep over
//          the taps (each tap is loaded once); per-pass accumulation order is unchanged, so the
//          results are bit-identical to the scalar reference algorithm.
//   FAST   separable contraction, dimension by dimension, with FMA.
// Compiled with -fmad=false: plain * and + never fuse.
#pragma once
#include <algorithm>

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
template <int ORDER, bool G, bool H>
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
        w[0] = ((t * t) * t) / 6;
        w[1] = 2.0 / 3.0 - (d2 * (2 - dx)) / 2;
        w[2] = 2.0 / 3.0 - ((u * u) * (1 + dx)) / 2;
        w[3] = (d2 * dx) / 6;
        if (G) {
            g[0] = (-(t * t)) / 2;
            g[1] = (3 * d2) / 2 - 2 * dx;
            g[2] = 0.5 + dx * (1 - 1.5 * dx);
            g[3] = d2 / 2;
        }
        if (H) { h[0] = 1 - dx; h[1] = 3 * dx - 2; h[2] = 1 - 3 * dx; h[3] = dx; }
    }
}
template <int ORDER, bool G, bool H>
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


// ---- tap loads -----------------------------------------------------------------------------------
// `base` is a generic pointer: the CTA's shared-memory brick (tiled mode, all taps inside) or the
// coefficient array in global memory.  Offsets are 32-bit element offsets (grids are limited to
// 2^32-1 elements at creation).
template <typename T> __device__ __forceinline__ bool nonfinite(T a);
template <> __device__ __forceinline__ bool nonfinite<double>(double a) { return (__double2hiint(a) & 0x7ff00000) == 0x7ff00000; }
template <> __device__ __forceinline__ bool nonfinite<float>(float a) { return (__float_as_int(a) & 0x7f800000) == 0x7f800000; }
// running maximum of |a|'s high word over the taps of a query: >= expo_all_ones() iff some tap was NaN/Inf
__device__ __forceinline__ u32 mag_hi(double a) { return (u32)__double2hiint(a) & 0x7fffffffu; }
__device__ __forceinline__ u32 mag_hi(float a) { return (u32)__float_as_int(a) & 0x7fffffffu; }
template <typename T> __device__ __forceinline__ constexpr u32 expo_all_ones() { return sizeof(T) == 8 ? 0x7ff00000u : 0x7f800000u; }

template <typename T, bool GUARD> __device__ __forceinline__ T load_tap(const T* __restrict__ base, u32 j, u32 total) {
    if constexpr (GUARD) return j < total ? base[j] : (T)0;  // D1: a tap past the array contributes c*0 = 0
    return base[j];
}

// ---- EXACT contraction -------------------------------------------------------------------------
// acc_k += c_k * a per tap, tap order dim-1-fastest, c_k = ((w_N*w_{N-1})*...)*w_1 per pass k.
// The reference's `c == 0 ? 0 : c*a` (src/interp.jl:504) differs from the plain product only when
// a is NaN/Inf (0*a is then NaN instead of 0); adding the resulting +-0 never changes an
// accumulator that started at +0.  So the hot loop forms the product unconditionally, with no
// branch between the tap loads, and records whether any tap was non-finite (`mag`); such a query
// is re-evaluated by eval_one_slow with the literal select.  Same bits, a third of the instructions.
// Taps are loaded a plane (L x L) at a time so that the loads of a plane are in flight together.
template <typename T, int N, int L, int NP, bool GUARD, int D> struct ExactRec {
    static __device__ __forceinline__ void run(const T* __restrict__ A, u32 base, u32 total, const u32 (&off)[N][L],
                                               const T (&w)[3][N][L], const T (&pp)[NP], T (&acc)[NP], u32& mag) {
        if constexpr (D <= 1) {
            constexpr int L1 = (D == 1) ? L : 1;
            T a[L1][L];
#pragma unroll
            for (int i1 = 0; i1 < L1; i1++)
#pragma unroll
                for (int i0 = 0; i0 < L; i0++) {
                    a[i1][i0] = load_tap<T, GUARD>(A, base + (D == 1 ? off[D][i1] : 0u) + off[0][i0], total);
                    mag = max(mag, mag_hi(a[i1][i0]));
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
                ExactRec<T, N, L, NP, GUARD, D - 1>::run(A, base + off[D][i], total, off, w, p, acc, mag);
            }
        }
    }
};

// ---- FAST contraction: separable, FMA --------------------------------------------------------
// Q = value, gradient components 0..D and Hessian entries (i<=j<=D) of the partial contraction
// over dimensions 0..D.  Taps are loaded a plane (L x L) at a time so that the loads of a plane
// are in flight together.  Non-finite taps are handled like in EXACT (mag -> eval_one_slow).
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
template <typename T, int N, int L, int OUTMODE, bool GUARD, int D> struct FastRec {
    static __device__ __forceinline__ FastQ<T, N, OUTMODE> run(const T* __restrict__ A, u32 base, u32 total,
                                                              const u32 (&off)[N][L], const T (&w)[3][N][L], u32& mag) {
        if constexpr (D == 0) {  // N == 1
            T a[L];
#pragma unroll
            for (int i = 0; i < L; i++) { a[i] = load_tap<T, GUARD>(A, base + off[0][i], total); mag = max(mag, mag_hi(a[i])); }
            return fast_row<T, N, L, OUTMODE>(a, w);
        } else if constexpr (D == 1) {  // one plane: L*L loads issued together
            T a[L][L];
#pragma unroll
            for (int i1 = 0; i1 < L; i1++)
#pragma unroll
                for (int i0 = 0; i0 < L; i0++) {
                    a[i1][i0] = load_tap<T, GUARD>(A, base + off[1][i1] + off[0][i0], total);
                    mag = max(mag, mag_hi(a[i1][i0]));
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
                const FastQ<T, N, OUTMODE> in = FastRec<T, N, L, OUTMODE, GUARD, D - 1>::run(A, base + off[D][i], total, off, w, mag);
                fast_fold<T, N, L, OUTMODE, D>(out, in, w, i);
            }
            return out;
        }
    }
};

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

// taps of one query: validity + 0-based tap coordinates, including the reference's offset-table
// addressing for InterpLinear.
// set_position (src/interp.jl:406-428): unless SOME dimension reported a wrap, the reference
// addresses taps through the precomputed offset table (ix + 0, ix + 1, ...) and ignores coord1d.
// The two differ only for InterpLinear exactly on a grid line at the upper edge (ix == len,
// dx == 0): the second tap is then ix+1 (its value weight is 0, its gradient weight is not).
// Reproduced for bit-parity; the D1 guard keeps the load in bounds.
template <typename T, int N, int ORDER, int BCF>
__device__ __forceinline__ bool query_taps(const EvalParams<N>& p, const double (&raw)[N], int (&c)[N][ORDER + 1], T (&dx)[N]) {
    bool valid = true, anywrap = false;
#pragma unroll
    for (int d = 0; d < N; d++) {
        const T x = index_coord<T, N>(p, d, raw[d]);
        bool iswrap;
        valid = dim_taps<T, ORDER, BCF>(x, p.len[d], c[d], dx[d], iswrap) && valid;
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

// ---- rare path: a query that touched a NaN/Inf coefficient ----------------------------------------
// Literal restatement of interp (src/interp.jl:501-505) with the `c == 0 ? 0 : c*a` select, rolled
// loops, global memory only, one pass per output.  Executed for (almost) no query; kept out of line.
template <typename T, int N, int ORDER, int BCF>
__device__ __noinline__ void eval_one_slow(const EvalParams<N>& p, const double (&raw)[N], const i64 q, const int outmode) {
    constexpr int L = ORDER + 1;
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    const T* __restrict__ A = (const T*)p.coefs;
    int c[N][L];
    T dx[N];
    T w[3][N][L];
    const bool valid = query_taps<T, N, ORDER, BCF>(p, raw, c, dx);
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
            if (cf != (T)0 && (!GUARD || (u64)j < (u64)p.total)) term = cf * A[j];
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
        if (k == 0) {
            if (p.val) ((T*)p.val)[q] = acc;
        } else if (k <= N) {
            if (p.grad) ((T*)p.grad)[q * N + (k - 1)] = valid ? scale_grad<T, N>(p, k - 1, acc) : acc;
        } else {
            T* h = (T*)p.hess + q * (N * N);
            h[hj * N + hi] = acc;
            h[hi * N + hj] = acc;
        }
    }
    if (!valid && p.first_invalid) atomicMin(p.first_invalid, (u64)(q + p.q0));
}

// ---- one query -------------------------------------------------------------------------------
// Box<N>: the coefficient brick a CTA staged in shared memory (tiled mode).  A query whose taps all
// fall inside the brick gathers from shared memory, any other query (wrapped taps at the grid
// boundary, taps past the last sample, plain mode) gathers from global memory.
template <int N> struct Box {
    int lo[N];    // first coordinate of the brick
    int clip[N];  // usable extent: min(tile + L - 1, len - lo)
    int sb[N];    // element strides inside the brick
    bool have;
};

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__device__ __forceinline__ void eval_one(const EvalParams<N>& p, const double (&raw)[N], const i64 q, const T* __restrict__ A, const T* box,
                                         const Box<N>& bx) {
    constexpr int L = ORDER + 1;
    constexpr int NP = num_passes(N, OUTMODE);
    // D1 guard: taps that can fall outside the array (upper-edge tap of Linear on the offset path,
    // 4th tap of Cubic at x == len-1).
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    u32 off[N][L];
    int c[N][L];
    T dx[N];
    T w[3][N][L];
    const bool valid = query_taps<T, N, ORDER, BCF>(p, raw, c, dx);
#pragma unroll
    for (int d = 0; d < N; d++) weights<ORDER, (OUTMODE >= 1), (OUTMODE >= 2)>(dx[d], w[0][d], w[1][d], w[2][d]);
    bool allin = bx.have;
    if (bx.have) {
#pragma unroll
        for (int d = 0; d < N; d++) {
#pragma unroll
            for (int i = 0; i < L; i++) allin = allin && (unsigned)(c[d][i] - bx.lo[d]) < (unsigned)bx.clip[d];
        }
    }
    const T* base = allin ? box : A;
    const u32 total = allin ? 0xffffffffu : (u32)p.total;
#pragma unroll
    for (int d = 0; d < N; d++) {
#pragma unroll
        for (int i = 0; i < L; i++) off[d][i] = allin ? (u32)((c[d][i] - bx.lo[d]) * bx.sb[d]) : (u32)c[d][i] * (u32)p.stride[d];
    }
    T acc[NP];
    u32 mag = 0;
    if constexpr (!FAST) {
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = (T)0;
        T pp[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) pp[k] = (T)1;
        ExactRec<T, N, L, NP, GUARD, N - 1>::run(base, 0u, total, off, w, pp, acc, mag);
    } else {
        const FastQ<T, N, OUTMODE> r = FastRec<T, N, L, OUTMODE, GUARD, N - 1>::run(base, 0u, total, off, w, mag);
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
    if (mag >= expo_all_ones<T>()) {  // a NaN/Inf coefficient was touched: literal zero-weight semantics
        eval_one_slow<T, N, ORDER, BCF>(p, raw, q, OUTMODE);
        return;
    }
    if (!valid) {
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = qnan<T>();
        if (p.first_invalid) atomicMin(p.first_invalid, (u64)(q + p.q0));
    }
    if (p.val) ((T*)p.val)[q] = acc[0];
    if constexpr (OUTMODE >= 1) {
        if (p.grad) {
            T* g = (T*)p.grad + q * N;
#pragma unroll
            for (int d = 0; d < N; d++) g[d] = valid ? scale_grad<T, N>(p, d, acc[1 + d]) : acc[1 + d];
        }
    }
    if constexpr (OUTMODE >= 2) {
        T* h = (T*)p.hess + q * (N * N);
#pragma unroll
        for (int k = 0; k < N * (N + 1) / 2; k++) {
            const int i = hess_i(N, k), j = hess_j(N, k);
            h[j * N + i] = acc[1 + N + k];
            h[i * N + j] = acc[1 + N + k];
        }
    }
}

// ---- the evaluation kernel ---------------------------------------------------------------------
// The batch is a list of work items, each a range [begin,end) of query slots handled by one CTA.
// plain mode : item w = slots [w*EVAL_BLOCK, (w+1)*EVAL_BLOCK) in the caller's order; taps are
//              gathered from global memory.
// tiled mode : the queries were binned by grid tile (tile_hist / tile_offsets / tile_scatter
//              below); perm[j] is the original index of the query in sorted slot j.  For each item the CTA stages the tile's coefficient brick (tile +
//              L-1 halo) in shared memory with coalesced loads, then evaluates the item's queries
//              out of it: random 8-byte gathers that would each cost an HBM sector become
//              shared-memory reads and the grid is read from HBM about once per step.
constexpr int EVAL_BLOCK = 256;
constexpr u32 EVAL_CHUNK = 2048;  // queries per work item in tiled mode

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(EVAL_BLOCK, 2) eval_kernel(const EvalParams<N> p) {
    constexpr int L = ORDER + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* box = reinterpret_cast<T*>(smem_raw);
    const T* __restrict__ A = (const T*)p.coefs;
    const bool tiled = p.tiled != 0;
    Box<N> bx;
    bx.have = false;
    int full[N];  // brick extents (shared-memory layout)
    int vol = 1, ntot = 1;
#pragma unroll
    for (int d = 0; d < N; d++) {
        full[d] = p.tile[d] + L - 1;
        bx.lo[d] = 0;
        bx.clip[d] = 0;
        bx.sb[d] = vol;
        vol *= full[d];
        ntot *= p.ntile[d];
    }
    const i64 nwork = tiled ? (i64)p.nwork[0] : (p.nq + EVAL_BLOCK - 1) / EVAL_BLOCK;
    __shared__ u32 s_item;
    for (i64 wi = blockIdx.x;; wi += gridDim.x) {
        if (tiled) {  // items differ in cost: CTAs pull them from a shared counter
            if (threadIdx.x == 0) s_item = atomicAdd(const_cast<u32*>(p.nwork) + 1, 1u);
            __syncthreads();
            wi = s_item;
        }
        if (wi >= nwork) break;
        i64 begin, end;
        if (tiled) {
            const u32 tile = p.work[3 * wi];
            begin = p.work[3 * wi + 1];
            end = p.work[3 * wi + 2];
            bx.have = tile < (u32)ntot;  // the last bin holds invalid queries: no brick needed
            if (bx.have) {
                u32 rem = tile;
#pragma unroll
                for (int d = 0; d < N; d++) {
                    const int t = rem % p.ntile[d];
                    rem /= p.ntile[d];
                    bx.lo[d] = t * p.tile[d];
                    const int left = p.len[d] - bx.lo[d];
                    bx.clip[d] = full[d] < left ? full[d] : left;
                }
#pragma unroll 4
                for (int e = threadIdx.x; e < vol; e += EVAL_BLOCK) {
                    int r = e;
                    u32 g = 0;
                    bool inside = true;
#pragma unroll
                    for (int d = 0; d < N; d++) {
                        const int rd = r % full[d];
                        r /= full[d];
                        inside = inside && rd < bx.clip[d];
                        g += (u32)(bx.lo[d] + rd) * (u32)p.stride[d];
                    }
                    box[e] = inside ? __ldg(A + g) : (T)0;
                }
            }
            __syncthreads();
        } else {
            begin = wi * EVAL_BLOCK;
            end = begin + EVAL_BLOCK < p.nq ? begin + EVAL_BLOCK : p.nq;
        }
        // software pipeline: the next query's coordinates are in flight while the current one is
        // evaluated (tiled mode: its original index perm[j] is fetched one step further ahead, the
        // coordinates are gathered through it)
        i64 j = begin + threadIdx.x;
        double raw_n[N];
        i64 q_n = 0, q_nn = 0;
        if (j < end) {
            q_n = tiled ? (i64)p.perm[j] : j;
#pragma unroll
            for (int d = 0; d < N; d++) raw_n[d] = load_raw<N>(p, d, q_n);
            if (j + EVAL_BLOCK < end) q_nn = tiled ? (i64)p.perm[j + EVAL_BLOCK] : j + EVAL_BLOCK;
        }
        while (j < end) {
            double raw[N];
#pragma unroll
            for (int d = 0; d < N; d++) raw[d] = raw_n[d];
            const i64 q = q_n;
            j += EVAL_BLOCK;
            if (j < end) {
                q_n = q_nn;
#pragma unroll
                for (int d = 0; d < N; d++) raw_n[d] = load_raw<N>(p, d, q_n);
                if (j + EVAL_BLOCK < end) q_nn = tiled ? (i64)p.perm[j + EVAL_BLOCK] : j + EVAL_BLOCK;
            }
            eval_one<T, N, ORDER, BCF, OUTMODE, FAST>(p, raw, q, A, box, bx);
        }
        if (tiled) __syncthreads();
    }
}

// ---- binning ----------------------------------------------------------------------------------------
// Counting sort of the batch by grid tile without global atomics.  SORT_CTAS CTAs each own a
// contiguous slice of the batch:
//   tile_hist     key of every query (tile of its lowest tap per dimension; a locality hint only --
//                 eval_one falls back to global memory for taps outside the brick) + a per-CTA
//                 histogram in shared memory, written bin-major to bh[bin][cta]
//   tile_offsets  (eval_tile.cu) per-bin totals, scan, work-item list, per-(bin,cta) write cursors
//   tile_scatter  (eval_tile.cu) each CTA replays its slice: slot = its cursor for the key
//                 (shared-memory atomic); writes perm[slot] = query index.  perm (4 B/query) stays
//                 L2-resident, so the scattered 4-byte writes merge before they reach HBM; the
//                 coordinates themselves are never moved (the evaluation gathers them through perm)
template <typename T, int N, int ORDER, int BCF>
__global__ void __launch_bounds__(256) tile_hist_kernel(const EvalParams<N> p, u32* __restrict__ keys, u32* __restrict__ bh, u32 nbins, i64 per_cta) {
    extern __shared__ u32 s_hist[];
    for (u32 b = threadIdx.x; b < nbins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    int ntot = 1;
#pragma unroll
    for (int d = 0; d < N; d++) ntot *= p.ntile[d];
    const i64 q0 = (i64)blockIdx.x * per_cta;
    const i64 q1 = q0 + per_cta < p.nq ? q0 + per_cta : p.nq;
    for (i64 q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        double raw[N];
#pragma unroll
        for (int d = 0; d < N; d++) raw[d] = load_raw<N>(p, d, q);
        int c[N][ORDER + 1];
        T dx[N];
        const bool valid = query_taps<T, N, ORDER, BCF>(p, raw, c, dx);
        u32 key = 0, mul = 1;
#pragma unroll
        for (int d = 0; d < N; d++) {
            int lo = c[d][0];
#pragma unroll
            for (int i = 1; i <= ORDER; i++) lo = c[d][i] < lo ? c[d][i] : lo;
            lo = lo < 0 ? 0 : (lo >= p.len[d] ? p.len[d] - 1 : lo);
            key += (u32)(lo / p.tile[d]) * mul;
            mul *= (u32)p.ntile[d];
        }
        if (!valid) key = (u32)ntot;
        keys[q] = key;
        atomicAdd(s_hist + key, 1u);
    }
    __syncthreads();
    for (u32 b = threadIdx.x; b < nbins; b += blockDim.x) bh[(size_t)b * gridDim.x + blockIdx.x] = s_hist[b];
}

__global__ void __launch_bounds__(256) tile_scatter_kernel(const u32* __restrict__ keys, i64 nq, const u32* __restrict__ boff, u32* __restrict__ perm,
                                                           u32 nbins, i64 per_cta);

// eval_tile.cu
int tile_offsets(const u32* bh, u32 nbins, u32 nctas, u32* hist, u32* offs, u32* boff, u32* work, u32* nwork, u32 chunk, cudaStream_t s);
bool choose_tile(int N, int L, size_t esz, const int* len, i64 nq, int* tile, int* ntile);
constexpr u32 SORT_MAX_BINS = 12000;  // per-CTA histogram lives in shared memory (<= 48 KB)

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
int launch_one(const EvalParams<N>& p_in, int flags, cudaStream_t s) {
    constexpr int L = ORDER + 1;
    auto kern = eval_kernel<T, N, ORDER, BCF, OUTMODE, FAST>;
    EvalParams<N> p = p_in;
    static int blocks_plain = 0;  // benign race: every thread computes the same value
    if (blocks_plain == 0) {
        int b = 0;
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, EVAL_BLOCK, 0));
        blocks_plain = b > 0 ? b : 1;
    }
    const size_t grid_bytes = (size_t)p.total * sizeof(T);
    bool tiled = (flags & GB_TILED) != 0;
    if (!(flags & (GB_TILED | GB_PLAIN))) tiled = N >= 2 && ORDER >= GB_LINEAR && p.nq >= (1ll << 20) && grid_bytes >= (48u << 20);
    if (p.nq >= 0xfffffff0ll) tiled = false;
    if (tiled) tiled = choose_tile(N, L, sizeof(T), p.len, (flags & GB_TILED) ? (i64)1 << 40 : p.nq, p.tile, p.ntile);
    p.tiled = 0;
    if (!tiled) {
        for (int d = 0; d < N; d++) { p.tile[d] = 1; p.ntile[d] = 1; }
        const i64 need = (p.nq + EVAL_BLOCK - 1) / EVAL_BLOCK;
        const i64 cap = (i64)num_sms() * blocks_plain * 4;  // a few waves, then CTAs stride over the items
        const int grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
        kern<<<grid, EVAL_BLOCK, 0, s>>>(p);
        count_launch();
        GB_CUDA(cudaGetLastError());
        return GB_OK;
    }
    // ---- tiled path: bin the queries by tile, then evaluate tile by tile out of shared memory ----
    u32 nbins = 1;
    size_t vol = 1;
    for (int d = 0; d < N; d++) { nbins *= (u32)p.ntile[d]; vol *= (size_t)(p.tile[d] + L - 1); }
    nbins += 1;  // + the bin of invalid queries
    const size_t smem = vol * sizeof(T);
    const u32 nctas = (u32)std::max<i64>(1, std::min<i64>((i64)num_sms() * 2, (p.nq + 4095) / 4096));
    const i64 per_cta = (p.nq + nctas - 1) / nctas;
    const size_t max_items = (size_t)nbins + (size_t)(p.nq / EVAL_CHUNK) + 2;
    const i64 pitch = (p.nq + 15) & ~(i64)15;
    // one scratch allocation: keys | perm | bh | boff | hist | offs | nwork | work
    const size_t nbc = (size_t)nbins * nctas;
    const size_t words = 2 * (size_t)pitch + 2 * nbc + 2 * (size_t)(nbins + 4) + 4 + 3 * max_items;
    u32* scratch = nullptr;
    GB_CUDA(cudaMallocAsync((void**)&scratch, words * sizeof(u32), s));
    u32* keys = scratch;
    u32* perm = keys + pitch;
    u32* bh = perm + pitch;
    u32* boff = bh + nbc;
    u32* hist = boff + nbc;
    u32* offs = hist + (nbins + 4);
    u32* nwork = offs + (nbins + 4);
    u32* work = nwork + 4;
    const size_t hsmem = (size_t)nbins * sizeof(u32);
    tile_hist_kernel<T, N, ORDER, BCF><<<nctas, 256, hsmem, s>>>(p, keys, bh, nbins, per_cta);
    count_launch();
    int rc = tile_offsets(bh, nbins, nctas, hist, offs, boff, work, nwork, EVAL_CHUNK, s);
    if (rc) { cudaFreeAsync(scratch, s); return rc; }
    tile_scatter_kernel<<<nctas, 256, hsmem, s>>>(keys, p.nq, boff, perm, nbins, per_cta);
    count_launch();
    GB_CUDA(cudaGetLastError());
    static size_t smem_set = 0;
    if (smem > smem_set) {
        GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int bt = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt, kern, EVAL_BLOCK, smem));
    if (bt < 1) { cudaFreeAsync(scratch, s); return fail(GB_ERR_CUDA, "tiled evaluation: brick does not fit in shared memory"); }
    p.tiled = 1;
    p.perm = perm;
    p.work = work;
    p.nwork = nwork;
    const i64 grid = std::min<i64>((i64)max_items, (i64)num_sms() * bt);
    kern<<<(int)grid, EVAL_BLOCK, smem, s>>>(p);
    count_launch();
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaFreeAsync(scratch, s));
    return GB_OK;
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
