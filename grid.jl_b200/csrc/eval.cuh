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

template <typename T> __device__ __forceinline__ T ldro(const T* p) { return __ldg(p); }

// ---- EXACT contraction -------------------------------------------------------------------------
template <typename T, int N, int L, int NP, bool GUARD, int D> struct ExactRec {
    static __device__ __forceinline__ void run(const T* __restrict__ A, i64 base, i64 total, const i64 (&off)[N][L],
                                               const T (&w)[3][N][L], const T (&pp)[NP], T (&acc)[NP]) {
#pragma unroll
        for (int i = 0; i < L; i++) {
            T p[NP];
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const T s = w[sel(N, k, D)][D][i];
                p[k] = (D == N - 1) ? s : pp[k] * s;
            }
            if constexpr (D == 0) {
                const i64 j = base + off[0][i];
                bool inb = true;
                if constexpr (GUARD) inb = (u64)j < (u64)total;  // D1
                const T a = inb ? ldro(A + j) : (T)0;
#pragma unroll
                for (int k = 0; k < NP; k++) {
                    const T c = p[k];
                    const T term = (c != (T)0 && inb) ? c * a : (T)0;  // zero-weight taps never contribute (:504)
                    acc[k] = acc[k] + term;
                }
            } else {
                ExactRec<T, N, L, NP, GUARD, D - 1>::run(A, base + off[D][i], total, off, w, p, acc);
            }
        }
    }
};

// ---- FAST contraction: separable, FMA --------------------------------------------------------
// Q<D> = value, gradient components 0..D and Hessian entries (i<=j<=D) of the partial contraction
// over dimensions 0..D.
template <typename T, int N, int OUTMODE> struct FastQ {
    T v;
    T g[OUTMODE >= 1 ? N : 1];
    T h[OUTMODE >= 2 ? N * N : 1];  // [i*N+j], i<=j
};
template <typename T, int N, int L, int OUTMODE, bool GUARD, int D> struct FastRec {
    static __device__ __forceinline__ FastQ<T, N, OUTMODE> run(const T* __restrict__ A, i64 base, i64 total,
                                                              const i64 (&off)[N][L], const T (&w)[3][N][L]) {
        FastQ<T, N, OUTMODE> out;
        out.v = (T)0;
        if constexpr (OUTMODE >= 1) {
#pragma unroll
            for (int k = 0; k <= D; k++) out.g[k] = (T)0;
        }
        if constexpr (OUTMODE >= 2) {
#pragma unroll
            for (int a = 0; a <= D; a++)
#pragma unroll
                for (int b = a; b <= D; b++) out.h[a * N + b] = (T)0;
        }
#pragma unroll
        for (int i = 0; i < L; i++) {
            const T wv = w[0][D][i];
            const bool pv = wv != (T)0;
            if constexpr (D == 0) {
                const i64 j = base + off[0][i];
                bool inb = true;
                if constexpr (GUARD) inb = (u64)j < (u64)total;
                const T a = inb ? ldro(A + j) : (T)0;
                if (pv && inb) out.v = fma(wv, a, out.v);
                if constexpr (OUTMODE >= 1) {
                    const T wg = w[1][0][i];
                    if (wg != (T)0 && inb) out.g[0] = fma(wg, a, out.g[0]);
                }
                if constexpr (OUTMODE >= 2) {
                    const T wh = w[2][0][i];
                    if (wh != (T)0 && inb) out.h[0] = fma(wh, a, out.h[0]);
                }
            } else {
                const FastQ<T, N, OUTMODE> in = FastRec<T, N, L, OUTMODE, GUARD, D - 1>::run(A, base + off[D][i], total, off, w);
                if (pv) out.v = fma(wv, in.v, out.v);
                if constexpr (OUTMODE >= 1) {
                    const T wg = w[1][D][i];
                    const bool pg = wg != (T)0;
#pragma unroll
                    for (int k = 0; k < D; k++)
                        if (pv) out.g[k] = fma(wv, in.g[k], out.g[k]);
                    if (pg) out.g[D] = fma(wg, in.v, out.g[D]);
                    if constexpr (OUTMODE >= 2) {
                        const T wh = w[2][D][i];
#pragma unroll
                        for (int a = 0; a < D; a++) {
#pragma unroll
                            for (int b = a; b < D; b++)
                                if (pv) out.h[a * N + b] = fma(wv, in.h[a * N + b], out.h[a * N + b]);
                            if (pg) out.h[a * N + D] = fma(wg, in.g[a], out.h[a * N + D]);
                        }
                        if (wh != (T)0) out.h[D * N + D] = fma(wh, in.v, out.h[D * N + D]);
                    }
                }
            }
        }
        return out;
    }
};

// ---- coordinate fetch: user coordinate -> index coordinate in T --------------------------------
// (coordlookup src/coord.jl:30-32, setx src/interp.jl:75-139), each operation in Julia's type.
template <typename T, int N> __device__ __forceinline__ T fetch_coord(const EvalParams<N>& p, int d, i64 q) {
    if (p.xdtype == GB_F64) {
        const double xq = __ldg((const double*)p.xp[d] + q * p.xqs);
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
        const float xq = __ldg((const float*)p.xp[d] + q * p.xqs);
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

// ---- the kernel ------------------------------------------------------------------------------
template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
__global__ void __launch_bounds__(128) eval_kernel(const EvalParams<N> p) {
    constexpr int L = ORDER + 1;
    constexpr int NP = num_passes(N, OUTMODE);
    // D1 guard: taps that can fall outside the array (upper-edge tap of Linear on the offset path,
    // 4th tap of Cubic at x == len-1).
    constexpr bool GUARD = (ORDER == GB_LINEAR) || (BCF == BCF_NIL && ORDER == GB_CUBIC);
    const T* __restrict__ A = (const T*)p.coefs;
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < p.nq; q += step) {
        i64 off[N][L];
        int c[N][L];
        T w[3][N][L];
        bool valid = true, anywrap = false;
#pragma unroll
        for (int d = 0; d < N; d++) {
            const T x = fetch_coord<T, N>(p, d, q);
            T dx;
            bool iswrap;
            valid = dim_taps<T, ORDER, BCF>(x, p.len[d], c[d], dx, iswrap) && valid;
            anywrap = anywrap || iswrap;
            weights<ORDER, (OUTMODE >= 1), (OUTMODE >= 2)>(dx, w[0][d], w[1][d], w[2][d]);
        }
        // set_position (src/interp.jl:406-428): unless SOME dimension reported a wrap, the reference
        // addresses taps through the precomputed offset table (ix + 0, ix + 1, ...) and ignores
        // coord1d.  The two differ only for InterpLinear exactly on a grid line at the upper edge
        // (ix == len, dx == 0): the second tap is then ix+1 (its value weight is 0, its gradient
        // weight is not).  Reproduced for bit-parity; the D1 guard keeps the load in bounds.
        if constexpr (ORDER == GB_LINEAR && BCF != BCF_NIL) {
            if (!anywrap) {
#pragma unroll
                for (int d = 0; d < N; d++) c[d][1] = c[d][0] + 1;
            }
        }
#pragma unroll
        for (int d = 0; d < N; d++) {
#pragma unroll
            for (int i = 0; i < L; i++) off[d][i] = (i64)c[d][i] * p.stride[d];
        }
        T acc[NP];
        if constexpr (!FAST) {
#pragma unroll
            for (int k = 0; k < NP; k++) acc[k] = (T)0;
            T pp[NP];
#pragma unroll
            for (int k = 0; k < NP; k++) pp[k] = (T)1;
            ExactRec<T, N, L, NP, GUARD, N - 1>::run(A, 0, p.total, off, w, pp, acc);
        } else {
            const FastQ<T, N, OUTMODE> r = FastRec<T, N, L, OUTMODE, GUARD, N - 1>::run(A, 0, p.total, off, w);
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
}

template <typename T, int N, int ORDER, int BCF, int OUTMODE, bool FAST>
int launch_one(const EvalParams<N>& p, cudaStream_t s) {
    auto kern = eval_kernel<T, N, ORDER, BCF, OUTMODE, FAST>;
    static int blocks_per_sm = 0;  // benign race: every thread computes the same value
    if (blocks_per_sm == 0) {
        int b = 0;
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, 128, 0));
        blocks_per_sm = b > 0 ? b : 1;
    }
    const i64 need = (p.nq + 127) / 128;
    const i64 cap = (i64)num_sms() * blocks_per_sm * 4;  // a few waves, grid-stride beyond that
    const int grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    kern<<<grid, 128, 0, s>>>(p);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

template <typename T, int N, int ORDER, int BCF>
int launch_mode(const EvalParams<N>& p, int outmode, int flags, cudaStream_t s) {
    const bool fast = (flags & GB_FAST) != 0;
    switch (outmode) {
    case 0: return fast ? launch_one<T, N, ORDER, BCF, 0, true>(p, s) : launch_one<T, N, ORDER, BCF, 0, false>(p, s);
    case 1: return fast ? launch_one<T, N, ORDER, BCF, 1, true>(p, s) : launch_one<T, N, ORDER, BCF, 1, false>(p, s);
    case 2:
        if constexpr (ORDER >= GB_QUADRATIC)
            return fast ? launch_one<T, N, ORDER, BCF, 2, true>(p, s) : launch_one<T, N, ORDER, BCF, 2, false>(p, s);
        else
            return fail(GB_ERR_UNSUPPORTED, "valgradhess is not defined for InterpNearest/InterpLinear (no interp_hcoefs_1d method)");
    }
    return fail(GB_ERR_ARG, "bad output mode");
}

template <typename T, int N, int ORDER> int launch_bc(const EvalParams<N>& p, int bcf, int outmode, int flags, cudaStream_t s) {
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

template <typename T, int N> int launch_eval_impl(const EvalParams<N>& p, int order, int bcf, int outmode, int flags, cudaStream_t s) {
    switch (order) {
    case GB_NEAREST: return launch_bc<T, N, GB_NEAREST>(p, bcf, outmode, flags, s);
    case GB_LINEAR: return launch_bc<T, N, GB_LINEAR>(p, bcf, outmode, flags, s);
    case GB_QUADRATIC: return launch_bc<T, N, GB_QUADRATIC>(p, bcf, outmode, flags, s);
    case GB_CUBIC: return launch_bc<T, N, GB_CUBIC>(p, bcf, outmode, flags, s);
    }
    return fail(GB_ERR_ARG, "bad interpolation order");
}

#define GB_DEFINE_EVAL(T, N) \
    int launch_eval_##T##_##N(const EvalParams<N>& p, int order, int bcf, int outmode, int flags, cudaStream_t s) { return launch_eval_impl<T, N>(p, order, bcf, outmode, flags, s); }

}  // namespace gb
