// multigrid.cu -- restrict / prolong (src/restrict_prolong.jl:10-48, 103-140) on sm_100a.
//
// The reference accumulates each output line with 3-4 strided BLAS axpy! calls into a zeroed
// array.  Per output element that is a fixed, short sequence   acc = 0; acc = acc + a_k*x_k ...
// in the order of the axpy! calls; the kernels below evaluate exactly that sequence (unfused
// multiply and add; the library is built with -fmad=false), one thread per output element.
// HBM-bound streaming stencils: no tensor cores, grid sized in multiples of the SM count.
//
// restrict3 / prolong3 fuse the reference's dim-by-dim application (mapdim, src/utilities.jl:7-20)
// for 3-D arrays into one kernel: a CTA stages the input brick it needs in shared memory and
// applies the three 1-D operators in the reference's order (dim 1, then 2, then 3), rounding every
// intermediate to T exactly where the staged reference stores it, so the result is bit-identical
// to the three separate passes while the fine grid is read from HBM once.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace gb {
namespace {

__host__ __device__ inline i64 rsize(i64 len) { return (len & 1) ? (len + 1) / 2 : len / 2 + 1; }

template <typename T> struct RCoef { T s1, s2, s75, s25; };
template <typename T> struct alignas(2 * sizeof(T)) Pair { T x, y; };  // two adjacent elements of a row (marching kernels)

// One restricted element j of a line of length L (stride st) starting at a.  axpy! order:
//  odd  L (:117-122): R[j] += s*A[2j] ; R[j] += s/2*A[2j+1] (j<=n-2) ; R[j] += s/2*A[2j-1] (j>=1)
//  even L (:130-136): R[j] += .75s*A[2j] (j<=n-2) ; += .25s*A[2j+1] (j<=n-2) ; += .75s*A[2j-1] (j>=1) ; += .25s*A[2j-2] (j>=1)
template <typename T, typename Load> __device__ __forceinline__ T restrict_elem(Load ld, i64 j, i64 L, i64 n, const RCoef<T>& c) {
    T acc = (T)0;
    if (L & 1) {
        acc = acc + c.s1 * ld(2 * j);
        if (j <= n - 2) acc = acc + c.s2 * ld(2 * j + 1);
        if (j >= 1) acc = acc + c.s2 * ld(2 * j - 1);
    } else {
        if (j <= n - 2) {
            acc = acc + c.s75 * ld(2 * j);
            acc = acc + c.s25 * ld(2 * j + 1);
        }
        if (j >= 1) {
            acc = acc + c.s75 * ld(2 * j - 1);
            acc = acc + c.s25 * ld(2 * j - 2);
        }
    }
    return acc;
}

// One prolonged element k of a line of n inputs to len outputs.
//  odd  len (:26-30): P[2i] = A[i] (copy!) ; P[2i+1] = (0 + .5*A[i]) + .5*A[i+1]
//  even len (:41-44): P[2i] = (0 + .75*A[i]) + .25*A[i+1] ; P[2i+1] = (0 + .25*A[i]) + .75*A[i+1]
template <typename T, typename Load> __device__ __forceinline__ T prolong_elem(Load ld, i64 k, i64 len) {
    const i64 i = k >> 1;
    if (len & 1) {
        if ((k & 1) == 0) return ld(i);
        T acc = (T)0;
        acc = acc + (T)0.5 * ld(i);
        acc = acc + (T)0.5 * ld(i + 1);
        return acc;
    }
    T acc = (T)0;
    if ((k & 1) == 0) {
        acc = acc + (T)0.75 * ld(i);
        acc = acc + (T)0.25 * ld(i + 1);
    } else {
        acc = acc + (T)0.25 * ld(i);
        acc = acc + (T)0.75 * ld(i + 1);
    }
    return acc;
}

// ---- per-dimension kernels: array viewed as (inner, extent, outer) -------------------------------
// Slab form: the input holds planes [in_first, in_first + Lloc) of a global extent L along the
// dimension, the output receives global planes [out_first, out_first + ocount).  The whole-array
// operators are the special case in_first = out_first = 0.  This is what a slab-sharded multigrid
// level runs after its halo exchange: same per-element arithmetic, so the assembled
// result is bit-identical to the single-device one.
template <typename T> __global__ void __launch_bounds__(256) restrict_dim_kernel(const T* __restrict__ A, T* __restrict__ R, i64 inner, i64 L, i64 n, i64 outer,
                                                                                 i64 Lloc, i64 in_first, i64 out_first, i64 ocount, RCoef<T> c) {
    const i64 total = inner * ocount * outer;
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        const i64 lo = e % inner, r = e / inner;
        const i64 j = out_first + r % ocount, o = r / ocount;
        const T* a = A + lo + o * inner * Lloc - in_first * inner;
        R[e] = restrict_elem<T>([&](i64 k) { return __ldg(a + k * inner); }, j, L, n, c);
    }
}
template <typename T> __global__ void __launch_bounds__(256) prolong_dim_kernel(const T* __restrict__ A, T* __restrict__ P, i64 inner, i64 n, i64 len, i64 outer,
                                                                                i64 nloc, i64 in_first, i64 out_first, i64 ocount) {
    const i64 total = inner * ocount * outer;
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        const i64 lo = e % inner, r = e / inner;
        const i64 k = out_first + r % ocount, o = r / ocount;
        const T* a = A + lo + o * inner * nloc - in_first * inner;
        P[e] = prolong_elem<T>([&](i64 i) { return __ldg(a + i * inner); }, k, len);
    }
}

template <typename T> RCoef<T> make_coef(double scale) {
    RCoef<T> c;
    c.s1 = (T)scale; c.s2 = (T)(scale / 2); c.s75 = (T)(0.75 * scale); c.s25 = (T)(0.25 * scale);
    return c;
}
inline int grid_for(i64 total, int block) {
    const i64 need = (total + block - 1) / block;
    const i64 cap = (i64)num_sms() * 16;
    return (int)std::max<i64>(1, std::min(need, cap));
}

// input planes an output range needs: restrict output j reads 2j-2..2j+1 (even L) / 2j-1..2j+1 (odd L)
inline void restrict_needs(i64 L, i64 n, i64 k0, i64 k1, i64* lo, i64* hi) {
    *lo = k0 == 0 ? 0 : ((L & 1) ? 2 * k0 - 1 : 2 * k0 - 2);
    *hi = (k1 - 1 <= n - 2) ? 2 * (k1 - 1) + 1 : L - 1;
    if (*hi > L - 1) *hi = L - 1;
}
inline void prolong_needs(i64 n, i64 len, i64 k0, i64 k1, i64* lo, i64* hi) {
    *lo = k0 >> 1;
    const i64 kl = k1 - 1;
    const bool copy_only = (len & 1) && !(kl & 1);  // P[2i] = A[i]
    *hi = std::min(n - 1, (kl >> 1) + (copy_only ? 0 : 1));
}

template <typename T> int restrict_slab_impl(int ndim, const i64* dims, const T* in, int dim, double scale, i64 L, i64 in_first, i64 out_first, i64 ocount,
                                             T* out, cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "restrict: dim out of range");
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < ndim; d++) outer *= dims[d];
    const i64 Lloc = dims[dim], n = rsize(L);
    if (ocount < 0 || out_first < 0 || out_first + ocount > n) return fail(GB_ERR_ARG, "restrict: output planes out of range");
    if (inner * ocount * outer == 0) return GB_OK;
    i64 lo, hi;
    restrict_needs(L, n, out_first, out_first + ocount, &lo, &hi);
    if (lo < in_first || hi >= in_first + Lloc) return fail(GB_ERR_ARG, "restrict: the slab does not hold every input plane the requested outputs need");
    restrict_dim_kernel<T><<<grid_for(inner * ocount * outer, 256), 256, 0, s>>>(in, out, inner, L, n, outer, Lloc, in_first, out_first, ocount, make_coef<T>(scale));
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
template <typename T> int restrict_impl(int ndim, const i64* dims, const T* in, int dim, double scale, T* out, cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "restrict: dim out of range");
    return restrict_slab_impl<T>(ndim, dims, in, dim, scale, dims[dim], 0, 0, rsize(dims[dim]), out, s);
}
template <typename T> int prolong_slab_impl(int ndim, const i64* dims, const T* in, int dim, i64 n, i64 len, i64 in_first, i64 out_first, i64 ocount, T* out,
                                            cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "prolong: dim out of range");
    const i64 newsz = (len & 1) ? 2 * n - 1 : 2 * (n - 1);
    if (newsz != len) return fail(GB_ERR_PROLONG_SIZE, "Cannot prolong a dimension of length " + std::to_string(n) + " to " + std::to_string(len));
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < ndim; d++) outer *= dims[d];
    const i64 nloc = dims[dim];
    if (ocount < 0 || out_first < 0 || out_first + ocount > len) return fail(GB_ERR_ARG, "prolong: output planes out of range");
    if (inner * ocount * outer == 0) return GB_OK;
    i64 lo, hi;
    prolong_needs(n, len, out_first, out_first + ocount, &lo, &hi);
    if (lo < in_first || hi >= in_first + nloc) return fail(GB_ERR_ARG, "prolong: the slab does not hold every input plane the requested outputs need");
    prolong_dim_kernel<T><<<grid_for(inner * ocount * outer, 256), 256, 0, s>>>(in, out, inner, n, len, outer, nloc, in_first, out_first, ocount);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
template <typename T> int prolong_impl(int ndim, const i64* dims, const T* in, int dim, i64 len, T* out, cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "prolong: dim out of range");
    return prolong_slab_impl<T>(ndim, dims, in, dim, dims[dim], len, 0, 0, len, out, s);
}

// ---- edge-corrected variants: restrictb / prolongb / restrict_extrap (src/restrict_prolong.jl:187-320) ----
// The reference computes the plain operator and then overwrites the two edge planes along `dim`
// (prolongb: the two outermost planes on each side), in a fixed assignment order; this kernel does
// the same on the finished output, one thread per 1-D line, same order (so degenerate short lines,
// where the planes coincide, end up with the value the reference's last assignment leaves).
// Constants are formed as the reference forms them: b = scale/sqrt(3) (Float64) converted to T,
// a = 2*b in T; (3/4)*x promotes a Float32 x to Float64 before the sum is stored.
enum { EDGE_RESTRICTB = 1, EDGE_RESTRICT_EXTRAP = 2, EDGE_PROLONGB = 3 };
template <typename T> struct EdgeCoef { T a, b, c2, c3, s; };
template <typename T> __device__ __forceinline__ T three_quarters_plus(T bx, T y) {  // b*A[i] + (3/4)*A[j]
    if constexpr (sizeof(T) == 4) return (T)((double)bx + 0.75 * (double)y);
    else return bx + 0.75 * y;
}
template <typename T> __global__ void __launch_bounds__(256) edge_kernel(int kind, const T* __restrict__ A, T* __restrict__ O, i64 inner, i64 Lin, i64 Lout, i64 outer,
                                                                         EdgeCoef<T> c) {
    const i64 total = inner * outer, st = inner;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 lo = e % inner, o = e / inner;
        const T* x = A + lo + o * inner * Lin;
        T* y = O + lo + o * inner * Lout;
        const T x0 = x[0], x1 = x[st], xe = x[(Lin - 1) * st], xe1 = x[(Lin - 2) * st];
        if (kind == EDGE_RESTRICTB) {  // :196-224
            y[0] = c.a * x0 + c.b * x1;
            y[(Lout - 1) * st] = c.a * xe + c.b * xe1;
        } else if (kind == EDGE_RESTRICT_EXTRAP) {  // :287-313
            if (Lin & 1) {
                y[0] = c.c2 * x0;
                y[(Lout - 1) * st] = c.c2 * xe;
            } else {
                y[0] = c.c3 * x0 - c.s * x1;
                y[(Lout - 1) * st] = c.c3 * xe - c.s * xe1;
            }
        } else {  // EDGE_PROLONGB :231-267
            if (Lout & 1) {
                y[0] = c.a * x0;
                y[st] = c.b * x0 + x1 / (T)2;
                y[(Lout - 1) * st] = c.a * xe;
                y[(Lout - 2) * st] = c.b * xe + xe1 / (T)2;
            } else {
                y[0] = c.a * x0 + x1 / (T)4;
                y[st] = three_quarters_plus<T>(c.b * x0, x1);
                y[(Lout - 1) * st] = c.a * xe + xe1 / (T)4;
                y[(Lout - 2) * st] = three_quarters_plus<T>(c.b * xe, xe1);
            }
        }
    }
}
template <typename T> int edge_impl(int kind, int ndim, const i64* dims, const T* in, int dim, double scale, i64 Lout, T* out, cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "dim out of range");
    const i64 Lin = dims[dim];
    if (Lin < 2) return fail(GB_ERR_ARG, "the edge-corrected operators need at least 2 samples along the dimension");
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < ndim; d++) outer *= dims[d];
    if (inner * outer == 0) return GB_OK;
    EdgeCoef<T> c;
    const T sc = (T)scale;  // the reference's default scale is one(T)
    c.s = sc;
    c.c2 = (T)2 * sc;
    c.c3 = (T)3 * sc;
    if (kind == EDGE_PROLONGB) {
        if (Lout & 1) { c.b = (T)(1.0 / std::sqrt(3.0)); c.a = (T)2 * c.b; }
        else { c.b = (T)(1.0 / (2.0 * std::sqrt(2.0))); c.a = (T)3 * c.b; }
    } else {
        if (Lin & 1) { c.b = (T)((double)sc / std::sqrt(3.0)); c.a = (T)2 * c.b; }
        else { c.b = (T)((double)sc / (2.0 * std::sqrt(2.0))); c.a = (T)3 * c.b; }
    }
    edge_kernel<T><<<grid_for(inner * outer, 256), 256, 0, s>>>(kind, in, out, inner, Lin, Lout, outer, c);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

// ---- fused 3-D restrict -----------------------------------------------------------------------
// Output brick (32,BY,BZ) per CTA of 8 warps; lane = output x.  Needed input range along a dim for
// outputs [j0, j0+B): 2*j0-2 .. 2*(j0+B-1)+1 -> 2B+2 elements (the odd-L stencil uses 3 of each 4).
// Stage 1 reduces dim 1 straight from global memory into S1[IZ][IY][32] (a warp per (y,z) row: the
// 4 taps of a lane are 2 aligned double2 when the row is 16-byte aligned); stage 2 reduces dim 2
// into S2[IZ][BY][32]; stage 3 reduces dim 3 and stores.  All index arithmetic is 32-bit and hoisted
// per row, so the kernel issues ~30 instructions per fine-grid point (the first version: 110, and it
// was issue-bound at 78 %).
// restrict_at: the axpy! sequence of restrict_elem with its four taps already in registers.
template <typename T> __device__ __forceinline__ T restrict_at(bool even, T am2, T am1, T a0, T ap1, bool has_lo, bool has_hi, const RCoef<T>& c) {
    T acc = (T)0;
    if (!even) {
        acc = acc + c.s1 * a0;
        if (has_hi) acc = acc + c.s2 * ap1;
        if (has_lo) acc = acc + c.s2 * am1;
    } else {
        if (has_hi) {
            acc = acc + c.s75 * a0;
            acc = acc + c.s25 * ap1;
        }
        if (has_lo) {
            acc = acc + c.s75 * am1;
            acc = acc + c.s25 * am2;
        }
    }
    return acc;
}
template <typename T, int BY, int BZ> __global__ void __launch_bounds__(256) restrict3_kernel(const T* __restrict__ A, T* __restrict__ R, int L1, int L2, int L3, int n1, int n2, int n3, RCoef<T> c) {
    constexpr int BX = 32, IY = 2 * BY + 2, IZ = 2 * BZ + 2, NW = 8;
    extern __shared__ unsigned char smem_raw[];
    T* S1 = reinterpret_cast<T*>(smem_raw);  // [IZ][IY][BX]
    T* S2 = S1 + IZ * IY * BX;                // [IZ][BY][BX]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j1 = blockIdx.x * BX, j2 = blockIdx.y * BY, j3 = blockIdx.z * BZ;
    const int y0 = 2 * j2 - 2, z0 = 2 * j3 - 2;  // input origin of the staged brick in dims 2,3
    const bool ev1 = !(L1 & 1), ev2 = !(L2 & 1), ev3 = !(L3 & 1);
    // ---- stage 1: dim-1 restriction for every (y,z) row of the brick
    {
        const int gx = j1 + lane;
        const bool vx = gx < n1, lo = vx && gx >= 1, hi = vx && gx <= n1 - 2;
        const bool need0 = vx && (ev1 ? hi : true);  // A[2j] is read unless (even L, last output)
        const bool vec = !(L1 & 1) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);  // rows 16-byte aligned
        constexpr int UN = 4;  // rows in flight per warp: all loads of UN rows are issued before the first use
        for (int rb = warp; rb < IZ * IY; rb += NW * UN) {
            T am2[UN], am1[UN], a0[UN], ap1[UN];
            bool ok[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int r = rb + u * NW;
                const int y = r % IY, z = r / IY;
                const int gy = y0 + y, gz = z0 + z;
                am2[u] = am1[u] = a0[u] = ap1[u] = (T)0;
                ok[u] = r < IZ * IY && gy >= 0 && gy < L2 && gz >= 0 && gz < L3;  // warp-uniform
                if (ok[u]) {
                    const T* a = A + ((size_t)gz * L2 + gy) * (size_t)L1 + 2 * gx;
                    bool done = false;
                    if constexpr (sizeof(T) == 8) {
                        if (vec) {
                            if (lo) { const double2 t = __ldg(reinterpret_cast<const double2*>(a - 2)); am2[u] = t.x; am1[u] = t.y; }
                            if (hi) { const double2 t = __ldg(reinterpret_cast<const double2*>(a)); a0[u] = t.x; ap1[u] = t.y; }
                            done = true;
                        }
                    }
                    if (!done) {
                        if (lo) { am1[u] = __ldg(a - 1); if (ev1) am2[u] = __ldg(a - 2); }
                        if (need0) a0[u] = __ldg(a);
                        if (hi) ap1[u] = __ldg(a + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int r = rb + u * NW;
                if (r < IZ * IY) S1[r * BX + lane] = (ok[u] && vx) ? restrict_at<T>(ev1, am2[u], am1[u], a0[u], ap1[u], lo, hi, c) : (T)0;
            }
        }
    }
    __syncthreads();
    // ---- stage 2: dim-2 restriction; taps of output y are S1 rows 2y .. 2y+3 of the same z
    for (int r = warp; r < IZ * BY; r += NW) {
        const int y = r % BY, z = r / BY;
        const int gy = j2 + y;
        T v = (T)0;
        if (gy < n2) {
            const T* a = S1 + (z * IY + 2 * y) * BX + lane;
            v = restrict_at<T>(ev2, a[0], a[BX], a[2 * BX], a[3 * BX], gy >= 1, gy <= n2 - 2, c);
        }
        S2[r * BX + lane] = v;
    }
    __syncthreads();
    // ---- stage 3: dim-3 restriction + store
    {
        const int gx = j1 + lane;
        for (int r = warp; r < BZ * BY; r += NW) {
            const int y = r % BY, z = r / BY;
            const int gy = j2 + y, gz = j3 + z;
            if (gx < n1 && gy < n2 && gz < n3) {
                const T* a = S2 + (2 * z * BY + y) * BX + lane;
                R[((size_t)gz * n2 + gy) * (size_t)n1 + gx] =
                    restrict_at<T>(ev3, a[0], a[BY * BX], a[2 * BY * BX], a[3 * BY * BX], gz >= 1, gz <= n3 - 2, c);
            }
        }
    }
}

template <typename T, int BY, int BZ> int restrict3_launch(const T* in, T* out, i64 L1, i64 L2, i64 L3, i64 n1, i64 n2, i64 n3, double scale, cudaStream_t s) {
    constexpr int BX = 32;
    constexpr int IY = 2 * BY + 2, IZ = 2 * BZ + 2;
    const size_t smem = sizeof(T) * (IZ * IY * BX + IZ * BY * BX);
    auto kern = restrict3_kernel<T, BY, BZ>;
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n1 + BX - 1) / BX), (unsigned)((n2 + BY - 1) / BY), (unsigned)((n3 + BZ - 1) / BZ));
    kern<<<grid, 256, smem, s>>>(in, out, (int)L1, (int)L2, (int)L3, (int)n1, (int)n2, (int)n3, make_coef<T>(scale));
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
// ---- fused 3-D restrict, marching along dim 3 ---------------------------------------------------
// The brick kernel above recomputes the dim-1 restriction of its halo rows: (2BY+2)(2BZ+2) staged rows for
// (2BY)(2BZ) new ones -- 1.56x at (32,4,4), and it is issue-bound (88 instructions per fine-grid point, 12 % of
// them FP64).  Here a small CTA (NW warps) owns an output tile of 64 x BY and walks a chunk of output planes: per
// INPUT plane it restricts dim 1 into S1[2BY+2][64] (stage A), dim 2 into one slot of a four-plane ring
// S2[4][BY][64] (stage B), and after every second plane restricts dim 3 across the ring and stores one output
// plane (stage C).  Nothing is recomputed along dims 1 and 3 (but for the two planes at a chunk boundary),
// (2BY+2)/(2BY) along dim 2.  Everything that depends on x is loop-invariant, and the interior case (every
// output of the tile has all its taps: a warp-uniform test) runs without per-lane branches.  Same restrict_at
// sequences in the same order as the brick kernel, hence the same bits.
// restrict_at with all four taps present (interior outputs): no per-tap conditions
template <typename T> __device__ __forceinline__ T restrict_full(bool even, T am2, T am1, T a0, T ap1, const RCoef<T>& c) {
    T acc = (T)0;
    if (!even) {
        acc = acc + c.s1 * a0;
        acc = acc + c.s2 * ap1;
        acc = acc + c.s2 * am1;
    } else {
        acc = acc + c.s75 * a0;
        acc = acc + c.s25 * ap1;
        acc = acc + c.s75 * am1;
        acc = acc + c.s25 * am2;
    }
    return acc;
}
// Slab form with PEER sides (multi-GPU, comm.cu): the input planes of a slab kernel come from up to three arrays -- the
// rank's own slab (A, planes [zin0, zin0 + nzin)), and the slabs of the ranks below and above, read IN PLACE through peer
// (NVLink) pointers: `lo` holds planes from lo0 on, `hi` planes from hi0 on.  [zmin, zmax) is the range of planes that exist
// in one of the three.  No sides: lo = hi = NULL, zmin = zin0, zmax = zin0 + nzin -- the single-array slab form.
template <typename T> struct Sides {
    const T *lo, *hi;
    int lo0, hi0, zmin, zmax;
};
template <typename T> __device__ __forceinline__ const T* slab_plane(const T* __restrict__ A, int zin0, int nzin, const Sides<T>& sd, int gz, size_t plane) {
    if (gz < zin0) return sd.lo + (size_t)(gz - sd.lo0) * plane;
    if (gz >= zin0 + nzin) return sd.hi + (size_t)(gz - sd.hi0) * plane;
    return A + (size_t)(gz - zin0) * plane;
}
template <typename T> Sides<T> no_sides(i64 zin0, i64 nzin) { return Sides<T>{nullptr, nullptr, 0, 0, (int)zin0, (int)(zin0 + nzin)}; }
// A lane owns outputs x = 2*lane, 2*lane+1 of the 64-wide tile: their eight taps are six consecutive inputs (three
// aligned 16-byte loads), and shared memory moves pairs (one column per lane: 2.94 instead of 2.63 ms per chain).
template <typename T, int BY, int NW, int UN, bool DO3, bool PEER> __global__ void __launch_bounds__(32 * NW) restrict3_march_kernel(const T* __restrict__ A, T* __restrict__ R, int L1, int L2, int L3, int n1,
                                                                                                        int n2, int n3, int zc, RCoef<T> c, int zin0, int nzin,
                                                                                                        int zout0, int ocount, int xtiles, const Sides<T> sd) {
    // slab form: A holds input planes [zin0, zin0 + nzin) of the L3 global ones, R output planes [zout0, zout0 + ocount)
    // of the n3 global ones (whole array: 0, L3, 0, n3); plane indices below are global.
    // DO3 == false: dims 1 and 2 only (2-D arrays, restrict(A, [true,true,false])): plane gz of R comes from plane gz of A.
    constexpr int HX = 32, BX = 64, IY = 2 * BY + 2;
    using V = Pair<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* S1 = reinterpret_cast<V*>(smem_raw);  // [IY][HX]
    V* S2 = S1 + IY * HX;                     // [4][BY][HX]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // xtiles != 0: (x, y) tiles linearised in blockIdx.x, plane chunks in blockIdx.y (tall 2-D arrays have more row tiles than
    // gridDim.y allows); otherwise the plain 3-D grid (measured 3 % faster on the C5 chain)
    const int j1 = (xtiles ? blockIdx.x % xtiles : blockIdx.x) * BX, j2 = (xtiles ? blockIdx.x / xtiles : blockIdx.y) * BY,
              j3 = zout0 + (xtiles ? blockIdx.y : blockIdx.z) * zc;
    const int nzo = min(zc, zout0 + ocount - j3);
    const int y0 = 2 * j2 - 2, z0 = DO3 ? 2 * j3 - 2 : j3;
    const bool ev1 = !(L1 & 1), ev2 = !(L2 & 1), ev3 = !(L3 & 1);
    // rows aligned to a pair (16 B for f64, 8 B for f32); PEER (compile time): the side arrays too -- the single-array kernel is unchanged
    const bool vec = !(L1 & 1) && (((reinterpret_cast<uintptr_t>(A) | (PEER ? reinterpret_cast<uintptr_t>(sd.lo) | reinterpret_cast<uintptr_t>(sd.hi) : 0)) &
                                    (sizeof(V) - 1)) == 0);
    const int gx = j1 + 2 * lane;
    const bool xfull = j1 >= 1 && j1 + BX - 1 <= n1 - 2;  // warp-uniform: every output of the tile has all four taps
    const bool vx0 = gx < n1, vx1 = gx + 1 < n1;
    const bool lo0 = vx0 && gx >= 1, hi0 = vx0 && gx <= n1 - 2, lo1 = vx1, hi1 = vx1 && gx + 1 <= n1 - 2;
    const size_t plane = (size_t)L2 * (size_t)L1;
    for (int p = 0; p < (DO3 ? 2 * nzo + 2 : nzo); p++) {
        const int gz = z0 + p;
        const bool zok = PEER ? (gz >= sd.zmin && gz < sd.zmax) : (gz >= zin0 && gz < zin0 + nzin);  // (planes a needed tap lies on are present: checked by the host)
        // ---- stage A: dim-1 restriction; inputs t[0..5] = A[2gx-2 .. 2gx+3]
        if (zok && j1 < n1) {
            const T* Ac = (PEER ? slab_plane<T>(A, zin0, nzin, sd, gz, plane) : A + (size_t)(gz - zin0) * plane) + 2 * gx;
#pragma unroll 1
            for (int k0 = 0; k0 < (IY + NW - 1) / NW; k0 += UN) {
                T t[UN][6];
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    const int y = warp + NW * (k0 + u), gy = y0 + y;
                    if (y < IY && gy >= 0 && gy < L2) {  // warp-uniform
                        const T* a = Ac + (ptrdiff_t)gy * (ptrdiff_t)L1;
                        if (vec && xfull) {
                            const V q0 = *reinterpret_cast<const V*>(a - 2), q1 = *reinterpret_cast<const V*>(a), q2 = *reinterpret_cast<const V*>(a + 2);
                            t[u][0] = q0.x; t[u][1] = q0.y; t[u][2] = q1.x; t[u][3] = q1.y; t[u][4] = q2.x; t[u][5] = q2.y;
                        } else {
#pragma unroll
                            for (int e = 0; e < 6; e++) t[u][e] = (T)0;
                            // output 0 taps t[0..3] (am2, am1, a0, ap1), output 1 taps t[2..5]; only taps the reference reads are loaded
                            if (lo0) { t[u][1] = __ldg(a - 1); if (ev1) t[u][0] = __ldg(a - 2); }
                            if (vx0 && (ev1 ? hi0 : true)) t[u][2] = __ldg(a);
                            if (hi0 || lo1) t[u][3] = __ldg(a + 1);
                            if (vx1 && (ev1 ? hi1 : true)) t[u][4] = __ldg(a + 2);
                            if (hi1) t[u][5] = __ldg(a + 3);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    const int y = warp + NW * (k0 + u), gy = y0 + y;
                    if (y < IY) {
                        V v{(T)0, (T)0};
                        if (gy >= 0 && gy < L2) {
                            if (xfull) {
                                v.x = restrict_full<T>(ev1, t[u][0], t[u][1], t[u][2], t[u][3], c);
                                v.y = restrict_full<T>(ev1, t[u][2], t[u][3], t[u][4], t[u][5], c);
                            } else {
                                if (vx0) v.x = restrict_at<T>(ev1, t[u][0], t[u][1], t[u][2], t[u][3], lo0, hi0, c);
                                if (vx1) v.y = restrict_at<T>(ev1, t[u][2], t[u][3], t[u][4], t[u][5], lo1, hi1, c);
                            }
                        }
                        S1[y * HX + lane] = v;
                    }
                }
            }
        }
        __syncthreads();
        // ---- stage B: dim-2 restriction into ring slot p & 3; taps of output y are S1 rows 2y .. 2y+3
        if (zok && j1 < n1) {
            V* dst = S2 + ((p & 3) * BY) * HX + lane;
            for (int y = warp; y < BY; y += NW) {
                const int gy = j2 + y;
                V v{(T)0, (T)0};
                if (gy < n2) {
                    const V* a = S1 + (2 * y) * HX + lane;
                    const V b0 = a[0], b1 = a[HX], b2 = a[2 * HX], b3 = a[3 * HX];
                    if (gy >= 1 && gy <= n2 - 2) {
                        v.x = restrict_full<T>(ev2, b0.x, b1.x, b2.x, b3.x, c);
                        v.y = restrict_full<T>(ev2, b0.y, b1.y, b2.y, b3.y, c);
                    } else {
                        v.x = restrict_at<T>(ev2, b0.x, b1.x, b2.x, b3.x, gy >= 1, gy <= n2 - 2, c);
                        v.y = restrict_at<T>(ev2, b0.y, b1.y, b2.y, b3.y, gy >= 1, gy <= n2 - 2, c);
                    }
                }
                if constexpr (DO3) dst[y * HX] = v;
                else if (gy < n2 && vx0) {  // dims 1, 2 only: this is the result
                    T* r = R + ((size_t)(gz - zout0) * (size_t)n2 + gy) * (size_t)n1 + gx;
                    r[0] = v.x;
                    if (vx1) r[1] = v.y;
                }
            }
        }
        __syncthreads();
        if constexpr (!DO3) continue;
        // ---- stage C: input planes p-3 .. p are the taps of output plane (p-3)/2 of this chunk
        if (p >= 3 && (p & 1) && vx0) {
            const int gzo = j3 + ((p - 3) >> 1);
            const bool zlo = gzo >= 1, zhi = gzo <= n3 - 2;
            const V* s0 = S2 + (((p + 1) & 3) * BY) * HX + lane;
            const V* s1 = S2 + (((p + 2) & 3) * BY) * HX + lane;
            const V* s2 = S2 + (((p + 3) & 3) * BY) * HX + lane;
            const V* s3 = S2 + ((p & 3) * BY) * HX + lane;
            T* Rz = R + (size_t)(gzo - zout0) * (size_t)n2 * (size_t)n1 + gx;
            for (int y = warp; y < BY; y += NW) {
                const int gy = j2 + y;
                if (gy < n2) {
                    const int o = y * HX;
                    const V b0 = s0[o], b1 = s1[o], b2 = s2[o], b3 = s3[o];
                    V v;
                    // taps outside the array are never used by restrict_at (zlo / zhi), whatever their slots hold
                    if (zlo && zhi) {
                        v.x = restrict_full<T>(ev3, b0.x, b1.x, b2.x, b3.x, c);
                        v.y = restrict_full<T>(ev3, b0.y, b1.y, b2.y, b3.y, c);
                    } else {
                        v.x = restrict_at<T>(ev3, b0.x, b1.x, b2.x, b3.x, zlo, zhi, c);
                        v.y = restrict_at<T>(ev3, b0.y, b1.y, b2.y, b3.y, zlo, zhi, c);
                    }
                    T* r = Rz + (size_t)gy * (size_t)n1;
                    r[0] = v.x;
                    if (vx1) r[1] = v.y;
                }
            }
        }
    }
}
template <typename T, int BY, int NW, int UN, bool DO3 = true>
int restrict3_march_launch(const T* in, T* out, i64 L1, i64 L2, i64 L3, i64 n1, i64 n2, i64 n3, double scale, i64 zin0, i64 nzin, i64 zout0, i64 ocount, cudaStream_t s,
                           const Sides<T>* sides = nullptr) {
    constexpr int BX = 64, IY = 2 * BY + 2;
    const size_t smem = sizeof(T) * (size_t)(IY * BX + 4 * BY * BX);
    auto kern = sides ? restrict3_march_kernel<T, BY, NW, UN, DO3, true> : restrict3_march_kernel<T, BY, NW, UN, DO3, false>;
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 gx = (n1 + BX - 1) / BX, gy = (n2 + BY - 1) / BY;
    i64 zc = 32;  // (16: 2.63, 32: 2.64, 64: 2.68, 128: 2.80, 256: 3.13 ms per chain)
    while (zc > 4 && gx * gy * ((ocount + zc - 1) / zc) < (i64)num_sms() * 4 * 4) zc /= 2;
    while ((ocount + zc - 1) / zc > 65535) zc *= 2;
    if (gx * gy >= (1ll << 31)) return fail(GB_ERR_UNSUPPORTED, "restrict: too many tiles");
    const bool lin = gy > 65535;
    const dim3 grid = lin ? dim3((unsigned)(gx * gy), (unsigned)((ocount + zc - 1) / zc), 1) : dim3((unsigned)gx, (unsigned)gy, (unsigned)((ocount + zc - 1) / zc));
    kern<<<grid, 32 * NW, smem, s>>>(in, out, (int)L1, (int)L2, (int)L3, (int)n1, (int)n2, (int)n3, (int)zc, make_coef<T>(scale), (int)zin0, (int)nzin, (int)zout0,
                                     (int)ocount, lin ? (int)gx : 0, sides ? *sides : no_sides<T>(zin0, nzin));
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

template <typename T> int restrict3_impl(const i64* dims, const T* in, double scale, T* out, cudaStream_t s) {
    const i64 L1 = dims[0], L2 = dims[1], L3 = dims[2];
    const i64 n1 = rsize(L1), n2 = rsize(L2), n3 = rsize(L3);
    if (L1 < 2 || L2 < 2 || L3 < 2) {  // degenerate: fall back to the staged path
        T *t1 = nullptr, *t2 = nullptr;
        GB_CUDA(cudaMallocAsync((void**)&t1, sizeof(T) * n1 * L2 * L3, s));
        GB_CUDA(cudaMallocAsync((void**)&t2, sizeof(T) * n1 * n2 * L3, s));
        i64 d0[3] = {L1, L2, L3}, d1[3] = {n1, L2, L3}, d2[3] = {n1, n2, L3};
        int rc = restrict_impl<T>(3, d0, in, 0, scale, t1, s);
        if (!rc) rc = restrict_impl<T>(3, d1, t1, 1, scale, t2, s);
        if (!rc) rc = restrict_impl<T>(3, d2, t2, 2, scale, out, s);
        cudaFreeAsync(t1, s);
        cudaFreeAsync(t2, s);
        return rc;
    }
    if (L1 >= (1ll << 30) || L2 >= (1ll << 30) || L3 >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "restrict3: extent too large");
    // brick (32,4,4): measured best of the bricks on 1024^3 f64 (3.6 ms per 1024->17 chain vs 3.9 for (32,4,8) / (32,8,4));
    // GB_RESTRICT3=brick selects it instead of the marching kernel (A/B measurements)
    static const bool brick = [] { const char* e = getenv("GB_RESTRICT3"); return e && e[0] == 'b'; }();
    if (brick) return restrict3_launch<T, 4, 4>(in, out, L1, L2, L3, n1, n2, n3, scale, s);
    // Tile 64 x 4 outputs per CTA of five warps (ten staged rows: two per warp, both in flight): measured best on the
    // 1024 -> 17 chain (2.63 ms; 64 x 4 / 4 warps 2.69, 64 x 8 / 4 warps 2.75-2.81, / 8 warps 2.77, 64 x 6 2.77, 64 x 4 / 8
    // warps 3.68; one column per lane: 32 x 8 / 2 warps 2.96, 128 x 8 / 8 warps 3.5-3.9; the (32,4,4) brick kernel 3.59)
    return restrict3_march_launch<T, 4, 5, 2>(in, out, L1, L2, L3, n1, n2, n3, scale, 0, L3, 0, n3, s);
}

// restrict along dims 1 and 2 only, fused: 2-D arrays (nz = 1) and restrict(A, [true,true,false]) on 3-D arrays
template <typename T> int restrict12_impl(i64 L1, i64 L2, i64 nz, const T* in, double scale, T* out, cudaStream_t s) {
    const i64 n1 = rsize(L1), n2 = rsize(L2);
    if (L1 < 2 || L2 < 2) return fail(GB_ERR_UNSUPPORTED, "fused 2-D restrict: extents below 2 (use gb_restrict per dimension)");
    if (L1 >= (1ll << 30) || L2 >= (1ll << 30) || nz >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "restrict: extent too large");
    if (nz == 0) return GB_OK;
    return restrict3_march_launch<T, 4, 5, 2, false>(in, out, L1, L2, nz, n1, n2, nz, scale, 0, nz, 0, nz, s);
}

// restrict(A, [true,true,true], scale) on ONE SLAB of a 3-D array sharded along dim 3: `in` holds input planes
// [in_first, in_first + dims[2]) of the L3 global ones, `out` receives output planes [out_first, out_first + ocount).
template <typename T> int restrict3_slab_impl(const i64* dims, const T* in, double scale, i64 L3, i64 in_first, i64 out_first, i64 ocount, T* out, cudaStream_t s,
                                              const Sides<T>* sides = nullptr) {
    const i64 L1 = dims[0], L2 = dims[1], nloc = dims[2];
    const i64 n1 = rsize(L1), n2 = rsize(L2), n3 = rsize(L3);
    if (L1 < 2 || L2 < 2 || L3 < 2) return fail(GB_ERR_UNSUPPORTED, "restrict3 slab: extents below 2 (use gb_restrict_slab per dimension)");
    if (L1 >= (1ll << 30) || L2 >= (1ll << 30) || L3 >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "restrict3: extent too large");
    if (ocount < 0 || out_first < 0 || out_first + ocount > n3) return fail(GB_ERR_ARG, "restrict: output planes out of range");
    if (in_first < 0 || in_first + nloc > L3) return fail(GB_ERR_ARG, "restrict: input planes out of range");
    if (ocount == 0) return GB_OK;
    i64 lo, hi;
    restrict_needs(L3, n3, out_first, out_first + ocount, &lo, &hi);
    const i64 have0 = sides ? sides->zmin : in_first, have1 = sides ? sides->zmax : in_first + nloc;
    if (lo < have0 || hi >= have1) return fail(GB_ERR_ARG, "restrict: the slab does not hold every input plane the requested outputs need");
    return restrict3_march_launch<T, 4, 5, 2>(in, out, L1, L2, L3, n1, n2, n3, scale, in_first, nloc, out_first, ocount, s, sides);
}

// ---- fused 3-D prolong ------------------------------------------------------------------------
// Output brick (64,BY,BZ) per CTA of 8 warps; a lane owns output x = lane and lane + 32.  Inputs
// needed along a dim for outputs [k0, k0+B) (k0 even): k0/2 .. k0/2 + B/2.  Same structure as
// restrict3: stage 1 prolongs dim 1 from global memory into S1[nz][ny][64], stage 2 prolongs dim 2
// into S2[nz][BY][64], stage 3 prolongs dim 3 and stores coalesced 512-byte rows.
// prolong_at: prolong_elem with its two taps in registers (a1 is ignored where the reference copies).
template <typename T> __device__ __forceinline__ T prolong_at(bool evenlen, bool kodd, T a0, T a1) {
    if (!evenlen && !kodd) return a0;
    const T w0 = !evenlen ? (T)0.5 : (kodd ? (T)0.25 : (T)0.75), w1 = !evenlen ? (T)0.5 : (kodd ? (T)0.75 : (T)0.25);
    T acc = (T)0;
    acc = acc + w0 * a0;
    acc = acc + w1 * a1;
    return acc;
}
template <typename T, int BY, int BZ> __global__ void __launch_bounds__(256) prolong3_kernel(const T* __restrict__ A, T* __restrict__ P, int n1, int n2, int n3, int m1, int m2, int m3, int do1, int do2, int do3) {
    constexpr int BX = 64, NW = 8;
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k1 = blockIdx.x * BX, k2 = blockIdx.y * BY, k3 = blockIdx.z * BZ;
    const int y0 = do2 ? k2 / 2 : k2, z0 = do3 ? k3 / 2 : k3;
    const int ny = do2 ? BY / 2 + 1 : BY, nz = do3 ? BZ / 2 + 1 : BZ;  // staged extents (input coordinates)
    T* S1 = reinterpret_cast<T*>(smem_raw);  // [nz][ny][BX]  after dim-1 prolongation
    T* S2 = S1 + nz * ny * BX;                // [nz][BY][BX]  after dim-2
    const bool ev1 = !(m1 & 1), ev2 = !(m2 & 1), ev3 = !(m3 & 1);
    // ---- stage 1
    for (int r = warp; r < nz * ny; r += NW) {
        const int y = r % ny, z = r / ny;
        const int gy = y0 + y, gz = z0 + z;
        T v[2] = {(T)0, (T)0};
        if (gy < n2 && gz < n3) {  // warp-uniform
            const T* a = A + ((size_t)gz * n2 + gy) * (size_t)n1;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int gx = k1 + lane + 32 * h;
                if (gx < m1) {
                    if (do1) {
                        const int i = gx >> 1;
                        const T a0 = __ldg(a + i), a1 = (i + 1 < n1) ? __ldg(a + i + 1) : (T)0;
                        v[h] = prolong_at<T>(ev1, gx & 1, a0, a1);
                    } else {
                        v[h] = __ldg(a + gx);
                    }
                }
            }
        }
        S1[r * BX + lane] = v[0];
        S1[r * BX + lane + 32] = v[1];
    }
    __syncthreads();
    // ---- stage 2
    for (int r = warp; r < nz * BY; r += NW) {
        const int y = r % BY, z = r / BY;
        const int gy = k2 + y;
        T v[2] = {(T)0, (T)0};
        if (gy < m2) {
            if (do2) {
                const int ly = y >> 1;
                const bool has1 = (gy >> 1) + 1 < n2;
                const T* a = S1 + (z * ny + ly) * BX + lane;
#pragma unroll
                for (int h = 0; h < 2; h++) v[h] = prolong_at<T>(ev2, gy & 1, a[32 * h], has1 ? a[BX + 32 * h] : (T)0);
            } else {
                const T* a = S1 + (z * ny + y) * BX + lane;
                v[0] = a[0];
                v[1] = a[32];
            }
        }
        S2[r * BX + lane] = v[0];
        S2[r * BX + lane + 32] = v[1];
    }
    __syncthreads();
    // ---- stage 3 + store
    for (int r = warp; r < BZ * BY; r += NW) {
        const int y = r % BY, z = r / BY;
        const int gy = k2 + y, gz = k3 + z;
        if (gy < m2 && gz < m3) {
            T* p = P + ((size_t)gz * m2 + gy) * (size_t)m1 + k1 + lane;
            if (do3) {
                const int lz = z >> 1;
                const bool has1 = (gz >> 1) + 1 < n3;
                const T* a = S2 + (lz * BY + y) * BX + lane;
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (k1 + lane + 32 * h < m1) p[32 * h] = prolong_at<T>(ev3, gz & 1, a[32 * h], has1 ? a[BY * BX + 32 * h] : (T)0);
            } else {
                const T* a = S2 + (z * BY + y) * BX + lane;
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (k1 + lane + 32 * h < m1) p[32 * h] = a[32 * h];
            }
        }
    }
}

// ---- fused 3-D prolong, marching along dim 3 (all three dimensions prolonged) ---------------------
// Same idea as restrict3_march: a small CTA owns a 64 x BY output tile and walks a chunk of output planes.  Per
// INPUT plane it prolongs dim 1 from global memory into S1[BY/2+1][64] (stage A) and dim 2 into one of two ring
// slots S2[2][BY][64] (stage B); two consecutive slots then give two output planes (stage C, coalesced 512-byte
// rows).  A lane owns the output pair x = 2*lane, 2*lane+1 for the whole kernel: both come from the same two
// inputs, the parities are compile-time, and shared memory and (for even row lengths) global stores move pairs.  Same prolong_at calls in the same order as the brick kernel, hence the same bits.
template <typename T, int BY, int NW, bool DO3, bool PEER> __global__ void __launch_bounds__(32 * NW) prolong3_march_kernel(const T* __restrict__ A, T* __restrict__ P, int n1, int n2, int n3, int m1, int m2,
                                                                                               int m3, int zc, int zin0, int nzin, int zout0, int ocount, int xtiles,
                                                                                               const Sides<T> sd) {
    // slab form: A holds input planes [zin0, zin0 + nzin) (of the n3 global ones), P output planes [zout0, zout0 + ocount) of the
    // m3 global ones (whole array: 0, 0, m3); plane indices below are global
    constexpr int BX = 64, NY = BY / 2 + 1, HX = BX / 2;  // a lane owns the output pair x = 2*lane, 2*lane + 1 (input column `xi`)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using V = Pair<T>;
    V* S1 = reinterpret_cast<V*>(smem_raw);  // [NY][HX]
    V* S2 = S1 + NY * HX;                     // [2][BY][HX]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k1 = (xtiles ? blockIdx.x % xtiles : blockIdx.x) * BX, k2 = (xtiles ? blockIdx.x / xtiles : blockIdx.y) * BY,  // (see restrict3_march_kernel)
              k3 = (DO3 ? (zout0 & ~1) : zout0) + (xtiles ? blockIdx.y : blockIdx.z) * zc;  // BX, BY, zc (and k3 when DO3) are even
    const int y0 = k2 >> 1, z0 = DO3 ? k3 >> 1 : k3;  // DO3 == false: dims 1 and 2 only, plane gz of P comes from plane gz of A
    const bool ev1 = !(m1 & 1), ev2 = !(m2 & 1), ev3 = !(m3 & 1);
    const int gx = k1 + 2 * lane, xi = gx >> 1;
    const bool xv0 = gx < m1, xv1 = gx + 1 < m1, xh = xv0 && xi + 1 < n1;
    const bool vst = xv1 && !(m1 & 1) && ((reinterpret_cast<uintptr_t>(P) & (2 * sizeof(T) - 1)) == 0);  // aligned pair stores
    const int nplanes = min(zc, zout0 + ocount - k3);  // output planes of this chunk (those below zout0 are skipped)
    const size_t iplane = (size_t)n2 * (size_t)n1, oplane = (size_t)m2 * (size_t)m1;
    for (int ip = 0; DO3 ? 2 * (ip - 1) < nplanes : ip < nplanes; ip++) {
        const int lz = z0 + ip;
        const bool zin = lz < n3 && lz < (PEER ? sd.zmax : zin0 + nzin);  // (planes a needed tap lies on are present: checked by the host)
        // ---- stage A: dim-1 prolongation of input plane lz, rows y0 .. y0+NY-1
        if (zin) {
            const T* Az = (PEER ? slab_plane<T>(A, zin0, nzin, sd, lz, iplane) : A + (size_t)(lz - zin0) * iplane) + xi;
            for (int y = warp; y < NY; y += NW) {
                const int gy = y0 + y;
                V v{(T)0, (T)0};
                if (gy < n2 && xv0) {
                    const T* a = Az + (size_t)gy * (size_t)n1;
                    const T a0 = __ldg(a), a1 = xh ? __ldg(a + 1) : (T)0;
                    v.x = prolong_at<T>(ev1, false, a0, a1);
                    if (xv1) v.y = prolong_at<T>(ev1, true, a0, a1);
                }
                S1[y * HX + lane] = v;
            }
        }
        __syncthreads();
        // ---- stage B: dim-2 prolongation into ring slot ip & 1
        if (zin) {
            V* dst = S2 + ((ip & 1) * BY) * HX + lane;
            for (int y = warp; y < BY; y += NW) {
                const int gy = k2 + y;
                V v{(T)0, (T)0};
                if (gy < m2) {
                    const bool has1 = (gy >> 1) + 1 < n2;
                    const V b0 = S1[(y >> 1) * HX + lane];
                    V b1{(T)0, (T)0};
                    if (has1) b1 = S1[((y >> 1) + 1) * HX + lane];
                    v.x = prolong_at<T>(ev2, gy & 1, b0.x, b1.x);
                    v.y = prolong_at<T>(ev2, gy & 1, b0.y, b1.y);
                }
                if constexpr (DO3) dst[y * HX] = v;
                else if (gy < m2 && xv0) {  // dims 1, 2 only: this is the result
                    T* p = P + ((size_t)(lz - zout0) * (size_t)m2 + gy) * (size_t)m1 + gx;
                    if (vst) *reinterpret_cast<V*>(p) = v;
                    else {
                        p[0] = v.x;
                        if (xv1) p[1] = v.y;
                    }
                }
            }
        }
        __syncthreads();
        if constexpr (!DO3) continue;
        // ---- stage C: output planes 2(lz-1), 2(lz-1)+1 from slots (ip-1) & 1 (plane lz-1) and ip & 1 (plane lz)
        if (ip >= 1 && xv0) {
            const V* s0 = S2 + (((ip - 1) & 1) * BY) * HX + lane;
            const V* s1 = S2 + ((ip & 1) * BY) * HX + lane;
#pragma unroll
            for (int par = 0; par < 2; par++) {
                const int gz = k3 + 2 * (ip - 1) + par;
                if (gz < k3 + nplanes && gz >= zout0) {  // uniform
                    T* Pz = P + (size_t)(gz - zout0) * oplane + gx;
                    for (int y = warp; y < BY; y += NW) {
                        const int gy = k2 + y;
                        if (gy < m2) {
                            const V c0 = s0[y * HX];
                            V c1{(T)0, (T)0};
                            if (zin) c1 = s1[y * HX];
                            V o;
                            o.x = prolong_at<T>(ev3, par, c0.x, c1.x);
                            o.y = prolong_at<T>(ev3, par, c0.y, c1.y);
                            T* p = Pz + (size_t)gy * (size_t)m1;
                            if (vst) *reinterpret_cast<V*>(p) = o;
                            else {
                                p[0] = o.x;
                                if (xv1) p[1] = o.y;
                            }
                        }
                    }
                }
            }
        }
    }
}
template <typename T, int BY, int NW, bool DO3 = true>
int prolong3_march_launch(const T* in, T* out, i64 n1, i64 n2, i64 n3, const i64* m, i64 zin0, i64 nzin, i64 zout0, i64 ocount, cudaStream_t s, const Sides<T>* sides = nullptr) {
    constexpr int BX = 64, NY = BY / 2 + 1;
    const size_t smem = sizeof(T) * (size_t)(NY * BX + 2 * BY * BX);
    auto kern = sides ? prolong3_march_kernel<T, BY, NW, DO3, true> : prolong3_march_kernel<T, BY, NW, DO3, false>;
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 gx = (m[0] + BX - 1) / BX, gy = (m[1] + BY - 1) / BY;
    i64 zc = 32;  // output planes per chunk (even): one extra input plane per chunk (16: 2.31, 32: 2.29, 64: 2.34, 128: 2.36, 256: 2.42 ms per chain)
    const i64 span = DO3 ? zout0 + ocount - (zout0 & ~(i64)1) : ocount;  // planes from the (even) chunk origin
    while (zc > 8 && gx * gy * ((span + zc - 1) / zc) < (i64)num_sms() * 4 * 4) zc /= 2;
    while ((span + zc - 1) / zc > 65535) zc *= 2;
    if (gx * gy >= (1ll << 31)) return fail(GB_ERR_UNSUPPORTED, "prolong: too many tiles");
    const bool lin = gy > 65535;
    const dim3 grid = lin ? dim3((unsigned)(gx * gy), (unsigned)((span + zc - 1) / zc), 1) : dim3((unsigned)gx, (unsigned)gy, (unsigned)((span + zc - 1) / zc));
    kern<<<grid, 32 * NW, smem, s>>>(in, out, (int)n1, (int)n2, (int)n3, (int)m[0], (int)m[1], (int)m[2], (int)zc, (int)zin0, (int)nzin, (int)zout0, (int)ocount,
                                     lin ? (int)gx : 0, sides ? *sides : no_sides<T>(zin0, nzin));
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

template <typename T> int prolong3_impl(const i64* dims, const T* in, const i64* odims, T* out, cudaStream_t s) {
    const i64 n1 = dims[0], n2 = dims[1], n3 = dims[2];
    i64 m[3];
    int doit[3];
    for (int d = 0; d < 3; d++) {  // prolong only if the target is larger (src/restrict_prolong.jl:94)
        doit[d] = odims[d] > dims[d];
        m[d] = dims[d];
        if (doit[d]) {
            const i64 newsz = (odims[d] & 1) ? 2 * dims[d] - 1 : 2 * (dims[d] - 1);
            if (newsz != odims[d]) return fail(GB_ERR_PROLONG_SIZE, "Cannot prolong a dimension of length " + std::to_string(dims[d]) + " to " + std::to_string(odims[d]));
            m[d] = odims[d];
        }
    }
    if (m[0] * m[1] * m[2] == 0) return GB_OK;
    for (int d = 0; d < 3; d++)
        if (m[d] >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "prolong3: extent too large");
    if (doit[0] && doit[1] && doit[2]) {  // the full prolongation marches (GB_PROLONG3=brick: A/B measurements)
        static const bool brick = [] { const char* e = getenv("GB_PROLONG3"); return e && e[0] == 'b'; }();
        // 64 x 8 tile, 4 warps: measured best on the 17 -> 1024 chain (2.29 ms; 2 warps 2.45, 8 warps 2.46, 64 x 4 2.37-2.62,
        // 64 x 12 2.31, 64 x 16 2.36-3.4; with one column per lane instead of a pair 2.78; the (64,8,8) brick kernel 3.06)
        if (!brick) return prolong3_march_launch<T, 8, 4>(in, out, n1, n2, n3, m, 0, n3, 0, m[2], s);
    }
    constexpr int BX = 64, BY = 8, BZ = 8;
    const int sny = doit[1] ? BY / 2 + 1 : BY, snz = doit[2] ? BZ / 2 + 1 : BZ;
    const size_t smem = sizeof(T) * ((size_t)snz * sny * BX + (size_t)snz * BY * BX);  // S1 + S2 (33 KB for a full f64 prolongation: 6 CTAs per SM)
    auto kern = prolong3_kernel<T, BY, BZ>;
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((m[0] + BX - 1) / BX), (unsigned)((m[1] + BY - 1) / BY), (unsigned)((m[2] + BZ - 1) / BZ));
    kern<<<grid, 256, smem, s>>>(in, out, (int)n1, (int)n2, (int)n3, (int)m[0], (int)m[1], (int)m[2], doit[0], doit[1], doit[2]);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

// prolong along dims 1 and 2 only, fused: 2-D arrays (nz = 1) and prolong(A, [m1,m2,n3]) on 3-D arrays
template <typename T> int prolong12_impl(i64 n1, i64 n2, i64 nz, const T* in, i64 m1, i64 m2, T* out, cudaStream_t s) {
    const i64 nn[2] = {n1, n2}, mm[2] = {m1, m2};
    for (int d = 0; d < 2; d++) {
        const i64 newsz = (mm[d] & 1) ? 2 * nn[d] - 1 : 2 * (nn[d] - 1);
        if (newsz != mm[d]) return fail(GB_ERR_PROLONG_SIZE, "Cannot prolong a dimension of length " + std::to_string(nn[d]) + " to " + std::to_string(mm[d]));
        if (mm[d] <= nn[d]) return fail(GB_ERR_UNSUPPORTED, "fused 2-D prolong: both dimensions must grow (use gb_prolong per dimension)");
        if (mm[d] >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "prolong: extent too large");
    }
    if (nz == 0 || nz >= (1ll << 30)) return nz == 0 ? GB_OK : fail(GB_ERR_UNSUPPORTED, "prolong: extent too large");
    const i64 m[3] = {m1, m2, nz};
    return prolong3_march_launch<T, 8, 4, false>(in, out, n1, n2, nz, m, 0, nz, 0, nz, s);
}

// prolong(A, [m1,m2,m3]) (every dimension prolonged) on ONE SLAB of a 3-D array sharded along dim 3: `in` holds coarse
// planes [in_first, in_first + dims[2]) of the n3 global ones, `out` receives fine planes [out_first, out_first + ocount).
template <typename T>
int prolong3_slab_impl(const i64* dims, const T* in, const i64* m, i64 n3, i64 in_first, i64 out_first, i64 ocount, T* out, cudaStream_t s, const Sides<T>* sides = nullptr) {
    const i64 n1 = dims[0], n2 = dims[1], nloc = dims[2];
    const i64 nn[3] = {n1, n2, n3};
    for (int d = 0; d < 3; d++) {
        const i64 newsz = (m[d] & 1) ? 2 * nn[d] - 1 : 2 * (nn[d] - 1);
        if (newsz != m[d]) return fail(GB_ERR_PROLONG_SIZE, "Cannot prolong a dimension of length " + std::to_string(nn[d]) + " to " + std::to_string(m[d]));
        if (m[d] <= nn[d]) return fail(GB_ERR_UNSUPPORTED, "prolong3 slab: every dimension must grow (use gb_prolong_slab per dimension)");
        if (m[d] >= (1ll << 30)) return fail(GB_ERR_UNSUPPORTED, "prolong3: extent too large");
    }
    if (ocount < 0 || out_first < 0 || out_first + ocount > m[2]) return fail(GB_ERR_ARG, "prolong: output planes out of range");
    if (in_first < 0 || in_first + nloc > n3) return fail(GB_ERR_ARG, "prolong: input planes out of range");
    if (ocount == 0) return GB_OK;
    i64 lo, hi;
    prolong_needs(n3, m[2], out_first, out_first + ocount, &lo, &hi);
    const i64 have0 = sides ? sides->zmin : in_first, have1 = sides ? sides->zmax : in_first + nloc;
    if (lo < have0 || hi >= have1) return fail(GB_ERR_ARG, "prolong: the slab does not hold every input plane the requested outputs need");
    return prolong3_march_launch<T, 8, 4>(in, out, n1, n2, n3, m, in_first, nloc, out_first, ocount, s, sides);
}

// ---- synthetic inputs ---------------------------------------------------------------------------
__device__ __forceinline__ u64 splitmix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename T> __global__ void __launch_bounds__(256) synth_kernel(T* out, i64 n, u64 seed, double lo, double hi, i64 first) {
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += step) {
        const double u = (double)(splitmix64(seed + (u64)(first + k)) >> 11) * (1.0 / 9007199254740992.0);
        out[k] = (T)(lo + (hi - lo) * u);
    }
}

}  // namespace

// channel-last (reference layout, element (i, c) at i + c*st) <-> channel-interleaved (i*nchan + c): the
// resident layout of multi-valued grids, so that the nchan values under one tap are contiguous.
template <typename T> __global__ void __launch_bounds__(256) interleave_kernel(const T* __restrict__ in, T* __restrict__ out, i64 st, i64 nchan, int to_interleaved) {
    const i64 total = st * nchan;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        if (to_interleaved) {
            const i64 i = e / nchan, c = e % nchan;
            out[e] = in[i + c * st];
        } else {
            const i64 c = e / st, i = e % st;
            out[e] = in[i * nchan + c];
        }
    }
}
int interleave_device(int dtype, const void* in, void* out, i64 st, i64 nchan, int to_interleaved, cudaStream_t s) {
    if (st * nchan == 0) return GB_OK;
    const int grid = grid_for(st * nchan, 256);
    if (dtype == GB_F64) interleave_kernel<double><<<grid, 256, 0, s>>>((const double*)in, (double*)out, st, nchan, to_interleaved);
    else if (dtype == GB_F32) interleave_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, st, nchan, to_interleaved);
    else return fail(GB_ERR_ARG, "bad dtype");
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

// column-major <-> blocked resident layout of a 2-D to 4-D coefficient array (tap_offset, eval.cuh): 4^N-element
// blocks in column-major order, Morton order inside a block; cells past the array's edges are zero.
i64 blocked_elems(int ndim, const i64* len) {
    i64 n = 1;
    for (int d = 0; d < ndim; d++) n *= 4 * ((len[d] + 3) / 4);
    return n;
}
struct BlockedGeo { i64 len[4]; i64 nb[4]; int ndim; };
template <typename T> __global__ void __launch_bounds__(256) blocked_kernel(const T* __restrict__ in, T* __restrict__ out, BlockedGeo g, i64 total, int to_blocked) {
    const int N = g.ndim;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 blk = e >> (2 * N), lin = 0, str = 1;
        bool inside = true;
        for (int d = 0; d < N; d++) {
            const i64 c = 4 * (blk % g.nb[d]) + ((e >> d) & 1) + 2 * ((e >> (N + d)) & 1);
            blk /= g.nb[d];
            inside = inside && c < g.len[d];
            lin += c * str;
            str *= g.len[d];
        }
        if (to_blocked) out[e] = inside ? in[lin] : (T)0;
        else if (inside) out[lin] = in[e];
    }
}
int blocked_device(int dtype, int ndim, const i64* len, const void* in, void* out, int to_blocked, cudaStream_t s) {
    if (ndim < 2 || ndim > 4) return fail(GB_ERR_ARG, "blocked layout: 2-D to 4-D arrays");
    BlockedGeo g;
    g.ndim = ndim;
    for (int d = 0; d < ndim; d++) { g.len[d] = len[d]; g.nb[d] = (len[d] + 3) / 4; }
    const i64 total = blocked_elems(ndim, len);
    if (total == 0) return GB_OK;
    const int grid = grid_for(total, 256);
    if (dtype == GB_F64) blocked_kernel<double><<<grid, 256, 0, s>>>((const double*)in, (double*)out, g, total, to_blocked);
    else if (dtype == GB_F32) blocked_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, g, total, to_blocked);
    else return fail(GB_ERR_ARG, "bad dtype");
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

int restrict_device(int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return restrict_impl<double>(ndim, dims, (const double*)in, dim, scale, (double*)out, s);
    if (dtype == GB_F32) return restrict_impl<float>(ndim, dims, (const float*)in, dim, scale, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
// variant: 1 restrictb, 2 restrict_extrap (restrict, then the edge planes); 3 prolongb (prolong, then the edge planes)
int edge_variant_device(int variant, int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, i64 len, void* out, cudaStream_t s) {
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "dim out of range");
    int rc;
    i64 Lout;
    if (variant == EDGE_PROLONGB) {
        Lout = len;
        rc = prolong_device(dtype, ndim, dims, in, dim, len, out, s);
    } else {
        Lout = rsize(dims[dim]);
        rc = restrict_device(dtype, ndim, dims, in, dim, scale, out, s);
    }
    if (rc) return rc;
    if (dtype == GB_F64) return edge_impl<double>(variant, ndim, dims, (const double*)in, dim, scale, Lout, (double*)out, s);
    return edge_impl<float>(variant, ndim, dims, (const float*)in, dim, scale, Lout, (float*)out, s);
}
int prolong_device(int dtype, int ndim, const i64* dims, const void* in, int dim, i64 len, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return prolong_impl<double>(ndim, dims, (const double*)in, dim, len, (double*)out, s);
    if (dtype == GB_F32) return prolong_impl<float>(ndim, dims, (const float*)in, dim, len, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int restrict_slab_device(int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, i64 L, i64 in_first, i64 out_first, i64 ocount, void* out,
                         cudaStream_t s) {
    if (dtype == GB_F64) return restrict_slab_impl<double>(ndim, dims, (const double*)in, dim, scale, L, in_first, out_first, ocount, (double*)out, s);
    if (dtype == GB_F32) return restrict_slab_impl<float>(ndim, dims, (const float*)in, dim, scale, L, in_first, out_first, ocount, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int restrict12_device(int dtype, i64 L1, i64 L2, i64 nz, const void* in, double scale, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return restrict12_impl<double>(L1, L2, nz, (const double*)in, scale, (double*)out, s);
    if (dtype == GB_F32) return restrict12_impl<float>(L1, L2, nz, (const float*)in, scale, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int prolong12_device(int dtype, i64 n1, i64 n2, i64 nz, const void* in, i64 m1, i64 m2, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return prolong12_impl<double>(n1, n2, nz, (const double*)in, m1, m2, (double*)out, s);
    if (dtype == GB_F32) return prolong12_impl<float>(n1, n2, nz, (const float*)in, m1, m2, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int restrict3_slab_device(int dtype, const i64* dims, const void* in, double scale, i64 L3, i64 in_first, i64 out_first, i64 ocount, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return restrict3_slab_impl<double>(dims, (const double*)in, scale, L3, in_first, out_first, ocount, (double*)out, s);
    if (dtype == GB_F32) return restrict3_slab_impl<float>(dims, (const float*)in, scale, L3, in_first, out_first, ocount, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int prolong3_slab_device(int dtype, const i64* dims, const void* in, const i64* m, i64 n3, i64 in_first, i64 out_first, i64 ocount, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return prolong3_slab_impl<double>(dims, (const double*)in, m, n3, in_first, out_first, ocount, (double*)out, s);
    if (dtype == GB_F32) return prolong3_slab_impl<float>(dims, (const float*)in, m, n3, in_first, out_first, ocount, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
// slab kernels whose halo planes are read in place from the neighbouring ranks' slabs (peer pointers; comm.cu): `lo` holds
// input planes from lo_first on (NULL: none below), `hi` planes from hi_first on (NULL: none above), [zmin, zmax) exist in all
template <typename T> Sides<T> make_sides(const void* lo, i64 lo_first, const void* hi, i64 hi_first, i64 zmin, i64 zmax) {
    return Sides<T>{(const T*)lo, (const T*)hi, (int)lo_first, (int)hi_first, (int)zmin, (int)zmax};
}
int restrict3_slab_peer_device(int dtype, const i64* dims, const void* in, double scale, i64 L3, i64 in_first, const void* lo, i64 lo_first, const void* hi, i64 hi_first,
                               i64 zmin, i64 zmax, i64 out_first, i64 ocount, void* out, cudaStream_t s) {
    if (dtype == GB_F64) {
        const Sides<double> sd = make_sides<double>(lo, lo_first, hi, hi_first, zmin, zmax);
        return restrict3_slab_impl<double>(dims, (const double*)in, scale, L3, in_first, out_first, ocount, (double*)out, s, &sd);
    }
    if (dtype == GB_F32) {
        const Sides<float> sd = make_sides<float>(lo, lo_first, hi, hi_first, zmin, zmax);
        return restrict3_slab_impl<float>(dims, (const float*)in, scale, L3, in_first, out_first, ocount, (float*)out, s, &sd);
    }
    return fail(GB_ERR_ARG, "bad dtype");
}
int prolong3_slab_peer_device(int dtype, const i64* dims, const void* in, const i64* m, i64 n3, i64 in_first, const void* lo, i64 lo_first, const void* hi, i64 hi_first,
                              i64 zmin, i64 zmax, i64 out_first, i64 ocount, void* out, cudaStream_t s) {
    if (dtype == GB_F64) {
        const Sides<double> sd = make_sides<double>(lo, lo_first, hi, hi_first, zmin, zmax);
        return prolong3_slab_impl<double>(dims, (const double*)in, m, n3, in_first, out_first, ocount, (double*)out, s, &sd);
    }
    if (dtype == GB_F32) {
        const Sides<float> sd = make_sides<float>(lo, lo_first, hi, hi_first, zmin, zmax);
        return prolong3_slab_impl<float>(dims, (const float*)in, m, n3, in_first, out_first, ocount, (float*)out, s, &sd);
    }
    return fail(GB_ERR_ARG, "bad dtype");
}
int prolong_slab_device(int dtype, int ndim, const i64* dims, const void* in, int dim, i64 n, i64 len, i64 in_first, i64 out_first, i64 ocount, void* out,
                        cudaStream_t s) {
    if (dtype == GB_F64) return prolong_slab_impl<double>(ndim, dims, (const double*)in, dim, n, len, in_first, out_first, ocount, (double*)out, s);
    if (dtype == GB_F32) return prolong_slab_impl<float>(ndim, dims, (const float*)in, dim, n, len, in_first, out_first, ocount, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int restrict3_device(int dtype, const i64* dims, const void* in, double scale, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return restrict3_impl<double>(dims, (const double*)in, scale, (double*)out, s);
    if (dtype == GB_F32) return restrict3_impl<float>(dims, (const float*)in, scale, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int prolong3_device(int dtype, const i64* dims, const void* in, const i64* odims, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return prolong3_impl<double>(dims, (const double*)in, odims, (double*)out, s);
    if (dtype == GB_F32) return prolong3_impl<float>(dims, (const float*)in, odims, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int synth_device(int dtype, void* out, i64 n, u64 seed, double lo, double hi, i64 first, cudaStream_t s) {
    if (n <= 0) return GB_OK;
    const int grid = grid_for(n, 256);
    if (dtype == GB_F64) synth_kernel<double><<<grid, 256, 0, s>>>((double*)out, n, seed, lo, hi, first);
    else if (dtype == GB_F32) synth_kernel<float><<<grid, 256, 0, s>>>((float*)out, n, seed, lo, hi, first);
    else return fail(GB_ERR_ARG, "bad dtype");
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

}  // namespace gb
