// prefilter.cu -- interp_invert! (src/interp.jl:909-1014) and pad1 (:1287-1310) on sm_100a.
//
// The system matrix of a sweep is the same for every 1-D line of that dimension, so the
// tridiagonal LU factors (the dgttrf transliteration Julia Base uses, linalg/lu.jl) and, where the
// reference needs one, the 2x2 Woodbury capacitance matrix Cp are computed ONCE per sweep on the
// host in O(n) and uploaded; the device does the per-line substitutions (dgtts2 order) for all
// lines in parallel.  The substitution order is the reference's, so results are bit-identical
// to the CPU oracle.
//
// Kernel shape: one thread per line (the substitutions are sequential along a line and must keep the
// reference's order), all lines of a sweep resident at once; a line is streamed in chunks of 8
// elements with the next chunk's loads in flight while the current one is computed (line IO
// policies below).  Sweeps along dims >= 2: 32 adjacent lines per warp, every access a coalesced
// 256-byte row.  Sweep along dim 1: 8 lanes fetch 64 contiguous bytes of one line and a
// shared-memory tile turns the chunk so that lane r owns line r.  The pivot divisions use the
// host-computed reciprocals with FMA refinement (exact_div: same bits as a / d, 5 dependent
// operations).  The factors are bit-constant a few rows in (LineSys f0..b1): chunks inside that run use register
// constants instead of per-element factor loads.  Measured on 258^3 f64 (cubic, Woodbury): 0.17 ms per sweep along
// dims 2/3, 0.26 ms along dim 1.  A variant that kept
// whole lines in shared memory (16 B/point of HBM traffic) was built and measured at 0.5 ms per
// sweep -- too few lines per SM in flight to hide the dependent FP64 chain -- and dropped.
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace gb {

namespace {

template <typename T> struct LineSys {
    const T* dl;   // n-1 multipliers
    const T* d;    // n
    const T* rd;   // n: correctly rounded 1/d[i] (for exact_div)
    const T* du;   // n-1
    const T* du2;  // n-2
    const int* swap;  // n-1 flags: row interchange at step i (ipiv[i] == i+1)
    int woodbury;
    int vcol0, vcol1;
    T cp[4];  // column-major 2x2
    // Stationary interior.  The factors of these (Toeplitz but for the ends) matrices reach a floating-point fixed
    // point a few rows in: forward steps [f0, f1) have swap == 0 and dl == cdl, backward rows [b0, b1) have
    // du == cdu, du2 == cdu2, d == cd (bit for bit, found by the host).  Chunks inside these ranges run the same
    // operations on register constants instead of six coefficient loads per element.
    i64 f0, f1, b0, b1;
    T cdl, cdu, cdu2, cd, crd;
};

// ---- line IO policies -------------------------------------------------------------------------------
// The substitutions are sequential along a line, so a thread hides memory latency only by having the
// NEXT elements of its line in flight: lines are processed in chunks of U elements; fetch() issues
// the global loads of a chunk into a staging register array (while the previous chunk is being
// computed), unpack() hands the chunk to the owning thread, store() writes a finished chunk back.
//   GlobalLine  element i of the line at b[i*s]; for sweeps along dims >= 2, where the 32 lines of a
//               warp are adjacent in memory and every access is coalesced.  unpack is a copy.
//   TileLine    contiguous lines (sweep along dim 1): U lanes read U consecutive elements of one
//               line, so a warp load covers 32/U lines and U loads cover the warp's 32 lines;
//               a shared-memory tile (odd pitch) turns the chunk so that lane r owns line r.
template <typename T, int U> struct GlobalLine {
    T* b;
    i64 s;
    __device__ __forceinline__ void fetch(T (&g)[U], i64 start, int dir, int cnt) const {
#pragma unroll
        for (int u = 0; u < U; u++)
            if (u < cnt) g[u] = b[(start + dir * u) * s];
    }
    __device__ __forceinline__ void unpack(T (&v)[U], const T (&g)[U]) const {
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = g[u];
    }
    __device__ __forceinline__ void store(const T (&v)[U], i64 start, int dir, int cnt) const {
#pragma unroll
        for (int u = 0; u < U; u++)
            if (u < cnt) b[(start + dir * u) * s] = v[u];
    }
    __device__ __forceinline__ T get(i64 i) const { return b[i * s]; }
    __device__ __forceinline__ void put(i64 i, T v) const { b[i * s] = v; }
};
template <typename T, int U> struct TileLine {  // U in {8, 16}; all 32 lanes of the warp call together
    static constexpr int LP = 32 / U;  // lines covered by one warp load (U lanes x sizeof(T) contiguous bytes per line)
    T* base;   // line 0 of this warp
    i64 n;     // line length (line r starts at base + r*n)
    int nl;    // lines this warp really has (<= 32)
    T* tile;   // shared memory [32][U+1]
    __device__ __forceinline__ void fetch(T (&g)[U], i64 start, int dir, int cnt) const {
        const int lane = threadIdx.x & 31, e = lane % U, l0 = lane / U;
#pragma unroll
        for (int r = 0; r < U; r++) {
            const int line = l0 + LP * r;
            if (line < nl && e < cnt) g[r] = base[(i64)line * n + start + dir * e];
        }
    }
    __device__ __forceinline__ void unpack(T (&v)[U], const T (&g)[U]) const {
        const int lane = threadIdx.x & 31, e = lane % U, l0 = lane / U;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < U; r++) tile[(l0 + LP * r) * (U + 1) + e] = g[r];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = tile[lane * (U + 1) + u];
    }
    __device__ __forceinline__ void store(const T (&v)[U], i64 start, int dir, int cnt) const {
        const int lane = threadIdx.x & 31, e = lane % U, l0 = lane / U;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; u++) tile[lane * (U + 1) + u] = v[u];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < U; r++) {
            const int line = l0 + LP * r;
            if (line < nl && e < cnt) base[(i64)line * n + start + dir * e] = tile[line * (U + 1) + e];
        }
    }
    __device__ __forceinline__ T get(i64 i) const {
        const int lane = threadIdx.x & 31;
        return lane < nl ? base[(i64)lane * n + i] : (T)0;
    }
    __device__ __forceinline__ void put(i64 i, T v) const {
        const int lane = threadIdx.x & 31;
        if (lane < nl) base[(i64)lane * n + i] = v;
    }
};

// ---- the substitutions: A_ldiv_B!(::LU{T,Tridiagonal{T}}, B), Julia Base linalg/lu.jl ("See dgtts2.f").
// forward elimination of a dense RHS held in `io` (in place)
template <typename T, int U, typename IO> __device__ __forceinline__ T lu_forward_dense(const IO& io, i64 n, const LineSys<T>& sys) {
    T cur = io.get(0);
    T g[U];
    if (n > 1) io.fetch(g, 1, +1, (int)min((i64)U, n - 1));
    for (i64 c0 = 0; c0 < n - 1; c0 += U) {
        const int cnt = (int)min((i64)U, n - 1 - c0);
        T v[U];
        io.unpack(v, g);
        if (c0 + U < n - 1) io.fetch(g, c0 + U + 1, +1, (int)min((i64)U, n - 1 - c0 - U));  // next chunk in flight
        if (cnt == U && c0 >= sys.f0 && c0 + U <= sys.f1) {  // stationary chunk
            const T m = sys.cdl;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const T nxt = v[u];
                v[u] = cur;
                cur = nxt - m * cur;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (u < cnt) {
                    const T m = __ldg(sys.dl + c0 + u);
                    const T nxt = v[u];
                    if (__ldg(sys.swap + c0 + u)) { v[u] = nxt; const T t = cur - m * nxt; cur = t; }
                    else { v[u] = cur; const T t = nxt - m * cur; cur = t; }
                }
            }
        }
        io.store(v, c0, +1, cnt);
    }
    return cur;  // the eliminated last element
}
// forward elimination of the Woodbury RHS k0*e_0 + k1*e_{n-1}; results go to `out`
template <typename T, int U, typename IO> __device__ __forceinline__ T lu_forward_sparse(const IO& out, i64 n, const LineSys<T>& sys, T k0, T k1) {
    T cur = k0;
    for (i64 c0 = 0; c0 < n - 1; c0 += U) {
        const int cnt = (int)min((i64)U, n - 1 - c0);
        T v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (u < cnt) {
                const T m = __ldg(sys.dl + c0 + u);
                const T nxt = (c0 + u + 1 == n - 1) ? k1 : (T)0;
                if (__ldg(sys.swap + c0 + u)) { v[u] = nxt; const T t = cur - m * nxt; cur = t; }
                else { v[u] = cur; const T t = nxt - m * cur; cur = t; }
            } else
                v[u] = (T)0;
        }
        out.store(v, c0, +1, cnt);
    }
    return cur;
}
// back substitution reading the eliminated RHS from `rhs`; SUB == false: results overwrite `dst`;
// SUB == true (Woodbury): dst[i] = dst[i] - x[i], i.e. B = tmpN1 - tmpN2 fused into the sweep.
template <typename T, int U, bool SUB, typename IO> __device__ __forceinline__ void lu_backward(const IO& rhs, const IO& dst, i64 n, const LineSys<T>& sys, T last) {
    T x1 = exact_div(last, __ldg(sys.d + (n - 1)), __ldg(sys.rd + (n - 1)));
    {
        const T o = SUB ? dst.get(n - 1) : (T)0;
        dst.put(n - 1, SUB ? o - x1 : x1);
    }
    if (n < 2) return;
    const T x0 = exact_div(rhs.get(n - 2) - __ldg(sys.du + (n - 2)) * x1, __ldg(sys.d + (n - 2)), __ldg(sys.rd + (n - 2)));
    {
        const T o = SUB ? dst.get(n - 2) : (T)0;
        dst.put(n - 2, SUB ? o - x0 : x0);
    }
    T x2 = x1;
    x1 = x0;
    T g[U], go[SUB ? U : 1];
    if (n > 2) {
        const int c = (int)min((i64)U, n - 2);
        rhs.fetch(g, n - 3, -1, c);
        if constexpr (SUB) dst.fetch(go, n - 3, -1, c);
    }
    for (i64 c0 = n - 3; c0 >= 0; c0 -= U) {
        const int cnt = (int)min((i64)U, c0 + 1);
        T v[U], o[SUB ? U : 1];
        rhs.unpack(v, g);
        if constexpr (SUB) dst.unpack(o, go);
        if (c0 - U >= 0) {  // next chunk in flight
            const int c = (int)min((i64)U, c0 - U + 1);
            rhs.fetch(g, c0 - U, -1, c);
            if constexpr (SUB) dst.fetch(go, c0 - U, -1, c);
        }
        if (cnt == U && c0 - U + 1 >= sys.b0 && c0 < sys.b1) {  // stationary chunk
#pragma unroll
            for (int u = 0; u < U; u++) {
                const T x = exact_div(v[u] - sys.cdu * x1 - sys.cdu2 * x2, sys.cd, sys.crd);
                if constexpr (SUB) v[u] = o[u] - x;
                else v[u] = x;
                x2 = x1;
                x1 = x;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (u < cnt) {
                    const i64 i = c0 - u;
                    const T x = exact_div(v[u] - __ldg(sys.du + i) * x1 - __ldg(sys.du2 + i) * x2, __ldg(sys.d + i), __ldg(sys.rd + i));
                    if constexpr (SUB) v[u] = o[u] - x;
                    else v[u] = x;
                    x2 = x1;
                    x1 = x;
                }
            }
        }
        dst.store(v, c0, -1, cnt);
    }
}

// ---- Woodbury second solve without the array-sized scratch -----------------------------------------
// The eliminated right-hand side t of the second solve (RHS k0*e_0 + k1*e_{n-1}) is a one-multiply
// recurrence, so instead of writing it to global memory on the way up and reading it back on the way
// down (16 of the sweep's 64 B/point), the forward pass only CHECKPOINTS the recurrence state at the
// start of every backward chunk (shared memory, [chunk][thread]) and the backward pass replays the U
// steps of a chunk from its checkpoint -- the same operations in the same order on the same inputs,
// hence the same bits.
template <typename T> __device__ __forceinline__ void sparse_step_stationary(const LineSys<T>& sys, T& cur, T& v) {  // f1 <= n-2: nxt == 0
    v = cur;
    cur = (T)0 - sys.cdl * cur;
}
template <typename T> __device__ __forceinline__ void sparse_step(const LineSys<T>& sys, i64 i, i64 n, T k1, T& cur, T& v) {
    const T m = __ldg(sys.dl + i);
    const T nxt = (i + 1 == n - 1) ? k1 : (T)0;
    if (__ldg(sys.swap + i)) { v = nxt; cur = cur - m * nxt; }
    else { v = cur; cur = nxt - m * cur; }
}
template <typename T, int U> __device__ __forceinline__ T lu_forward_sparse_ck(i64 n, const LineSys<T>& sys, T k0, T k1, T* ck, int ckstride, T& t_nm2) {
    T cur = k0;
    t_nm2 = (T)0;
    const i64 nck = (n - 2 + U - 1) / U;  // backward chunks cover indices 0 .. n-3
    i64 j = nck - 1;
    i64 sj = 0;  // start index of chunk j: max(0, n-2-(j+1)U); the lowest chunk starts at 0
    for (i64 i = 0; i < n - 1; i++) {
        if (j >= 0 && i == sj) {
            ck[j * ckstride] = cur;
            j--;
            sj = n - 2 - (j + 1) * U;
        }
        T v;
        if (i >= sys.f0 && i < sys.f1) sparse_step_stationary<T>(sys, cur, v);
        else sparse_step<T>(sys, i, n, k1, cur, v);
        if (i == n - 2) t_nm2 = v;
    }
    return cur;
}
template <typename T, int U, typename IO>
__device__ __forceinline__ void lu_backward_sub_ck(const IO& dst, i64 n, const LineSys<T>& sys, T last, T t_nm2, T k1, const T* ck, int ckstride) {
    T x1 = exact_div(last, __ldg(sys.d + (n - 1)), __ldg(sys.rd + (n - 1)));
    dst.put(n - 1, dst.get(n - 1) - x1);
    if (n < 2) return;
    const T x0 = exact_div(t_nm2 - __ldg(sys.du + (n - 2)) * x1, __ldg(sys.d + (n - 2)), __ldg(sys.rd + (n - 2)));
    dst.put(n - 2, dst.get(n - 2) - x0);
    T x2 = x1;
    x1 = x0;
    T go[U];
    if (n > 2) dst.fetch(go, n - 3, -1, (int)min((i64)U, n - 2));
    i64 j = 0;
    for (i64 c0 = n - 3; c0 >= 0; c0 -= U, j++) {
        const int cnt = (int)min((i64)U, c0 + 1);
        T v[U], o[U], tt[U];
        dst.unpack(o, go);
        if (c0 - U >= 0) dst.fetch(go, c0 - U, -1, (int)min((i64)U, c0 - U + 1));  // next chunk in flight
        // replay the chunk's U forward steps from its checkpoint (positions below index 0 are skipped)
        T cur = ck[j * ckstride];
        const i64 s = c0 - U + 1;
        if (s >= sys.f0 && s + U <= sys.f1) {
#pragma unroll
            for (int u = 0; u < U; u++) sparse_step_stationary<T>(sys, cur, tt[u]);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                tt[u] = (T)0;
                if (s + u >= 0) sparse_step<T>(sys, s + u, n, k1, cur, tt[u]);
            }
        }
        if (cnt == U && s >= sys.b0 && c0 < sys.b1) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const T x = exact_div(tt[U - 1 - u] - sys.cdu * x1 - sys.cdu2 * x2, sys.cd, sys.crd);
                v[u] = o[u] - x;
                x2 = x1;
                x1 = x;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (u < cnt) {
                    const i64 i = c0 - u;
                    const T x = exact_div(tt[U - 1 - u] - __ldg(sys.du + i) * x1 - __ldg(sys.du2 + i) * x2, __ldg(sys.d + i), __ldg(sys.rd + i));
                    v[u] = o[u] - x;
                    x2 = x1;
                    x1 = x;
                }
            }
        }
        dst.store(v, c0, -1, cnt);
    }
}

template <typename T, int U, typename IO> __device__ __forceinline__ void solve_line(const IO& b, const IO& t, i64 n, const LineSys<T>& sys, T* ck, int ckstride) {
    const T last = lu_forward_dense<T, U>(b, n, sys);
    lu_backward<T, U, false>(b, b, n, sys, last);
    if (sys.woodbury) {                                            // WoodburyMatrices A_ldiv_B!
        const T k10 = b.get(sys.vcol0), k11 = b.get(sys.vcol1);    // tmpk1 = V*tmpN1
        const T k20 = ((T)0 + sys.cp[0] * k10) + sys.cp[2] * k11;  // tmpk2 = Cp*tmpk1 (gemv column order)
        const T k21 = ((T)0 + sys.cp[1] * k10) + sys.cp[3] * k11;
        if (ck && n >= 3) {  // A_ldiv_B!(W.A, tmpN2 = U*tmpk2), B = tmpN1 - tmpN2; tmpN2 never leaves the chip
            T t_nm2;
            const T last2 = lu_forward_sparse_ck<T, U>(n, sys, k20, k21, ck, ckstride, t_nm2);
            lu_backward_sub_ck<T, U>(b, n, sys, last2, t_nm2, k21, ck, ckstride);
        } else {
            const T last2 = lu_forward_sparse<T, U>(t, n, sys, k20, k21);
            lu_backward<T, U, true>(t, b, n, sys, last2);
        }
    }
}

// sweeps along dims >= 2: one thread per line, 32 adjacent lines per warp -> coalesced chunk IO
constexpr int PF_U = 8;
template <typename T> __global__ void __launch_bounds__(128, 4) line_solve_global(T* A, T* tmp, i64 n, i64 s, i64 inner, i64 nlines, const LineSys<T> sys, int use_ck) {
    const i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const i64 lo = l % inner, hi = l / inner;
    const i64 base = lo + hi * inner * n;
    const GlobalLine<T, PF_U> b{A + base, s}, t{tmp ? tmp + base : nullptr, s};
    extern __shared__ __align__(16) unsigned char pf_dyn[];
    T* ck = use_ck ? reinterpret_cast<T*>(pf_dyn) + threadIdx.x : nullptr;
    solve_line<T, PF_U>(b, t, n, sys, ck, (int)blockDim.x);
}
// sweep along dim 1 (contiguous lines): one warp per 32 lines, shared-memory transposing tile
constexpr int PF_WARPS = 4;
template <typename T, int TU> __global__ void __launch_bounds__(32 * PF_WARPS, TU == 8 ? 4 : 3) line_solve_tile(T* A, T* tmp, i64 n, i64 nlines, const LineSys<T> sys, int use_ck) {
    __shared__ T tiles[PF_WARPS][32 * (TU + 1)];
    const int warp = threadIdx.x >> 5;
    const i64 first = ((i64)blockIdx.x * PF_WARPS + warp) * 32;
    if (first >= nlines) return;
    const int nl = (int)min((i64)32, nlines - first);
    const TileLine<T, TU> b{A + first * n, n, nl, tiles[warp]}, t{tmp ? tmp + first * n : nullptr, n, nl, tiles[warp]};
    extern __shared__ __align__(16) unsigned char pf_dyn[];
    T* ck = use_ck ? reinterpret_cast<T*>(pf_dyn) + threadIdx.x : nullptr;
    solve_line<T, TU>(b, t, n, sys, ck, (int)blockDim.x);
}

// ---- host: factorisation (dgttrf order) + Woodbury capacitance ---------------------------------
template <typename T> struct HostSys {
    std::vector<T> dl, d, du, du2;
    std::vector<int> swap;
    bool woodbury = false;
    i64 vcol[2] = {0, 0};
    T cp[4] = {0, 0, 0, 0};
    i64 n = 0;

    void factor() {
        du2.assign(n > 2 ? n - 2 : 0, (T)0);
        swap.assign(n > 1 ? n - 1 : 0, 0);
        for (i64 i = 0; i + 1 < n; i++) {
            const bool last = (i == n - 2);
            if (std::fabs(d[i]) >= std::fabs(dl[i])) {
                if (d[i] != (T)0) {
                    T fact = dl[i] / d[i];
                    dl[i] = fact;
                    d[i + 1] = d[i + 1] - fact * du[i];
                    if (!last) du2[i] = (T)0;
                }
            } else {
                T fact = d[i] / dl[i];
                d[i] = dl[i];
                dl[i] = fact;
                T tmp = du[i];
                du[i] = d[i + 1];
                d[i + 1] = tmp - fact * d[i + 1];
                if (!last) {
                    du2[i] = du[i + 1];
                    du[i + 1] = -fact * du[i + 1];
                }
                swap[i] = 1;
            }
        }
    }
    void solve(T* b) const {
        for (i64 i = 0; i + 1 < n; i++) {
            T cur = b[i], nxt = b[i + 1];
            if (swap[i]) { b[i] = nxt; b[i + 1] = cur - dl[i] * nxt; }
            else b[i + 1] = nxt - dl[i] * cur;
        }
        b[n - 1] = b[n - 1] / d[n - 1];
        if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
        for (i64 i = n - 3; i >= 0; i--) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
    }
    // _interp_invert_matrix, src/interp.jl:955-1014
    int setup(int bc, int order, i64 n_) {
        n = n_;
        const int fam = bc_family(bc);
        if (order == GB_QUADRATIC) {
            if (n < 3) return fail(GB_ERR_ARG, "InterpQuadratic needs at least 3 points per dimension");
            du.assign(n - 1, (T)(1.0 / 8));
            d.assign(n, (T)(3.0 / 4));
            dl = du;
            if (fam == BCF_NIL) {
                d[0] = d[n - 1] = (T)(9.0 / 8);
                dl[n - 2] = du[0] = (T)(-1.0 / 4);
                woodbury = true;
                vcol[0] = 2; vcol[1] = n - 3;
            } else if (fam == BCF_REFLECT) {
                d[0] = (T)((double)d[0] + 1.0 / 8);
                d[n - 1] = (T)((double)d[n - 1] + 1.0 / 8);
            } else if (fam == BCF_PERIODIC) {
                woodbury = true;
                vcol[0] = n - 1; vcol[1] = 0;
            } else {
                du[0] = (T)((double)du[0] + 1.0 / 8);
                dl[n - 2] = (T)((double)dl[n - 2] + 1.0 / 8);
            }
        } else {
            if (fam != BCF_NIL) return fail(GB_ERR_UNSUPPORTED, "InterpCubic supports only BCnil/BCnan/BCna");
            if (n < 4) return fail(GB_ERR_ARG, "InterpCubic needs at least 4 points per (padded) dimension");
            du.assign(n - 1, (T)(1.0 / 6));
            d.assign(n, (T)(4.0 / 6));
            dl = du;
            d[0] = d[1] = d[n - 2] = d[n - 1] = (T)1;
            du[0] = dl[n - 2] = (T)-2;
            du[1] = dl[n - 3] = du[n - 2] = dl[0] = (T)0;
            woodbury = true;
            vcol[0] = 2; vcol[1] = n - 3;
        }
        factor();
        if (woodbury) {  // Cp = inv(inv(C) .+ V*(A\U)), WoodburyMatrices.jl constructor
            const T cinv = (order == GB_QUADRATIC) ? (T)1 / (T)(1.0 / 8) : (T)1;
            std::vector<T> u1(n, (T)0), u2(n, (T)0);
            u1[0] = (T)1; u2[n - 1] = (T)1;
            solve(u1.data());
            solve(u2.data());
            T a11 = cinv + u1[vcol[0]], a21 = (T)0 + u1[vcol[1]], a12 = (T)0 + u2[vcol[0]], a22 = cinv + u2[vcol[1]];
            // inv(::Matrix) = LAPACK getrf (partial pivoting, reciprocal scaling) + getri, unblocked
            const bool sw = std::fabs(a21) > std::fabs(a11);
            if (sw) { std::swap(a11, a21); std::swap(a12, a22); }
            const T r = (T)1 / a11;
            const T l21 = a21 * r;
            const T u22 = a22 - l21 * a12;
            const T i11 = (T)1 / a11, i22 = (T)1 / u22;
            const T i12 = (-i22) * (i11 * a12);
            T b11 = i11 - i12 * l21, b21 = (T)0 - i22 * l21, b12 = i12, b22 = i22;
            if (sw) { std::swap(b11, b12); std::swap(b21, b22); }
            cp[0] = b11; cp[1] = b21; cp[2] = b12; cp[3] = b22;
        }
        return GB_OK;
    }
};

// ---- factorisation cache ---------------------------------------------------------------------------------------------
// The factors of a sweep depend on (element type, bc family, order, n) only: computed and uploaded once per device, then
// reused by every later sweep of that shape (InterpGrid(...) on 258^3 spends more time here than in its kernels otherwise).
template <typename T> struct CachedSys {
    LineSys<T> sys;
    int woodbury = 0;
};
template <typename T> int get_line_sys(int bc, int order, i64 n, CachedSys<T>& out) {
    struct Key {
        int dev, fam, order;
        i64 n;
        bool operator==(const Key& o) const { return dev == o.dev && fam == o.fam && order == o.order && n == o.n; }
    };
    static std::mutex mu;
    static std::vector<std::pair<Key, CachedSys<T>>> cache;
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    const Key key{dev, bc_family(bc), order, n};
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : cache)
        if (e.first == key) { out = e.second; return GB_OK; }
    HostSys<T> hs;
    int rc = hs.setup(bc, order, n);
    if (rc) return rc;
    // one device blob: dl | d | du | du2 | rd | swap
    const size_t nT = (size_t)(5 * n);
    const size_t bytes = nT * sizeof(T) + (size_t)n * sizeof(int);
    std::vector<unsigned char> blob(bytes, 0);
    T* hb = reinterpret_cast<T*>(blob.data());
    std::copy(hs.dl.begin(), hs.dl.end(), hb);
    std::copy(hs.d.begin(), hs.d.end(), hb + n);
    std::copy(hs.du.begin(), hs.du.end(), hb + 2 * n);
    std::copy(hs.du2.begin(), hs.du2.end(), hb + 3 * n);
    for (i64 i = 0; i < n; i++) hb[4 * n + i] = (T)1 / hs.d[i];  // correctly rounded reciprocals of the pivots
    int* hsw = reinterpret_cast<int*>(blob.data() + nT * sizeof(T));
    std::copy(hs.swap.begin(), hs.swap.end(), hsw);
    unsigned char* dblob = nullptr;
    GB_CUDA(cudaMalloc((void**)&dblob, bytes));
    cudaError_t e = cudaMemcpy(dblob, blob.data(), bytes, cudaMemcpyHostToDevice);  // synchronous: valid for every stream afterwards
    if (e != cudaSuccess) { cudaFree(dblob); return cuda_fail(e, "cudaMemcpy (prefilter factors)"); }
    CachedSys<T> cs;
    LineSys<T>& sys = cs.sys;
    T* db = reinterpret_cast<T*>(dblob);
    sys.dl = db; sys.d = db + n; sys.du = db + 2 * n; sys.du2 = db + 3 * n; sys.rd = db + 4 * n;
    sys.swap = reinterpret_cast<const int*>(dblob + nT * sizeof(T));
    sys.woodbury = hs.woodbury ? 1 : 0;
    cs.woodbury = sys.woodbury;
    sys.vcol0 = (int)hs.vcol[0]; sys.vcol1 = (int)hs.vcol[1];
    for (int i = 0; i < 4; i++) sys.cp[i] = hs.cp[i];
    {  // stationary interior: longest runs of bit-identical factors (forward steps end before n-2: sparse_step's k1 row)
        auto same = [](T a, T b) { return std::memcmp(&a, &b, sizeof(T)) == 0; };
        auto longest = [&](i64 lo, i64 hi, auto eq, i64& r0, i64& r1) {
            r0 = r1 = 0;
            for (i64 i = lo; i < hi;) {
                i64 j = i + 1;
                while (j < hi && eq(i, j)) j++;
                if (j - i > r1 - r0) { r0 = i; r1 = j; }
                i = j;
            }
        };
        longest(0, n - 2, [&](i64 i, i64 j) { return !hs.swap[i] && !hs.swap[j] && same(hs.dl[i], hs.dl[j]); }, sys.f0, sys.f1);
        if (sys.f1 - sys.f0 < 2 || hs.swap[sys.f0]) sys.f0 = sys.f1 = 0;
        longest(0, n - 2, [&](i64 i, i64 j) { return same(hs.du[i], hs.du[j]) && same(hs.du2[i], hs.du2[j]) && same(hs.d[i], hs.d[j]); }, sys.b0, sys.b1);
        if (sys.b1 - sys.b0 < 2) sys.b0 = sys.b1 = 0;
        sys.cdl = n > 1 ? hs.dl[sys.f0] : (T)0;
        sys.cdu = n > 2 ? hs.du[sys.b0] : (T)0;
        sys.cdu2 = n > 2 ? hs.du2[sys.b0] : (T)0;
        sys.cd = hs.d[sys.b0];
        sys.crd = (T)1 / hs.d[sys.b0];
    }
    if (cache.size() >= 256) {  // bounded: forget the oldest shape
        cudaFree(const_cast<T*>(cache.front().second.sys.dl));
        cache.erase(cache.begin());
    }
    cache.emplace_back(key, cs);
    out = cs;
    return GB_OK;
}

// one exact sweep (the reference's LU / Woodbury substitution order) along dimension idim, in place
template <typename T> int sweep_exact(T* A, int ndim, const i64* dims, int bc, int order, int idim, cudaStream_t s) {
    if (idim < 0 || idim >= ndim) return fail(GB_ERR_ARG, "dimlist entry out of range");
    i64 total = 1, str = 1;
    for (int d = 0; d < ndim; d++) {
        if (d == idim) str = total;
        total *= dims[d];
    }
    if (total == 0) return GB_OK;
    const i64 n = dims[idim], nlines = total / n;
    CachedSys<T> cs;
    int rc = get_line_sys<T>(bc, order, n, cs);
    if (rc) return rc;
    const LineSys<T>& sys = cs.sys;
    // Woodbury: checkpoints of the second solve's RHS in shared memory when they fit, else an array-sized scratch
    const size_t ck_bytes = cs.woodbury && n >= 3 ? (size_t)((n - 2 + PF_U - 1) / PF_U) * 128 * sizeof(T) : 0;
    const int use_ck = cs.woodbury && ck_bytes > 0 && ck_bytes <= 64 * 1024;
    const size_t dyn = use_ck ? ck_bytes : 0;
    AsyncBuf tmp(s);
    if (cs.woodbury && !use_ck) GB_CUDA(tmp.alloc((size_t)total * sizeof(T)));
    if (str == 1 && nlines >= 32) {
        const i64 grid = (nlines + 32 * PF_WARPS - 1) / (32 * PF_WARPS);
        auto kern = line_solve_tile<T, 8>;  // 8: measured faster than 16
        if (dyn) GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<(unsigned)grid, 32 * PF_WARPS, dyn, s>>>(A, (T*)tmp.p, n, nlines, sys, use_ck);
    } else {
        const i64 grid = (nlines + 127) / 128;
        auto kern = line_solve_global<T>;
        if (dyn) GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<(unsigned)grid, 128, dyn, s>>>(A, (T*)tmp.p, n, str, str, nlines, sys, use_ck);
    }
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

template <typename T> int invert_impl(T* A, int ndim, const i64* dims, int bc, int order, const int* dimlist, int ndl, cudaStream_t s) {
    if (order < GB_QUADRATIC) return GB_OK;  // src/interp.jl:909-912
    for (int k = 0; k < ndl; k++) {
        const int rc = sweep_exact<T>(A, ndim, dims, bc, order, dimlist[k], s);
        if (rc) return rc;
    }
    return GB_OK;
}

// filledges!(A, val) (src/utilities.jl:22-33): one thread per element, only edge elements are written
struct EdgeDims { i64 d[MAXG]; int n; };
template <typename T> __global__ void __launch_bounds__(256) filledges_kernel(T* __restrict__ A, i64 total, EdgeDims e, T val) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
        i64 rem = k;
        bool edge = false;
        for (int d = 0; d < e.n; d++) {
            const i64 c = rem % e.d[d];
            rem /= e.d[d];
            edge = edge || c == 0 || c == e.d[d] - 1;
        }
        if (edge) A[k] = val;
    }
}

template <typename T, int N> struct PadParams { i64 od[N]; i64 id[N]; };
// One warp per output row (dim-0 line): the row's coordinates are decoded once, the lanes stream the row.
template <typename T, int N> __global__ void __launch_bounds__(256) pad1_kernel(const T* __restrict__ in, T* __restrict__ out, i64 orows, PadParams<T, N> p, T f) {
    const int lane = threadIdx.x & 31;
    const i64 wstep = (i64)gridDim.x * (blockDim.x >> 5);
    const int n0 = (int)p.od[0], i0 = (int)p.id[0];
    for (i64 r = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < orows; r += wstep) {
        i64 rem = r, src = 0, mul = p.id[0];
        bool inside = true;
#pragma unroll
        for (int d = 1; d < N; d++) {
            const i64 c = rem % p.od[d];
            rem /= p.od[d];
            inside = inside && c >= 1 && c <= p.id[d];
            src += (c - 1) * mul;
            mul *= p.id[d];
        }
        T* o = out + r * n0;
        const T* a = in + src - 1;  // a[x] for output column x in [1, i0]
        for (int x = lane; x < n0; x += 32) o[x] = (inside && x >= 1 && x <= i0) ? a[x] : f;
    }
}
template <typename T, int N> int pad1_n(const i64* dims, const T* in, T f, T* out, cudaStream_t s) {
    PadParams<T, N> p;
    i64 ototal = 1;
    for (int d = 0; d < N; d++) { p.id[d] = dims[d]; p.od[d] = dims[d] + 2; ototal *= p.od[d]; }
    if (p.od[0] >= (1ll << 31)) return fail(GB_ERR_UNSUPPORTED, "pad1: extent too large");
    const i64 orows = ototal / p.od[0];
    const i64 need = (orows + 7) / 8;
    const int grid = (int)std::max<i64>(1, std::min<i64>(need, (i64)num_sms() * 32));
    pad1_kernel<T, N><<<grid, 256, 0, s>>>(in, out, orows, p, f);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
template <typename T> int pad1_impl(int ndim, const i64* dims, const T* in, T f, T* out, cudaStream_t s) {
    switch (ndim) {
    case 1: return pad1_n<T, 1>(dims, in, f, out, s);
    case 2: return pad1_n<T, 2>(dims, in, f, out, s);
    case 3: return pad1_n<T, 3>(dims, in, f, out, s);
    case 4: return pad1_n<T, 4>(dims, in, f, out, s);
    case 5: return pad1_n<T, 5>(dims, in, f, out, s);
    case 6: return pad1_n<T, 6>(dims, in, f, out, s);
    case 7: return pad1_n<T, 7>(dims, in, f, out, s);
    case 8: return pad1_n<T, 8>(dims, in, f, out, s);
    }
    return fail(GB_ERR_UNSUPPORTED, "pad1: ndim must be 1..8");
}

}  // namespace

int invert_device(int dtype, int ndim, const i64* dims, void* data, int bc, int order, const int* dimlist, int ndl, cudaStream_t s) {
    if (dtype == GB_F64) return invert_impl<double>((double*)data, ndim, dims, bc, order, dimlist, ndl, s);
    if (dtype == GB_F32) return invert_impl<float>((float*)data, ndim, dims, bc, order, dimlist, ndl, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int invert_sweep_exact(int dtype, int ndim, const i64* dims, void* data, int bc, int order, int idim, cudaStream_t s) {
    if (order < GB_QUADRATIC) return GB_OK;
    if (dtype == GB_F64) return sweep_exact<double>((double*)data, ndim, dims, bc, order, idim, s);
    if (dtype == GB_F32) return sweep_exact<float>((float*)data, ndim, dims, bc, order, idim, s);
    return fail(GB_ERR_ARG, "bad dtype");
}
int filledges_device(int dtype, int ndim, const i64* dims, void* data, double val, cudaStream_t s) {
    if (ndim < 1 || ndim > MAXG) return fail(GB_ERR_UNSUPPORTED, "filledges: ndim must be 1..8");
    EdgeDims e;
    e.n = ndim;
    i64 total = 1;
    for (int d = 0; d < ndim; d++) { e.d[d] = dims[d]; total *= dims[d]; }
    if (total <= 0) return GB_OK;
    const int grid = (int)std::max<i64>(1, std::min<i64>((total + 255) / 256, (i64)num_sms() * 32));
    if (dtype == GB_F64) filledges_kernel<double><<<grid, 256, 0, s>>>((double*)data, total, e, val);
    else if (dtype == GB_F32) filledges_kernel<float><<<grid, 256, 0, s>>>((float*)data, total, e, (float)val);
    else return fail(GB_ERR_ARG, "bad dtype");
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
int pad1_device(int dtype, int ndim, const i64* dims, const void* in, double fill, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return pad1_impl<double>(ndim, dims, (const double*)in, fill, (double*)out, s);
    if (dtype == GB_F32) return pad1_impl<float>(ndim, dims, (const float*)in, (float)fill, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}

}  // namespace gb
