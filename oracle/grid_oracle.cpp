// grid_oracle.cpp -- CPU restatement of timholy/Grid.jl's evaluation / prefilter /
// restrict-prolong algorithms.
//
// *** TEST INFRASTRUCTURE ONLY ***  Nothing in the product path (grid.jl_b200/, the C-ABI
// library) links, imports or executes this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load liboracle.so, and there only as
// the checker / the CPU baseline.
//
// Parity status: the evaluation path (index rules, weights, product association, summation
// order, zero-weight skip) and restrict/prolong are pinned by the reference's own known-answer
// tests (tests/test_oracle_kats.py replays test/grid.jl, test/cubic_interpolation.jl,
// test/coord.jl and the README values).  The LAST BITS of the prefilter are "parity unpinned":
// they depend on Julia Base's tridiagonal LU, on the un-vendored WoodburyMatrices.jl (REQUIRE:3,
// no version pinned) and on the LAPACK/BLAS build behind `inv` and `gemv`; this file restates
// the published algorithms (dgttrf/dgtts2 as transliterated in Julia Base linalg/lu.jl,
// Woodbury's A_ldiv_B!, unblocked dgetf2/dgetri) without FMA contraction.  Likewise the
// restrict/prolong summation order is the order of the reference's axpy! calls, unfused.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
// All citations are file:line under /root/reference/.
//
// Documented resolutions of undefined behaviour in the reference (shared with the CUDA path,
// see DESIGN.md "Divergences"):
//   D1  a tap whose flat index falls outside the coefficient array is skipped (the reference
//       reads out of bounds under @inbounds, src/interp.jl:504).
//   D2  InterpCubic: stored coordinate x < 2 is invalid (reference accepts [1,2) and then
//       dereferences index 0).
//   D3  InterpNearest with BCnil/nan/na: the rounded index is clamped to [1,len] (x = 0.5 or
//       len+0.5 round to 0 / len+1 in the reference).
//   D4  non-finite or |x| >= 2^62 coordinates are invalid (Julia throws InexactError).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

enum { BC_NIL = 0, BC_NAN = 1, BC_NA = 2, BC_REFLECT = 3, BC_PERIODIC = 4, BC_NEAREST = 5, BC_FILL = 6 };
enum { IT_NEAREST = 0, IT_LINEAR = 1, IT_QUADRATIC = 2, IT_CUBIC = 3 };
enum { OUT_VAL = 1, OUT_GRAD = 2, OUT_HESS = 4 };
enum { ERR_OK = 0, ERR_ARG = 1, ERR_INVALID_LOCATION = 2, ERR_UNSUPPORTED = 3 };

typedef int64_t i64;

inline bool nilfam(int bc) { return bc == BC_NIL || bc == BC_NAN || bc == BC_NA; }

// Julia's mod (floored).  src/boundaryconditions.jl:23,25 use it.
inline i64 jmod(i64 a, i64 m) {
    i64 r = a % m;
    if (r != 0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

// src/boundaryconditions.jl:22-26
inline i64 wrap(int bc, i64 pos, i64 len) {
    switch (bc) {
    case BC_REFLECT: {
        i64 r = jmod(pos - 1, 2 * len);
        return r < len ? r + 1 : 2 * len - r;
    }
    case BC_PERIODIC: return jmod(pos - 1, len) + 1;
    case BC_NEAREST:
    case BC_FILL: return pos < 1 ? 1 : (pos > len ? len : pos);
    default: return pos;
    }
}

// src/boundaryconditions.jl:20-21
inline bool isvalid_pos(int bc, i64 pos, i64 mn, i64 mx) {
    if (bc == BC_NAN || bc == BC_NA) return mn <= pos && pos <= mx;
    return true;
}

inline int npoints(int order) { return order + 1; }  // src/interp.jl:577-582

template <typename T> inline bool representable(T x) {  // D4
    return std::isfinite(x) && std::fabs((double)x) < 4611686018427387904.0;
}

// src/interp.jl:585-591 (+D2, D4)
template <typename T> bool isvalid_x(int bc, int order, T x, i64 len) {
    if (!representable(x)) return false;
    if (!nilfam(bc)) return true;
    switch (order) {
    case IT_NEAREST: {
        // x >= 0.5 && x-0.5 <= len ; for Float32 the literal promotes to Float64
        double xm = std::is_same<T, float>::value ? (double)x - 0.5 : (double)(T)(x - (T)0.5);
        return (double)x >= 0.5 && xm <= (double)len;
    }
    case IT_LINEAR:
    case IT_QUADRATIC: return x >= (T)1 && (double)x <= (double)len;
    case IT_CUBIC: return (double)x >= 2.0 && (double)x <= (double)(len - 1);  // D2: ref says 1 <= x
    }
    return false;
}

// src/interp.jl:594-783.  Returns ix; fills coord[0..l-1] (1-based grid coordinates), dx, iswrap.
template <typename T> i64 coords_1d(int bc, int order, T x, i64 len, i64* coord, T& dx, bool& iswrap) {
    switch (order) {
    case IT_NEAREST: {  // :620-624
        i64 ix = wrap(bc, (i64)std::nearbyint(x), len);
        if (nilfam(bc)) ix = ix < 1 ? 1 : (ix > len ? len : ix);  // D3
        coord[0] = ix;
        dx = (T)0;
        iswrap = false;
        return ix;
    }
    case IT_LINEAR: {
        if (bc == BC_REFLECT) {  // :642-663
            i64 fx = (i64)std::floor(x);
            i64 ix = jmod(fx - 1, 2 * len);
            dx = x - (T)fx;
            i64 ixp;
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
            coord[0] = ix;
            coord[1] = wrap(BC_REFLECT, ixp, len);
            iswrap = (ix == ixp && dx > (T)0);
            return ix;
        }
        // :632-641
        i64 ifx = (i64)std::floor(x);
        dx = x - (T)ifx;
        ifx -= (x < (T)ifx) ? 1 : 0;
        i64 ix = wrap(bc, ifx, len);
        coord[0] = ix;
        iswrap = (ix == len && dx > (T)0) || (ix == 1 && ifx < 1);
        coord[1] = iswrap ? ix : wrap(bc, ix + 1, len);
        return ix;
    }
    case IT_QUADRATIC: {
        if (nilfam(bc)) {  // :689-702
            i64 ix;
            // x+0.5: Float64 arithmetic for either T (literal 0.5 is Float64)
            double xp = (double)x + 0.5;
            if ((double)x > 1.5 && xp < (double)len) ix = (i64)std::nearbyint(x);
            else if ((double)x <= 1.5) ix = 2;
            else ix = len - 1;
            dx = x - (T)ix;
            coord[0] = ix - 1;
            coord[1] = ix;
            coord[2] = ix + 1;
            iswrap = false;
            return ix;
        }
        if (bc == BC_NEAREST || bc == BC_FILL) {  // :704-736
            double xp = (double)x + 0.5;
            i64 ix;
            if ((double)x > 1.5 && xp < (double)len) {
                ix = (i64)std::nearbyint(x);
                coord[0] = ix - 1;
                coord[1] = ix;
                coord[2] = ix + 1;
                iswrap = false;
                dx = x - (T)ix;
            } else if ((double)x <= 1.5) {
                ix = 1;
                coord[0] = 2;
                coord[1] = 1;
                coord[2] = 2;
                iswrap = true;
                dx = (x < (T)1) ? (T)0 : x - (T)ix;
            } else {
                ix = len;
                coord[0] = len - 1;
                coord[1] = len;
                coord[2] = len - 1;
                iswrap = true;
                dx = ((double)x > (double)len) ? (T)0 : x - (T)ix;
            }
            return ix;
        }
        if (bc == BC_REFLECT) {  // :737-757
            i64 ix = (i64)std::nearbyint(x);
            dx = x - (T)ix;
            ix = jmod(ix - 1, 2 * len);
            if (ix < len) ix += 1;
            else {
                dx = -dx;
                ix = 2 * len - ix;
            }
            coord[1] = ix;
            iswrap = (ix == 1 || ix == len);
            if (iswrap) {
                coord[0] = wrap(BC_REFLECT, ix - 1, len);
                coord[2] = wrap(BC_REFLECT, ix + 1, len);
            } else {
                coord[0] = ix - 1;
                coord[2] = ix + 1;
            }
            return ix;
        }
        // generic :672-686 (reached by BCperiodic only)
        i64 ix = (i64)std::nearbyint(x);
        dx = x - (T)ix;
        ix = wrap(bc, ix, len);
        coord[1] = ix;
        iswrap = (ix == 1 || ix == len);
        if (iswrap) {
            coord[0] = wrap(bc, ix - 1, len);
            coord[2] = wrap(bc, ix + 1, len);
        } else {
            coord[0] = ix - 1;
            coord[2] = ix + 1;
        }
        return ix;
    }
    case IT_CUBIC: {  // :767-783
        i64 ix = (i64)std::trunc(x);
        dx = x - (T)ix;
        coord[1] = ix;
        iswrap = isvalid_pos(bc, ix, 2, len - 1);
        coord[0] = wrap(bc, ix - 1, len);
        coord[2] = wrap(bc, ix + 1, len);
        coord[3] = wrap(bc, ix + 2, len);
        return ix;
    }
    }
    return 0;
}

// "version for offsets" src/interp.jl:594-596,616-618,627-630,666-670,760-765
inline void offsets_1d(int order, i64* c) {
    switch (order) {
    case IT_NEAREST: c[0] = 0; break;
    case IT_LINEAR: c[0] = 0; c[1] = 1; break;
    case IT_QUADRATIC: c[0] = -1; c[1] = 0; c[2] = 1; break;
    case IT_CUBIC: c[0] = -1; c[1] = 0; c[2] = 1; c[3] = 2; break;
    }
}

// src/interp.jl:785-847.  Julia promotion rules are spelled out: a Float64 literal (0.5, 0.75,
// 2/3, 1/2, 3/2, 1.0) promotes a Float32 operand to Float64 and the store rounds back to T;
// Int literals keep T.  x^2 = x*x, x^3 = (x*x)*x (powi).
inline void coefs_1d(int order, double dx, double* c) {
    switch (order) {
    case IT_NEAREST: c[0] = 1; break;
    case IT_LINEAR: c[0] = 1.0 - dx; c[1] = dx; break;
    case IT_QUADRATIC: {
        double a = dx - 0.5, b = dx + 0.5;
        c[0] = (a * a) / 2;
        c[1] = 0.75 - dx * dx;
        c[2] = (b * b) / 2;
        break;
    }
    case IT_CUBIC: {
        double t = 1 - dx;
        c[0] = ((t * t) * t) / 6;
        c[1] = 2.0 / 3.0 - ((dx * dx) * (2 - dx)) / 2;
        double u = dx - 1;
        c[2] = 2.0 / 3.0 - ((u * u) * (1 + dx)) / 2;
        c[3] = ((dx * dx) * dx) / 6;
        break;
    }
    }
}
inline void coefs_1d(int order, float dx, float* c) {
    switch (order) {
    case IT_NEAREST: c[0] = 1; break;
    case IT_LINEAR: c[0] = (float)(1.0 - (double)dx); c[1] = dx; break;
    case IT_QUADRATIC: {
        double a = (double)dx - 0.5, b = (double)dx + 0.5;
        c[0] = (float)((a * a) / 2);
        float d2 = dx * dx;
        c[1] = (float)(0.75 - (double)d2);
        c[2] = (float)((b * b) / 2);
        break;
    }
    case IT_CUBIC: {
        float t = 1.0f - dx;
        c[0] = ((t * t) * t) / 6.0f;
        float p = ((dx * dx) * (2.0f - dx)) / 2.0f;
        c[1] = (float)(2.0 / 3.0 - (double)p);
        float u = dx - 1.0f;
        float q = ((u * u) * (1.0f + dx)) / 2.0f;
        c[2] = (float)(2.0 / 3.0 - (double)q);
        c[3] = ((dx * dx) * dx) / 6.0f;
        break;
    }
    }
}
inline void gcoefs_1d(int order, double dx, double* c) {
    switch (order) {
    case IT_NEAREST: c[0] = 0; break;
    case IT_LINEAR: c[0] = -1; c[1] = 1; break;
    case IT_QUADRATIC: c[0] = dx - 0.5; c[1] = -2 * dx; c[2] = dx + 0.5; break;
    case IT_CUBIC: {
        double t = 1 - dx;
        c[0] = (-(t * t)) / 2;
        c[1] = (3 * (dx * dx)) / 2 - 2 * dx;
        c[2] = 0.5 + dx * (1 - 1.5 * dx);
        c[3] = (dx * dx) / 2;
        break;
    }
    }
}
inline void gcoefs_1d(int order, float dx, float* c) {
    switch (order) {
    case IT_NEAREST: c[0] = 0; break;
    case IT_LINEAR: c[0] = -1; c[1] = 1; break;
    case IT_QUADRATIC:
        c[0] = (float)((double)dx - 0.5);
        c[1] = -2.0f * dx;
        c[2] = (float)((double)dx + 0.5);
        break;
    case IT_CUBIC: {
        float t = 1.0f - dx;
        c[0] = (-(t * t)) / 2.0f;
        c[1] = (3.0f * (dx * dx)) / 2.0f - 2.0f * dx;
        double d = (double)dx;
        c[2] = (float)(0.5 + d * (1 - 1.5 * d));
        c[3] = (dx * dx) / 2.0f;
        break;
    }
    }
}
template <typename T> inline void hcoefs_1d(int order, T dx, T* c) {
    switch (order) {
    case IT_QUADRATIC: c[0] = 1; c[1] = -2; c[2] = 1; break;
    case IT_CUBIC: c[0] = (T)1 - dx; c[1] = (T)3 * dx - (T)2; c[2] = (T)1 - (T)3 * dx; c[3] = dx; break;
    default: break;  // no method in the reference (MethodError); callers reject it
    }
}

// InterpGridCoefs, src/interp.jl:17-31,542-571
template <typename T> struct IC {
    int N = 0, order = 0, L = 0;
    std::vector<i64> dims, strides;
    std::vector<std::array<i64, 4>> coord1d;
    std::vector<std::array<T, 4>> coef1d, gcoef1d, hcoef1d;
    std::vector<const T*> c1d;
    std::vector<i64> offset, index;
    i64 offset_base = 0;
    i64 total = 0;  // number of addressable elements behind the array (for D1)
    bool wrap = false, valid = false;
    std::vector<T> coef;

    void init(int order_, int N_, const i64* dims_, const i64* strides_) {
        N = N_; order = order_; L = npoints(order_);
        dims.assign(dims_, dims_ + N);
        strides.assign(strides_, strides_ + N);
        coord1d.resize(N); coef1d.resize(N); gcoef1d.resize(N); hcoef1d.resize(N); c1d.resize(N);
        size_t n = 1;
        for (int d = 0; d < N; d++) n *= L;
        offset.assign(n, 0); index.assign(n, 0); coef.assign(n, (T)0);
        total = 1;
        for (int d = 0; d < N; d++) total += (dims[d] - 1) * strides[d];
        for (int d = 0; d < N; d++) offsets_1d(order, coord1d[d].data());
        interp_index();
    }
    // src/interp.jl:1029-1041 (dim-1-fastest enumeration, test/counter.jl:4-10)
    void interp_index() {
        std::vector<int> c(N, 0);
        for (size_t i = 0; i < offset.size(); i++) {
            i64 ind = 0;
            for (int d = 0; d < N; d++) ind += strides[d] * coord1d[d][c[d]];
            offset[i] = ind;
            for (int d = 0; d < N; d++) {
                if (++c[d] < L) break;
                c[d] = 0;
            }
        }
    }
    // src/interp.jl:1130-1183 / 1240-1283: coef = ((w_N*w_{N-1})*...)*w_1, dim 1 fastest.
    // with_index: src/interp.jl:1044-1129 / 1189-1239 (index = 1 + sum (c_d-1)*stride_d).
    void build(const T* const* w, bool with_index) {
        std::vector<int> c(N, 0);
        std::vector<T> p(N);
        std::vector<i64> off(N);
        auto refresh = [&](int from) {
            for (int d = from; d >= 0; d--) {
                if (d == N - 1) {
                    p[d] = w[d][c[d]];
                    off[d] = (coord1d[d][c[d]] - 1) * strides[d];
                } else {
                    p[d] = p[d + 1] * w[d][c[d]];
                    off[d] = off[d + 1] + (coord1d[d][c[d]] - 1) * strides[d];
                }
            }
        };
        refresh(N - 1);
        for (size_t i = 0; i < coef.size(); i++) {
            coef[i] = p[0];
            if (with_index) index[i] = off[0] + 1;
            int d = 0;
            for (; d < N; d++) {
                if (++c[d] < L) break;
                c[d] = 0;
            }
            if (d < N) refresh(d);
        }
    }
    // src/interp.jl:390-429.  Returns false when BCnil meets an invalid location (:402-404).
    bool set_position(int bc, bool calc_grad, bool calc_hess, const T* x) {
        bool v = true;
        for (int d = 0; d < N; d++)
            if (!isvalid_x<T>(bc, order, x[d], dims[d])) { v = false; break; }
        valid = v;
        if (!v && bc == BC_NIL) return false;
        bool wr = false;
        if (v) {
            i64 ib = 0;
            for (int d = 0; d < N; d++) {
                T dx = (T)0; bool tw = false;
                i64 ix = coords_1d<T>(bc, order, x[d], dims[d], coord1d[d].data(), dx, tw);
                wr = wr | tw;
                ib += (ix - 1) * strides[d];
                coefs_1d(order, dx, coef1d[d].data());
                if (calc_grad) gcoefs_1d(order, dx, gcoef1d[d].data());
                if (calc_hess) hcoefs_1d<T>(order, dx, hcoef1d[d].data());
            }
            for (int d = 0; d < N; d++) c1d[d] = coef1d[d].data();
            if (wr) build(c1d.data(), true);
            else { offset_base = ib; build(c1d.data(), false); }
        }
        wrap = wr;
        return true;
    }
    // src/interp.jl:443-459
    void set_gradient_coordinate(int i) {
        if (!valid) return;
        for (int d = 0; d < N; d++) c1d[d] = (d == i) ? gcoef1d[d].data() : coef1d[d].data();
        build(c1d.data(), false);
    }
    // src/interp.jl:460-484
    void set_hessian_coordinate(int i1, int i2) {
        if (!valid) return;
        for (int d = 0; d < N; d++) {
            if (i1 == i2) c1d[d] = (d == i1) ? hcoef1d[d].data() : coef1d[d].data();
            else c1d[d] = (d == i1 || d == i2) ? gcoef1d[d].data() : coef1d[d].data();
        }
        build(c1d.data(), false);
    }
    // src/interp.jl:489-508 (index is the 1-based channel start; D1 guards the dereference)
    T interp(const T* A, i64 idx1 = 1) const {
        if (!valid) return std::numeric_limits<T>::quiet_NaN();
        const std::vector<i64>& off = wrap ? index : offset;
        i64 base = wrap ? idx1 - 1 : idx1 + offset_base;
        T val = (T)0;
        for (size_t i = 0; i < coef.size(); i++) {
            T c = coef[i];
            T term = (T)0;
            if (c != (T)0) {
                i64 j = off[i] + base;  // 1-based
                if (j >= 1 && j <= total) term = c * A[j - 1];
            }
            val = val + term;
        }
        return val;
    }
};

inline bool needs_shift(int bc, int order) {  // src/interp.jl:97-139
    return bc == BC_FILL || (order == IT_CUBIC && nilfam(bc));
}

// affine rows: [kind, start, step, divisor]; kind 0 FloatRange, 1 StepRange, 2 UnitRange
// (src/coord.jl:30-32).  Returns the index coordinate converted to T, including the +1 shift of
// setx, each operation in the type Julia would use.
template <typename T, typename XT> inline T lookup_coord(XT xq, const double* aff, bool shift) {
    if (aff) {
        int kind = (int)aff[0];
        if (kind == 0) {
            double u = (aff[3] * (double)xq - aff[1]) / aff[2] + 1.0;
            if (shift) u = u + 1;
            return (T)u;
        } else if (kind == 1) {
            XT t = (xq - (XT)aff[1]) / (XT)aff[2];
            double u = (double)t + 1.0;
            if (shift) u = u + 1;
            return (T)u;
        } else {
            XT t = (xq - (XT)aff[1]) + (XT)1;
            if (shift) t = t + (XT)1;
            return (T)t;
        }
    }
    XT t = xq;
    if (shift) t = t + (XT)1;
    return (T)t;
}
template <typename T> inline T scale_grad(T g, const double* aff) {  // src/coord.jl:54-56
    if (!aff) return g;
    int kind = (int)aff[0];
    if (kind == 0) return (T)(((double)g * aff[3]) / aff[2]);
    if (kind == 1) return g / (T)aff[2];
    return g;
}

// Batched evaluation == a loop of the reference's scalar calls (getindex src/interp.jl:142-156,
// _valgrad :158-171, _valgradhess :205-230; CoordInterpGrid src/coord.jl:35-69).
// x: AoS (N x nq column-major, xs_q = N, xs_d = 1) or SoA via strides.
template <typename T, typename XT>
int eval_impl(const T* coefs, int N, const i64* sdims, int bc, int order, i64 nq, const XT* x, i64 xs_q, i64 xs_d,
              const double* affine, int out_mask, T* val, T* grad, T* hess, i64* first_invalid, int nthreads) {
    if ((out_mask & OUT_HESS) && order < IT_QUADRATIC) return ERR_UNSUPPORTED;  // MethodError in the reference
    if (order == IT_CUBIC && !nilfam(bc)) return ERR_UNSUPPORTED;
    std::vector<i64> strides(N);
    strides[0] = 1;
    for (int d = 1; d < N; d++) strides[d] = strides[d - 1] * sdims[d - 1];
    bool shift = needs_shift(bc, order);
    i64 bad = nq;
    const T nan = std::numeric_limits<T>::quiet_NaN();
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads) reduction(min : bad)
    {
        IC<T> ic;
        ic.init(order, N, sdims, strides.data());
        std::vector<T> xt(N);
#pragma omp for schedule(static)
        for (i64 q = 0; q < nq; q++) {
            for (int d = 0; d < N; d++)
                xt[d] = lookup_coord<T, XT>(x[q * xs_q + d * xs_d], affine ? affine + 4 * d : nullptr, shift);
            bool ok = ic.set_position(bc, (out_mask & (OUT_GRAD | OUT_HESS)) != 0, (out_mask & OUT_HESS) != 0, xt.data());
            if (!ok) {  // BCnil: ErrorException("Invalid location"); we record and emit NaN
                if (q < bad) bad = q;
                ic.valid = false;
            }
            T v = ic.interp(coefs);
            if (val) val[q] = v;
            if (out_mask & (OUT_GRAD | OUT_HESS)) {
                for (int i = 0; i < N; i++) {
                    if (ic.valid) ic.set_gradient_coordinate(i);
                    T g = ic.interp(coefs);
                    if (grad) grad[q * N + i] = ic.valid ? scale_grad<T>(g, affine ? affine + 4 * i : nullptr) : nan;
                    if (out_mask & OUT_HESS) {
                        ic.set_hessian_coordinate(i, i);
                        hess[q * N * N + i * N + i] = ic.interp(coefs);
                        for (int j = i + 1; j < N; j++) {
                            ic.set_hessian_coordinate(i, j);
                            T h = ic.interp(coefs);
                            hess[q * N * N + j * N + i] = h;  // h[i,j], column-major
                            hess[q * N * N + i * N + j] = h;  // h[j,i]
                        }
                    }
                }
            }
        }
    }
    if (first_invalid) *first_invalid = (bad < nq) ? bad : -1;
    return (bad < nq) ? ERR_INVALID_LOCATION : ERR_OK;
}

// ---------------------------------------------------------------------------------------------
// Extended-precision evaluator ("truth" for ulp accounting; OURS, not in the reference).  Same taps
// as the reference (set_position: coord1d / dx in double), but the B-spline weights
// (src/interp.jl:785-847, exact formulas) and the sum over the l^N taps are carried in long double
// (x87 80-bit: 64-bit significand).  Also returns, per output, the magnitude scale
// S = sum_i |c_i * a_i| that makes "k ulp" meaningful for a cancelling sum.
inline void truth_weights(int order, long double dx, long double* w, long double* g) {
    switch (order) {
    case IT_NEAREST: w[0] = 1; g[0] = 0; break;
    case IT_LINEAR: w[0] = 1 - dx; w[1] = dx; g[0] = -1; g[1] = 1; break;
    case IT_QUADRATIC:
        w[0] = (dx - 0.5L) * (dx - 0.5L) / 2; w[1] = 0.75L - dx * dx; w[2] = (dx + 0.5L) * (dx + 0.5L) / 2;
        g[0] = dx - 0.5L; g[1] = -2 * dx; g[2] = dx + 0.5L;
        break;
    case IT_CUBIC: {
        const long double t = 1 - dx, two3 = 2.0L / 3.0L;
        w[0] = t * t * t / 6; w[1] = two3 - dx * dx * (2 - dx) / 2; w[2] = two3 - (dx - 1) * (dx - 1) * (1 + dx) / 2; w[3] = dx * dx * dx / 6;
        g[0] = -(t * t) / 2; g[1] = 3 * dx * dx / 2 - 2 * dx; g[2] = 0.5L + dx * (1 - 1.5L * dx); g[3] = dx * dx / 2;
        break;
    }
    }
}
// out: val[nq], grad[nq*N] (rounded to double), sval[nq], sgrad[nq*N] (scales).  No affine map.
int truth_impl(const double* coefs, int N, const i64* sdims, int bc, int order, i64 nq, const double* x, double* val, double* grad,
               double* sval, double* sgrad, int nthreads) {
    std::vector<i64> strides(N);
    strides[0] = 1;
    for (int d = 1; d < N; d++) strides[d] = strides[d - 1] * sdims[d - 1];
    const bool shift = needs_shift(bc, order);
    const int L = npoints(order);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        IC<double> ic;
        ic.init(order, N, sdims, strides.data());
        std::vector<double> xt(N);
        std::vector<long double> w(4 * N), g(4 * N);
        std::vector<int> c(N);
#pragma omp for schedule(static)
        for (i64 q = 0; q < nq; q++) {
            for (int d = 0; d < N; d++) xt[d] = lookup_coord<double, double>(x[q * N + d], nullptr, shift);
            const bool ok = ic.set_position(bc, true, false, xt.data());
            if (!ok || !ic.valid) {
                val[q] = sval[q] = std::numeric_limits<double>::quiet_NaN();
                for (int d = 0; d < N; d++) grad[q * N + d] = sgrad[q * N + d] = std::numeric_limits<double>::quiet_NaN();
                continue;
            }
            // dx per dimension as the reference computes it (one double subtraction, exact by Sterbenz
            // in the interior); recovered from the value weights' argument: recompute through coords_1d
            for (int d = 0; d < N; d++) {
                double dx = 0; bool tw = false; i64 tmp[4];
                coords_1d<double>(bc, order, xt[d], sdims[d], tmp, dx, tw);
                truth_weights(order, (long double)dx, &w[4 * d], &g[4 * d]);
            }
            const std::vector<i64>& off = ic.wrap ? ic.index : ic.offset;
            const i64 base = ic.wrap ? 0 : 1 + ic.offset_base;
            for (int pass = 0; pass <= N; pass++) {
                long double acc = 0, sc = 0;
                std::fill(c.begin(), c.end(), 0);
                for (size_t i = 0; i < ic.coef.size(); i++) {
                    long double cf = 1;
                    for (int d = 0; d < N; d++) cf *= (pass == d + 1) ? g[4 * d + c[d]] : w[4 * d + c[d]];
                    const i64 j = off[i] + base;  // 1-based
                    if (cf != 0 && j >= 1 && j <= ic.total) {
                        const long double t = cf * (long double)coefs[j - 1];
                        acc += t;
                        sc += fabsl(t);
                    }
                    for (int d = 0; d < N; d++) {
                        if (++c[d] < L) break;
                        c[d] = 0;
                    }
                }
                if (pass == 0) { val[q] = (double)acc; sval[q] = (double)sc; }
                else { grad[q * N + pass - 1] = (double)acc; sgrad[q * N + pass - 1] = (double)sc; }
            }
        }
    }
    return ERR_OK;
}

// ---------------------------------------------------------------------------------------------
// Prefilter.  Julia Base linalg/lu.jl lufact!(::Tridiagonal) ("See dgttrf.f") and
// A_ldiv_B!(::LU{T,Tridiagonal{T}}, B) ("See dgtts2.f"), called from src/interp.jl:926,946,961,
// 975,979,993,1004; WoodburyMatrices.jl Woodbury(A,U,C,V) / A_ldiv_B! from :969,987,1013.
template <typename T> struct TriLU {
    i64 n = 0;
    std::vector<T> dl, d, du, du2;
    std::vector<i64> ipiv;  // 1-based
    void factor() {
        n = (i64)d.size();
        du2.assign(n > 2 ? n - 2 : 0, (T)0);
        ipiv.resize(n);
        for (i64 i = 0; i < n; i++) ipiv[i] = i + 1;
        for (i64 i = 0; i < n - 2; i++) {
            if (std::fabs(d[i]) >= std::fabs(dl[i])) {
                if (d[i] != (T)0) {
                    T fact = dl[i] / d[i];
                    dl[i] = fact;
                    d[i + 1] = d[i + 1] - fact * du[i];
                    du2[i] = (T)0;
                }
            } else {
                T fact = d[i] / dl[i];
                d[i] = dl[i];
                dl[i] = fact;
                T tmp = du[i];
                du[i] = d[i + 1];
                d[i + 1] = tmp - fact * d[i + 1];
                du2[i] = du[i + 1];
                du[i + 1] = -fact * du[i + 1];
                ipiv[i] = i + 2;
            }
        }
        if (n > 1) {
            i64 i = n - 2;
            if (std::fabs(d[i]) >= std::fabs(dl[i])) {
                if (d[i] != (T)0) {
                    T fact = dl[i] / d[i];
                    dl[i] = fact;
                    d[i + 1] = d[i + 1] - fact * du[i];
                }
            } else {
                T fact = d[i] / dl[i];
                d[i] = dl[i];
                dl[i] = fact;
                T tmp = du[i];
                du[i] = d[i + 1];
                d[i + 1] = tmp - fact * d[i + 1];
                ipiv[i] = i + 2;
            }
        }
    }
    // in-place solve on a strided vector b[0], b[s], ...
    void solve(T* b, i64 s) const {
        for (i64 i = 0; i < n - 1; i++) {
            i64 ip = ipiv[i] - 1;            // i or i+1
            i64 other = i + 1 - ip + i;      // the row that is not ip
            T tmp = b[other * s] - dl[i] * b[ip * s];
            b[i * s] = b[ip * s];
            b[(i + 1) * s] = tmp;
        }
        b[(n - 1) * s] = b[(n - 1) * s] / d[n - 1];
        if (n > 1) b[(n - 2) * s] = (b[(n - 2) * s] - du[n - 2] * b[(n - 1) * s]) / d[n - 2];
        for (i64 i = n - 3; i >= 0; i--)
            b[i * s] = (b[i * s] - du[i] * b[(i + 1) * s] - du2[i] * b[(i + 2) * s]) / d[i];
    }
};

// inv of a 2x2 via unblocked LAPACK dgetf2 (partial pivoting, reciprocal scaling) + dgetri.
template <typename T> void inv2x2(const T a_in[4] /*col-major*/, T out[4]) {
    T a[4] = {a_in[0], a_in[1], a_in[2], a_in[3]};  // a11 a21 a12 a22
    bool swap = std::fabs(a[1]) > std::fabs(a[0]);
    if (swap) { std::swap(a[0], a[1]); std::swap(a[2], a[3]); }
    T r = (T)1 / a[0];
    T l21 = a[1] * r;
    T u11 = a[0], u12 = a[2];
    T u22 = a[3] - l21 * u12;
    // dtrti2 on U
    T i11 = (T)1 / u11;
    T i22 = (T)1 / u22;
    T i12 = (-i22) * (i11 * u12);
    // dgetri: inv(A)*L = inv(U), j = 1
    T b11 = i11 - i12 * l21;
    T b21 = (T)0 - i22 * l21;
    T b12 = i12, b22 = i22;
    if (swap) { std::swap(b11, b12); std::swap(b21, b22); }  // column interchange
    out[0] = b11; out[1] = b21; out[2] = b12; out[3] = b22;
}

template <typename T> struct Prefilter {
    TriLU<T> lu;
    bool woodbury = false;
    i64 vcol[2] = {0, 0};  // 0-based column of the single 1 in each row of V
    T Cp[4] = {0, 0, 0, 0};  // col-major
    i64 n = 0;
    std::vector<T> t1, t2;

    // src/interp.jl:919-922 / 939-942 + 955-1014
    int setup(int bc, int order, i64 n_) {
        n = n_;
        if (order == IT_QUADRATIC) {
            if (n < 3) return ERR_ARG;
            lu.du.assign(n - 1, (T)(1.0 / 8));
            lu.d.assign(n, (T)(3.0 / 4));
            lu.dl = lu.du;
            if (nilfam(bc)) {  // :955-970
                lu.d[0] = lu.d[n - 1] = (T)(9.0 / 8);
                lu.dl[n - 2] = lu.du[0] = (T)(-1.0 / 4);
                woodbury = true;
                vcol[0] = 2; vcol[1] = n - 3;
            } else if (bc == BC_REFLECT) {  // :971-976
                lu.d[0] = (T)((double)lu.d[0] + 1.0 / 8);
                lu.d[n - 1] = (T)((double)lu.d[n - 1] + 1.0 / 8);
            } else if (bc == BC_PERIODIC) {  // :977-988
                woodbury = true;
                vcol[0] = n - 1; vcol[1] = 0;
            } else {  // nearest / fill :989-994
                lu.du[0] = (T)((double)lu.du[0] + 1.0 / 8);
                lu.dl[n - 2] = (T)((double)lu.dl[n - 2] + 1.0 / 8);
            }
        } else if (order == IT_CUBIC) {
            if (!nilfam(bc)) return ERR_UNSUPPORTED;
            if (n < 4) return ERR_ARG;
            lu.du.assign(n - 1, (T)(1.0 / 6));
            lu.d.assign(n, (T)(4.0 / 6));
            lu.dl = lu.du;
            // :1001-1003
            lu.d[0] = lu.d[1] = lu.d[n - 2] = lu.d[n - 1] = (T)1;
            lu.du[0] = lu.dl[n - 2] = (T)-2;
            lu.du[1] = lu.dl[n - 3] = lu.du[n - 2] = lu.dl[0] = (T)0;
            woodbury = true;
            vcol[0] = 2; vcol[1] = n - 3;
        } else
            return ERR_OK;
        lu.factor();
        if (woodbury) {
            // Cp = inv(inv(C) .+ V*(A\U)); C = c*I with c = 1/8 (quadratic) or 1 (cubic)
            T cinv = (order == IT_QUADRATIC) ? (T)1 / (T)(1.0 / 8) : (T)1;
            std::vector<T> u1(n, (T)0), u2(n, (T)0);
            u1[0] = (T)1; u2[n - 1] = (T)1;
            lu.solve(u1.data(), 1);
            lu.solve(u2.data(), 1);
            T M[4];
            M[0] = cinv + u1[vcol[0]];
            M[1] = (T)0 + u1[vcol[1]];
            M[2] = (T)0 + u2[vcol[0]];
            M[3] = cinv + u2[vcol[1]];
            inv2x2<T>(M, Cp);
            t1.resize(n); t2.resize(n);
        }
        return ERR_OK;
    }
    // one line, in place (src/interp.jl:926 / :946)
    void solve_line(T* b, i64 s) {
        if (!woodbury) { lu.solve(b, s); return; }
        for (i64 i = 0; i < n; i++) t1[i] = b[i * s];       // copy!(W.tmpN1, B)
        lu.solve(t1.data(), 1);                                // A_ldiv_B!(W.A, tmpN1)
        T k1[2] = {t1[vcol[0]], t1[vcol[1]]};                  // tmpk1 = V*tmpN1
        T k2[2];                                               // tmpk2 = Cp*tmpk1 (gemv, column order)
        k2[0] = ((T)0 + Cp[0] * k1[0]) + Cp[2] * k1[1];
        k2[1] = ((T)0 + Cp[1] * k1[0]) + Cp[3] * k1[1];
        for (i64 i = 0; i < n; i++) t2[i] = (T)0;              // tmpN2 = U*tmpk2
        t2[0] = k2[0]; t2[n - 1] = k2[1];
        lu.solve(t2.data(), 1);                                // A_ldiv_B!(W.A, tmpN2)
        for (i64 i = 0; i < n; i++) b[i * s] = t1[i] - t2[i];
    }
};

// interp_invert!(A, BC, IT, dimlist), src/interp.jl:909-953
template <typename T> int invert_impl(T* A, int N, const i64* dims, int bc, int order, const int* dimlist, int ndl, int nthreads) {
    if (order < IT_QUADRATIC) return ERR_OK;  // :909-912
    std::vector<i64> strides(N);
    strides[0] = 1;
    i64 total = dims[0];
    for (int d = 1; d < N; d++) { strides[d] = strides[d - 1] * dims[d - 1]; total *= dims[d]; }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    for (int k = 0; k < ndl; k++) {
        int idim = dimlist[k];
        if (idim < 0 || idim >= N) return ERR_ARG;
        i64 n = dims[idim], s = strides[idim];
        Prefilter<T> proto;
        int rc = proto.setup(bc, order, n);
        if (rc) return rc;
        i64 nlines = total / n;
        i64 inner = s;  // lines enumerate: lo in [0,inner), hi in [0,nlines/inner)
#pragma omp parallel num_threads(nthreads)
        {
            Prefilter<T> pf = proto;
#pragma omp for schedule(static)
            for (i64 l = 0; l < nlines; l++) {
                i64 lo = l % inner, hi = l / inner;
                pf.solve_line(A + lo + hi * inner * n, s);
            }
        }
    }
    return ERR_OK;
}

// pad1(A, f, 1:N), src/interp.jl:1287-1310
template <typename T> void pad1_impl(const T* in, int N, const i64* dims, T f, T* out) {
    std::vector<i64> od(N), c(N, 0);
    i64 ototal = 1, itotal = 1;
    for (int d = 0; d < N; d++) { od[d] = dims[d] + 2; ototal *= od[d]; itotal *= dims[d]; }
    for (i64 i = 0; i < ototal; i++) out[i] = f;
    for (i64 i = 0; i < itotal; i++) {
        i64 o = 0, st = 1;
        for (int d = 0; d < N; d++) { o += (c[d] + 1) * st; st *= od[d]; }
        out[o] = in[i];
        for (int d = 0; d < N; d++) { if (++c[d] < dims[d]) break; c[d] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------
// restrict / prolong, BLAS path.  y[i] += a*x[i] per axpy!, unfused, in the reference's order.
inline i64 restrict_size1(i64 len) { return (len & 1) ? (len + 1) / 2 : len / 2 + 1; }  // src/restrict_prolong.jl:334

template <typename T> inline void axpy(T a, const T* x, i64 sx, T* y, i64 sy, i64 n) {
    for (i64 i = 0; i < n; i++) y[i * sy] = y[i * sy] + a * x[i * sx];
}

// src/restrict_prolong.jl:103-140
template <typename T> int restrict_impl(const T* A, int N, const i64* dims, int dim, double scale, T* R, int nthreads) {
    if (dim < 0 || dim >= N) return ERR_ARG;
    i64 L = dims[dim], n = restrict_size1(L);
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < N; d++) outer *= dims[d];
    i64 skipA = inner, skipR = inner;
    std::fill(R, R + inner * n * outer, (T)0);
    const T s1 = (T)scale, s2 = (T)(scale / 2), s75 = (T)(0.75 * scale), s25 = (T)(0.25 * scale);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for collapse(2) num_threads(nthreads) schedule(static)
    for (i64 o = 0; o < outer; o++)
        for (i64 i = 0; i < inner; i++) {
            const T* a = A + i + o * inner * L;
            T* r = R + i + o * inner * n;
            if (L & 1) {
                axpy<T>(s1, a, 2 * skipA, r, skipR, n);
                axpy<T>(s2, a + skipA, 2 * skipA, r, skipR, n - 1);
                axpy<T>(s2, a + skipA, 2 * skipA, r + skipR, skipR, n - 1);
            } else {
                axpy<T>(s75, a, 2 * skipA, r, skipR, n - 1);
                axpy<T>(s25, a + skipA, 2 * skipA, r, skipR, n - 1);
                axpy<T>(s75, a + skipA, 2 * skipA, r + skipR, skipR, n - 1);
                axpy<T>(s25, a, 2 * skipA, r + skipR, skipR, n - 1);
            }
        }
    return ERR_OK;
}

// src/restrict_prolong.jl:10-48 ; len must be 2n-1 or 2(n-1) (:322-332)
template <typename T> int prolong_impl(const T* A, int N, const i64* dims, int dim, i64 len, T* P, int nthreads) {
    if (dim < 0 || dim >= N) return ERR_ARG;
    i64 n = dims[dim];
    i64 newsz = (len & 1) ? 2 * n - 1 : 2 * (n - 1);
    if (newsz != len) return ERR_ARG;
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < N; d++) outer *= dims[d];
    i64 skipA = inner, skipP = inner;
    std::fill(P, P + inner * len * outer, (T)0);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for collapse(2) num_threads(nthreads) schedule(static)
    for (i64 o = 0; o < outer; o++)
        for (i64 i = 0; i < inner; i++) {
            const T* a = A + i + o * inner * n;
            T* p = P + i + o * inner * len;
            if (len & 1) {
                for (i64 k = 0; k < n; k++) p[k * 2 * skipP] = a[k * skipA];  // copy!
                axpy<T>((T)0.5, a, skipA, p + skipP, 2 * skipP, n - 1);
                axpy<T>((T)0.5, a + skipA, skipA, p + skipP, 2 * skipP, n - 1);
            } else {
                axpy<T>((T)0.75, a, skipA, p, 2 * skipP, n - 1);
                axpy<T>((T)0.25, a + skipA, skipA, p, 2 * skipP, n - 1);
                axpy<T>((T)0.25, a, skipA, p + skipP, 2 * skipP, n - 1);
                axpy<T>((T)0.75, a + skipA, skipA, p + skipP, 2 * skipP, n - 1);
            }
        }
    return ERR_OK;
}

// restrictb (src/restrict_prolong.jl:196-224), restrict_extrap (:287-313), prolongb (:231-267):
// the plain operator, then the edge planes along `dim` overwritten in the reference's assignment
// order.  variant 1 restrictb, 2 restrict_extrap, 3 prolongb.  Constants as the reference forms them:
// `local a::T, b::T; b = scale/sqrt(3); a = 2*b` (Float64 quotient converted to T, product in T);
// `(3/4)*x` promotes a Float32 x to Float64 and the sum is rounded when stored.
template <typename T> inline T q34(T bx, T y) {
    if (sizeof(T) == 4) return (T)((double)bx + 0.75 * (double)y);
    return (T)(bx + (T)0.75 * y);
}
template <typename T> int edge_variant_impl(int variant, const T* A, int N, const i64* dims, int dim, double scale, i64 len, T* O, int nthreads) {
    if (dim < 0 || dim >= N) return ERR_ARG;
    const i64 Lin = dims[dim];
    if (Lin < 2) return ERR_ARG;  // the reference reads A[index+Astep]
    int rc;
    i64 Lout;
    if (variant == 3) { Lout = len; rc = prolong_impl<T>(A, N, dims, dim, len, O, nthreads); }
    else { Lout = restrict_size1(Lin); rc = restrict_impl<T>(A, N, dims, dim, scale, O, nthreads); }
    if (rc) return rc;
    i64 inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= dims[d];
    for (int d = dim + 1; d < N; d++) outer *= dims[d];
    const T sc = (T)scale;
    T a, b;
    if (variant == 3) {
        if (Lout & 1) { b = (T)(1.0 / std::sqrt(3.0)); a = (T)2 * b; }
        else { b = (T)(1.0 / (2.0 * std::sqrt(2.0))); a = (T)3 * b; }
    } else {
        if (Lin & 1) { b = (T)((double)sc / std::sqrt(3.0)); a = (T)2 * b; }
        else { b = (T)((double)sc / (2.0 * std::sqrt(2.0))); a = (T)3 * b; }
    }
    const i64 st = inner;
    for (i64 o = 0; o < outer; o++)
        for (i64 i = 0; i < inner; i++) {
            const T* x = A + i + o * inner * Lin;
            T* y = O + i + o * inner * Lout;
            const i64 e = (Lin - 1) * st, er = (Lout - 1) * st;
            if (variant == 1) {
                y[0] = a * x[0] + b * x[st];
                y[er] = a * x[e] + b * x[e - st];
            } else if (variant == 2) {
                if (Lin & 1) {
                    y[0] = ((T)2 * sc) * x[0];
                    y[er] = ((T)2 * sc) * x[e];
                } else {
                    y[0] = ((T)3 * sc) * x[0] - sc * x[st];
                    y[er] = ((T)3 * sc) * x[e] - sc * x[e - st];
                }
            } else if (Lout & 1) {
                y[0] = a * x[0];
                y[st] = b * x[0] + x[st] / (T)2;
                y[er] = a * x[e];
                y[er - st] = b * x[e] + x[e - st] / (T)2;
            } else {
                y[0] = a * x[0] + x[st] / (T)4;
                y[st] = q34<T>(b * x[0], x[st]);
                y[er] = a * x[e] + x[e - st] / (T)4;
                y[er - st] = q34<T>(b * x[e], x[e - st]);
            }
        }
    return ERR_OK;
}

// InterpIrregular (src/interp.jl:318-377): 1-D, sorted knots.  it: 0 nearest, 1 linear, 4 forward, 5 backward.
// Returns ERR_INVALID_LOCATION (BoundsError under BCnil) with the first offending query; outputs NaN there.
template <typename T> inline bool jl_isless(T a, T b) {  // Julia isless: NaN last, -0.0 < +0.0
    const bool an = a != a, bn = b != b;
    if (an || bn) return !an && bn;
    if (a == (T)0 && b == (T)0) return std::signbit(a) && !std::signbit(b);
    return a < b;
}
template <typename T> int irregular_impl(const T* g, const T* coefs, i64 n, int bc, int it, double fill, const T* x, T* out, i64 nq, i64* first_invalid) {
    if (n < 1) return ERR_ARG;
    if (bc == BC_REFLECT || bc == BC_PERIODIC) return ERR_UNSUPPORTED;  // :335-337
    for (i64 k = 0; k + 1 < n; k++)
        if (jl_isless(g[k + 1], g[k])) return ERR_ARG;  // issorted, :340-342
    const T fv = (bc == BC_FILL) ? (T)fill : std::numeric_limits<T>::quiet_NaN();  // :348, :351-355
    i64 bad = -1;
    for (i64 q = 0; q < nq; q++) {
        const T xq = x[q];
        i64 i;  // 1-based, :352
        if (xq == g[0]) i = 2;
        else {  // searchsortedfirst(g, x)
            i = n + 1;
            i64 lo = 0, hi = n + 1;
            while (lo < hi - 1) {
                const i64 m = (lo + hi) >> 1;
                if (jl_isless(g[m - 1], xq)) lo = m;
                else hi = m;
            }
            i = hi;
        }
        T v;
        if (i == 1 || i == n + 1) {
            if (bc == BC_NEAREST) v = (i == 1) ? coefs[0] : coefs[n - 1];  // :361-365
            else {
                v = fv;
                if (bc == BC_NIL && bad < 0) bad = q;  // throw(BoundsError()), :356-360
            }
        } else {  // _interpu, :371-377
            const T g0 = g[i - 2], g1 = g[i - 1], c0 = coefs[i - 2], c1 = coefs[i - 1];
            if (it == 4) v = c1;
            else if (it == 5) v = c0;
            else if (it == IT_NEAREST) v = (xq - g0 < g1 - xq) ? c0 : c1;
            else {
                const T f = (xq - g0) / (g1 - g0);
                v = ((T)1 - f) * c0 + f * c1;
            }
        }
        out[q] = v;
    }
    if (first_invalid) *first_invalid = bad;
    return bad >= 0 ? ERR_INVALID_LOCATION : ERR_OK;
}

// Synthetic inputs: u(splitmix64(seed + k)) in [0,1) from the top 53 bits.
inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename T> void synth_impl(T* out, i64 n, uint64_t seed, double lo, double hi, i64 first) {
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < n; k++) {
        double u = (double)(splitmix64(seed + (uint64_t)(first + k)) >> 11) * (1.0 / 9007199254740992.0);
        out[k] = (T)(lo + (hi - lo) * u);
    }
}

}  // namespace

// =============================================================================================
extern "C" {

int oracle_wrap(int bc, int64_t pos, int64_t len, int64_t* out) { *out = wrap(bc, pos, len); return 0; }
int64_t oracle_restrict_size(int64_t len) { return restrict_size1(len); }

// Low-level KAT access (test/grid.jl:17-48): set_position on a zero-initialised IC and dump state.
// coord1d/coef1d/gcoef1d/hcoef1d: N*4 ; offset/index/coef: l^N ; flags: [wrap, valid, offset_base]
int oracle_set_position_f64(int order, int N, const int64_t* dims, const int64_t* strides, int bc, int calc_grad,
                            int calc_hess, const double* x, int64_t* coord1d, double* coef1d, double* gcoef1d,
                            double* hcoef1d, int64_t* offset, int64_t* index, double* coef, int64_t* flags) {
    IC<double> ic;
    ic.init(order, N, dims, strides);
    bool ok = ic.set_position(bc, calc_grad != 0, calc_hess != 0, x);
    for (int d = 0; d < N; d++)
        for (int k = 0; k < 4; k++) {
            coord1d[d * 4 + k] = ic.coord1d[d][k];
            coef1d[d * 4 + k] = ic.coef1d[d][k];
            gcoef1d[d * 4 + k] = ic.gcoef1d[d][k];
            hcoef1d[d * 4 + k] = ic.hcoef1d[d][k];
        }
    for (size_t i = 0; i < ic.coef.size(); i++) { offset[i] = ic.offset[i]; index[i] = ic.index[i]; coef[i] = ic.coef[i]; }
    flags[0] = ic.wrap; flags[1] = ic.valid; flags[2] = ic.offset_base;
    return ok ? ERR_OK : ERR_INVALID_LOCATION;
}

// Low-level interp on raw coefficients without the InterpGrid padding/shift (test/grid.jl:123-154,
// test/cubic_interpolation.jl:12-27): out = [val, g_1..g_N, h_11, h_21, ...] as requested.
int oracle_lowlevel_f64(const double* A, int order, int N, const int64_t* dims, int bc, const double* x, int out_mask,
                        double* val, double* grad, double* hess) {
    std::vector<i64> strides(N);
    strides[0] = 1;
    for (int d = 1; d < N; d++) strides[d] = strides[d - 1] * dims[d - 1];
    IC<double> ic;
    ic.init(order, N, dims, strides.data());
    if (!ic.set_position(bc, true, (out_mask & OUT_HESS) != 0, x)) return ERR_INVALID_LOCATION;
    *val = ic.interp(A);
    for (int i = 0; i < N; i++) {
        ic.set_gradient_coordinate(i);
        grad[i] = ic.interp(A);
        if (out_mask & OUT_HESS)
            for (int j = i; j < N; j++) {
                ic.set_hessian_coordinate(i, j);
                double h = ic.interp(A);
                hess[j * N + i] = h;
                hess[i * N + j] = h;
            }
    }
    return ERR_OK;
}

int oracle_eval_truth_f64(const double* coefs, int N, const int64_t* sdims, int bc, int order, int64_t nq, const double* x, double* val,
                          double* grad, double* sval, double* sgrad, int nthreads) {
    if (order == IT_CUBIC && !nilfam(bc)) return ERR_UNSUPPORTED;
    return truth_impl(coefs, N, sdims, bc, order, nq, x, val, grad, sval, sgrad, nthreads);
}

// dtype: 0 f32, 1 f64 ; xdtype likewise.
int oracle_eval(int dtype, int xdtype, const void* coefs, int N, const int64_t* sdims, int bc, int order, int64_t nq,
                const void* x, int64_t xs_q, int64_t xs_d, const double* affine, int out_mask, void* val, void* grad,
                void* hess, int64_t* first_invalid, int nthreads) {
    if (dtype == 1 && xdtype == 1)
        return eval_impl<double, double>((const double*)coefs, N, sdims, bc, order, nq, (const double*)x, xs_q, xs_d, affine, out_mask, (double*)val, (double*)grad, (double*)hess, first_invalid, nthreads);
    if (dtype == 1 && xdtype == 0)
        return eval_impl<double, float>((const double*)coefs, N, sdims, bc, order, nq, (const float*)x, xs_q, xs_d, affine, out_mask, (double*)val, (double*)grad, (double*)hess, first_invalid, nthreads);
    if (dtype == 0 && xdtype == 1)
        return eval_impl<float, double>((const float*)coefs, N, sdims, bc, order, nq, (const double*)x, xs_q, xs_d, affine, out_mask, (float*)val, (float*)grad, (float*)hess, first_invalid, nthreads);
    if (dtype == 0 && xdtype == 0)
        return eval_impl<float, float>((const float*)coefs, N, sdims, bc, order, nq, (const float*)x, xs_q, xs_d, affine, out_mask, (float*)val, (float*)grad, (float*)hess, first_invalid, nthreads);
    return ERR_ARG;
}

int oracle_invert(int dtype, void* A, int N, const int64_t* dims, int bc, int order, const int* dimlist, int ndl, int nthreads) {
    if (dtype == 1) return invert_impl<double>((double*)A, N, dims, bc, order, dimlist, ndl, nthreads);
    if (dtype == 0) return invert_impl<float>((float*)A, N, dims, bc, order, dimlist, ndl, nthreads);
    return ERR_ARG;
}

int oracle_pad1(int dtype, const void* in, int N, const int64_t* dims, double f, void* out) {
    if (dtype == 1) pad1_impl<double>((const double*)in, N, dims, f, (double*)out);
    else if (dtype == 0) pad1_impl<float>((const float*)in, N, dims, (float)f, (float*)out);
    else return ERR_ARG;
    return ERR_OK;
}

// Prefilter system pieces, for checking the matrices against dense solves: returns the factored
// LU vectors and Cp for (bc, order, n).
int oracle_prefilter_factors_f64(int bc, int order, int64_t n, double* dl, double* d, double* du, double* du2,
                                 int64_t* ipiv, double* Cp, int64_t* vcol, int* woodbury) {
    Prefilter<double> pf;
    int rc = pf.setup(bc, order, n);
    if (rc) return rc;
    std::copy(pf.lu.dl.begin(), pf.lu.dl.end(), dl);
    std::copy(pf.lu.d.begin(), pf.lu.d.end(), d);
    std::copy(pf.lu.du.begin(), pf.lu.du.end(), du);
    std::copy(pf.lu.du2.begin(), pf.lu.du2.end(), du2);
    std::copy(pf.lu.ipiv.begin(), pf.lu.ipiv.end(), ipiv);
    for (int i = 0; i < 4; i++) Cp[i] = pf.Cp[i];
    vcol[0] = pf.vcol[0]; vcol[1] = pf.vcol[1];
    *woodbury = pf.woodbury;
    return ERR_OK;
}

int oracle_restrict(int dtype, const void* A, int N, const int64_t* dims, int dim, double scale, void* R, int nthreads) {
    if (dtype == 1) return restrict_impl<double>((const double*)A, N, dims, dim, scale, (double*)R, nthreads);
    if (dtype == 0) return restrict_impl<float>((const float*)A, N, dims, dim, scale, (float*)R, nthreads);
    return ERR_ARG;
}
int oracle_prolong(int dtype, const void* A, int N, const int64_t* dims, int dim, int64_t len, void* P, int nthreads) {
    if (dtype == 1) return prolong_impl<double>((const double*)A, N, dims, dim, len, (double*)P, nthreads);
    if (dtype == 0) return prolong_impl<float>((const float*)A, N, dims, dim, len, (float*)P, nthreads);
    return ERR_ARG;
}

int oracle_edge_variant(int variant, int dtype, const void* A, int N, const int64_t* dims, int dim, double scale, int64_t len, void* O, int nthreads) {
    if (variant < 1 || variant > 3) return ERR_ARG;
    if (dtype == 1) return edge_variant_impl<double>(variant, (const double*)A, N, dims, dim, scale, len, (double*)O, nthreads);
    if (dtype == 0) return edge_variant_impl<float>(variant, (const float*)A, N, dims, dim, scale, len, (float*)O, nthreads);
    return ERR_ARG;
}

int oracle_irregular(int dtype, const void* g, const void* coefs, int64_t n, int bc, int it, double fill, const void* x, void* out, int64_t nq,
                     int64_t* first_invalid) {
    if (dtype == 1) return irregular_impl<double>((const double*)g, (const double*)coefs, n, bc, it, fill, (const double*)x, (double*)out, nq, first_invalid);
    if (dtype == 0) return irregular_impl<float>((const float*)g, (const float*)coefs, n, bc, it, fill, (const float*)x, (float*)out, nq, first_invalid);
    return ERR_ARG;
}

int oracle_synth(int dtype, void* out, int64_t n, uint64_t seed, double lo, double hi, int64_t first) {
    if (dtype == 1) synth_impl<double>((double*)out, n, seed, lo, hi, first);
    else if (dtype == 0) synth_impl<float>((float*)out, n, seed, lo, hi, first);
    else return ERR_ARG;
    return ERR_OK;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
