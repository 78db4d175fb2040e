// common.cuh -- shared definitions of libgridb200 (sm_100a only).
//
// Numerics contract: every translation unit of this library is compiled with -fmad=false, so a
// plain `a*b + c` is a rounded multiply followed by a rounded add exactly as the reference's
// Julia code evaluates it.  The GB_FAST paths call fma() explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/gridb200.h"

namespace gb {

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int MAXD = 4;  // dimension-specialised kernels (the reference unrolls 1..4, src/interp.jl:1047-1125)
constexpr int MAXG = 8;  // the rolled generic kernel (src/interp.jl:1189-1283) covers N up to 8

// BC families with identical index rules (src/interp.jl:594-783): nil/nan/na share one
// (they differ only in how an invalid location is reported), nearest/fill share one.
enum { BCF_NIL = 0, BCF_REFLECT = 1, BCF_PERIODIC = 2, BCF_CLAMP = 3 };
__host__ __device__ constexpr int bc_family(int bc) {
    return bc <= GB_BC_NA ? BCF_NIL : bc == GB_BC_REFLECT ? BCF_REFLECT : bc == GB_BC_PERIODIC ? BCF_PERIODIC : BCF_CLAMP;
}
inline bool needs_shift(int bc, int order) {  // setx, src/interp.jl:97-139
    return bc == GB_BC_FILL || (order == GB_CUBIC && bc <= GB_BC_NA);
}

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define GB_CUDA(expr)                                              \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return gb::cuda_fail(_e, #expr);    \
    } while (0)

// stream-ordered scratch released on every exit path (cudaFreeAsync on the stream it was allocated on)
struct AsyncBuf {
    void* p = nullptr;
    cudaStream_t s;
    explicit AsyncBuf(cudaStream_t s_) : s(s_) {}
    AsyncBuf(const AsyncBuf&) = delete;
    AsyncBuf& operator=(const AsyncBuf&) = delete;
    ~AsyncBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, s); }
};

extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int num_sms();
void prof_mark(int stage_boundary, cudaStream_t s);
void prof_mode(const unsigned int* nwork, cudaStream_t s);  // records which path (bins / plain for a coherent batch) the tiled call took
int prof_enable(int on);
int prof_read(double* ms, int n, int* nstages);  // eval_tile.cu: records boundary event i (0..6) when profiling is on

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = GB_F32; };
template <> struct dtype_of<double> { static constexpr int value = GB_F64; };

// ---- a / d, correctly rounded, without the division routine ------------------------------------------
// Divisions by a value known in advance (the prefilter's pivots d[i]): rd = RN(1/d) is supplied and the quotient is refined with exact FMA residuals (Markstein): q0 = RN(a*rd)
// is within 2 ulp of a/d, one correction makes it faithful, the second one rounds correctly -- the
// same bits as a / d (checked against the hardware division over 5e8 random numerators per pivot,
// the check_exact_div.c program of the test infrastructure) in 5 dependent operations instead of ~30.  Numerators whose exponent
// is far from 1 (possible underflow of the residuals, Inf/NaN) take the real division; a zero
// numerator returns the correctly signed zero directly.
__device__ __forceinline__ bool div_safe(double a) { return (unsigned)(((unsigned)__double2hiint(a) >> 20) & 0x7ffu) - 200u < 1600u; }
__device__ __forceinline__ bool div_safe(float a) { return (unsigned)(((unsigned)__float_as_int(a) >> 23) & 0xffu) - 40u < 170u; }
template <typename T> __device__ __forceinline__ T exact_div(T a, T d, T rd) {
    const T q0 = a * rd;
    if (a == (T)0) return q0;  // +-0 with the sign of a/d (rd has the sign of d): zero lines (padding planes) stay cheap
    if (div_safe(a)) {
        const T r0 = fma(-d, q0, a);
        const T q1 = fma(r0, rd, q0);
        const T r1 = fma(-d, q1, a);
        return fma(r1, rd, q1);
    }
    return a / d;
}

// x / 6 of the cubic weights (src/interp.jl: `/6`), bit for bit, in three operations instead of the ~12 plus a special-case
// branch of the IEEE division sequence -- six of them per 3-D cubic query on the FP64-bound brick kernel.  Correctly rounded for
// every finite x away from the subnormal range (the weights' arguments are 0 or >= 2^-159): x is a multiple of 4 or 8 ulp(q), so
// x / 6 is a multiple of ulp(q) / 3 -- never within ulp(q) / 6 of a rounding midpoint; r is exact; q + r * RN(1/6) is within
// 2^-53 ulp(q) of x / 6.  Checked against the hardware division on 10^9 numerators by tests/native/div6_check.c.  (x = -0 would
// give +0; the arguments are products of non-negative factors.)
__host__ __device__ __forceinline__ double div6_exact(double x) {
    const double z = 1.0 / 6.0;
    const double q = x * z;
    const double r = fma(-6.0, q, x);
    return fma(r, z, q);
}

// ---- brick geometry of the tiled evaluation (eval.cuh, eval_tile.cu) -------------------------------
// fixed leading tile extents (cells), so that brick strides are compile-time constants; the extent
// along the LAST dimension is chosen at run time (choose_tile): N=1 -, N=2 64, N=3 16x16, N=4 8x8x8
__host__ __device__ constexpr int brick_tile(int N, int d) { return N == 2 ? 64 : N == 3 ? 16 : 8; }
// extent of the brick along dimension e (tile + L-1 halo cells); the contiguous dimension is padded to a multiple of 16 bytes
// (esz = element size) so that a brick row is a legal TMA box row (cp.async.bulk.tensor: boxDim[0]*esz % 16 == 0; the box must
// also START on a 16-byte boundary -- an odd corner raises "illegal instruction" -- which tile-aligned corners do)
__host__ __device__ constexpr int brick_extent(int N, int L, int e, int esz) {
    const int f = brick_tile(N, e) + L - 1, q = 16 / esz;
    return e == 0 ? (f + q - 1) / q * q : f;
}
__host__ __device__ constexpr int brick_stride(int N, int L, int d, int esz) {
    int s = 1;
    for (int e = 0; e < d; e++) s *= brick_extent(N, L, e, esz);
    return s;
}
// the brick (tiled) path is compiled for the neighbourhoods heavy enough for it ever to win (launch_one's heuristic needs
// >= 256 B of taps per query; GB_TILED on anything else silently takes the plain path -- scheduling only, same results)
__host__ __device__ constexpr bool brick_compiled(int N, int order) { return (N == 2 || N == 3) && order >= GB_QUADRATIC; }

// ---- evaluation launch interface (eval_inst.cu, one translation unit per (T, N, order)) ---------
template <int N> struct EvalParams {
    const void* coefs;
    i64 total;            // elements in the stored array
    int len[N];           // stored dims
    i64 stride[N];        // element strides (column-major)
    const void* xp[N];    // per-dimension coordinate base pointers
    i64 xqs;              // element stride between consecutive queries
    int xdtype;           // GB_F32 / GB_F64
    int has_aff;
    double aff[N][4];     // kind, start, step, divisor
    int shift;            // +1 of setx
    void* val;
    void* grad;
    void* hess;
    i64 nq;
    i64 q0;               // index of this launch's first query in the caller's batch
    u64* first_invalid;   // device, atomicMin target (BCnil) or NULL
    int blocked;          // coefficients stored as 4^N-element Morton blocks (tap_offset, eval.cuh); N = 2..4
    u32 bstride[N];       // elements between consecutive blocks along each dimension
    int nchan;            // channels sharing the taps/weights of a query (multi-valued grids); 1 = scalar grid
    i64 cstride;          // elements between consecutive channels of the coefficient array
    // brick (spatially binned) evaluation, filled in by launch_one / launch_tiled
    int aos;              // != 0: outputs are records of `aos` words per sorted slot at `val`
    int tile[N];          // cells per tile
    int ntile[N];         // tiles per dimension
    const u32* work;      // 4 words per work item: tile, first slot, one past the last, -
    const u32* nwork;     // device: [0] number of work items, [1] the ticket counter CTAs pull items from, [2] 1 = the batch arrived
                          // spatially coherent: the binning kernels exit and the plain kernel evaluates it in place
    int use_tma;          // bricks are staged by cp.async.bulk.tensor through the kernel's tensor map (else element-wise cp.async)
    int by_query;         // brick path: the output record of a query goes to record q (its index in the batch) instead of its sorted slot
};

// One entry per (dtype, N, order) -- eval_inst.cu; each dispatches on (bc family, output mode, flags).
// outmode: 0 value, 1 value+gradient, 2 value+gradient+hessian.
#define GB_DECL_EVAL(T, N, O)                                                                                          \
    int launch_eval_##T##_##N##_##O(const EvalParams<N>& p, int bcf, int outmode, int flags, cudaStream_t s);        \
    int launch_outer_##T##_##N##_##O(const EvalParams<N>& p, int bcf, const i64* lens, int flags, void* out, cudaStream_t s);
#define GB_DECL_EVAL_N(T, N) GB_DECL_EVAL(T, N, 0) GB_DECL_EVAL(T, N, 1) GB_DECL_EVAL(T, N, 2) GB_DECL_EVAL(T, N, 3)
GB_DECL_EVAL_N(float, 1) GB_DECL_EVAL_N(float, 2) GB_DECL_EVAL_N(float, 3) GB_DECL_EVAL(float, 4, 0) GB_DECL_EVAL(float, 4, 1)
GB_DECL_EVAL_N(double, 1) GB_DECL_EVAL_N(double, 2) GB_DECL_EVAL_N(double, 3) GB_DECL_EVAL(double, 4, 0) GB_DECL_EVAL(double, 4, 1)
#undef GB_DECL_EVAL_N
#undef GB_DECL_EVAL

// ---- rolled generic N-d kernel (eval_generic.cu): N >= 5 and 4-D Quadratic/Cubic ---------------------
struct GenericParams {
    const void* coefs;
    i64 total;
    int N, order, bcf;
    int len[MAXG];
    i64 stride[MAXG];
    const void* xp[MAXG];
    i64 xqs;
    int xdtype, has_aff, shift;
    double aff[MAXG][4];
    void *val, *grad, *hess;
    i64 nq, q0;
    u64* first_invalid;
    int outmode;
};
int launch_eval_generic(const GenericParams& p, int dtype, cudaStream_t s);

// ---- prefilter (prefilter.cu) -------------------------------------------------------------------
int invert_device(int dtype, int ndim, const i64* dims, void* data, int bc, int order, const int* dimlist, int ndl,
                  cudaStream_t s);
// GB_FAST prefilter (prefilter_fast.cu): sweeps ping-pong between `data` and `tmp` (same size); *in_tmp tells where the
// result ended up.  Sweeps the fast solver does not take (short lines, many lines, TMA-ineligible pitch) run the exact
// in-place kernels on the current buffer.
int invert_fast_device(int dtype, int ndim, const i64* dims, void* data, void* tmp, int bc, int order, const int* dimlist, int ndl, int* in_tmp, cudaStream_t s);
int fast_sweep_count(int dtype, int ndim, const i64* dims, int bc, int order, const int* dimlist, int ndl);  // sweeps that will flip buffers
int invert_sweep_exact(int dtype, int ndim, const i64* dims, void* data, int bc, int order, int idim, cudaStream_t s);
int pad1_device(int dtype, int ndim, const i64* dims, const void* in, double fill, void* out, cudaStream_t s);
int filledges_device(int dtype, int ndim, const i64* dims, void* data, double val, cudaStream_t s);

// ---- restrict / prolong (multigrid.cu) ------------------------------------------------------------
int restrict_device(int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, void* out, cudaStream_t s);
int prolong_device(int dtype, int ndim, const i64* dims, const void* in, int dim, i64 len, void* out, cudaStream_t s);
int edge_variant_device(int variant, int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, i64 len, void* out, cudaStream_t s);
int restrict_slab_device(int dtype, int ndim, const i64* dims, const void* in, int dim, double scale, i64 L, i64 in_first, i64 out_first, i64 ocount, void* out,
                         cudaStream_t s);
int prolong_slab_device(int dtype, int ndim, const i64* dims, const void* in, int dim, i64 n, i64 len, i64 in_first, i64 out_first, i64 ocount, void* out,
                        cudaStream_t s);
int restrict3_device(int dtype, const i64* dims, const void* in, double scale, void* out, cudaStream_t s);
int prolong3_device(int dtype, const i64* dims, const void* in, const i64* odims, void* out, cudaStream_t s);

int restrict12_device(int dtype, i64 L1, i64 L2, i64 nz, const void* in, double scale, void* out, cudaStream_t s);
int prolong12_device(int dtype, i64 n1, i64 n2, i64 nz, const void* in, i64 m1, i64 m2, void* out, cudaStream_t s);
int restrict3_slab_device(int dtype, const i64* dims, const void* in, double scale, i64 L3, i64 in_first, i64 out_first, i64 ocount, void* out, cudaStream_t s);
int prolong3_slab_device(int dtype, const i64* dims, const void* in, const i64* m, i64 n3, i64 in_first, i64 out_first, i64 ocount, void* out, cudaStream_t s);
int restrict3_slab_peer_device(int dtype, const i64* dims, const void* in, double scale, i64 L3, i64 in_first, const void* lo, i64 lo_first, const void* hi, i64 hi_first,
                               i64 zmin, i64 zmax, i64 out_first, i64 ocount, void* out, cudaStream_t s);
int prolong3_slab_peer_device(int dtype, const i64* dims, const void* in, const i64* m, i64 n3, i64 in_first, const void* lo, i64 lo_first, const void* hi, i64 hi_first,
                              i64 zmin, i64 zmax, i64 out_first, i64 ocount, void* out, cudaStream_t s);
int irregular_device(int dtype, const void* knots, const void* coefs, i64 n, int bc, int it, double fill, const void* x, void* out, i64 nq, u64* first_invalid,
                     cudaStream_t s);
int interleave_device(int dtype, const void* in, void* out, i64 st, i64 nchan, int to_interleaved, cudaStream_t s);
int blocked_device(int dtype, int ndim, const i64* len, const void* in, void* out, int to_blocked, cudaStream_t s);
i64 blocked_elems(int ndim, const i64* len);
int synth_device(int dtype, void* out, i64 n, u64 seed, double lo, double hi, i64 first, cudaStream_t s);

}  // namespace gb
