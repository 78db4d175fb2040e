// eval_generic.cu -- rolled N-d evaluation kernel (the reference's "generic" path,
// src/interp.jl:1189-1283): any N up to 8, any order / BC, one thread per query, taps enumerated
// with an odometer exactly like interp_coef_generic, one pass per requested output (value, each
// gradient component, each Hessian entry) like _valgrad / _valgradhess (src/interp.jl:158-230).
// Used for N >= 5 and for 4-D Quadratic / Cubic, whose fully unrolled kernels (81 / 256 taps times up
// to 15 passes) cost more compile time and instruction cache than they are worth.  Same taps,
// weights, product association and summation order as the specialised kernels -> bit-identical
// results; GB_FAST is accepted and evaluated exactly.
#include "eval.cuh"

namespace gb {

namespace {

// index_coord / load_raw of eval.cuh take EvalParams<N>; feed them through a one-dimensional view.
template <typename T> __device__ __forceinline__ T g_index_coord(const GenericParams& p, int d, i64 q) {
    EvalParams<1> v;  // only the fields index_coord reads
    v.xdtype = p.xdtype;
    v.has_aff = p.has_aff;
    v.shift = p.shift;
#pragma unroll
    for (int k = 0; k < 4; k++) v.aff[0][k] = p.aff[d][k];
    v.xp[0] = p.xp[d];
    v.xqs = p.xqs;
    const double raw = load_raw<1>(v, 0, q);
    return index_coord<T, 1>(v, 0, raw);
}
template <typename T> __device__ __forceinline__ T g_scale_grad(const GenericParams& p, int d, T g) {
    if (!p.has_aff) return g;
    const int kind = (int)p.aff[d][0];
    if (kind == 0) return (T)(((double)g * p.aff[d][3]) / p.aff[d][2]);
    if (kind == 1) return g / (T)p.aff[d][2];
    return g;
}

template <typename T, int ORDER, int BCF>
__device__ __noinline__ bool taps_weights(T x, int len, int* c, T* w, T* g, T* h, bool& iswrap) {
    int cc[ORDER + 1];
    T dx;
    const bool valid = dim_taps<T, ORDER, BCF>(x, len, cc, dx, iswrap);
    T ww[ORDER + 1], gg[ORDER + 1], hh[ORDER + 1];
#pragma unroll
    for (int i = 0; i <= ORDER; i++) { gg[i] = (T)0; hh[i] = (T)0; }
    weights<ORDER, true, (ORDER >= GB_QUADRATIC)>(dx, ww, gg, hh);
#pragma unroll
    for (int i = 0; i <= ORDER; i++) { c[i] = cc[i]; w[i] = ww[i]; g[i] = gg[i]; h[i] = hh[i]; }
    return valid;
}

template <typename T> __device__ bool taps_weights_rt(int order, int bcf, T x, int len, int* c, T* w, T* g, T* h, bool& iswrap) {
#define GB_TW(O, B) return taps_weights<T, O, B>(x, len, c, w, g, h, iswrap)
#define GB_TWB(O)                               \
    switch (bcf) {                              \
    case BCF_NIL: GB_TW(O, BCF_NIL);            \
    case BCF_REFLECT: GB_TW(O, BCF_REFLECT);    \
    case BCF_PERIODIC: GB_TW(O, BCF_PERIODIC);  \
    default: GB_TW(O, BCF_CLAMP);               \
    }
    switch (order) {
    case GB_NEAREST: GB_TWB(GB_NEAREST)
    case GB_LINEAR: GB_TWB(GB_LINEAR)
    case GB_QUADRATIC: GB_TWB(GB_QUADRATIC)
    default: GB_TW(GB_CUBIC, BCF_NIL);
    }
#undef GB_TWB
#undef GB_TW
}

template <typename T> __global__ void __launch_bounds__(128) eval_generic_kernel(const GenericParams p) {
    const T* __restrict__ A = (const T*)p.coefs;
    const int N = p.N, L = p.order + 1;
    const bool guard = (p.order == GB_LINEAR) || (p.bcf == BCF_NIL && p.order == GB_CUBIC);
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < p.nq; q += step) {
        int c[MAXG][4];
        T w[3][MAXG][4];
        bool valid = true, anywrap = false;
        for (int d = 0; d < N; d++) {
            const T x = g_index_coord<T>(p, d, q);
            bool iswrap;
            valid = taps_weights_rt<T>(p.order, p.bcf, x, p.len[d], c[d], w[0][d], w[1][d], w[2][d], iswrap) && valid;
            anywrap = anywrap || iswrap;
        }
        if (p.order == GB_LINEAR && p.bcf != BCF_NIL && !anywrap)
            for (int d = 0; d < N; d++) c[d][1] = c[d][0] + 1;  // offset-table addressing, see eval.cuh
        const int npass = num_passes(N, p.outmode);
        for (int k = 0; k < npass; k++) {
            // which weight set each dimension uses in this pass (src/interp.jl:443-484)
            int s[MAXG];
            for (int d = 0; d < N; d++) s[d] = 0;
            int hi = -1, hj = -1;
            if (k >= 1 && k <= N) s[k - 1] = 1;
            else if (k > N) {
                int rem = k - 1 - N, i = 0;
                while (rem >= N - i) { rem -= N - i; i++; }
                hi = i; hj = i + rem;
                if (hi == hj) s[hi] = 2;
                else { s[hi] = 1; s[hj] = 1; }
            }
            // interp_coef_generic odometer (src/interp.jl:1240-1283): p[d] = p[d+1]*w_d, dim 1 fastest
            int idx[MAXG];
            T pr[MAXG];
            i64 off[MAXG];
            for (int d = 0; d < N; d++) idx[d] = 0;
            for (int d = N - 1; d >= 0; d--) {
                const T sw = w[s[d]][d][0];
                pr[d] = (d == N - 1) ? sw : pr[d + 1] * sw;
                off[d] = (d == N - 1 ? 0 : off[d + 1]) + (i64)c[d][0] * p.stride[d];
            }
            T acc = (T)0;
            while (true) {
                const T cf = pr[0];
                const i64 j = off[0];
                const bool inb = !guard || (u64)j < (u64)p.total;
                const T a = inb ? __ldg(A + j) : (T)0;
                T term = cf * a;
                if (nonfinite(a) && cf == (T)0) term = (T)0;  // zero-weight taps never contribute (:504)
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
                if (p.grad) ((T*)p.grad)[q * N + (k - 1)] = valid ? g_scale_grad<T>(p, k - 1, acc) : acc;
            } else {
                T* h = (T*)p.hess + q * (i64)(N * N);
                h[hj * N + hi] = acc;
                h[hi * N + hj] = acc;
            }
        }
        if (!valid && p.first_invalid) atomicMin(p.first_invalid, (u64)(q + p.q0));
    }
}

}  // namespace

int launch_eval_generic(const GenericParams& p, int dtype, cudaStream_t s) {
    const i64 need = (p.nq + 127) / 128;
    const int grid = (int)std::max<i64>(1, std::min<i64>(need, (i64)num_sms() * 32));
    if (dtype == GB_F64) eval_generic_kernel<double><<<grid, 128, 0, s>>>(p);
    else eval_generic_kernel<float><<<grid, 128, 0, s>>>(p);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

}  // namespace gb
