// irregular.cu -- InterpIrregular (src/interp.jl:318-377): 1-D interpolation on sorted, non-uniform
// knots; InterpForward / InterpBackward / InterpNearest / InterpLinear; BCnil / BCnan / BCna / BCfill /
// BCnearest (the reference rejects BCreflect and BCperiodic, :335-337).  One thread per query:
//   i = (x == g[1]) ? 2 : searchsortedfirst(g, x)        (:352, Julia's isless ordering)
//   i == 1 or i == n+1  ->  fill value / NaN / first, last sample / BoundsError (:351-366)
//   otherwise _interpu (:371-377).
// Knots, samples and queries share one element type T, so every operation is in T as in the reference.
#include "common.cuh"

namespace gb {
namespace {

// Julia's isless for floats: NaN sorts after everything, -0.0 before +0.0.
template <typename T> __device__ __forceinline__ bool jl_isless(T a, T b) {
    const bool an = a != a, bn = b != b;
    if (an || bn) return !an && bn;
    if (a == (T)0 && b == (T)0) return signbit(a) && !signbit(b);
    return a < b;
}

template <typename T> __global__ void __launch_bounds__(256) irregular_kernel(const T* __restrict__ g, const T* __restrict__ coefs, i64 n, int bc, int it, T fill,
                                                                              const T* __restrict__ x, T* __restrict__ out, i64 nq, u64* first_invalid) {
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (i64)gridDim.x * blockDim.x) {
        const T xq = x[q];
        i64 i;  // 1-based
        if (xq == __ldg(g)) {
            i = 2;
        } else {  // searchsortedfirst: first index whose knot is not less than x
            i64 lo = 0, hi = n + 1;
            while (lo < hi - 1) {
                const i64 m = (lo + hi) >> 1;
                if (jl_isless(__ldg(g + m - 1), xq)) lo = m;
                else hi = m;
            }
            i = hi;
        }
        T v;
        if (i == 1 || i == n + 1) {
            if (bc == GB_BC_NEAREST) v = __ldg(coefs + (i == 1 ? 0 : n - 1));
            else {
                v = fill;
                if (bc == GB_BC_NIL && first_invalid) atomicMin(first_invalid, (u64)q);
            }
        } else {
            const T g0 = __ldg(g + i - 2), g1 = __ldg(g + i - 1), c0 = __ldg(coefs + i - 2), c1 = __ldg(coefs + i - 1);
            if (it == GB_FORWARD) v = c1;
            else if (it == GB_BACKWARD) v = c0;
            else if (it == GB_NEAREST) v = (xq - g0 < g1 - xq) ? c0 : c1;
            else {
                const T f = (xq - g0) / (g1 - g0);
                v = ((T)1 - f) * c0 + f * c1;
            }
        }
        out[q] = v;
    }
}

template <typename T> __global__ void sorted_check_kernel(const T* __restrict__ g, i64 n, int* bad) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k + 1 < n; k += (i64)gridDim.x * blockDim.x)
        if (jl_isless(g[k + 1], g[k])) *bad = 1;
}

template <typename T> int irregular_impl(const T* g, const T* coefs, i64 n, int bc, int it, double fill, const T* x, T* out, i64 nq, u64* first_invalid, cudaStream_t s) {
    int* d_bad = nullptr;
    GB_CUDA(cudaMallocAsync((void**)&d_bad, sizeof(int), s));
    GB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
    const int cg = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)num_sms() * 8));
    sorted_check_kernel<T><<<cg, 256, 0, s>>>(g, n, d_bad);
    int bad = 0;
    GB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    GB_CUDA(cudaFreeAsync(d_bad, s));
    count_launch();
    if (bad) return fail(GB_ERR_ARG, "Coordinates must be supplied in increasing order");  // :340-342
    if (nq == 0) return GB_OK;
    const int grid = (int)std::max<i64>(1, std::min<i64>((nq + 255) / 256, (i64)num_sms() * 16));
    irregular_kernel<T><<<grid, 256, 0, s>>>(g, coefs, n, bc, it, (T)fill, x, out, nq, first_invalid);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

}  // namespace

int irregular_device(int dtype, const void* g, const void* coefs, i64 n, int bc, int it, double fill, const void* x, void* out, i64 nq, u64* first_invalid,
                     cudaStream_t s) {
    if (dtype == GB_F64) return irregular_impl<double>((const double*)g, (const double*)coefs, n, bc, it, fill, (const double*)x, (double*)out, nq, first_invalid, s);
    if (dtype == GB_F32) return irregular_impl<float>((const float*)g, (const float*)coefs, n, bc, it, fill, (const float*)x, (float*)out, nq, first_invalid, s);
    return fail(GB_ERR_ARG, "bad dtype");
}

}  // namespace gb
