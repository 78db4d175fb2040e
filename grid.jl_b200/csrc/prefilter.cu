// prefilter.cu -- interp_invert! (src/interp.jl:909-1014) and pad1 (:1287-1310) on sm_100a.
//
// The system matrix of a sweep is the same for every 1-D line of that dimension, so the
// tridiagonal LU factors (the dgttrf transliteration Julia Base uses, linalg/lu.jl) and, where the
// reference needs one, the 2x2 Woodbury capacitance matrix Cp are computed ONCE per sweep on the
// host in O(n) and uploaded; the device does the per-line substitutions (dgtts2 order) for all
// lines in parallel.  The substitution order is the reference's, so results are bit-identical
// to the CPU oracle.
//
// Kernel shape: one thread per line, lines enumerated with the contiguous dimension fastest, so
// for sweeps along dims >= 2 a warp touches 32 consecutive elements per step (coalesced).  The
// sweep along dim 1 walks 32 different lines per warp; each thread then consumes its own 128-byte
// line over 16 (f64) / 32 (f32) consecutive steps out of L1.
#include <vector>

#include "common.cuh"

namespace gb {

namespace {

template <typename T> struct LineSys {
    const T* dl;   // n-1 multipliers
    const T* d;    // n
    const T* du;   // n-1
    const T* du2;  // n-2
    const int* swap;  // n-1 flags: row interchange at step i (ipiv[i] == i+1)
    int woodbury;
    int vcol0, vcol1;
    T cp[4];  // column-major 2x2
};

// forward + backward substitution of one strided line held in global memory.
// RHS: b (in place).  A_ldiv_B!(::LU{T,Tridiagonal{T}}, B), Julia Base linalg/lu.jl ("See dgtts2.f").
template <typename T> __device__ __forceinline__ void lu_solve_inplace(T* b, i64 s, i64 n, const LineSys<T>& sys) {
    T cur = b[0];
    for (i64 i = 0; i < n - 1; i++) {
        const T nxt = b[(i + 1) * s];
        const T m = __ldg(sys.dl + i);
        T keep, tmp;
        if (__ldg(sys.swap + i)) { tmp = cur - m * nxt; keep = nxt; }
        else { tmp = nxt - m * cur; keep = cur; }
        b[i * s] = keep;
        cur = tmp;
    }
    T x1 = cur / __ldg(sys.d + (n - 1));
    b[(n - 1) * s] = x1;
    if (n > 1) {
        T x0 = (b[(n - 2) * s] - __ldg(sys.du + (n - 2)) * x1) / __ldg(sys.d + (n - 2));
        b[(n - 2) * s] = x0;
        T x2 = x1;
        x1 = x0;
        for (i64 i = n - 3; i >= 0; i--) {
            const T y = b[i * s];
            const T x = (y - __ldg(sys.du + i) * x1 - __ldg(sys.du2 + i) * x2) / __ldg(sys.d + i);
            b[i * s] = x;
            x2 = x1;
            x1 = x;
        }
    }
}

// Woodbury second solve: RHS = k2[0]*e_0 + k2[1]*e_{n-1}; forward results go to t (same layout as
// b), the backward pass is fused with B[i] = tmpN1[i] - tmpN2[i] (WoodburyMatrices A_ldiv_B!).
template <typename T> __device__ __forceinline__ void woodbury_correct(T* b, T* t, i64 s, i64 n, const LineSys<T>& sys) {
    const T k10 = b[(i64)sys.vcol0 * s], k11 = b[(i64)sys.vcol1 * s];  // tmpk1 = V*tmpN1
    const T k20 = ((T)0 + sys.cp[0] * k10) + sys.cp[2] * k11;           // tmpk2 = Cp*tmpk1 (gemv column order)
    const T k21 = ((T)0 + sys.cp[1] * k10) + sys.cp[3] * k11;
    T cur = k20;
    for (i64 i = 0; i < n - 1; i++) {
        const T nxt = (i + 1 == n - 1) ? k21 : (T)0;
        const T m = __ldg(sys.dl + i);
        T keep, tmp;
        if (__ldg(sys.swap + i)) { tmp = cur - m * nxt; keep = nxt; }
        else { tmp = nxt - m * cur; keep = cur; }
        t[i * s] = keep;
        cur = tmp;
    }
    T x1 = cur / __ldg(sys.d + (n - 1));
    b[(n - 1) * s] = b[(n - 1) * s] - x1;
    if (n > 1) {
        T x0 = (t[(n - 2) * s] - __ldg(sys.du + (n - 2)) * x1) / __ldg(sys.d + (n - 2));
        b[(n - 2) * s] = b[(n - 2) * s] - x0;
        T x2 = x1;
        x1 = x0;
        for (i64 i = n - 3; i >= 0; i--) {
            const T y = t[i * s];
            const T x = (y - __ldg(sys.du + i) * x1 - __ldg(sys.du2 + i) * x2) / __ldg(sys.d + i);
            b[i * s] = b[i * s] - x;
            x2 = x1;
            x1 = x;
        }
    }
}

// one thread per line; the line lives in global memory (L1/L2 resident while it is being swept).
template <typename T> __global__ void __launch_bounds__(128) line_solve_global(T* A, T* tmp, i64 n, i64 s, i64 inner, i64 nlines, const LineSys<T> sys) {
    const i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const i64 lo = l % inner, hi = l / inner;
    const i64 base = lo + hi * inner * n;
    lu_solve_inplace<T>(A + base, s, n, sys);
    if (sys.woodbury) woodbury_correct<T>(A + base, tmp + base, s, n, sys);
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

template <typename T> int invert_impl(T* A, int ndim, const i64* dims, int bc, int order, const int* dimlist, int ndl, cudaStream_t s) {
    if (order < GB_QUADRATIC) return GB_OK;  // src/interp.jl:909-912
    i64 total = 1;
    std::vector<i64> strides(ndim);
    for (int d = 0; d < ndim; d++) { strides[d] = total; total *= dims[d]; }
    if (total == 0) return GB_OK;
    T* tmp = nullptr;
    int rc = GB_OK;
    for (int k = 0; k < ndl && rc == GB_OK; k++) {
        const int idim = dimlist[k];
        if (idim < 0 || idim >= ndim) { rc = fail(GB_ERR_ARG, "dimlist entry out of range"); break; }
        const i64 n = dims[idim], str = strides[idim], nlines = total / n;
        HostSys<T> hs;
        rc = hs.setup(bc, order, n);
        if (rc) break;
        // one device blob: dl | d | du | du2 | swap
        const size_t nT = (size_t)(4 * n);
        const size_t bytes = nT * sizeof(T) + (size_t)n * sizeof(int);
        std::vector<unsigned char> blob(bytes, 0);
        T* hb = reinterpret_cast<T*>(blob.data());
        std::copy(hs.dl.begin(), hs.dl.end(), hb);
        std::copy(hs.d.begin(), hs.d.end(), hb + n);
        std::copy(hs.du.begin(), hs.du.end(), hb + 2 * n);
        std::copy(hs.du2.begin(), hs.du2.end(), hb + 3 * n);
        int* hsw = reinterpret_cast<int*>(blob.data() + nT * sizeof(T));
        std::copy(hs.swap.begin(), hs.swap.end(), hsw);
        unsigned char* dblob = nullptr;
        GB_CUDA(cudaMallocAsync((void**)&dblob, bytes, s));
        // blob is pageable: cudaMemcpyAsync stages it before returning, so the vector may die.
        GB_CUDA(cudaMemcpyAsync(dblob, blob.data(), bytes, cudaMemcpyHostToDevice, s));
        LineSys<T> sys;
        T* db = reinterpret_cast<T*>(dblob);
        sys.dl = db; sys.d = db + n; sys.du = db + 2 * n; sys.du2 = db + 3 * n;
        sys.swap = reinterpret_cast<const int*>(dblob + nT * sizeof(T));
        sys.woodbury = hs.woodbury ? 1 : 0;
        sys.vcol0 = (int)hs.vcol[0]; sys.vcol1 = (int)hs.vcol[1];
        for (int i = 0; i < 4; i++) sys.cp[i] = hs.cp[i];

        if (hs.woodbury && !tmp) GB_CUDA(cudaMallocAsync((void**)&tmp, (size_t)total * sizeof(T), s));
        const i64 grid = (nlines + 127) / 128;
        line_solve_global<T><<<(unsigned)grid, 128, 0, s>>>(A, tmp, n, str, str, nlines, sys);
        count_launch();
        GB_CUDA(cudaGetLastError());
        GB_CUDA(cudaFreeAsync(dblob, s));
    }
    if (tmp) GB_CUDA(cudaFreeAsync(tmp, s));
    return rc;
}

template <typename T, int N> struct PadParams { i64 od[N]; i64 id[N]; };
template <typename T, int N> __global__ void pad1_kernel(const T* __restrict__ in, T* __restrict__ out, i64 ototal, PadParams<T, N> p, T f) {
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x; o < ototal; o += step) {
        i64 rem = o, src = 0, mul = 1;
        bool inside = true;
#pragma unroll
        for (int d = 0; d < N; d++) {
            const i64 c = rem % p.od[d];
            rem /= p.od[d];
            inside = inside && c >= 1 && c <= p.id[d];
            src += (c - 1) * mul;
            mul *= p.id[d];
        }
        out[o] = inside ? in[src] : f;
    }
}
template <typename T, int N> int pad1_n(const i64* dims, const T* in, T f, T* out, cudaStream_t s) {
    PadParams<T, N> p;
    i64 ototal = 1;
    for (int d = 0; d < N; d++) { p.id[d] = dims[d]; p.od[d] = dims[d] + 2; ototal *= p.od[d]; }
    const i64 need = (ototal + 255) / 256;
    const int grid = (int)std::min<i64>(need, (i64)num_sms() * 32);
    pad1_kernel<T, N><<<grid, 256, 0, s>>>(in, out, ototal, p, f);
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
int pad1_device(int dtype, int ndim, const i64* dims, const void* in, double fill, void* out, cudaStream_t s) {
    if (dtype == GB_F64) return pad1_impl<double>(ndim, dims, (const double*)in, fill, (double*)out, s);
    if (dtype == GB_F32) return pad1_impl<float>(ndim, dims, (const float*)in, (float)fill, (float*)out, s);
    return fail(GB_ERR_ARG, "bad dtype");
}

}  // namespace gb
