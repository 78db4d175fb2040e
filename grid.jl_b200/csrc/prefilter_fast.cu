// prefilter_fast.cu -- GB_FAST interp_invert! (src/interp.jl:909-1014): the line-parallel solver of prefilter_fast.cuh on
// sm_100a.  Every sweep reads one array and writes another (segments overlap their neighbours' rows, so the update
// cannot be in place); the host ping-pongs between the caller's array and a spare one.
//
//   sweeps along dims >= 2   fast_strided_kernel: one thread per (line, segment); the 32 lines of a warp are adjacent in
//                            memory, every access a coalesced row; E rows are in flight per thread.
//   sweep along dim 1        fast_tma_kernel: a warp owns 32 adjacent lines and one segment; lane = line.  The lines are
//                            contiguous in memory, so the tile (32 lines x 128 bytes) is TRANSPOSED on its way in and
//                            out by the TMA unit: cp.async.bulk.tensor.2d loads it with the 128-byte swizzle, which lets
//                            lane r read row r as eight conflict-free 16-byte shared loads; results leave the same way
//                            (swizzled tile -> cp.async.bulk.tensor store).  Four load stages and two store stages per
//                            warp, mbarrier completion, no CTA-wide barrier anywhere.
//   both                     fast_ends_*_kernel: the K rows at each end of every line from the dense end blocks.
// A sweep the solver does not take -- lines shorter than 8K, so many lines that the exact kernel already fills the
// machine, a row pitch TMA cannot address -- runs the exact in-place kernel (prefilter.cu) on the current buffer.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "prefilter_fast.cuh"

namespace gb {

namespace {

// ---- strided lines -------------------------------------------------------------------------------------------------
template <typename T, int E> struct StridedIO {
    const T* __restrict__ src;
    T* __restrict__ dst;
    i64 s, n;
    int periodic;
    __device__ __forceinline__ void load(i64 row0, T (&v)[E]) const {
#pragma unroll
        for (int e = 0; e < E; e++) {
            i64 r = row0 + e;
            if (periodic) r = r < 0 ? r + n : (r >= n ? r - n : r);
            v[e] = (r >= 0 && r < n) ? __ldg(src + r * s) : (T)0;
        }
    }
    __device__ __forceinline__ void store(i64 row0, const T (&v)[E]) const {
#pragma unroll
        for (int e = 0; e < E; e++)
            if (row0 + e < n) dst[(row0 + e) * s] = v[e];
    }
};
template <typename T, int E, int KB>
__global__ void __launch_bounds__(128) fast_strided_kernel(const T* __restrict__ src, T* __restrict__ dst, i64 n, i64 s, i64 inner, i64 nlines, int nseg,
                                                           i64 seg_rows, T l, T rd, int periodic) {
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nlines * nseg) return;
    const i64 line = idx % nlines, seg = idx / nlines;
    const i64 base = (line % inner) + (line / inner) * inner * n;
    const i64 r0 = seg * seg_rows, r1 = min(n, r0 + seg_rows);
    if (r0 >= r1) return;
    StridedIO<T, E> io{src + base, dst + base, s, n, periodic};
    fast_segment<T, E, KB>(io, r0, r1, l, rd);
}

// ---- the K rows at both ends of every line ----------------------------------------------------------------------------
// out[i] = sum_j W[i][j] f[j], double accumulation (FAST_K rows from FAST_COLS inputs per end).
// Strided lines (the 32 lines of a warp adjacent in memory): a thread takes 8 rows of one end of one line -- every f[j] is a
// coalesced row across the warp, W is warp-uniform.
template <typename T>
__global__ void __launch_bounds__(128) fast_ends_strided_kernel(const T* __restrict__ src, T* __restrict__ dst, i64 n, i64 s, i64 inner, i64 nlines,
                                                                const double* __restrict__ wtop, const double* __restrict__ wbot) {
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * (FAST_K / 8) * nlines) return;
    const i64 line = idx % nlines;
    const int part = (int)(idx / nlines), end = part / (FAST_K / 8), i0 = (part % (FAST_K / 8)) * 8;
    const i64 base = (line % inner) + (line / inner) * inner * n;
    const T* f = src + base + (end ? n - FAST_COLS : 0) * s;
    T* o = dst + base + ((end ? n - FAST_K : 0) + i0) * s;
    const double* W = (end ? wbot : wtop) + (size_t)i0 * FAST_COLS;
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.0;
#pragma unroll 4
    for (int j = 0; j < FAST_COLS; j++) {
        const double fj = (double)__ldg(f + (i64)j * s);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = fma(__ldg(W + k * FAST_COLS + j), fj, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) o[(i64)k * s] = (T)acc[k];
}
// Contiguous lines: a warp takes one end of one line, lane = output row (FAST_K == 32): the inputs are one contiguous,
// warp-uniform run of the line, the block is read transposed (wt[j][i]) so that the lanes' loads coalesce.
static_assert(FAST_K == 32, "fast_ends_contig_kernel maps one output row to each lane of a warp");
template <typename T>
__global__ void __launch_bounds__(128) fast_ends_contig_kernel(const T* __restrict__ src, T* __restrict__ dst, i64 n, i64 nlines, const double* __restrict__ wtop_t,
                                                               const double* __restrict__ wbot_t) {
    const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2 * nlines) return;
    const i64 line = w >> 1;
    const int end = (int)(w & 1);
    const T* f = src + line * n + (end ? n - FAST_COLS : 0);
    const double* Wt = end ? wbot_t : wtop_t;
    double acc = 0.0;
#pragma unroll 8
    for (int j = 0; j < FAST_COLS; j++) acc = fma(__ldg(Wt + j * FAST_K + lane), (double)__ldg(f + j), acc);
    dst[line * n + (end ? n - FAST_K : 0) + lane] = (T)acc;
}

// ---- contiguous lines: TMA-transposed tiles ---------------------------------------------------------------------------
constexpr int TMA_WARPS = 4, TMA_NST = 4, TMA_OST = 2, TMA_TILE = 4096;  // 32 lines x 128 bytes

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fmbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void fmbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fmbar_wait(u64* bar, u32 parity) {
    const u32 a = smem_u32(bar);
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, void* dst, u64* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(m),
                 "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(x), "r"(y) : "memory");
}

// lane = line (row of the tile); 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4) (CU_TENSOR_MAP_SWIZZLE_128B)
template <typename T, int E> struct TmaIO {
    const CUtensorMap* msrc;
    const CUtensorMap* mdst;
    unsigned char* in_st;   // TMA_NST tiles of this warp
    unsigned char* out_st;  // TMA_OST tiles
    u64* bars;              // TMA_NST barriers
    i64 n;
    int line0, periodic;
    i64 first_row;          // row of box 0 of this warp's sequence
    int nboxes;             // boxes this warp will load
    int next_issue = 0, next_use = 0, nstore = 0;
    __device__ __forceinline__ int box_row(int i) const {
        i64 r = first_row + (i64)i * E;
        if (periodic) r = r < 0 ? r + n : (r >= n ? r - n : r);
        return (int)r;
    }
    __device__ __forceinline__ void issue(int i) {  // lane 0
        u64* bar = bars + (i % TMA_NST);
        fmbar_expect_tx(bar, TMA_TILE);
        tma_load_2d(msrc, in_st + (i % TMA_NST) * TMA_TILE, bar, box_row(i), line0);
    }
    __device__ __forceinline__ void start() {
        const int lane = threadIdx.x & 31;
        if (lane == 0) {
            for (int i = 0; i < TMA_NST && i < nboxes; i++) issue(i);
        }
        next_issue = TMA_NST < nboxes ? TMA_NST : nboxes;
    }
    // boxes arrive strictly in sequence: the row argument is implied (checked by construction in fast_segment's order)
    __device__ __forceinline__ void load(i64, T (&v)[E]) {
        const int lane = threadIdx.x & 31;
        const int i = next_use++, st = i % TMA_NST;
        fmbar_wait(bars + st, (u32)((i / TMA_NST) & 1));
        const unsigned char* row = in_st + st * TMA_TILE + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const float4 q = *reinterpret_cast<const float4*>(row + ((c ^ (lane & 7)) << 4));
            reinterpret_cast<float4*>(v)[c] = q;
        }
        __syncwarp();  // every lane has read the stage: refill it
        if (lane == 0 && next_issue < nboxes) issue(next_issue);
        next_issue = next_issue < nboxes ? next_issue + 1 : next_issue;
    }
    __device__ __forceinline__ void store(i64 row0, const T (&v)[E]) {
        const int lane = threadIdx.x & 31;
        const int st = nstore % TMA_OST;
        if (nstore >= TMA_OST) {  // the store that used this stage TMA_OST boxes ago has read it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TMA_OST - 1) : "memory");
            __syncwarp();
        }
        unsigned char* row = out_st + st * TMA_TILE + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<float4*>(row + ((c ^ (lane & 7)) << 4)) = reinterpret_cast<const float4*>(v)[c];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(mdst, out_st + st * TMA_TILE, (int)row0, line0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        nstore++;
    }
    __device__ __forceinline__ void finish() {
        if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
};
template <typename T, int E, int KB>
__global__ void __launch_bounds__(32 * TMA_WARPS) fast_tma_kernel(const __grid_constant__ CUtensorMap msrc, const __grid_constant__ CUtensorMap mdst, i64 n, i64 nlines,
                                                                  int nseg, i64 seg_rows, T l, T rd, int periodic) {
    extern __shared__ __align__(1024) unsigned char fsm[];
    __shared__ __align__(8) u64 s_bar[TMA_WARPS][TMA_NST];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 wid = (i64)blockIdx.x * TMA_WARPS + warp;
    const i64 ngroups = (nlines + 31) / 32;
    if (wid >= ngroups * nseg) return;
    const i64 grp = wid % ngroups, seg = wid / ngroups;
    const i64 r0 = seg * seg_rows, r1 = min(n, r0 + seg_rows);
    if (r0 >= r1) return;
    if (lane == 0) {
        for (int i = 0; i < TMA_NST; i++) fmbar_init(&s_bar[warp][i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    TmaIO<T, E> io;
    io.msrc = &msrc;
    io.mdst = &mdst;
    unsigned char* tiles = fsm + ((1024u - (smem_u32(fsm) & 1023u)) & 1023u);  // swizzled tiles need 1024-byte alignment
    io.in_st = tiles + (size_t)warp * (TMA_NST + TMA_OST) * TMA_TILE;
    io.out_st = io.in_st + TMA_NST * TMA_TILE;
    io.bars = s_bar[warp];
    io.n = n;
    io.line0 = (int)(grp * 32);
    io.periodic = periodic;
    io.first_row = r0 - (i64)KB * E;
    io.nboxes = (int)((r1 - r0 + E - 1) / E) + 2 * KB;
    io.start();
    fast_segment<T, E, KB>(io, r0, r1, l, rd);
    io.finish();
}

// ---- host ----------------------------------------------------------------------------------------------------------
template <typename T> int get_fast_sys(int bc, int order, i64 n, FastSys<T>& out) {
    struct Key { int dev, fam, order; };
    static std::mutex mu;
    static std::vector<std::pair<Key, FastSys<T>>> cache;  // (the end blocks do not depend on n once n >= 2 FAST_M)
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    const int fam = bc_family(bc);
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : cache)
        if (e.first.dev == dev && e.first.fam == fam && e.first.order == order) { out = e.second; return GB_OK; }
    FastHost<T> h;
    h.setup(fam, order, std::max<i64>(n, 2 * FAST_M));
    FastSys<T> s;
    s.l = h.l; s.rd = h.rd; s.periodic = h.periodic;
    double* d = nullptr;
    const size_t words = (size_t)FAST_K * FAST_COLS;
    std::vector<double> all(4 * words);  // wtop | wbot | wtop transposed [j][i] | wbot transposed
    for (size_t i = 0; i < (size_t)FAST_K; i++)
        for (size_t j = 0; j < (size_t)FAST_COLS; j++) {
            all[i * FAST_COLS + j] = h.wtop[i * FAST_COLS + j];
            all[words + i * FAST_COLS + j] = h.wbot[i * FAST_COLS + j];
            all[2 * words + j * FAST_K + i] = h.wtop[i * FAST_COLS + j];
            all[3 * words + j * FAST_K + i] = h.wbot[i * FAST_COLS + j];
        }
    GB_CUDA(cudaMalloc((void**)&d, 4 * words * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, all.data(), 4 * words * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "cudaMemcpy (fast prefilter end blocks)"); }
    s.wtop = d;
    s.wbot = d + words;
    cache.emplace_back(Key{dev, fam, order}, s);
    out = s;
    return GB_OK;
}

i64 fast_max_lines() {
    static const i64 v = [] { const char* e = getenv("GB_FAST_PREFILTER_MAX_LINES"); const long long x = e ? atoll(e) : 0; return (i64)(x > 0 ? x : 32768); }();
    return v;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tensor_map_encoder() {
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeFn)fn;
    }();
    return encode;
}
// can the dim-1 sweep of lines of length n go through the TMA tiles?
bool tma_lines_ok(size_t es, i64 n, i64 nlines, bool periodic) {
    static const bool off = [] { const char* e = getenv("GB_NO_TMA"); return e && *e && *e != '0'; }();
    if (off || !tensor_map_encoder()) return false;
    const i64 E = 128 / (i64)es;
    if (((size_t)n * es) % 16 != 0 || n >= (1ll << 31) || nlines >= (1ll << 31)) return false;
    if (periodic && n % E != 0) return false;  // ring boxes must not straddle the seam
    return true;
}
int make_line_map(CUtensorMap* m, int dtype, void* base, i64 n, i64 nlines) {
    const size_t es = dtype == GB_F64 ? 8 : 4;
    if (((uintptr_t)base & 15) != 0) return GB_ERR_UNSUPPORTED;
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)nlines};
    const cuuint64_t gstr[1] = {(cuuint64_t)n * es};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / es), 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = tensor_map_encoder()(m, dtype == GB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GB_OK : GB_ERR_UNSUPPORTED;
}

struct SweepGeo {
    i64 n, str, nlines, total;
    bool fast;
};
SweepGeo sweep_geo(int dtype, int ndim, const i64* dims, int bc, int order, int idim) {
    SweepGeo g{0, 1, 0, 1, false};
    for (int d = 0; d < ndim; d++) {
        if (d == idim) g.str = g.total;
        g.total *= dims[d];
    }
    g.n = dims[idim];
    g.nlines = g.n ? g.total / g.n : 0;
    const size_t es = dtype == GB_F64 ? 8 : 4;
    g.fast = order >= GB_QUADRATIC && g.total > 0 && g.n >= FAST_MIN_N && g.nlines <= fast_max_lines() &&
             (g.str > 1 || g.nlines < 32 || tma_lines_ok(es, g.n, g.nlines, bc_family(bc) == BCF_PERIODIC));
    if (g.str == 1 && g.nlines < 32 && g.nlines > 1) g.fast = false;  // a handful of contiguous lines: neither tile path fits
    return g;
}

template <typename T> int sweep_fast(const T* src, T* dst, int dtype, const SweepGeo& g, int bc, int order, cudaStream_t s) {
    constexpr int E = 128 / (int)sizeof(T), KB = FAST_K / E;
    FastSys<T> fs;
    int rc = get_fast_sys<T>(bc, order, g.n, fs);
    if (rc) return rc;
    const i64 boxes = (g.n + E - 1) / E;
    if (g.str > 1 || g.nlines == 1) {
        // enough threads to fill the machine, segments of at least 8 boxes
        const i64 want = (i64)num_sms() * 1024;
        i64 nseg = std::max<i64>(1, std::min<i64>((want + g.nlines - 1) / g.nlines, boxes / 8));
        const i64 seg_rows = (boxes + nseg - 1) / nseg * E;
        nseg = (g.n + seg_rows - 1) / seg_rows;
        const i64 threads = g.nlines * nseg;
        fast_strided_kernel<T, E, KB><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, dst, g.n, g.str, g.str, g.nlines, (int)nseg, seg_rows, fs.l, fs.rd,
                                                                                      fs.periodic);
    } else {
        CUtensorMap ms, md;
        if (make_line_map(&ms, dtype, const_cast<T*>(src), g.n, g.nlines) != GB_OK || make_line_map(&md, dtype, dst, g.n, g.nlines) != GB_OK)
            return fail(GB_ERR_CUDA, "fast prefilter: cuTensorMapEncodeTiled failed");
        const i64 groups = (g.nlines + 31) / 32;
        const i64 want = (i64)num_sms() * 16;  // warps
        i64 nseg = std::max<i64>(1, std::min<i64>((want + groups - 1) / groups, boxes / 8));
        const i64 seg_rows = (boxes + nseg - 1) / nseg * E;
        nseg = (g.n + seg_rows - 1) / seg_rows;
        const i64 warps = groups * nseg;
        auto kern = fast_tma_kernel<T, E, KB>;
        const size_t smem = (size_t)TMA_WARPS * (TMA_NST + TMA_OST) * TMA_TILE + 1024;
        GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((warps + TMA_WARPS - 1) / TMA_WARPS), 32 * TMA_WARPS, smem, s>>>(ms, md, g.n, g.nlines, (int)nseg, seg_rows, fs.l, fs.rd, fs.periodic);
    }
    count_launch();
    GB_CUDA(cudaGetLastError());
    if (!fs.periodic) {
        const size_t words = (size_t)FAST_K * FAST_COLS;
        if (g.str == 1 && g.nlines > 1) {
            const i64 threads = 2 * g.nlines * 32;
            fast_ends_contig_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, dst, g.n, g.nlines, fs.wtop + 2 * words, fs.wtop + 3 * words);
        } else {
            const i64 threads = 2 * (FAST_K / 8) * g.nlines;
            fast_ends_strided_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, dst, g.n, g.str, g.str, g.nlines, fs.wtop, fs.wbot);
        }
        count_launch();
        GB_CUDA(cudaGetLastError());
    }
    return GB_OK;
}

}  // namespace

int fast_sweep_count(int dtype, int ndim, const i64* dims, int bc, int order, const int* dimlist, int ndl) {
    int k = 0;
    for (int i = 0; i < ndl; i++)
        if (dimlist[i] >= 0 && dimlist[i] < ndim && sweep_geo(dtype, ndim, dims, bc, order, dimlist[i]).fast) k++;
    return k;
}

int invert_fast_device(int dtype, int ndim, const i64* dims, void* data, void* tmp, int bc, int order, const int* dimlist, int ndl, int* in_tmp, cudaStream_t s) {
    void* cur = data;
    void* other = tmp;
    if (in_tmp) *in_tmp = 0;
    for (int i = 0; i < ndl; i++) {
        const int idim = dimlist[i];
        if (idim < 0 || idim >= ndim) return fail(GB_ERR_ARG, "dimlist entry out of range");
        const SweepGeo g = sweep_geo(dtype, ndim, dims, bc, order, idim);
        if (g.total == 0) return GB_OK;
        int rc;
        if (g.fast) {
            if (dtype == GB_F64) rc = sweep_fast<double>((const double*)cur, (double*)other, dtype, g, bc, order, s);
            else rc = sweep_fast<float>((const float*)cur, (float*)other, dtype, g, bc, order, s);
            if (rc) return rc;
            std::swap(cur, other);
        } else {
            rc = invert_sweep_exact(dtype, ndim, dims, cur, bc, order, idim, s);
            if (rc) return rc;
        }
    }
    if (in_tmp) *in_tmp = cur == tmp ? 1 : 0;
    return GB_OK;
}

}  // namespace gb
