// eval_tile.cu -- order-independent pieces of the binned ("tiled") evaluation path: the histogram
// scan that turns per-tile counts into bin offsets and a work-item list, and the tile-shape heuristic.  See eval.cuh for the kernels that
// compute the keys and evaluate the bricks.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>

#include <cuda.h>  // CUtensorMap / cuTensorMapEncodeTiled types (the entry point is resolved at run time)

#include "common.cuh"

namespace gb {

// ---- stage profiler (gb_profile_enable / gb_profile_read) ---------------------------------------------
namespace {
constexpr int PROF_NB = 7;  // boundaries around 6 stages
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
cudaEvent_t g_prof_ev[PROF_NB];
bool g_prof_init = false, g_prof_have = false;
u32* g_prof_mode = nullptr;  // pinned: nwork[2] of the last profiled tiled call (1 = coherent batch, evaluated in place by the plain kernel)
}  // namespace
void prof_mode(const u32* nwork, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_mode && cudaHostAlloc((void**)&g_prof_mode, sizeof(u32), cudaHostAllocDefault) != cudaSuccess) { g_prof_mode = nullptr; return; }
    cudaMemcpyAsync(g_prof_mode, nwork + 2, sizeof(u32), cudaMemcpyDeviceToHost, s);
}
void prof_mark(int i, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed) || i < 0 || i >= PROF_NB) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_init) {
        for (int k = 0; k < PROF_NB; k++)
            if (cudaEventCreate(&g_prof_ev[k]) != cudaSuccess) return;
        g_prof_init = true;
    }
    cudaEventRecord(g_prof_ev[i], s);
    if (i == PROF_NB - 1) g_prof_have = true;
}
int prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on.store(on ? 1 : 0);
    g_prof_have = false;
    return GB_OK;
}
int prof_read(double* ms, int n, int* nstages) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (nstages) *nstages = 0;
    if (!g_prof_have) return GB_OK;
    GB_CUDA(cudaEventSynchronize(g_prof_ev[PROF_NB - 1]));
    for (int k = 0; k + 1 < PROF_NB && k < n; k++) {
        float t = 0.f;
        GB_CUDA(cudaEventElapsedTime(&t, g_prof_ev[k], g_prof_ev[k + 1]));
        ms[k] = (double)t;
    }
    if (PROF_NB - 1 < n) ms[PROF_NB - 1] = g_prof_mode ? (double)*g_prof_mode : 0.0;  // [6]: 1 = the batch arrived coherent and the plain kernel evaluated it in place
    if (nstages) *nstages = PROF_NB - 1;
    return GB_OK;
}

// One CTA of 1024 threads.  Each thread owns a contiguous slice of at most SCAN_PER bins (held in
// registers: one pass over global memory); warp-shuffle scans of (queries per slice) and (work items per
// slice), then a 32-entry scan of the warp totals, give every thread its starting offsets.
constexpr int SCAN_PER = 12;  // bins per thread and segment (1024 * 12 >= sort_smem_bins(): one segment for the shared-memory sort)
struct TileGeo { int N, tile[MAXD], ntile[MAXD]; };
// work item of tile `bin`: {corner dim 1, first slot, one past the last, corner dims 2, 3 | no-brick flag} (eval.cuh: make_item)
__device__ __forceinline__ uint4 tile_item(const TileGeo& g, u32 bin, u32 ntot, u32 beg, u32 end) {
    if (bin >= ntot) return make_uint4(0u, beg, end, 0x80000000u);  // the bin of invalid locations: no brick
    u32 rem = bin, lo[MAXD] = {0, 0, 0, 0};
    for (int d = 0; d < g.N; d++) {
        lo[d] = (rem % (u32)g.ntile[d]) * (u32)g.tile[d];
        rem /= (u32)g.ntile[d];
    }
    u32 w = 0;
    if (g.N == 2) w = lo[1] & 0x7fffffffu;
    if (g.N >= 3) w = (lo[1] & 0x7fffu) | ((lo[2] & 0x7fffu) << 15);
    return make_uint4(lo[0], beg, end, w);
}
__global__ void __launch_bounds__(1024) tile_scan_kernel(const u32* __restrict__ hist, u32 nbins, u32* __restrict__ offs, u32* __restrict__ work,
                                                         u32* __restrict__ nwork, u32 chunk, const TileGeo geo) {
    if (nwork[2] != 0) return;  // coherent batch (coherence_probe_kernel): the plain kernel takes it
    __shared__ u32 s_wc[32], s_wi[32], s_base[2];
    const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) { s_base[0] = 0; s_base[1] = 0; }
    __syncthreads();
    // segments of 1024 * SCAN_PER bins (one for the shared-memory sort; more for tile_sort_global's bin counts)
    for (u32 seg = 0; seg < nbins; seg += 1024 * SCAN_PER) {
        const u32 nb = min(nbins - seg, 1024u * SCAN_PER);
        const u32 per = (nb + 1023) / 1024;  // <= SCAN_PER
        const u32 b0 = min(t * per, nb);
        u32 h[SCAN_PER];
        u32 cnt = 0, itm = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            h[k] = ((u32)k < per && b < nb) ? hist[seg + b] : 0u;
            cnt += h[k];
            itm += (h[k] + chunk - 1) / chunk;
        }
        u32 ic = cnt, ii = itm;  // inclusive scans within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, ic, o), b = __shfl_up_sync(0xffffffffu, ii, o);
            if ((int)lane >= o) { ic += a; ii += b; }
        }
        if (lane == 31) { s_wc[warp] = ic; s_wi[warp] = ii; }
        __syncthreads();
        if (warp == 0) {
            u32 wc = s_wc[lane], wi = s_wi[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 a = __shfl_up_sync(0xffffffffu, wc, o), b = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)lane >= o) { wc += a; wi += b; }
            }
            s_wc[lane] = wc;  // inclusive over warps
            s_wi[lane] = wi;
        }
        __syncthreads();
        u32 o = s_base[0] + (warp ? s_wc[warp - 1] : 0u) + ic - cnt;  // exclusive offsets of this thread's slice
        u32 w = s_base[1] + (warp ? s_wi[warp - 1] : 0u) + ii - itm;
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            if ((u32)k < per && b < nb) {
                offs[seg + b] = o;
                const u32 n = (h[k] + chunk - 1) / chunk;            // items of this bin ...
                const u32 sz = n ? (h[k] + n - 1) / n : 0;           // ... of equal size
                for (u32 j = 0; j < n; j++) {
                    reinterpret_cast<uint4*>(work)[w] = tile_item(geo, seg + b, nbins - 1, o + j * sz, o + min((j + 1) * sz, h[k]));
                    w++;
                }
                o += h[k];
            }
        }
        __syncthreads();
        if (t == 1023) { s_base[0] += s_wc[31]; s_base[1] += s_wi[31]; }
        __syncthreads();
    }
    if (t == 1023) { nwork[0] = s_base[1]; nwork[1] = 0; }  // item count, and the counter CTAs pull items from
}

// ---- single-pass scatter (eval.cuh: tile_scatter_capped_kernel) --------------------------------------------------------
// exclusive scan of one value per thread across a 1024-thread CTA; s_w[32] scratch; returns the exclusive prefix, total in s_w[31]
__device__ __forceinline__ u32 cta_scan_1024(u32 v, u32* s_w) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += a;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, w, o);
            if ((int)lane >= o) w += a;
        }
        s_w[lane] = w;
    }
    __syncthreads();
    return (warp ? s_w[warp - 1] : 0u) + inc - v;
}
// room of a bin whose sample count is h: the estimate plus four standard deviations of the sampling noise (the count of a
// bin in the sample is ~Poisson(h)), rounded up.  Shared by the kernel and the host bound.
__host__ __device__ __forceinline__ u32 bin_capacity(u32 h, float scale) { return (u32)(scale * ((float)h + 4.0f * sqrtf((float)h + 1.0f) + 4.0f)) + 1u; }
// cap[b], offs[b] (exclusive scan of cap), cursor[b] = 0 from the probe's sampled histogram (nwork[6] = queries sampled)
__global__ void __launch_bounds__(1024) tile_capacity_kernel(const u32* __restrict__ shist, u32 nbins, const u32* __restrict__ nwork, i64 nq, u32* __restrict__ offs,
                                                             u32* __restrict__ cap, u32* __restrict__ cursor) {
    if (nwork[2] != 0) return;
    __shared__ u32 s_w[32], s_base;
    const u32 t = threadIdx.x;
    if (t == 0) s_base = 0;
    __syncthreads();
    const u32 ns = nwork[6] ? nwork[6] : 1u;
    const float scale = (float)((double)nq / (double)ns);
    for (u32 seg = 0; seg < nbins; seg += 1024 * SCAN_PER) {
        const u32 nb = min(nbins - seg, 1024u * SCAN_PER);
        const u32 per = (nb + 1023) / 1024;
        const u32 b0 = min(t * per, nb);
        u32 c[SCAN_PER], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            c[k] = ((u32)k < per && b < nb) ? bin_capacity(shist[seg + b], scale) : 0u;
            sum += c[k];
        }
        u32 o = s_base + cta_scan_1024(sum, s_w);
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            if ((u32)k < per && b < nb) {
                offs[seg + b] = o;
                cap[seg + b] = c[k];
                cursor[seg + b] = 0;
                o += c[k];
            }
        }
        __syncthreads();
        if (t == 1023) s_base += s_w[31];
        __syncthreads();
    }
}
// work items after the scatter: bin b holds min(cursor[b], cap[b]) records from offs[b]
__global__ void __launch_bounds__(1024) tile_items_kernel(const u32* __restrict__ cursor, const u32* __restrict__ cap, const u32* __restrict__ offs, u32 nbins,
                                                          u32* __restrict__ work, u32* __restrict__ nwork, u32 chunk, const TileGeo geo) {
    if (nwork[2] != 0) return;
    __shared__ u32 s_w[32], s_base;
    const u32 t = threadIdx.x;
    if (t == 0) s_base = 0;
    __syncthreads();
    for (u32 seg = 0; seg < nbins; seg += 1024 * SCAN_PER) {
        const u32 nb = min(nbins - seg, 1024u * SCAN_PER);
        const u32 per = (nb + 1023) / 1024;
        const u32 b0 = min(t * per, nb);
        u32 h[SCAN_PER], itm = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            h[k] = ((u32)k < per && b < nb) ? min(cursor[seg + b], cap[seg + b]) : 0u;
            itm += (h[k] + chunk - 1) / chunk;
        }
        u32 w = s_base + cta_scan_1024(itm, s_w);
#pragma unroll
        for (int k = 0; k < SCAN_PER; k++) {
            const u32 b = b0 + k;
            if ((u32)k < per && b < nb) {
                const u32 o = offs[seg + b];
                const u32 n = (h[k] + chunk - 1) / chunk;
                const u32 sz = n ? (h[k] + n - 1) / n : 0;
                for (u32 j = 0; j < n; j++) {
                    reinterpret_cast<uint4*>(work)[w] = tile_item(geo, seg + b, nbins - 1, o + j * sz, o + min((j + 1) * sz, h[k]));
                    w++;
                }
            }
        }
        __syncthreads();
        if (t == 1023) s_base += s_w[31];
        __syncthreads();
    }
    if (t == 1023) { nwork[0] = s_base; nwork[1] = 0; }
}
int tile_capacity(const u32* shist, u32 nbins, const u32* nwork, i64 nq, u32* offs, u32* cap, u32* cursor, cudaStream_t s) {
    tile_capacity_kernel<<<1, 1024, 0, s>>>(shist, nbins, nwork, nq, offs, cap, cursor);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
int tile_items(const u32* cursor, const u32* cap, const u32* offs, u32 nbins, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s) {
    tile_items_kernel<<<1, 1024, 0, s>>>(cursor, cap, offs, nbins, work, nwork, chunk, geo);
    count_launch();
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
// Upper bound of sum_b bin_capacity(h_b) over any histogram with sum h_b = nsample: sum sqrt(h_b + 1) <= sqrt(nbins (nsample + nbins)) (Cauchy-Schwarz).
size_t tile_capacity_bound(i64 nq, i64 nsample, u32 nbins) {
    if (nsample < 1) nsample = 1;
    const double scale = (double)nq / (double)nsample * (1.0 + 1e-6);  // the kernel's scale is rounded to float
    const double b = (double)nbins;
    return (size_t)(scale * ((double)nsample + 4.0 * sqrt(b * ((double)nsample + b)) + 4.0 * b) + 2.0 * b + 64.0);
}
// GB_SORT_PASSES=1 selects the single-pass scatter where it applies (measured 2 % faster on C2 for 1.8x the record scratch:
// opt-in); default: the two-pass counting sort (histogram, scatter)
bool sort_single_pass() {
    static const bool v = [] { const char* e = getenv("GB_SORT_PASSES"); return e && atoi(e) == 1; }();
    return v;
}

// hist[b] = sum over CTAs of bh[b][cta]; one warp per bin.
__global__ void __launch_bounds__(256) bin_totals_kernel(const u32* __restrict__ bh, u32 nbins, u32 nctas, u32* __restrict__ hist, const u32* __restrict__ nwork) {
    if (nwork[2] != 0) return;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nbins) return;
    u32 sum = 0;
    for (u32 c = lane; c < nctas; c += 32) sum += bh[(size_t)warp * nctas + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) hist[warp] = sum;
}

// boff[b][cta] = offs[b] + (queries of bin b owned by CTAs < cta); one warp per bin.
__global__ void __launch_bounds__(256) bin_cursors_kernel(const u32* __restrict__ bh, u32 nbins, u32 nctas, const u32* __restrict__ offs, u32* __restrict__ boff,
                                                          const u32* __restrict__ nwork) {
    if (nwork[2] != 0) return;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nbins) return;
    u32 carry = offs[warp];
    for (u32 c0 = 0; c0 < nctas; c0 += 32) {
        const u32 c = c0 + lane;
        const u32 v = c < nctas ? bh[(size_t)warp * nctas + c] : 0;
        u32 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (c < nctas) boff[(size_t)warp * nctas + c] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

int tile_offsets(const u32* bh, u32 nbins, u32 nctas, u32* hist, u32* offs, u32* boff, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s) {
    const u32 wgrid = (nbins * 32 + 255) / 256;
    bin_totals_kernel<<<wgrid, 256, 0, s>>>(bh, nbins, nctas, hist, nwork);
    tile_scan_kernel<<<1, 1024, 0, s>>>(hist, nbins, offs, work, nwork, chunk, geo);
    bin_cursors_kernel<<<wgrid, 256, 0, s>>>(bh, nbins, nctas, offs, boff, nwork);
    count_launch(3);
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}

// global-histogram sort (tile_sort_global_kernel): hist holds the totals already; after the scan the write cursors
// are the bin offsets themselves (cursor may alias hist)
int tile_offsets_global(const u32* hist, u32 nbins, u32* offs, u32* cursor, u32* work, u32* nwork, u32 chunk, const TileGeo& geo, cudaStream_t s) {
    tile_scan_kernel<<<1, 1024, 0, s>>>(hist, nbins, offs, work, nwork, chunk, geo);
    count_launch();
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaMemcpyAsync(cursor, offs, (size_t)nbins * sizeof(u32), cudaMemcpyDeviceToDevice, s));
    return GB_OK;
}

// GB_COHERENT_PERCENT: share of the probed 32-query leaves that must be spatially coherent for a batch on the tiled path
// to be handed to the plain kernel instead (eval.cuh: coherence_probe_kernel); default 88 (70 when the caller passed
// GB_SORTED); values above 100 disable the probe.
u32 coherent_need_percent(bool hinted) {
    static const long env = [] { const char* e = getenv("GB_COHERENT_PERCENT"); return e ? atol(e) : -1L; }();
    if (env >= 0) return (u32)env;
    return hinted ? 70u : 88u;
}

// Tensor map of the coefficient array for the brick kernel's TMA loads: rank N, box = the brick.  Fails (-> cp.async
// staging) when the array breaks TMA's rules: strides of dims >= 1 must be multiples of 16 bytes, box extents <= 256.
int make_brick_tensor_map(CUtensorMap* m, int dtype, int N, const void* base, const int* len, const i64* stride, const int* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeFn)fn;
    }();
    if (!encode || N < 2 || N > 3) return GB_ERR_UNSUPPORTED;
    const size_t es = dtype == GB_F64 ? 8 : 4;
    if (((uintptr_t)base & 15) != 0) return GB_ERR_UNSUPPORTED;
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t bdim[3], estr[3] = {1, 1, 1};
    for (int d = 0; d < N; d++) {
        gdim[d] = (cuuint64_t)len[d];
        if (box[d] < 1 || box[d] > 256) return GB_ERR_UNSUPPORTED;
        bdim[d] = (cuuint32_t)box[d];
        if (d >= 1) {
            gstr[d - 1] = (cuuint64_t)stride[d] * es;
            if (gstr[d - 1] % 16 != 0) return GB_ERR_UNSUPPORTED;
        }
    }
    if ((box[0] * es) % 16 != 0) return GB_ERR_UNSUPPORTED;
    const CUresult r = encode(m, dtype == GB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)N, const_cast<void*>(base), gdim, gstr,
                              bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GB_OK : GB_ERR_UNSUPPORTED;
}

// Tile shape.  The leading extents are fixed (brick_tile); the extent along the last dimension is
// the largest power of two whose brick (tile + L-1 halo per dimension) fits the shared-memory
// budget per buffer (GB_TILE_KB, default 40 KB: two buffers x two 256-thread CTAs per SM) -- grown
// further only if the number of tiles would not fit the per-CTA shared-memory histogram of the
// counting sort; when no brick achieves that, the budget brick is kept and the sort runs on a global
// histogram.  Returns false when tiling cannot pay off (brick too large, or fewer than ~32 queries per
// tile to amortise staging a brick) -> plain path.
// bins the shared-memory sort handles (GB_SORT_SMEM_BINS lowers it: tests of the global-histogram sort)
u32 sort_smem_bins() {
    static const u32 v = [] {
        const char* e = getenv("GB_SORT_SMEM_BINS");
        const long x = e ? atol(e) : 0;
        return (u32)(x >= 2 && x <= 12000 ? x : 12000);
    }();
    return v;
}
bool choose_tile(int N, int L, size_t esz, const int* len, i64 nq, bool forced, int* tile, int* ntile) {
    const u32 MAX_BINS = sort_smem_bins() - 1;        // less the bin of invalid queries
    constexpr u32 MAX_BINS_GLOBAL = (1u << 24) - 1;
    size_t budget = 40 * 1024;
    if (const char* e = getenv("GB_TILE_KB")) {
        long v = atol(e);
        if (v >= 4 && v <= 100) budget = (size_t)v * 1024;
    }
    size_t lead = esz;
    unsigned long long nt = 1;
    for (int d = 0; d < N - 1; d++) {
        tile[d] = brick_tile(N, d);
        ntile[d] = (len[d] + tile[d] - 1) / tile[d];
        lead *= (size_t)brick_extent(N, L, d, (int)esz);
        nt *= (unsigned long long)ntile[d];
    }
    const int last = len[N - 1];
    int t = 1;
    while (t * 2 <= last && lead * (size_t)(t * 2 + L - 1) <= budget) t *= 2;
    if (lead * (size_t)(t + L - 1) > 100 * 1024) return false;
    const int t_fit = t;
    while (nt * (unsigned long long)((last + t - 1) / t) > MAX_BINS) {
        if (t >= last || lead * (size_t)(t * 2 + L - 1) > 100 * 1024) {
            // no brick makes the tiles few enough for the shared-memory histograms: keep the brick that fits
            // the budget and sort through the global histogram (tile_sort_global_kernel)
            t = t_fit;
            if (nt * (unsigned long long)((last + t - 1) / t) > MAX_BINS_GLOBAL) return false;
            break;
        }
        t *= 2;
    }
    tile[N - 1] = t;
    ntile[N - 1] = (last + t - 1) / t;
    nt *= (unsigned long long)ntile[N - 1];
    return forced || nq / (i64)nt >= 32;  // measured: 110 queries per tile (1M-query host chunks on C2) still beat plain 1.6x end to end
}

}  // namespace gb
