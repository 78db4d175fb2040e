// abi.cu -- the extern "C" surface of libgridb200.so (see include/gridb200.h).
// Host-side logic only: argument checking in the reference's error vocabulary, the grid handle,
// host<->device staging (chunked, copy/compute overlapped over a small ring of streams) and
// dispatch to the kernels in eval_*.cu / prefilter.cu / multigrid.cu.  No CPU compute fallback.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gb {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }
int fail(int code, const std::string& msg) {
    t_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return GB_ERR_CUDA;
}
int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = 148;
    }
    return n;
}

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default pool.  Keep a bounded amount cached between
// calls -- staging buffers and sort scratch are reused call after call -- instead of everything forever: PyTorch's
// allocator (plain cudaMalloc) cannot reuse what this pool holds.  Once per device; GB_POOL_KEEP_MB overrides 2048.
static void pool_keep_memory() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    done[dev] = true;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
    const char* e = getenv("GB_POOL_KEEP_MB");
    unsigned long long keep = (unsigned long long)(e ? std::max(0.0, atof(e)) : 2048.0) * 1048576ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}

static size_t esize(int dtype) { return dtype == GB_F64 ? 8 : 4; }

// Coefficient arrays come from the stream-ordered pool (a cudaMalloc / cudaFree pair of 137 MB costs ~0.4 ms, more than
// half the kernels of InterpGrid(...) on 258^3); the pool keeps the memory of a destroyed grid for the next one.
static cudaError_t coefs_alloc(void** p, size_t bytes, cudaStream_t s) { return cudaMallocAsync(p, bytes ? bytes : 1, s); }
static void coefs_free(void* p) {
    if (!p) return;
    cudaDeviceSynchronize();  // (what cudaFree did implicitly: no stream may still be reading the grid)
    cudaFreeAsync(p, nullptr);
}
static bool ok_dtype(int d) { return d == GB_F32 || d == GB_F64; }

}  // namespace gb

using namespace gb;

struct gb_grid {
    int dtype, ndim, bc, order;
    double fill;
    i64 user_dims[MAXG], stored_dims[MAXG];
    i64 total;   // stored elements of ONE channel
    i64 nchan;   // channels (multi-valued grids), 1 for a scalar grid; the channel is the slowest dimension
    void* coefs;
    int blocked;  // resident layout: 4^N-element Morton blocks instead of column-major (wants_blocked below)
    int device;
    int shift;
    cudaEvent_t ready;  // recorded on the creation stream once the coefficients are final: every reader's stream waits for it
};
// Order a reader's stream after the grid's construction (device-mode creation returns before its kernels have run, and
// readers may use other, non-blocking streams).
static int wait_ready(const gb_grid* g, cudaStream_t s) {
    if (g->ready) GB_CUDA(cudaStreamWaitEvent(s, g->ready, 0));
    return GB_OK;
}
static void mark_ready(gb_grid* g, cudaStream_t s) {
    if (!g->ready && cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming) != cudaSuccess) { g->ready = nullptr; cudaStreamSynchronize(s); return; }
    if (cudaEventRecord(g->ready, s) != cudaSuccess) cudaStreamSynchronize(s);
}

// every entry point that touches a grid's memory runs on the grid's device, whatever the calling thread's current one is
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

static int check_grid_args(int dtype, int ndim, const int64_t* dims, int bc, int order) {
    if (!ok_dtype(dtype)) return fail(GB_ERR_ARG, "dtype must be GB_F32 or GB_F64");
    if (ndim < 1) return fail(GB_ERR_ARG, "ndim must be >= 1");
    if (ndim > MAXG) return fail(GB_ERR_UNSUPPORTED, "ndim > 8 is not supported");
    if (bc < GB_BC_NIL || bc > GB_BC_FILL) return fail(GB_ERR_ARG, "bad boundary condition");
    if (order < GB_NEAREST || order > GB_CUBIC) return fail(GB_ERR_ARG, "bad interpolation order");
    if (order == GB_CUBIC && bc > GB_BC_NA) return fail(GB_ERR_UNSUPPORTED, "InterpCubic supports only BCnil/BCnan/BCna (README.md:146)");
    if (!dims) return fail(GB_ERR_ARG, "dims is NULL");
    for (int d = 0; d < ndim; d++)
        if (dims[d] < 1 || dims[d] >= (1ll << 30) - 2) return fail(GB_ERR_ARG, "every dimension must be in [1, 2^30 - 3] (the reflect fold forms 2*len in 32 bits)");
    double tot = 1;
    for (int d = 0; d < ndim; d++) tot *= (double)(dims[d] + 2);
    if (tot > 2147483648.0) return fail(GB_ERR_ARG, "grids are limited to 2^31 stored elements (32-bit tap offsets)");
    return GB_OK;
}

static int min_len(int bc, int order) {
    if (order == GB_CUBIC) return 4;
    if (order == GB_QUADRATIC) return bc <= GB_BC_NA ? 3 : (bc == GB_BC_PERIODIC || bc == GB_BC_REFLECT ? 3 : 3);
    return 1;
}

extern "C" {

int gb_version(void) { return GB_VERSION; }
const char* gb_last_error(void) { return t_error.c_str(); }
int64_t gb_launch_count(void) { return g_launches.load(); }
int gb_profile_enable(int on) { return gb::prof_enable(on); }
int gb_profile_read(double* ms, int n, int* nstages) {
    if (!ms || n < 0) return fail(GB_ERR_ARG, "ms is NULL");
    return gb::prof_read(ms, n, nstages);
}

int gb_device_count(int* n) {
    if (!n) return fail(GB_ERR_ARG, "n is NULL");
    GB_CUDA(cudaGetDeviceCount(n));
    return GB_OK;
}
int gb_set_device(int device) {
    GB_CUDA(cudaSetDevice(device));
    return GB_OK;
}

int gb_device_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(GB_ERR_ARG, "ptr is NULL");
    GB_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return GB_OK;
}
int gb_device_free(void* ptr) {
    GB_CUDA(cudaFree(ptr));
    return GB_OK;
}
int gb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    GB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return GB_OK;
}
int gb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    GB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return GB_OK;
}
int gb_host_alloc_pinned(void** ptr, size_t bytes) {
    if (!ptr) return fail(GB_ERR_ARG, "ptr is NULL");
    GB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return GB_OK;
}
int gb_host_free_pinned(void* ptr) {
    GB_CUDA(cudaFreeHost(ptr));
    return GB_OK;
}
int gb_host_register(void* ptr, size_t bytes) {
    GB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return GB_OK;
}
int gb_host_unregister(void* ptr) {
    GB_CUDA(cudaHostUnregister(ptr));
    return GB_OK;
}
int gb_stream_synchronize(void* stream) {
    GB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GB_OK;
}

int gb_synth_uniform(int dtype, void* device_out, int64_t n, uint64_t seed, double lo, double hi, int64_t first, void* stream) {
    if (!ok_dtype(dtype)) return fail(GB_ERR_ARG, "bad dtype");
    return synth_device(dtype, device_out, n, seed, lo, hi, first, (cudaStream_t)stream);
}

// ---- prefilter ------------------------------------------------------------------------------
// interp_invert!(A, BC, IT, dimlist) in place; flags: GB_EXACT (the reference's substitution order) or GB_FAST (line-parallel
// solver where it applies, see prefilter_fast.cu)
int gb_invert_opt(int dtype, int ndim, const int64_t* dims, void* data, int mem, int bc, int order, const int* dimlist, int ndimlist, int flags,
                  void* stream) {
    if (!ok_dtype(dtype)) return fail(GB_ERR_ARG, "bad dtype");
    if (ndim < 1 || ndim > MAXG || !dims || !data) return fail(GB_ERR_ARG, "bad array arguments");
    if (bc < GB_BC_NIL || bc > GB_BC_FILL || order < GB_NEAREST || order > GB_CUBIC) return fail(GB_ERR_ARG, "bad bc/order");
    if (flags & ~GB_FAST) return fail(GB_ERR_ARG, "gb_invert_opt: flags must be GB_EXACT or GB_FAST");
    if (order < GB_QUADRATIC) return GB_OK;
    std::vector<int> all;
    if (!dimlist) {
        for (int d = 0; d < ndim; d++) all.push_back(d);
        dimlist = all.data();
        ndimlist = ndim;
    }
    for (int k = 0; k < ndimlist; k++)
        if (dimlist[k] < 0 || dimlist[k] >= ndim) return fail(GB_ERR_ARG, "dimlist entry out of range");
    i64 total = 1;
    for (int d = 0; d < ndim; d++) total *= dims[d];
    if (total == 0) return GB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    const size_t bytes = (size_t)total * esize(dtype);
    const int nflip = (flags & GB_FAST) ? fast_sweep_count(dtype, ndim, (const i64*)dims, bc, order, dimlist, ndimlist) : 0;
    AsyncBuf dev(s), spare(s);
    void* work = data;
    if (mem != GB_DEVICE) {
        GB_CUDA(dev.alloc(bytes));
        GB_CUDA(cudaMemcpyAsync(dev.p, data, bytes, cudaMemcpyHostToDevice, s));
        work = dev.p;
    }
    int rc = GB_OK;
    void* result = work;
    if (nflip > 0) {
        GB_CUDA(spare.alloc(bytes));
        int in_tmp = 0;
        rc = invert_fast_device(dtype, ndim, (const i64*)dims, work, spare.p, bc, order, dimlist, ndimlist, &in_tmp, s);
        if (rc) return rc;
        if (in_tmp) result = spare.p;
    } else {
        rc = invert_device(dtype, ndim, (const i64*)dims, work, bc, order, dimlist, ndimlist, s);
        if (rc) return rc;
    }
    if (mem == GB_DEVICE) {
        if (result != data) GB_CUDA(cudaMemcpyAsync(data, result, bytes, cudaMemcpyDeviceToDevice, s));  // odd number of ping-pong sweeps
        return GB_OK;
    }
    GB_CUDA(cudaMemcpyAsync(data, result, bytes, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
}
int gb_invert(int dtype, int ndim, const int64_t* dims, void* data, int mem, int bc, int order, const int* dimlist,
              int ndimlist, void* stream) {
    return gb_invert_opt(dtype, ndim, dims, data, mem, bc, order, dimlist, ndimlist, GB_EXACT, stream);
}

// filledges!(A, val), src/utilities.jl:22-33
int gb_filledges(int dtype, int ndim, const int64_t* dims, void* data, double val, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || ndim > MAXG || !dims || !data) return fail(GB_ERR_ARG, "bad arguments");
    i64 total = 1;
    for (int d = 0; d < ndim; d++) {
        if (dims[d] < 0) return fail(GB_ERR_ARG, "bad dims");
        total *= dims[d];
    }
    if (total == 0) return GB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return filledges_device(dtype, ndim, (const i64*)dims, data, val, s);
    const size_t bytes = (size_t)total * esize(dtype);
    AsyncBuf dev(s);
    GB_CUDA(dev.alloc(bytes));
    GB_CUDA(cudaMemcpyAsync(dev.p, data, bytes, cudaMemcpyHostToDevice, s));
    int rc = filledges_device(dtype, ndim, (const i64*)dims, dev.p, val, s);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(data, dev.p, bytes, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
}

int gb_pad1(int dtype, int ndim, const int64_t* dims, const void* in, double fill, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || ndim > MAXG || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return pad1_device(dtype, ndim, (const i64*)dims, in, fill, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= dims[d] + 2; }
    AsyncBuf di(s), dout(s);
    GB_CUDA(di.alloc(it * esize(dtype)));
    GB_CUDA(dout.alloc(ot * esize(dtype)));
    GB_CUDA(cudaMemcpyAsync(di.p, in, it * esize(dtype), cudaMemcpyHostToDevice, s));
    int rc = pad1_device(dtype, ndim, (const i64*)dims, di.p, fill, dout.p, s);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(out, dout.p, ot * esize(dtype), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
}

// ---- grid lifetime ----------------------------------------------------------------------------
static int new_grid(gb_grid** out, int dtype, int ndim, int bc, int order, double fill) {
    gb_grid* g = new gb_grid();
    g->dtype = dtype; g->ndim = ndim; g->bc = bc; g->order = order; g->fill = fill;
    g->coefs = nullptr; g->total = 0; g->nchan = 1; g->blocked = 0; g->ready = nullptr;
    g->shift = needs_shift(bc, order) ? 1 : 0;
    for (int d = 0; d < MAXG; d++) g->user_dims[d] = g->stored_dims[d] = 1;
    if (cudaGetDevice(&g->device) != cudaSuccess) g->device = 0;
    *out = g;
    return GB_OK;
}

// Blocked resident layout (tap_offset, eval.cuh): 2-D to 4-D scalar grids that L2 cannot hold and whose
// neighbourhoods are too light for the brick path ever to be chosen (l^N * sizeof(T) < 256 B, launch_one), so
// every evaluation is a random gather bound by DRAM granules.  GB_BLOCK_MIN_MB overrides the 48 MB threshold
// (tests use 0).
static bool wants_blocked(const gb_grid* g) {
    static const double min_mb = [] { const char* e = getenv("GB_BLOCK_MIN_MB"); return e ? atof(e) : 48.0; }();
    if (g->nchan != 1 || g->ndim < 2 || g->ndim > MAXD || (g->ndim == 4 && g->order >= GB_QUADRATIC)) return false;
    double tap_bytes = (double)esize(g->dtype);
    for (int d = 0; d < g->ndim; d++) tap_bytes *= g->order + 1;
    const i64 be = blocked_elems(g->ndim, g->stored_dims);
    return tap_bytes < 256 && (double)g->total * (double)esize(g->dtype) >= min_mb * 1048576.0 && be < (1ll << 32) &&
           (double)be <= 1.25 * (double)g->total + 4096;
}
// g->coefs holds the column-major array: replace it by the blocked one
static int to_blocked_layout(gb_grid* g, cudaStream_t s) {
    void* fin = nullptr;
    const size_t bytes = (size_t)blocked_elems(g->ndim, g->stored_dims) * esize(g->dtype);
    if (coefs_alloc(&fin, bytes, s) != cudaSuccess) {  // no room for the second copy the conversion needs: stay column-major
        cudaGetLastError();
        return GB_OK;
    }
    int rc = blocked_device(g->dtype, g->ndim, g->stored_dims, g->coefs, fin, 1, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc || e != cudaSuccess) { coefs_free(fin); return rc ? rc : cuda_fail(e, "cudaStreamSynchronize"); }
    coefs_free(g->coefs);
    g->coefs = fin;
    g->blocked = 1;
    return GB_OK;
}

// InterpGrid(A, BC, IT) for `nchan` arrays sharing one geometry (channel = slowest dimension of
// `samples`): copy (or pad by one sample per interpolated dimension, src/interp.jl:58-72), then
// interp_invert! along the interpolated dimensions only (the README's dimlist = 1:3 usage).
static int grid_create_impl(gb_grid** out, int dtype, int ndim, const int64_t* dims, int64_t nchan, const void* samples, int mem, int bc,
                            int order, double fill, int flags, void* stream) {
    if (!out || !samples) return fail(GB_ERR_ARG, "NULL argument");
    *out = nullptr;
    int rc = check_grid_args(dtype, ndim, dims, bc, order);
    if (rc) return rc;
    if (nchan < 1) return fail(GB_ERR_ARG, "nchan must be >= 1");
    if (nchan > 1 && ndim + 1 > MAXG) return fail(GB_ERR_UNSUPPORTED, "multi-valued grids: ndim + 1 must be <= 8");
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    const bool pad = needs_shift(bc, order);              // src/interp.jl:58-72
    const int pf_bc = bc == GB_BC_FILL ? GB_BC_NEAREST : bc;  // :68
    const double padval = bc == GB_BC_FILL ? fill : 0.0;      // :60,:67
    i64 ut = 1, st = 1, sd[MAXG + 1];
    for (int d = 0; d < ndim; d++) {
        sd[d] = dims[d] + (pad ? 2 : 0);
        ut *= dims[d];
        st *= sd[d];
        if (sd[d] < min_len(bc, order)) return fail(GB_ERR_ARG, "grid too small for this interpolation order");
    }
    if ((double)st * (double)nchan > 2147483648.0) return fail(GB_ERR_ARG, "multi-valued grids are limited to 2^31 stored elements over all channels (32-bit tap offsets)");
    gb_grid* g = nullptr;
    new_grid(&g, dtype, ndim, bc, order, fill);
    for (int d = 0; d < ndim; d++) { g->user_dims[d] = dims[d]; g->stored_dims[d] = sd[d]; }
    g->total = st;
    g->nchan = nchan;
    const size_t es = esize(dtype);
    cudaError_t e = coefs_alloc(&g->coefs, (size_t)st * (size_t)nchan * es, s);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMalloc(coefs)"); }
    void* staged = nullptr;
    const void* src = samples;
    auto cleanup = [&](int code) {
        if (staged) cudaFreeAsync(staged, s);
        cudaStreamSynchronize(s);
        coefs_free(g->coefs);
        delete g;
        return code;
    };
    std::vector<int> all;
    for (int d = 0; d < ndim; d++) all.push_back(d);
    sd[ndim] = nchan;  // lines of every channel are solved in the same sweeps
    const int pf_nd = nchan > 1 ? ndim + 1 : ndim;
    // GB_FAST prefilter: its sweeps ping-pong between two arrays; start in the spare one when their number is odd, so
    // that the result lands in g->coefs without a final copy
    void* spare = nullptr;
    const int nflip = (flags & GB_FAST) && order >= GB_QUADRATIC ? fast_sweep_count(dtype, pf_nd, sd, pf_bc, order, all.data(), ndim) : 0;
    if (nflip > 0) {
        e = cudaMallocAsync(&spare, (size_t)st * (size_t)nchan * es, s);
        if (e != cudaSuccess) return cleanup(cuda_fail(e, "cudaMallocAsync"));
    }
    void* first = (nflip & 1) ? spare : g->coefs;  // where padding / the copy writes
    auto cleanup2 = [&](int code) {
        if (spare) cudaFreeAsync(spare, s);
        return cleanup(code);
    };
    if (pad) {
        if (mem == GB_HOST) {
            e = cudaMallocAsync(&staged, (size_t)ut * (size_t)nchan * es, s);
            if (e != cudaSuccess) return cleanup2(cuda_fail(e, "cudaMallocAsync"));
            e = cudaMemcpyAsync(staged, samples, (size_t)ut * (size_t)nchan * es, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return cleanup2(cuda_fail(e, "cudaMemcpyAsync H2D"));
            src = staged;
        }
        for (i64 c = 0; c < nchan; c++) {
            rc = pad1_device(dtype, ndim, (const i64*)dims, (const char*)src + (size_t)c * ut * es, padval, (char*)first + (size_t)c * st * es, s);
            if (rc) return cleanup2(rc);
        }
    } else {
        e = cudaMemcpyAsync(first, samples, (size_t)ut * (size_t)nchan * es, mem == GB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return cleanup2(cuda_fail(e, "cudaMemcpyAsync"));
    }
    if (nflip > 0) {
        int in_tmp = 0;
        rc = invert_fast_device(dtype, pf_nd, sd, first, first == spare ? g->coefs : spare, pf_bc, order, all.data(), ndim, &in_tmp, s);
        if (rc == GB_OK && ((in_tmp != 0) != (first == spare))) rc = fail(GB_ERR_CUDA, "internal: fast prefilter ended in the wrong buffer");
        cudaFreeAsync(spare, s);
        spare = nullptr;
    } else {
        rc = invert_device(dtype, pf_nd, sd, g->coefs, pf_bc, order, all.data(), ndim, s);
    }
    if (rc) return cleanup(rc);
    if (nchan > 1) {  // resident layout of multi-valued grids: channel-interleaved (see interleave_device)
        void* fin = nullptr;
        e = coefs_alloc(&fin, (size_t)st * (size_t)nchan * es, s);
        if (e != cudaSuccess) return cleanup(cuda_fail(e, "cudaMalloc(coefs)"));
        rc = interleave_device(dtype, g->coefs, fin, st, nchan, 1, s);
        if (rc) { coefs_free(fin); return cleanup(rc); }
        cudaStreamSynchronize(s);
        coefs_free(g->coefs);
        g->coefs = fin;
    }
    if (wants_blocked(g)) {
        rc = to_blocked_layout(g, s);
        if (rc) return cleanup(rc);
    }
    if (staged) cudaFreeAsync(staged, s);
    if (mem == GB_HOST) {
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { staged = nullptr; return cleanup(cuda_fail(e, "cudaStreamSynchronize")); }
    } else {
        mark_ready(g, s);
    }
    *out = g;
    return GB_OK;
}
static int grid_from_coefs_impl(gb_grid** out, int dtype, int ndim, const int64_t* stored_dims, int64_t nchan, const void* coefs, int mem,
                                int bc, int order, double fill, int raw_coords, void* stream) {
    if (!out || !coefs) return fail(GB_ERR_ARG, "NULL argument");
    *out = nullptr;
    int rc = check_grid_args(dtype, ndim, stored_dims, bc, order);
    if (rc) return rc;
    if (nchan < 1) return fail(GB_ERR_ARG, "nchan must be >= 1");
    const bool pad = needs_shift(bc, order) && !raw_coords;
    i64 st = 1;
    for (int d = 0; d < ndim; d++) {
        if (stored_dims[d] < min_len(bc, order) || (pad && stored_dims[d] < 3)) return fail(GB_ERR_ARG, "stored grid too small");
        st *= stored_dims[d];
    }
    if ((double)st * (double)nchan > 2147483648.0) return fail(GB_ERR_ARG, "multi-valued grids are limited to 2^31 stored elements over all channels (32-bit tap offsets)");
    gb_grid* g = nullptr;
    new_grid(&g, dtype, ndim, bc, order, fill);
    if (raw_coords) g->shift = 0;  // low-level interface: coordinates index the array as stored (no setx shift)
    for (int d = 0; d < ndim; d++) { g->stored_dims[d] = stored_dims[d]; g->user_dims[d] = stored_dims[d] - (pad ? 2 : 0); }
    g->total = st;
    g->nchan = nchan;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)st * (size_t)nchan * esize(dtype);
    cudaError_t e = coefs_alloc(&g->coefs, bytes, s);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMalloc(coefs)"); }
    if (nchan > 1) {  // adopt channel-last coefficients into the channel-interleaved resident layout
        void* tmp = nullptr;
        const void* src = coefs;
        if (mem == GB_HOST) {
            e = cudaMallocAsync(&tmp, bytes, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, coefs, bytes, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) { if (tmp) cudaFreeAsync(tmp, s); coefs_free(g->coefs); delete g; return cuda_fail(e, "copy coefs"); }
            src = tmp;
        }
        rc = interleave_device(dtype, src, g->coefs, st, nchan, 1, s);
        if (tmp) cudaFreeAsync(tmp, s);
        e = (mem == GB_HOST) ? cudaStreamSynchronize(s) : cudaSuccess;
        if (rc || e != cudaSuccess) { coefs_free(g->coefs); delete g; return rc ? rc : cuda_fail(e, "copy coefs"); }
        if (mem != GB_HOST) mark_ready(g, s);
        *out = g;
        return GB_OK;
    }
    e = cudaMemcpyAsync(g->coefs, coefs, bytes, mem == GB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && mem == GB_HOST) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { coefs_free(g->coefs); delete g; return cuda_fail(e, "copy coefs"); }
    if (wants_blocked(g)) {
        rc = to_blocked_layout(g, s);
        if (rc) { coefs_free(g->coefs); delete g; return rc; }
    } else if (mem != GB_HOST) {
        mark_ready(g, s);
    }
    *out = g;
    return GB_OK;
}

int gb_grid_create(gb_grid** out, int dtype, int ndim, const int64_t* dims, const void* samples, int mem, int bc,
                   int order, double fill, void* stream) {
    return grid_create_impl(out, dtype, ndim, dims, 1, samples, mem, bc, order, fill, GB_EXACT, stream);
}
int gb_grid_create_opt(gb_grid** out, int dtype, int ndim, const int64_t* dims, const void* samples, int mem, int bc,
                       int order, double fill, int flags, void* stream) {
    if (flags & ~GB_FAST) return fail(GB_ERR_ARG, "gb_grid_create_opt: flags must be GB_EXACT or GB_FAST");
    return grid_create_impl(out, dtype, ndim, dims, 1, samples, mem, bc, order, fill, flags, stream);
}
int gb_grid_create_from_coefs(gb_grid** out, int dtype, int ndim, const int64_t* stored_dims, const void* coefs, int mem,
                              int bc, int order, double fill, void* stream) {
    return grid_from_coefs_impl(out, dtype, ndim, stored_dims, 1, coefs, mem, bc, order, fill, 0, stream);
}
int gb_grid_create_channels(gb_grid** out, int dtype, int ndim, const int64_t* dims, int64_t nchan, const void* samples, int mem,
                            int bc, int order, double fill, void* stream) {
    return grid_create_impl(out, dtype, ndim, dims, nchan, samples, mem, bc, order, fill, GB_EXACT, stream);
}
int gb_grid_create_from_coefs_channels(gb_grid** out, int dtype, int ndim, const int64_t* stored_dims, int64_t nchan, const void* coefs,
                                       int mem, int bc, int order, double fill, int raw_coords, void* stream) {
    return grid_from_coefs_impl(out, dtype, ndim, stored_dims, nchan, coefs, mem, bc, order, fill, raw_coords, stream);
}
int gb_grid_channels(const gb_grid* g, int64_t* nchan) {
    if (!g || !nchan) return fail(GB_ERR_ARG, "NULL argument");
    *nchan = g->nchan;
    return GB_OK;
}

int gb_grid_info(const gb_grid* g, int* dtype, int* ndim, int64_t* user_dims, int64_t* stored_dims, int* bc, int* order, double* fill) {
    if (!g) return fail(GB_ERR_ARG, "grid is NULL");
    if (dtype) *dtype = g->dtype;
    if (ndim) *ndim = g->ndim;
    if (bc) *bc = g->bc;
    if (order) *order = g->order;
    if (fill) *fill = g->fill;
    for (int d = 0; d < g->ndim; d++) {
        if (user_dims) user_dims[d] = g->user_dims[d];
        if (stored_dims) stored_dims[d] = g->stored_dims[d];
    }
    return GB_OK;
}
int gb_grid_coefs_device(const gb_grid* g, void** device_ptr) {
    if (!g || !device_ptr) return fail(GB_ERR_ARG, "NULL argument");
    if (g->blocked || g->nchan > 1)
        return fail(GB_ERR_UNSUPPORTED, "the resident coefficients of this grid are not in the reference layout (blocked / channel-interleaved); use gb_grid_coefs_copy");
    *device_ptr = g->coefs;
    return GB_OK;
}
// G.coefs in the reference layout written to a caller-provided DEVICE buffer (any resident layout)
int gb_grid_coefs_copy(const gb_grid* g, void* device_out, void* stream) {
    if (!g || !device_out) return fail(GB_ERR_ARG, "NULL argument");
    DeviceGuard guard(g->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)g->total * (size_t)g->nchan * esize(g->dtype);
    int rcw = wait_ready(g, s);
    if (rcw) return rcw;
    if (g->blocked) return blocked_device(g->dtype, g->ndim, g->stored_dims, g->coefs, device_out, 0, s);
    if (g->nchan > 1) return interleave_device(g->dtype, g->coefs, device_out, g->total, g->nchan, 0, s);
    GB_CUDA(cudaMemcpyAsync(device_out, g->coefs, bytes, cudaMemcpyDeviceToDevice, s));
    return GB_OK;
}
int gb_grid_coefs_download(const gb_grid* g, void* host_out) {
    if (!g || !host_out) return fail(GB_ERR_ARG, "NULL argument");
    DeviceGuard guard(g->device);
    const size_t bytes = (size_t)g->total * (size_t)g->nchan * esize(g->dtype);
    if (g->nchan > 1 || g->blocked) {  // resident layout is channel-interleaved / blocked; the caller gets the reference's array
        void* tmp = nullptr;
        GB_CUDA(cudaMalloc(&tmp, bytes));
        int rc = gb_grid_coefs_copy(g, tmp, nullptr);
        cudaError_t e = rc ? cudaSuccess : cudaMemcpy(host_out, tmp, bytes, cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        if (rc) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy D2H");
        return GB_OK;
    }
    int rcw = wait_ready(g, nullptr);
    if (rcw) return rcw;
    GB_CUDA(cudaMemcpyAsync(host_out, g->coefs, bytes, cudaMemcpyDeviceToHost, nullptr));
    GB_CUDA(cudaStreamSynchronize(nullptr));
    return GB_OK;
}
int gb_grid_destroy(gb_grid* g) {
    if (!g) return GB_OK;
    DeviceGuard guard(g->device);
    coefs_free(g->coefs);
    if (g->ready) cudaEventDestroy(g->ready);
    delete g;
    return GB_OK;
}

}  // extern "C"

// ---- evaluation dispatch ------------------------------------------------------------------------
namespace {

struct EvalCall {
    const gb_grid* g;
    int outmode, flags, xdtype;
    const double* affine;
    i64 nq;
    const void* xp[MAXG];  // device pointers
    i64 xqs;
    void *val, *grad, *hess;  // device
    u64* first_invalid;       // device or NULL
    i64 q0;
    int packed = 0;           // GB_OUT_PACKED: words per output record at val (0: separate arrays)
};

template <int N> void fill_params(const EvalCall& c, EvalParams<N>& p) {
    const gb_grid* g = c.g;
    p.coefs = g->coefs;
    p.total = g->total;
    i64 str = 1;
    for (int d = 0; d < N; d++) {
        p.len[d] = (int)g->stored_dims[d];
        p.stride[d] = str;
        str *= g->stored_dims[d];
        p.xp[d] = c.xp[d];
        for (int k = 0; k < 4; k++) p.aff[d][k] = c.affine ? c.affine[4 * d + k] : 0.0;
    }
    p.xqs = c.xqs;
    p.xdtype = c.xdtype;
    p.has_aff = c.affine ? 1 : 0;
    p.shift = g->shift;
    p.val = c.val; p.grad = c.grad; p.hess = c.hess;
    p.nq = c.nq;
    p.q0 = c.q0;
    p.first_invalid = c.first_invalid;
    p.aos = c.packed;
    p.blocked = g->blocked;
    {
        u32 bs = 1u << (2 * N);
        for (int d = 0; d < N; d++) {
            p.bstride[d] = bs;
            bs *= (u32)((g->stored_dims[d] + 3) / 4);
        }
    }
    p.nchan = (int)g->nchan;
    p.cstride = g->total;
    if (g->nchan > 1) {  // channel-interleaved resident layout: tap (i, c) at i*nchan + c
        for (int d = 0; d < N; d++) p.stride[d] *= g->nchan;
        p.total = g->total * g->nchan;
        p.cstride = 1;
    }
    p.work = nullptr; p.nwork = nullptr; p.use_tma = 0; p.by_query = 0;
    for (int d = 0; d < N; d++) { p.tile[d] = 1; p.ntile[d] = 1; }
}
template <int N> int launch_n(const EvalCall& c, cudaStream_t s) {
    const gb_grid* g = c.g;
    EvalParams<N> p;
    fill_params<N>(c, p);
    const int bcf = bc_family(g->bc);
#define GB_CALL(T, O) return launch_eval_##T##_##O(p, bcf, c.outmode, c.flags, s)
#define GB_ORDERS(TN)                                    \
    switch (g->order) {                                  \
    case GB_NEAREST: GB_CALL(TN, 0);                     \
    case GB_LINEAR: GB_CALL(TN, 1);                      \
    case GB_QUADRATIC: GB_CALL(TN, 2);                   \
    case GB_CUBIC: GB_CALL(TN, 3);                       \
    }
    if (g->dtype == GB_F64) {
        if constexpr (N == 1) { GB_ORDERS(double_1) }
        if constexpr (N == 2) { GB_ORDERS(double_2) }
        if constexpr (N == 3) { GB_ORDERS(double_3) }
        if constexpr (N == 4) { if (g->order == GB_NEAREST) GB_CALL(double_4, 0); GB_CALL(double_4, 1); }
    } else {
        if constexpr (N == 1) { GB_ORDERS(float_1) }
        if constexpr (N == 2) { GB_ORDERS(float_2) }
        if constexpr (N == 3) { GB_ORDERS(float_3) }
        if constexpr (N == 4) { if (g->order == GB_NEAREST) GB_CALL(float_4, 0); GB_CALL(float_4, 1); }
    }
#undef GB_ORDERS
#undef GB_CALL
    return GB_ERR_ARG;
}

// separable tensor-product getindex: c.xp[d] = device vector of axis d, lens[d] its length
template <int N> int launch_outer_n(const EvalCall& c, const i64* lens, void* out, cudaStream_t s) {
    const gb_grid* g = c.g;
    EvalParams<N> p;
    fill_params<N>(c, p);
    p.xqs = 1;
    p.val = out;
    const int bcf = bc_family(g->bc);
#define GB_CALL(T, O) return launch_outer_##T##_##O(p, bcf, lens, c.flags, out, s)
#define GB_ORDERS(TN)                                    \
    switch (g->order) {                                  \
    case GB_NEAREST: GB_CALL(TN, 0);                     \
    case GB_LINEAR: GB_CALL(TN, 1);                      \
    case GB_QUADRATIC: GB_CALL(TN, 2);                   \
    case GB_CUBIC: GB_CALL(TN, 3);                       \
    }
    if (g->dtype == GB_F64) {
        if constexpr (N == 1) { GB_ORDERS(double_1) }
        if constexpr (N == 2) { GB_ORDERS(double_2) }
        if constexpr (N == 3) { GB_ORDERS(double_3) }
        if constexpr (N == 4) { if (g->order == GB_NEAREST) GB_CALL(double_4, 0); GB_CALL(double_4, 1); }
    } else {
        if constexpr (N == 1) { GB_ORDERS(float_1) }
        if constexpr (N == 2) { GB_ORDERS(float_2) }
        if constexpr (N == 3) { GB_ORDERS(float_3) }
        if constexpr (N == 4) { if (g->order == GB_NEAREST) GB_CALL(float_4, 0); GB_CALL(float_4, 1); }
    }
#undef GB_ORDERS
#undef GB_CALL
    return GB_ERR_ARG;
}
static bool outer_separable(const gb_grid* g) {
    return g->nchan == 1 && g->ndim <= MAXD && !(g->ndim == 4 && g->order >= GB_QUADRATIC);
}

int launch_generic(const EvalCall& c, cudaStream_t s) {
    const gb_grid* g = c.g;
    if (g->nchan > 1) return fail(GB_ERR_UNSUPPORTED, "multi-valued grids are supported for N <= 4 (N = 4: InterpNearest / InterpLinear)");
    GenericParams p;
    p.coefs = g->coefs;
    p.total = g->total;
    p.N = g->ndim; p.order = g->order; p.bcf = bc_family(g->bc);
    i64 str = 1;
    for (int d = 0; d < g->ndim; d++) {
        p.len[d] = (int)g->stored_dims[d];
        p.stride[d] = str;
        str *= g->stored_dims[d];
        p.xp[d] = c.xp[d];
        for (int k = 0; k < 4; k++) p.aff[d][k] = c.affine ? c.affine[4 * d + k] : 0.0;
    }
    p.xqs = c.xqs; p.xdtype = c.xdtype; p.has_aff = c.affine ? 1 : 0; p.shift = g->shift;
    p.val = c.val; p.grad = c.grad; p.hess = c.hess;
    p.nq = c.nq; p.q0 = c.q0; p.first_invalid = c.first_invalid; p.outmode = c.outmode;
    return launch_eval_generic(p, g->dtype, s);
}

int launch_eval(const EvalCall& c, cudaStream_t s) {
    if (c.nq <= 0) return GB_OK;
    if (c.g->ndim > MAXD || (c.g->ndim == 4 && c.g->order >= GB_QUADRATIC)) return launch_generic(c, s);
    switch (c.g->ndim) {
    case 1: return launch_n<1>(c, s);
    case 2: return launch_n<2>(c, s);
    case 3: return launch_n<3>(c, s);
    case 4: return launch_n<4>(c, s);
    }
    return fail(GB_ERR_UNSUPPORTED, "ndim > 4");
}

int outmode_of(const gb_grid* g, int out_mask, void* val, void* grad, void* hess, int* outmode) {
    if (out_mask & ~(GB_OUT_VAL | GB_OUT_GRAD | GB_OUT_HESS | GB_OUT_PACKED) || (out_mask & ~GB_OUT_PACKED) == 0) return fail(GB_ERR_ARG, "bad out_mask");
    *outmode = (out_mask & GB_OUT_HESS) ? 2 : (out_mask & GB_OUT_GRAD) ? 1 : 0;
    if (out_mask & GB_OUT_PACKED) {  // one record per query at val
        if ((out_mask & GB_OUT_HESS) && g->order < GB_QUADRATIC)
            return fail(GB_ERR_UNSUPPORTED, "valgradhess is not defined for InterpNearest/InterpLinear (no interp_hcoefs_1d method)");
        if (!val) return fail(GB_ERR_ARG, "val is NULL");
        if (g->ndim > MAXD || (g->ndim == 4 && g->order >= GB_QUADRATIC)) return fail(GB_ERR_UNSUPPORTED, "GB_OUT_PACKED is provided for the dimension-specialised kernels (N <= 4; N = 4: InterpNearest / InterpLinear)");
        return GB_OK;
    }
    if ((out_mask & GB_OUT_HESS) && g->order < GB_QUADRATIC)
        return fail(GB_ERR_UNSUPPORTED, "valgradhess is not defined for InterpNearest/InterpLinear (no interp_hcoefs_1d method)");
    if ((out_mask & GB_OUT_VAL) && !val) return fail(GB_ERR_ARG, "val is NULL");
    if ((out_mask & GB_OUT_GRAD) && !grad) return fail(GB_ERR_ARG, "Wrong number of components for the gradient");
    if ((out_mask & GB_OUT_HESS) && !hess) return fail(GB_ERR_ARG, "Wrong number of components for the hessian");
    return GB_OK;
}

int packed_words(const gb_grid* g, int out_mask, int outmode) {
    if (!(out_mask & GB_OUT_PACKED)) return 0;
    return 1 + (outmode >= 1 ? g->ndim : 0) + (outmode >= 2 ? g->ndim * g->ndim : 0);
}

int check_affine(const gb_grid* g, const double* affine) {
    if (!affine) return GB_OK;
    for (int d = 0; d < g->ndim; d++) {
        const double kind = affine[4 * d];
        if (!(kind == 0.0 || kind == 1.0 || kind == 2.0)) return fail(GB_ERR_ARG, "affine kind must be 0 (FloatRange), 1 (StepRange) or 2 (UnitRange)");
        if (kind != 2.0 && affine[4 * d + 2] == 0.0) return fail(GB_ERR_ARG, "range step cannot be zero");
    }
    return GB_OK;
}

// page-locks host ranges that are not pinned yet; released (and only those) when the call returns
struct HostPins {
    std::vector<void*> mine;
    void add(const void* p, size_t bytes) {
        if (!p || bytes == 0) return;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) return;  // pinned / managed already
        cudaGetLastError();
        if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) mine.push_back(const_cast<void*>(p));
        else cudaGetLastError();  // (overlapping ranges, read-only mappings, ...: fall back to pageable copies)
    }
    ~HostPins() {
        for (void* p : mine) cudaHostUnregister(p);
    }
};

// ---- pageable host buffers ---------------------------------------------------------------------------------------------
// A Julia Array or a numpy array is pageable memory: cudaMemcpyAsync then goes through the driver's single staging thread and
// blocks the caller (measured on the C2 batch: 40-46 ms against 7.5 ms with pinned buffers), and page-locking the caller's
// ranges for one call costs more than it saves (cudaHostRegister: ~9 MB per ms).  So the library keeps its own small pinned
// staging slots (allocated once, two per worker) and a few host threads do what the driver's one thread does: worker w takes
// chunks w, w + T, ... of the batch -- memcpy of the chunk into a pinned slot, H2D, evaluation, D2H into a pinned slot, memcpy
// out -- double-buffered, so its copies overlap its own GPU work and every other worker's.  profiles/README.md has the numbers.
struct StageSlot {
    void *in = nullptr, *out = nullptr;
    size_t in_bytes = 0, out_bytes = 0;
};
struct StageCache {
    std::mutex mu;  // one staged call at a time per device
    std::vector<StageSlot> slots;
};
constexpr int STAGE_MAX_DEV = 64;
StageCache g_stage_dev[STAGE_MAX_DEV];  // per device: gb_eval_sharded runs one host-buffer call per device at the same time
std::atomic<int> g_stage_calls{0};      // staged calls in flight (they share the host's cores)
bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
int stage_threads() {
    static const int v = [] {
        const char* e = getenv("GB_HOST_THREADS");
        if (e) return std::max(0, std::min(32, atoi(e)));  // 0: the driver's pageable copies
        const unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(8u, hc / 2));
    }();
    return v;
}
int eval_host_staged(const gb_grid* g, int out_mask, int outmode, i64 nq, const void* const* xcols, const void* xaos, int xdtype, const double* affine,
                     void* val, void* grad, void* hess, int flags, int64_t* first_invalid, int nthreads) {
    const int N = g->ndim;
    const size_t xs = esize(xdtype), es = esize(g->dtype) * (size_t)g->nchan;
    const int packed = packed_words(g, out_mask, outmode);
    static const i64 chunk_env = [] {
        const char* e = getenv("GB_STAGE_CHUNK");
        long long v = e ? atoll(e) : 0;
        return (i64)(v > 0 ? v : (1ll << 18));
    }();
    const i64 chunk = std::max<i64>(1, std::min<i64>(chunk_env, nq));
    const i64 nchunk = (nq + chunk - 1) / chunk;
    struct InFlight {
        int others;
        InFlight() : others(g_stage_calls.fetch_add(1)) {}
        ~InFlight() { g_stage_calls.fetch_sub(1); }
    } inflight;
    // (concurrent calls -- one per device under gb_eval_sharded -- split the host threads between them, two at least each)
    const int T = (int)std::max<i64>(1, std::min<i64>(std::max(2, nthreads / (inflight.others + 1)), nchunk));
    StageCache& g_stage = g_stage_dev[g->device >= 0 && g->device < STAGE_MAX_DEV ? g->device : 0];
    // one output slot: [val | grad | hess] or the packed records
    const size_t vw = packed ? (size_t)packed : ((out_mask & GB_OUT_VAL) ? 1 : 0), gw = packed ? 0 : ((out_mask & GB_OUT_GRAD) ? (size_t)N : 0),
                 hw = packed ? 0 : ((out_mask & GB_OUT_HESS) ? (size_t)N * N : 0);
    const size_t in_bytes = (size_t)chunk * N * xs, out_bytes = (size_t)chunk * (vw + gw + hw) * es;
    const size_t off_g = (size_t)chunk * vw * es, off_h = off_g + (size_t)chunk * gw * es;
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (g_stage.slots.size() < (size_t)2 * T) g_stage.slots.resize((size_t)2 * T);
    for (int i = 0; i < 2 * T; i++) {  // (grown lazily, kept for the life of the process: a few MB per slot)
        StageSlot& sl = g_stage.slots[i];
        if (sl.in_bytes < in_bytes) {
            if (sl.in) cudaFreeHost(sl.in);
            sl.in = nullptr; sl.in_bytes = 0;
            if (cudaHostAlloc(&sl.in, in_bytes, cudaHostAllocDefault) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaHostAlloc (staging)");
            sl.in_bytes = in_bytes;
        }
        if (sl.out_bytes < out_bytes) {
            if (sl.out) cudaFreeHost(sl.out);
            sl.out = nullptr; sl.out_bytes = 0;
            if (cudaHostAlloc(&sl.out, out_bytes, cudaHostAllocDefault) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaHostAlloc (staging)");
            sl.out_bytes = out_bytes;
        }
    }
    u64* d_bad = nullptr;
    if (first_invalid && g->bc == GB_BC_NIL) {
        GB_CUDA(cudaMalloc((void**)&d_bad, sizeof(u64)));
        GB_CUDA(cudaMemset(d_bad, 0xFF, sizeof(u64)));
    }
    std::vector<int> rcs(T, GB_OK);
    std::vector<std::string> msgs(T);
    auto worker = [&](int w) {
        int rc = GB_OK;
        auto cu = [&](cudaError_t e, const char* what) {
            if (e != cudaSuccess && rc == GB_OK) rc = cuda_fail(e, what);
            return e == cudaSuccess;
        };
        cu(cudaSetDevice(g->device), "cudaSetDevice");
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        void *din[2] = {nullptr, nullptr}, *dout[2] = {nullptr, nullptr};
        if (rc == GB_OK) cu(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
        if (rc == GB_OK && g->ready) cu(cudaStreamWaitEvent(st, g->ready, 0), "cudaStreamWaitEvent");
        for (int i = 0; i < 2 && rc == GB_OK; i++) {
            cu(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync), "cudaEventCreate");
            cu(cudaMallocAsync(&din[i], in_bytes, st), "cudaMallocAsync");
            cu(cudaMallocAsync(&dout[i], out_bytes, st), "cudaMallocAsync");
        }
        i64 pend[2] = {-1, -1};
        auto copy_out = [&](int sl) {  // pinned slot -> the caller's arrays
            const i64 c = pend[sl];
            if (c < 0) return;
            cu(cudaEventSynchronize(ev[sl]), "cudaEventSynchronize");
            const i64 q0 = c * chunk, n = std::min(chunk, nq - q0);
            const char* src = (const char*)g_stage.slots[2 * w + sl].out;
            if (rc == GB_OK) {
                if (vw) std::memcpy((char*)val + (size_t)q0 * vw * es, src, (size_t)n * vw * es);
                if (gw) std::memcpy((char*)grad + (size_t)q0 * gw * es, src + off_g, (size_t)n * gw * es);
                if (hw) std::memcpy((char*)hess + (size_t)q0 * hw * es, src + off_h, (size_t)n * hw * es);
            }
            pend[sl] = -1;
        };
        int k = 0;
        for (i64 c = w; c < nchunk && rc == GB_OK; c += T, k++) {
            const int sl = k & 1;
            copy_out(sl);  // (waits for the chunk that used this slot two rounds ago)
            const i64 q0 = c * chunk, n = std::min(chunk, nq - q0);
            char* pin = (char*)g_stage.slots[2 * w + sl].in;
            EvalCall call;
            call.g = g; call.outmode = outmode; call.flags = flags; call.xdtype = xdtype; call.affine = affine;
            call.nq = n; call.q0 = q0; call.first_invalid = d_bad;
            if (xaos) {
                std::memcpy(pin, (const char*)xaos + (size_t)q0 * N * xs, (size_t)n * N * xs);
                for (int d = 0; d < N; d++) call.xp[d] = (const char*)din[sl] + (size_t)d * xs;
                call.xqs = N;
            } else {
                for (int d = 0; d < N; d++) {
                    std::memcpy(pin + (size_t)d * n * xs, (const char*)xcols[d] + (size_t)q0 * xs, (size_t)n * xs);
                    call.xp[d] = (const char*)din[sl] + (size_t)d * n * xs;
                }
                call.xqs = 1;
            }
            cu(cudaMemcpyAsync(din[sl], pin, (size_t)n * N * xs, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync H2D");
            call.val = vw ? dout[sl] : nullptr;
            call.grad = gw ? (char*)dout[sl] + off_g : nullptr;
            call.hess = hw ? (char*)dout[sl] + off_h : nullptr;
            call.packed = packed;
            if (rc != GB_OK) break;
            rc = launch_eval(call, st);
            if (rc != GB_OK) break;
            char* pout = (char*)g_stage.slots[2 * w + sl].out;
            if (vw) cu(cudaMemcpyAsync(pout, dout[sl], (size_t)n * vw * es, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H");
            if (gw) cu(cudaMemcpyAsync(pout + off_g, (char*)dout[sl] + off_g, (size_t)n * gw * es, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H");
            if (hw) cu(cudaMemcpyAsync(pout + off_h, (char*)dout[sl] + off_h, (size_t)n * hw * es, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H");
            cu(cudaEventRecord(ev[sl], st), "cudaEventRecord");
            pend[sl] = c;
        }
        copy_out(k & 1);        // the older of the two chunks still in flight ...
        copy_out((k & 1) ^ 1);  // ... and the last one
        if (st) {
            for (int i = 0; i < 2; i++) {
                if (din[i]) cudaFreeAsync(din[i], st);
                if (dout[i]) cudaFreeAsync(dout[i], st);
            }
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        for (int i = 0; i < 2; i++)
            if (ev[i]) cudaEventDestroy(ev[i]);
        rcs[w] = rc;
        if (rc != GB_OK) msgs[w] = gb_last_error();
    };
    std::vector<std::thread> th;
    std::vector<int> inline_w;  // workers whose thread could not be created run here, one after the other
    for (int w = 1; w < T; w++) {
        try {
            th.emplace_back(worker, w);
        } catch (const std::exception&) {
            inline_w.push_back(w);
        }
    }
    worker(0);
    for (int w : inline_w) worker(w);
    for (auto& t : th) t.join();
    int rc = GB_OK;
    for (int w = 0; w < T && rc == GB_OK; w++)
        if (rcs[w] != GB_OK) rc = fail(rcs[w], msgs[w]);
    if (first_invalid) *first_invalid = -1;
    if (d_bad) {
        u64 bad = ~0ull;
        if (cudaMemcpy(&bad, d_bad, sizeof(u64), cudaMemcpyDeviceToHost) != cudaSuccess && rc == GB_OK) rc = cuda_fail(cudaGetLastError(), "cudaMemcpy");
        cudaFree(d_bad);
        if (rc == GB_OK && bad != ~0ull) {
            *first_invalid = (int64_t)bad;
            rc = fail(GB_ERR_INVALID_LOCATION, "Invalid location");
        }
    }
    return rc;
}

// ---- small host batches: the scalar API ----------------------------------------------------------------------------------
// G[x], valgrad(G, x) and short batches arrive here one C call at a time (GridB200.jl evaluates a scalar query per ccall):
// three stream creations, a dozen pool allocations and as many frees cost 150-170 us per call, ten times the work itself.
// Up to SMALL_MAX queries therefore go through a context that is created once and reused -- one stream, one pinned and one
// device staging buffer (layout: coordinates | first-invalid word | outputs): memcpy in, ONE copy to the device, the kernel,
// ONE copy back, synchronise, memcpy out.  Contexts live in a per-device free list (a call takes one, returns it): as many
// exist as calls ever ran at the same time, whatever the number of host threads that come and go.  (Never destroyed: nothing
// may call CUDA at process exit.)
constexpr i64 SMALL_MAX = 1 << 15;
struct SmallCtx {
    cudaStream_t s;
    char *pin, *dev;
    size_t bytes;
};
struct SmallPool {
    std::mutex mu;
    std::vector<SmallCtx*> free_list;
};
SmallPool g_small[STAGE_MAX_DEV];
struct SmallLease {
    SmallPool& pool;
    SmallCtx* c;
    explicit SmallLease(SmallPool& p) : pool(p), c(nullptr) {
        std::lock_guard<std::mutex> lk(pool.mu);
        if (!pool.free_list.empty()) {
            c = pool.free_list.back();
            pool.free_list.pop_back();
        }
        if (!c) c = new SmallCtx{nullptr, nullptr, nullptr, 0};
    }
    ~SmallLease() {
        std::lock_guard<std::mutex> lk(pool.mu);
        pool.free_list.push_back(c);
    }
};
int eval_host_small(const gb_grid* g, int out_mask, int outmode, i64 nq, const void* const* xcols, const void* xaos, int xdtype, const double* affine,
                    void* val, void* grad, void* hess, int flags, int64_t* first_invalid) {
    if (g->device < 0 || g->device >= STAGE_MAX_DEV) return fail(GB_ERR_ARG, "device index out of range");
    SmallLease lease(g_small[g->device]);
    SmallCtx& c = *lease.c;
    const int N = g->ndim;
    const size_t xs = esize(xdtype), es = esize(g->dtype) * (size_t)g->nchan;
    const int packed = packed_words(g, out_mask, outmode);
    const size_t vw = packed ? (size_t)packed : ((out_mask & GB_OUT_VAL) ? 1 : 0), gw = packed ? 0 : ((out_mask & GB_OUT_GRAD) ? (size_t)N : 0),
                 hw = packed ? 0 : ((out_mask & GB_OUT_HESS) ? (size_t)N * N : 0);
    const size_t bytes_x = (size_t)nq * N * xs;
    const size_t off_bad = (bytes_x + 15) & ~(size_t)15, off_v = off_bad + 16;
    const size_t off_g = off_v + (((size_t)nq * vw * es + 15) & ~(size_t)15), off_h = off_g + (((size_t)nq * gw * es + 15) & ~(size_t)15);
    const size_t total = off_h + (size_t)nq * hw * es;
    if (!c.s) GB_CUDA(cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking));
    if (c.bytes < total) {
        size_t want = 1 << 16;
        while (want < total) want <<= 1;
        if (c.pin) cudaFreeHost(c.pin);
        if (c.dev) cudaFree(c.dev);
        c.pin = c.dev = nullptr;
        c.bytes = 0;
        GB_CUDA(cudaHostAlloc((void**)&c.pin, want, cudaHostAllocDefault));
        GB_CUDA(cudaMalloc((void**)&c.dev, want));
        c.bytes = want;
    }
    const bool check = first_invalid && g->bc == GB_BC_NIL;
    // A handful of queries (the scalar API): no copies at all -- the kernel reads the coordinates from, and writes the results to,
    // the pinned buffer itself (mapped into the device's address space: a few dozen bytes over PCIe each way), which saves
    // two copy submissions per call.  BCnil's first-invalid word needs a device atomic: such calls copy.
    static const i64 zc_max = [] { const char* e = getenv("GB_ZERO_COPY_MAX"); return (i64)(e ? atoll(e) : 64); }();
    const bool zero_copy = nq <= zc_max && !check;
    char* const base = zero_copy ? c.pin : c.dev;  // what the kernel reads and writes
    EvalCall call;
    call.g = g; call.outmode = outmode; call.flags = flags; call.xdtype = xdtype; call.affine = affine;
    call.nq = nq; call.q0 = 0;
    call.first_invalid = check ? reinterpret_cast<u64*>(c.dev + off_bad) : nullptr;
    if (xaos) {
        std::memcpy(c.pin, xaos, bytes_x);
        for (int d = 0; d < N; d++) call.xp[d] = base + (size_t)d * xs;
        call.xqs = N;
    } else {
        for (int d = 0; d < N; d++) {
            std::memcpy(c.pin + (size_t)d * nq * xs, xcols[d], (size_t)nq * xs);
            call.xp[d] = base + (size_t)d * nq * xs;
        }
        call.xqs = 1;
    }
    *reinterpret_cast<u64*>(c.pin + off_bad) = ~0ull;
    call.val = vw ? base + off_v : nullptr;
    call.grad = gw ? base + off_g : nullptr;
    call.hess = hw ? base + off_h : nullptr;
    call.packed = packed;
    if (g->ready) GB_CUDA(cudaStreamWaitEvent(c.s, g->ready, 0));  // a grid built on another stream
    int rc;
    if (zero_copy) {
        rc = launch_eval(call, c.s);
    } else {
        GB_CUDA(cudaMemcpyAsync(c.dev, c.pin, off_v, cudaMemcpyHostToDevice, c.s));
        rc = launch_eval(call, c.s);
        if (rc == GB_OK) {
            cudaError_t e = cudaMemcpyAsync(c.pin + off_bad, c.dev + off_bad, total - off_bad, cudaMemcpyDeviceToHost, c.s);
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H");
        }
    }
    const cudaError_t es_ = cudaStreamSynchronize(c.s);
    if (rc == GB_OK && es_ != cudaSuccess) rc = cuda_fail(es_, "cudaStreamSynchronize");
    if (first_invalid) *first_invalid = -1;
    if (rc != GB_OK) return rc;
    if (vw) std::memcpy(val, c.pin + off_v, (size_t)nq * vw * es);
    if (gw) std::memcpy(grad, c.pin + off_g, (size_t)nq * gw * es);
    if (hw) std::memcpy(hess, c.pin + off_h, (size_t)nq * hw * es);
    if (check) {
        const u64 bad = *reinterpret_cast<const u64*>(c.pin + off_bad);
        if (bad != ~0ull) {
            *first_invalid = (int64_t)bad;
            return fail(GB_ERR_INVALID_LOCATION, "Invalid location");
        }
    }
    return GB_OK;
}

// Host-resident evaluation: chunks of the batch flow H2D -> kernel -> D2H over a ring of streams so
// PCIe traffic in both directions overlaps the kernels.
int eval_host(const gb_grid* g, int out_mask, int outmode, i64 nq, const void* const* xcols, const void* xaos, int xdtype,
              const double* affine, void* val, void* grad, void* hess, int flags, int64_t* first_invalid) {
    const int N = g->ndim;
    const size_t xs = esize(xdtype), es = esize(g->dtype) * (size_t)g->nchan;  // es: bytes of one query's values (all channels)
    static const i64 chunk_env = [] {
        const char* e = getenv("GB_EVAL_CHUNK");
        long long v = e ? atoll(e) : 0;
        return (i64)(v > 0 ? v : (1ll << 20));
    }();
    const i64 chunk = std::max<i64>(1, std::min<i64>(chunk_env, nq));
    {
        static const bool small_off = [] { const char* e = getenv("GB_NO_SMALL_PATH"); return e && *e == '1'; }();
        if (nq <= SMALL_MAX && !small_off) return eval_host_small(g, out_mask, outmode, nq, xcols, xaos, xdtype, affine, val, grad, hess, flags, first_invalid);
    }
    // pageable buffers, a batch worth the threads: the library's own pinned staging slots (eval_host_staged)
    {
        static const int reg_env0 = [] { const char* e = getenv("GB_HOST_REGISTER"); return e ? atoi(e) : 0; }();
        const int nt = stage_threads();
        if (nt > 0 && reg_env0 != 1 && nq >= (1ll << 19) && is_pageable(xaos ? xaos : xcols[0]) &&
            is_pageable(val ? val : (grad ? grad : hess)))
            return eval_host_staged(g, out_mask, outmode, nq, xcols, xaos, xdtype, affine, val, grad, hess, flags, first_invalid, nt);
    }
    // Pageable host buffers (a Julia Array, a numpy array): every cudaMemcpyAsync then goes through the driver's staging
    // buffer and blocks the host, so the chunks stop overlapping (measured on C2: 44 ms against 7.7 ms with pinned buffers).
    // Page-locking the caller's ranges for the duration of ONE call does not pay (cudaHostRegister pins ~9 MB per ms on the
    // GPU box: 65 ms for the C2 batch, profiles/README.md) -- callers that reuse their buffers should pin them once
    // (gb_host_register; GridB200.pin!).  GB_HOST_REGISTER=1 makes the call do it anyway (buffers that are not pinned yet).
    HostPins pins;
    {
        static const int reg_env = [] { const char* e = getenv("GB_HOST_REGISTER"); return e ? atoi(e) : 0; }();
        const size_t xbytes = (size_t)nq * N * xs, vbytes = (size_t)nq * es;
        const bool want = reg_env == 1;
        if (want) {
            if (xaos) pins.add(xaos, xbytes);
            else for (int d = 0; d < N; d++) pins.add(xcols[d], (size_t)nq * xs);
            const int pw = packed_words(g, out_mask, outmode);
            if (pw) pins.add(val, vbytes * pw);
            else {
                if (out_mask & GB_OUT_VAL) pins.add(val, vbytes);
                if (out_mask & GB_OUT_GRAD) pins.add(grad, vbytes * N);
                if (out_mask & GB_OUT_HESS) pins.add(hess, vbytes * N * N);
            }
        }
    }
    constexpr int NSLOT = 3;
    const int nslot = (int)std::min<i64>(NSLOT, (nq + chunk - 1) / chunk);
    struct Slot {
        cudaStream_t s = nullptr;
        void *x = nullptr, *v = nullptr, *gr = nullptr, *h = nullptr;
    } slots[NSLOT];
    u64* d_bad = nullptr;
    int rc = GB_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == GB_OK) rc = cuda_fail(e, what);
        return e == cudaSuccess;
    };
    const int packed = packed_words(g, out_mask, outmode);
    for (int i = 0; i < nslot && rc == GB_OK; i++) {
        Slot& sl = slots[i];
        if (!cu(cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking), "cudaStreamCreate")) break;
        if (g->ready) cu(cudaStreamWaitEvent(sl.s, g->ready, 0), "cudaStreamWaitEvent");  // a grid built on another stream
        cu(cudaMallocAsync(&sl.x, (size_t)chunk * N * xs, sl.s), "cudaMallocAsync");
        if (packed) {
            cu(cudaMallocAsync(&sl.v, (size_t)chunk * packed * es, sl.s), "cudaMallocAsync");
            continue;
        }
        if (out_mask & GB_OUT_VAL) cu(cudaMallocAsync(&sl.v, (size_t)chunk * es, sl.s), "cudaMallocAsync");
        if (out_mask & GB_OUT_GRAD) cu(cudaMallocAsync(&sl.gr, (size_t)chunk * N * es, sl.s), "cudaMallocAsync");
        if (out_mask & GB_OUT_HESS) cu(cudaMallocAsync(&sl.h, (size_t)chunk * N * N * es, sl.s), "cudaMallocAsync");
    }
    if (rc == GB_OK && first_invalid && g->bc == GB_BC_NIL) {
        cu(cudaMallocAsync((void**)&d_bad, sizeof(u64), slots[0].s), "cudaMallocAsync");
        cu(cudaMemsetAsync(d_bad, 0xFF, sizeof(u64), slots[0].s), "cudaMemsetAsync");
        cu(cudaStreamSynchronize(slots[0].s), "cudaStreamSynchronize");
    }
    // GB_EVAL_TRACE=1: CUDA events around every chunk's copy-in, evaluation and copy-out; the timeline goes to stderr
    // (profiles/trace_e2e.py, profiles/r2_trace_e2e.txt)
    static const bool trace = [] { const char* e = getenv("GB_EVAL_TRACE"); return e && *e == '1'; }();
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); tev.push_back(e); }
    };
    for (i64 q0 = 0, c = 0; q0 < nq && rc == GB_OK; q0 += chunk, c++) {
        Slot& sl = slots[c % nslot];
        const i64 n = std::min(chunk, nq - q0);
        mark(sl.s);
        EvalCall call;
        call.g = g; call.outmode = outmode; call.flags = flags; call.xdtype = xdtype; call.affine = affine;
        call.nq = n; call.q0 = q0; call.first_invalid = d_bad;
        if (xaos) {  // N contiguous coordinates per query
            cu(cudaMemcpyAsync(sl.x, (const char*)xaos + (size_t)q0 * N * xs, (size_t)n * N * xs, cudaMemcpyHostToDevice, sl.s), "cudaMemcpyAsync H2D");
            for (int d = 0; d < N; d++) call.xp[d] = (const char*)sl.x + (size_t)d * xs;
            call.xqs = N;
        } else {  // one vector per dimension
            for (int d = 0; d < N; d++) {
                cu(cudaMemcpyAsync((char*)sl.x + (size_t)d * chunk * xs, (const char*)xcols[d] + (size_t)q0 * xs, (size_t)n * xs, cudaMemcpyHostToDevice, sl.s), "cudaMemcpyAsync H2D");
                call.xp[d] = (const char*)sl.x + (size_t)d * chunk * xs;
            }
            call.xqs = 1;
        }
        call.val = sl.v; call.grad = sl.gr; call.hess = sl.h;
        call.packed = packed;
        if (rc != GB_OK) break;
        mark(sl.s);
        rc = launch_eval(call, sl.s);
        if (rc != GB_OK) break;
        mark(sl.s);
        if (packed) {
            cu(cudaMemcpyAsync((char*)val + (size_t)q0 * packed * es, sl.v, (size_t)n * packed * es, cudaMemcpyDeviceToHost, sl.s), "cudaMemcpyAsync D2H");
        } else {
            if (sl.v) cu(cudaMemcpyAsync((char*)val + (size_t)q0 * es, sl.v, (size_t)n * es, cudaMemcpyDeviceToHost, sl.s), "cudaMemcpyAsync D2H");
            if (sl.gr) cu(cudaMemcpyAsync((char*)grad + (size_t)q0 * N * es, sl.gr, (size_t)n * N * es, cudaMemcpyDeviceToHost, sl.s), "cudaMemcpyAsync D2H");
            if (sl.h) cu(cudaMemcpyAsync((char*)hess + (size_t)q0 * N * N * es, sl.h, (size_t)n * N * N * es, cudaMemcpyDeviceToHost, sl.s), "cudaMemcpyAsync D2H");
        }
        mark(sl.s);
    }
    for (int i = 0; i < nslot; i++) {
        Slot& sl = slots[i];
        if (!sl.s) continue;
        if (sl.x) cudaFreeAsync(sl.x, sl.s);
        if (sl.v) cudaFreeAsync(sl.v, sl.s);
        if (sl.gr) cudaFreeAsync(sl.gr, sl.s);
        if (sl.h) cudaFreeAsync(sl.h, sl.s);
        cu(cudaStreamSynchronize(sl.s), "cudaStreamSynchronize");
    }
    if (trace && tev.size() >= 4) {  // (every slot stream has been synchronised above)
        fprintf(stderr, "[gb_eval trace] chunk: copy-in start .. end | eval end | copy-out end   (ms since the first copy-in)\n");
        for (size_t c = 0; c + 3 < tev.size(); c += 4) {
            float t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], tev[0], tev[c + k]);
            fprintf(stderr, "[gb_eval trace] %2zu: %7.3f .. %7.3f | %7.3f | %7.3f\n", c / 4, t[0], t[1], t[2], t[3]);
        }
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    if (first_invalid) *first_invalid = -1;
    if (d_bad) {
        u64 bad = ~0ull;
        cu(cudaMemcpy(&bad, d_bad, sizeof(u64), cudaMemcpyDeviceToHost), "cudaMemcpy");
        cudaFree(d_bad);
        if (rc == GB_OK && bad != ~0ull) {
            *first_invalid = (int64_t)bad;
            rc = fail(GB_ERR_INVALID_LOCATION, "Invalid location");
        }
    }
    for (int i = 0; i < nslot; i++)
        if (slots[i].s) cudaStreamDestroy(slots[i].s);
    return rc;
}

int eval_device(const gb_grid* g, int outmode, i64 nq, const void* const* xp, i64 xqs, int xdtype, const double* affine,
                void* val, void* grad, void* hess, int flags, int64_t* first_invalid, cudaStream_t s, int packed = 0) {
    EvalCall call;
    call.packed = packed;
    call.g = g; call.outmode = outmode; call.flags = flags; call.xdtype = xdtype; call.affine = affine;
    call.nq = nq; call.q0 = 0; call.xqs = xqs;
    for (int d = 0; d < g->ndim; d++) call.xp[d] = xp[d];
    call.val = val; call.grad = grad; call.hess = hess;
    {
        const int rcw = wait_ready(g, s);
        if (rcw) return rcw;
    }
    u64* d_bad = nullptr;
    const bool check = first_invalid && g->bc == GB_BC_NIL;
    if (check) {
        GB_CUDA(cudaMallocAsync((void**)&d_bad, sizeof(u64), s));
        GB_CUDA(cudaMemsetAsync(d_bad, 0xFF, sizeof(u64), s));
    }
    call.first_invalid = d_bad;
    int rc = launch_eval(call, s);
    if (first_invalid) *first_invalid = -1;
    if (check) {
        u64 bad = ~0ull;
        cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(u64), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFreeAsync(d_bad, s);
        if (rc == GB_OK && e != cudaSuccess) rc = cuda_fail(e, "read first_invalid");
        if (rc == GB_OK && bad != ~0ull) {
            *first_invalid = (int64_t)bad;
            rc = fail(GB_ERR_INVALID_LOCATION, "Invalid location");
        }
    }
    return rc;
}

// out[i1,i2,...] = G[v1[i1], v2[i2], ...]: expand the vectors into per-dimension coordinate columns.
template <typename XT> __global__ void expand_outer_kernel(XT* __restrict__ cols, i64 nq, int N, const XT* const* __restrict__ vecs, const i64* __restrict__ lens) {
    const i64 step = (i64)gridDim.x * blockDim.x;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += step) {
        i64 rem = q;
        for (int d = 0; d < N; d++) {
            const i64 i = rem % lens[d];
            rem /= lens[d];
            cols[(i64)d * nq + q] = vecs[d][i];
        }
    }
}

}  // namespace

extern "C" {

int gb_eval(const gb_grid* g, int out_mask, int64_t nq, const void* x, int xdtype, int64_t x_qstride, int64_t x_dstride,
            const double* affine, void* val, void* grad, void* hess, int mem, int flags, int64_t* first_invalid, void* stream) {
    if (!g) return fail(GB_ERR_ARG, "grid is NULL");
    if (nq < 0) return fail(GB_ERR_ARG, "nq < 0");
    if (!ok_dtype(xdtype)) return fail(GB_ERR_ARG, "xdtype must be GB_F32 or GB_F64");
    int outmode = 0;
    int rc = outmode_of(g, out_mask, val, grad, hess, &outmode);
    if (rc) return rc;
    rc = check_affine(g, affine);
    if (rc) return rc;
    if (first_invalid) *first_invalid = -1;
    if (nq == 0) return GB_OK;
    if (!x) return fail(GB_ERR_ARG, "x is NULL");
    DeviceGuard guard(g->device);
    pool_keep_memory();
    const int N = g->ndim;
    const void* cols[MAXG];
    for (int d = 0; d < N; d++) cols[d] = (const char*)x + (size_t)d * x_dstride * esize(xdtype);
    if (mem == GB_DEVICE) {
        const int pw = packed_words(g, out_mask, outmode);
        return eval_device(g, outmode, nq, cols, x_qstride, xdtype, affine, (pw || (out_mask & GB_OUT_VAL)) ? val : nullptr,
                           (!pw && (out_mask & GB_OUT_GRAD)) ? grad : nullptr, (!pw && (out_mask & GB_OUT_HESS)) ? hess : nullptr, flags, first_invalid,
                           (cudaStream_t)stream, pw);
    }
    if (x_qstride == N && x_dstride == 1)
        return eval_host(g, out_mask, outmode, nq, nullptr, x, xdtype, affine, val, grad, hess, flags, first_invalid);
    if (x_qstride == 1)
        return eval_host(g, out_mask, outmode, nq, cols, nullptr, xdtype, affine, val, grad, hess, flags, first_invalid);
    return fail(GB_ERR_ARG, "host query batches must be N x nq (x_qstride=N, x_dstride=1) or column vectors (x_qstride=1)");
}

int gb_eval_soa(const gb_grid* g, int out_mask, int64_t nq, const void* const* xs, int xdtype, const double* affine, void* val,
                void* grad, void* hess, int mem, int flags, int64_t* first_invalid, void* stream) {
    if (!g) return fail(GB_ERR_ARG, "grid is NULL");
    if (nq < 0) return fail(GB_ERR_ARG, "nq < 0");
    if (!ok_dtype(xdtype)) return fail(GB_ERR_ARG, "xdtype must be GB_F32 or GB_F64");
    int outmode = 0;
    int rc = outmode_of(g, out_mask, val, grad, hess, &outmode);
    if (rc) return rc;
    rc = check_affine(g, affine);
    if (rc) return rc;
    if (first_invalid) *first_invalid = -1;
    if (nq == 0) return GB_OK;
    if (!xs) return fail(GB_ERR_ARG, "xs is NULL");
    for (int d = 0; d < g->ndim; d++)
        if (!xs[d]) return fail(GB_ERR_ARG, "Incorrect number of dimensions supplied");
    DeviceGuard guard(g->device);
    pool_keep_memory();
    if (mem == GB_DEVICE) {
        const int pw = packed_words(g, out_mask, outmode);
        return eval_device(g, outmode, nq, xs, 1, xdtype, affine, (pw || (out_mask & GB_OUT_VAL)) ? val : nullptr,
                           (!pw && (out_mask & GB_OUT_GRAD)) ? grad : nullptr, (!pw && (out_mask & GB_OUT_HESS)) ? hess : nullptr, flags, first_invalid,
                           (cudaStream_t)stream, pw);
    }
    return eval_host(g, out_mask, outmode, nq, xs, nullptr, xdtype, affine, val, grad, hess, flags, first_invalid);
}

int gb_eval_outer(const gb_grid* g, const int64_t* lens, const void* const* vecs, int xdtype, const double* affine, void* out,
                  int mem, int flags, void* stream) {
    if (!g || !lens || !vecs || !out) return fail(GB_ERR_ARG, "NULL argument");
    if (!ok_dtype(xdtype)) return fail(GB_ERR_ARG, "xdtype must be GB_F32 or GB_F64");
    int rc = check_affine(g, affine);
    if (rc) return rc;
    if (g->nchan > 1) return fail(GB_ERR_UNSUPPORTED, "tensor-product getindex is not provided for multi-valued grids");
    const int N = g->ndim;
    i64 nq = 1;
    for (int d = 0; d < N; d++) {
        if (lens[d] < 0 || !vecs[d]) return fail(GB_ERR_ARG, "Dimensionality mismatch");
        nq *= lens[d];
    }
    if (nq == 0) return GB_OK;
    DeviceGuard guard(g->device);
    pool_keep_memory();
    cudaStream_t s = (cudaStream_t)stream;
    rc = wait_ready(g, s);
    if (rc) return rc;
    const size_t xs = esize(xdtype), es = esize(g->dtype) * (size_t)g->nchan;
    // device copies of the vectors (host mode), the pointer table, lens, the expanded columns
    void* dvec[MAXG] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const void* vptr[MAXG];
    void *d_tab = nullptr, *d_lens = nullptr, *d_cols = nullptr, *d_out = nullptr;
    auto cleanup = [&](int code) {
        for (int d = 0; d < N; d++)
            if (dvec[d]) cudaFreeAsync(dvec[d], s);
        if (d_tab) cudaFreeAsync(d_tab, s);
        if (d_lens) cudaFreeAsync(d_lens, s);
        if (d_cols) cudaFreeAsync(d_cols, s);
        if (d_out) cudaFreeAsync(d_out, s);
        if (mem == GB_HOST) cudaStreamSynchronize(s);
        return code;
    };
#define GB_TRY(expr)                                                   \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return cleanup(cuda_fail(_e, #expr));   \
    } while (0)
    for (int d = 0; d < N; d++) {
        if (mem == GB_HOST) {
            GB_TRY(cudaMallocAsync(&dvec[d], std::max<size_t>(1, lens[d] * xs), s));
            GB_TRY(cudaMemcpyAsync(dvec[d], vecs[d], lens[d] * xs, cudaMemcpyHostToDevice, s));
            vptr[d] = dvec[d];
        } else
            vptr[d] = vecs[d];
    }
    if (outer_separable(g)) {  // weights / taps once per axis point, then a pure gather-contract kernel
        void* dst = out;
        if (mem == GB_HOST) {
            GB_TRY(cudaMallocAsync(&d_out, (size_t)nq * es, s));
            dst = d_out;
        }
        u64* d_bad = nullptr;
        if (g->bc == GB_BC_NIL) {
            GB_TRY(cudaMallocAsync((void**)&d_bad, sizeof(u64), s));
            GB_TRY(cudaMemsetAsync(d_bad, 0xFF, sizeof(u64), s));
        }
        EvalCall call;
        call.g = g; call.outmode = 0; call.flags = flags; call.xdtype = xdtype; call.affine = affine;
        call.nq = nq; call.q0 = 0; call.xqs = 1; call.first_invalid = d_bad;
        for (int d = 0; d < N; d++) call.xp[d] = vptr[d];
        call.val = dst; call.grad = nullptr; call.hess = nullptr;
        switch (N) {
        case 1: rc = launch_outer_n<1>(call, (const i64*)lens, dst, s); break;
        case 2: rc = launch_outer_n<2>(call, (const i64*)lens, dst, s); break;
        case 3: rc = launch_outer_n<3>(call, (const i64*)lens, dst, s); break;
        default: rc = launch_outer_n<4>(call, (const i64*)lens, dst, s); break;
        }
        if (rc == GB_OK && mem == GB_HOST) GB_TRY(cudaMemcpyAsync(out, d_out, (size_t)nq * es, cudaMemcpyDeviceToHost, s));
        if (d_bad) {
            u64 bad = ~0ull;
            if (rc == GB_OK) {
                GB_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(u64), cudaMemcpyDeviceToHost, s));
                GB_TRY(cudaStreamSynchronize(s));
            }
            cudaFreeAsync(d_bad, s);
            if (rc == GB_OK && bad != ~0ull) rc = fail(GB_ERR_INVALID_LOCATION, "Invalid location");
        }
        return cleanup(rc);
    }
    GB_TRY(cudaMallocAsync(&d_tab, sizeof(void*) * MAXG, s));
    GB_TRY(cudaMemcpyAsync(d_tab, vptr, sizeof(void*) * N, cudaMemcpyHostToDevice, s));
    GB_TRY(cudaMallocAsync(&d_lens, sizeof(i64) * MAXG, s));
    GB_TRY(cudaMemcpyAsync(d_lens, lens, sizeof(i64) * N, cudaMemcpyHostToDevice, s));
    GB_TRY(cudaMallocAsync(&d_cols, (size_t)nq * N * xs, s));
    const int grid = (int)std::max<i64>(1, std::min<i64>((nq + 255) / 256, (i64)num_sms() * 16));
    if (xdtype == GB_F64) expand_outer_kernel<double><<<grid, 256, 0, s>>>((double*)d_cols, nq, N, (const double* const*)d_tab, (const i64*)d_lens);
    else expand_outer_kernel<float><<<grid, 256, 0, s>>>((float*)d_cols, nq, N, (const float* const*)d_tab, (const i64*)d_lens);
    count_launch();
    GB_TRY(cudaGetLastError());
    void* dst = out;
    if (mem == GB_HOST) {
        GB_TRY(cudaMallocAsync(&d_out, (size_t)nq * es, s));
        dst = d_out;
    }
    const void* cols[MAXG];
    for (int d = 0; d < N; d++) cols[d] = (const char*)d_cols + (size_t)d * nq * xs;
    int64_t bad = -1;
    rc = eval_device(g, 0, nq, cols, 1, xdtype, affine, dst, nullptr, nullptr, flags, g->bc == GB_BC_NIL ? &bad : nullptr, s);
    if (rc == GB_OK && mem == GB_HOST) GB_TRY(cudaMemcpyAsync(out, d_out, (size_t)nq * es, cudaMemcpyDeviceToHost, s));
#undef GB_TRY
    return cleanup(rc);
}

// ---- InterpIrregular -------------------------------------------------------------------------------
int gb_interp_irregular(int dtype, int64_t n, const void* knots, const void* samples, int bc, int order, double fill, int64_t nq, const void* x, void* out,
                        int mem, int64_t* first_invalid, void* stream) {
    if (!ok_dtype(dtype)) return fail(GB_ERR_ARG, "dtype must be GB_F32 or GB_F64");
    if (n < 1 || !knots || !samples) return fail(GB_ERR_ARG, "InterpIrregular needs at least one knot");
    if (nq < 0 || (nq > 0 && (!x || !out))) return fail(GB_ERR_ARG, "bad query arguments");
    if (bc == GB_BC_REFLECT || bc == GB_BC_PERIODIC) return fail(GB_ERR_UNSUPPORTED, "Sorry, reflecting or periodic boundary conditions not yet supported");
    if (bc < GB_BC_NIL || bc > GB_BC_FILL) return fail(GB_ERR_ARG, "bad boundary condition");
    if (!(order == GB_FORWARD || order == GB_BACKWARD || order == GB_NEAREST || order == GB_LINEAR))
        return fail(GB_ERR_UNSUPPORTED, "InterpIrregular supports InterpForward, InterpBackward, InterpNearest and InterpLinear");
    if (first_invalid) *first_invalid = -1;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    const size_t es = esize(dtype);
    const double fv = (bc == GB_BC_FILL) ? fill : std::nan("");
    void *dk = nullptr, *dc = nullptr, *dx = nullptr, *dout = nullptr;
    u64* d_bad = nullptr;
    int rc = GB_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == GB_OK) rc = cuda_fail(e, what);
        return e == cudaSuccess;
    };
    const void *pk = knots, *pc = samples, *px = x;
    void* po = out;
    if (mem == GB_HOST) {
        cu(cudaMallocAsync(&dk, n * es, s), "cudaMallocAsync");
        cu(cudaMallocAsync(&dc, n * es, s), "cudaMallocAsync");
        cu(cudaMallocAsync(&dx, std::max<size_t>(1, nq * es), s), "cudaMallocAsync");
        cu(cudaMallocAsync(&dout, std::max<size_t>(1, nq * es), s), "cudaMallocAsync");
        if (rc == GB_OK) {
            cu(cudaMemcpyAsync(dk, knots, n * es, cudaMemcpyHostToDevice, s), "H2D");
            cu(cudaMemcpyAsync(dc, samples, n * es, cudaMemcpyHostToDevice, s), "H2D");
            if (nq) cu(cudaMemcpyAsync(dx, x, nq * es, cudaMemcpyHostToDevice, s), "H2D");
        }
        pk = dk; pc = dc; px = dx; po = dout;
    }
    if (rc == GB_OK && bc == GB_BC_NIL) {
        cu(cudaMallocAsync((void**)&d_bad, sizeof(u64), s), "cudaMallocAsync");
        cu(cudaMemsetAsync(d_bad, 0xFF, sizeof(u64), s), "cudaMemsetAsync");
    }
    if (rc == GB_OK) rc = irregular_device(dtype, pk, pc, n, bc, order, fv, px, po, nq, d_bad, s);
    if (rc == GB_OK && mem == GB_HOST && nq) cu(cudaMemcpyAsync(out, dout, nq * es, cudaMemcpyDeviceToHost, s), "D2H");
    u64 bad = ~0ull;
    if (rc == GB_OK && d_bad) {
        cu(cudaMemcpyAsync(&bad, d_bad, sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H");
        cu(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    }
    for (void* p : {dk, dc, dx, dout, (void*)d_bad})
        if (p) cudaFreeAsync(p, s);
    if (mem == GB_HOST) cu(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    if (rc == GB_OK && bad != ~0ull) {
        if (first_invalid) *first_invalid = (int64_t)bad;
        rc = fail(GB_ERR_INVALID_LOCATION, "BoundsError");
    }
    return rc;
}

// ---- restrict / prolong -----------------------------------------------------------------------
int64_t gb_restrict_size(int64_t len) { return (len & 1) ? (len + 1) / 2 : len / 2 + 1; }
int gb_prolong_size(int64_t n, int64_t len, int64_t* newlen) {
    const int64_t newsz = (len & 1) ? 2 * n - 1 : 2 * (n - 1);
    if (newsz != len) return fail(GB_ERR_PROLONG_SIZE, "Cannot prolong a dimension of length " + std::to_string(n) + " to " + std::to_string(len));
    if (newlen) *newlen = newsz;
    return GB_OK;
}

static int mg_host(int dtype, i64 in_elems, i64 out_elems, const void* in, void* out, cudaStream_t s,
                   int (*run)(void*, const void*, void*, cudaStream_t), void* ctx) {
    AsyncBuf di(s), dout(s);
    GB_CUDA(di.alloc(in_elems * esize(dtype)));
    GB_CUDA(dout.alloc(out_elems * esize(dtype)));
    GB_CUDA(cudaMemcpyAsync(di.p, in, in_elems * esize(dtype), cudaMemcpyHostToDevice, s));
    int rc = run(ctx, di.p, dout.p, s);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(out, dout.p, out_elems * esize(dtype), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
}

int gb_restrict(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "restrict: dim out of range");
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return restrict_device(dtype, ndim, (const i64*)dims, in, dim, scale, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= (d == dim) ? gb_restrict_size(dims[d]) : dims[d]; }
    struct Ctx { int dtype, ndim, dim; const i64* dims; double scale; } ctx{dtype, ndim, dim, (const i64*)dims, scale};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return restrict_device(x->dtype, x->ndim, x->dims, i, x->dim, x->scale, o, st);
    }, &ctx);
}

int gb_prolong(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t len, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "prolong: dim out of range");
    int rc = gb_prolong_size(dims[dim], len, nullptr);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return prolong_device(dtype, ndim, (const i64*)dims, in, dim, len, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= (d == dim) ? len : dims[d]; }
    struct Ctx { int dtype, ndim, dim; const i64* dims; i64 len; } ctx{dtype, ndim, dim, (const i64*)dims, len};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return prolong_device(x->dtype, x->ndim, x->dims, i, x->dim, x->len, o, st);
    }, &ctx);
}

// restrictb / restrict_extrap / prolongb (src/restrict_prolong.jl:187-320): the plain operator, then the edge planes
static int edge_variant(int variant, int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, int64_t len, void* out, int mem,
                        void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "dim out of range");
    if (variant == 3) {
        int rc = gb_prolong_size(dims[dim], len, nullptr);
        if (rc) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return edge_variant_device(variant, dtype, ndim, (const i64*)dims, in, dim, scale, len, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= (d == dim) ? (variant == 3 ? len : gb_restrict_size(dims[d])) : dims[d]; }
    struct Ctx { int variant, dtype, ndim, dim; const i64* dims; double scale; i64 len; } ctx{variant, dtype, ndim, dim, (const i64*)dims, scale, len};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return edge_variant_device(x->variant, x->dtype, x->ndim, x->dims, i, x->dim, x->scale, x->len, o, st);
    }, &ctx);
}
int gb_restrictb(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out, int mem, void* stream) {
    return edge_variant(1, dtype, ndim, dims, in, dim, scale, 0, out, mem, stream);
}
int gb_restrict_extrap(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out, int mem, void* stream) {
    return edge_variant(2, dtype, ndim, dims, in, dim, scale, 0, out, mem, stream);
}
int gb_prolongb(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t len, void* out, int mem, void* stream) {
    return edge_variant(3, dtype, ndim, dims, in, dim, 1.0, len, out, mem, stream);
}

int gb_restrict_slab(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, int64_t global_len, int64_t in_first,
                     int64_t out_first, int64_t out_count, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "restrict: dim out of range");
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return restrict_slab_device(dtype, ndim, (const i64*)dims, in, dim, scale, global_len, in_first, out_first, out_count, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= (d == dim) ? out_count : dims[d]; }
    struct Ctx { int dtype, ndim, dim; const i64* dims; double scale; i64 L, f, o, c; } ctx{dtype, ndim, dim, (const i64*)dims, scale, global_len, in_first, out_first, out_count};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return restrict_slab_device(x->dtype, x->ndim, x->dims, i, x->dim, x->scale, x->L, x->f, x->o, x->c, o, st);
    }, &ctx);
}

int gb_prolong_slab(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t global_n, int64_t global_len, int64_t in_first,
                    int64_t out_first, int64_t out_count, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || ndim < 1 || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (dim < 0 || dim >= ndim) return fail(GB_ERR_ARG, "prolong: dim out of range");
    int rc = gb_prolong_size(global_n, global_len, nullptr);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return prolong_slab_device(dtype, ndim, (const i64*)dims, in, dim, global_n, global_len, in_first, out_first, out_count, out, s);
    i64 it = 1, ot = 1;
    for (int d = 0; d < ndim; d++) { it *= dims[d]; ot *= (d == dim) ? out_count : dims[d]; }
    struct Ctx { int dtype, ndim, dim; const i64* dims; i64 n, len, f, o, c; } ctx{dtype, ndim, dim, (const i64*)dims, global_n, global_len, in_first, out_first, out_count};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return prolong_slab_device(x->dtype, x->ndim, x->dims, i, x->dim, x->n, x->len, x->f, x->o, x->c, o, st);
    }, &ctx);
}

// fused 3-D operators on one slab of an array sharded along dim 3 (device memory: the multi-GPU path keeps slabs resident)
int gb_restrict3_slab(int dtype, const int64_t* dims, const void* in, double scale, int64_t global_len3, int64_t in_first, int64_t out_first,
                      int64_t out_count, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    if (mem != GB_DEVICE) return fail(GB_ERR_UNSUPPORTED, "gb_restrict3_slab takes device memory");
    pool_keep_memory();
    return restrict3_slab_device(dtype, (const i64*)dims, in, scale, global_len3, in_first, out_first, out_count, out, (cudaStream_t)stream);
}
int gb_prolong3_slab(int dtype, const int64_t* dims, const void* in, const int64_t* out_dims, int64_t global_n3, int64_t in_first, int64_t out_first,
                     int64_t out_count, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out || !out_dims) return fail(GB_ERR_ARG, "bad arguments");
    if (mem != GB_DEVICE) return fail(GB_ERR_UNSUPPORTED, "gb_prolong3_slab takes device memory");
    pool_keep_memory();
    return prolong3_slab_device(dtype, (const i64*)dims, in, (const i64*)out_dims, global_n3, in_first, out_first, out_count, out, (cudaStream_t)stream);
}

int gb_restrict3(int dtype, const int64_t* dims, const void* in, double scale, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    for (int d = 0; d < 3; d++)
        if (dims[d] < 1) return fail(GB_ERR_ARG, "bad dims");
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return restrict3_device(dtype, (const i64*)dims, in, scale, out, s);
    i64 it = dims[0] * dims[1] * dims[2];
    i64 ot = gb_restrict_size(dims[0]) * gb_restrict_size(dims[1]) * gb_restrict_size(dims[2]);
    struct Ctx { int dtype; const i64* dims; double scale; } ctx{dtype, (const i64*)dims, scale};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return restrict3_device(x->dtype, x->dims, i, x->scale, o, st);
    }, &ctx);
}

int gb_prolong3(int dtype, const int64_t* dims, const void* in, const int64_t* out_dims, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !out_dims || !in || !out) return fail(GB_ERR_ARG, "bad arguments");
    i64 it = 1, ot = 1;
    for (int d = 0; d < 3; d++) {
        if (dims[d] < 1) return fail(GB_ERR_ARG, "bad dims");
        i64 m = dims[d];
        if (out_dims[d] > dims[d]) {
            int rc = gb_prolong_size(dims[d], out_dims[d], nullptr);
            if (rc) return rc;
            m = out_dims[d];
        }
        it *= dims[d];
        ot *= m;
    }
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return prolong3_device(dtype, (const i64*)dims, in, (const i64*)out_dims, out, s);
    struct Ctx { int dtype; const i64* dims; const i64* od; } ctx{dtype, (const i64*)dims, (const i64*)out_dims};
    return mg_host(dtype, it, ot, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return prolong3_device(x->dtype, x->dims, i, x->od, o, st);
    }, &ctx);
}

// Several fused 3-D levels in one call: restrict(A, flags) with one all-true column per level / prolong(A, sizes) with one
// column of sizes per level (mapdim, src/utilities.jl:7-20).  The intermediate levels live in two stream-ordered scratch
// arrays and never leave the device (host arrays: one copy in, one copy out, instead of a round trip per level).
// dims_levels: 3 x (nlevels + 1) extents, column l = the array that enters level l.
static int levels_device(int kind, int dtype, const i64* dims_levels, int nlevels, const void* in, double scale, void* out, cudaStream_t s) {
    AsyncBuf t0(s), t1(s);
    AsyncBuf* t[2] = {&t0, &t1};
    size_t need[2] = {0, 0};
    for (int lev = 0; lev + 1 < nlevels; lev++) {
        const i64* d = dims_levels + 3 * (lev + 1);
        need[lev & 1] = std::max(need[lev & 1], (size_t)d[0] * (size_t)d[1] * (size_t)d[2] * esize(dtype));
    }
    for (int k = 0; k < 2; k++)
        if (need[k]) GB_CUDA(t[k]->alloc(need[k]));
    const void* cur = in;
    for (int lev = 0; lev < nlevels; lev++) {
        void* dst = lev + 1 == nlevels ? out : t[lev & 1]->p;
        const int rc = kind == 0 ? restrict3_device(dtype, dims_levels + 3 * lev, cur, scale, dst, s)
                                 : prolong3_device(dtype, dims_levels + 3 * lev, cur, dims_levels + 3 * (lev + 1), dst, s);
        if (rc) return rc;
        cur = dst;
    }
    return GB_OK;
}
static int levels_call(int kind, int dtype, const std::vector<i64>& dl, int nlevels, const void* in, double scale, void* out, int mem, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return levels_device(kind, dtype, dl.data(), nlevels, in, scale, out, s);
    const i64* d0 = dl.data();
    const i64* dn = dl.data() + 3 * nlevels;
    struct Ctx { int kind, dtype, nlevels; const i64* dl; double scale; } ctx{kind, dtype, nlevels, dl.data(), scale};
    return mg_host(dtype, d0[0] * d0[1] * d0[2], dn[0] * dn[1] * dn[2], in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return levels_device(x->kind, x->dtype, x->dl, x->nlevels, i, x->scale, o, st);
    }, &ctx);
}
int gb_restrict3_levels(int dtype, const int64_t* dims, int nlevels, const void* in, double scale, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out || nlevels < 1 || nlevels > 64) return fail(GB_ERR_ARG, "bad arguments");
    std::vector<i64> dl(3 * (size_t)(nlevels + 1));
    for (int d = 0; d < 3; d++) dl[d] = dims[d];
    for (int l = 0; l < nlevels; l++)
        for (int d = 0; d < 3; d++) {
            if (dl[3 * l + d] < 1) return fail(GB_ERR_ARG, "bad dims");
            dl[3 * (l + 1) + d] = gb_restrict_size(dl[3 * l + d]);
        }
    return levels_call(0, dtype, dl, nlevels, in, scale, out, mem, stream);
}
int gb_prolong3_levels(int dtype, const int64_t* dims, int nlevels, const int64_t* out_dims, const void* in, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !out_dims || !in || !out || nlevels < 1 || nlevels > 64) return fail(GB_ERR_ARG, "bad arguments");
    std::vector<i64> dl(3 * (size_t)(nlevels + 1));
    for (int d = 0; d < 3; d++) dl[d] = dims[d];
    for (int l = 0; l < nlevels; l++)
        for (int d = 0; d < 3; d++) {
            const i64 from = dl[3 * l + d], to = out_dims[3 * l + d];
            if (from < 1) return fail(GB_ERR_ARG, "bad dims");
            i64 m = from;
            if (to > from) {  // (a dimension whose target is not larger is kept, src/restrict_prolong.jl:94)
                int rc = gb_prolong_size(from, to, nullptr);
                if (rc) return rc;
                m = to;
            }
            dl[3 * (l + 1) + d] = m;
        }
    return levels_call(1, dtype, dl, nlevels, in, 1.0, out, mem, stream);
}

// restrict / prolong along dims 1 and 2 of a 2-D array, or of every plane of a 3-D array, in one pass (the marching
// kernels without their dim-3 stage); same bits as gb_restrict / gb_prolong applied to dim 0, then dim 1
int gb_restrict12(int dtype, int ndim, const int64_t* dims, const void* in, double scale, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out || ndim < 2 || ndim > 3) return fail(GB_ERR_ARG, "bad arguments");
    for (int d = 0; d < ndim; d++)
        if (dims[d] < 1) return fail(GB_ERR_ARG, "bad dims");
    const i64 nz = ndim == 3 ? dims[2] : 1;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return restrict12_device(dtype, dims[0], dims[1], nz, in, scale, out, s);
    struct Ctx { int dtype; i64 L1, L2, nz; double scale; } ctx{dtype, dims[0], dims[1], nz, scale};
    return mg_host(dtype, dims[0] * dims[1] * nz, gb_restrict_size(dims[0]) * gb_restrict_size(dims[1]) * nz, in, out, s,
                   [](void* c, const void* i, void* o, cudaStream_t st) {
                       Ctx* x = (Ctx*)c;
                       return restrict12_device(x->dtype, x->L1, x->L2, x->nz, i, x->scale, o, st);
                   }, &ctx);
}
int gb_prolong12(int dtype, int ndim, const int64_t* dims, const void* in, int64_t m1, int64_t m2, void* out, int mem, void* stream) {
    if (!ok_dtype(dtype) || !dims || !in || !out || ndim < 2 || ndim > 3) return fail(GB_ERR_ARG, "bad arguments");
    for (int d = 0; d < ndim; d++)
        if (dims[d] < 1) return fail(GB_ERR_ARG, "bad dims");
    int rc = gb_prolong_size(dims[0], m1, nullptr);
    if (!rc) rc = gb_prolong_size(dims[1], m2, nullptr);
    if (rc) return rc;
    const i64 nz = ndim == 3 ? dims[2] : 1;
    cudaStream_t s = (cudaStream_t)stream;
    pool_keep_memory();
    if (mem == GB_DEVICE) return prolong12_device(dtype, dims[0], dims[1], nz, in, m1, m2, out, s);
    struct Ctx { int dtype; i64 n1, n2, nz, m1, m2; } ctx{dtype, dims[0], dims[1], nz, m1, m2};
    return mg_host(dtype, dims[0] * dims[1] * nz, m1 * m2 * nz, in, out, s, [](void* c, const void* i, void* o, cudaStream_t st) {
        Ctx* x = (Ctx*)c;
        return prolong12_device(x->dtype, x->n1, x->n2, x->nz, i, x->m1, x->m2, o, st);
    }, &ctx);
}

}  // extern "C"
