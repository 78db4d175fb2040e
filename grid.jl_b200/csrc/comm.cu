// comm.cu -- multi-GPU behind the C ABI: communicators, coefficient replication,
// query sharding of host batches, and slab-sharded restrict / prolong with halo exchange.
//
// A gb_comm owns one or more LOCAL ranks (device, NCCL communicator, a communication stream, a
// work stream):
//   gb_comm_init_all   single process drives every GPU of the box (ncclCommInitAll): what a Julia host
//                      uses -- one ccall fans out over the devices;
//   gb_comm_init_rank  one process per GPU (torchrun / MPI style): the 128-byte NCCL id travels through the
//                      launcher's own channel;
//   gb_comm_init_loopback  `nranks` virtual ranks on the CURRENT device, transfers are device copies -- runs the
//                      whole sharded code path (plan, halos, interior/boundary split) on a one-GPU box (tests).
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process, e.g. PyTorch's, or the system
// one), so the library itself has no link-time dependency and loads on boxes without NCCL.
//
// Sharded restrict / prolong (the reference's restrict(A, [true,true,true]) / prolong(A, sizes),
// src/restrict_prolong.jl:181-185, 87-100): the array is split into slabs along dim 3.  Per call and local rank:
//   1. the comm stream (highest stream priority) exchanges the <= 2-3 halo planes per side with the neighbours --
//      grouped ncclSend / ncclRecv directly out of the slabs' plane ranges into a small halo buffer; enqueued FIRST
//      so that its kernels hold their SM slots before the interior kernel fills the machine;
//   2. meanwhile the INTERIOR output planes -- those whose stencil stays inside the rank's own planes -- are computed
//      on the work stream straight from the caller's slab (no staging copy);
//   3. the few BOUNDARY output planes run after the halo has landed, from a small assembled buffer
//      (halo planes + the adjoining own planes).
// The element code is that of gb_restrict3_slab / gb_prolong3_slab, so the assembled result is bit-identical
// to the unsharded one.
//
// PEER mode (one process, every rank local, peer access between the devices -- gb_comm_init_all on an NVLink box; loopback):
// there is no exchange at all.  The boundary kernel reads the halo planes IN PLACE from the neighbouring ranks' slabs through
// peer pointers (multigrid.cu: Sides / slab_plane) -- the transfer is the kernel's own loads over NVLink, a few planes of a
// tile at a time, overlapped with its arithmetic; no NCCL group, no halo buffer, no assembly copies (which is what the five
// small levels of a chain pay for: ~40 us of fixed cost per level).  Falls back to the exchange when a rank's stencil reaches
// past its immediate neighbours (slabs a plane or two thin) or a slab lives in a memory pool peers cannot map.
#include <dlfcn.h>
#include <nccl.h>  // types only; every entry point is resolved with dlsym
#include <cuda.h>  // types only (cuPointerGetAttribute is resolved at run time)

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct gb_grid;

namespace gb {

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    static bool ok = false;
    std::call_once(once, [] {
        const char* names[] = {getenv("GB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) return;
#define GB_SYM(name) *(void**)(&api.name) = dlsym(api.h, "nccl" #name)
        GB_SYM(GetVersion); GB_SYM(GetUniqueId); GB_SYM(CommInitRank); GB_SYM(CommInitAll); GB_SYM(CommDestroy);
        GB_SYM(GroupStart); GB_SYM(GroupEnd); GB_SYM(Send); GB_SYM(Recv); GB_SYM(Broadcast); GB_SYM(GetErrorString);
#undef GB_SYM
        ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv &&
             api.Broadcast && api.GetErrorString;
    });
    return ok ? &api : nullptr;
}

int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
    return fail(GB_ERR_CUDA, std::string("NCCL error: ") + (a && a->GetErrorString ? a->GetErrorString(r) : "?") + " in " + what);
}
#define GB_NCCL(api, expr)                                              \
    do {                                                                \
        ncclResult_t _r = (expr);                                       \
        if (_r != ncclSuccess) return nccl_fail(api, _r, #expr);        \
    } while (0)

inline i64 rsize1(i64 len) { return (len & 1) ? (len + 1) / 2 : len / 2 + 1; }
inline void shard(i64 n, int r, int w, i64* lo, i64* hi) {  // contiguous, balanced
    const i64 base = n / w, rem = n % w;
    *lo = r * base + std::min<i64>(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}
// input planes (inclusive range) that output planes [k0, k1) read (src/restrict_prolong.jl:117-136, 26-44)
inline void needs(int kind, i64 n_in, i64 n_out, i64 k0, i64 k1, i64* lo, i64* hi) {
    if (k1 <= k0) { *lo = 0; *hi = -1; return; }
    if (kind == 0) {  // restrict: n_in = L, n_out = rsize(L)
        *lo = k0 == 0 ? 0 : ((n_in & 1) ? 2 * k0 - 1 : 2 * k0 - 2);
        *hi = (k1 - 1 <= n_out - 2) ? 2 * (k1 - 1) + 1 : n_in - 1;
        if (*hi > n_in - 1) *hi = n_in - 1;
        if (*lo < 0) *lo = 0;
    } else {  // prolong: n_in = n, n_out = len
        *lo = k0 >> 1;
        const i64 kl = k1 - 1;
        const bool copy_only = (n_out & 1) && !(kl & 1);
        *hi = std::min(n_in - 1, (kl >> 1) + (copy_only ? 0 : 1));
    }
}

struct Transfer { int src, dst; i64 first, count; };
struct Plan {
    int kind, world;
    i64 n_in, n_out;
    std::vector<i64> own0, own1, out0, out1, need0, need1;  // need: inclusive, need1 < need0 when the rank has no outputs
    std::vector<Transfer> xfers;
    Plan(int kind_, i64 n_in_, i64 n_out_, int world_) : kind(kind_), world(world_), n_in(n_in_), n_out(n_out_) {
        own0.resize(world); own1.resize(world); out0.resize(world); out1.resize(world); need0.resize(world); need1.resize(world);
        for (int r = 0; r < world; r++) {
            shard(n_in, r, world, &own0[r], &own1[r]);
            shard(n_out, r, world, &out0[r], &out1[r]);
            needs(kind, n_in, n_out, out0[r], out1[r], &need0[r], &need1[r]);
        }
        for (int dst = 0; dst < world; dst++) {
            const i64 lo = need0[dst], hi = need1[dst];
            if (hi < lo) continue;
            for (int src = 0; src < world; src++) {
                if (src == dst) continue;
                const i64 a = std::max(lo, own0[src]), b = std::min(hi + 1, own1[src]);
                const i64 seg[2][2] = {{a, std::min(b, own0[dst])}, {std::max(a, own1[dst]), b}};  // only planes outside dst's own range travel
                for (auto& s : seg)
                    if (s[1] > s[0]) xfers.push_back({src, dst, s[0], s[1] - s[0]});
            }
        }
    }
    // planes held after the exchange: own range widened to cover need
    void ext(int r, i64* e0, i64* e1) const {
        *e0 = own0[r]; *e1 = own1[r];
        if (need1[r] >= need0[r]) { *e0 = std::min(*e0, need0[r]); *e1 = std::max(*e1, need1[r] + 1); }
    }
};

// Host threads for communicators that drive several devices from one process: the ~20 runtime calls per device and level of
// a sharded multigrid call are enqueued by one thread per device instead of one thread for all (which is what limits the
// small levels of a chain: their kernels take 20-30 us).  Workers sleep between calls; inside a call (begin .. end) they
// spin on a phase counter, so a phase costs a cache-line round trip, not a futex wake-up.
class DevicePool {
public:
    explicit DevicePool(int n) : n_(n) {
        for (int k = 1; k < n; k++) {
            try {
                th_.emplace_back([this, k] { loop(k); });
            } catch (const std::exception&) {
                break;  // fewer workers: run() covers the missing indices on the calling thread
            }
        }
    }
    ~DevicePool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void begin() {
        {
            std::lock_guard<std::mutex> lk(m_);
            base_phase_ = phase_.load(std::memory_order_relaxed);
            open_.store(true, std::memory_order_release);
        }
        cv_.notify_all();
    }
    void end() { open_.store(false, std::memory_order_release); }
    // fn(i) for every i in [0, n): worker k takes i = k, the caller i = 0 and the indices without a worker
    void run(const std::function<void(int)>& fn) {
        const int nw = (int)th_.size();
        job_ = &fn;
        done_.store(0, std::memory_order_relaxed);
        phase_.fetch_add(1, std::memory_order_release);
        fn(0);
        for (int i = nw + 1; i < n_; i++) fn(i);
        while (done_.load(std::memory_order_acquire) < nw) cpu_relax();
    }

private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void loop(int k) {
        for (;;) {
            int seen;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || open_.load(std::memory_order_acquire); });
                if (stop_) return;
                seen = base_phase_;
            }
            while (open_.load(std::memory_order_acquire)) {
                const int p = phase_.load(std::memory_order_acquire);
                if (p != seen) {
                    seen = p;
                    (*job_)(k);
                    done_.fetch_add(1, std::memory_order_release);
                } else {
                    cpu_relax();
                }
            }
        }
    }
    const int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
    int base_phase_ = 0;
    std::atomic<bool> open_{false};
    std::atomic<int> phase_{0}, done_{0};
    const std::function<void(int)>* job_ = nullptr;
};

struct LocalRank {
    int dev = 0, rank = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t cs = nullptr, ws = nullptr;      // communication stream, default work stream
    cudaEvent_t ev_in = nullptr, ev_halo = nullptr;  // inputs ready (work -> comm), halo landed (comm -> work)
    cudaEvent_t ev_t0 = nullptr, ev_int = nullptr, ev_t1 = nullptr, ev_h1 = nullptr;  // timing (created with timing enabled)
    cudaEvent_t ev_done = nullptr;  // peer mode: this rank's kernels have read their neighbours' slabs
    void* scratch[2] = {nullptr, nullptr};  // intermediate levels of the multi-level calls (sharded_levels)
    size_t scratch_cap[2] = {0, 0};
    cudaEvent_t ev_scratch = nullptr;  // end of the last multi-level call on this rank
    bool scratch_used = false;
    void* halo = nullptr;  // [lo planes | hi planes]
    size_t halo_cap = 0;
    double halo_bytes = 0;  // last call
    bool timed = false;
};

}  // namespace

}  // namespace gb

using namespace gb;

struct gb_comm {
    int nranks = 0, first_rank = 0;
    int loopback = 0;
    int peer_capable = 0, peer = 0;  // peer-pointer halo reads possible / in use (gb_comm_peer_halo)
    int last_peer = 0;               // the last sharded call ran in peer mode
    std::unique_ptr<gb::DevicePool> pool;  // one host thread per local rank (communicators with several local ranks)
    NcclApi* api = nullptr;
    std::vector<LocalRank> loc;
    std::mutex mu;
};

namespace {

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

int init_local(LocalRank& l) {
    DevGuard g(l.dev);
    // the communication stream outranks the work streams: the halo kernels' few CTAs are scheduled as soon as SM slots free up,
    // instead of queueing behind every remaining block of the interior kernel (measured: the halo of level 0 landed 0.2 ms
    // AFTER the interior kernel at 8 GPUs before this, profiles/README.md)
    int lo_pri = 0, hi_pri = 0;
    GB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    GB_CUDA(cudaStreamCreateWithPriority(&l.cs, cudaStreamNonBlocking, hi_pri));
    GB_CUDA(cudaStreamCreateWithFlags(&l.ws, cudaStreamNonBlocking));
    GB_CUDA(cudaEventCreateWithFlags(&l.ev_in, cudaEventDisableTiming));
    GB_CUDA(cudaEventCreateWithFlags(&l.ev_halo, cudaEventDisableTiming));
    GB_CUDA(cudaEventCreate(&l.ev_t0));
    GB_CUDA(cudaEventCreate(&l.ev_int));
    GB_CUDA(cudaEventCreate(&l.ev_t1));
    GB_CUDA(cudaEventCreate(&l.ev_h1));
    GB_CUDA(cudaEventCreateWithFlags(&l.ev_done, cudaEventDisableTiming));
    GB_CUDA(cudaEventCreateWithFlags(&l.ev_scratch, cudaEventDisableTiming));
    return GB_OK;
}
bool host_threads_default() {  // GB_COMM_THREADS=0: one host thread enqueues for every local rank
    static const bool v = [] { const char* e = getenv("GB_COMM_THREADS"); return !(e && *e == '0'); }();
    return v;
}
bool peer_default() {
    static const bool v = [] { const char* e = getenv("GB_PEER_HALO"); return !(e && *e == '0'); }();
    return v;
}
// memory from a stream-ordered pool is mapped on peers only if the pool was told so: such slabs take the exchange
bool peer_mappable(const void* p) {
    typedef CUresult (*Fn)(void*, CUpointer_attribute, CUdeviceptr);
    static Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (Fn)f;
    }();
    if (!fn || !p) return false;
    CUmemoryPool pool = nullptr;
    if (fn(&pool, CU_POINTER_ATTRIBUTE_MEMPOOL_HANDLE, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return false;
    return pool == nullptr;
}

size_t esz(int dtype) { return dtype == GB_F64 ? 8 : 4; }

int ensure_halo(LocalRank& l, size_t bytes) {
    if (bytes <= l.halo_cap) return GB_OK;
    DevGuard g(l.dev);
    if (l.halo) {
        // earlier calls may still read the old buffer on the work / comm streams
        cudaStreamSynchronize(l.cs);
        cudaDeviceSynchronize();
        cudaFree(l.halo);
        l.halo = nullptr;
        l.halo_cap = 0;
    }
    GB_CUDA(cudaMalloc(&l.halo, bytes));
    l.halo_cap = bytes;
    return GB_OK;
}

// One sharded multigrid level.  kind 0: restrict3 (in: fine slabs, L = global fine extents), kind 1: prolong3 (in: coarse
// slabs, n = global coarse extents, m = global target extents).
// fn(i) -> status for every local rank, one host thread per rank when the communicator has a pool; the first failure (by rank)
int par_for(gb_comm* c, int nl, const std::function<int(int)>& fn) {
    std::vector<int> rcs(nl, GB_OK);
    std::vector<std::string> msgs(nl);
    auto body = [&](int i) {
        rcs[i] = fn(i);
        if (rcs[i]) msgs[i] = gb_last_error();
    };
    if (c->pool && nl > 1) c->pool->run(body);
    else for (int i = 0; i < nl; i++) body(i);
    for (int i = 0; i < nl; i++)
        if (rcs[i]) return fail(rcs[i], msgs[i]);
    return GB_OK;
}
// the pool's workers spin only while a sharded call is in progress
struct PoolRegion {
    gb::DevicePool* p;
    explicit PoolRegion(gb_comm* c, bool nested) : p(nested ? nullptr : c->pool.get()) { if (p) p->begin(); }
    ~PoolRegion() { if (p) p->end(); }
};

// (the caller holds c->mu; sync_end: with library-owned work streams, wait for the results before returning)
int sharded_level_locked(gb_comm* c, int kind, int dtype, const i64* gin, const i64* gout, const void* const* in_slabs, double scale, void* const* out_slabs,
                         void* const* streams, bool sync_end, bool in_region = false) {
    if (!c || !gin || !gout || !in_slabs || !out_slabs) return fail(GB_ERR_ARG, "NULL argument");
    if (dtype != GB_F32 && dtype != GB_F64) return fail(GB_ERR_ARG, "bad dtype");
    PoolRegion region(c, in_region);
    const int W = c->nranks, nl = (int)c->loc.size();
    const i64 n_in = gin[2], n_out = gout[2];
    const Plan plan(kind, n_in, n_out, W);
    const size_t es = esz(dtype);
    const size_t in_plane = (size_t)gin[0] * (size_t)gin[1] * es, out_plane = (size_t)gout[0] * (size_t)gout[1] * es;
    NcclApi* api = c->api;
    std::vector<cudaStream_t> ws(nl);
    // ---- 1. inputs ready; which output planes are interior ------------------------------------------------------
    struct Piece { i64 k0, k1; bool interior; };
    std::vector<std::vector<Piece>> pieces(nl);
    const bool maybe_peer = c->peer != 0 && nl == W && W > 1;
    int rc1 = par_for(c, nl, [&](int i) -> int {
        LocalRank& l = c->loc[i];
        const int r = l.rank;
        DevGuard g(l.dev);
        ws[i] = streams ? (cudaStream_t)streams[i] : l.ws;
        i64 e0, e1;
        plan.ext(r, &e0, &e1);
        const i64 o0 = plan.own0[r], o1 = plan.own1[r], k0 = plan.out0[r], k1 = plan.out1[r];
        if (!maybe_peer) {  // (peer mode needs no halo buffer; a call that falls back allocates it below)
            int rc = ensure_halo(l, std::max<size_t>(1, (size_t)((o0 - e0) + (e1 - o1)) * in_plane));
            if (rc) return rc;
        }
        GB_CUDA(cudaEventRecord(l.ev_t0, ws[i]));
        GB_CUDA(cudaEventRecord(l.ev_in, ws[i]));
        l.timed = true;
        l.halo_bytes = 0;
        // interior outputs [ka, kb): needs within [o0, o1)
        i64 ka = k0, kb = k1;
        while (ka < k1) {
            i64 lo, hi;
            needs(kind, n_in, n_out, ka, ka + 1, &lo, &hi);
            if (lo >= o0 && hi < o1) break;
            ka++;
        }
        while (kb > ka) {
            i64 lo, hi;
            needs(kind, n_in, n_out, kb - 1, kb, &lo, &hi);
            if (lo >= o0 && hi < o1) break;
            kb--;
        }
        if (ka < kb) {
            // the low boundary outputs only lack low planes and the high ones only high planes?  Not when the slab is a few
            // planes thin -- the boundary pieces below assemble whatever their stencil reads, so any split is valid.
            if (k0 < ka) pieces[i].push_back({k0, ka, false});
            pieces[i].push_back({ka, kb, true});
            if (kb < k1) pieces[i].push_back({kb, k1, false});
        } else if (k0 < k1) {
            pieces[i].push_back({k0, k1, false});
        }
        return GB_OK;
    });
    if (rc1) return rc1;
    // ---- peer mode?  every rank local, non-empty, its boundary stencils within the immediate neighbours' slabs, slabs mappable
    bool use_peer = maybe_peer;
    for (int i = 0; i < nl && use_peer; i++) {
        const int r = c->loc[i].rank;
        const i64 o0 = plan.own0[r], o1 = plan.own1[r];
        if (o1 <= o0 || r != i) { use_peer = false; break; }
        const i64 zmin = i > 0 ? plan.own0[r - 1] : o0, zmax = i < nl - 1 ? plan.own1[r + 1] : o1;
        for (const Piece& p : pieces[i]) {
            if (p.interior) continue;
            i64 lo, hi;
            needs(kind, n_in, n_out, p.k0, p.k1, &lo, &hi);
            if (lo < zmin || hi >= zmax) use_peer = false;
        }
        if (!c->loopback && !peer_mappable(in_slabs[i])) use_peer = false;
    }
    c->last_peer = use_peer ? 1 : 0;
    if (maybe_peer && !use_peer) {  // fell back to the exchange: the halo buffers after all
        for (int i = 0; i < nl; i++) {
            LocalRank& l = c->loc[i];
            const int r = l.rank;
            DevGuard g(l.dev);
            i64 e0, e1;
            plan.ext(r, &e0, &e1);
            int rc = ensure_halo(l, std::max<size_t>(1, (size_t)((plan.own0[r] - e0) + (e1 - plan.own1[r])) * in_plane));
            if (rc) return rc;
        }
    }
    // ---- 2. halo exchange on the comm streams ------------------------------------------------------------------
    for (int i = 0; i < nl && !use_peer; i++) {  // a send reads the source rank's slab: every comm stream waits for every local input
        DevGuard g(c->loc[i].dev);
        for (int j = 0; j < nl; j++) GB_CUDA(cudaStreamWaitEvent(c->loc[i].cs, c->loc[j].ev_in, 0));
    }
    auto local_of = [&](int rank) { return (rank >= c->first_rank && rank < c->first_rank + nl) ? rank - c->first_rank : -1; };
    auto halo_ptr = [&](int i, i64 plane) -> char* {  // where plane `plane` of the halo of local rank i lives
        const int r = c->loc[i].rank;
        i64 e0, e1;
        plan.ext(r, &e0, &e1);
        const i64 o0 = plan.own0[r], o1 = plan.own1[r];
        if (plane < o0) return (char*)c->loc[i].halo + (size_t)(plane - e0) * in_plane;
        return (char*)c->loc[i].halo + (size_t)((o0 - e0) + (plane - o1)) * in_plane;
    };
    if (!plan.xfers.empty() && !use_peer) {
        if (!c->loopback) GB_NCCL(api, api->GroupStart());
        for (const Transfer& t : plan.xfers) {
            const int si = local_of(t.src), di = local_of(t.dst);
            const size_t bytes = (size_t)t.count * in_plane;
            if (c->loopback) {
                LocalRank& d = c->loc[di];
                const char* src = (const char*)in_slabs[si] + (size_t)(t.first - plan.own0[t.src]) * in_plane;
                GB_CUDA(cudaMemcpyAsync(halo_ptr(di, t.first), src, bytes, cudaMemcpyDeviceToDevice, d.cs));
                d.halo_bytes += (double)bytes;
                continue;
            }
            if (si >= 0) {
                LocalRank& s = c->loc[si];
                DevGuard g(s.dev);
                const char* src = (const char*)in_slabs[si] + (size_t)(t.first - plan.own0[t.src]) * in_plane;
                GB_NCCL(api, api->Send(src, bytes, ncclChar, t.dst, s.comm, s.cs));
            }
            if (di >= 0) {
                LocalRank& d = c->loc[di];
                DevGuard g(d.dev);
                GB_NCCL(api, api->Recv(halo_ptr(di, t.first), bytes, ncclChar, t.src, d.comm, d.cs));
                d.halo_bytes += (double)bytes;
            }
        }
        if (!c->loopback) GB_NCCL(api, api->GroupEnd());
    }
    // ---- peer mode: no exchange.  Per rank (one host thread each): the BOUNDARY planes on the priority stream, by a kernel
    // that reads the neighbours' planes in place through peer pointers; the interior planes on the work stream meanwhile; join.
    if (use_peer) {
        int rc2 = par_for(c, nl, [&](int i) -> int {
            LocalRank& l = c->loc[i];
            const int r = l.rank;
            DevGuard g(l.dev);
            const i64 o0 = plan.own0[r], o1 = plan.own1[r], k0 = plan.out0[r];
            const i64 zmin = i > 0 ? plan.own0[r - 1] : o0, zmax = i < nl - 1 ? plan.own1[r + 1] : o1;
            const i64 dims[3] = {gin[0], gin[1], o1 - o0};
            GB_CUDA(cudaStreamWaitEvent(l.cs, l.ev_in, 0));  // own slab and the neighbours' slabs are complete
            if (i > 0) GB_CUDA(cudaStreamWaitEvent(l.cs, c->loc[i - 1].ev_in, 0));
            if (i < nl - 1) GB_CUDA(cudaStreamWaitEvent(l.cs, c->loc[i + 1].ev_in, 0));
            for (const Piece& p : pieces[i]) {
                if (p.interior) continue;
                i64 lo, hi;
                needs(kind, n_in, n_out, p.k0, p.k1, &lo, &hi);
                l.halo_bytes += (double)((size_t)(std::max<i64>(0, o0 - lo) + std::max<i64>(0, hi + 1 - o1)) * in_plane);  // read over the link
                const void* plo = i > 0 ? in_slabs[i - 1] : nullptr;
                const void* phi = i < nl - 1 ? in_slabs[i + 1] : nullptr;
                const i64 lo_first = i > 0 ? plan.own0[r - 1] : 0, hi_first = i < nl - 1 ? plan.own0[r + 1] : 0;
                char* out = (char*)out_slabs[i] + (size_t)(p.k0 - k0) * out_plane;
                int rc;
                if (kind == 0) rc = restrict3_slab_peer_device(dtype, dims, in_slabs[i], scale, n_in, o0, plo, lo_first, phi, hi_first, zmin, zmax, p.k0, p.k1 - p.k0, out, l.cs);
                else rc = prolong3_slab_peer_device(dtype, dims, in_slabs[i], gout, n_in, o0, plo, lo_first, phi, hi_first, zmin, zmax, p.k0, p.k1 - p.k0, out, l.cs);
                if (rc) return rc;
            }
            GB_CUDA(cudaEventRecord(l.ev_halo, l.cs));
            GB_CUDA(cudaEventRecord(l.ev_h1, l.cs));
            for (const Piece& p : pieces[i]) {
                if (!p.interior) continue;
                char* out = (char*)out_slabs[i] + (size_t)(p.k0 - k0) * out_plane;
                int rc;
                if (kind == 0) rc = restrict3_slab_device(dtype, dims, in_slabs[i], scale, n_in, o0, p.k0, p.k1 - p.k0, out, ws[i]);
                else rc = prolong3_slab_device(dtype, dims, in_slabs[i], gout, n_in, o0, p.k0, p.k1 - p.k0, out, ws[i]);
                if (rc) return rc;
            }
            GB_CUDA(cudaEventRecord(l.ev_int, ws[i]));
            GB_CUDA(cudaStreamWaitEvent(ws[i], l.ev_halo, 0));  // join: results complete when both streams are
            GB_CUDA(cudaEventRecord(l.ev_t1, ws[i]));
            GB_CUDA(cudaEventRecord(l.ev_done, ws[i]));
            return GB_OK;
        });
        if (rc2) return rc2;
        rc2 = par_for(c, nl, [&](int i) -> int {  // a slab may change once the neighbours that read it are done
            DevGuard g(c->loc[i].dev);
            if (i > 0) GB_CUDA(cudaStreamWaitEvent(ws[i], c->loc[i - 1].ev_done, 0));
            if (i < nl - 1) GB_CUDA(cudaStreamWaitEvent(ws[i], c->loc[i + 1].ev_done, 0));
            return GB_OK;
        });
        if (rc2) return rc2;
    }
    for (int i = 0; i < nl && !use_peer; i++) {
        LocalRank& l = c->loc[i];
        DevGuard g(l.dev);
        GB_CUDA(cudaEventRecord(l.ev_halo, l.cs));
        GB_CUDA(cudaEventRecord(l.ev_h1, l.cs));
    }
    // ---- 2b. interior planes on the work streams, while the halo travels (the exchange was enqueued FIRST: its kernels
    // take their SM slots before the interior kernel fills the machine) ------------------------------------------------
    for (int i = 0; i < nl && !use_peer; i++) {
        LocalRank& l = c->loc[i];
        const int r = l.rank;
        DevGuard g(l.dev);
        const i64 o0 = plan.own0[r], o1 = plan.own1[r], k0 = plan.out0[r];
        for (const Piece& p : pieces[i]) {
            if (!p.interior) continue;
            const i64 dims[3] = {gin[0], gin[1], o1 - o0};
            char* out = (char*)out_slabs[i] + (size_t)(p.k0 - k0) * out_plane;
            int rc;
            if (kind == 0) rc = restrict3_slab_device(dtype, dims, in_slabs[i], scale, n_in, o0, p.k0, p.k1 - p.k0, out, ws[i]);
            else rc = prolong3_slab_device(dtype, dims, in_slabs[i], gout, n_in, o0, p.k0, p.k1 - p.k0, out, ws[i]);
            if (rc) return rc;
        }
        GB_CUDA(cudaEventRecord(l.ev_int, ws[i]));
    }
    // ---- 3. boundary planes after the halo has landed ------------------------------------------------------------
    for (int i = 0; i < nl && !use_peer; i++) {
        LocalRank& l = c->loc[i];
        const int r = l.rank;
        DevGuard g(l.dev);
        GB_CUDA(cudaStreamWaitEvent(ws[i], l.ev_halo, 0));
        const i64 o0 = plan.own0[r], o1 = plan.own1[r], k0 = plan.out0[r];
        for (const Piece& p : pieces[i]) {
            if (p.interior) continue;
            i64 lo, hi;
            needs(kind, n_in, n_out, p.k0, p.k1, &lo, &hi);
            const i64 np = hi - lo + 1;
            void* ext = nullptr;
            GB_CUDA(cudaMallocAsync(&ext, (size_t)np * in_plane, ws[i]));
            int rc = GB_OK;
            // runs of planes by source: halo (below own), own, halo (above own)
            const i64 seg[3][2] = {{lo, std::min(hi + 1, o0)}, {std::max(lo, o0), std::min(hi + 1, o1)}, {std::max(lo, o1), hi + 1}};
            for (int sidx = 0; sidx < 3 && rc == GB_OK; sidx++) {
                const i64 a = seg[sidx][0], b = seg[sidx][1];
                if (b <= a) continue;
                const char* src = sidx == 1 ? (const char*)in_slabs[i] + (size_t)(a - o0) * in_plane : halo_ptr(i, a);
                cudaError_t e = cudaMemcpyAsync((char*)ext + (size_t)(a - lo) * in_plane, src, (size_t)(b - a) * in_plane, cudaMemcpyDeviceToDevice, ws[i]);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync (halo assembly)");
            }
            if (rc == GB_OK) {
                const i64 dims[3] = {gin[0], gin[1], np};
                char* out = (char*)out_slabs[i] + (size_t)(p.k0 - k0) * out_plane;
                if (kind == 0) rc = restrict3_slab_device(dtype, dims, ext, scale, n_in, lo, p.k0, p.k1 - p.k0, out, ws[i]);
                else rc = prolong3_slab_device(dtype, dims, ext, gout, n_in, lo, p.k0, p.k1 - p.k0, out, ws[i]);
            }
            cudaFreeAsync(ext, ws[i]);
            if (rc) return rc;
        }
        GB_CUDA(cudaEventRecord(l.ev_t1, ws[i]));
    }
    if (!use_peer && c->loopback) {  // the copies out of a source slab ran on the DESTINATION's comm stream: the source's work stream
        for (const Transfer& t : plan.xfers) {  // (whatever the caller does to that slab next) waits for them
            const int si = local_of(t.src), di = local_of(t.dst);
            if (si < 0 || di < 0) continue;
            DevGuard g(c->loc[si].dev);
            GB_CUDA(cudaStreamWaitEvent(ws[si], c->loc[di].ev_halo, 0));
        }
    }
    if (!streams && sync_end) {  // library-owned work streams: the call returns with the results in place
        for (int i = 0; i < nl; i++) {
            DevGuard g(c->loc[i].dev);
            GB_CUDA(cudaStreamSynchronize(ws[i]));
        }
    }
    return GB_OK;
}
int sharded_level(gb_comm* c, int kind, int dtype, const i64* gin, const i64* gout, const void* const* in_slabs, double scale, void* const* out_slabs,
                  void* const* streams) {
    if (!c) return fail(GB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    return sharded_level_locked(c, kind, dtype, gin, gout, in_slabs, scale, out_slabs, streams, true);
}

// Several levels in one call -- the reference's restrict(A, flags) with one all-true column per level / prolong(A, sizes) with
// one column of sizes per level (mapdim, src/utilities.jl:7-20): the intermediate levels live in two scratch arrays per rank
// that belong to the communicator (plain cudaMalloc, so neighbours can map them in peer mode), nothing returns to the host
// language between levels, and only the last level lands in the caller's slabs.  dims_levels: 3 x (nlevels + 1) global
// extents, column l = the array that enters level l (column nlevels = the result).
int sharded_levels(gb_comm* c, int kind, int dtype, const i64* dims_levels, int nlevels, const void* const* in_slabs, double scale, void* const* out_slabs,
                   void* const* streams) {
    if (!c || !dims_levels || !in_slabs || !out_slabs) return fail(GB_ERR_ARG, "NULL argument");
    if (nlevels < 1) return fail(GB_ERR_ARG, "nlevels < 1");
    if (dtype != GB_F32 && dtype != GB_F64) return fail(GB_ERR_ARG, "bad dtype");
    std::lock_guard<std::mutex> lk(c->mu);
    const int W = c->nranks, nl = (int)c->loc.size();
    const size_t es = esz(dtype);
    // scratch for the intermediate levels: level l's output goes to scratch[l & 1]
    if (nlevels > 1) {
        for (int i = 0; i < nl; i++) {
            LocalRank& l = c->loc[i];
            size_t need[2] = {0, 0};
            for (int lev = 0; lev + 1 < nlevels; lev++) {
                const i64* d = dims_levels + 3 * (lev + 1);
                i64 a, b;
                shard(d[2], l.rank, W, &a, &b);
                need[lev & 1] = std::max(need[lev & 1], (size_t)d[0] * (size_t)d[1] * (size_t)(b - a) * es);
            }
            DevGuard g(l.dev);
            for (int k = 0; k < 2; k++) {
                if (need[k] <= l.scratch_cap[k]) continue;
                if (l.scratch[k]) {  // earlier calls (on any stream) may still use the old array
                    GB_CUDA(cudaDeviceSynchronize());
                    cudaFree(l.scratch[k]);
                    l.scratch[k] = nullptr;
                    l.scratch_cap[k] = 0;
                }
                GB_CUDA(cudaMalloc(&l.scratch[k], need[k]));
                l.scratch_cap[k] = need[k];
            }
            // the scratch arrays may still be read by the previous call's last level (another stream, a neighbour)
            cudaStream_t w = streams ? (cudaStream_t)streams[i] : l.ws;
            if (l.scratch_used) GB_CUDA(cudaStreamWaitEvent(w, l.ev_scratch, 0));
        }
    }
    std::vector<const void*> ins(nl);
    std::vector<void*> outs(nl);
    for (int i = 0; i < nl; i++) ins[i] = in_slabs[i];
    int all_peer = 1;
    PoolRegion region(c, false);  // the workers stay awake across the levels
    for (int lev = 0; lev < nlevels; lev++) {
        const bool last = lev + 1 == nlevels;
        for (int i = 0; i < nl; i++) outs[i] = last ? out_slabs[i] : c->loc[i].scratch[lev & 1];
        const i64* gin = dims_levels + 3 * lev;
        const i64* gout = dims_levels + 3 * (lev + 1);
        // (a rank that owns no plane of a level passes whatever pointer it has: nothing is read or written through it)
        int rc = sharded_level_locked(c, kind, dtype, gin, gout, ins.data(), scale, outs.data(), streams, last, true);
        if (rc) return rc;
        all_peer &= c->last_peer;
        for (int i = 0; i < nl; i++) ins[i] = outs[i];
    }
    c->last_peer = all_peer;
    if (nlevels > 1) {
        for (int i = 0; i < nl; i++) {
            LocalRank& l = c->loc[i];
            DevGuard g(l.dev);
            GB_CUDA(cudaEventRecord(l.ev_scratch, streams ? (cudaStream_t)streams[i] : l.ws));
            l.scratch_used = true;
        }
    }
    return GB_OK;
}

}  // namespace

extern "C" {

int gb_shard_range(int64_t n, int rank, int nranks, int64_t* lo, int64_t* hi) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) return fail(GB_ERR_ARG, "gb_shard_range: bad arguments");
    i64 a, b;
    shard(n, rank, nranks, &a, &b);
    *lo = a;
    *hi = b;
    return GB_OK;
}

int gb_slab_plan(int kind, int64_t n_in, int64_t n_out, int rank, int nranks, int64_t* own, int64_t* out, int64_t* need, int64_t* ext, int64_t* transfers,
                 int max_transfers, int* ntransfers) {
    if ((kind != 0 && kind != 1) || n_in < 1 || n_out < 1 || nranks < 1 || rank < 0 || rank >= nranks) return fail(GB_ERR_ARG, "gb_slab_plan: bad arguments");
    const Plan p(kind, n_in, n_out, nranks);
    if (own) { own[0] = p.own0[rank]; own[1] = p.own1[rank]; }
    if (out) { out[0] = p.out0[rank]; out[1] = p.out1[rank]; }
    if (need) { need[0] = p.need0[rank]; need[1] = p.need1[rank]; }
    if (ext) { i64 a, b; p.ext(rank, &a, &b); ext[0] = a; ext[1] = b; }
    if (ntransfers) *ntransfers = (int)p.xfers.size();
    if (transfers) {
        for (int i = 0; i < (int)p.xfers.size() && i < max_transfers; i++) {
            transfers[4 * i] = p.xfers[i].src; transfers[4 * i + 1] = p.xfers[i].dst;
            transfers[4 * i + 2] = p.xfers[i].first; transfers[4 * i + 3] = p.xfers[i].count;
        }
    }
    return GB_OK;
}

int gb_comm_unique_id(void* id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) return fail(GB_ERR_ARG, "gb_comm_unique_id: the buffer must hold 128 bytes");
    NcclApi* api = nccl_api();
    if (!api) return fail(GB_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded; set GB_NCCL_LIB");
    ncclUniqueId u;
    GB_NCCL(api, api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return GB_OK;
}

int gb_comm_init_rank(gb_comm** out, int nranks, int rank, const void* id, size_t bytes) {
    if (!out || !id || bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) return fail(GB_ERR_ARG, "gb_comm_init_rank: bad arguments");
    *out = nullptr;
    NcclApi* api = nccl_api();
    if (!api) return fail(GB_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded; set GB_NCCL_LIB");
    gb_comm* c = new gb_comm();
    c->nranks = nranks; c->first_rank = rank; c->api = api;
    c->loc.resize(1);
    LocalRank& l = c->loc[0];
    l.rank = rank;
    if (cudaGetDevice(&l.dev) != cudaSuccess) l.dev = 0;
    int rc = init_local(l);
    if (rc) { delete c; return rc; }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclResult_t r = api->CommInitRank(&l.comm, nranks, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(api, r, "ncclCommInitRank"); }
    *out = c;
    return GB_OK;
}

int gb_comm_init_all(gb_comm** out, int ndev, const int* dev_ids) {
    if (!out || ndev < 1) return fail(GB_ERR_ARG, "gb_comm_init_all: bad arguments");
    *out = nullptr;
    int have = 0;
    GB_CUDA(cudaGetDeviceCount(&have));
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; i++) {
        devs[i] = dev_ids ? dev_ids[i] : i;
        if (devs[i] < 0 || devs[i] >= have) return fail(GB_ERR_ARG, "gb_comm_init_all: device id out of range");
    }
    NcclApi* api = nccl_api();
    if (!api) return fail(GB_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded; set GB_NCCL_LIB");
    gb_comm* c = new gb_comm();
    c->nranks = ndev; c->first_rank = 0; c->api = api;
    c->loc.resize(ndev);
    std::vector<ncclComm_t> comms(ndev);
    ncclResult_t r = api->CommInitAll(comms.data(), ndev, devs.data());
    if (r != ncclSuccess) { delete c; return nccl_fail(api, r, "ncclCommInitAll"); }
    for (int i = 0; i < ndev; i++) {
        c->loc[i].dev = devs[i];
        c->loc[i].rank = i;
        c->loc[i].comm = comms[i];
        int rc = init_local(c->loc[i]);
        if (rc) return rc;
    }
    // peer access between every pair of devices (NVLink / NVSwitch boxes): the boundary kernels then read halo planes in place
    bool ok = ndev > 1;
    for (int i = 0; i < ndev && ok; i++) {
        DevGuard g(devs[i]);
        for (int j = 0; j < ndev && ok; j++) {
            if (i == j || devs[i] == devs[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) { ok = false; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
    }
    c->peer_capable = ok ? 1 : 0;
    c->peer = ok && peer_default() ? 1 : 0;
    if (ndev > 1 && host_threads_default()) c->pool.reset(new gb::DevicePool(ndev));
    *out = c;
    return GB_OK;
}

int gb_comm_init_loopback(gb_comm** out, int nranks) {
    if (!out || nranks < 1 || nranks > 64) return fail(GB_ERR_ARG, "gb_comm_init_loopback: bad arguments");
    gb_comm* c = new gb_comm();
    c->nranks = nranks; c->first_rank = 0; c->loopback = 1;
    c->loc.resize(nranks);
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < nranks; i++) {
        c->loc[i].dev = dev;
        c->loc[i].rank = i;
        int rc = init_local(c->loc[i]);
        if (rc) return rc;
    }
    c->peer_capable = 1;  // one device: every "peer" pointer is local
    c->peer = peer_default() ? 1 : 0;
    if (nranks > 1 && host_threads_default()) c->pool.reset(new gb::DevicePool(nranks));
    *out = c;
    return GB_OK;
}

// Peer-pointer halo reads (see the header of this file): enable = 1 / 0 switches them on (if the communicator can) / off,
// anything else leaves the setting; *enabled (may be NULL) receives the setting in force.
int gb_comm_peer_halo(gb_comm* c, int enable, int* enabled) {
    if (!c) return fail(GB_ERR_ARG, "comm is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    if (enable == 0) c->peer = 0;
    else if (enable == 1) c->peer = c->peer_capable;
    if (enabled) *enabled = c->peer;
    return GB_OK;
}

int gb_comm_last_path(gb_comm* c, int* peer) {
    if (!c || !peer) return fail(GB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *peer = c->last_peer;
    return GB_OK;
}

int gb_comm_size(const gb_comm* c, int* nranks, int* nlocal, int* first_rank) {
    if (!c) return fail(GB_ERR_ARG, "comm is NULL");
    if (nranks) *nranks = c->nranks;
    if (nlocal) *nlocal = (int)c->loc.size();
    if (first_rank) *first_rank = c->first_rank;
    return GB_OK;
}

int gb_comm_destroy(gb_comm* c) {
    if (!c) return GB_OK;
    for (LocalRank& l : c->loc) {
        DevGuard g(l.dev);
        if (l.cs) cudaStreamSynchronize(l.cs);
        if (l.ws) cudaStreamSynchronize(l.ws);
        if (l.comm && c->api) c->api->CommDestroy(l.comm);
        if (l.halo) cudaFree(l.halo);
        for (void* p : l.scratch)
            if (p) cudaFree(p);
        for (cudaEvent_t e : {l.ev_in, l.ev_halo, l.ev_t0, l.ev_int, l.ev_t1, l.ev_h1, l.ev_done, l.ev_scratch})
            if (e) cudaEventDestroy(e);
        if (l.cs) cudaStreamDestroy(l.cs);
        if (l.ws) cudaStreamDestroy(l.ws);
    }
    delete c;
    return GB_OK;
}

int gb_comm_stats(gb_comm* c, double* out, int n, int* nout) {
    if (!c || !out) return fail(GB_ERR_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    int k = 0;
    for (LocalRank& l : c->loc) {
        if (k + 4 > n) break;
        double v[4] = {0, 0, 0, l.halo_bytes};
        if (l.timed) {
            DevGuard g(l.dev);
            GB_CUDA(cudaEventSynchronize(l.ev_t1));
            GB_CUDA(cudaEventSynchronize(l.ev_h1));
            float total = 0, interior = 0, halo = 0;
            GB_CUDA(cudaEventElapsedTime(&total, l.ev_t0, l.ev_t1));
            GB_CUDA(cudaEventElapsedTime(&interior, l.ev_t0, l.ev_int));
            GB_CUDA(cudaEventElapsedTime(&halo, l.ev_t0, l.ev_h1));
            v[0] = total; v[1] = interior; v[2] = halo > interior ? halo - interior : 0.0;
        }
        for (int j = 0; j < 4; j++) out[k + j] = v[j];
        k += 4;
    }
    if (nout) *nout = k;
    return GB_OK;
}

int gb_restrict3_sharded(gb_comm* c, int dtype, const int64_t* global_dims, const void* const* in_slabs, double scale, void* const* out_slabs,
                         void* const* streams) {
    if (!global_dims) return fail(GB_ERR_ARG, "NULL argument");
    const i64 gout[3] = {rsize1(global_dims[0]), rsize1(global_dims[1]), rsize1(global_dims[2])};
    for (int d = 0; d < 3; d++)
        if (global_dims[d] < 2) return fail(GB_ERR_UNSUPPORTED, "gb_restrict3_sharded: extents below 2 (use gb_restrict_slab per dimension)");
    return sharded_level(c, 0, dtype, (const i64*)global_dims, gout, in_slabs, scale, out_slabs, streams);
}

int gb_prolong3_sharded(gb_comm* c, int dtype, const int64_t* global_dims, const void* const* in_slabs, const int64_t* out_dims, void* const* out_slabs,
                        void* const* streams) {
    if (!global_dims || !out_dims) return fail(GB_ERR_ARG, "NULL argument");
    for (int d = 0; d < 3; d++) {
        int rc = gb_prolong_size(global_dims[d], out_dims[d], nullptr);
        if (rc) return rc;
        if (out_dims[d] <= global_dims[d]) return fail(GB_ERR_UNSUPPORTED, "gb_prolong3_sharded: every dimension must grow (use gb_prolong_slab per dimension)");
    }
    return sharded_level(c, 1, dtype, (const i64*)global_dims, (const i64*)out_dims, in_slabs, 1.0, out_slabs, streams);
}

int gb_restrict3_sharded_levels(gb_comm* c, int dtype, const int64_t* global_dims, int nlevels, const void* const* in_slabs, double scale, void* const* out_slabs,
                                void* const* streams) {
    if (!global_dims || nlevels < 1 || nlevels > 64) return fail(GB_ERR_ARG, "gb_restrict3_sharded_levels: bad arguments");
    std::vector<i64> dims(3 * (size_t)(nlevels + 1));
    for (int d = 0; d < 3; d++) dims[d] = global_dims[d];
    for (int l = 0; l < nlevels; l++)
        for (int d = 0; d < 3; d++) {
            if (dims[3 * l + d] < 2) return fail(GB_ERR_UNSUPPORTED, "gb_restrict3_sharded_levels: extents below 2 (use gb_restrict_slab per dimension)");
            dims[3 * (l + 1) + d] = rsize1(dims[3 * l + d]);
        }
    return sharded_levels(c, 0, dtype, dims.data(), nlevels, in_slabs, scale, out_slabs, streams);
}

int gb_prolong3_sharded_levels(gb_comm* c, int dtype, const int64_t* global_dims, int nlevels, const int64_t* out_dims, const void* const* in_slabs,
                               void* const* out_slabs, void* const* streams) {
    if (!global_dims || !out_dims || nlevels < 1 || nlevels > 64) return fail(GB_ERR_ARG, "gb_prolong3_sharded_levels: bad arguments");
    std::vector<i64> dims(3 * (size_t)(nlevels + 1));
    for (int d = 0; d < 3; d++) dims[d] = global_dims[d];
    for (int l = 0; l < nlevels; l++)
        for (int d = 0; d < 3; d++) {
            const i64 from = dims[3 * l + d], to = out_dims[3 * l + d];
            int rc = gb_prolong_size(from, to, nullptr);
            if (rc) return rc;
            if (to <= from) return fail(GB_ERR_UNSUPPORTED, "gb_prolong3_sharded_levels: every dimension must grow at every level");
            dims[3 * (l + 1) + d] = to;
        }
    return sharded_levels(c, 1, dtype, dims.data(), nlevels, in_slabs, 1.0, out_slabs, streams);
}

// Replicate a grid's coefficients from the rank that holds it to every rank: one ncclBroadcast of the stored array.
// grids[i] belongs to local rank i: the source rank's entry must be a grid, the others receive new handles.
int gb_grid_replicate(gb_comm* c, int src_rank, gb_grid** grids, int dtype, int ndim, const int64_t* stored_dims, int bc, int order, double fill) {
    if (!c || !grids || !stored_dims || ndim < 1 || ndim > MAXG) return fail(GB_ERR_ARG, "gb_grid_replicate: bad arguments");
    if (src_rank < 0 || src_rank >= c->nranks) return fail(GB_ERR_ARG, "gb_grid_replicate: bad source rank");
    std::lock_guard<std::mutex> lk(c->mu);
    const int nl = (int)c->loc.size();
    size_t bytes = esz(dtype);
    for (int d = 0; d < ndim; d++) bytes *= (size_t)stored_dims[d];
    std::vector<void*> buf(nl, nullptr);
    int rc = GB_OK;
    auto cleanup = [&](int code) {
        for (int i = 0; i < nl; i++)
            if (buf[i]) { DevGuard g(c->loc[i].dev); cudaStreamSynchronize(c->loc[i].cs); cudaFree(buf[i]); }
        return code;
    };
    const int si = (src_rank >= c->first_rank && src_rank < c->first_rank + nl) ? src_rank - c->first_rank : -1;
    for (int i = 0; i < nl; i++) {
        DevGuard g(c->loc[i].dev);
        cudaError_t e = cudaMalloc(&buf[i], bytes);
        if (e != cudaSuccess) return cleanup(cuda_fail(e, "cudaMalloc (replicate)"));
    }
    if (si >= 0) {
        if (!grids[si]) return cleanup(fail(GB_ERR_ARG, "gb_grid_replicate: the source rank's grid is NULL"));
        DevGuard g(c->loc[si].dev);
        rc = gb_grid_coefs_copy(grids[si], buf[si], c->loc[si].cs);
        if (rc) return cleanup(rc);
    }
    if (c->loopback) {
        for (int i = 0; i < nl; i++) {
            if (i == si) continue;
            cudaStreamSynchronize(c->loc[si].cs);
            cudaError_t e = cudaMemcpyAsync(buf[i], buf[si], bytes, cudaMemcpyDeviceToDevice, c->loc[i].cs);
            if (e != cudaSuccess) return cleanup(cuda_fail(e, "cudaMemcpyAsync (replicate)"));
        }
    } else if (c->nranks > 1) {
        NcclApi* api = c->api;
        ncclResult_t r = api->GroupStart();
        for (int i = 0; i < nl && r == ncclSuccess; i++) {
            DevGuard g(c->loc[i].dev);
            r = api->Broadcast(buf[i], buf[i], bytes, ncclChar, src_rank, c->loc[i].comm, c->loc[i].cs);
        }
        ncclResult_t r2 = api->GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) return cleanup(nccl_fail(api, r != ncclSuccess ? r : r2, "ncclBroadcast"));
    }
    for (int i = 0; i < nl; i++) {
        if (i == si) continue;
        DevGuard g(c->loc[i].dev);
        gb_grid* ng = nullptr;
        rc = gb_grid_create_from_coefs(&ng, dtype, ndim, stored_dims, buf[i], GB_DEVICE, bc, order, fill, c->loc[i].cs);
        if (rc) return cleanup(rc);
        grids[i] = ng;
    }
    return cleanup(GB_OK);
}

// A HOST query batch split over the local devices (contiguous ranges): every device runs the
// pipelined host-buffer path of gb_eval on its range, concurrently.
int gb_eval_sharded(gb_comm* c, gb_grid* const* grids, int out_mask, int64_t nq, const void* x, int xdtype, const double* affine, void* val, void* grad,
                    void* hess, int flags, int64_t* first_invalid) {
    if (!c || !grids) return fail(GB_ERR_ARG, "NULL argument");
    if (nq < 0) return fail(GB_ERR_ARG, "nq < 0");
    if (first_invalid) *first_invalid = -1;
    const int nl = (int)c->loc.size();
    for (int i = 0; i < nl; i++)
        if (!grids[i]) return fail(GB_ERR_ARG, "gb_eval_sharded: a grid handle is NULL (gb_grid_replicate first)");
    int dtype = 0, ndim = 0;
    int64_t nchan = 1;
    int rc = gb_grid_info(grids[0], &dtype, &ndim, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    gb_grid_channels(grids[0], &nchan);
    const size_t xs = xdtype == GB_F64 ? 8 : 4, es = esz(dtype) * (size_t)nchan;
    // GB_OUT_PACKED: one record of pw elements per query at val
    const size_t pw = (out_mask & GB_OUT_PACKED) ? (size_t)(1 + ((out_mask & (GB_OUT_GRAD | GB_OUT_HESS)) ? ndim : 0) + ((out_mask & GB_OUT_HESS) ? ndim * ndim : 0)) : 1;
    std::vector<int> rcs(nl, GB_OK);
    std::vector<int64_t> bad(nl, -1);
    std::vector<std::string> msgs(nl);
    std::vector<std::thread> th;
    auto shard_eval = [&](int i) {
        {
            i64 q0, q1;
            shard(nq, i, nl, &q0, &q1);
            if (q1 <= q0) return;
            cudaSetDevice(c->loc[i].dev);
            const char* xp = (const char*)x + (size_t)q0 * ndim * xs;
            rcs[i] = gb_eval(grids[i], out_mask, q1 - q0, xp, xdtype, ndim, 1, affine, val ? (char*)val + (size_t)q0 * es * pw : nullptr,
                             grad ? (char*)grad + (size_t)q0 * ndim * es : nullptr, hess ? (char*)hess + (size_t)q0 * ndim * ndim * es : nullptr, GB_HOST, flags,
                             first_invalid ? &bad[i] : nullptr, nullptr);
            if (rcs[i]) msgs[i] = gb_last_error();
            if (bad[i] >= 0) bad[i] += q0;
        }
    };
    std::vector<int> inline_i;  // (a thread that cannot be created: that shard runs on the calling thread)
    for (int i = 1; i < nl; i++) {
        try {
            th.emplace_back(shard_eval, i);
        } catch (const std::exception&) {
            inline_i.push_back(i);
        }
    }
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    shard_eval(0);
    for (int i : inline_i) shard_eval(i);
    cudaSetDevice(prev_dev);
    for (auto& t : th) t.join();
    for (int i = 0; i < nl; i++) {
        if (first_invalid && bad[i] >= 0 && (*first_invalid < 0 || bad[i] < *first_invalid)) *first_invalid = bad[i];
    }
    for (int i = 0; i < nl; i++)
        if (rcs[i] && rcs[i] != GB_ERR_INVALID_LOCATION) return fail(rcs[i], msgs[i]);
    for (int i = 0; i < nl; i++)
        if (rcs[i]) return fail(rcs[i], msgs[i]);
    return GB_OK;
}

}  // extern "C"
