/* gridb200.h -- C ABI of libgridb200.so, the B200 (sm_100a) engine behind Grid.jl's batched
 * evaluation hot path.
 *
 * The reference (timholy/Grid.jl) is pure Julia and has no FFI boundary of its own; its
 * "operator API" is the set of exported generic functions (src/Grid.jl:28-70).  Each entry point
 * below names the reference function(s) it replaces (file:line under the reference tree); the
 * Julia `ccall` stubs that bind them are in INTEGRATION.md and grid.jl_b200/julia/GridB200.jl.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (GB_OK == 0) and never
 *     throws or aborts; gb_last_error() gives a thread-local message for the last failure.
 *   - arrays are dense, column-major (Julia layout, dim 1 contiguous), dims are int64.
 *   - `mem` says where EVERY data pointer of that call lives: GB_HOST (the library stages the
 *     copies itself; the call returns when the results are in the caller's buffers) or GB_DEVICE
 *     (pointers are device pointers valid on the current device; work is enqueued on `stream`,
 *     a cudaStream_t passed as void*, NULL = default stream, and the call does not synchronise
 *     unless it has to report a per-query error).
 *   - query batches: coordinate d of query q is read at x[q*x_qstride + d*x_dstride] (elements of
 *     `xdtype`), so a Julia N x nq Matrix is (x_qstride=N, x_dstride=1) and an nq x N one is (1, nq).
 *   - outputs: val[nq], grad[N*nq] (N contiguous per query), hess[N*N*nq] (full symmetric N x N
 *     per query, both triangles stored, src/interp.jl:223-227).
 *   - handles are immutable after creation: any number of host threads may evaluate one grid.
 */
#ifndef GRIDB200_H
#define GRIDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_VERSION 10100 /* 1.1.0 */

typedef struct gb_grid gb_grid; /* opaque: InterpGrid{T,N,BC,IT} (src/interp.jl:42-47) on the device */

enum { GB_F32 = 0, GB_F64 = 1 };
/* src/boundaryconditions.jl:5-12 */
enum { GB_BC_NIL = 0, GB_BC_NAN = 1, GB_BC_NA = 2, GB_BC_REFLECT = 3, GB_BC_PERIODIC = 4, GB_BC_NEAREST = 5, GB_BC_FILL = 6 };
/* src/interpflags.jl:4-7 (npoints = order+1, src/interp.jl:579-582) */
enum { GB_NEAREST = 0, GB_LINEAR = 1, GB_QUADRATIC = 2, GB_CUBIC = 3,
       GB_FORWARD = 4, GB_BACKWARD = 5 /* InterpForward / InterpBackward: gb_interp_irregular only */ };
enum { GB_OUT_VAL = 1, GB_OUT_GRAD = 2, GB_OUT_HESS = 4,
       /* with GB_OUT_PACKED the requested outputs of a query form ONE record at val: value, gradient (N), Hessian (N x N,
          both triangles) -- W = 1 [+ N [+ N*N]] contiguous elements per query, i.e. a Julia W x nq Matrix (W x nchan x nq
          for multi-valued grids); grad / hess are ignored.  A new batched layout (absent upstream): one full-sector write
          per query instead of two or three partial ones, and the tiled path skips its unpack pass. */
       GB_OUT_PACKED = 8 };
enum { GB_HOST = 0, GB_DEVICE = 1 };
enum {
    GB_OK = 0,
    GB_ERR_ARG = 1,              /* ErrorException: arity / shape (src/interp.jl:88-90,159-161,206-211) */
    GB_ERR_INVALID_LOCATION = 2, /* ErrorException("Invalid location"), BCnil only (src/interp.jl:402-404) */
    GB_ERR_UNSUPPORTED = 3,      /* MethodError in the reference: Cubic with reflect/periodic/nearest/fill,
                                    valgradhess on Nearest/Linear, ndim > 8 */
    GB_ERR_CUDA = 4,
    GB_ERR_DIM = 5,              /* DimensionMismatch (src/coord.jl:7) */
    GB_ERR_PROLONG_SIZE = 6      /* "Cannot prolong a dimension of length n to len" (src/restrict_prolong.jl:328-330) */
};
/* gb_eval flags */
enum {
    GB_EXACT = 0, /* default: the reference's operation order (weight products, sequential sum,
                     no FMA contraction) -> bit-identical to the CPU oracle */
    GB_FAST = 1,  /* separable dimension-by-dimension contraction with FMA; same taps and weights,
                     different rounding (validated to a few ulp of sum|w_i a_i|) */
    /* Scheduling only -- results are bit-identical either way.  By default the library bins large
       batches with heavy neighbourhoods (>= 256 bytes of taps per query, e.g. 3-D cubic f64) on big
       grids by grid tile and evaluates them out of shared-memory bricks ("tiled"); everything else
       gathers straight from global memory ("plain").  Callers whose queries are spatially coherent
       (lattices, image warps, points pre-sorted by cell) should pass GB_PLAIN: measured 1.6-2.3x
       faster than the default on such batches (profiles/README.md). */
    GB_TILED = 2, /* force the binned shared-memory path (compiled for 2-D / 3-D Quadratic and Cubic; ignored elsewhere) */
    GB_PLAIN = 4, /* force the plain path */
    /* Hint: the batch is spatially coherent (a lattice in raster order, points sorted by cell, tile or Morton code).  The
       tiled path probes a sample of every batch by itself -- bounding boxes of the first taps of 32 consecutive queries --
       and hands coherent batches to the plain kernel, which then beats the bins (its gathers hit L1 / L2); the hint
       only lowers the share of coherent leaves required (70 % instead of 88 %; GB_COHERENT_PERCENT overrides both,
       > 100 disables the probe; GB_TILED skips it).  Results are bit-identical in every case. */
    GB_SORTED = 8,
    /* gb_invert_opt / gb_grid_create_opt with GB_FAST: the prefilter runs as a line-parallel solver (one line per
       warp, parallel cyclic reduction of the interior recurrences) instead of the reference's sequential LU sweep;
       same solution to a few ulp of max|f| instead of bit for bit. */
};

int gb_version(void);
const char* gb_last_error(void);
int gb_device_count(int* n);
int gb_set_device(int device); /* device used by subsequent calls from the calling thread */
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t gb_launch_count(void);
/* Stage timing of the tiled evaluation (measurement aid for bench.py's roofline line; no reference
   counterpart).  gb_profile_enable(1) makes every tiled gb_eval* call record CUDA events on its
   launch stream around its stages; gb_profile_read waits for the last such call and returns the
   stage durations in milliseconds: ms[0] tile_hist (+ the coherence probe), [1] tile_offsets (3 small
   kernels), [2] tile_scatter, [3] eval_brick, [4] eval_deferred, [5] unpermute (+ the plain kernel, which does all the
   work of a coherent batch while the other stages exit at once, see GB_SORTED).  n = capacity of ms;
   *nstages = 6, or 0 when no tiled call ran since profiling was enabled.  With n >= 7, ms[6] = 1.0 when the call
   found the batch coherent and evaluated it in place on the plain path, else 0.0. */
int gb_profile_enable(int on);
int gb_profile_read(double* ms, int n, int* nstages);

/* ---- grid lifetime ------------------------------------------------------------------------ */
/* InterpGrid(A, BC, IT) / InterpGrid(A, fill::Number, IT)  (src/interp.jl:48-72): copies the
 * samples, pads by one element per side where the reference does (BCfill with `fill`; Cubic with
 * BCnil/nan/na with 0), runs interp_invert! on the device and keeps the coefficients resident.
 * bc == GB_BC_FILL selects the fill constructor (prefilter uses the BCnearest system, :68). */
int gb_grid_create(gb_grid** out, int dtype, int ndim, const int64_t* dims, const void* samples, int mem,
                   int bc, int order, double fill, void* stream);
/* Adopt an already padded + prefiltered coefficient array (a copy is made).  This is how a
 * replica is built on every other GPU after one rank has prefiltered and broadcast the
 * coefficients; `stored_dims` are the padded dims. */
int gb_grid_create_from_coefs(gb_grid** out, int dtype, int ndim, const int64_t* stored_dims, const void* coefs,
                              int mem, int bc, int order, double fill, void* stream);
/* size(G) (src/interp.jl:853-856) and the stored (padded) size; any out pointer may be NULL */
/* Multi-valued grids -- the reference's low-level interface (README.md:151-209; InterpGridCoefs /
   set_position / interp(ic, A, index), src/interp.jl:390-429,489-508): `nchan` arrays share one
   geometry, the channel is the slowest dimension of `samples` / `coefs` (element stride = product of
   the stored dims), taps and weights are computed once per query and reused for every channel.
   gb_eval* on such a handle write nchan values per query, channel fastest:
   val[q*nchan + c], grad[(q*nchan + c)*N + d], hess[(q*nchan + c)*N*N + ...].
   gb_grid_create_channels  = InterpGrid(A[..., c], BC, IT) for every c (padding + prefilter along the
                              interpolated dims only, i.e. interp_invert!(A, BC, IT, 1:N)).
   gb_grid_create_from_coefs_channels  adopts prefiltered coefficients; raw_coords != 0 gives the
                              low-level semantics: query coordinates index the stored array directly
                              (set_position's x), no +1 shift of setx.  N <= 4 (N = 4: nearest/linear). */
int gb_grid_create_channels(gb_grid** out, int dtype, int ndim, const int64_t* dims, int64_t nchan, const void* samples,
                            int mem, int bc, int order, double fill, void* stream);
int gb_grid_create_from_coefs_channels(gb_grid** out, int dtype, int ndim, const int64_t* stored_dims, int64_t nchan,
                                       const void* coefs, int mem, int bc, int order, double fill, int raw_coords,
                                       void* stream);
int gb_grid_channels(const gb_grid* g, int64_t* nchan);
int gb_grid_info(const gb_grid* g, int* dtype, int* ndim, int64_t* user_dims, int64_t* stored_dims, int* bc,
                 int* order, double* fill);
int gb_grid_coefs_device(const gb_grid* g, void** device_ptr); /* borrowed, read-only */
int gb_grid_coefs_download(const gb_grid* g, void* host_out);  /* G.coefs */
/* G.coefs in the reference layout into a caller-provided DEVICE buffer.  Works for every grid; gb_grid_coefs_device
   is refused (GB_ERR_UNSUPPORTED) when the resident layout differs from the reference's (multi-valued grids are kept
   channel-interleaved; 2-D to 4-D grids of >= 48 MB whose neighbourhoods are lighter than 256 B are kept in 4^N-element
   Morton blocks -- a pure addressing change, results are bit-identical; GB_BLOCK_MIN_MB overrides the threshold). */
int gb_grid_coefs_copy(const gb_grid* g, void* device_out, void* stream);
int gb_grid_destroy(gb_grid* g);

/* ---- prefilter ---------------------------------------------------------------------------- */
/* interp_invert!(A, BC, IT, dimlist) in place (src/interp.jl:909-953); dimlist is 0-based;
 * a no-op for Nearest/Linear (:909-912). */
int gb_invert(int dtype, int ndim, const int64_t* dims, void* data, int mem, int bc, int order,
              const int* dimlist, int ndimlist, void* stream);
/* The same with flags: GB_EXACT (= gb_invert: the reference's LU / Woodbury operation order, bit-identical to the
 * oracle) or GB_FAST (line-parallel solver: every line is solved by a whole warp -- the tridiagonal system's two
 * one-term recurrences are evaluated by warp-level scans, the boundary rows / Woodbury corners by a rank-2
 * correction -- so grids with few long lines, e.g. an 8192 x 8192 image, fill the machine; agrees with the exact
 * solve to a few ulp of max|f|, the tolerance north_star states). */
int gb_invert_opt(int dtype, int ndim, const int64_t* dims, void* data, int mem, int bc, int order,
                  const int* dimlist, int ndimlist, int flags, void* stream);
/* gb_grid_create with flags for the prefilter (GB_EXACT / GB_FAST as in gb_invert_opt). */
int gb_grid_create_opt(gb_grid** out, int dtype, int ndim, const int64_t* dims, const void* samples, int mem,
                       int bc, int order, double fill, int flags, void* stream);
/* filledges!(A, val) (src/utilities.jl:22-33): every element with an index equal to 1 or size(A, d) in some
 * dimension d is set to val, in place. */
int gb_filledges(int dtype, int ndim, const int64_t* dims, void* data, double val, int mem, void* stream);
/* pad1(A, f, 1:N) (src/interp.jl:1305-1310): out has dims+2 */
int gb_pad1(int dtype, int ndim, const int64_t* dims, const void* in, double fill, void* out, int mem, void* stream);

/* ---- evaluation --------------------------------------------------------------------------- */
/* Batched getindex / valgrad / valgradhess == a loop of the scalar calls
 * (src/interp.jl:142-156, 158-203, 205-270).  x holds USER coordinates (the +1 of setx for
 * BCfill / padded Cubic, :97-139, is applied inside, in the coordinate's own type).
 * affine: NULL, or ndim rows [kind, start, step, divisor] for CoordInterpGrid
 *   (kind 0 FloatRange, 1 StepRange, 2 UnitRange: coordlookup src/coord.jl:30-32; gradients are
 *   scaled by coordiscale :54-56).
 * val/grad/hess may be NULL when not in out_mask.  Invalid locations produce NaN in every
 * requested output; for GB_BC_NIL the call additionally returns GB_ERR_INVALID_LOCATION and the
 * index of the first offending query in *first_invalid (-1 if none; pass NULL to skip the check
 * and keep a GB_DEVICE call fully asynchronous).
 * mem == GB_HOST: the batch is streamed in chunks (copy-in, evaluation, copy-out overlapped).  Pinned buffers
 * (gb_host_alloc_pinned / gb_host_register) run at the PCIe rate; pageable buffers of 2^19 queries or more are staged by
 * the library through its own pinned slots with several host threads (environment: GB_HOST_THREADS, 0 = the driver's
 * pageable copies; GB_EVAL_CHUNK / GB_STAGE_CHUNK = queries per chunk; GB_EVAL_TRACE=1 prints the chunk timeline). */
int gb_eval(const gb_grid* g, int out_mask, int64_t nq, const void* x, int xdtype, int64_t x_qstride,
            int64_t x_dstride, const double* affine, void* val, void* grad, void* hess, int mem, int flags,
            int64_t* first_invalid, void* stream);
/* Same, coordinates as ndim separate vectors (a Julia tuple of Vectors): xs[d][q]. */
int gb_eval_soa(const gb_grid* g, int out_mask, int64_t nq, const void* const* xs, int xdtype,
                const double* affine, void* val, void* grad, void* hess, int mem, int flags,
                int64_t* first_invalid, void* stream);
/* Tensor-product getindex(G, xvec, yvec, ...) (src/interp.jl:273-306): out[i1,i2,...] =
 * G[vecs[0][i1], vecs[1][i2], ...], column-major, lens[d] = length of vecs[d]. */
int gb_eval_outer(const gb_grid* g, const int64_t* lens, const void* const* vecs, int xdtype,
                  const double* affine, void* out, int mem, int flags, void* stream);

/* ---- InterpIrregular (src/interp.jl:318-377) -----------------------------------------------------
   1-D interpolation on `n` sorted, non-uniform knots: out[q] = G[x[q]] for G = InterpIrregular(knots,
   samples, BC, IT).  order: GB_FORWARD, GB_BACKWARD, GB_NEAREST or GB_LINEAR; bc: GB_BC_NIL, GB_BC_NAN,
   GB_BC_NA, GB_BC_FILL (fill value `fill`; NaN for the BC-type constructor) or GB_BC_NEAREST
   (BCreflect / BCperiodic: GB_ERR_UNSUPPORTED like the reference's error, :335-337).  Knots, samples,
   queries and outputs share one element type.  Unsorted knots: GB_ERR_ARG ("Coordinates must be supplied
   in increasing order").  BCnil outside the knots: outputs NaN, GB_ERR_INVALID_LOCATION (the reference's
   BoundsError) with the first offending query in *first_invalid. */
int gb_interp_irregular(int dtype, int64_t n, const void* knots, const void* samples, int bc, int order, double fill,
                        int64_t nq, const void* x, void* out, int mem, int64_t* first_invalid, void* stream);

/* ---- restrict / prolong ------------------------------------------------------------------- */
int64_t gb_restrict_size(int64_t len);                           /* src/restrict_prolong.jl:334 */
int gb_prolong_size(int64_t n, int64_t len, int64_t* newlen);    /* :322-332, GB_ERR_PROLONG_SIZE on mismatch */
/* restrict(A, dim, scale) (:103-140), dim 0-based; out has dims with dims[dim] -> restrict_size */
int gb_restrict(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out,
                int mem, void* stream);
/* prolong(A, dim, len) (:10-48); out has dims with dims[dim] -> len */
int gb_prolong(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t len, void* out,
               int mem, void* stream);
/* "Balanced" and edge-extrapolating variants (src/restrict_prolong.jl:187-320): the plain operator
   followed by the reference's overwrite of the edge planes along `dim` (constants 1/sqrt(3), 2/sqrt(3),
   1/(2 sqrt(2)), 3/(2 sqrt(2)) formed as the reference forms them).  Same array conventions as
   gb_restrict / gb_prolong; the extent along `dim` must be >= 2. */
int gb_restrictb(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out, int mem,
                 void* stream);
int gb_restrict_extrap(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, void* out,
                       int mem, void* stream);
int gb_prolongb(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t len, void* out, int mem,
                void* stream);
/* Slab forms for a grid sharded along `dim`: `in` holds planes
 * [in_first, in_first + dims[dim]) of an array whose full extent along dim is global_len (restrict)
 * / global_n (prolong); the call writes output planes [out_first, out_first + out_count) of the
 * full result (out has extent out_count along dim).  Every input plane those outputs read must be
 * in the slab (own planes + the halo received from the neighbours), else GB_ERR_ARG.  Per-element
 * arithmetic is that of gb_restrict / gb_prolong, so the assembled slabs equal the unsharded result
 * bit for bit. */
int gb_restrict_slab(int dtype, int ndim, const int64_t* dims, const void* in, int dim, double scale, int64_t global_len,
                     int64_t in_first, int64_t out_first, int64_t out_count, void* out, int mem, void* stream);
int gb_prolong_slab(int dtype, int ndim, const int64_t* dims, const void* in, int dim, int64_t global_n, int64_t global_len,
                    int64_t in_first, int64_t out_first, int64_t out_count, void* out, int mem, void* stream);
/* The fused 3-D operators (below) on ONE SLAB of an array sharded along its last dimension: `in` holds planes
 * [in_first, in_first + dims[2]) of the global array (extent global_len3 / global_n3 along dim 3), `out` receives output
 * planes [out_first, out_first + out_count).  Device memory only.  The slab must hold every plane the requested outputs
 * read (GB_ERR_ARG otherwise); gb_prolong3_slab requires every dimension to grow (out_dims = the GLOBAL target). */
int gb_restrict3_slab(int dtype, const int64_t* dims, const void* in, double scale, int64_t global_len3, int64_t in_first,
                      int64_t out_first, int64_t out_count, void* out, int mem, void* stream);
int gb_prolong3_slab(int dtype, const int64_t* dims, const void* in, const int64_t* out_dims, int64_t global_n3,
                     int64_t in_first, int64_t out_first, int64_t out_count, void* out, int mem, void* stream);
/* restrict(A, [true,true,true], scale) on a 3-D array as ONE pass over the fine grid; the
 * per-dim intermediates are rounded exactly as the staged reference does, so the result is
 * bit-identical to gb_restrict applied to dims 0,1,2 in turn (src/utilities.jl:7-20). */
int gb_restrict3(int dtype, const int64_t* dims, const void* in, double scale, void* out, int mem, void* stream);
/* nlevels fused 3-D levels in one call (restrict(A, flags) with one all-true column per level / prolong(A, sizes) with one
 * column of sizes per level -- out_dims: 3 x nlevels): the intermediate levels stay on the device; `out` receives the last
 * level (extents: gb_restrict_size applied nlevels times / the last column of out_dims).  Same bits as nlevels calls. */
int gb_restrict3_levels(int dtype, const int64_t* dims, int nlevels, const void* in, double scale, void* out, int mem, void* stream);
int gb_prolong3_levels(int dtype, const int64_t* dims, int nlevels, const int64_t* out_dims, const void* in, void* out, int mem, void* stream);
/* prolong(A, [s1,s2,s3]) on a 3-D array as one pass (src/restrict_prolong.jl:87-100). */
int gb_prolong3(int dtype, const int64_t* dims, const void* in, const int64_t* out_dims, void* out, int mem,
                void* stream);

/* restrict / prolong along dims 1 and 2 in ONE pass: a 2-D array (ndim = 2: restrict(A, [true,true], scale),
 * prolong(A, [m1,m2]) -- image pyramids), or every plane of a 3-D array (ndim = 3: flags [true,true,false]).  Same bits
 * as the per-dimension calls.  Extents below 2 (restrict) or a dimension that does not grow (prolong): GB_ERR_UNSUPPORTED,
 * use the per-dimension calls. */
int gb_restrict12(int dtype, int ndim, const int64_t* dims, const void* in, double scale, void* out, int mem, void* stream);
int gb_prolong12(int dtype, int ndim, const int64_t* dims, const void* in, int64_t m1, int64_t m2, void* out, int mem,
                 void* stream);

/* ---- multi-GPU (SURVEY.md 8b "*_init(ndev, dev_ids) / *_shutdown", 8e) ---------------------------------------
 * The reference is single-process and single-threaded; these entry points have no reference counterpart beyond the
 * functions whose work they spread: getindex / valgrad / valgradhess batches (src/interp.jl:142-270) shard by query
 * range over replicated coefficients, restrict(A, [true,true,true]) / prolong(A, sizes)
 * (src/restrict_prolong.jl:181-185, 87-100) shard by slabs along dim 3 with one halo exchange per call.
 *
 * A communicator owns one or more LOCAL ranks (a device, an NCCL communicator, a communication stream):
 *   gb_comm_init_all       one process drives `ndev` GPUs (ncclCommInitAll; dev_ids NULL = 0..ndev-1): a Julia host
 *                          reaches the whole box with one ccall per operation;
 *   gb_comm_init_rank      one process per GPU: rank 0 calls gb_comm_unique_id and ships the 128 bytes through the
 *                          launcher's channel (torch.distributed, MPI, a file); every process passes its rank; the
 *                          calling thread's current device becomes the rank's device;
 *   gb_comm_init_loopback  `nranks` virtual ranks on the current device, halo transfers are device copies: the whole
 *                          sharded path on a one-GPU box (tests; no NCCL involved).
 * NCCL is bound at run time (libnccl.so.2 as already loaded in the process, else from the loader path; GB_NCCL_LIB
 * overrides); without it these calls return GB_ERR_UNSUPPORTED and everything else works.
 * Per-rank array arguments (`grids`, `in_slabs`, `out_slabs`, `streams`) have one entry per LOCAL rank, in rank order. */
typedef struct gb_comm gb_comm;
int gb_comm_unique_id(void* id, size_t bytes /* >= 128 */);
int gb_comm_init_rank(gb_comm** out, int nranks, int rank, const void* id, size_t bytes);
int gb_comm_init_all(gb_comm** out, int ndev, const int* dev_ids);
int gb_comm_init_loopback(gb_comm** out, int nranks);
int gb_comm_size(const gb_comm* c, int* nranks, int* nlocal, int* first_rank);
int gb_comm_destroy(gb_comm* c);
/* Host-only geometry (no GPU needed): rank r of n owns the contiguous, balanced range [lo, hi). */
int gb_shard_range(int64_t n, int rank, int nranks, int64_t* lo, int64_t* hi);
/* One level of slab-sharded restrict (kind 0: n_in = fine extent along dim 3, n_out = restrict_size(n_in)) or prolong
 * (kind 1: n_in coarse, n_out fine): own[2] = input planes rank holds, out[2] = output planes it produces, need[2] =
 * INCLUSIVE input range those outputs read (need[1] < need[0]: none), ext[2] = planes held after the exchange;
 * transfers = rows (src, dst, first_plane, count) of the WHOLE exchange (identical on every rank), at most max_transfers
 * rows written, *ntransfers = their number.  Any out pointer may be NULL. */
int gb_slab_plan(int kind, int64_t n_in, int64_t n_out, int rank, int nranks, int64_t* own, int64_t* out, int64_t* need,
                 int64_t* ext, int64_t* transfers, int max_transfers, int* ntransfers);
/* One ncclBroadcast of the stored coefficient array from `src_rank` to every rank.  grids[i]: the source rank's
 * entry is its grid (any resident layout); every other local entry receives a new handle (gb_grid_destroy it).
 * Every rank passes the same dtype / stored_dims / bc / order / fill (gb_grid_info of the source). */
int gb_grid_replicate(gb_comm* c, int src_rank, gb_grid** grids, int dtype, int ndim, const int64_t* stored_dims, int bc,
                      int order, double fill);
/* gb_eval on a HOST batch (N x nq layout: x_qstride = N, x_dstride = 1) split into contiguous ranges over the LOCAL
 * devices, each running the pipelined host-buffer path concurrently; *first_invalid is the smallest offending index. */
int gb_eval_sharded(gb_comm* c, gb_grid* const* grids, int out_mask, int64_t nq, const void* x, int xdtype,
                    const double* affine, void* val, void* grad, void* hess, int flags, int64_t* first_invalid);
/* restrict(A, [true,true,true], scale) / prolong(A, out_dims) of a 3-D array sharded by slabs along dim 3: local rank i
 * (global rank r) passes DEVICE planes gb_shard_range(global_dims[2], r) in in_slabs[i] and receives planes
 * gb_shard_range(out extent along dim 3, r) in out_slabs[i].  Interior planes are computed while the halo planes
 * travel (grouped ncclSend/ncclRecv on the communication stream); boundary planes follow.  Bit-identical to the
 * unsharded gb_restrict3 / gb_prolong3.  streams: NULL (library streams; the call returns when the results are in
 * place) or one cudaStream_t per local rank (work is enqueued; later work on that stream is ordered after it). */
int gb_restrict3_sharded(gb_comm* c, int dtype, const int64_t* global_dims, const void* const* in_slabs, double scale,
                         void* const* out_slabs, void* const* streams);
int gb_prolong3_sharded(gb_comm* c, int dtype, const int64_t* global_dims, const void* const* in_slabs,
                        const int64_t* out_dims, void* const* out_slabs, void* const* streams);
/* Timing of the last sharded restrict / prolong call, 4 doubles per local rank: total ms, interior-kernel ms, time the
 * halo landed after the interior kernel had finished (ms; 0 = fully hidden), halo bytes received.  Waits for the call. */
/* Several levels in one call: restrict(A, flags) with one all-true column per level / prolong(A, sizes) with one column of
 * sizes per level (mapdim, src/utilities.jl:7-20).  in_slabs hold the rank's planes of the array that enters the first level
 * (global_dims), out_slabs receive its planes of the LAST level's result; the levels in between live in scratch arrays
 * owned by the communicator and nothing returns to the host between levels.  gb_prolong3_sharded_levels: out_dims holds
 * 3 x nlevels target extents, column l = the sizes after level l.  Bit-identical to nlevels single-level calls. */
int gb_restrict3_sharded_levels(gb_comm* c, int dtype, const int64_t* global_dims, int nlevels, const void* const* in_slabs, double scale,
                                void* const* out_slabs, void* const* streams);
int gb_prolong3_sharded_levels(gb_comm* c, int dtype, const int64_t* global_dims, int nlevels, const int64_t* out_dims, const void* const* in_slabs,
                               void* const* out_slabs, void* const* streams);
int gb_comm_stats(gb_comm* c, double* out, int n, int* nout);
/* Peer mode of the sharded restrict / prolong: when one process drives every rank and the devices can map each other's
 * memory (gb_comm_init_all on an NVLink box; loopback), there is no halo exchange -- the boundary kernel reads the
 * neighbouring ranks' planes in place through peer pointers, on the priority stream, next to the interior kernel.  On by
 * default where possible (GB_PEER_HALO=0 in the environment: off); enable = 1 / 0 switches it, any other value only
 * queries; *enabled (may be NULL) receives the setting in force.  Calls whose stencils reach past the immediate neighbours,
 * or whose slabs live in a memory pool peers cannot map (cudaMallocAsync), take the exchange on their own. */
int gb_comm_peer_halo(gb_comm* c, int enable, int* enabled);
int gb_comm_last_path(gb_comm* c, int* peer); /* 1: the last sharded restrict / prolong ran in peer mode, 0: it exchanged halos */

/* ---- device / pinned memory helpers (device-resident query and output buffers) -------------- */
int gb_device_alloc(void** ptr, size_t bytes);
int gb_device_free(void* ptr);
int gb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int gb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int gb_host_alloc_pinned(void** ptr, size_t bytes);
int gb_host_free_pinned(void* ptr);
int gb_host_register(void* ptr, size_t bytes);
int gb_host_unregister(void* ptr);
int gb_stream_synchronize(void* stream);

/* ---- synthetic inputs ------------------------------------------------------ */
/* out[k] = (T)(lo + (hi-lo)*u(splitmix64(seed + first + k))), u from the top 53 bits; replayable
 * on the CPU for any index subset. */
int gb_synth_uniform(int dtype, void* device_out, int64_t n, uint64_t seed, double lo, double hi, int64_t first,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRIDB200_H */
