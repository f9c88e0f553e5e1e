/*
 * stencil_b200.h -- C-ABI of the B200-native dispersion-objective library (libstencil_b200.so).
 *
 * Drop-in boundary for ONE hot path of Ablinne/optimize-stencil: the objective that
 * Optimize.optimize() minimises.  The reference has no FFI of its own (it is pure
 * Python/NumPy); the entry points below are what a ctypes binding placed inside the
 * reference's extended_stencil/dispersion.py would call instead of its NumPy grid passes.
 * Citations are file:line into the reference's extended_stencil/ package.
 *
 *   sb200_create            <- Dispersion2D.init2 / Dispersion3D.init2   dispersion.py:268-281, 339-357
 *                              (the caller computes the per-axis tables with NumPy exactly as
 *                               the reference does and hands them over; the library owns all
 *                               device memory and never materialises the dense k array)
 *   sb200_set_weight        <- Dispersion.w / weight.py:4-33             dispersion.py:40
 *   sb200_objective_batch   <- Dispersion.norm / .stencil_ok / .dt_ok    dispersion.py:235-255, 284-307,
 *                              for a BATCH of coefficient vectors          360-385, 129-154
 *                              (replaces the mp.Pool scan, optimize.py:169-190)
 *   sb200_export            <- Dispersion.omega / .sqrtarg / .sqrtres    dispersion.py:164-196, 309-320,
 *                                                                         387-399, 111-121
 *   sb200_group_velocity    <- Dispersion.omega_output / omega_interp    dispersion.py:199-232, 402-413
 *                              (numpy.gradient of omega, |vg|, vph, k on the device)
 *   sb200_objective_batch_ex  the same with options: strided x-plane lists (load-balanced k-slabs), SEGMENTED
 *                              batches (many independent small batches in one launch, each bit-identical to its
 *                              own stand-alone launch: the lock-stepped multi-start scan of optimize.py:147-205)
 *                              and the start-grid screening variant (optimize.py:118-134)
 *   sb200_group_*           <- mp.Pool(nproc).imap(...)                  optimize.py:169-190
 *                              (one host thread, one context per visible GPU, one stream per GPU: the batch is
 *                               sharded by candidates or by k-slabs across the devices of ONE process,
 *                               SURVEY.md 8b.1 "device list", 8e "process model")
 *   sb200_destroy, sb200_last_error
 *
 * Conventions
 *   - plain C, plain pointers and sizes; all floating point is IEEE binary64.
 *   - every call returns 0 on success, a negative SB200_E* code otherwise; the text is available
 *     from sb200_last_error().  No C++ exception crosses the ABI.  There is NO CPU fallback:
 *     without a CUDA device every compute entry point fails with SB200_ECUDA.
 *   - a context is not re-entrant: one context per host thread.  The asynchronous *_device entry points
 *     share the context's scratch buffers (partial-sum slots, tickets): issue them on ONE stream, or
 *     synchronise between calls that use different streams.  The host-buffer entry points run on the
 *     context's own stream and first wait (on the device) for the last *_device call issued through this
 *     context, so mixing the two kinds needs no host synchronisation.
 *   - coefficient vectors are in the reference's own field order (stencil.py:209, :275):
 *       dim 2:  dt, alphax, alphay, betaxy, betayx, deltax, deltay                     (7 doubles)
 *       dim 3:  dt, alphax, alphay, alphaz, betaxy, betaxz, betayx, betayz, betazx,
 *               betazy, deltax, deltay, deltaz                                         (13 doubles)
 *     omega = (2/dt) asin(dt sqrt(sqrtarg)) is even in dt bit for bit, so the device works with |dt|; dt = 0 gives
 *     S = NaN and nan > 0 (0/0 at every k-point, as in the reference).
 *   - the grid is C-ordered with the first axis (x) slowest, like the reference's arrays.
 *   - results are the UN-GATED grid reductions; the reference's gating (stencil_ok < 0, dt_ok < 0,
 *     the 1e6 sentinel, the ValueErrors) is applied by the caller from (S, Amin, Amax, nan),
 *     see SURVEY.md section 8(a) "Gate".
 *   - results are deterministic: no floating-point atomics, fixed reduction order for a given
 *     (grid, batch size, x-range).
 */
#ifndef STENCIL_B200_H
#define STENCIL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 2

/* status codes */
#define SB200_OK        0
#define SB200_EINVAL   (-1)  /* bad argument */
#define SB200_ECUDA    (-2)  /* CUDA runtime error (incl. "no device") */
#define SB200_ENOMEM   (-3)  /* host or device allocation failed */

/* weight kinds (sb200_set_weight) */
#define SB200_W_UNIT    0    /* w == 1 everywhere: the reference's default `equal` weight        */
#define SB200_W_PLANE   1    /* dim 3 only: w[ny*nz], independent of x (weight.py:25-27 shapes)   */
#define SB200_W_DENSE   2    /* w[nx*ny*nz] (dim 3) or w[nx*ny] (dim 2), C order                  */

/* what sb200_export writes */
#define SB200_X_OMEGA   0    /* (2/dt) * asin(dt * sqrt(sqrtarg))            dispersion.py:191 */
#define SB200_X_SQRTARG 1    /* the term under the square root               dispersion.py:319, 398 */
#define SB200_X_SQRTRES 2    /* sqrt(sqrtarg)                                dispersion.py:116 */

typedef struct sb200_ctx sb200_ctx;

/* Grid description.  Table a (a = 0..dim-1) has n[a] entries:
 *   cos_kappa[a][i] = cos(kappa_a[i])            dispersion.py:277-278, 351-353
 *   s2[a][i]        = 0.5*(1 - cos_kappa[a][i])  dispersion.py:279-280, 354-356
 *   kcomp[a][i]     = kappa_a[i] / (1, Y, Z)[a]  dispersion.py:272-273, 344-346
 * Y, Z are the cell aspect ratios (Z ignored for dim 2).  Y2, Z2 are the divisors (Y*dx)**2 and
 * (Z*dx)**2 of dispersion.py:319, 398 EXACTLY as the caller's host language evaluated them -- NumPy's
 * and libm's x**2 is not always the correctly rounded x*x, and sqrtarg must divide by the same bits.
 * Pass 0 to let the library use Y*Y and Z*Z.  `device` is the CUDA ordinal. */
typedef struct sb200_grid_desc {
    int32_t dim;
    int32_t n[3];
    const double* cos_kappa[3];
    const double* s2[3];
    const double* kcomp[3];
    double Y, Z;
    double Y2, Z2;
    int32_t device;
} sb200_grid_desc;

int sb200_abi_version(void);

/* Number of visible CUDA devices (0 and SB200_ECUDA when there is no usable driver). */
int sb200_device_count(int32_t* count);

int sb200_create(sb200_ctx** out, const sb200_grid_desc* desc);
void sb200_destroy(sb200_ctx* ctx);

/* Text of the last error on this context (ctx may be NULL: last error of a failed create). */
const char* sb200_last_error(const sb200_ctx* ctx);

/* Set the weight function.  `w` is a host pointer (ignored for SB200_W_UNIT) holding `count`
 * doubles; count must match the kind.  The data is copied to the device. */
int sb200_set_weight(sb200_ctx* ctx, int32_t kind, const double* w, int64_t count);

/* Batched objective.  coeffs: HOST pointer, nbatch rows of ncoef(dim) doubles, row-major.
 * Only x-planes [x_begin, x_end) of the first axis are visited (k-slab sharding; pass 0, n[0]
 * for the whole grid).  Outputs are HOST arrays of nbatch entries each:
 *   S[b]    = sum over the visited points of w * (omega - k)^2       (NaN if any omega is NaN)
 *   Amin[b] = min(sqrtarg U {+0}),  Amax[b] = max(sqrtarg U {+0})  over the visited points, bit-exact:
 *             every value that can be an extremum is evaluated in the reference's operation order without
 *             FMA contraction (sb200_export does so at every point).
 *             sqrtarg is exactly 0 at k = 0, so for any range containing plane 0 these ARE
 *             numpy.min / numpy.max of the reference's array (up to the sign of a zero).  All the
 *             reference needs (dispersion.py:292-302, 149) is "is the minimum negative, and if so
 *             what is it" and the maximum when it is non-negative.  Both are NaN if any sqrtarg is.
 *   nan[b]  = 0 when no visited omega is NaN, a positive number otherwise.  It is a FLAG, not the count of
 *             NaN points (SURVEY.md 8b.2 asked for a count; the reference only ever tests it against zero,
 *             dispersion.py:193): it is derived from the partial sums, which a NaN omega poisons.
 *             sb200_export returns the exact count.
 * Synchronous: on return the outputs are valid. */
int sb200_objective_batch(sb200_ctx* ctx, const double* coeffs, int64_t nbatch,
                          int32_t x_begin, int32_t x_end,
                          double* S, double* Amin, double* Amax, int64_t* nan);

/* Options of the *_ex entry points.  Zero-initialise, then set what is needed.
 *   planes    the visited x-planes are x_begin + i * x_stride for i in [0, n_planes); n_planes == 0 means
 *             "x_begin, x_begin + x_stride, ... while < n[0]" and x_stride == 0 means 1.  dim 2: whole grid only.
 *   seg_size  0: one batch.  > 0: nbatch must be a multiple of seg_size; the launch then consists of
 *             nbatch / seg_size independent SEGMENTS, and the results of every segment are bit-identical to those
 *             of sb200_objective_batch on that segment alone (same tiling, same candidates sharing a thread, same
 *             summation order) -- the property that lets many SLSQP searches share one launch per iterate
 *             without changing a single search.
 *   flags     SB200_F_SCAN: use the start-grid screening variant (a warp whose k-points all have s > 1 for all of
 *             its candidates skips the evaluation: omega is NaN there whatever the bits).  Results are identical
 *             to the plain variant; it is faster when most candidates violate the CFL limit and ~1.5 % slower
 *             otherwise.  SB200_F_AUTO_SCAN (host-buffer entry points only): decide per call from the
 *             candidates' corner values.
 *             SB200_F_FAST: evaluate asin with the shorter series (relative error 3.7e-15 instead of 2.3e-16; omega
 *             stays 25x inside the 1e-13 bar, the objective 500x inside 1e-12) -- 4 % more throughput, meant for
 *             large batches.  Without it the sums agree with NumPy's to a few 1e-16, which keeps an SLSQP search
 *             on the reference's trajectory for as long as floating point allows; the screening variant always
 *             uses the shorter series. */
#define SB200_F_SCAN       1
#define SB200_F_AUTO_SCAN  2
#define SB200_F_FAST       4
typedef struct sb200_batch_opts {
    int32_t x_begin, x_stride, n_planes;
    int32_t flags;
    int64_t seg_size;
} sb200_batch_opts;

int sb200_objective_batch_ex(sb200_ctx* ctx, const double* coeffs, int64_t nbatch, const sb200_batch_opts* opts,
                             double* S, double* Amin, double* Amax, int64_t* nan);

/* Same, with DEVICE pointers and an explicit stream: a cudaStream_t passed as void* and used
 * literally (NULL is CUDA's default stream -- what PyTorch's default stream is; use
 * sb200_get_stream() for the context's own stream).  d_coeffs: nbatch*ncoef doubles.  d_out:
 * 4*nbatch doubles laid out as [S | Amin | Amax | nan-count-as-double].  Asynchronous on `stream`. */
int sb200_objective_batch_device(sb200_ctx* ctx, const double* d_coeffs, int64_t nbatch,
                                 int32_t x_begin, int32_t x_end, double* d_out, void* stream);

int sb200_objective_batch_device_ex(sb200_ctx* ctx, const double* d_coeffs, int64_t nbatch,
                                    const sb200_batch_opts* opts, double* d_out, void* stream);

/* Fixed-order combine of k-slab partials on the device ("one all-gather of 4 doubles per rank
 * followed by a fixed-rank-order reduce"): d_parts[world][4][nbatch] (the gathered d_out blocks of `world`
 * shards) -> d_out[4][nbatch]; S and nan are summed shard 0 first, Amin / Amax by min / max, a NaN in any
 * shard's extrema poisons both.  Asynchronous on `stream`. */
int sb200_combine_slabs_device(sb200_ctx* ctx, const double* d_parts, int32_t world, int64_t nbatch,
                               double* d_out, void* stream);

/* Dense export for ONE coefficient vector over x-planes [x_begin, x_end): writes
 * (x_end-x_begin)*n[1]*n[2] doubles (C order) to the HOST buffer `out`; also returns the grid
 * extrema and the NaN count of the exported field (any of the three may be NULL). */
int sb200_export(sb200_ctx* ctx, const double* coeffs, int32_t what, int32_t x_begin, int32_t x_end,
                 double* out, double* Amin, double* Amax, int64_t* nan);

/* Same with a DEVICE output buffer and stream (used literally, see above); d_stats (may be NULL)
 * receives 3 doubles [Amin, Amax, nan-count].  coeffs is a HOST pointer (one vector). */
int sb200_export_device(sb200_ctx* ctx, const double* coeffs, int32_t what, int32_t x_begin, int32_t x_end,
                        double* d_out, double* d_stats, void* stream);

/* Group-velocity fields of ONE coefficient vector on the whole grid -- what Dispersion.omega_output and
 * Dispersion3D.omega_interp derive from omega (dispersion.py:208-220, 408-410):
 *   omega                                    prod(n) doubles
 *   vgc[a] = numpy.gradient(omega, *h, edge_order=2)[a]   dim arrays of prod(n), concatenated; h[a] = dks[a]
 *   vg     = sqrt(sum_a vgc[a]^2)            (the caller applies the reference's vg[:3,:3(,:3)] = 1 patch)
 *   vph    = omega / k                       (k = 0 at the origin gives NaN, as in NumPy; the caller sets it to 1)
 *   k      = sqrt(sum_a kcomp[a]^2)
 * in NumPy's own arithmetic (bit-identical to numpy.gradient on the exported omega).  Every output is a
 * HOST buffer or NULL.  Needs n[a] >= 3 on every axis (numpy.gradient's own requirement). */
int sb200_group_velocity(sb200_ctx* ctx, const double* coeffs, const double* h,
                         double* omega, double* vgc, double* vg, double* vph, double* k, int64_t* nan);

/* ---- device groups: the batch sharded over several GPUs of ONE process ---------------------------------------
 * Replaces the reference's mp.Pool scan (optimize.py:169-190) behind the same Python call: a group owns one
 * context (tables, weights, stream, staging) per device; one host thread enqueues the shards on all devices
 * before it waits for any of them.  Results do not depend on the number of devices in candidate mode (every
 * candidate's reduction is the single-device one when the shards keep the single-device tiling, which they do
 * for seg_size > 0; for one big batch the tiling follows the shard size and the sums agree to the last few
 * bits); in slab mode the per-device partials are combined on the host in device order (sum / min / max / sum,
 * NaN-poisoning), deterministically. */
#define SB200_SHARD_AUTO        0   /* candidates when the batch is large, slabs when the grid is, else one device */
#define SB200_SHARD_CANDIDATES  1   /* device g evaluates a contiguous block of candidates (whole segments)        */
#define SB200_SHARD_SLABS       2   /* device g evaluates all candidates on planes g, g+G, g+2G, ... of the list    */
#define SB200_SHARD_NONE        3   /* the first device only                                                       */
#define SB200_SHARD_SLABS_CONTIGUOUS 4 /* slabs as contiguous x-ranges [gN/G, (g+1)N/G) (SURVEY.md 8e wording)     */

typedef struct sb200_group sb200_group;

/* devices: CUDA ordinals (ndevices >= 1); NULL means "all visible devices". desc->device is ignored. */
int sb200_group_create(sb200_group** out, const sb200_grid_desc* desc, const int32_t* devices, int32_t ndevices);
void sb200_group_destroy(sb200_group* grp);
const char* sb200_group_last_error(const sb200_group* grp);   /* grp may be NULL: last failed create */
int sb200_group_size(const sb200_group* grp, int32_t* ndevices);
/* Borrow member i's context (owned by the group) for the single-device entry points (exports, streams). */
int sb200_group_member(sb200_group* grp, int32_t i, sb200_ctx** ctx);
int sb200_group_set_weight(sb200_group* grp, int32_t kind, const double* w, int64_t count);
/* As sb200_objective_batch_ex (opts may be NULL), sharded as `shard` says.  used (may be NULL) receives the
 * sharding that was applied and the number of devices that took part: used[0] = SB200_SHARD_*, used[1] = count. */
int sb200_group_objective_batch(sb200_group* grp, const double* coeffs, int64_t nbatch, int32_t shard,
                                const sb200_batch_opts* opts, double* S, double* Amin, double* Amax, int64_t* nan,
                                int32_t* used);

/* The context's own (non-blocking) stream as a cudaStream_t, for callers of the *_device variants. */
int sb200_get_stream(const sb200_ctx* ctx, void** stream);

/* Bookkeeping for benchmarks: kernels launched by this context so far, and the CUDA-event
 * duration (ms) of the most recent sb200_objective_batch (host-buffer variant). */
int sb200_launch_count(const sb200_ctx* ctx, int64_t* launches);
int sb200_last_batch_ms(const sb200_ctx* ctx, double* ms);

/* Tile configuration override for experiments (0 = automatic): candidates per thread (1,2,3,4,8),
 * y-tile, x-tile.  Affects only speed and the (deterministic) summation order. */
int sb200_set_tiling(sb200_ctx* ctx, int32_t cand_per_thread, int32_t tile_y, int32_t tile_x);

/* Register-resident DFMA throughput of `device` in TFLOP/s (2 flop per DFMA), measured with CUDA
 * events over about `seconds` of back-to-back launches, best of several (chains per thread, resident warps)
 * shapes.  The roofline denominator "of measured". */
int sb200_fp64_peak(int32_t device, double seconds, double* tflops);

/* End-to-end rate of the most recent dense export / group-velocity call of this context: bytes delivered to
 * the caller's host buffers and the wall-clock seconds from the first launch to the last byte (chunked,
 * double-buffered D2H through pinned staging). */
int sb200_last_export_stats(const sb200_ctx* ctx, double* bytes, double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* STENCIL_B200_H */
