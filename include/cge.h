/*
 * cge.h -- C ABI of libcge.so: the B200-native (sm_100a) grid-estimation hot path.
 *
 * Drop-in boundary for the one data-parallel path of nablabits/cuda-grid-estimation
 * (reference paths below are relative to cpp/grid-algorithm/ in that repository):
 *
 *     src/grid.cu:79-127   computeDensitiesWrapper -> computeLikesWrapper ->
 *                          computePosteriorCuda -> computeExpectationsWrapper
 *
 * The reference has no FFI layer; its "API" is a set of header-defined C++ functions.
 * This header is what an FFI for that path binds: plain C, plain pointers and sizes, an int
 * status from every call, no C++ or torch types.  The C++ shim with the reference's own
 * function names on top of this ABI is include/cge_compat.h.
 *
 * Conventions
 *   - posterior layout is the reference's: posterior[i * grid_size + j] = P(mu_i, sigma_j)
 *     (inc/kernels.h:92; pinned by src/tests.cu:93-95).
 *   - ranges use the reference's float32 formula start + i*(end-start)/(size-1)
 *     (inc/kernels.h:63), bit for bit.
 *   - flat prior (src/grid.cu:97-101): the posterior is likes * (1/Z) (inc/utils.h:82-84).
 *   - every `*_dev` pointer is device (or managed) memory on the context's device; `stream` is
 *     a cudaStream_t passed as void*, with CUDA's own meaning of NULL (the legacy default
 *     stream).  Calls that take a stream only enqueue work unless stated otherwise.  The
 *     synchronous entry points run on a context-owned non-blocking stream: inputs produced
 *     on another stream must be complete before they are called.
 *   - there is NO CPU fallback: without a usable sm_100 device every call fails with
 *     CGE_ERR_NO_DEVICE / CGE_ERR_CUDA.
 */
#ifndef CGE_H
#define CGE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGE_VERSION 2

/* status codes (the reference returns void and prints; src/grid.cu:15-17 is its only check) */
enum {
  CGE_OK = 0,
  CGE_ERR_BAD_ARG = 1,    /* NULL pointer, size < 1, sigma range not > 0, shard out of range */
  CGE_ERR_NO_DEVICE = 2,  /* no CUDA device / not sm_100 */
  CGE_ERR_CUDA = 3,       /* a CUDA runtime call failed; see cge_last_error() */
  CGE_ERR_DEGENERATE = 4, /* normaliser Z is zero or not finite (all cells underflowed) */
  CGE_ERR_PEER = 5,       /* the exchange between GPUs cannot be set up (no peer access between two
                             devices of a cge_multi, or a CUDA IPC handle that cannot be opened) */
  CGE_ERR_SIZE = 6        /* stage API: densitiesSize != rows*cols (inc/wrappers.h:110-113) */
};

/* arithmetic mode of the likelihood kernel */
enum {
  CGE_MODE_PRODUCT_F32 = 0, /* running float32 product of per-observation pdf factors (2^t from
                               MUFU ex2, or derived from a neighbouring cell's on the FMA pipe),
                               one global per-factor rescale (posterior is scale-invariant) --
                               BASELINE.json "float32 product path" */
  CGE_MODE_LOGSPACE = 1     /* sum of log-pdf (float32x2 tile sums of centred terms folded into
                               float64), log-sum-exp normalisation -- for observation counts where
                               any product underflows */
};

/* element type of the posterior the caller receives (cge_params.posterior_dtype) */
enum {
  CGE_POSTERIOR_F32 = 0,  /* float32 -- BASELINE.json "float32 product path"; normalised in place:
                             12 bytes of HBM traffic per cell */
  CGE_POSTERIOR_F64 = 1   /* float64, the reference's type (thrust::device_vector<double>,
                             inc/cuda_functions.h:139-159, src/grid.cu:111-112): the likelihoods are
                             the same float32 values (kept in context scratch), widened and normalised
                             in float64 on the way out -- 16 bytes per cell instead of the 24 a
                             float64 unnormalised buffer would cost */
};

typedef struct cge_context cge_context; /* opaque: device, stream, scratch */
typedef struct cge_multi cge_multi;     /* opaque: one context per GPU of a single-process group */

typedef struct cge_params {
  int32_t n_obs;        /* observations (reference: rvsSize, src/grid.cu:34) */
  int32_t grid_size;    /* N: points per axis (reference: vecSize, src/grid.cu:65) */
  float mu_start;       /* reference: startMu..endSigma, src/grid.cu:67-70 */
  float mu_end;
  float sigma_start;
  float sigma_end;
  int32_t mode;         /* CGE_MODE_* */
  int32_t row_begin;    /* shard: this call owns mu rows [row_begin, row_end) of the global */
  int32_t row_end;      /* grid; 0,0 means the whole grid (0, grid_size) */
  int32_t variant;      /* CGE_VARIANT_*: 0 = library default; other values force one arithmetic
                           policy of the likelihood kernel (same results within tolerance) */
  int32_t posterior_dtype; /* CGE_POSTERIOR_*: what `posterior` points to */
} cge_params;

/* likelihood-kernel variants for cge_params.variant */
enum {
  CGE_VARIANT_DEFAULT = 0,   /* neighbour rows on fine grids (one exponential on MUFU.EX2 per 7 or
                                3 rows, the others derived from it on the FMA pipe: the tall /
                                wide 8-column policies for a whole grid, the 4-column shared-slope
                                one or a per-cell degree-2 / degree-3 polynomial per 128-column
                                strip), hybrid otherwise; chosen on the device from the ranges and
                                the observations */
  CGE_VARIANT_MUFU_ONLY = 1, /* scalar arithmetic, every 2^t on MUFU.EX2 (the MUFU roofline case) */
  CGE_VARIANT_PAIRED = 3,    /* neighbour rows with a per-cell polynomial, forced: packed f32x2
                                arithmetic; of three vertically adjacent cells the middle one's
                                factor is 2^t on MUFU.EX2, the two others are derived from it,
                                e * 2^u with 2^u - 1 a degree-2 FMA-pipe polynomial (only accurate
                                while the row spacing keeps |u| <= 0.011) */
  CGE_VARIANT_PAIRED3 = 8,   /* the same with the degree-3 polynomial, forced (|u| <= 0.05) */
  CGE_VARIANT_SLOPE = 9,     /* neighbour rows with the polynomial shared by a thread's four adjacent
                                columns, forced (only accurate on the finest grids) */
  CGE_VARIANT_SLOPE_WIDE = 10, /* the same with eight adjacent columns per thread (32-byte stores),
                                forced; the default kernel takes it for the whole grid when the
                                grid-level bounds allow it */
  CGE_VARIANT_SLOPE_TALL = 11, /* seven rows x eight adjacent columns per thread, one exponential
                                per seven rows, forced; the default on the finest grids */
  CGE_VARIANT_LOG_F64 = 5,   /* log-space mode only: float64 accumulation, one DFMA per pdf-eval (the
                                default log-space kernel sums centred terms in float32x2 tiles of 256
                                observations and folds the tiles into float64) */
  CGE_VARIANT_UNFUSED = 7,   /* the default kernels, but cge_grid_posterior always takes the
                                three-launch path (by default grids up to 256 points per axis run
                                as ONE launch: set-up, likelihood, fold, results and scale in a
                                single kernel) */
  CGE_VARIANT_HYBRID = 4     /* hybrid, forced: packed f32x2 arithmetic; of the 16 exponentials a
                                thread evaluates per observation, 12 use MUFU.EX2 and 4 an FMA-pipe
                                degree-5 polynomial with Cody-Waite range reduction */
};

/* per-phase device timings of the last synchronous call, milliseconds (CUDA events) */
typedef struct cge_timing {
  float setup_ms;       /* staging of > 16384 observations (0 otherwise: the set-up is part of the
                           likelihood kernel) */
  float likelihood_ms;  /* fused likelihood kernel (set-up, K1-K5) */
  float reduce_ms;      /* fold of the partial sums + publish */
  float scale_ms;       /* finish kernel: (peer sum,) marginals, expectations, posterior *= 1/Z */
  float total_ms;
  float h2d_ms;         /* host-buffer entry points only */
  float d2h_ms;
} cge_timing;

/* ---- context ------------------------------------------------------------------------- */
int cge_create(int device, cge_context **out);
int cge_destroy(cge_context *ctx);
const char *cge_last_error(void);       /* thread-local message of the last failing call */
int cge_version(void);
int cge_device_sm_count(cge_context *ctx, int *sm_count, int *cc_major, int *cc_minor);

/* ---- the fused hot path: replaces src/grid.cu:79-127 in one call ------------------------
 * cge_grid_posterior: host or device buffers (detected with cudaPointerGetAttributes), runs
 * synchronously, returns when every requested output is written.
 *   obs            n_obs float32, any order                       (reference: observations)
 *   posterior      (row_end-row_begin) * grid_size elements of p->posterior_dtype, or NULL to keep
 *                  it on device only (cge_posterior_dev)
 *   marg_mu        grid_size float64 (row sums; only [row_begin,row_end) are this shard's, the
 *                  rest are written as 0), or NULL               (inc/wrappers.h:169)
 *   marg_sigma     grid_size float64 (column sums of this shard), or NULL (inc/wrappers.h:170)
 *   expectations   2 float64 {E[mu], E[sigma]}, or NULL          (inc/wrappers.h:173-176)
 *   timing         optional
 * With a shard (row_begin/row_end not covering the grid) the normaliser is the SHARD's unless
 * the caller uses the phase API below and sums the partials over ranks, or cge_grid_posterior_multi.
 * Product mode takes up to 16384 observations (the running float32 product stays in range for any
 * input order: the kernel consumes the observations in a range-balanced order of its own); more
 * need CGE_MODE_LOGSPACE.
 */
int cge_grid_posterior(cge_context *ctx, const cge_params *p, const float *obs,
                       void *posterior, double *marg_mu, double *marg_sigma,
                       double *expectations, cge_timing *timing);

/* Asynchronous form of the fused call for device buffers: enqueues the whole step on `stream`
 * (one launch for grids up to 256 points per axis, three otherwise) and returns without waiting.
 * Results stay on the device (cge_results / cge_read_results, which waits for `stream`);
 * posterior_dev receives the normalised posterior.  Whole grid only (row_begin/row_end must cover
 * it): a shard needs the exchange between the two phases below. */
int cge_grid_posterior_async(cge_context *ctx, const cge_params *p, const float *obs_dev,
                             void *posterior_dev, void *stream);

/* ---- phase API (multi-GPU: one process per GPU, the caller owns the collective) ----------
 * phase 1: likelihood (scale statistics and range values derived inside the kernel) + local
 *          deterministic reduction.  Leaves the packed partial vector
 *          {Z, sum_i mu_i*rowsum_i, colsum[0..N)} (N+2 float64) in a context-owned device buffer
 *          (cge_partials) and the UNNORMALISED float32 likelihoods in posterior_dev (float32
 *          posterior) or in context scratch (float64 posterior).
 * ...      the caller sums the partial vector over ranks in place (ncclAllReduce / torch
 *          all_reduce, float64 sum) -- nothing else crosses NVLink.
 * phase 2: marginals, expectations and the scale by 1/Z into posterior_dev, one launch.
 *          Results stay on the device in the context's result block (cge_results).
 * A context runs ONE step at a time: both phases of a step, and consecutive steps, must be
 * enqueued on the same stream (or be ordered by the caller); the scratch and the "last block"
 * tickets are per context, not per stream.
 * Pointers returned by cge_partials / cge_results / cge_posterior_dev stay valid until the next
 * call on the context with a LARGER grid or shard (scratch is grow-only and reallocates then), a
 * cge_set_prior, or cge_destroy.
 */
/* Peer exchange (optional; one process per GPU on ONE node): replaces "the caller sums the partial
 * vector" by peer loads over NVLink fused into phase 2.  Each rank allocates an exchange block and
 * exports its CUDA IPC handle (cge_peer_local_handle), the caller gathers the handles of all ranks
 * (any transport) and hands the concatenation, in rank order, to cge_peer_connect.  From then on
 * cge_likelihood_partials publishes this rank's vector to the peers and cge_finalize waits for
 * theirs and adds all of them in rank order inside the finish kernel -- no library collective,
 * bit-identical totals on every rank.  Every rank must make the same sequence of phase calls.
 * Ranks must not destroy/disconnect while a peer may still be reading (barrier first).
 * (A single process driving several GPUs uses cge_create_multi instead: same kernels, plain peer
 * pointers.) */
#define CGE_IPC_HANDLE_BYTES 64
int cge_peer_local_handle(cge_context *ctx, void *handle /* CGE_IPC_HANDLE_BYTES out */,
                          int max_grid_size);
int cge_peer_connect(cge_context *ctx, int rank, int world,
                     const void *handles /* world * CGE_IPC_HANDLE_BYTES, rank order */);
int cge_peer_disconnect(cge_context *ctx);

int cge_likelihood_partials(cge_context *ctx, const cge_params *p, const float *obs_dev,
                            void *posterior_dev, void *stream);
/* Lifetime of the device pointers returned by cge_partials, cge_results and cge_posterior_dev:
 * they point into the context's grow-only scratch and stay valid until the next call on this
 * context with a LARGER grid, shard or observation count (which may re-allocate), cge_set_prior,
 * or cge_destroy.  Their CONTENTS are overwritten by the next step.  One step at a time per
 * context: the phase calls of two steps must not overlap on different streams (the context's
 * scratch, accumulators and tickets are shared). */
int cge_partials(cge_context *ctx, double **partials_dev, size_t *count);
int cge_finalize(cge_context *ctx, const cge_params *p, void *posterior_dev, void *stream);
/* result block on device: {E[mu], E[sigma], Z, 1/Z} then marg_sigma[N] then marg_mu[N]
 * (marg_mu holds this shard's rows at their global index, 0 elsewhere) */
int cge_results(cge_context *ctx, double **results_dev, size_t *count);
/* blocking convenience: copies the result block pieces to host pointers (any may be NULL) */
int cge_read_results(cge_context *ctx, const cge_params *p, double *expectations,
                     double *marg_mu, double *marg_sigma, double *z, void *stream);
/* device posterior the context keeps when cge_grid_posterior got posterior == NULL or a host pointer */
int cge_posterior_dev(cge_context *ctx, void **posterior_dev, size_t *count);
int cge_last_timing(cge_context *ctx, cge_timing *timing);
/* per-step CUDA-event capture around every phase of the next `max_steps` phase-API steps (events
 * are recorded on the stream the kernels are launched on); cge_profile_end synchronises and
 * returns the mean per-phase device times.  h2d_ms/d2h_ms are not filled.  (An event between two
 * kernels keeps the second from being set up while the first drains, so a profiled step is a few
 * microseconds longer than an unprofiled one.) */
int cge_profile_begin(cge_context *ctx, int max_steps);
int cge_profile_end(cge_context *ctx, cge_timing *mean, int *steps_captured);
/* number of kernels the last step launched */
int cge_launch_count(cge_context *ctx, int *launches);

/* ---- several GPUs from ONE process (the reference's caller is a single-process main(),
 * src/grid.cu:19-142) ------------------------------------------------------------------------
 * cge_create_multi: one context per listed device (devices == NULL: 0 .. n_gpus-1), peer access
 * enabled between every pair, exchange blocks cross-mapped.  The grid is row-sharded over the
 * devices in list order (balanced, contiguous mu rows); each step issues, from the calling thread,
 * likelihood + fold on every device and then the finish kernel on every device; the finish kernel
 * of each GPU reads the other GPUs' partial vectors over NVLink (plain peer pointers) and adds them
 * in device-list order.  A device may be listed more than once (the shards then share one stream
 * and run back to back: the exchange kernels can be exercised on a single-GPU box).
 *
 * cge_grid_posterior_multi: the fused call.  obs is host (or managed) memory.
 *   posterior_host   grid_size*grid_size elements of p->posterior_dtype gathered to host, or NULL
 *   the normalised shards always stay on their devices: cge_multi_shard returns them
 *   marg_mu, marg_sigma, expectations, timing: as cge_grid_posterior (whole grid)
 * p->row_begin / row_end must be 0 (the whole grid). */
int cge_create_multi(int n_gpus, const int *devices, cge_multi **out);
int cge_destroy_multi(cge_multi *m);
int cge_multi_size(cge_multi *m, int *n_gpus);
int cge_grid_posterior_multi(cge_multi *m, const cge_params *p, const float *obs,
                             void *posterior_host, double *marg_mu, double *marg_sigma,
                             double *expectations, cge_timing *timing);
/* shard `index` of the last cge_grid_posterior_multi: its device, row range and device pointer */
int cge_multi_shard(cge_multi *m, int index, int *device, int *row_begin, int *row_end,
                    void **posterior_dev);
/* the same step without host synchronisation or result copies, for timing loops: observations
 * already on every device (obs_dev[i] on device i) */
int cge_multi_step_async(cge_multi *m, const cge_params *p, const float *const *obs_dev);
int cge_multi_synchronize(cge_multi *m);

/* ---- stage API: the reference's L1 functions, one C entry each (device/managed pointers) --
 * These exist so the reference's stage-by-stage callers and tests keep working; they
 * materialise what the reference materialises and are meant for the reference's sizes.
 */
/* inc/cuda_functions.h:47-62 linspaceCuda */
int cge_linspace(cge_context *ctx, float *array_dev, int size, float start, float end);
/* inc/cuda_functions.h:65-92 create3dGridCuda */
int cge_meshgrid3(cge_context *ctx, const float *vec_x, const float *vec_y, const float *vec_z,
                  float *grid_x, float *grid_y, float *grid_z, int vec_xy_size, int vec_z_size);
/* inc/cuda_functions.h:95-115 computeDensitiesCuda: densities[f] = pdf(grid_z[f]; grid_x[f], grid_y[f]) */
int cge_densities(cge_context *ctx, float *densities, const float *grid_x, const float *grid_y,
                  const float *grid_z, size_t grid_size);
/* inc/wrappers.h:19-92 computeDensitiesWrapper minus its two linspace calls: densities straight
 * from the three vectors, densities[(i*N + j)*n + k] = pdf(obs[k]; mu[i], sigma[j]); the three
 * N*N*n meshgrid temporaries of the reference are never built */
int cge_densities_grid(cge_context *ctx, float *densities, const float *mu, const float *sigma,
                       const float *obs, int grid_size, int n_obs);
/* inc/cuda_functions.h:118-136 computeLikesCuda on the reference's row-pointer matrix (double**,
 * inc/utils.h:53-63): likes[r] = prod_k rows_ptr[r][k]; the matrix is left untouched */
int cge_row_products_rows(cge_context *ctx, double *likes, double *const *rows_ptr, int rows,
                          int cols);
/* inc/cuda_functions.h:161-182 marginalizeCuda on a row-pointer matrix, same axis meaning */
int cge_marginalize_rows(cge_context *ctx, double *marginal, double *const *rows_ptr, int rows,
                         int cols, int axis);
/* inc/wrappers.h:95-121 computeLikesWrapper: likes[r] = prod_k densities[r*cols+k], float64 */
int cge_row_products(cge_context *ctx, const float *densities, double *likes,
                     size_t densities_size, size_t likes_size);
/* inc/cuda_functions.h:139-159 computePosteriorCuda: posterior = likes * (1/sum(likes)) */
int cge_normalize(cge_context *ctx, const double *likes, double *posterior, size_t size);
/* inc/cuda_functions.h:161-182 marginalizeCuda on a dense row-major rows x cols matrix;
 * axis=1 -> row sums (mu marginal), axis=0 -> column sums (sigma marginal) */
int cge_marginalize(cge_context *ctx, double *marginal, const double *matrix, int rows,
                    int cols, int axis);
/* inc/cuda_functions.h:185-221 computeExpectationsCuda */
int cge_expectation(cge_context *ctx, const double *marginal, const float *vector, int size,
                    double *expectation_host);

/* ---- non-flat prior -------------------------------------------------
 * The reference only ever uses a flat prior and says a custom one "would have needed" the joint
 * product before normalisation (inc/cuda_functions.h:145-148; notebook cell 4: posterior =
 * likes * prior).  cge_set_prior installs, for the calls that follow on this context, separable
 * factors prior_mu[grid_size], prior_sigma[grid_size] (host or device; copied) and/or a full
 * device table laid out like the posterior shard ([rows][grid_size]; not copied, must stay
 * alive).  posterior = likes * prior / sum(likes * prior).  All-NULL restores the flat prior.
 * Prior values should be O(1): they multiply a float32 likelihood rescaled to peak at 2^40. */
int cge_set_prior(cge_context *ctx, const float *prior_mu, const float *prior_sigma, int grid_size,
                  const float *prior_full_dev);

/* ---- around the path (SURVEY section 8(f) "next" rows) ---------------------------------------
 * cge_generate_normal: the reference's observation generator -- cuRAND XORWOW, curand_init(seed,
 * i, 0) per observation i, out[i] = fma(sigma, curand_normal, mu) -- inc/kernels.h:7-49 with the
 * launch shape of inc/cuda_functions.h:13-44 (seed 42 there).  flags & 1: draw from an all-zero
 * generator state instead, which is what src/tests.cu:21-46 exercises (expects 0.00459315).
 * Input synthesis, not part of the hot path; provided so the reference binary's own observations
 * (obs[0] = 18.5689, src/grid.cu:41-47) can be reproduced bit for bit. */
int cge_generate_normal(cge_context *ctx, float *out_dev, int n, float mu, float sigma,
                        unsigned long long seed, int flags);
/* cge_quantiles: for each probability p, the first index whose cumulative marginal mass exceeds p
 * (fixed-order device prefix sum) -- the credible-interval step the notebook suggests
 * (notebooks/quick-prototyping.ipynb cell 5).  cdf_dev (n float64) is optional. */
int cge_quantiles(cge_context *ctx, const double *marginal_dev, int n, const double *probs_host,
                  int nprobs, int *indices_host, double *cdf_dev);

/* Scale statistics the likelihood kernel derived for the last call (diagnostics; waits for
 * `stream`): out[0..11] = log2 of the per-factor rescale, log-space shift, log2 of the largest
 * likelihood on the grid, sample mean, sum of squared deviations, min_i sum_k (x_k - mu_i)^2,
 * log-space centre mu_c, sum_k (x_k - mu_c)^2, umax (largest exponent step between vertically
 * adjacent cells, over the whole grid), the bounds on umax below which the paired-rows product
 * kernel runs with its degree-2 and degree-3 polynomial, ustep (umax per unit |a_j|: a column
 * band's own bound is max |a_j| over the band times ustep -- the kernel chooses per band),
 * out[12..14] = the grid's relative sigma step |dsigma| / min |sigma|, the largest shared-slope
 * error term the slope policies accept (2^-24), and 1 / 2 when the step ran the 8-column wide /
 * tall policy. */
int cge_scale_info(cge_context *ctx, double *out, int count, void *stream);

/* ---- roofline micro-benchmarks (bench.py): measured pipe peaks on this device ------------
 * kind: 0 = MUFU ex2.approx.ftz.f32, 1 = FFMA, 2 = DFMA, 3 = FFMA2 (fma.rn.f32x2);
 * 4-7 = instruction mixes per loop iteration: 4 FFMA2 + 4 FFMA, 4 FFMA2 + 8 FFMA,
 * 6 FFMA2 + 2 MUFU, 4 FFMA2 + 4 FFMA + 2 MUFU (how packed, scalar and MUFU work overlap);
 * 8-10 = the 6 FFMA2 + 2 MUFU mix at 8 / 12 / 16 resident warps per SM instead of 64.
 * Returns operations per second (FMA lane-ops; a packed FFMA2 counts as two; for the mixes the
 * MUFU ops are not counted). */
int cge_microbench(cge_context *ctx, int kind, double *ops_per_second);

#ifdef __cplusplus
}
#endif
#endif /* CGE_H */
