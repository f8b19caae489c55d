// cge_api.cu -- the C ABI of libcge.so (include/cge.h) on top of the sm_100a kernels.
// Host-side runtime: context (device, stream, grow-only scratch), launch planning, the fused
// and phase entry points, the stage API, and the pipe-peak micro-benchmarks.
// There is no CPU fallback anywhere in this file: every entry point needs the device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "cge_host.cuh"
#include "cge_kernels_misc.cuh"

#define CGE_EXPORT extern "C" __attribute__((visibility("default")))

namespace cge_host {

thread_local std::string g_last_error;

int set_error(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

}  // namespace cge_host

using cge_host::Plan;
using cge_host::set_error;

namespace {

// grow-only device buffer
template <class T>
struct DevBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t count) {
    if (count <= cap) return CGE_OK;
    if (ptr) {
      CK(cudaFree(ptr));
      ptr = nullptr;
      cap = 0;
    }
    size_t want = count + count / 8 + 64;
    CK(cudaMalloc(&ptr, want * sizeof(T)));
    cap = want;
    return CGE_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// grow-only pinned host buffer (staging for small host<->device copies: a copy from or to
// pageable memory costs ~10 us of driver staging each, a pinned one is a plain DMA)
template <class T>
struct PinnedBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t count) {
    if (count <= cap) return CGE_OK;
    if (ptr) {
      CK(cudaFreeHost(ptr));
      ptr = nullptr;
      cap = 0;
    }
    size_t want = count + count / 8 + 64;
    CK(cudaMallocHost(&ptr, want * sizeof(T)));
    cap = want;
    return CGE_OK;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

}  // namespace

struct cge_context {
  int device = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10] = {};
  DevBuf<float> mu, sigma, obs_pad, obs_in, posterior;
  DevBuf<double> ad, cd0, rowpart, colpart, partials, rowsum, zpart, smupart, espart, results, tmp;
  DevBuf<unsigned int> tickets;  // [0] fold, [1] results; zero between launches
  DevBuf<float> prior_mu, prior_sigma;   // owned copies of the separable prior factors
  bool has_prior_mu = false, has_prior_sigma = false;
  int prior_n = 0;
  const float *prior_full = nullptr;     // caller-owned device table, shard layout
  DevBuf<cge::ScaleInfo> scale;
  DevBuf<float> sink;
  PinnedBuf<float> host_obs;       // staging for host observations
  PinnedBuf<double> host_results;  // staging for {E, Z, 1/Z}, both marginals
  bool host_results_valid = false;  // the last step's kernel wrote host_results itself
  Plan plan;
  cge_timing timing = {};
  int launches = 0;
  // per-step event capture: 7 marks per step
  //   0 before tables | 1 after setup | 2 after likelihood | 3 after fold |
  //   4 before results | 5 after results | 6 after scale
  // peer exchange (CUDA IPC): block = [flag word, padded to 256 B][buf 0: cap f64][buf 1: cap f64]
  bool peer_on = false;
  int peer_rank = 0, peer_world = 1;
  size_t peer_cap = 0;                 // doubles per buffer
  unsigned char *peer_block[cge::kMaxPeers] = {};  // [peer_rank] is this context's own allocation
  unsigned long long peer_step = 0;    // steps completed; every rank advances in lockstep
  std::vector<cudaEvent_t> prof_events;
  int prof_cap = 0;       // steps that fit
  int prof_step = 0;      // steps captured so far
  bool prof_on = false;
};

namespace {

int bind(cge_context *ctx) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  CK(cudaSetDevice(ctx->device));
  return CGE_OK;
}

int check_params(const cge_params *p, int *rb, int *re) {
  if (!p) return set_error(CGE_ERR_BAD_ARG, "params is NULL");
  if (p->n_obs < 1) return set_error(CGE_ERR_BAD_ARG, "n_obs must be >= 1 (got %d)", p->n_obs);
  if (p->grid_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "grid_size must be >= 1 (got %d)", p->grid_size);
  if (!(p->sigma_start > 0.0f) || !(p->sigma_end > 0.0f))
    return set_error(CGE_ERR_BAD_ARG, "sigma range must be > 0 (got [%g, %g])", p->sigma_start,
                     p->sigma_end);
  if (!std::isfinite(p->mu_start) || !std::isfinite(p->mu_end) || !std::isfinite(p->sigma_start) ||
      !std::isfinite(p->sigma_end))
    return set_error(CGE_ERR_BAD_ARG, "ranges must be finite");
  if (p->mode != CGE_MODE_PRODUCT_F32 && p->mode != CGE_MODE_LOGSPACE)
    return set_error(CGE_ERR_BAD_ARG, "unknown mode %d", p->mode);
  switch (p->variant) {
    case CGE_VARIANT_DEFAULT: case CGE_VARIANT_MUFU_ONLY: case CGE_VARIANT_PAIRED: case CGE_VARIANT_PAIRED3:
    case CGE_VARIANT_HYBRID: case CGE_VARIANT_LOG_F64: case CGE_VARIANT_UNFUSED:
    case 6: case 31: case 39: case 45: case 381: break;
    default: return set_error(CGE_ERR_BAD_ARG, "unknown kernel variant %d", p->variant);
  }
  int b = p->row_begin, e = p->row_end;
  if (b == 0 && e == 0) e = p->grid_size;
  if (b < 0 || e > p->grid_size || b >= e)
    return set_error(CGE_ERR_BAD_ARG, "row shard [%d, %d) outside grid of %d rows", b, e,
                     p->grid_size);
  *rb = b;
  *re = e;
  return CGE_OK;
}

// log-space kernels other than the float64 ones read centred observations (x_k - mu_c)
int centred_obs(const cge_params *p) {
  return (p->mode == CGE_MODE_LOGSPACE && p->variant != CGE_VARIANT_LOG_F64) ? 1 : 0;
}

// tile shape and grid for one shard
void make_plan(const cge_context *ctx, const cge_params *p, int rb, int re, const float *post,
               Plan *pl) {
  pl->N = p->grid_size;
  pl->n = p->n_obs;
  pl->n_pad = (p->n_obs + 3) & ~3;
  pl->mode = p->mode;
  pl->variant = p->variant;
  pl->row_begin = rb;
  pl->row_end = re;
  pl->rows_local = re - rb;
  // columns per warp strip: 32 lanes x Policy::kCols (log-space keeps 8 columns per thread)
  const bool wide_patch = p->mode == CGE_MODE_PRODUCT_F32 && p->variant == 381;
  const int strip = 32 * (p->mode == CGE_MODE_LOGSPACE ? cge::LogSpaceTiled<>::kCols : wide_patch ? 8 : 4);
  pl->ncw = (pl->N + strip - 1) / strip;
  if (pl->ncw >= 8) { pl->wm = 1; pl->wn = 8; }
  else if (pl->ncw >= 4) { pl->wm = 2; pl->wn = 4; }
  else if (pl->ncw >= 2) { pl->wm = 4; pl->wn = 2; }
  else { pl->wm = 8; pl->wn = 1; }
  if (const char *ov = std::getenv("CGE_WARP_ROWS")) {  // tuning knob, not part of the ABI
    const int v = std::atoi(ov);
    if (v == 1 || v == 2 || v == 4 || v == 8) { pl->wm = v; pl->wn = 8 / v; }
  }
  pl->bands = (pl->ncw + pl->wn - 1) / pl->wn;
  // tile heights are multiples of the patch rows of every warp row and of the global alignment
  const int patch_rows = p->mode == CGE_MODE_LOGSPACE ? cge::LogSpaceF64::kRows : cge::kRM;
  int unit = pl->wm * patch_rows;
  if (unit % cge::kRM) unit *= cge::kRM / patch_rows;
  // Tile height: tall enough that the per-tile column partials stay small (the fold kernel reads
  // row_tiles*WM*N doubles) and the per-CTA prologue is amortised, short enough that the grid is
  // still >= ~8 CTAs per resident slot.  Measured sweep: profiles/r01_tile_rows_sweep.jsonl
  // (16..48 rows is flat for the likelihood kernel; the fold time falls with the height).
  const int span = re - (rb & ~(cge::kRM - 1));  // tiles start at a multiple of kRM (see kernel)
  const long long want_ctas = (long long)ctx->sm_count * 2 * 8;
  long long tr = ((long long)span * pl->bands) / want_ctas;
  if (tr < 16) tr = 16;
  if (tr > 48) tr = 48;
  if (tr > span) tr = span;
  tr = ((tr + unit - 1) / unit) * unit;
  if (const char *ov = std::getenv("CGE_TILE_ROWS")) {  // tuning knob, not part of the ABI
    long long v = std::atoll(ov);
    if (v >= unit) tr = (v / unit) * unit;
  }
  pl->tile_rows = (int)tr;
  pl->row_tiles = (span + pl->tile_rows - 1) / pl->tile_rows;
  pl->ncolparts = pl->row_tiles * pl->wm;
  pl->nb_col = (pl->N + 31) / 32;
  pl->nb_row = (pl->rows_local + 7) / 8;
  pl->vec4 = (pl->N % 4 == 0) && ((reinterpret_cast<uintptr_t>(post) & 15u) == 0);
  const bool streamed = pl->n_pad > cge::kObsSingleMax;
  pl->smem = 128 + (streamed ? 2 * (size_t)cge::kObsChunk * 4 : (size_t)pl->n_pad * 4);
  pl->cluster = false;
  pl->valid = true;
}

// ---------------------------------------------------------------------------------------
// single-launch path for small grids (cge::small_grid_kernel): whole grid, no exchange
// ---------------------------------------------------------------------------------------
constexpr int kSmallGridMax = 256;
constexpr int kZeroCopyObsMax = 1024;  // host observations read in place (mapped pinned) up to this n
int mark(cge_context *ctx, int k, cudaStream_t st);

bool small_grid_eligible(const cge_context *ctx, const cge_params *p, int rb, int re) {
  return p->variant == CGE_VARIANT_DEFAULT && p->grid_size <= kSmallGridMax && rb == 0 &&
         re == p->grid_size && !ctx->peer_on && ((p->n_obs + 3) & ~3) <= cge::kObsSingleMax;
}

int reserve_step_buffers(cge_context *ctx, const Plan &pl, cudaStream_t st) {
  const int N = pl.N;
  CKRC(ctx->mu.reserve((size_t)N + 4));
  CKRC(ctx->sigma.reserve(N));
  CKRC(ctx->ad.reserve(N));
  CKRC(ctx->cd0.reserve(N));
  CKRC(ctx->obs_pad.reserve(pl.n_pad));
  CKRC(ctx->scale.reserve(1));
  CKRC(ctx->espart.reserve((size_t)(N + 255) / 256));
  if (!ctx->tickets.ptr) {
    CKRC(ctx->tickets.reserve(3));
    CK(cudaMemsetAsync(ctx->tickets.ptr, 0, 3 * sizeof(unsigned int), st));
  }
  CKRC(ctx->rowpart.reserve((size_t)pl.rows_local * pl.ncw));
  CKRC(ctx->colpart.reserve((size_t)pl.ncolparts * N));
  CKRC(ctx->partials.reserve((size_t)N + 2));
  CKRC(ctx->rowsum.reserve(pl.rows_local));
  CKRC(ctx->zpart.reserve(pl.nb_row));
  CKRC(ctx->smupart.reserve(pl.nb_row));
  CKRC(ctx->results.reserve(4 + 2 * (size_t)N));
  return CGE_OK;
}

int fill_like_args(cge_context *ctx, const Plan &pl, float *posterior_dev, cge::LikeArgs *args) {
  const int N = pl.N;
  args->obs_pad = ctx->obs_pad.ptr;
  args->mu = ctx->mu.ptr;
  args->ad = ctx->ad.ptr;
  args->cd0 = ctx->cd0.ptr;
  args->scale = ctx->scale.ptr;
  if ((ctx->has_prior_mu || ctx->has_prior_sigma) && ctx->prior_n != N)
    return set_error(CGE_ERR_BAD_ARG, "prior was set for grid_size %d, call has %d", ctx->prior_n, N);
  args->prior_mu = ctx->has_prior_mu ? ctx->prior_mu.ptr : nullptr;
  args->prior_sigma = ctx->has_prior_sigma ? ctx->prior_sigma.ptr : nullptr;
  args->prior_full = ctx->prior_full;
  args->post = posterior_dev;
  args->rowpart = ctx->rowpart.ptr;
  args->colpart = ctx->colpart.ptr;
  args->n = pl.n;
  args->n_pad = pl.n_pad;
  args->N = N;
  args->row_begin = pl.row_begin;
  args->row_end = pl.row_end;
  args->tile_rows = pl.tile_rows;
  args->ncw = pl.ncw;
  return CGE_OK;
}

// the whole step in one launch; leaves the context as cge_likelihood_partials + cge_finalize would
int small_grid_step(cge_context *ctx, const cge_params *p, const float *obs_dev,
                    float *posterior_dev, cudaStream_t st, bool results_to_host) {
  Plan &pl = ctx->plan;
  pl.valid = false;
  make_plan(ctx, p, 0, p->grid_size, posterior_dev, &pl);
  // One column band (always the case up to 256 columns) and at most 8 row tiles (the portable
  // cluster size) -> the grid runs as one thread-block cluster: taller tiles if need be.
  if (pl.bands == 1 && !std::getenv("CGE_NO_CLUSTER")) {
    const int unit = pl.wm * cge::kRM;
    if (pl.row_tiles > 8) {
      pl.tile_rows = (((pl.N + 7) / 8 + unit - 1) / unit) * unit;
      pl.row_tiles = (pl.N + pl.tile_rows - 1) / pl.tile_rows;
      pl.ncolparts = pl.row_tiles * pl.wm;
    }
    pl.cluster = pl.row_tiles <= 8;
  }
  CKRC(reserve_step_buffers(ctx, pl, st));
  cge::SmallArgs a = {};
  CKRC(fill_like_args(ctx, pl, posterior_dev, &a.like));
  a.range = cge::RangeArgs{pl.N, p->mu_start, p->mu_end, p->sigma_start, p->sigma_end};
  a.obs = obs_dev;
  a.mu = ctx->mu.ptr;
  a.sigma = ctx->sigma.ptr;
  a.ad = ctx->ad.ptr;
  a.cd0 = ctx->cd0.ptr;
  a.scale = ctx->scale.ptr;
  a.centred = centred_obs(p);
  a.ncolparts = pl.ncolparts;
  a.partials = ctx->partials.ptr;
  a.rowsum = ctx->rowsum.ptr;
  a.results = ctx->results.ptr;
  a.results_host = nullptr;
  if (results_to_host) {
    CKRC(ctx->host_results.reserve(4 + 2 * (size_t)pl.N));
    a.results_host = ctx->host_results.ptr;
  }
  ctx->host_results_valid = results_to_host;
  a.ticket = ctx->tickets.ptr + 2;
  ctx->launches = 0;
  CKRC(mark(ctx, 0, st));
  CKRC(mark(ctx, 1, st));
  if (p->mode == CGE_MODE_PRODUCT_F32)
    CKRC(cge_host::launch_small_product(pl, a, st));
  else
    CKRC(cge_host::launch_small_logspace(pl, a, st));
  ctx->launches += 1;
  for (int k = 2; k <= 6; ++k) CKRC(mark(ctx, k, st));
  if (ctx->prof_on && ctx->prof_step < ctx->prof_cap) ctx->prof_step++;
  return CGE_OK;
}

int peer_release(cge_context *ctx);
constexpr size_t kPeerHeader = 256;
double *peer_buf(const cge_context *ctx, int rank, unsigned long long step) {
  return reinterpret_cast<double *>(ctx->peer_block[rank] + kPeerHeader) + (step & 1) * ctx->peer_cap;
}
unsigned long long *peer_flag(const cge_context *ctx, int rank) {
  return reinterpret_cast<unsigned long long *>(ctx->peer_block[rank]);
}

// record profiling mark `k` of the current step (no-op unless capture is on and has room)
int mark(cge_context *ctx, int k, cudaStream_t st) {
  if (!ctx->prof_on || ctx->prof_step >= ctx->prof_cap) return CGE_OK;
  CK(cudaEventRecord(ctx->prof_events[(size_t)ctx->prof_step * 7 + k], st));
  return CGE_OK;
}

int prof_reserve(cge_context *ctx, int steps) {
  while ((int)ctx->prof_events.size() < steps * 7) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    ctx->prof_events.push_back(e);
  }
  ctx->prof_cap = steps;
  ctx->prof_step = 0;
  return CGE_OK;
}

// mean per-phase times over the captured steps (synchronises on the last mark)
int prof_collect(cge_context *ctx, cge_timing *out, int *steps) {
  cge_timing t = {};
  const int n = ctx->prof_step;
  if (n > 0) CK(cudaEventSynchronize(ctx->prof_events[(size_t)(n - 1) * 7 + 6]));
  for (int s = 0; s < n; ++s) {
    cudaEvent_t *e = &ctx->prof_events[(size_t)s * 7];
    float a = 0, b = 0, c = 0, d = 0, f = 0, tot = 0;
    CK(cudaEventElapsedTime(&a, e[0], e[1]));
    CK(cudaEventElapsedTime(&b, e[1], e[2]));
    CK(cudaEventElapsedTime(&c, e[2], e[3]));
    CK(cudaEventElapsedTime(&d, e[4], e[5]));
    CK(cudaEventElapsedTime(&f, e[5], e[6]));
    CK(cudaEventElapsedTime(&tot, e[0], e[6]));
    t.setup_ms += a; t.likelihood_ms += b; t.reduce_ms += c + d; t.scale_ms += f; t.total_ms += tot;
  }
  if (n > 0) {
    t.setup_ms /= n; t.likelihood_ms /= n; t.reduce_ms /= n; t.scale_ms /= n; t.total_ms /= n;
  }
  if (out) *out = t;
  if (steps) *steps = n;
  return CGE_OK;
}

bool is_device_pointer(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_version(void) { return CGE_VERSION; }

CGE_EXPORT const char *cge_last_error(void) { return cge_host::g_last_error.c_str(); }

CGE_EXPORT int cge_create(int device, cge_context **out) {
  if (!out) return set_error(CGE_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return set_error(CGE_ERR_NO_DEVICE, "no CUDA device (%s); libcge has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count)
    return set_error(CGE_ERR_BAD_ARG, "device %d out of range (have %d)", device, count);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(CGE_ERR_NO_DEVICE,
                     "device %d is sm_%d%d; libcge is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
  CK(cudaSetDevice(device));
  cge_context *ctx = new (std::nothrow) cge_context();
  if (!ctx) return set_error(CGE_ERR_CUDA, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (se != cudaSuccess) {
    delete ctx;
    return set_error(CGE_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
  }
  for (auto &ev : ctx->ev) {
    cudaError_t ee = cudaEventCreate(&ev);
    if (ee != cudaSuccess) {
      delete ctx;
      return set_error(CGE_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(ee));
    }
  }
  *out = ctx;
  return CGE_OK;
}

CGE_EXPORT int cge_destroy(cge_context *ctx) {
  if (!ctx) return CGE_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaDeviceSynchronize();
  peer_release(ctx);
  ctx->mu.release(); ctx->sigma.release(); ctx->obs_pad.release();
  ctx->obs_in.release(); ctx->espart.release(); ctx->tickets.release(); ctx->prior_mu.release(); ctx->prior_sigma.release();
  ctx->posterior.release(); ctx->ad.release(); ctx->cd0.release(); ctx->rowpart.release();
  ctx->colpart.release(); ctx->partials.release(); ctx->rowsum.release(); ctx->zpart.release();
  ctx->smupart.release(); ctx->results.release(); ctx->tmp.release(); ctx->scale.release();
  ctx->sink.release();
  ctx->host_obs.release();
  ctx->host_results.release();
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : ctx->prof_events) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CGE_OK;
}

CGE_EXPORT int cge_device_sm_count(cge_context *ctx, int *sm_count, int *cc_major, int *cc_minor) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// phase 1: tables + scale + fused likelihood + local fold
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_likelihood_partials(cge_context *ctx, const cge_params *p, const float *obs_dev,
                                       float *posterior_dev, void *stream) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (!obs_dev || !posterior_dev)
    return set_error(CGE_ERR_BAD_ARG, "obs_dev and posterior_dev must be device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  Plan &pl = ctx->plan;
  pl.valid = false;
  ctx->host_results_valid = false;
  make_plan(ctx, p, rb, re, posterior_dev, &pl);
  const int N = pl.N;

  CKRC(reserve_step_buffers(ctx, pl, st));

  ctx->launches = 0;
  CKRC(mark(ctx, 0, st));
  const cge::RangeArgs range = {N, p->mu_start, p->mu_end, p->sigma_start, p->sigma_end};
  cge::setup_kernel<<<(N + 3 + 255) / 256, 256, 0, st>>>(
      ctx->mu.ptr, ctx->sigma.ptr, ctx->ad.ptr, ctx->cd0.ptr, range, obs_dev, ctx->obs_pad.ptr, pl.n,
      pl.n_pad, centred_obs(p), ctx->scale.ptr);
  CK(cudaGetLastError());
  ctx->launches += 1;
  CKRC(mark(ctx, 1, st));

  cge::LikeArgs args;
  CKRC(fill_like_args(ctx, pl, posterior_dev, &args));
  if (p->mode == CGE_MODE_PRODUCT_F32) {
    CKRC(cge_host::launch_product(pl, args, st));
  } else {
    CKRC(cge_host::launch_logspace(pl, args, st));
  }
  ctx->launches += 1;
  CKRC(mark(ctx, 2, st));

  double *fold_out = ctx->partials.ptr;
  unsigned long long *ready = nullptr;
  if (ctx->peer_on) {  // this step's vector goes where the peers can read it
    if ((size_t)N + 2 > ctx->peer_cap)
      return set_error(CGE_ERR_BAD_ARG, "grid_size %d exceeds the peer exchange capacity %zu", N,
                       ctx->peer_cap - 2);
    fold_out = peer_buf(ctx, ctx->peer_rank, ctx->peer_step);
    ready = peer_flag(ctx, ctx->peer_rank);
  }
  cge::fold_partials_kernel<<<pl.nb_col + pl.nb_row, 256, 0, st>>>(
      ctx->colpart.ptr, pl.ncolparts, ctx->rowpart.ptr, pl.ncw, pl.rows_local, rb, N, ctx->mu.ptr,
      fold_out, ctx->rowsum.ptr, ctx->zpart.ptr, ctx->smupart.ptr, pl.nb_col, pl.nb_row,
      ctx->tickets.ptr, ready, ctx->peer_step + 1);
  CK(cudaGetLastError());
  ctx->launches += 1;
  CKRC(mark(ctx, 3, st));
  return CGE_OK;
}

// optional non-flat prior for the following calls (NULL, NULL, n, NULL restores the flat prior)
CGE_EXPORT int cge_set_prior(cge_context *ctx, const float *prior_mu, const float *prior_sigma,
                             int grid_size, const float *prior_full_dev) {
  CKRC(bind(ctx));
  if ((prior_mu || prior_sigma) && grid_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "set_prior: grid_size must be >= 1");
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->has_prior_mu = ctx->has_prior_sigma = false;
  ctx->prior_n = grid_size;
  if (prior_mu) {
    CKRC(ctx->prior_mu.reserve(grid_size));
    CK(cudaMemcpy(ctx->prior_mu.ptr, prior_mu, (size_t)grid_size * sizeof(float), cudaMemcpyDefault));
    ctx->has_prior_mu = true;
  }
  if (prior_sigma) {
    CKRC(ctx->prior_sigma.reserve(grid_size));
    CK(cudaMemcpy(ctx->prior_sigma.ptr, prior_sigma, (size_t)grid_size * sizeof(float),
                  cudaMemcpyDefault));
    ctx->has_prior_sigma = true;
  }
  if (prior_full_dev && !is_device_pointer(prior_full_dev))
    return set_error(CGE_ERR_BAD_ARG, "set_prior: the full prior table must be device memory");
  ctx->prior_full = prior_full_dev;
  return CGE_OK;
}

CGE_EXPORT int cge_partials(cge_context *ctx, double **partials_dev, size_t *count) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  if (!ctx->plan.valid) return set_error(CGE_ERR_BAD_ARG, "no likelihood call has been made");
  if (partials_dev) *partials_dev = ctx->partials.ptr;
  if (count) *count = (size_t)ctx->plan.N + 2;
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// phase 2: results + in-place scale
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_finalize(cge_context *ctx, const cge_params *p, float *posterior_dev,
                            void *stream) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  const Plan &pl = ctx->plan;
  if (!pl.valid || pl.N != p->grid_size || pl.row_begin != rb || pl.row_end != re)
    return set_error(CGE_ERR_BAD_ARG, "cge_finalize does not match the last cge_likelihood_partials");
  if (!posterior_dev) return set_error(CGE_ERR_BAD_ARG, "posterior_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl.N;
  CKRC(mark(ctx, 4, st));
  if (ctx->peer_on) {
    cge::PeerView pv = {};
    pv.world = ctx->peer_world;
    pv.want = ctx->peer_step + 1;
    for (int r = 0; r < ctx->peer_world; ++r) {
      pv.buf[r] = peer_buf(ctx, r, ctx->peer_step);
      pv.flag[r] = peer_flag(ctx, r);
    }
    cge::results_peer_kernel<<<(N + 255) / 256, 256, 0, st>>>(
        pv, ctx->partials.ptr, ctx->rowsum.ptr, ctx->sigma.ptr, N, rb, re, ctx->results.ptr,
        ctx->espart.ptr, ctx->tickets.ptr + 1);
    ctx->peer_step++;
  } else {
    cge::results_kernel<<<(N + 255) / 256, 256, 0, st>>>(ctx->partials.ptr, ctx->rowsum.ptr,
                                                         ctx->sigma.ptr, N, rb, re,
                                                         ctx->results.ptr, ctx->espart.ptr,
                                                         ctx->tickets.ptr + 1);
  }
  CK(cudaGetLastError());
  CKRC(mark(ctx, 5, st));
  const size_t cells = (size_t)pl.rows_local * N;
  const bool aligned = (reinterpret_cast<uintptr_t>(posterior_dev) & 15u) == 0;
  // enough CTAs to keep every SM's memory pipeline full, capped so tiny grids stay tiny
  size_t want = aligned ? (cells / 4 + 1023) / 1024 : (cells + 255) / 256;
  size_t cap = (size_t)ctx->sm_count * 16;
  int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  if (aligned)
    cge::scale_posterior_kernel<<<blocks, 256, 0, st>>>(posterior_dev, cells, ctx->partials.ptr);
  else
    cge::scale_posterior_scalar_kernel<<<blocks, 256, 0, st>>>(posterior_dev, cells,
                                                               ctx->partials.ptr);
  CK(cudaGetLastError());
  ctx->launches += 2;
  CKRC(mark(ctx, 6, st));
  if (ctx->prof_on && ctx->prof_step < ctx->prof_cap) ctx->prof_step++;
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// peer exchange over CUDA IPC (one process per GPU, one node)
// ---------------------------------------------------------------------------------------
namespace {
int peer_release(cge_context *ctx) {
  for (int r = 0; r < cge::kMaxPeers; ++r) {
    if (!ctx->peer_block[r]) continue;
    if (r == ctx->peer_rank) cudaFree(ctx->peer_block[r]);
    else cudaIpcCloseMemHandle(ctx->peer_block[r]);
    ctx->peer_block[r] = nullptr;
  }
  ctx->peer_on = false;
  ctx->peer_world = 1;
  ctx->peer_rank = 0;
  ctx->peer_cap = 0;
  ctx->peer_step = 0;
  return CGE_OK;
}
}  // namespace

CGE_EXPORT int cge_peer_local_handle(cge_context *ctx, void *handle, int max_grid_size) {
  CKRC(bind(ctx));
  static_assert(sizeof(cudaIpcMemHandle_t) == CGE_IPC_HANDLE_BYTES, "IPC handle size");
  if (!handle || max_grid_size < 1) return set_error(CGE_ERR_BAD_ARG, "peer_local_handle: bad argument");
  CK(cudaStreamSynchronize(ctx->stream));
  peer_release(ctx);
  ctx->peer_cap = (((size_t)max_grid_size + 2) + 31) & ~(size_t)31;
  const size_t bytes = kPeerHeader + 2 * ctx->peer_cap * sizeof(double);
  unsigned char *block = nullptr;
  CK(cudaMalloc(&block, bytes));
  CK(cudaMemset(block, 0, bytes));
  CK(cudaDeviceSynchronize());
  ctx->peer_block[0] = block;  // parked at slot 0 until cge_peer_connect tells us our rank
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, block));
  std::memcpy(handle, &h, sizeof h);
  return CGE_OK;
}

CGE_EXPORT int cge_peer_connect(cge_context *ctx, int rank, int world, const void *handles) {
  CKRC(bind(ctx));
  if (!handles || world < 2 || world > cge::kMaxPeers || rank < 0 || rank >= world)
    return set_error(CGE_ERR_BAD_ARG, "peer_connect: need 2..%d ranks", cge::kMaxPeers);
  if (!ctx->peer_block[0] || ctx->peer_on)
    return set_error(CGE_ERR_BAD_ARG, "peer_connect: call cge_peer_local_handle first");
  unsigned char *mine = ctx->peer_block[0];
  ctx->peer_block[0] = nullptr;
  ctx->peer_rank = rank;
  ctx->peer_block[rank] = mine;
  const unsigned char *hb = static_cast<const unsigned char *>(handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + (size_t)r * sizeof h, sizeof h);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ctx->peer_world = world;
      peer_release(ctx);
      return set_error(CGE_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r,
                       cudaGetErrorString(e));
    }
    ctx->peer_block[r] = static_cast<unsigned char *>(p);
  }
  ctx->peer_world = world;
  ctx->peer_step = 0;
  ctx->peer_on = true;
  return CGE_OK;
}

CGE_EXPORT int cge_peer_disconnect(cge_context *ctx) {
  CKRC(bind(ctx));
  CK(cudaDeviceSynchronize());
  return peer_release(ctx);
}

CGE_EXPORT int cge_results(cge_context *ctx, double **results_dev, size_t *count) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  if (!ctx->plan.valid) return set_error(CGE_ERR_BAD_ARG, "no likelihood call has been made");
  if (results_dev) *results_dev = ctx->results.ptr;
  if (count) *count = 4 + 2 * (size_t)ctx->plan.N;
  return CGE_OK;
}

CGE_EXPORT int cge_read_results(cge_context *ctx, const cge_params *p, double *expectations,
                                double *marg_mu, double *marg_sigma, double *z, void *stream) {
  CKRC(bind(ctx));
  if (!ctx->plan.valid || !p || ctx->plan.N != p->grid_size)
    return set_error(CGE_ERR_BAD_ARG, "cge_read_results does not match the last call");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)ctx->plan.N;
  double head[4];
  const bool dev_out = (marg_sigma && is_device_pointer(marg_sigma)) ||
                       (marg_mu && is_device_pointer(marg_mu)) ||
                       (expectations && is_device_pointer(expectations));
  if (!dev_out) {
    // host destinations: one DMA of the whole results block into pinned staging, then host copies
    const size_t count = 4 + ((marg_mu || marg_sigma) ? 2 * N : 0);
    if (!ctx->host_results_valid) {  // (small_grid_kernel writes the staging block itself)
      CKRC(ctx->host_results.reserve(count));
      CK(cudaMemcpyAsync(ctx->host_results.ptr, ctx->results.ptr, count * sizeof(double),
                         cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    const double *h = ctx->host_results.ptr;
    std::memcpy(head, h, sizeof head);
    if (marg_sigma) std::memcpy(marg_sigma, h + 4, N * sizeof(double));
    if (marg_mu) std::memcpy(marg_mu, h + 4 + N, N * sizeof(double));
    if (expectations) {
      expectations[0] = head[0];
      expectations[1] = head[1];
    }
  } else {
    CK(cudaMemcpyAsync(head, ctx->results.ptr, sizeof head, cudaMemcpyDeviceToHost, st));
    if (marg_sigma)
      CK(cudaMemcpyAsync(marg_sigma, ctx->results.ptr + 4, N * sizeof(double), cudaMemcpyDefault, st));
    if (marg_mu)
      CK(cudaMemcpyAsync(marg_mu, ctx->results.ptr + 4 + N, N * sizeof(double), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    if (expectations) {
      if (is_device_pointer(expectations))
        CK(cudaMemcpy(expectations, head, 2 * sizeof(double), cudaMemcpyHostToDevice));
      else {
        expectations[0] = head[0];
        expectations[1] = head[1];
      }
    }
  }
  if (z) *z = head[2];
  if (!(head[2] > 0.0) || !std::isfinite(head[2]))
    return set_error(CGE_ERR_DEGENERATE,
                     "normaliser Z = %g: every cell underflowed or the inputs are not finite "
                     "(use CGE_MODE_LOGSPACE for large n_obs)", head[2]);
  return CGE_OK;
}

CGE_EXPORT int cge_last_timing(cge_context *ctx, cge_timing *timing) {
  if (!ctx || !timing) return set_error(CGE_ERR_BAD_ARG, "NULL argument");
  *timing = ctx->timing;
  return CGE_OK;
}

CGE_EXPORT int cge_profile_begin(cge_context *ctx, int max_steps) {
  CKRC(bind(ctx));
  if (max_steps < 1 || max_steps > 65536) return set_error(CGE_ERR_BAD_ARG, "max_steps out of range");
  CKRC(prof_reserve(ctx, max_steps));
  ctx->prof_on = true;
  return CGE_OK;
}

CGE_EXPORT int cge_profile_end(cge_context *ctx, cge_timing *mean, int *steps_captured) {
  CKRC(bind(ctx));
  int rc = prof_collect(ctx, mean, steps_captured);
  ctx->prof_on = false;
  return rc;
}

CGE_EXPORT int cge_launch_count(cge_context *ctx, int *launches) {
  if (!ctx || !launches) return set_error(CGE_ERR_BAD_ARG, "NULL argument");
  *launches = ctx->launches;
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// fused synchronous entry: host or device buffers
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_grid_posterior(cge_context *ctx, const cge_params *p, const float *obs,
                                  float *posterior, double *marg_mu, double *marg_sigma,
                                  double *expectations, cge_timing *timing) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (!obs) return set_error(CGE_ERR_BAD_ARG, "obs is NULL");
  cudaStream_t st = ctx->stream;
  const size_t cells = (size_t)(re - rb) * p->grid_size;
  ctx->timing = cge_timing{};
  const bool was_on = ctx->prof_on;
  if (timing && !was_on) {
    CKRC(prof_reserve(ctx, 1));
    ctx->prof_on = true;
  }

  const float *obs_dev = obs;
  const bool small = small_grid_eligible(ctx, p, rb, re);
  if (timing) CK(cudaEventRecord(ctx->ev[7], st));
  if (!is_device_pointer(obs)) {
    // (the previous call's use of this staging buffer has completed: every call ends with a
    // stream synchronisation in cge_read_results)
    CKRC(ctx->host_obs.reserve(p->n_obs));
    std::memcpy(ctx->host_obs.ptr, obs, (size_t)p->n_obs * sizeof(float));
    if (small && p->n_obs <= kZeroCopyObsMax) {
      obs_dev = ctx->host_obs.ptr;  // mapped pinned memory: the kernel reads it in place
    } else {
      CKRC(ctx->obs_in.reserve(p->n_obs));
      CK(cudaMemcpyAsync(ctx->obs_in.ptr, ctx->host_obs.ptr, (size_t)p->n_obs * sizeof(float),
                         cudaMemcpyHostToDevice, st));
      obs_dev = ctx->obs_in.ptr;
    }
  }
  if (timing) CK(cudaEventRecord(ctx->ev[0], st));
  float *post_dev = posterior;
  const bool post_to_host = posterior && !is_device_pointer(posterior);
  if (!posterior || post_to_host) {
    CKRC(ctx->posterior.reserve(cells));
    post_dev = ctx->posterior.ptr;
  }
  int rc;
  if (small) {
    const bool host_out = !(marg_mu && is_device_pointer(marg_mu)) &&
                          !(marg_sigma && is_device_pointer(marg_sigma)) &&
                          !(expectations && is_device_pointer(expectations));
    rc = small_grid_step(ctx, p, obs_dev, post_dev, st, host_out);
  } else {
    rc = cge_likelihood_partials(ctx, p, obs_dev, post_dev, st);
    if (rc == CGE_OK) rc = cge_finalize(ctx, p, post_dev, st);
  }
  if (rc != CGE_OK) {
    if (!was_on) ctx->prof_on = false;
    return rc;
  }
  if (timing) CK(cudaEventRecord(ctx->ev[1], st));
  if (post_to_host)
    CK(cudaMemcpyAsync(posterior, post_dev, cells * sizeof(float), cudaMemcpyDeviceToHost, st));
  rc = cge_read_results(ctx, p, expectations, marg_mu, marg_sigma, nullptr, st);
  if (timing) {
    cudaEventRecord(ctx->ev[2], st);
    cudaEventSynchronize(ctx->ev[2]);
    cge_timing t = {};
    if (!was_on) {
      prof_collect(ctx, &t, nullptr);
      ctx->prof_on = false;
    }
    cudaEventElapsedTime(&t.h2d_ms, ctx->ev[7], ctx->ev[0]);
    cudaEventElapsedTime(&t.d2h_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&t.total_ms, ctx->ev[7], ctx->ev[2]);
    ctx->timing = t;
    *timing = t;
  }
  return rc;
}

// asynchronous fused entry: device buffers, whole grid, nothing waited for
CGE_EXPORT int cge_grid_posterior_async(cge_context *ctx, const cge_params *p, const float *obs_dev,
                                        float *posterior_dev, void *stream) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (rb != 0 || re != p->grid_size)
    return set_error(CGE_ERR_BAD_ARG, "cge_grid_posterior_async needs the whole grid; a shard [%d, %d) "
                     "goes through cge_likelihood_partials / cge_finalize", rb, re);
  if (!obs_dev || !posterior_dev || !is_device_pointer(obs_dev) || !is_device_pointer(posterior_dev))
    return set_error(CGE_ERR_BAD_ARG, "cge_grid_posterior_async needs device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  if (small_grid_eligible(ctx, p, rb, re)) return small_grid_step(ctx, p, obs_dev, posterior_dev, st, false);
  CKRC(cge_likelihood_partials(ctx, p, obs_dev, posterior_dev, st));
  return cge_finalize(ctx, p, posterior_dev, st);
}

// posterior kept by the context when cge_grid_posterior was called with posterior == NULL or a
// host pointer
CGE_EXPORT int cge_posterior_dev(cge_context *ctx, float **posterior_dev, size_t *count) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  if (posterior_dev) *posterior_dev = ctx->posterior.ptr;
  if (count) *count = ctx->plan.valid ? (size_t)ctx->plan.rows_local * ctx->plan.N : 0;
  return CGE_OK;
}

// scale statistics of the last likelihood call (waits for `stream`)
CGE_EXPORT int cge_scale_info(cge_context *ctx, double *out, int count, void *stream) {
  CKRC(bind(ctx));
  if (!out || count < 1) return set_error(CGE_ERR_BAD_ARG, "scale_info: bad argument");
  if (!ctx->plan.valid || !ctx->scale.ptr)
    return set_error(CGE_ERR_BAD_ARG, "no likelihood call has been made");
  cge::ScaleInfo h;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(&h, ctx->scale.ptr, sizeof h, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const double v[12] = {h.log2s, h.shift, h.m2, h.xbar, h.ss, h.smin, h.mu_c, h.s_c, h.umax,
                        cge::kPairedUmax, cge::kPairedUmax3, h.ustep};
  for (int i = 0; i < count && i < 12; ++i) out[i] = v[i];
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// stage API
// ---------------------------------------------------------------------------------------
static int grid_for(size_t count, int threads, int sm_count) {
  size_t want = (count + threads - 1) / threads;
  size_t cap = (size_t)sm_count * 32;
  if (want < 1) want = 1;
  return (int)(want > cap ? cap : want);
}

CGE_EXPORT int cge_linspace(cge_context *ctx, float *array_dev, int size, float start, float end) {
  CKRC(bind(ctx));
  if (!array_dev || size < 1) return set_error(CGE_ERR_BAD_ARG, "linspace: bad array/size");
  cge::linspace_kernel<<<(size + 255) / 256, 256, 0, ctx->stream>>>(array_dev, size, start, end);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_meshgrid3(cge_context *ctx, const float *vec_x, const float *vec_y,
                             const float *vec_z, float *grid_x, float *grid_y, float *grid_z,
                             int vec_xy_size, int vec_z_size) {
  CKRC(bind(ctx));
  if (!vec_x || !vec_y || !vec_z || !grid_x || !grid_y || !grid_z || vec_xy_size < 1 ||
      vec_z_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "meshgrid3: bad argument");
  size_t total = (size_t)vec_xy_size * vec_xy_size * vec_z_size;
  cge::meshgrid3_kernel<<<grid_for(total, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      vec_x, vec_y, vec_z, grid_x, grid_y, grid_z, vec_xy_size, vec_z_size);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_densities(cge_context *ctx, float *densities, const float *grid_x,
                             const float *grid_y, const float *grid_z, size_t grid_size) {
  CKRC(bind(ctx));
  if (!densities || !grid_x || !grid_y || !grid_z || grid_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "densities: bad argument");
  cge::densities_kernel<<<grid_for(grid_size, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      densities, grid_x, grid_y, grid_z, grid_size);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_densities_grid(cge_context *ctx, float *densities, const float *mu,
                                  const float *sigma, const float *obs, int grid_size, int n_obs) {
  CKRC(bind(ctx));
  if (!densities || !mu || !sigma || !obs || grid_size < 1 || n_obs < 1)
    return set_error(CGE_ERR_BAD_ARG, "densities_grid: bad argument");
  size_t total = (size_t)grid_size * grid_size * n_obs;
  cge::densities_grid_kernel<<<grid_for(total, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      densities, mu, sigma, obs, grid_size, n_obs);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_row_products_rows(cge_context *ctx, double *likes, double *const *rows_ptr,
                                     int rows, int cols) {
  CKRC(bind(ctx));
  if (!likes || !rows_ptr || rows < 1 || cols < 1)
    return set_error(CGE_ERR_BAD_ARG, "row_products_rows: bad argument");
  cge::row_products_pp_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(likes, rows_ptr, rows, cols);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_marginalize_rows(cge_context *ctx, double *marginal, double *const *rows_ptr,
                                    int rows, int cols, int axis) {
  CKRC(bind(ctx));
  if (!marginal || !rows_ptr || rows < 1 || cols < 1 || (axis != 0 && axis != 1))
    return set_error(CGE_ERR_BAD_ARG, "marginalize_rows: bad argument");
  const int outs = axis == 1 ? rows : cols;
  cge::marginalize_pp_kernel<<<(outs + 7) / 8, 256, 0, ctx->stream>>>(marginal, rows_ptr, rows, cols,
                                                                      axis);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_row_products(cge_context *ctx, const float *densities, double *likes,
                                size_t densities_size, size_t likes_size) {
  CKRC(bind(ctx));
  if (!densities || !likes || likes_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "row_products: bad argument");
  size_t cols = densities_size / likes_size;
  if (cols * likes_size != densities_size || cols < 1 || cols > 0x7fffffff)
    return set_error(CGE_ERR_SIZE, "likesSize != rows * cols (densities %zu, likes %zu)",
                     densities_size, likes_size);
  size_t blocks = (likes_size + 7) / 8;
  if (blocks > 0x7fffffff) return set_error(CGE_ERR_BAD_ARG, "row_products: too many rows");
  cge::row_products_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(densities, likes, likes_size,
                                                                      (int)cols);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_normalize(cge_context *ctx, const double *likes, double *posterior, size_t size) {
  CKRC(bind(ctx));
  if (!likes || !posterior || size < 1) return set_error(CGE_ERR_BAD_ARG, "normalize: bad argument");
  int nb = grid_for(size, 1024, ctx->sm_count);
  if (nb > 1024) nb = 1024;
  CKRC(ctx->tmp.reserve((size_t)nb + 1));
  cge::block_sums_kernel<<<nb, 1024, 0, ctx->stream>>>(likes, size, ctx->tmp.ptr + 1);
  CK(cudaGetLastError());
  cge::final_sum_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->tmp.ptr + 1, nb, ctx->tmp.ptr);
  CK(cudaGetLastError());
  cge::scale_f64_kernel<<<grid_for(size, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      likes, posterior, size, ctx->tmp.ptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_marginalize(cge_context *ctx, double *marginal, const double *matrix, int rows,
                               int cols, int axis) {
  CKRC(bind(ctx));
  if (!marginal || !matrix || rows < 1 || cols < 1 || (axis != 0 && axis != 1))
    return set_error(CGE_ERR_BAD_ARG, "marginalize: bad argument");
  int blocks = axis == 1 ? (rows + 7) / 8 : (cols + 255) / 256;
  cge::marginalize_kernel<<<blocks, 256, 0, ctx->stream>>>(marginal, matrix, rows, cols, axis);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

CGE_EXPORT int cge_expectation(cge_context *ctx, const double *marginal, const float *vector,
                               int size, double *expectation_host) {
  CKRC(bind(ctx));
  if (!marginal || !vector || size < 1 || !expectation_host)
    return set_error(CGE_ERR_BAD_ARG, "expectation: bad argument");
  CKRC(ctx->tmp.reserve(2));
  cge::expectation_kernel<<<1, 1024, 0, ctx->stream>>>(marginal, vector, size, ctx->tmp.ptr);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(expectation_host, ctx->tmp.ptr, sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// hooks for the other translation unit (cge_extras.cu); not part of the public ABI
// ---------------------------------------------------------------------------------------
extern "C" int cge_internal_stream(cge_context *ctx, void **stream, int *sm_count) {
  CKRC(bind(ctx));
  if (stream) *stream = ctx->stream;
  if (sm_count) *sm_count = ctx->sm_count;
  return CGE_OK;
}
extern "C" int cge_internal_error(int code, const char *msg) { return set_error(code, "%s", msg); }

// ---------------------------------------------------------------------------------------
// pipe-peak micro-benchmarks
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_microbench(cge_context *ctx, int kind, double *ops_per_second) {
  CKRC(bind(ctx));
  if (!ops_per_second || kind < 0 || kind > 10)
    return set_error(CGE_ERR_BAD_ARG, "microbench: bad argument");
  CKRC(ctx->sink.reserve(4));
  const int iters = 4096, chains = 8;
  // kinds 8, 9, 10: the 6 FFMA2 + 2 MUFU mix at 2, 3 and 4 CTAs per SM (8, 12, 16 warps per SM)
  const int blocks = ctx->sm_count * (kind == 8 ? 2 : kind == 9 ? 3 : kind == 10 ? 4 : 8), threads = 256;
  cudaStream_t st = ctx->stream;
  auto launch = [&]() {
    switch (kind) {
      case 0: cge::microbench_kernel<0><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 1: cge::microbench_kernel<1><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 2: cge::microbench_kernel<2><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 3: cge::microbench_kernel<3><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 4: cge::microbench_kernel<4><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 5: cge::microbench_kernel<5><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 6: cge::microbench_kernel<6><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      case 7: cge::microbench_kernel<7><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
      default: cge::microbench_kernel<6><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
    }
  };
  for (int w = 0; w < 3; ++w) launch();
  CK(cudaGetLastError());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(ctx->ev[0], st));
    launch();
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  // kinds 0-3: one instruction class, 8 chains; kinds 4-7: FMA lane-ops of the mix per iteration
  // (packed count as 2; the MUFU chains' bounding FFMA counts as 1; MUFU ops are not counted)
  static const double per_iter[11] = {8, 8, 8, 16, 4 * 2 + 4, 4 * 2 + 8, 6 * 2 + 2, 4 * 2 + 4 + 2,
                                      6 * 2 + 2, 6 * 2 + 2, 6 * 2 + 2};
  (void)chains;
  double ops = (double)blocks * threads * (double)iters * per_iter[kind];
  *ops_per_second = ops / (best * 1e-3);
  return CGE_OK;
}
