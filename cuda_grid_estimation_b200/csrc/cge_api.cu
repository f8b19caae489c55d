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
#include <map>
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

constexpr int kCopyChunksMax = 32;  // row chunks of a posterior that is copied to the host

struct cge_context {
  int device = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev[10] = {};
  cudaStream_t copy_stream = nullptr;          // host-posterior path: chunked device-to-host copies
  cudaEvent_t copy_ev[kCopyChunksMax + 1] = {};
  DevBuf<float> obs_pad, obs_in, posterior, like32;
  DevBuf<double> posterior64;
  DevBuf<double> rowpart, partials, rowsum, zpart, smupart, espart, results, tmp;
  DevBuf<unsigned long long> colacc;  // fixed-point column sums; zero between steps
  DevBuf<int> strip_next;             // claim counters; zero between steps
  float prior_mu_max = 1.0f, prior_sigma_max = 1.0f, prior_full_max = 1.0f;  // largest |factor| of each piece
  size_t prior_full_cells = 0;        // cells of the full table prior_full_max was measured over
  DevBuf<unsigned int> tickets;  // [0] fold, [1] finish, [2] small-grid kernel; zero between steps
  DevBuf<unsigned long long> zbcast;  // peer exchange: {step flag, Z, Smu} the finish kernel's block 0 publishes
  DevBuf<float> prior_mu, prior_sigma;   // owned copies of the separable prior factors
  bool has_prior_mu = false, has_prior_sigma = false;
  int prior_n = 0;
  const float *prior_full = nullptr;     // caller-owned device table, shard layout
  DevBuf<cge::ScaleInfo> scale;
  DevBuf<float> sink;
#ifdef CGE_DEBUG_TIMES
  DevBuf<unsigned long long> debug_times;
#endif
  PinnedBuf<float> host_obs;       // staging for host observations
  PinnedBuf<double> host_results;  // staging for {E, Z, 1/Z}, both marginals
  bool host_results_valid = false;  // the last step's kernel wrote host_results itself
  bool want_host_results = false;   // set by cge_grid_posterior around its finalize call
  Plan plan;
  cge_host::KernelChoice kernel;   // the likelihood kernel of the current plan
  cge_timing timing = {};
  int launches = 0;
  // tuning knobs, read from the environment ONCE at cge_create (not part of the ABI)
  int knob_ctas_per_sm = 0;  // CGE_CTAS_PER_SM: cap the persistent grid
  int knob_grid_mult = 1;    // CGE_GRID_MULT: CTAs per resident slot (1 = persistent)
  int knob_chunk = 0;        // CGE_ITEM_STEPS: CTA steps per claimed item
  bool knob_no_wide = false;     // CGE_NO_WIDE: never take the 8-column product policies
  bool knob_no_tall = false;     // CGE_NO_TALL: never take the 7-row one of them
  bool knob_no_cluster = false;  // CGE_NO_CLUSTER: small grids without the cluster launch
  bool knob_no_pdl = false;      // CGE_NO_PDL: ordinary stream-ordered launches
  // resident CTAs per SM of each likelihood kernel at its shared-memory size (occupancy query)
  std::map<std::pair<const void *, size_t>, int> occupancy;
  std::map<const void *, size_t> smem_optin;  // largest dynamic shared memory each kernel was opted in for
  // per-step event capture: 6 marks per step
  //   0 start | 1 after the streamed set-up (== 0 otherwise) | 2 after likelihood |
  //   3 after fold + publish | 4 before finish | 5 after finish
  // peer exchange: block = [flag word, padded to 256 B][buf 0: cap f64][buf 1: cap f64]
  bool peer_on = false;
  bool peer_ipc = false;               // peers mapped through CUDA IPC (else: same-process pointers)
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
  cudaGetLastError();  // a non-sticky error left by an unrelated earlier call must not be blamed on
                       // the first launch check of this one
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
  if (p->posterior_dtype != CGE_POSTERIOR_F32 && p->posterior_dtype != CGE_POSTERIOR_F64)
    return set_error(CGE_ERR_BAD_ARG, "unknown posterior_dtype %d", p->posterior_dtype);
  switch (p->variant) {
    case CGE_VARIANT_DEFAULT: case CGE_VARIANT_UNFUSED: break;
    case CGE_VARIANT_MUFU_ONLY: case CGE_VARIANT_PAIRED: case CGE_VARIANT_PAIRED3: case CGE_VARIANT_HYBRID:
    case CGE_VARIANT_SLOPE: case CGE_VARIANT_SLOPE_WIDE: case CGE_VARIANT_SLOPE_TALL:
      if (p->mode != CGE_MODE_PRODUCT_F32)
        return set_error(CGE_ERR_BAD_ARG, "variant %d belongs to the product mode", p->variant);
      break;
    case CGE_VARIANT_LOG_F64:
      if (p->mode != CGE_MODE_LOGSPACE)
        return set_error(CGE_ERR_BAD_ARG, "variant %d belongs to the log-space mode", p->variant);
      break;
    default: return set_error(CGE_ERR_BAD_ARG, "unknown kernel variant %d", p->variant);
  }
  if (p->mode == CGE_MODE_PRODUCT_F32 && ((p->n_obs + 3) & ~3) > cge::kObsSingleMax)
    return set_error(CGE_ERR_BAD_ARG,
                     "the float32 product path takes at most %d observations (got %d): use "
                     "CGE_MODE_LOGSPACE", cge::kObsSingleMax, p->n_obs);
  int b = p->row_begin, e = p->row_end;
  if (b == 0 && e == 0) e = p->grid_size;
  if (b < 0 || e > p->grid_size || b >= e)
    return set_error(CGE_ERR_BAD_ARG, "row shard [%d, %d) outside grid of %d rows", b, e,
                     p->grid_size);
  *rb = b;
  *re = e;
  return CGE_OK;
}

// log-space kernels other than the float64 one read centred observations (x_k - mu_c)
int centred_obs(const cge_params *p) {
  return (p->mode == CGE_MODE_LOGSPACE && p->variant != CGE_VARIANT_LOG_F64) ? 1 : 0;
}


// CTA shape and unit decomposition for one shard (everything but the persistent grid size, which
// needs the kernel: size_grid below).  `post` is where the float32 cells go.
void make_plan(const cge_context *ctx, const cge_params *p, int rb, int re, const float *post,
               Plan *pl) {
  pl->N = p->grid_size;
  pl->n = p->n_obs;
  pl->n_pad = (p->n_obs + 3) & ~3;
  pl->mode = p->mode;
  pl->variant = p->variant;
  pl->post_f64 = p->posterior_dtype == CGE_POSTERIOR_F64;
  pl->row_begin = rb;
  pl->row_end = re;
  pl->rows_local = re - rb;
  // columns per warp strip: 32 lanes x Policy::kCols (log-space keeps 8 columns per thread); rows
  // per group: Policy::kRows
  pl->kcols = p->mode == CGE_MODE_LOGSPACE ? 8 : 4;
  pl->krows = p->mode == CGE_MODE_LOGSPACE ? cge::kRowsLog : cge::kRowsProduct;
  const int strip = 32 * pl->kcols;
  pl->ncw = (pl->N + strip - 1) / strip;
  // row groups start at a multiple of krows rows (see run_strip)
  const int span = re - (rb - rb % pl->krows);
  const int groups = (span + pl->krows - 1) / pl->krows;
  pl->groups = groups;
  // fixed-point column sums: at most `groups` addends per column, each a krows-row sum
  // (col_fixed_shift)
  int lc = 0;
  while ((1ll << lc) < (long long)groups + 1) ++lc;
  pl->fixed_bits = 58 - lc - (pl->krows > 4 ? 1 : 0);
  pl->nb_col = (pl->N + 255) / 256;
  pl->nb_row = (pl->rows_local + 31) / 32;
  pl->nb_res = (pl->N + 255) / 256;
  pl->vec4 = (pl->N % 4 == 0) && ((reinterpret_cast<uintptr_t>(post) & 15u) == 0);
  // a shard that fits in L2 with room to spare is stored with the default cache policy, so the
  // scale pass reads it back from there; larger ones stream past L2 (st.global.cs)
  pl->keep_l2 = (size_t)pl->rows_local * pl->N * sizeof(float) * 4 <= ctx->l2_bytes * 3;
  pl->streamed = pl->n_pad > cge::kObsSingleMax;
  pl->reorder = p->mode == CGE_MODE_PRODUCT_F32 && pl->n <= cge::kReorderMax;
  pl->cluster = false;
  pl->nctas = 0;
}

// dynamic shared memory of a likelihood / small-grid launch: header + observation stage + the
// larger of the per-thread float64 slots and the sort scratch of the set-up
size_t like_smem(const Plan &pl, const cge_host::KernelChoice &kc) {
  size_t slots = (size_t)kc.smem_bytes * kc.threads;
  size_t sort = pl.reorder ? (size_t)pl.n_pad * sizeof(float) : 0;
  return cge::kSmemHeader + cge::stage_bytes(pl.n_pad) + (slots > sort ? slots : sort) + 16;
}

// fixes the persistent grid: one CTA per resident slot of the device (occupancy of this kernel at
// this shared-memory size), never more than there are groups for; and the claim size
int size_grid(cge_context *ctx, Plan *pl, const cge_host::KernelChoice &kc, int max_ctas) {
  pl->smem = like_smem(*pl, kc);
  const auto key = std::make_pair(kc.fn, pl->smem);
  if (pl->smem > 40 * 1024) {  // static + dynamic beyond 48 KB needs the opt-in; it only ever grows
    size_t &have = ctx->smem_optin[kc.fn];  // (a smaller value set later would make a cached larger
    if (pl->smem > have) {                  // launch shape invalid)
      CK(cudaFuncSetAttribute(kc.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
      have = pl->smem;
    }
  }
  auto it = ctx->occupancy.find(key);
  if (it == ctx->occupancy.end()) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kc.fn, kc.threads, pl->smem));
    if (per_sm < 1)
      return set_error(CGE_ERR_CUDA, "likelihood kernel does not fit an SM with %zu B of shared memory",
                       pl->smem);
    it = ctx->occupancy.emplace(key, per_sm).first;
  }
  int per_sm = it->second;
  if (ctx->knob_ctas_per_sm > 0 && ctx->knob_ctas_per_sm < per_sm) per_sm = ctx->knob_ctas_per_sm;
  long long P = (long long)ctx->sm_count * per_sm * ctx->knob_grid_mult;
  if (max_ctas > 0 && P > max_ctas) P = max_ctas;
  const long long total = (long long)pl->ncw * pl->groups;
  if (!pl->streamed) {
    // every CTA owns total / P groups (its pool) and its 8 warps take `chunk` of them at a time from
    // a shared-memory counter: ~48 claims per warp, so that the last claim of the slowest warp is a
    // small tail; a claim covers at most 32 rows (run_strip: one range value per lane)
    if (P > total) P = total;
    const int warps = kc.threads / 32;
    long long chunk = total / (P * warps * 48);
    const long long cap = 32 / pl->krows;
    if (chunk < 1) chunk = 1;
    if (chunk > cap) chunk = cap;
    if (ctx->knob_chunk > 0) chunk = ctx->knob_chunk < cap ? ctx->knob_chunk : cap;
    pl->chunk = (int)chunk;
    pl->items_per_strip = 0;
  } else {
    // streamed observations: the CTA is the worker; items of 8 * chunk groups of one strip are
    // claimed from the strip's counter in global memory (a step is hundreds of microseconds there)
    long long chunk = total / (P * 32 * cge::kWarps);
    const long long cap = 32 / pl->krows;
    if (chunk < 1) chunk = 1;
    if (chunk > cap) chunk = cap;
    pl->chunk = (int)chunk;
    pl->items_per_strip = (int)((pl->groups + chunk * cge::kWarps - 1) / (chunk * cge::kWarps));
    const long long items = (long long)pl->ncw * pl->items_per_strip;
    if (P > items) P = items;
  }
  pl->nctas = (int)P;
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// single-launch path for small grids (cge::small_grid_kernel): whole grid, no exchange
// ---------------------------------------------------------------------------------------
constexpr int kSmallGridMax = 256;
constexpr int kZeroCopyObsMax = 1024;  // host observations read in place (mapped pinned) up to this n
int mark(cge_context *ctx, int k, cudaStream_t st);

bool small_grid_eligible(const cge_context *ctx, const cge_params *p, int rb, int re) {
  return p->variant == CGE_VARIANT_DEFAULT && p->grid_size <= kSmallGridMax && rb == 0 &&
         re == p->grid_size && !ctx->peer_on && p->posterior_dtype == CGE_POSTERIOR_F32 &&
         ((p->n_obs + 3) & ~3) <= cge::kObsSingleMax;
}

int reserve_step_buffers(cge_context *ctx, const Plan &pl, cudaStream_t st) {
  const int N = pl.N;
  CKRC(ctx->scale.reserve(1));
  CKRC(ctx->espart.reserve((size_t)pl.nb_res));
  if (!ctx->tickets.ptr) {
    CKRC(ctx->tickets.reserve(3));
    CK(cudaMemsetAsync(ctx->tickets.ptr, 0, 3 * sizeof(unsigned int), st));
  }
  if (!ctx->zbcast.ptr) {
    CKRC(ctx->zbcast.reserve(4));
    CK(cudaMemsetAsync(ctx->zbcast.ptr, 0, 4 * sizeof(unsigned long long), st));
  }
  CKRC(ctx->rowpart.reserve((size_t)pl.rows_local * pl.ncw));
  if ((size_t)N > ctx->colacc.cap) {   // (re)allocated zeroed; every step leaves it zeroed
    CKRC(ctx->colacc.reserve(N));
    CK(cudaMemsetAsync(ctx->colacc.ptr, 0, ctx->colacc.cap * sizeof(unsigned long long), st));
  }
  if ((size_t)pl.ncw > ctx->strip_next.cap) {
    CKRC(ctx->strip_next.reserve(pl.ncw));
    CK(cudaMemsetAsync(ctx->strip_next.ptr, 0, ctx->strip_next.cap * sizeof(int), st));
  }
  CKRC(ctx->partials.reserve((size_t)N + 2));
  CKRC(ctx->rowsum.reserve(pl.rows_local));
  CKRC(ctx->zpart.reserve(pl.nb_row));
  CKRC(ctx->smupart.reserve(pl.nb_row));
  CKRC(ctx->results.reserve(4 + 2 * (size_t)N));
  if (pl.streamed) CKRC(ctx->obs_pad.reserve(pl.n_pad));
  return CGE_OK;
}

int device_absmax(cge_context *ctx, const float *v_dev, size_t count, float *out);

int fill_like_args(cge_context *ctx, const cge_params *p, const Plan &pl, const float *obs_dev,
                   float *posterior_dev, cge::LikeArgs *args) {
  const int N = pl.N;
  *args = cge::LikeArgs{};
  args->obs = pl.streamed ? ctx->obs_pad.ptr : obs_dev;
  args->scale_in = pl.streamed ? ctx->scale.ptr : nullptr;
  args->scale_out = ctx->scale.ptr;
  args->range = cge::RangeArgs{N, p->mu_start, p->mu_end, p->sigma_start, p->sigma_end};
  args->centred = centred_obs(p);
  args->reorder = pl.reorder ? 1 : 0;
  if ((ctx->has_prior_mu || ctx->has_prior_sigma) && ctx->prior_n != N)
    return set_error(CGE_ERR_BAD_ARG, "prior was set for grid_size %d, call has %d", ctx->prior_n, N);
  args->prior_mu = ctx->has_prior_mu ? ctx->prior_mu.ptr : nullptr;
  args->prior_sigma = ctx->has_prior_sigma ? ctx->prior_sigma.ptr : nullptr;
  args->prior_full = ctx->prior_full;
  args->post = posterior_dev;
  args->rowpart = ctx->rowpart.ptr;
  args->colacc = ctx->colacc.ptr;
  args->strip_next = ctx->strip_next.ptr;
  args->tickets = ctx->tickets.ptr;
  args->prior_log2max = 0.0f;
  if (args->prior_mu || args->prior_sigma || args->prior_full) {
    const size_t cells = (size_t)pl.rows_local * N;
    if (args->prior_full && ctx->prior_full_cells != cells) {  // once per table and shard shape
      CKRC(device_absmax(ctx, ctx->prior_full, cells, &ctx->prior_full_max));
      ctx->prior_full_cells = cells;
    }
    const double bound = (args->prior_mu ? (double)ctx->prior_mu_max : 1.0) *
                         (args->prior_sigma ? (double)ctx->prior_sigma_max : 1.0) *
                         (args->prior_full ? (double)ctx->prior_full_max : 1.0);
    args->prior_log2max = bound > 0.0 ? (float)std::ceil(std::log2(bound)) : -1000.0f;
  }
  args->n = pl.n;
  args->n_pad = pl.n_pad;
  args->N = N;
  args->row_begin = pl.row_begin;
  args->row_end = pl.row_end;
  args->groups = pl.groups;
  args->chunk = pl.chunk;
  args->items_per_strip = pl.items_per_strip;
  args->fixed_bits = pl.fixed_bits;
  args->ncw = pl.ncw;
  args->vec4 = pl.vec4 ? 1 : 0;
  args->keep_l2 = pl.keep_l2 ? 1 : 0;
  args->vec8 = (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(posterior_dev) & 31u) == 0) ? 1 : 0;
  args->wide = p->variant == CGE_VARIANT_SLOPE_WIDE ? 1 : p->variant == CGE_VARIANT_SLOPE_TALL ? 2
               : ctx->knob_no_wide ? -1 : ctx->knob_no_tall ? -2 : 0;
#ifdef CGE_DEBUG_TIMES
  CKRC(ctx->debug_times.reserve((size_t)(4 + 32) * pl.nctas));
  CK(cudaMemset(ctx->debug_times.ptr, 0, (size_t)(4 + 32) * pl.nctas * sizeof(unsigned long long)));
  args->debug = ctx->debug_times.ptr;
#endif
  return CGE_OK;
}

// one launch through the extensible API: optional programmatic dependent launch (the kernel may
// be set up while its predecessor in the stream drains; it orders itself with griddepcontrol.wait)
// and optional cluster shape
int launch_ex(const cge_context *ctx, const void *fn, unsigned grid, unsigned block, size_t smem,
              cudaStream_t st, void *arg, bool pdl, unsigned cluster) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (pdl && !ctx->knob_no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  void *args[1] = {arg};
  const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess)
    return set_error(CGE_ERR_CUDA, "kernel launch failed: %s (grid %u, block %u, %zu B dynamic smem, "
                     "cluster %u, pdl %d)", cudaGetErrorString(e), grid, block, smem, cluster,
                     (int)(pdl && !ctx->knob_no_pdl));
  return CGE_OK;
}

// the whole step in one launch; leaves the context as cge_likelihood_partials + cge_finalize would
int small_grid_step(cge_context *ctx, const cge_params *p, const float *obs_dev,
                    float *posterior_dev, cudaStream_t st, bool results_to_host) {
  Plan &pl = ctx->plan;
  pl.valid = false;
  make_plan(ctx, p, 0, p->grid_size, posterior_dev, &pl);
  // at most 8 CTAs (the portable cluster size) -> the grid runs as one thread-block cluster
  pl.cluster = !ctx->knob_no_cluster;
  cge_host::KernelChoice kc;
  if (p->mode == CGE_MODE_PRODUCT_F32) CKRC(cge_host::choose_small_product(pl, &kc));
  else CKRC(cge_host::choose_small_logspace(pl, &kc));
  CKRC(size_grid(ctx, &pl, kc, pl.cluster ? 8 : 0));
  CKRC(reserve_step_buffers(ctx, pl, st));
  cge::SmallArgs a = {};
  CKRC(fill_like_args(ctx, p, pl, obs_dev, posterior_dev, &a.like));
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
  CKRC(launch_ex(ctx, kc.fn, pl.nctas, kc.threads, pl.smem, st, &a, false, pl.cluster ? pl.nctas : 0));
  ctx->launches += 1;
  for (int k = 2; k <= 5; ++k) CKRC(mark(ctx, k, st));
  if (ctx->prof_on && ctx->prof_step < ctx->prof_cap) ctx->prof_step++;
  ctx->kernel = kc;
  pl.valid = true;
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
constexpr int kMarks = 6;
int mark(cge_context *ctx, int k, cudaStream_t st) {
  if (!ctx->prof_on || ctx->prof_step >= ctx->prof_cap) return CGE_OK;
  CK(cudaEventRecord(ctx->prof_events[(size_t)ctx->prof_step * kMarks + k], st));
  return CGE_OK;
}

int prof_reserve(cge_context *ctx, int steps) {
  while ((int)ctx->prof_events.size() < steps * kMarks) {
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
  if (n > 0) CK(cudaEventSynchronize(ctx->prof_events[(size_t)(n - 1) * kMarks + 5]));
  for (int s = 0; s < n; ++s) {
    cudaEvent_t *e = &ctx->prof_events[(size_t)s * kMarks];
    float a = 0, b = 0, c = 0, f = 0, tot = 0;
    CK(cudaEventElapsedTime(&a, e[0], e[1]));
    CK(cudaEventElapsedTime(&b, e[1], e[2]));
    CK(cudaEventElapsedTime(&c, e[2], e[3]));
    CK(cudaEventElapsedTime(&f, e[4], e[5]));
    CK(cudaEventElapsedTime(&tot, e[0], e[5]));
    t.setup_ms += a; t.likelihood_ms += b; t.reduce_ms += c; t.scale_ms += f; t.total_ms += tot;
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
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  // tuning knobs: read once here, never on the step path
  if (const char *v = std::getenv("CGE_CTAS_PER_SM")) ctx->knob_ctas_per_sm = std::atoi(v);
  if (const char *v = std::getenv("CGE_GRID_MULT")) ctx->knob_grid_mult = std::atoi(v) > 0 ? std::atoi(v) : 1;
  if (const char *v = std::getenv("CGE_ITEM_STEPS")) ctx->knob_chunk = std::atoi(v);
  ctx->knob_no_wide = std::getenv("CGE_NO_WIDE") != nullptr;
  ctx->knob_no_tall = std::getenv("CGE_NO_TALL") != nullptr;
  ctx->knob_no_cluster = std::getenv("CGE_NO_CLUSTER") != nullptr;
  ctx->knob_no_pdl = std::getenv("CGE_NO_PDL") != nullptr;
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
  ctx->obs_pad.release(); ctx->obs_in.release(); ctx->espart.release(); ctx->tickets.release(); ctx->zbcast.release();
  ctx->prior_mu.release(); ctx->prior_sigma.release(); ctx->posterior.release();
  ctx->like32.release(); ctx->posterior64.release(); ctx->rowpart.release();
  ctx->colacc.release(); ctx->strip_next.release(); ctx->partials.release(); ctx->rowsum.release(); ctx->zpart.release();
  ctx->smupart.release(); ctx->results.release(); ctx->tmp.release(); ctx->scale.release();
  ctx->sink.release();
  ctx->host_obs.release();
  ctx->host_results.release();
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : ctx->prof_events) cudaEventDestroy(ev);
  for (auto &e : ctx->copy_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
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
// phase 1: fused likelihood (set-up included) + local fold + publish
// ---------------------------------------------------------------------------------------
namespace {
// float64 output: the kernels write float32 likelihoods to context scratch, the finish kernel
// widens and normalises them into the caller's float64 buffer
int like_target(cge_context *ctx, const cge_params *p, int rb, int re, void *posterior_dev,
                float **target) {
  if (p->posterior_dtype == CGE_POSTERIOR_F64) {
    CKRC(ctx->like32.reserve((size_t)(re - rb) * p->grid_size));
    *target = ctx->like32.ptr;
  } else {
    *target = static_cast<float *>(posterior_dev);
  }
  return CGE_OK;
}
}  // namespace

CGE_EXPORT int cge_likelihood_partials(cge_context *ctx, const cge_params *p, const float *obs_dev,
                                       void *posterior_dev, void *stream) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (!obs_dev || !posterior_dev)
    return set_error(CGE_ERR_BAD_ARG, "obs_dev and posterior_dev must be device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  Plan &pl = ctx->plan;
  pl.valid = false;
  ctx->host_results_valid = false;
  const int N = p->grid_size;
  // everything that can fail is checked before anything is launched: a rank that returned an
  // error after publishing nothing would leave its peers polling
  if (ctx->peer_on && (size_t)N + 2 > ctx->peer_cap)
    return set_error(CGE_ERR_BAD_ARG, "grid_size %d exceeds the peer exchange capacity %zu", N,
                     ctx->peer_cap - 2);
  float *like = nullptr;
  CKRC(like_target(ctx, p, rb, re, posterior_dev, &like));
  make_plan(ctx, p, rb, re, like, &pl);
  cge_host::KernelChoice kc;
  if (p->mode == CGE_MODE_PRODUCT_F32) CKRC(cge_host::choose_product(pl, &kc));
  else CKRC(cge_host::choose_logspace(pl, &kc));
  CKRC(size_grid(ctx, &pl, kc, 0));
  CKRC(reserve_step_buffers(ctx, pl, st));
  cge::LikeArgs args;
  CKRC(fill_like_args(ctx, p, pl, obs_dev, like, &args));

  ctx->launches = 0;
  CKRC(mark(ctx, 0, st));
  if (pl.streamed) {  // > 16384 observations: stage them (padded, centred) once in global memory
    cge::setup_kernel<<<1, 256, 0, st>>>(args.range, obs_dev, ctx->obs_pad.ptr, pl.n, pl.n_pad,
                                         args.centred, ctx->scale.ptr);
    CK(cudaGetLastError());
    ctx->launches += 1;
  }
  CKRC(mark(ctx, 1, st));
  CKRC(launch_ex(ctx, kc.fn, pl.nctas, kc.threads, pl.smem, st, &args, true, 0));
  ctx->launches += 1;
  CKRC(mark(ctx, 2, st));

  cge::FoldArgs f = {};
  f.colacc = ctx->colacc.ptr;
  f.strip_next = ctx->strip_next.ptr;
  f.scale = ctx->scale.ptr;
  f.rowpart = ctx->rowpart.ptr;
  f.out = ctx->partials.ptr;
  f.ready_flag = nullptr;
  if (ctx->peer_on) {  // this step's vector goes where the peers can read it
    f.out = peer_buf(ctx, ctx->peer_rank, ctx->peer_step);
    f.ready_flag = peer_flag(ctx, ctx->peer_rank);
    f.ready_value = ctx->peer_step + 1;
  }
  f.rowsum = ctx->rowsum.ptr;
  f.zpart = ctx->zpart.ptr;
  f.smupart = ctx->smupart.ptr;
  f.ticket = ctx->tickets.ptr;
  f.range = args.range;
  f.prior_log2max = args.prior_log2max;
  f.fixed_bits = pl.fixed_bits;
  f.n = pl.n;
  f.ncw = pl.ncw;
  f.rows_local = pl.rows_local;
  f.row_begin = rb;
  f.N = N;
  f.nb_col = pl.nb_col;
  f.nb_row = pl.nb_row;
  CKRC(launch_ex(ctx, (const void *)cge::fold_publish_kernel, pl.nb_col + pl.nb_row, 256, 0, st, &f,
                 true, 0));
  ctx->launches += 1;
  CKRC(mark(ctx, 3, st));
  ctx->kernel = kc;
  pl.valid = true;
  return CGE_OK;
}

// optional non-flat prior for the following calls (NULL, NULL, n, NULL restores the flat prior).
// The largest |factor| of each piece is kept: the fixed-point column sums are scaled from a bound
// on the largest cell (col_fixed_shift), and the prior multiplies the cells.
namespace {
int device_absmax(cge_context *ctx, const float *v_dev, size_t count, float *out) {
  CKRC(ctx->tmp.reserve(2));
  unsigned int *slot = reinterpret_cast<unsigned int *>(ctx->tmp.ptr);
  CK(cudaMemsetAsync(slot, 0, sizeof(unsigned int), ctx->stream));
  size_t blocks = (count + 1023) / 1024;
  if (blocks > (size_t)ctx->sm_count * 8) blocks = (size_t)ctx->sm_count * 8;
  cge::absmax_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(v_dev, count, slot);
  CK(cudaGetLastError());
  unsigned int bits = 0;
  CK(cudaMemcpyAsync(&bits, slot, sizeof bits, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::memcpy(out, &bits, sizeof bits);
  return CGE_OK;
}
}  // namespace

CGE_EXPORT int cge_set_prior(cge_context *ctx, const float *prior_mu, const float *prior_sigma,
                             int grid_size, const float *prior_full_dev) {
  CKRC(bind(ctx));
  if ((prior_mu || prior_sigma) && grid_size < 1)
    return set_error(CGE_ERR_BAD_ARG, "set_prior: grid_size must be >= 1");
  if (prior_full_dev && !is_device_pointer(prior_full_dev))
    return set_error(CGE_ERR_BAD_ARG, "set_prior: the full prior table must be device memory");
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->has_prior_mu = ctx->has_prior_sigma = false;
  ctx->prior_n = grid_size;
  ctx->prior_mu_max = ctx->prior_sigma_max = 1.0f;
  if (prior_mu) {
    CKRC(ctx->prior_mu.reserve(grid_size));
    CK(cudaMemcpy(ctx->prior_mu.ptr, prior_mu, (size_t)grid_size * sizeof(float), cudaMemcpyDefault));
    CKRC(device_absmax(ctx, ctx->prior_mu.ptr, grid_size, &ctx->prior_mu_max));
    ctx->has_prior_mu = true;
  }
  if (prior_sigma) {
    CKRC(ctx->prior_sigma.reserve(grid_size));
    CK(cudaMemcpy(ctx->prior_sigma.ptr, prior_sigma, (size_t)grid_size * sizeof(float),
                  cudaMemcpyDefault));
    CKRC(device_absmax(ctx, ctx->prior_sigma.ptr, grid_size, &ctx->prior_sigma_max));
    ctx->has_prior_sigma = true;
  }
  ctx->prior_full = prior_full_dev;
  ctx->prior_full_cells = 0;  // its largest factor is measured at the first call that knows its shape
  ctx->prior_full_max = 1.0f;
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
// phase 2: (peer sum) + results + scale, one launch
// ---------------------------------------------------------------------------------------
// `scale_cells`: how many cells of the shard this launch scales (0 = all of them; the chunked
// host-posterior path of cge_grid_posterior scales the rest itself)
static int finalize_impl(cge_context *ctx, const cge_params *p, void *posterior_dev, void *stream,
                         size_t scale_cells) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  const Plan &pl = ctx->plan;
  if (!pl.valid || pl.N != p->grid_size || pl.row_begin != rb || pl.row_end != re ||
      pl.post_f64 != (p->posterior_dtype == CGE_POSTERIOR_F64))
    return set_error(CGE_ERR_BAD_ARG, "cge_finalize does not match the last cge_likelihood_partials");
  if (!posterior_dev) return set_error(CGE_ERR_BAD_ARG, "posterior_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl.N;
  CKRC(mark(ctx, 4, st));
  cge::FinishArgs a = {};
  a.peers.world = 0;
  if (ctx->peer_on) {
    a.peers.world = ctx->peer_world;
    a.peers.want = ctx->peer_step + 1;
    for (int r = 0; r < ctx->peer_world; ++r) {
      a.peers.buf[r] = peer_buf(ctx, r, ctx->peer_step);
      a.peers.flag[r] = peer_flag(ctx, r);
    }
  }
  a.partials = ctx->partials.ptr;
  a.rowsum = ctx->rowsum.ptr;
  a.results = ctx->results.ptr;
  a.results_host = nullptr;
  if (ctx->want_host_results) {  // cge_grid_posterior with host outputs: written in place over PCIe
    CKRC(ctx->host_results.reserve(4 + 2 * (size_t)N));
    a.results_host = ctx->host_results.ptr;
    ctx->host_results_valid = true;
  }
  a.espart = ctx->espart.ptr;
  a.ticket = ctx->tickets.ptr + 1;
  a.bcast = ctx->zbcast.ptr;
  a.cells = (size_t)pl.rows_local * N;
  if (scale_cells > 0 && scale_cells < a.cells) a.cells = scale_cells;
  if (pl.post_f64) {
    a.post = ctx->like32.ptr;
    a.post64 = static_cast<double *>(posterior_dev);
    a.aligned = ((reinterpret_cast<uintptr_t>(a.post) | reinterpret_cast<uintptr_t>(a.post64)) & 15u) == 0;
  } else {
    a.post = static_cast<float *>(posterior_dev);
    a.post64 = nullptr;
    a.aligned = (reinterpret_cast<uintptr_t>(a.post) & 15u) == 0;
  }
  a.range = cge::RangeArgs{N, p->mu_start, p->mu_end, p->sigma_start, p->sigma_end};
  a.N = N;
  a.row_begin = rb;
  a.row_end = re;
  a.nb_res = pl.nb_res;
  // enough CTAs to keep every SM's memory pipeline full, capped so tiny grids stay tiny; never
  // fewer than the blocks that write the marginals
  size_t want = a.aligned ? (a.cells / 4 + 1023) / 1024 : (a.cells + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  if (want > cap) want = cap;
  if (want < (size_t)pl.nb_res) want = pl.nb_res;
  CKRC(launch_ex(ctx, (const void *)cge::finish_kernel, (unsigned)want, 256, 0, st, &a, true, 0));
  if (ctx->peer_on) ctx->peer_step++;
  ctx->launches += 1;
  CKRC(mark(ctx, 5, st));
  if (ctx->prof_on && ctx->prof_step < ctx->prof_cap) ctx->prof_step++;
  return CGE_OK;
}

CGE_EXPORT int cge_finalize(cge_context *ctx, const cge_params *p, void *posterior_dev,
                            void *stream) {
  return finalize_impl(ctx, p, posterior_dev, stream, 0);
}

// Phase 2 with the posterior going to host memory: the shard is scaled in row chunks of ~64 MB
// (the first by the finish kernel, the others by scale_chunk_kernel) and each chunk's PCIe copy,
// on a second stream, overlaps the scale pass of the next ones.
static int finalize_to_host(cge_context *ctx, const cge_params *p, void *post_dev, void *post_host,
                            cudaStream_t st) {
  const Plan &pl = ctx->plan;
  const size_t N = pl.N, rows = pl.rows_local, cell_bytes = pl.post_f64 ? sizeof(double) : sizeof(float);
  size_t chunk_rows = ((size_t)64 << 20) / (N * cell_bytes);
  if (chunk_rows < 1) chunk_rows = 1;
  if ((rows + chunk_rows - 1) / chunk_rows > kCopyChunksMax) chunk_rows = (rows + kCopyChunksMax - 1) / kCopyChunksMax;
  if (chunk_rows >= rows) {  // one chunk: nothing to overlap
    CKRC(finalize_impl(ctx, p, post_dev, st, 0));
    CK(cudaMemcpyAsync(post_host, post_dev, rows * N * cell_bytes, cudaMemcpyDeviceToHost, st));
    return CGE_OK;
  }
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto &e : ctx->copy_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  float *src32 = pl.post_f64 ? ctx->like32.ptr : static_cast<float *>(post_dev);
  double *dst64 = pl.post_f64 ? static_cast<double *>(post_dev) : nullptr;
  int k = 0;
  for (size_t r0 = 0; r0 < rows; r0 += chunk_rows, ++k) {
    const size_t nr = rows - r0 < chunk_rows ? rows - r0 : chunk_rows, off = r0 * N, cells = nr * N;
    if (k == 0) {
      CKRC(finalize_impl(ctx, p, post_dev, st, cells));
    } else {
      cge::ScaleChunkArgs a = {};
      a.results = ctx->results.ptr;
      a.post = src32 + off;
      a.post64 = dst64 ? dst64 + off : nullptr;
      a.cells = cells;
      a.aligned = ((reinterpret_cast<uintptr_t>(a.post) | reinterpret_cast<uintptr_t>(a.post64)) & 15u) == 0;
      size_t want = a.aligned ? (cells / 4 + 1023) / 1024 : (cells + 255) / 256;
      const size_t cap = (size_t)ctx->sm_count * 16;
      if (want > cap) want = cap;
      CKRC(launch_ex(ctx, (const void *)cge::scale_chunk_kernel, (unsigned)want, 256, 0, st, &a, true, 0));
      ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->copy_ev[k], st));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[k], 0));
    CK(cudaMemcpyAsync(static_cast<char *>(post_host) + off * cell_bytes,
                       static_cast<char *>(post_dev) + off * cell_bytes, cells * cell_bytes,
                       cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  CK(cudaEventRecord(ctx->copy_ev[kCopyChunksMax], ctx->copy_stream));
  CK(cudaStreamWaitEvent(st, ctx->copy_ev[kCopyChunksMax], 0));
  return CGE_OK;
}

// ---------------------------------------------------------------------------------------
// peer exchange: CUDA IPC (one process per GPU, one node) or plain peer pointers (one process)
// ---------------------------------------------------------------------------------------
namespace {
// the finish kernel's local {flag, Z, Smu} word counts steps like peer_step: restart it with it
void reset_bcast(cge_context *ctx) {
  if (!ctx->zbcast.ptr) return;
  int cur = -1;
  cudaGetDevice(&cur);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaMemset(ctx->zbcast.ptr, 0, 4 * sizeof(unsigned long long));
  if (cur >= 0) cudaSetDevice(cur);
}
int peer_release(cge_context *ctx) {
  for (int r = 0; r < cge::kMaxPeers; ++r) {
    if (!ctx->peer_block[r]) continue;
    if (r == ctx->peer_rank || !ctx->peer_on) cudaFree(ctx->peer_block[r]);  // (parked at slot 0 before connect)
    else if (ctx->peer_ipc) cudaIpcCloseMemHandle(ctx->peer_block[r]);
    ctx->peer_block[r] = nullptr;
  }
  ctx->peer_on = false;
  ctx->peer_ipc = false;
  ctx->peer_world = 1;
  ctx->peer_rank = 0;
  ctx->peer_cap = 0;
  ctx->peer_step = 0;
  reset_bcast(ctx);  // the local flag counts steps too
  return CGE_OK;
}

// this context's exchange block, parked at slot 0 until the rank is known
int peer_alloc(cge_context *ctx, int max_grid_size) {
  CK(cudaStreamSynchronize(ctx->stream));
  peer_release(ctx);
  ctx->peer_cap = (((size_t)max_grid_size + 2) + 31) & ~(size_t)31;
  const size_t bytes = kPeerHeader + 2 * ctx->peer_cap * sizeof(double);
  unsigned char *block = nullptr;
  CK(cudaMalloc(&block, bytes));
  CK(cudaMemset(block, 0, bytes));
  CK(cudaDeviceSynchronize());
  ctx->peer_block[0] = block;
  return CGE_OK;
}
}  // namespace

CGE_EXPORT int cge_peer_local_handle(cge_context *ctx, void *handle, int max_grid_size) {
  CKRC(bind(ctx));
  static_assert(sizeof(cudaIpcMemHandle_t) == CGE_IPC_HANDLE_BYTES, "IPC handle size");
  if (!handle || max_grid_size < 1) return set_error(CGE_ERR_BAD_ARG, "peer_local_handle: bad argument");
  CKRC(peer_alloc(ctx, max_grid_size));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->peer_block[0]));
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
  ctx->peer_on = true;   // (from here peer_release frees only slot `rank` and closes the others)
  ctx->peer_ipc = true;
  const unsigned char *hb = static_cast<const unsigned char *>(handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + (size_t)r * sizeof h, sizeof h);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      peer_release(ctx);
      return set_error(CGE_ERR_PEER, "cudaIpcOpenMemHandle for rank %d failed: %s", r,
                       cudaGetErrorString(e));
    }
    ctx->peer_block[r] = static_cast<unsigned char *>(p);
  }
  ctx->peer_world = world;
  ctx->peer_step = 0;
  reset_bcast(ctx);  // the local flag counts steps too
  return CGE_OK;
}

CGE_EXPORT int cge_peer_disconnect(cge_context *ctx) {
  CKRC(bind(ctx));
  CK(cudaDeviceSynchronize());
  return peer_release(ctx);
}

// ---- same-process peers (cge_multi.cu): the contexts of one cge_multi exchange plain pointers
extern "C" int cge_internal_peer_alloc(cge_context *ctx, int max_grid_size, unsigned char **block) {
  CKRC(bind(ctx));
  CKRC(peer_alloc(ctx, max_grid_size));
  *block = ctx->peer_block[0];
  return CGE_OK;
}

extern "C" int cge_internal_peer_attach(cge_context *ctx, int rank, int world,
                                        unsigned char *const *blocks) {
  CKRC(bind(ctx));
  if (world < 2 || world > cge::kMaxPeers || rank < 0 || rank >= world || !ctx->peer_block[0])
    return set_error(CGE_ERR_BAD_ARG, "peer_attach: bad argument");
  unsigned char *mine = ctx->peer_block[0];
  ctx->peer_block[0] = nullptr;
  for (int r = 0; r < world; ++r) ctx->peer_block[r] = blocks[r];
  if (ctx->peer_block[rank] != mine) return set_error(CGE_ERR_BAD_ARG, "peer_attach: block mismatch");
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  ctx->peer_step = 0;
  reset_bcast(ctx);  // the local flag counts steps too
  ctx->peer_ipc = false;
  ctx->peer_on = true;
  return CGE_OK;
}

// the context stops owning its peers' blocks (their owners free them) before it is destroyed
extern "C" int cge_internal_peer_detach(cge_context *ctx) {
  CKRC(bind(ctx));
  CK(cudaDeviceSynchronize());
  for (int r = 0; r < cge::kMaxPeers; ++r)
    if (r != ctx->peer_rank) ctx->peer_block[r] = nullptr;
  return peer_release(ctx);
}

// all contexts of a cge_multi whose devices repeat share one stream (see cge_multi.cu)
extern "C" int cge_internal_share_stream(cge_context *ctx, void *stream) {
  CKRC(bind(ctx));
  if (ctx->stream && ctx->own_stream) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamDestroy(ctx->stream));
  }
  ctx->stream = (cudaStream_t)stream;
  ctx->own_stream = false;
  return CGE_OK;
}

// cge_multi.cu, host outputs: `on` makes the finish kernel of the following cge_finalize write the
// result block into this context's pinned staging block as well (what cge_grid_posterior does for
// host destinations: no device-to-host copy afterwards); `block` receives that staging block once
// such a finish kernel has been issued (valid after the stream has been synchronised)
extern "C" int cge_internal_host_results(cge_context *ctx, int on, const double **block) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  ctx->want_host_results = on != 0;
  if (block) *block = ctx->host_results_valid ? ctx->host_results.ptr : nullptr;
  return CGE_OK;
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
                                  void *posterior, double *marg_mu, double *marg_sigma,
                                  double *expectations, cge_timing *timing) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (!obs) return set_error(CGE_ERR_BAD_ARG, "obs is NULL");
  cudaStream_t st = ctx->stream;
  const size_t cells = (size_t)(re - rb) * p->grid_size;
  const bool f64 = p->posterior_dtype == CGE_POSTERIOR_F64;
  const size_t cell_bytes = f64 ? sizeof(double) : sizeof(float);
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
    if (p->n_obs <= kZeroCopyObsMax && !(((p->n_obs + 3) & ~3) > cge::kObsSingleMax)) {
      obs_dev = ctx->host_obs.ptr;  // mapped pinned memory: the kernels read it in place
    } else {
      CKRC(ctx->obs_in.reserve(p->n_obs));
      CK(cudaMemcpyAsync(ctx->obs_in.ptr, ctx->host_obs.ptr, (size_t)p->n_obs * sizeof(float),
                         cudaMemcpyHostToDevice, st));
      obs_dev = ctx->obs_in.ptr;
    }
  }
  if (timing) CK(cudaEventRecord(ctx->ev[0], st));
  void *post_dev = posterior;
  const bool post_to_host = posterior && !is_device_pointer(posterior);
  if (!posterior || post_to_host) {
    if (f64) {
      CKRC(ctx->posterior64.reserve(cells));
      post_dev = ctx->posterior64.ptr;
    } else {
      CKRC(ctx->posterior.reserve(cells));
      post_dev = ctx->posterior.ptr;
    }
  }
  int rc;
  if (small) {
    const bool host_out = !(marg_mu && is_device_pointer(marg_mu)) &&
                          !(marg_sigma && is_device_pointer(marg_sigma)) &&
                          !(expectations && is_device_pointer(expectations));
    rc = small_grid_step(ctx, p, obs_dev, static_cast<float *>(post_dev), st, host_out);
  } else {
    rc = cge_likelihood_partials(ctx, p, obs_dev, post_dev, st);
    if (rc == CGE_OK) {
      // host destinations for the results: the finish kernel writes them into the mapped staging
      // block itself (no D2H copy after it)
      ctx->want_host_results = !(marg_mu && is_device_pointer(marg_mu)) &&
                               !(marg_sigma && is_device_pointer(marg_sigma)) &&
                               !(expectations && is_device_pointer(expectations));
      rc = post_to_host ? finalize_to_host(ctx, p, post_dev, posterior, st) : cge_finalize(ctx, p, post_dev, st);
      ctx->want_host_results = false;
    }
  }
  if (rc != CGE_OK) {
    if (!was_on) ctx->prof_on = false;
    return rc;
  }
  if (timing) CK(cudaEventRecord(ctx->ev[1], st));
  if (post_to_host && small)
    CK(cudaMemcpyAsync(posterior, post_dev, cells * cell_bytes, cudaMemcpyDeviceToHost, st));
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
                                        void *posterior_dev, void *stream) {
  CKRC(bind(ctx));
  int rb, re;
  CKRC(check_params(p, &rb, &re));
  if (rb != 0 || re != p->grid_size)
    return set_error(CGE_ERR_BAD_ARG, "cge_grid_posterior_async needs the whole grid; a shard [%d, %d) "
                     "goes through cge_likelihood_partials / cge_finalize", rb, re);
  if (!obs_dev || !posterior_dev || !is_device_pointer(obs_dev) || !is_device_pointer(posterior_dev))
    return set_error(CGE_ERR_BAD_ARG, "cge_grid_posterior_async needs device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  if (small_grid_eligible(ctx, p, rb, re))
    return small_grid_step(ctx, p, obs_dev, static_cast<float *>(posterior_dev), st, false);
  CKRC(cge_likelihood_partials(ctx, p, obs_dev, posterior_dev, st));
  return cge_finalize(ctx, p, posterior_dev, st);
}

// posterior kept by the context when cge_grid_posterior was called with posterior == NULL or a
// host pointer (float32 or float64 according to that call's posterior_dtype)
CGE_EXPORT int cge_posterior_dev(cge_context *ctx, void **posterior_dev, size_t *count) {
  if (!ctx) return set_error(CGE_ERR_BAD_ARG, "context is NULL");
  if (posterior_dev)
    *posterior_dev = ctx->plan.post_f64 ? (void *)ctx->posterior64.ptr : (void *)ctx->posterior.ptr;
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
  const double v[15] = {h.log2s, h.shift, h.m2, h.xbar, h.ss, h.smin, h.mu_c, h.s_c, h.umax,
                        cge::kPairedUmax, cge::kPairedUmax3, h.ustep, h.seps, cge::kSlopeErr, h.wide};
  for (int i = 0; i < count && i < 15; ++i) out[i] = v[i];
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
// `count` doubles of context-owned device scratch (grow-only; valid until the next call that asks)
extern "C" int cge_internal_scratch(cge_context *ctx, size_t count, double **ptr) {
  CKRC(bind(ctx));
  CKRC(ctx->tmp.reserve(count));
  *ptr = ctx->tmp.ptr;
  return CGE_OK;
}
#ifdef CGE_DEBUG_TIMES
// diagnostic build only: per-CTA {start ns, end ns, smid, set-up done ns} of the last likelihood
// launch, followed by the end time of each of its warps (32 slots per CTA, unused ones zero);
// `out` holds 36 * max_ctas
CGE_EXPORT int cge_debug_times(cge_context *ctx, unsigned long long *out, int max_ctas, int *nctas) {
  CKRC(bind(ctx));
  CK(cudaDeviceSynchronize());
  if (ctx->plan.nctas > max_ctas) return set_error(CGE_ERR_BAD_ARG, "debug_times: buffer too small");
  const int n = ctx->plan.nctas;
  CK(cudaMemcpy(out, ctx->debug_times.ptr, (size_t)(4 + 32) * n * sizeof(unsigned long long),
                cudaMemcpyDeviceToHost));
  *nctas = n;
  return CGE_OK;
}
#endif

// ---------------------------------------------------------------------------------------
// pipe-peak micro-benchmarks
// ---------------------------------------------------------------------------------------
CGE_EXPORT int cge_microbench(cge_context *ctx, int kind, double *ops_per_second) {
  CKRC(bind(ctx));
  if (!ops_per_second || kind < 0 || kind > 12)
    return set_error(CGE_ERR_BAD_ARG, "microbench: bad argument");
  CKRC(ctx->sink.reserve(4));
  const int iters = 4096, chains = 8;
  // kinds 8, 9, 10: the 6 FFMA2 + 2 MUFU mix at 2, 3 and 4 CTAs per SM (8, 12, 16 warps per SM)
  // kinds 11, 12: the 10 FFMA2 + 2 MUFU mix at 64 and at 24 resident warps per SM
  const int blocks = ctx->sm_count * (kind == 8 ? 2 : kind == 9 || kind == 12 ? 3 : kind == 10 ? 4 : 8), threads = 256;
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
      case 11: case 12: cge::microbench_kernel<11><<<blocks, threads, 0, st>>>(ctx->sink.ptr, iters, 0.5f); break;
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
  static const double per_iter[13] = {8, 8, 8, 16, 4 * 2 + 4, 4 * 2 + 8, 6 * 2 + 2, 4 * 2 + 4 + 2,
                                      6 * 2 + 2, 6 * 2 + 2, 6 * 2 + 2, 10 * 2 + 2, 10 * 2 + 2};
  (void)chains;
  double ops = (double)blocks * threads * (double)iters * per_iter[kind];
  *ops_per_second = ops / (best * 1e-3);
  return CGE_OK;
}
