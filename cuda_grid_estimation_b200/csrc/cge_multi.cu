// cge_multi.cu -- several GPUs driven from ONE process through the C ABI (include/cge.h,
// cge_create_multi / cge_grid_posterior_multi): the single-process counterpart of the
// one-process-per-GPU phase API.  The reference's caller is a single-process main()
// (cpp/grid-algorithm/src/grid.cu:19-142); this is what lets such a caller use a whole 8-GPU box
// with one call.
//
// One cge_context per listed device; mu rows are sharded contiguously in list order; per step the
// calling thread issues likelihood + fold on every device, then the finish kernel on every device.
// The exchange is the same as in the IPC case -- each finish kernel polls the other devices'
// "vector complete" flags and reads their N+2 float64 over NVLink, adding them in list order --
// but with plain peer pointers (cudaDeviceEnablePeerAccess), no IPC handles.  Nothing else
// crosses NVLink; the observations go to every device (a few hundred bytes, read in place from
// portable pinned memory when small).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "cge_host.cuh"

#define CGE_EXPORT extern "C" __attribute__((visibility("default")))

extern "C" int cge_internal_stream(cge_context *ctx, void **stream, int *sm_count);
extern "C" int cge_internal_peer_alloc(cge_context *ctx, int max_grid_size, unsigned char **block);
extern "C" int cge_internal_peer_attach(cge_context *ctx, int rank, int world,
                                        unsigned char *const *blocks);
extern "C" int cge_internal_peer_detach(cge_context *ctx);
extern "C" int cge_internal_share_stream(cge_context *ctx, void *stream);
extern "C" int cge_internal_host_results(cge_context *ctx, int on, const double **block);

using cge_host::set_error;

struct cge_multi {
  int n = 0;
  std::vector<int> devices;
  std::vector<cge_context *> ctx;
  std::vector<cudaStream_t> stream;  // each context's stream (one shared stream when devices repeat)
  bool shared_stream = false;
  cudaStream_t shared = nullptr;
  int shared_device = 0;
  int peer_grid = 0;                 // grid size the exchange blocks were sized for (0: none yet)
  std::vector<int> rb, re;           // shard rows of the last call
  std::vector<void *> post;          // posterior shards (device), grow-only
  std::vector<size_t> post_bytes;
  std::vector<float *> obs_dev;      // observations on each device, grow-only
  std::vector<size_t> obs_cap;
  float *obs_pinned = nullptr;       // portable mapped pinned staging of the observations
  size_t obs_pinned_cap = 0;
  std::vector<const double *> res_host;  // each context's pinned result block {head[4],
                                         // marg_sigma[N], marg_mu[N]}, written by its finish kernel
  std::vector<cudaEvent_t> ev0, ev1;
  int last_dtype = 0, last_N = 0;
};

namespace {

constexpr int kZeroCopyObsMax = 1024;

void shard_rows(int N, int world, int rank, int *b, int *e) {
  const int base = N / world, rem = N % world;
  *b = rank * base + (rank < rem ? rank : rem);
  *e = *b + base + (rank < rem ? 1 : 0);
}

int release_peers(cge_multi *m) {
  if (!m->peer_grid) return CGE_OK;
  for (int i = 0; i < m->n; ++i) CKRC(cge_internal_peer_detach(m->ctx[i]));
  m->peer_grid = 0;
  return CGE_OK;
}

// (re)allocate every device's exchange block for grids up to N and cross-map them
int connect_peers(cge_multi *m, int N) {
  if (m->n < 2 || N <= m->peer_grid) return CGE_OK;
  CKRC(release_peers(m));
  std::vector<unsigned char *> blocks(m->n, nullptr);
  for (int i = 0; i < m->n; ++i) CKRC(cge_internal_peer_alloc(m->ctx[i], N, &blocks[i]));
  for (int i = 0; i < m->n; ++i) CKRC(cge_internal_peer_attach(m->ctx[i], i, m->n, blocks.data()));
  m->peer_grid = N;
  return CGE_OK;
}

int check_multi_params(const cge_multi *m, const cge_params *p) {
  if (!m) return set_error(CGE_ERR_BAD_ARG, "multi context is NULL");
  if (!p) return set_error(CGE_ERR_BAD_ARG, "params is NULL");
  if (p->row_begin != 0 || (p->row_end != 0 && p->row_end != p->grid_size))
    return set_error(CGE_ERR_BAD_ARG, "cge_grid_posterior_multi shards the whole grid itself: "
                                      "row_begin/row_end must be 0");
  if (p->grid_size < m->n)
    return set_error(CGE_ERR_BAD_ARG, "grid_size %d < %d GPUs: a shard would be empty", p->grid_size,
                     m->n);
  return CGE_OK;
}

// shard rows + posterior buffers for this call
int prepare(cge_multi *m, const cge_params *p) {
  const size_t cell = p->posterior_dtype == CGE_POSTERIOR_F64 ? sizeof(double) : sizeof(float);
  for (int i = 0; i < m->n; ++i) {
    shard_rows(p->grid_size, m->n, i, &m->rb[i], &m->re[i]);
    const size_t need = (size_t)(m->re[i] - m->rb[i]) * p->grid_size * cell;
    if (need > m->post_bytes[i]) {
      CK(cudaSetDevice(m->devices[i]));
      if (m->post[i]) CK(cudaFree(m->post[i]));
      m->post[i] = nullptr;
      m->post_bytes[i] = 0;
      CK(cudaMalloc(&m->post[i], need + need / 8));
      m->post_bytes[i] = need + need / 8;
    }
  }
  m->last_dtype = p->posterior_dtype;
  m->last_N = p->grid_size;
  return connect_peers(m, p->grid_size);
}

// both phases on every device; obs[i] is readable from device i.  host_results: every finish
// kernel also writes its result block into its context's pinned staging block (m->res_host[i])
int issue_step(cge_multi *m, const cge_params *p, const float *const *obs, bool host_results) {
  cge_params q = *p;
  for (int i = 0; i < m->n; ++i) {  // phase 1 everywhere first: every flag is on its way before any
    q.row_begin = m->rb[i];         // finish kernel starts to poll
    q.row_end = m->re[i];
    CKRC(cge_likelihood_partials(m->ctx[i], &q, obs[i], m->post[i], m->stream[i]));
  }
  for (int i = 0; i < m->n; ++i) {
    q.row_begin = m->rb[i];
    q.row_end = m->re[i];
    if (host_results) CKRC(cge_internal_host_results(m->ctx[i], 1, nullptr));
    const int rc = cge_finalize(m->ctx[i], &q, m->post[i], m->stream[i]);
    if (host_results) cge_internal_host_results(m->ctx[i], 0, &m->res_host[i]);
    if (rc != CGE_OK) return rc;
    if (host_results && !m->res_host[i])
      return set_error(CGE_ERR_CUDA, "shard %d: the finish kernel has no host result block", i);
  }
  return CGE_OK;
}

}  // namespace

CGE_EXPORT int cge_create_multi(int n_gpus, const int *devices, cge_multi **out) {
  if (!out) return set_error(CGE_ERR_BAD_ARG, "out is NULL");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > cge::kMaxPeers)
    return set_error(CGE_ERR_BAD_ARG, "n_gpus must be 1..%d (got %d)", cge::kMaxPeers, n_gpus);
  cge_multi *m = new (std::nothrow) cge_multi();
  if (!m) return set_error(CGE_ERR_CUDA, "out of host memory");
  m->n = n_gpus;
  m->devices.resize(n_gpus);
  for (int i = 0; i < n_gpus; ++i) m->devices[i] = devices ? devices[i] : i;
  for (int i = 0; i < n_gpus; ++i)
    for (int j = 0; j < i; ++j)
      if (m->devices[i] == m->devices[j]) m->shared_stream = true;
  m->ctx.assign(n_gpus, nullptr);
  m->stream.assign(n_gpus, nullptr);
  m->rb.assign(n_gpus, 0);
  m->re.assign(n_gpus, 0);
  m->post.assign(n_gpus, nullptr);
  m->post_bytes.assign(n_gpus, 0);
  m->obs_dev.assign(n_gpus, nullptr);
  m->obs_cap.assign(n_gpus, 0);
  m->res_host.assign(n_gpus, nullptr);
  m->ev0.assign(n_gpus, nullptr);
  m->ev1.assign(n_gpus, nullptr);
  auto fail = [&](int rc) {
    cge_destroy_multi(m);
    return rc;
  };
  for (int i = 0; i < n_gpus; ++i) {
    int rc = cge_create(m->devices[i], &m->ctx[i]);
    if (rc != CGE_OK) return fail(rc);
  }
  // peer access between every pair of distinct devices
  for (int i = 0; i < n_gpus; ++i)
    for (int j = 0; j < n_gpus; ++j) {
      const int di = m->devices[i], dj = m->devices[j];
      if (di == dj) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, di, dj) != cudaSuccess || !can) {
        cudaGetLastError();
        return fail(set_error(CGE_ERR_PEER, "device %d cannot access device %d (no peer access)", di, dj));
      }
      cudaSetDevice(di);
      cudaError_t e = cudaDeviceEnablePeerAccess(dj, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess)
        return fail(set_error(CGE_ERR_PEER, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", di, dj,
                              cudaGetErrorString(e)));
    }
  if (m->shared_stream) {
    // a device is listed more than once: every shard runs on ONE stream in issue order, so a
    // finish kernel never waits for a flag whose fold kernel is queued behind it
    for (int i = 1; i < n_gpus; ++i)
      if (m->devices[i] != m->devices[0])
        return fail(set_error(CGE_ERR_BAD_ARG, "a repeated device must be the only device of the group"));
    m->shared_device = m->devices[0];
    cudaSetDevice(m->shared_device);
    cudaError_t e = cudaStreamCreateWithFlags(&m->shared, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(set_error(CGE_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)));
    for (int i = 0; i < n_gpus; ++i) {
      int rc = cge_internal_share_stream(m->ctx[i], m->shared);
      if (rc != CGE_OK) return fail(rc);
    }
  }
  for (int i = 0; i < n_gpus; ++i) {
    void *st = nullptr;
    int rc = cge_internal_stream(m->ctx[i], &st, nullptr);
    if (rc != CGE_OK) return fail(rc);
    m->stream[i] = (cudaStream_t)st;
    cudaSetDevice(m->devices[i]);
    if (cudaEventCreate(&m->ev0[i]) != cudaSuccess || cudaEventCreate(&m->ev1[i]) != cudaSuccess)
      return fail(set_error(CGE_ERR_CUDA, "cudaEventCreate failed"));
  }
  *out = m;
  return CGE_OK;
}

CGE_EXPORT int cge_destroy_multi(cge_multi *m) {
  if (!m) return CGE_OK;
  for (int i = 0; i < m->n; ++i) {
    if (!m->ctx[i]) continue;
    cudaSetDevice(m->devices[i]);
    cudaDeviceSynchronize();
  }
  if (m->peer_grid)
    for (int i = 0; i < m->n; ++i)
      if (m->ctx[i]) cge_internal_peer_detach(m->ctx[i]);
  for (int i = 0; i < m->n; ++i) {
    if (!m->ctx[i]) continue;  // (creation failed before this entry: nothing was allocated on it)
    cudaSetDevice(m->devices[i]);
    if (m->post[i]) cudaFree(m->post[i]);
    if (m->obs_dev[i]) cudaFree(m->obs_dev[i]);
    if (m->ev0[i]) cudaEventDestroy(m->ev0[i]);
    if (m->ev1[i]) cudaEventDestroy(m->ev1[i]);
    if (m->ctx[i]) cge_destroy(m->ctx[i]);
  }
  if (m->shared) {
    cudaSetDevice(m->shared_device);
    cudaStreamDestroy(m->shared);
  }
  if (m->obs_pinned) cudaFreeHost(m->obs_pinned);
  cudaGetLastError();
  delete m;
  return CGE_OK;
}

CGE_EXPORT int cge_multi_size(cge_multi *m, int *n_gpus) {
  if (!m || !n_gpus) return set_error(CGE_ERR_BAD_ARG, "NULL argument");
  *n_gpus = m->n;
  return CGE_OK;
}

CGE_EXPORT int cge_multi_shard(cge_multi *m, int index, int *device, int *row_begin, int *row_end,
                               void **posterior_dev) {
  if (!m || index < 0 || index >= m->n) return set_error(CGE_ERR_BAD_ARG, "multi_shard: bad index");
  if (device) *device = m->devices[index];
  if (row_begin) *row_begin = m->rb[index];
  if (row_end) *row_end = m->re[index];
  if (posterior_dev) *posterior_dev = m->post[index];
  return CGE_OK;
}

CGE_EXPORT int cge_multi_step_async(cge_multi *m, const cge_params *p, const float *const *obs_dev) {
  CKRC(check_multi_params(m, p));
  if (!obs_dev) return set_error(CGE_ERR_BAD_ARG, "obs_dev is NULL");
  CKRC(prepare(m, p));
  return issue_step(m, p, obs_dev, false);
}

CGE_EXPORT int cge_multi_synchronize(cge_multi *m) {
  if (!m) return set_error(CGE_ERR_BAD_ARG, "multi context is NULL");
  for (int i = 0; i < m->n; ++i) {
    CK(cudaSetDevice(m->devices[i]));
    CK(cudaStreamSynchronize(m->stream[i]));
  }
  return CGE_OK;
}

CGE_EXPORT int cge_grid_posterior_multi(cge_multi *m, const cge_params *p, const float *obs,
                                        void *posterior_host, double *marg_mu, double *marg_sigma,
                                        double *expectations, cge_timing *timing) {
  CKRC(check_multi_params(m, p));
  if (!obs) return set_error(CGE_ERR_BAD_ARG, "obs is NULL");
  if (p->n_obs < 1) return set_error(CGE_ERR_BAD_ARG, "n_obs must be >= 1 (got %d)", p->n_obs);
  const int N = p->grid_size, n = p->n_obs;
  CKRC(prepare(m, p));

  // observations: one pinned staging copy every device can read; small ones are read in place by
  // the kernels, larger ones are copied to each device on its own stream
  if ((size_t)n > m->obs_pinned_cap) {
    if (m->obs_pinned) CK(cudaFreeHost(m->obs_pinned));
    m->obs_pinned = nullptr;
    m->obs_pinned_cap = 0;
    CK(cudaHostAlloc(&m->obs_pinned, ((size_t)n + 64) * sizeof(float),
                     cudaHostAllocPortable | cudaHostAllocMapped));
    m->obs_pinned_cap = (size_t)n + 64;
  }
  std::memcpy(m->obs_pinned, obs, (size_t)n * sizeof(float));
  std::vector<const float *> obs_ptr(m->n, nullptr);
  const bool in_place = n <= kZeroCopyObsMax;
  for (int i = 0; i < m->n; ++i) {
    CK(cudaSetDevice(m->devices[i]));
    if (timing) CK(cudaEventRecord(m->ev0[i], m->stream[i]));
    if (in_place) {
      obs_ptr[i] = m->obs_pinned;
    } else {
      if ((size_t)n > m->obs_cap[i]) {
        if (m->obs_dev[i]) CK(cudaFree(m->obs_dev[i]));
        m->obs_dev[i] = nullptr;
        m->obs_cap[i] = 0;
        CK(cudaMalloc(&m->obs_dev[i], ((size_t)n + 64) * sizeof(float)));
        m->obs_cap[i] = (size_t)n + 64;
      }
      CK(cudaMemcpyAsync(m->obs_dev[i], m->obs_pinned, (size_t)n * sizeof(float),
                         cudaMemcpyHostToDevice, m->stream[i]));
      obs_ptr[i] = m->obs_dev[i];
    }
  }
  // results: every finish kernel writes its result block straight into its context's pinned
  // host block (no device-to-host copies to issue); optionally the posterior shards are copied,
  // each over its own stream
  CKRC(issue_step(m, p, obs_ptr.data(), true));
  const size_t cell = p->posterior_dtype == CGE_POSTERIOR_F64 ? sizeof(double) : sizeof(float);
  if (timing || posterior_host)
    for (int i = 0; i < m->n; ++i) {
      CK(cudaSetDevice(m->devices[i]));
      if (timing) CK(cudaEventRecord(m->ev1[i], m->stream[i]));
      if (posterior_host)
        CK(cudaMemcpyAsync(static_cast<unsigned char *>(posterior_host) + (size_t)m->rb[i] * N * cell,
                           m->post[i], (size_t)(m->re[i] - m->rb[i]) * N * cell, cudaMemcpyDeviceToHost,
                           m->stream[i]));
    }
  CKRC(cge_multi_synchronize(m));
  // head and sigma marginal from shard 0 (identical on every shard: the same rank-ordered sums),
  // each shard's own rows of the mu marginal
  const double *h = m->res_host[0];
  if (expectations) {
    expectations[0] = h[0];
    expectations[1] = h[1];
  }
  if (marg_sigma) std::memcpy(marg_sigma, h + 4, (size_t)N * sizeof(double));
  if (marg_mu)
    for (int i = 0; i < m->n; ++i)
      std::memcpy(marg_mu + m->rb[i], m->res_host[i] + 4 + N + m->rb[i],
                  (size_t)(m->re[i] - m->rb[i]) * sizeof(double));
  if (timing) {
    cge_timing t = {};
    for (int i = 0; i < m->n; ++i) {
      float ms = 0;
      CK(cudaSetDevice(m->devices[i]));
      CK(cudaEventElapsedTime(&ms, m->ev0[i], m->ev1[i]));
      if (ms > t.total_ms) t.total_ms = ms;
    }
    *timing = t;
  }
  if (!(h[2] > 0.0) || !std::isfinite(h[2]))
    return set_error(CGE_ERR_DEGENERATE,
                     "normaliser Z = %g: every cell underflowed, the inputs are not finite, or a "
                     "peer did not answer (use CGE_MODE_LOGSPACE for large n_obs)", h[2]);
  return CGE_OK;
}
