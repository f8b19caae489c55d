// cge_extras.cu -- the rows SURVEY section 8(f) lists as "next" around the hot path:
//   * observation synthesis with the reference's generator (cuRAND XORWOW, seed 42, one
//     subsequence per observation) so the reference binary's own inputs can be reproduced;
//   * credible intervals from a marginal (the notebook's suggested follow-up).
// Separate translation unit: curand_kernel.h is heavy and nothing on the hot path needs it.
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include <cstring>

#include "../../include/cge.h"
#include "cge_device.cuh"

#define CGE_EXPORT extern "C" __attribute__((visibility("default")))

extern "C" int cge_internal_stream(cge_context *ctx, void **stream, int *sm_count);
extern "C" int cge_internal_error(int code, const char *msg);
extern "C" int cge_internal_scratch(cge_context *ctx, size_t count, double **ptr);

namespace {

// reference: inc/kernels.h:7-49 (setup_kernel + generate_normal_kernel), launch shape
// inc/cuda_functions.h:25-43: ceil(n/64) blocks of 64 threads, so every observation i is drawn
// by thread i from a state initialised as curand_init(42, i, 0): one draw per subsequence.
// mu + sigma*rv contracts to one FMA under the reference's default -fmad=true.
__global__ void generate_normal_kernel(float *out, unsigned n, float mu, float sigma,
                                       unsigned long long seed, int zero_state) {
  const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  curandState st;
  if (zero_state) {
    // src/tests.cu:21-46 runs the draw kernel on a zero-filled managed curandState
    memset(&st, 0, sizeof st);
  } else {
    curand_init(seed, id, 0, &st);
  }
  // the reference re-reads the (never written back) initial state for every element a thread
  // serves (kernels.h:44-47); with its launch shape each thread serves exactly one
  out[id] = __fmaf_rn(sigma, curand_normal(&st), mu);
}

// inclusive prefix sum of `n` doubles by one CTA in a fixed order (thread t owns a contiguous
// chunk; the 1024 chunk totals are scanned by shuffles inside each warp and then across the 32
// warp totals -- a fixed tree, no serial pass), then the first index whose cumulative mass exceeds
// each probability -- notebooks/quick-prototyping.ipynb cell 5 (credible_interval_90).
struct QuantileProbs {
  double p[16];
};
__global__ void __launch_bounds__(1024)
    quantile_kernel(const double *__restrict__ marginal, int n, const QuantileProbs probs, int nprobs,
                    int *__restrict__ idx_out, double *__restrict__ cdf_out) {
  __shared__ double warp_total[32];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + nt - 1) / nt;
  const int b = tid * per, e = min(n, b + per);
  double s = 0.0;
  for (int i = b; i < e; ++i) s += marginal[i];
  double incl = s;  // inclusive scan of the chunk totals inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += up;
  }
  if (lane == 31) warp_total[warp] = incl;
  if (tid < nprobs) idx_out[tid] = n;  // "never exceeded"
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the warp totals
    const double w = warp_total[lane];
    double wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += up;
    }
    warp_total[lane] = wi - w;
  }
  __syncthreads();
  double run = warp_total[warp] + (incl - s);  // mass before this thread's chunk
  for (int i = b; i < e; ++i) {
    const double prev = run;
    run += marginal[i];
    if (cdf_out) cdf_out[i] = run;
    for (int q = 0; q < nprobs; ++q)
      if (prev <= probs.p[q] && run > probs.p[q]) atomicMin(&idx_out[q], i);
  }
}

}  // namespace

CGE_EXPORT int cge_generate_normal(cge_context *ctx, float *out_dev, int n, float mu, float sigma,
                                   unsigned long long seed, int flags) {
  void *st = nullptr;
  int sms = 0;
  int rc = cge_internal_stream(ctx, &st, &sms);
  if (rc != CGE_OK) return rc;
  if (!out_dev || n < 1) return cge_internal_error(CGE_ERR_BAD_ARG, "generate_normal: bad argument");
  cudaStream_t stream = (cudaStream_t)st;
  generate_normal_kernel<<<(n + 63) / 64, 64, 0, stream>>>(out_dev, (unsigned)n, mu, sigma, seed,
                                                           flags & 1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cge_internal_error(CGE_ERR_CUDA, cudaGetErrorString(e));
  return CGE_OK;
}

CGE_EXPORT int cge_quantiles(cge_context *ctx, const double *marginal_dev, int n,
                             const double *probs_host, int nprobs, int *indices_host,
                             double *cdf_dev) {
  void *st = nullptr;
  int sms = 0;
  int rc = cge_internal_stream(ctx, &st, &sms);
  if (rc != CGE_OK) return rc;
  if (!marginal_dev || n < 1 || !probs_host || nprobs < 1 || nprobs > 16 || !indices_host)
    return cge_internal_error(CGE_ERR_BAD_ARG, "quantiles: bad argument");
  cudaStream_t stream = (cudaStream_t)st;
  double *scratch = nullptr;  // context-owned: no allocation on the call path
  rc = cge_internal_scratch(ctx, 16, &scratch);
  if (rc != CGE_OK) return rc;
  int *idx_dev = reinterpret_cast<int *>(scratch);
  QuantileProbs probs = {};
  for (int q = 0; q < nprobs; ++q) probs.p[q] = probs_host[q];
  quantile_kernel<<<1, 1024, 0, stream>>>(marginal_dev, n, probs, nprobs, idx_dev, cdf_dev);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(indices_host, idx_dev, nprobs * sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cge_internal_error(CGE_ERR_CUDA, cudaGetErrorString(e));
  return CGE_OK;
}
