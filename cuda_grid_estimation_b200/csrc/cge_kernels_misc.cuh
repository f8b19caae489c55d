// cge_kernels_misc.cuh -- the non-template kernels of the step (setup, fold, results, scale),
// the stage-API kernels and the pipe micro-benchmarks.  Included by cge_api.cu only (they are
// ordinary __global__ definitions: one translation unit must own them).
#pragma once
#include "cge_kernels.cuh"

namespace cge {

__global__ void __launch_bounds__(256)
    setup_kernel(float *__restrict__ mu, float *__restrict__ sigma, double *__restrict__ ad,
                 double *__restrict__ cd0, const RangeArgs r, const float *__restrict__ obs,
                 float *__restrict__ obs_pad, int n, int n_pad, int centred,
                 ScaleInfo *__restrict__ out) {
  __shared__ double scratch[32];
  setup_point(blockIdx.x * blockDim.x + threadIdx.x, r, mu, sigma, ad, cd0);
  if (blockIdx.x != 0) return;
  const ScaleInfo sc = setup_scale(r, obs, obs_pad, n, n_pad, centred, scratch);
  if (threadIdx.x == 0) *out = sc;
}

// ---------------------------------------------------------------------------------------
// K6-K8 replacement, stage 1: fold the partial buffers in a fixed order.
//   blocks [0, nb_col): 32 columns each; 8 slot groups stride the ncolparts slots (4 loads in
//                       flight per thread), then the groups are summed 0..7       -> colsum
//   blocks [nb_col, ..): warp per row, lanes stride the ncw strips, shuffle tree  -> rowsum;
//                        per-block partials of Z and sum mu_i*rowsum_i
//   the last block to finish (threadfence + ticket) adds the block partials in index order:
//                        partials[0] = Z, partials[1] = sum_i mu_i*rowsum_i
// The ticket only decides WHO does the final sum, never its order: bit-reproducible.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    fold_partials_kernel(const double *__restrict__ colpart, int ncolparts,
                         const double *__restrict__ rowpart, int ncw, int rows_local,
                         int row_begin, int N, const float *__restrict__ mu,
                         double *__restrict__ partials, double *__restrict__ rowsum,
                         double *__restrict__ zpart, double *__restrict__ smupart, int nb_col,
                         int nb_row, unsigned int *__restrict__ ticket,
                         unsigned long long *ready_flag, unsigned long long ready_value) {
  __shared__ double sh[8][33];
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x < nb_col) {
    const int j = blockIdx.x * 32 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (j < N) {
      const double *src = colpart + j;
      int t = warp;
      for (; t + 24 < ncolparts; t += 32) {
        s0 += src[(size_t)t * N];
        s1 += src[(size_t)(t + 8) * N];
        s2 += src[(size_t)(t + 16) * N];
        s3 += src[(size_t)(t + 24) * N];
      }
      for (; t < ncolparts; t += 8) s0 += src[(size_t)t * N];
    }
    sh[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && j < N) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sh[w][lane];
      partials[2 + j] = s;
    }
  } else {
    const int rb = blockIdx.x - nb_col;
    const int il = rb * 8 + warp;
    double s = 0.0;
    if (il < rows_local) {
      const double *src = rowpart + (size_t)il * ncw;
      for (int w = lane; w < ncw; w += 32) s += src[w];
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (il < rows_local) rowsum[il] = s;
      sh[0][warp] = s;
      sh[1][warp] = il < rows_local ? s * (double)mu[row_begin + il] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
      double z = 0.0, m = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        z += sh[0][w];
        m += sh[1][w];
      }
      zpart[rb] = z;
      smupart[rb] = m;
    }
  }
  // ---- last block: Z and sum mu*rowsum over the row-block partials, in index order
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double z = 0.0, m = 0.0;
  for (int b = tid; b < nb_row; b += 256) {
    z += __ldcg(zpart + b);
    m += __ldcg(smupart + b);
  }
  z = warp_sum(z);
  m = warp_sum(m);
  __syncthreads();
  if (lane == 0) {
    sh[2][warp] = z;
    sh[3][warp] = m;
  }
  __syncthreads();
  if (tid == 0) {
    double zt = 0.0, mt = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      zt += sh[2][w];
      mt += sh[3][w];
    }
    partials[0] = zt;
    partials[1] = mt;
    *ticket = 0u;  // ready for the next launch on this stream
    if (ready_flag) {
      // peer exchange: every block's column sums are ordered before this point by the ticket
      // (fence + atomic); publish "the partial vector of step `ready_value - 1` is complete"
      // to the other GPUs, which poll this word over NVLink
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ready_flag), "l"(ready_value)
                   : "memory");
    }
  }
}

// stage 2, after the cross-rank sum of `partials`: marginals and expectations.
//   marg_sigma_j = colsum_j * (1/Z)   E[sigma] = sum_j marg_sigma_j * sigma_j   (cuda_functions.h:201-218)
//   marg_mu_i    = rowsum_i * (1/Z)   E[mu]    = (sum_i mu_i*rowsum_i) * (1/Z)
// 256 grid points per block; the last block adds the per-block E[sigma] partials in index order.
__global__ void __launch_bounds__(256)
    results_kernel(const double *__restrict__ partials, const double *__restrict__ rowsum,
                   const float *__restrict__ sigma, int N, int row_begin, int row_end,
                   double *__restrict__ results, double *__restrict__ espart,
                   unsigned int *__restrict__ ticket) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  const int tid = threadIdx.x;
  const double z = partials[0];
  const double inv = 1 / z;  // reciprocal multiply, as inc/utils.h:82-84
  double *marg_sigma = results + 4, *marg_mu = results + 4 + N;
  const int idx = blockIdx.x * 256 + tid;
  double es = 0.0;
  if (idx < N) {
    const double m = partials[2 + idx] * inv;
    marg_sigma[idx] = m;
    es = m * (double)sigma[idx];
    marg_mu[idx] = (idx >= row_begin && idx < row_end) ? rowsum[idx - row_begin] * inv : 0.0;
  }
  es = block_sum(es, scratch);
  if (tid == 0) espart[blockIdx.x] = es;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double e = 0.0;
  for (int b = tid; b < (int)gridDim.x; b += 256) e += __ldcg(espart + b);
  e = block_sum(e, scratch);
  if (tid == 0) {
    results[0] = partials[1] * inv;
    results[1] = e;
    results[2] = z;
    results[3] = inv;
    *ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// stage 2 with the exchange fused in (one process per GPU, one node): instead of a library
// all-reduce between the fold and this kernel, every rank reads the peers' partial vectors
// straight out of their HBM over NVLink (CUDA IPC mappings) and adds them in RANK ORDER -- the
// sum is bit-identical on every rank and independent of any collective algorithm.
//   peers.buf[r]   rank r's partial vector for this step (N+2 f64; this rank's own included)
//   peers.flag[r]  rank r's "step s complete" word; polled with ld.acquire.sys until >= want
// A rank may overwrite its vector of step s at step s+2: by then it has itself waited for every
// peer's step-(s+1) flag, which each peer raises only after its own step-s results kernel.
// The poll gives up after ~4e9 cycles and poisons the result (Z = NaN -> CGE_ERR_DEGENERATE).
// ---------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerView {
  const double *buf[kMaxPeers];
  const unsigned long long *flag[kMaxPeers];
  int world;
  unsigned long long want;
};

__device__ __forceinline__ double ld_peer(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
    results_peer_kernel(const PeerView peers, double *__restrict__ partials,
                        const double *__restrict__ rowsum, const float *__restrict__ sigma, int N,
                        int row_begin, int row_end, double *__restrict__ results,
                        double *__restrict__ espart, unsigned int *__restrict__ ticket) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  __shared__ int timed_out;
  const int tid = threadIdx.x;
  if (tid == 0) timed_out = 0;
  __syncthreads();
  if (tid < peers.world) {  // one lane per peer polls that peer's flag
    const long long t0 = clock64();
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(peers.flag[tid]) : "memory");
      if (v >= peers.want) break;
      if (clock64() - t0 > 4000000000LL) { timed_out = 1; break; }
      __nanosleep(100);
    } while (true);
  }
  __syncthreads();
  __threadfence_system();
  double z = 0.0, smu = 0.0;
  for (int r = 0; r < peers.world; ++r) {  // rank order: same bits on every rank
    z += ld_peer(peers.buf[r]);
    smu += ld_peer(peers.buf[r] + 1);
  }
  if (timed_out) z = __longlong_as_double(0x7ff8000000000000LL);
  const double inv = 1 / z;
  double *marg_sigma = results + 4, *marg_mu = results + 4 + N;
  const int idx = blockIdx.x * 256 + tid;
  double es = 0.0;
  if (idx < N) {
    double c = 0.0;
    for (int r = 0; r < peers.world; ++r) c += ld_peer(peers.buf[r] + 2 + idx);
    partials[2 + idx] = c;  // the summed vector, for callers that read cge_partials afterwards
    const double m = c * inv;
    marg_sigma[idx] = m;
    es = m * (double)sigma[idx];
    marg_mu[idx] = (idx >= row_begin && idx < row_end) ? rowsum[idx - row_begin] * inv : 0.0;
  }
  if (blockIdx.x == 0 && tid == 0) {
    partials[0] = z;  // the scale pass reads Z from here
    partials[1] = smu;
  }
  es = block_sum(es, scratch);
  if (tid == 0) espart[blockIdx.x] = es;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double e = 0.0;
  for (int b = tid; b < (int)gridDim.x; b += 256) e += __ldcg(espart + b);
  e = block_sum(e, scratch);
  if (tid == 0) {
    results[0] = smu * inv;
    results[1] = e;
    results[2] = z;
    results[3] = inv;
    *ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// scale pass: posterior *= 1/Z in place.  Pure HBM streaming: 4 B read + 4 B written per cell.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    scale_posterior_kernel(float *__restrict__ post, size_t count, const double *__restrict__ partials) {
  const float inv = (float)(1 / partials[0]);
  const size_t nvec = count >> 2;
  float4 *p4 = reinterpret_cast<float4 *>(post);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float4 a = p4[i], b = p4[i + stride], c = p4[i + 2 * stride], d = p4[i + 3 * stride];
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv;
    c.x *= inv; c.y *= inv; c.z *= inv; c.w *= inv;
    d.x *= inv; d.y *= inv; d.z *= inv; d.w *= inv;
    p4[i] = a; p4[i + stride] = b; p4[i + 2 * stride] = c; p4[i + 3 * stride] = d;
  }
  for (; i < nvec; i += stride) {
    float4 a = p4[i];
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    p4[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) post[(nvec << 2) + threadIdx.x] *= inv;
}

// same pass for a posterior pointer that is not 16-byte aligned (caller-owned sub-buffers)
__global__ void __launch_bounds__(256)
    scale_posterior_scalar_kernel(float *__restrict__ post, size_t count,
                                  const double *__restrict__ partials) {
  const float inv = (float)(1 / partials[0]);
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < count;
       f += (size_t)gridDim.x * blockDim.x)
    post[f] *= inv;
}

// ---------------------------------------------------------------------------------------
// stage API kernels (the reference's L1 functions at the reference's sizes)
// ---------------------------------------------------------------------------------------
__global__ void linspace_kernel(float *__restrict__ out, int size, float start, float end) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < size) {
    // the reference divides by (size-1) even when size == 1 (0/0); keep its value for size>1
    float span = __fsub_rn(end, start);
    float num = __fmul_rn(__int2float_rn(idx), span);
    out[idx] = __fadd_rn(start, __fdiv_rn(num, __int2float_rn(size - 1)));
  }
}

// coalesced along the flat output index (the reference maps threadIdx.x to the slowest index)
__global__ void meshgrid3_kernel(const float *__restrict__ vx, const float *__restrict__ vy,
                                 const float *__restrict__ vz, float *__restrict__ gx,
                                 float *__restrict__ gy, float *__restrict__ gz, int nxy, int nz) {
  const size_t total = (size_t)nxy * nxy * nz;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < total;
       f += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(f % nz);
    const size_t ij = f / nz;
    gx[f] = vx[ij / nxy];
    gy[f] = vy[ij % nxy];
    gz[f] = vz[k];
  }
}

// K3/K4 with the reference's mixed precision: f32 z, square, exp; f64 sqrt(2*pi)*sigma and
// divide; truncate to f32 (inc/kernels.h:111).
__global__ void densities_kernel(float *__restrict__ dens, const float *__restrict__ gx,
                                 const float *__restrict__ gy, const float *__restrict__ gz,
                                 size_t count) {
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < count;
       f += (size_t)gridDim.x * blockDim.x) {
    const float x = gz[f], m = gx[f], s = gy[f];
    const float z = __fdiv_rn(__fsub_rn(x, m), s);
    const float e = expf(__fmul_rn(-0.5f, __fmul_rn(z, z)));
    const double denom = (double)s * 2.5066282746310002;  // sqrt(2.0f * M_PI) in f64
    dens[f] = (float)((double)e / denom);
  }
}

// K5 replacement for dense [rows][cols] f32 densities: warp per row, f64 products,
// fixed-order shuffle tree.
__global__ void __launch_bounds__(256)
    row_products_kernel(const float *__restrict__ dens, double *__restrict__ likes, size_t rows,
                        int cols) {
  const int lane = threadIdx.x & 31;
  const size_t r = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  double p = 1.0;
  const float *src = dens + r * (size_t)cols;
  for (int k = lane; k < cols; k += 32) p *= (double)src[k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, off);
  if (lane == 0) likes[r] = p;
}

// a2+a4 fused for the stage wrapper (inc/wrappers.h:19-92): densities[(i*N + j)*n + k] =
// pdf(obs_k; mu_i, sigma_j) straight from the three vectors -- the reference's three N*N*n
// meshgrid copies are never built.
__global__ void densities_grid_kernel(float *__restrict__ dens, const float *__restrict__ mu,
                                      const float *__restrict__ sigma,
                                      const float *__restrict__ obs, int N, int n) {
  const size_t total = (size_t)N * N * n;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < total;
       f += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(f % n);
    const size_t ij = f / n;
    const float x = obs[k], m = mu[ij / N], s = sigma[ij % N];
    const float z = __fdiv_rn(__fsub_rn(x, m), s);
    const float e = expf(__fmul_rn(-0.5f, __fmul_rn(z, z)));
    dens[f] = (float)((double)e / ((double)s * 2.5066282746310002));
  }
}

// the reference's row-pointer matrices (double**, one allocation per row: inc/utils.h:53-63):
// K5 / K7 on that layout, warp per output element, fixed-order shuffle tree.
__global__ void __launch_bounds__(256)
    row_products_pp_kernel(double *__restrict__ likes, double *const *__restrict__ rows_ptr,
                           int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double *src = rows_ptr[r];
  double p = 1.0;
  for (int k = lane; k < cols; k += 32) p *= src[k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, off);
  if (lane == 0) likes[r] = p;
}

__global__ void __launch_bounds__(256)
    marginalize_pp_kernel(double *__restrict__ marginal, double *const *__restrict__ rows_ptr,
                          int rows, int cols, int axis) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  double s = 0.0;
  if (axis == 1) {  // row sums
    if (o >= rows) return;
    const double *src = rows_ptr[o];
    for (int j = lane; j < cols; j += 32) s += src[j];
  } else {          // column sums
    if (o >= cols) return;
    for (int i = lane; i < rows; i += 32) s += rows_ptr[i][o];
  }
  s = warp_sum(s);
  if (lane == 0) marginal[o] = s;
}

// K6: two-level fixed-order sum, then the reciprocal multiply
__global__ void __launch_bounds__(1024)
    block_sums_kernel(const double *__restrict__ v, size_t count, double *__restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < count;
       f += (size_t)gridDim.x * blockDim.x)
    s += v[f];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
    final_sum_kernel(const double *__restrict__ parts, int nparts, double *__restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int b = threadIdx.x; b < nparts; b += blockDim.x) s += parts[b];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void scale_f64_kernel(const double *__restrict__ likes, double *__restrict__ post,
                                 size_t count, const double *__restrict__ zsum) {
  const double inv = 1 / zsum[0];
  for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < count;
       f += (size_t)gridDim.x * blockDim.x)
    post[f] = likes[f] * inv;
}

// K7: axis=1 -> row sums (warp per row); axis=0 -> column sums (thread per column)
__global__ void __launch_bounds__(256)
    marginalize_kernel(double *__restrict__ marginal, const double *__restrict__ mat, int rows,
                       int cols, int axis) {
  if (axis == 1) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    double s = 0.0;
    for (int j = lane; j < cols; j += 32) s += mat[(size_t)r * cols + j];
    s = warp_sum(s);
    if (lane == 0) marginal[r] = s;
  } else {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    double s = 0.0;
    for (int i = 0; i < rows; ++i) s += mat[(size_t)i * cols + j];
    marginal[j] = s;
  }
}

// K8: sum_i marginal_i * double(vector_i)
__global__ void __launch_bounds__(1024)
    expectation_kernel(const double *__restrict__ marginal, const float *__restrict__ vec, int size,
                       double *__restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < size; i += blockDim.x) s += marginal[i] * (double)vec[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}

// ---------------------------------------------------------------------------------------
// pipe-peak micro-benchmarks: 8 independent dependency chains per thread, no memory traffic
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(float *out, int iters, float seed) {
  constexpr int C = 8;
  if (KIND == 0) {  // MUFU.EX2
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = seed + 0.001f * (threadIdx.x + c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = ex2_approx(v[c]) - 1.0f;  // one FADD per MUFU keeps it bounded
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;
  } else if (KIND == 1) {  // FFMA
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = seed + 0.001f * (threadIdx.x + c);
    const float a = 0.999f + seed * 1e-6f, b = 0.001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = fmaf(v[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;
  } else if (KIND == 2) {  // DFMA
    double v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = (double)seed + 0.001 * (threadIdx.x + c);
    const double a = 0.999 + (double)seed * 1e-6, b = 0.001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = fma(v[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += v[c];
    if (s == 123.456) out[0] = (float)s;
  } else if (KIND >= 4) {
    // mixes: NP packed FFMA2 chains + NS scalar FFMA chains (+ NM MUFU chains) per iteration --
    // do packed (heavy half of the FMA pipe only) and scalar (either half) instructions add up to
    // more than either alone, and how much FMA work fits beside a saturated XU pipe?
    //   4: 4 FFMA2 + 4 FFMA    5: 4 FFMA2 + 8 FFMA    6: 6 FFMA2 + 2 MUFU    7: 4 FFMA2 + 4 FFMA + 2 MUFU
    constexpr int NP = KIND == 6 ? 6 : 4;
    constexpr int NS = KIND == 4 ? 4 : KIND == 5 ? 8 : KIND == 7 ? 4 : 0;
    constexpr int NM = KIND >= 6 ? 2 : 0;
    unsigned long long v[NP];
    float w[NS > 0 ? NS : 1], m[NM > 0 ? NM : 1];
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      float lo = seed + 0.001f * (threadIdx.x + c), hi = lo + 0.5f;
      asm("mov.b64 %0, {%1, %2};" : "=l"(v[c]) : "f"(lo), "f"(hi));
    }
#pragma unroll
    for (int c = 0; c < NS; ++c) w[c] = seed + 0.002f * (threadIdx.x + c);
#pragma unroll
    for (int c = 0; c < NM; ++c) m[c] = seed + 0.003f * (threadIdx.x + c);
    unsigned long long a2, b2;
    const float a = 0.999f + seed * 1e-6f, b = 0.001f;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < (NP > NS ? NP : NS); ++c) {
        if (c < NP) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(a2), "l"(b2));
        if (c < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(w[c]) : "f"(a), "f"(b));
        if (c < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
      }
#pragma unroll
      for (int c = 0; c < NM; ++c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[c]) : "f"(a), "f"(-1.0f));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c]));
      s += lo + hi;
    }
#pragma unroll
    for (int c = 0; c < NS; ++c) s += w[c];
#pragma unroll
    for (int c = 0; c < NM; ++c) s += m[c];
    if (s == 123.456f) out[0] = s;
  } else {  // packed FFMA2: fma.rn.f32x2, two lanes-ops per instruction
    unsigned long long v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float lo = seed + 0.001f * (threadIdx.x + c), hi = lo + 0.5f;
      asm("mov.b64 %0, {%1, %2};" : "=l"(v[c]) : "f"(lo), "f"(hi));
    }
    unsigned long long a2, b2;
    {
      float a = 0.999f + seed * 1e-6f, b = 0.001f;
      asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
      asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(a2), "l"(b2));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c]));
      s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
  }
}

}  // namespace cge
