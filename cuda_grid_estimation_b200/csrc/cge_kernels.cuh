// cge_kernels.cuh -- the sm_100a kernels of the grid-estimation hot path.
//
// Reference path being replaced (paths relative to cpp/grid-algorithm/ of the reference):
//   K1 linspaceKernel (inc/kernels.h:51-65), K2 create3dGridKernel (:68-99), K3 normalPdf
//   (:102-112), K4 computeDensitiesKernel (:115-141), K5 computeLikesKernel (:144-205),
//   K6 thrust reduce+transform (inc/cuda_functions.h:155-158), K7 marginalizeKernel
//   (inc/kernels.h:208-259), K8 computeExpectationsCuda (inc/cuda_functions.h:185-221).
//
// Here K1-K5 are ONE fused kernel (grid_likelihood_kernel) that never materialises the
// N*N*n tensors, and K6-K8 are a fixed-order partial reduction (two small kernels) plus one
// HBM-bound in-place scale pass.  A step is 5 launches: setup, likelihood, fold, results, scale.
//
// Data layout in HBM
//   posterior  float32 [rows_local][N], sigma fastest (the reference layout, kernels.h:92)
//   mu, sigma  float32 [N]      range tables, bit-identical to the reference's linspaceKernel
//   ad         float64 [N]      a_j  = -0.5*log2(e)/sigma_j^2
//   cd0        float64 [N]      c0_j = -log2(sigma_j*sqrt(2*pi))
//   obs_pad    float32 [n_pad]  observations, zero-padded to a multiple of 4, 16 B aligned
//   rowpart    float64 [rows_local][ncw]   per (row, warp column strip) partial row sums
//   colpart    float64 [ncolparts][N]      per (row tile x warp row) partial column sums
//   partials   float64 [N+2]    {Z, sum_i mu_i*rowsum_i, colsum[0..N)}  <- the all-reduce payload
//   results    float64 [4+2N]   {E[mu], E[sigma], Z, 1/Z}, marg_sigma[N], marg_mu[N]
#pragma once
#include "cge_device.cuh"

namespace cge {

// Peak of the rescaled likelihood: the largest cell ends the product at 2^kPeakLog2.  f32 spans
// 2^-126..2^127: 87 bits of headroom for partial products that overshoot the final value and
// 166 bits (1e-50 relative to the peak) below it before a cell flushes to zero.
constexpr double kPeakLog2 = 40.0;

constexpr int kRM = 4;            // register patch rows (mu) per thread for the product policies;
                                  // also the row alignment of every tile (Policy::kRows divides it)
constexpr int kRN = 4;            // columns (sigma) per thread and float4 store; a policy owns
                                  // Policy::kCols = 4 or 8 columns = 1 or 2 such groups
constexpr int kGroupCols = 32 * kRN;  // 128 columns: one float4 per lane across a warp

// column of patch element `cidx` for a thread whose first column is jbase: groups of 4 adjacent
// columns, successive groups 128 columns apart (so every float4 store of a warp is one 512-byte run)
__device__ __forceinline__ int patch_col(int jbase, int cidx) {
  return jbase + (cidx >> 2) * kGroupCols + (cidx & 3);
}
constexpr int kObsChunk = 8192;   // floats per streamed observation chunk (32 KB)
constexpr int kObsSingleMax = 16384;  // up to this many (padded) observations live in smem at once

struct ScaleInfo {
  double log2s;  // product mode: log2 of the per-factor rescale s
  double shift;  // log-space mode: added to every cell's log2-likelihood
  double m2;     // log2 of the largest likelihood on the grid (closed form, scale choice only)
  double xbar;
  double ss;
  double smin;
  double mu_c;   // log-space: grid value nearest the sample mean; obs_pad holds x_k - mu_c
  double s_c;    // log-space: sum_k (x_k - mu_c)^2
  double umax;   // bound on |t_(i+1,j,k) - t_(i,j,k)| over adjacent mu rows, whole grid
  double ustep;  // the same bound per unit |a_j|: umax of a column band = max |a_j| in it * ustep
};

struct LikeArgs {
  const float *obs_pad;
  const float *mu;
  const double *ad;
  const double *cd0;
  const ScaleInfo *scale;
  // optional prior (NULL = flat, the reference's case): separable factors indexed by the global
  // mu row / sigma column, and/or a full table laid out like this shard of the posterior
  const float *prior_mu;
  const float *prior_sigma;
  const float *prior_full;
  float *post;      // shard base (row `row_begin` at offset 0)
  double *rowpart;  // [rows_local][ncw]
  double *colpart;  // [ncolparts][N]
  int n, n_pad, N;  // n observations; n_pad staged floats (multiple of 4)
  int row_begin, row_end;
  int tile_rows;    // rows per CTA tile, multiple of WM*Policy::kRows and of kRM
  int ncw;          // column strips per row: ceil(N / (32 * Policy::kCols))
};

// per-thread column coefficients, fetched before the policy is known (dual kernel)
template <int KC>
struct ColumnCoefs {
  double a[KC], c0[KC];
  __device__ __forceinline__ void load(const LikeArgs &g, int j0) {
#pragma unroll
    for (int cidx = 0; cidx < KC; ++cidx) {
      const int j = patch_col(j0, cidx);
      a[cidx] = j < g.N ? g.ad[j] : 0.0;
      c0[cidx] = j < g.N ? g.cd0[j] : 0.0;
    }
  }
};

template <int WN>
__device__ __forceinline__ int first_column(int kcols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return (blockIdx.x * WN + warp % WN) * (32 * kcols) + lane * kRN;
}

// ---------------------------------------------------------------------------------------
// setup_kernel: K1 twice + the per-sigma coefficients (every block), and in block 0 the
// observation staging and the global rescale.
//
// The rescale only positions the float32 dynamic range; it cancels in the posterior.  It needs
// max_ij log2 L_ij = max_j (a_j * S_min + n*c0_j) with S_min = min_i [SS + n (xbar - mu_i)^2].
// S(mu) is a parabola and f(sigma) = a(sigma) S + n c0(sigma) is unimodal (peak at
// sigma^2 = S/n), so the grid maxima sit next to the continuous ones: a handful of candidate
// grid points (the neighbours of the continuous optimum plus both ends) is exact, O(1).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double coef_a(double sd) { return -0.5 * 1.4426950408889634074 / (sd * sd); }
__device__ __forceinline__ double coef_c0(double sd) { return -log2(sd * 2.5066282746310005024); }

struct RangeArgs {
  int N;
  float mu0, mu1, sg0, sg1;
};

// one grid point of the four tables; idx in [N, N+3) pads mu for the kernels' float4 row loads
__device__ __forceinline__ void setup_point(int idx, const RangeArgs &r, float *__restrict__ mu,
                                            float *__restrict__ sigma, double *__restrict__ ad,
                                            double *__restrict__ cd0) {
  if (idx >= r.N && idx < r.N + 3) mu[idx] = range_value(r.N - 1, r.N, r.mu0, r.mu1);
  if (idx < r.N) {
    mu[idx] = range_value(idx, r.N, r.mu0, r.mu1);
    const float s = range_value(idx, r.N, r.sg0, r.sg1);
    sigma[idx] = s;
    ad[idx] = coef_a((double)s);
    cd0[idx] = coef_c0((double)s);
  }
}

// Block-wide (every thread of the block must call it, every thread gets the result): stage the
// observations -- zero-padded to n_pad, raw or centred (x_k - mu_c) -- into `stage` (global or
// shared) and derive the scale statistics.  `scratch` holds >= 32 doubles of shared memory.
__device__ __forceinline__ ScaleInfo setup_scale(const RangeArgs &r, const float *__restrict__ obs,
                                                 float *stage, int n, int n_pad, int centred,
                                                 double *scratch) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int N = r.N;
  const float mu0 = r.mu0, mu1 = r.mu1, sg0 = r.sg0, sg1 = r.sg1;
  double sum = 0.0, xlo = 1.0 / 0.0, xhi = -1.0 / 0.0;
  for (int k = tid; k < n_pad; k += nt) {
    const float x = k < n ? obs[k] : 0.0f;
    stage[k] = x;
    sum += (double)x;
    if (k < n) {
      xlo = fmin(xlo, (double)x);
      xhi = fmax(xhi, (double)x);
    }
  }
  const double xbar = block_sum(sum, scratch) / (double)n;
  xlo = block_min(xlo, scratch);
  xhi = block_max(xhi, scratch);
  // centre of the log-space kernel: the grid value nearest the sample mean (depends on the global
  // grid only, so every shard uses the same one).  In that mode the staged observations are
  // x_k - mu_c, rounded once to float32.
  double mu_c = 0.0;
  {
    const double span = (double)mu1 - (double)mu0;
    double pos = span != 0.0 ? (xbar - (double)mu0) / span * (double)(N - 1) : 0.0;
    if (!(pos == pos) || pos < 0.0) pos = 0.0;
    if (pos > (double)(N - 1)) pos = (double)(N - 1);
    mu_c = (double)range_value((int)(pos + 0.5), N, mu0, mu1);
  }
  // (from here on the observations are read back from `stage`, not from `obs`, which may be
  // mapped host memory: each thread re-reads exactly the elements it staged itself)
  double q = 0.0;
  for (int k = tid; k < n; k += nt) {
    const double d = (double)stage[k] - xbar;
    q += d * d;
  }
  const double ss = block_sum(q, scratch);
  if (centred)
    for (int k = tid; k < n; k += nt) stage[k] = (float)((double)stage[k] - mu_c);

  // candidate grid indices around a continuous optimum `pos`: both ends + 4 neighbours;
  // lane t < 6 of every warp evaluates candidate t, the other lanes idle through the shuffles
  auto candidate = [N](double pos, int t) {
    if (!(pos == pos) || pos < 0.0) pos = 0.0;          // NaN / below range
    if (pos > (double)(N - 1)) pos = (double)(N - 1);
    const int f = (int)pos;
    if (t == 0) return 0;
    if (t == 1) return N - 1;
    return min(max(f - 3 + t, 0), N - 1);
  };
  const double mspan = (double)mu1 - (double)mu0;
  const double mpos = mspan != 0.0 ? (xbar - (double)mu0) / mspan * (double)(N - 1) : 0.0;
  double smin = 1.0 / 0.0;
  if (lane < 6) {
    const double d = xbar - (double)range_value(candidate(mpos, lane), N, mu0, mu1);
    smin = ss + (double)n * d * d;
  }
  smin = warp_min(smin);
  const double sspan = (double)sg1 - (double)sg0;
  const double spos =
      sspan != 0.0 ? (sqrt(smin / (double)n) - (double)sg0) / sspan * (double)(N - 1) : 0.0;
  double m2 = -1.0 / 0.0;
  if (lane < 6) {
    const double sd = (double)range_value(candidate(spos, lane), N, sg0, sg1);
    m2 = coef_a(sd) * smin + (double)n * coef_c0(sd);
  }
  m2 = warp_max(m2);
  ScaleInfo out;
  out.log2s = (kPeakLog2 - m2) / (double)n;
  out.shift = kPeakLog2 - m2;
  out.m2 = m2;
  out.xbar = xbar;
  out.ss = ss;
  out.smin = smin;
  out.mu_c = mu_c;
  out.s_c = ss + (double)n * (xbar - mu_c) * (xbar - mu_c);
  // Largest exponent step between vertically adjacent cells,
  //   |t_(i+1,j,k) - t_(i,j,k)| = |a_j| |mu_(i+1) - mu_i| |mu_i + mu_(i+1) - 2 x_k|,
  // bounded over the whole grid and every observation.  The row spacing gets 2 ulp of slack for
  // the float32 rounding of the range values.  ProductPaired is only used below kPairedUmax.
  {
    const double mlo = fmin((double)mu0, (double)mu1), mhi = fmax((double)mu0, (double)mu1);
    const double reach = fmax(fmax(fabs(xhi - mlo), fabs(xlo - mlo)), fmax(fabs(xhi - mhi), fabs(xlo - mhi)));
    const double dmu = N > 1 ? (mhi - mlo) / (double)(N - 1) + 2.4e-7 * fmax(fabs(mlo), fabs(mhi)) : 0.0;
    const double sdmin = fmin((double)sg0, (double)sg1);
    out.ustep = dmu * 2.0 * reach;
    out.umax = fabs(coef_a(sdmin)) * out.ustep;
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// Arithmetic policies of the fused likelihood kernel.  Each owns the per-thread state for a
// kRM x kRN patch of cells and consumes one observation at a time.
// ---------------------------------------------------------------------------------------
template <class Policy>
__device__ __forceinline__ void consume_scalar_obs(Policy &pol, const float *xs, int count,
                                                   const float (&mu)[Policy::kRows]);

// float32 product path: per pdf-eval one FFMA (exponent), one MUFU.EX2, one FMUL (running
// product); the subtract and square are shared by the kRN columns of a row.
//   factor_k = 2^(a_j*(x_k-mu_i)^2 + c_j),  c_j = c0_j + log2 s   (constants FMA-folded)
struct ProductF32 {
  static constexpr bool kColSmem = false;
  static constexpr int kRows = kRM;
  static constexpr int kCols = 4;
  static constexpr int kObsStride = 1;  // floats of shared memory per observation
  static __device__ __forceinline__ void consume(ProductF32 &pol, const float *xs, int count,
                                                 const float (&mu)[kRM]) {
    consume_scalar_obs(pol, xs, count, mu);
  }
  float a[kRN], c[kRN];
  float p[kRM][kRN];

  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf) {
#pragma unroll
    for (int cidx = 0; cidx < kRN; ++cidx) {
      const bool ok = j0 + cidx < g.N;
      a[cidx] = (float)cf.a[cidx];
      c[cidx] = ok ? (float)(cf.c0[cidx] + sc.log2s) : 0.0f;
    }
  }
  __device__ __forceinline__ void begin_rows() {
#pragma unroll
    for (int r = 0; r < kRM; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kRN; ++cidx) p[r][cidx] = 1.0f;
  }
  __device__ __forceinline__ void step(float x, const float (&mu)[kRM]) {
    float dd[kRM];
#pragma unroll
    for (int r = 0; r < kRM; ++r) {
      float d = x - mu[r];  // same f32 rounding as the reference's (x - mu), kernels.h:111
      dd[r] = d * d;
    }
#pragma unroll
    for (int r = 0; r < kRM; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kRN; ++cidx)
        p[r][cidx] *= ex2_approx(fmaf(dd[r], a[cidx], c[cidx]));
  }
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx) const {
    return p[r][cidx];
  }
};

// log-space path: float64 sum of log2-pdf terms (one DFMA per pdf-eval), then a single
// exponential per cell against the global maximum: L_ij = 2^(sum_k a_j d_ik^2 + n c0_j + shift).
// The FP64 pipe is the bound, so everything that is not the per-eval DFMA is kept off it:
// (x - mu)^2 is formed in float32 (2^-24 relative rounding per term, random: ~1e-6 relative on a
// cell after 10^4 observations) and widened once per (row, observation); 8 columns per thread
// amortise that widening over 8 evaluations.
struct LogSpaceF64 {
  static constexpr bool kColSmem = false;
  static constexpr int kRows = 4;   // 4 x 8 patch (a 2 x 8 patch measured 4 % slower: the bound is
                                    // the FP64 pipe's operand bandwidth, not occupancy)
  static constexpr int kCols = 8;
  static constexpr int kObsStride = 1;
  static __device__ __forceinline__ void consume(LogSpaceF64 &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const int n4 = count >> 2;
#pragma unroll 1
    for (int q = 0; q < n4; ++q) {  // 4 observations x 32 independent DFMA chains per iteration
      const float4 x = xs4[q];
      pol.step(x.x, mu);
      pol.step(x.y, mu);
      pol.step(x.z, mu);
      pol.step(x.w, mu);
    }
    for (int k = n4 << 2; k < count; ++k) pol.step(xs[k], mu);
  }
  double a[kCols];
  double acc[kRows][kCols];
  double shift_;
  int jbase_;

  __device__ __forceinline__ void load_columns(const LikeArgs &, int jbase, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf) {
    jbase_ = jbase;
    shift_ = sc.shift;
#pragma unroll
    for (int cidx = 0; cidx < kCols; ++cidx) a[cidx] = cf.a[cidx];
  }
  __device__ __forceinline__ void begin_rows() {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) acc[r][cidx] = 0.0;
  }
  __device__ __forceinline__ void step(float x, const float (&mu)[kRows]) {
    double dd[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float d = x - mu[r];
      dd[r] = (double)(d * d);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) acc[r][cidx] = fma(dd[r], a[cidx], acc[r][cidx]);
  }
  // n*c0_j + shift is only needed here, once per cell: read c0_j back instead of holding it
  __device__ __forceinline__ float value(const LikeArgs &g, int r, int cidx) const {
    const int j = patch_col(jbase_, cidx);
    const double c = j < g.N ? (double)g.n * g.cd0[j] + shift_ : 0.0;
    return ex2_approx((float)(acc[r][cidx] + c));
  }
};

// ---------------------------------------------------------------------------------------
// log-space path, float32x2 tiles: the same sum of log2-pdf terms, one FMA per pdf-eval, but on
// the FP32 pipe (packed FFMA2, twice the FP64 pipe's rate) with the float64 work reduced to one
// add per cell and tile of observations.  What makes float32 accumulation accurate enough is
// centring: with mu_c the grid value nearest the sample mean, x' = x - mu_c, mu' = mu_i - mu_c,
//   (x_k - mu_i)^2 = (x_k - mu_c)^2 + e_ik,    e_ik = mu'_i (mu'_i - 2 x'_k)      (exact identity)
//   log2 L_ij = [a_j S_c + n c0_j] + sum_k a_j e_ik,        S_c = sum_k (x_k - mu_c)^2
// The bracket is float64, once per cell (S_c from setup_kernel).  The per-eval terms a_j e_ik are
// small wherever the posterior is not negligible (|mu'| <= ~sqrt(66 / (n |a|)) for cells within
// 1e-20 of the peak), so float32 tile sums (TILE = 256 observations) folded into float64 lose
// ~7e-6 relative on a cell at n = 10^4 -- a plain float32 sum of a_j (x - mu)^2 would lose 1e-4.
// Per (row, observation): one scalar FFMA (e = x' * (-2 mu') + mu'^2) shared by the 8 columns,
// then 4 FFMA2.
// ---------------------------------------------------------------------------------------
template <int TILE = 256>
struct LogSpaceTiled {
  static constexpr double kUmax = 1.0 / 0.0;
  static constexpr bool kColSmem = false;
  static constexpr int kRows = 4;
  static constexpr int kCols = 8;
  static constexpr int kObsStride = 1;
  static constexpr int kTile = TILE;  // observations per float32 tile sum
  static constexpr int kPairs = kCols / 2;
  f32x2 a2[kPairs];
  double acc[kRows][kCols];
  double s_c_, shift_, mu_c_;
  int jbase_;

  __device__ __forceinline__ void load_columns(const LikeArgs &, int jbase, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf) {
    jbase_ = jbase;
    s_c_ = sc.s_c;
    shift_ = sc.shift;
    mu_c_ = sc.mu_c;
#pragma unroll
    for (int h = 0; h < kPairs; ++h) a2[h] = pack2((float)cf.a[2 * h], (float)cf.a[2 * h + 1]);
  }
  __device__ __forceinline__ void begin_rows() {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) acc[r][cidx] = 0.0;
  }
  __device__ __forceinline__ void step(float xc, const float (&gm)[kRows], const float (&gb)[kRows],
                                       f32x2 (&t2)[kRows][kPairs]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float e = fmaf(xc, gm[r], gb[r]);
      const f32x2 e2 = pack2(e, e);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) t2[r][h] = fma2(e2, a2[h], t2[r][h]);
    }
  }
  __device__ __forceinline__ void fold(f32x2 (&t2)[kRows][kPairs]) {
    const f32x2 zero = pack2(0.0f, 0.0f);
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        float lo, hi;
        unpack2(t2[r][h], lo, hi);
        acc[r][2 * h] += (double)lo;
        acc[r][2 * h + 1] += (double)hi;
        t2[r][h] = zero;
      }
  }
  static __device__ __forceinline__ void consume(LogSpaceTiled &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    float gm[kRows], gb[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float m = (float)((double)mu[r] - pol.mu_c_);
      gm[r] = -2.0f * m;
      gb[r] = m * m;
    }
    f32x2 t2[kRows][kPairs];
    const f32x2 zero = pack2(0.0f, 0.0f);
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) t2[r][h] = zero;
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const int ntiles = count / kTile;
#pragma unroll 1
    for (int tIdx = 0; tIdx < ntiles; ++tIdx) {
#pragma unroll 1
      for (int q = 0; q < kTile / 4; ++q) {  // 4 observations x 16 independent FFMA2 chains
        const float4 x = xs4[tIdx * (kTile / 4) + q];
        pol.step(x.x, gm, gb, t2);
        pol.step(x.y, gm, gb, t2);
        pol.step(x.z, gm, gb, t2);
        pol.step(x.w, gm, gb, t2);
      }
      pol.fold(t2);
    }
    const int rest = ntiles * kTile;
    if (rest < count) {
      for (int k = rest; k < count; ++k) pol.step(xs[k], gm, gb, t2);
      pol.fold(t2);
    }
  }
  __device__ __forceinline__ float value(const LikeArgs &g, int r, int cidx) const {
    const int j = patch_col(jbase_, cidx);
    if (j >= g.N) return 0.0f;
    const double base = g.ad[j] * s_c_ + (double)g.n * g.cd0[j] + shift_;
    return ex2_approx((float)(acc[r][cidx] + base));
  }
};

// consume `count` observations from shared memory (16-byte aligned), four per LDS.128 broadcast
template <class Policy>
__device__ __forceinline__ void consume_scalar_obs(Policy &pol, const float *xs, int count,
                                                   const float (&mu)[Policy::kRows]) {
  const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
  const int n4 = count >> 2;
#pragma unroll 2
  for (int q = 0; q < n4; ++q) {
    float4 x = xs4[q];
    pol.step(x.x, mu);
    pol.step(x.y, mu);
    pol.step(x.z, mu);
    pol.step(x.w, mu);
  }
  for (int k = n4 << 2; k < count; ++k) pol.step(xs[k], mu);
}

// ---------------------------------------------------------------------------------------
// float32 product path, hybrid exponential: the same per-pdf-eval work as ProductF32
// (exponent FMA, 2^t, running product) but
//   * the per-cell arithmetic runs as packed f32x2 (FFMA2/FMUL2): a thread's 4 columns are two
//     register pairs; the per-row (x - mu)^2 stays scalar and enters FFMA2 as a broadcast operand;
//   * of the 8 column pairs a thread updates per observation, NPA (even observations) or NPB
//     (odd observations) pairs take 2^t from the degree-5 FMA-pipe polynomial (ex2_poly2)
//     instead of MUFU.EX2.  With MUFU-only code the XU pipe is 96 % busy while the FMA pipe
//     idles at 30 % and half the issue slots are free (profiles/r01_ncu_full_16384x50.csv);
//     moving ~5/16 of the exponentials over balances the two pipes.
// Which cells use which pipe is fixed at compile time, so results stay bit-reproducible.
// ---------------------------------------------------------------------------------------
template <int NPA, int NPB, bool /*unused*/ = false, int OBS_PER_ITER = 4>
struct ProductHybrid {
  static constexpr double kUmax = 1.0 / 0.0;  // any grid
  static constexpr bool kColSmem = false;
  static constexpr int kRows = kRM;
  static constexpr int kCols = 4;
  static constexpr int kObsStride = 1;
  f32x2 a2[kRN / 2], c2[kRN / 2];
  f32x2 p2[kRM][kRN / 2];

  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf) {
#pragma unroll
    for (int h = 0; h < kRN / 2; ++h) {
      float av[2], cv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + 2 * h + e;
        av[e] = (float)cf.a[2 * h + e];
        cv[e] = j < g.N ? (float)(cf.c0[2 * h + e] + sc.log2s) : 0.0f;
      }
      a2[h] = pack2(av[0], av[1]);
      c2[h] = pack2(cv[0], cv[1]);
    }
  }
  __device__ __forceinline__ void begin_rows() {
    const f32x2 one = pack2(1.0f, 1.0f);
#pragma unroll
    for (int r = 0; r < kRM; ++r)
#pragma unroll
      for (int h = 0; h < kRN / 2; ++h) p2[r][h] = one;
  }
  // one observation; the first NP (row, pair) slots take 2^t from the polynomial
  template <int NP>
  __device__ __forceinline__ void step_np(float x, const float (&mu)[kRM]) {
#pragma unroll
    for (int r = 0; r < kRM; ++r) {
      const float d = x - mu[r];  // same f32 rounding as the reference's (x - mu), kernels.h:111
      const float dd = d * d;
      const f32x2 dd2 = pack2(dd, dd);  // becomes a scalar-broadcast operand of FFMA2
#pragma unroll
      for (int h = 0; h < kRN / 2; ++h) {
        const f32x2 t2 = fma2(dd2, a2[h], c2[h]);
        // spread the polynomial slots over the rows: (r, h) = (0,0),(1,1),(2,0),(3,1), then the rest
        const int slot = ((h + r) & 1) * kRM + r;
        const f32x2 e2 = slot < NP ? ex2_poly2(t2) : ex2_mufu2(t2);
        p2[r][h] = mul2(p2[r][h], e2);
      }
    }
  }
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductHybrid &pol, const float *xs, int count,
                                                 const float (&mu)[kRM]) {
    if (OBS_PER_ITER == 4) {
      const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
      const int n4 = count >> 2;
      for (int q = 0; q < n4; ++q) {
        const float4 x = xs4[q];
        pol.template step_np<NPA>(x.x, mu);
        pol.template step_np<NPB>(x.y, mu);
        pol.template step_np<NPA>(x.z, mu);
        pol.template step_np<NPB>(x.w, mu);
      }
      for (int k = n4 << 2; k < count; ++k) {
        if (k & 1) pol.template step_np<NPB>(xs[k], mu);
        else pol.template step_np<NPA>(xs[k], mu);
      }
    } else if (OBS_PER_ITER == 2) {
      const float2 *xs2 = reinterpret_cast<const float2 *>(xs);
      const int n2 = count >> 1;
#pragma unroll 1
      for (int q = 0; q < n2; ++q) {
        const float2 x = xs2[q];
        pol.template step_np<NPA>(x.x, mu);
        pol.template step_np<NPB>(x.y, mu);
      }
      if (count & 1) pol.template step_np<NPA>(xs[count - 1], mu);
    } else {
#pragma unroll 1
      for (int k = 0; k < count; ++k) pol.template step_np<NPA>(xs[k], mu);
    }
  }
};

// ---------------------------------------------------------------------------------------
// float32 product path, paired rows: of every two vertically adjacent cells (mu_i, mu_(i+1) at
// the same sigma_j) one takes its factor from MUFU.EX2 and the other derives its own from it,
//   pdf(x_k; mu_(i+1), sigma_j) = pdf(x_k; mu_i, sigma_j) * 2^u,
//   u = a_j * g_ik,   g_ik = (mu_(i+1) - mu_i) * (mu_i + mu_(i+1) - 2 x_k)      (exact identity)
// with 2^u - 1 from a degree-2 polynomial on the FMA pipe: on a fine grid |u| is tiny (6e-3 at
// 16384 rows for the README ranges), so no range reduction is needed.  Every pdf-eval still forms
// its own factor and multiplies it into its own running product; what changes is the pipe mix:
//   MUFU row   : FFMA2 (exponent), 2 MUFU.EX2, FMUL2 (product)            per column pair
//   derived row: FFMA2 (a_j k1 + g a_j^2 k2), FMUL2 (* g), FFMA2 (e + e v), FMUL2 (product)
// i.e. half the MUFU work of ProductF32 for ~3 packed FMA-pipe instructions more per 4 evals;
// ncu showed the MUFU-only kernel with the XU pipe 97 % busy and the FMA pipe at 30 %.
// Polynomial: 2^u = 1 + u (k1 + k2 u) with k2 = ln(2)^2/2 and the cubic term folded into
// k1 = ln 2 + (ln(2)^3/6)(3/4) U^2 (Chebyshev economisation on [-U, U], U = ScaleInfo::umax):
// error <= 0.0139 U^3 = 1.8e-8 at the eligibility bound kPairedUmax, below the float32 rounding
// of the factor itself.  The derived factor is e*(1+v) evaluated as fma(e, v, e): one rounding.
// DEG = 3 adds the cubic term, 2^u - 1 = u (k1 + u (k2 + u k3)), k3 = ln(2)^3/6 and the quartic
// folded into k2 = ln(2)^2/2 + (ln(2)^4/24) U^2: error <= 0.0024 U^4 = 1.5e-8 at its bound
// kPairedUmax3 = 0.05, which takes grids 4.5x coarser (README ranges: N >= 690) for one more
// FFMA2 per derived pair.  Grids too coarse for that run ProductHybrid instead (chosen on the
// device, see grid_likelihood_select_kernel); which rows are MUFU rows depends on the global row
// index only.
// ---------------------------------------------------------------------------------------
constexpr double kPairedUmax = 0.011;
constexpr double kPairedUmax3 = 0.05;

template <int KC = 4, int OBS_PER_ITER = 1, bool PACKED_ROWS = false, bool COL_SMEM = false,
          int DEG = 2>
struct ProductPaired {
  static_assert(DEG == 2 || DEG == 3, "polynomial degree of 2^u - 1");
  static constexpr double kUmax = DEG == 2 ? kPairedUmax : kPairedUmax3;  // eligibility bound
  // COL_SMEM: the tile's float64 column sums live in shared memory instead of 2*KC registers
  // (they are touched once per row group), which is what lets the kernel fit 80 registers
  static constexpr bool kColSmem = COL_SMEM;
  static constexpr int kRows = kRM;
  static constexpr int kCols = KC;
  static constexpr int kObsStride = 1;
  static constexpr int kPairs = KC / 2;
  f32x2 a2[kPairs], c2[kPairs];    // exponent coefficients of the MUFU rows
  f32x2 k1a[kPairs], k2a[kPairs];  // a_j k1, a_j^2 k2 of the derived rows
  f32x2 k3a[DEG == 3 ? kPairs : 1];  // a_j^3 k3
  f32x2 p2[kRM][kPairs];

  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf) {
    const double U = sc.umax;
    const double k1 = 0.69314718055994530942 + (DEG == 2 ? 0.0555041086648215799 * 0.75 * U * U : 0.0);
    const double k2 = 0.24022650695910071233 + (DEG == 3 ? 0.00961812910762847716 * U * U : 0.0);
    const double k3 = 0.0555041086648215799;
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      float av[2], cv[2], k1v[2], k2v[2], k3v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = patch_col(j0, 2 * h + e);
        const double a = cf.a[2 * h + e];
        av[e] = (float)a;
        cv[e] = j < g.N ? (float)(cf.c0[2 * h + e] + sc.log2s) : 0.0f;
        k1v[e] = (float)(a * k1);
        k2v[e] = (float)(a * a * k2);
        k3v[e] = (float)(a * a * a * k3);
      }
      a2[h] = pack2(av[0], av[1]);
      c2[h] = pack2(cv[0], cv[1]);
      k1a[h] = pack2(k1v[0], k1v[1]);
      k2a[h] = pack2(k2v[0], k2v[1]);
      if (DEG == 3) k3a[h] = pack2(k3v[0], k3v[1]);
    }
  }
  __device__ __forceinline__ void begin_rows() {
    const f32x2 one = pack2(1.0f, 1.0f);
#pragma unroll
    for (int r = 0; r < kRM; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) p2[r][h] = one;
  }
  // one observation: rows (0,1) and (2,3) of the patch are the pairs; gm/gb give g = x*gm + gb
  __device__ __forceinline__ void step(float x, const float (&mu)[kRM], const float (&gm)[kRM / 2],
                                       const float (&gb)[kRM / 2]) {
    float ddv[kRM / 2], ggv[kRM / 2];
    if (PACKED_ROWS) {  // the two row pairs' scalars as one packed FADD2 / FMUL2 / FFMA2 each
      const f32x2 x2 = pack2(x, x);
      const f32x2 d2 = sub2(x2, pack2(mu[0], mu[2]));
      unpack2(mul2(d2, d2), ddv[0], ddv[1]);
      unpack2(fma2(x2, pack2(gm[0], gm[1]), pack2(gb[0], gb[1])), ggv[0], ggv[1]);
    } else {
#pragma unroll
      for (int rp = 0; rp < kRM / 2; ++rp) {
        const float d = x - mu[2 * rp];  // same f32 rounding as the reference's (x - mu), kernels.h:111
        ddv[rp] = d * d;
        ggv[rp] = fmaf(x, gm[rp], gb[rp]);
      }
    }
#pragma unroll
    for (int rp = 0; rp < kRM / 2; ++rp) {
      const float dd = ddv[rp], gg = ggv[rp];
      const f32x2 dd2 = pack2(dd, dd);
      const f32x2 g2 = pack2(gg, gg);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        const f32x2 e2 = ex2_mufu2(fma2(dd2, a2[h], c2[h]));
        p2[2 * rp][h] = mul2(p2[2 * rp][h], e2);
        const f32x2 q2 = DEG == 3 ? fma2(g2, k3a[h], k2a[h]) : k2a[h];
        const f32x2 v2 = mul2(g2, fma2(g2, q2, k1a[h]));  // 2^u - 1
        const f32x2 f2 = fma2(e2, v2, e2);                    // neighbour's factor
        p2[2 * rp + 1][h] = mul2(p2[2 * rp + 1][h], f2);
      }
    }
  }
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductPaired &pol, const float *xs, int count,
                                                 const float (&mu)[kRM]) {
    float gm[kRM / 2], gb[kRM / 2];
#pragma unroll
    for (int rp = 0; rp < kRM / 2; ++rp) {
      const float delta = __fsub_rn(mu[2 * rp + 1], mu[2 * rp]);  // exact (adjacent range values)
      gm[rp] = -2.0f * delta;
      gb[rp] = __fmul_rn(delta, __fadd_rn(mu[2 * rp], mu[2 * rp + 1]));
    }
    if (OBS_PER_ITER == 4) {
      const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
      const int n4 = count >> 2;
#pragma unroll 1
      for (int q = 0; q < n4; ++q) {
        const float4 x = xs4[q];
        pol.step(x.x, mu, gm, gb);
        pol.step(x.y, mu, gm, gb);
        pol.step(x.z, mu, gm, gb);
        pol.step(x.w, mu, gm, gb);
      }
      for (int k = n4 << 2; k < count; ++k) pol.step(xs[k], mu, gm, gb);
    } else if (OBS_PER_ITER == 2) {
      const float2 *xs2 = reinterpret_cast<const float2 *>(xs);
      const int n2 = count >> 1;
#pragma unroll 1
      for (int q = 0; q < n2; ++q) {
        const float2 x = xs2[q];
        pol.step(x.x, mu, gm, gb);
        pol.step(x.y, mu, gm, gb);
      }
      if (count & 1) pol.step(xs[count - 1], mu, gm, gb);
    } else {
#pragma unroll 1
      for (int k = 0; k < count; ++k) pol.step(xs[k], mu, gm, gb);
    }
  }
};

// ---------------------------------------------------------------------------------------
// grid_likelihood_kernel: K1-K5 fused.
//   CTA = WM x WN warps.  A warp owns a strip of 32*kCols sigma columns (lane -> kCols/4 groups
//   of 4 adjacent columns, one 16-byte store per group and row) and walks down its rows kRM at
//   a time.  blockIdx.x = column band (WN strips), blockIdx.y = row tile (tile_rows rows).
//   Observations reach shared memory by TMA bulk copy (cp.async.bulk + mbarrier): once per CTA
//   when they fit, else streamed in double-buffered 32 KB chunks.
//   Per row group the warp emits its partial row sums (fixed-order shuffle tree) and keeps
//   float64 column sums in registers until the tile ends: no atomics anywhere.
// ---------------------------------------------------------------------------------------
// PRELOADED: the caller has already staged all n_pad observations at smem_raw + 128 and
// synchronised the block (small_grid_kernel); no mbarrier / bulk copy is used then.
template <class Policy, int WM, int WN, bool VEC4, bool PRELOADED = false>
__device__ __forceinline__ void likelihood_body(const LikeArgs &g, const ScaleInfo &sc,
                                                const ColumnCoefs<Policy::kCols> &cf,
                                                unsigned char *smem_raw) {
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);         // 2 mbarriers
  float *obs_s = reinterpret_cast<float *>(smem_raw + 128);         // stage 0 (and 1 if streamed)

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wn = warp % WN, wm = warp / WN;
  constexpr int kRows = Policy::kRows;            // rows per thread: 4 or 2
  constexpr int kCols = Policy::kCols;            // columns per thread: 4 or 8
  constexpr int kGroups = kCols / kRN;            // float4 groups per thread
  const int cw = blockIdx.x * WN + wn;            // global column-strip index (32*kCols columns)
  const int j0 = cw * (32 * kCols) + lane * kRN;  // first column; group g starts 128*g further
  // tiles start at a multiple of kRM below row_begin, so a cell's position inside the register
  // patch (hence which exponential path it takes) depends on its GLOBAL row only: a row shard
  // computes bit-identical unnormalised cells to the single-GPU run
  const int tile_row0 = (g.row_begin & ~(kRM - 1)) + blockIdx.y * g.tile_rows;
  const int groups = g.tile_rows / (WM * kRows);    // row groups this warp walks
  constexpr int kStride = Policy::kObsStride;        // staged floats per observation
  constexpr int kChunkObs = kObsChunk / kStride;     // observations per streamed chunk
  const bool streamed = !PRELOADED && g.n_pad > kObsSingleMax;  // n_pad counts staged floats
  const int nchunks = streamed ? (g.n_pad + kObsChunk - 1) / kObsChunk : 1;

  if (!PRELOADED) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_init_fence();
    }
    __syncthreads();
  }

  auto issue = [&](int seq) {  // thread 0 only: start the bulk copy of load number `seq`
    const int ch = seq % nchunks, st = seq & 1;
    const int floats = streamed ? min(kObsChunk, g.n_pad - ch * kObsChunk) : g.n_pad;
    const uint32_t bytes = (uint32_t)floats * 4u;
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_copy_g2s(obs_s + (size_t)st * kObsChunk, g.obs_pad + (size_t)ch * kObsChunk, bytes,
                  &bars[st]);
  };
  if (!PRELOADED && tid == 0) issue(0);

  Policy pol;
  pol.load_columns(g, j0, sc, cf);
  // float64 column sums of this thread's columns over the tile: registers, or (Policy::kColSmem)
  // shared memory behind the observation stage, element cidx of thread tid at [cidx*nthreads + tid]
  constexpr bool kColSmem = Policy::kColSmem;
  double colreg[kColSmem ? 1 : kCols];
  double *colsm = reinterpret_cast<double *>(
      smem_raw + 128 + (((streamed ? 2 * kObsChunk : g.n_pad) * 4 + 15) & ~15));
#pragma unroll
  for (int cidx = 0; cidx < kCols; ++cidx) {
    if (kColSmem) colsm[cidx * (WM * WN * 32) + tid] = 0.0;
    else colreg[cidx] = 0.0;
  }
  const bool col_any = j0 < g.N;  // columns only grow with cidx
  const bool has_prior = g.prior_mu || g.prior_sigma || g.prior_full;

  float mu[kRows];
  int row = 0;
  bool rows_any = false;

  auto begin_group = [&](int grp) {
    row = tile_row0 + (grp * WM + wm) * kRows;
    rows_any = row < g.row_end && row + kRows > g.row_begin;
    // `row` is a multiple of 4 and setup_kernel pads the table with 3 copies of its last value,
    // so the 4 rows of the patch are one aligned 16-byte load
    const float4 m4 = *reinterpret_cast<const float4 *>(g.mu + min(row, (g.N - 1) & ~3));
    mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
    pol.begin_rows();
  };

  // Epilogue of one row group: store the kRows x kCols patch, add it into the row / column sums.
  // Within the patch the sums are float32 in a fixed pairwise order (4 or 8 terms); everything
  // across threads, groups and tiles is float64 in a fixed order.  The kRows row sums of a warp
  // are reduced together by a transposing butterfly (12 shuffles instead of 40): lane 8*q ends up
  // with the total of row q.
  auto end_group = [&]() {
    if (!rows_any) return;  // warp-uniform
    float v[kRows][kCols];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) v[r][cidx] = pol.value(g, r, cidx);
    const bool interior =
        row >= g.row_begin && row + kRows <= g.row_end && patch_col(j0, kCols - 1) < g.N;
    if (!interior) {  // shard / grid edges: cells outside contribute exactly zero
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const bool rv = row + r >= g.row_begin && row + r < g.row_end;
#pragma unroll
        for (int cidx = 0; cidx < kCols; ++cidx)
          if (!(rv && patch_col(j0, cidx) < g.N)) v[r][cidx] = 0.0f;
      }
    }
    if (has_prior) {  // likelihood x prior before normalisation (cuda_functions.h:145-148)
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const bool rv = row + r >= g.row_begin && row + r < g.row_end;
#pragma unroll
        for (int cidx = 0; cidx < kCols; ++cidx) {
          const int j = patch_col(j0, cidx);
          if (!(rv && j < g.N)) continue;
          if (g.prior_mu) v[r][cidx] *= g.prior_mu[row + r];
          if (g.prior_sigma) v[r][cidx] *= g.prior_sigma[j];
          if (g.prior_full)
            v[r][cidx] *= g.prior_full[(size_t)(row + r - g.row_begin) * (size_t)g.N + j];
        }
      }
    }
    float *rowp = g.post + (ptrdiff_t)(row - g.row_begin) * (ptrdiff_t)g.N + j0;
    if (VEC4 && __all_sync(0xffffffffu, interior)) {
      // the whole warp is inside the shard and the grid (all but the edge tiles): straight-line
      // stores, one 512-byte run per instruction
#pragma unroll
      for (int r = 0; r < kRows; ++r, rowp += g.N)
#pragma unroll
        for (int grp4 = 0; grp4 < kGroups; ++grp4)
          st_stream_f4(rowp + grp4 * kGroupCols, v[r][4 * grp4], v[r][4 * grp4 + 1],
                       v[r][4 * grp4 + 2], v[r][4 * grp4 + 3]);
    } else {
#pragma unroll
      for (int r = 0; r < kRows; ++r, rowp += g.N) {
        const bool rv = interior || (row + r >= g.row_begin && row + r < g.row_end);
#pragma unroll
        for (int grp4 = 0; grp4 < kGroups; ++grp4) {
          const int jg = j0 + grp4 * kGroupCols;
          if (VEC4) {
            if (rv && jg < g.N)
              st_stream_f4(rowp + grp4 * kGroupCols, v[r][4 * grp4], v[r][4 * grp4 + 1],
                           v[r][4 * grp4 + 2], v[r][4 * grp4 + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < kRN; ++e)
              if (rv && jg + e < g.N) st_stream_f1(rowp + grp4 * kGroupCols + e, v[r][4 * grp4 + e]);
          }
        }
      }
    }
    // column sums over the patch rows, row sums over the patch columns (float32, pairwise)
    static_assert(kRows == 4, "the patch reductions below are written for 4 rows");
#pragma unroll
    for (int cidx = 0; cidx < kCols; ++cidx)
    {
      const double add = (double)((v[0][cidx] + v[1][cidx]) + (v[2][cidx] + v[3][cidx]));
      if (kColSmem) colsm[cidx * (WM * WN * 32) + tid] += add;
      else colreg[cidx] += add;
    }
    double rs[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float t = (v[r][0] + v[r][1]) + (v[r][2] + v[r][3]);
      if (kCols == 8) t += (v[r][4] + v[r][5]) + (v[r][6] + v[r][7]);
      rs[r] = (double)t;
    }
    const bool up16 = lane & 16, up8 = lane & 8;
    const double k0 = (up16 ? rs[2] : rs[0]) + __shfl_xor_sync(0xffffffffu, up16 ? rs[0] : rs[2], 16);
    const double k1 = (up16 ? rs[3] : rs[1]) + __shfl_xor_sync(0xffffffffu, up16 ? rs[1] : rs[3], 16);
    double tot = (up8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
    tot += __shfl_xor_sync(0xffffffffu, tot, 4);
    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    const int rq = row + (lane >> 3);  // lanes 0, 8, 16, 24 publish rows 0..3 of the patch
    if ((lane & 7) == 0 && rq >= g.row_begin && rq < g.row_end && cw < g.ncw)
      g.rowpart[(size_t)(rq - g.row_begin) * (size_t)g.ncw + cw] = tot;
  };

  if (!streamed) {
    if (!PRELOADED) mbar_wait(&bars[0], 0);
    for (int grp = 0; grp < groups; ++grp) {
      begin_group(grp);
      // (no skip for groups or columns outside the grid: they are masked in the epilogue, and a
      // single control path keeps the patch initialisation from being emitted per branch)
      Policy::consume(pol, obs_s, g.n, mu);
      end_group();
    }
  } else {
    const int total = groups * nchunks;
    for (int seq = 0; seq < total; ++seq) {
      if (tid == 0 && seq + 1 < total) issue(seq + 1);
      mbar_wait(&bars[seq & 1], (seq >> 1) & 1);
      const int ch = seq % nchunks;
      if (ch == 0) begin_group(seq / nchunks);
      const int count = min(kChunkObs, g.n - ch * kChunkObs);
      if (rows_any && col_any)
        Policy::consume(pol, obs_s + (size_t)(seq & 1) * kObsChunk, count, mu);
      if (ch == nchunks - 1) end_group();
      __syncthreads();  // every warp is done with this stage before it is refilled
    }
  }

  // partial column sums of this warp's rows: one slot per (row tile, warp row)
  if (col_any) {
    double *dst = g.colpart + (size_t)(blockIdx.y * WM + wm) * (size_t)g.N;
#pragma unroll
    for (int cidx = 0; cidx < kCols; ++cidx) {
      const int j = patch_col(j0, cidx);
      if (j < g.N) dst[j] = kColSmem ? colsm[cidx * (WM * WN * 32) + tid] : colreg[cidx];
    }
  }
}

// ScaleInfo as this CTA's column band sees it: `umax` becomes the bound over the band's own
// columns (|a_j| is monotone along the sigma axis, so its maximum sits at one end of the band).
// The exponent step scales with |a_j| = 0.72/sigma_j^2, so on a grid with a wide sigma range only
// the bands at small sigma need the costlier arithmetic.  A function of the band (columns) and of
// the global statistics only: the same for every row tile, shard and rank.
template <int WN>
__device__ __forceinline__ ScaleInfo band_scale(const LikeArgs &g, int kcols) {
  ScaleInfo sc = *g.scale;
  const int jfirst = blockIdx.x * WN * 32 * kcols;
  const int jlast = min(jfirst + WN * 32 * kcols, g.N) - 1;
  sc.umax = fmax(fabs(g.ad[jfirst]), fabs(g.ad[jlast])) * sc.ustep;
  return sc;
}

template <class Policy, int WM, int WN, bool VEC4, int MINB = 1>
__global__ void __launch_bounds__(WM *WN * 32, MINB)
    grid_likelihood_kernel(const LikeArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColumnCoefs<Policy::kCols> cf;
  cf.load(g, first_column<WN>(Policy::kCols));
  const ScaleInfo sc = band_scale<WN>(g, Policy::kCols);
  likelihood_body<Policy, WM, WN, VEC4>(g, sc, cf, smem_raw);
}

// The default product kernel: ProductPaired with the degree-2 polynomial where the grid is fine
// enough for it (the band's umax, from the statistics setup_kernel derives on the device out of
// the ranges and the observations), with the degree-3 polynomial up to 4.5x coarser,
// ProductHybrid otherwise.  The choice is made per column band and depends on (ranges, N,
// observations, band) only -- never on the row tile or the shard -- so every rank of a sharded
// run takes the same path for the same cells and results stay bit-reproducible.  The policies
// share the tiling.
template <class Fine, class Mid, class Coarse, int WM, int WN, bool VEC4, int MINB = 1>
__global__ void __launch_bounds__(WM *WN * 32, MINB)
    grid_likelihood_select_kernel(const LikeArgs g) {
  static_assert(Fine::kCols == Coarse::kCols && Fine::kRows == Coarse::kRows, "same tiling");
  static_assert(Fine::kCols == Mid::kCols && Fine::kRows == Mid::kRows, "same tiling");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColumnCoefs<Fine::kCols> cf;  // in flight together with the statistics, before the path is known
  cf.load(g, first_column<WN>(Fine::kCols));
  const ScaleInfo sc = band_scale<WN>(g, Fine::kCols);
  if (sc.umax <= Fine::kUmax)
    likelihood_body<Fine, WM, WN, VEC4>(g, sc, cf, smem_raw);
  else if (sc.umax <= Mid::kUmax)
    likelihood_body<Mid, WM, WN, VEC4>(g, sc, cf, smem_raw);
  else
    likelihood_body<Coarse, WM, WN, VEC4>(g, sc, cf, smem_raw);
}

// ---------------------------------------------------------------------------------------
// small_grid_kernel: the whole step in ONE launch for grids of a few hundred points per axis,
// where five launches are all latency (the README workload, N = 101, is 510 050 pdf-evals:
// ~2 us of arithmetic).  Every CTA
//   1. fills the four range tables and derives the scale statistics itself (identical values
//      from every CTA: a benign race) and stages the observations straight into its shared
//      memory -- setup_kernel's work, redundantly, instead of a launch and a dependency;
//   2. runs the same fused likelihood body on its tile (unnormalised cells + partial sums);
//   3. takes a ticket; the last CTA to finish folds the partials in a fixed order, writes Z, the
//      marginals and the expectations, and scales the whole posterior by 1/Z (it fits in L2).
// Observations may be read from, and the results block also written to, mapped pinned host
// memory, which removes both small DMA copies (~6 us and ~10 us) from the call's latency.
// No grid-wide barrier, so no co-residency requirement.  Sums are in a fixed order (index
// order within a thread, then the fixed block tree): bit-reproducible run to run.
// ---------------------------------------------------------------------------------------
constexpr int kSmallColChunk = 256;  // columns folded per pass of the last CTA (8 x 256 f64 of smem)

struct SmallArgs {
  LikeArgs like;  // tables written by this kernel; row_begin = 0, row_end = N
  RangeArgs range;
  const float *obs;  // raw observations (device)
  float *mu, *sigma;
  double *ad, *cd0;
  ScaleInfo *scale;
  int centred;
  int ncolparts;
  double *partials, *rowsum, *results;
  double *results_host;  // optional mapped pinned copy of `results` (saves the D2H copy)
  unsigned int *ticket;
};

// in place: post[0..count) *= finv, reading through L2 (other CTAs wrote some of it), several
// independent loads in flight per thread; `post` is 16-byte aligned when VEC4
template <bool VEC4>
__device__ __forceinline__ void scale_cells_cta(float *post, size_t count, float finv) {
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int kDepth = 8;
  if (VEC4) {
    float4 *p4 = reinterpret_cast<float4 *>(post);
    const size_t nvec = count >> 2;  // VEC4 implies N % 4 == 0, so count % 4 == 0
    for (size_t base = tid; base < nvec; base += (size_t)nt * kDepth) {
      float4 v[kDepth];
#pragma unroll
      for (int u = 0; u < kDepth; ++u)
        if (base + (size_t)u * nt < nvec) v[u] = __ldcg(p4 + base + (size_t)u * nt);
#pragma unroll
      for (int u = 0; u < kDepth; ++u)
        if (base + (size_t)u * nt < nvec) {
          v[u].x *= finv; v[u].y *= finv; v[u].z *= finv; v[u].w *= finv;
          p4[base + (size_t)u * nt] = v[u];
        }
    }
  } else {
    for (size_t base = tid; base < count; base += (size_t)nt * kDepth) {
      float v[kDepth];
#pragma unroll
      for (int u = 0; u < kDepth; ++u)
        if (base + (size_t)u * nt < count) v[u] = __ldcg(post + base + (size_t)u * nt);
#pragma unroll
      for (int u = 0; u < kDepth; ++u)
        if (base + (size_t)u * nt < count) post[base + (size_t)u * nt] = v[u] * finv;
    }
  }
}

// hardware barrier over the thread-block cluster; release/acquire at cluster scope orders the
// global-memory writes made before it against the reads made after it in the other CTAs
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}

// CLUSTER = false: any grid; the last CTA to take the ticket does K6-K8 and scales the whole
//   posterior.
// CLUSTER = true: the grid is ONE thread-block cluster (one column band, <= 8 row tiles: every
//   grid up to 256 x 256), so its CTAs are co-scheduled and a hardware cluster barrier replaces
//   the ticket: after it every CTA folds the row partials itself (same order everywhere, so the
//   same Z bit for bit) and scales ITS OWN tile in parallel; CTA 0 also writes the marginals and
//   expectations.  No single CTA walks the whole posterior.
template <class Fine, class Mid, class Coarse, int WM, int WN, bool VEC4, bool CLUSTER = false>
__global__ void __launch_bounds__(WM *WN * 32)
    small_grid_kernel(const SmallArgs a) {
  static_assert(Fine::kCols == Coarse::kCols && Fine::kRows == Coarse::kRows, "same tiling");
  static_assert(Fine::kCols == Mid::kCols && Fine::kRows == Mid::kRows, "same tiling");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double scratch[32];
  __shared__ double colred[WM * WN][kSmallColChunk];
  __shared__ bool is_last;
  const LikeArgs &g = a.like;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = g.N;
  float *obs_s = reinterpret_cast<float *>(smem_raw + 128);

  for (int idx = tid; idx < N + 3; idx += nt) setup_point(idx, a.range, a.mu, a.sigma, a.ad, a.cd0);
  ScaleInfo sc = setup_scale(a.range, a.obs, obs_s, g.n, g.n_pad, a.centred, scratch);
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *a.scale = sc;
  __syncthreads();  // this CTA's own table writes and staged observations are visible to it

  ColumnCoefs<Fine::kCols> cf;
  cf.load(g, first_column<WN>(Fine::kCols));
  {  // the band's own bound, as band_scale() computes it from the tables
    const int jfirst = blockIdx.x * WN * 32 * Fine::kCols;
    const int jlast = min(jfirst + WN * 32 * Fine::kCols, N) - 1;
    sc.umax = fmax(fabs(a.ad[jfirst]), fabs(a.ad[jlast])) * sc.ustep;
  }
  if (sc.umax <= Fine::kUmax)
    likelihood_body<Fine, WM, WN, VEC4, true>(g, sc, cf, smem_raw);
  else if (sc.umax <= Mid::kUmax)
    likelihood_body<Mid, WM, WN, VEC4, true>(g, sc, cf, smem_raw);
  else
    likelihood_body<Coarse, WM, WN, VEC4, true>(g, sc, cf, smem_raw);

  bool writer;  // the CTA that writes partials / results: the last one, or cluster rank 0
  if (CLUSTER) {
    cluster_barrier();
    writer = blockIdx.y == 0;
  } else {
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    writer = true;
  }

  // ---- K6-K8 on the partial buffers of every CTA (read through L2).  Every loop keeps several
  // independent loads in flight: one CTA is latency-bound, not bandwidth-bound.
  double z = 0.0, smu = 0.0;
  for (int i = tid; i < N; i += nt) {
    double s = 0.0;
    for (int w = 0; w < g.ncw; ++w) s += __ldcg(g.rowpart + (size_t)i * g.ncw + w);
    if (writer) a.rowsum[i] = s;
    z += s;
    smu += s * (double)a.mu[i];
  }
  if (writer) {  // column sums: warp w takes slots w, w+NW, ... of a column (all its loads in
                 // flight at once), then the NW partial sums are added in warp order
    constexpr int NW = WM * WN;
    const int lane = tid & 31, warp = tid >> 5;
    for (int j0c = 0; j0c < N; j0c += kSmallColChunk) {
      for (int j = j0c + lane; j < min(N, j0c + kSmallColChunk); j += 32) {
        const double *src = g.colpart + j;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int t = warp;
        for (; t + 3 * NW < a.ncolparts; t += 4 * NW) {
          s0 += __ldcg(src + (size_t)t * N);
          s1 += __ldcg(src + (size_t)(t + NW) * N);
          s2 += __ldcg(src + (size_t)(t + 2 * NW) * N);
          s3 += __ldcg(src + (size_t)(t + 3 * NW) * N);
        }
        for (; t < a.ncolparts; t += NW) s0 += __ldcg(src + (size_t)t * N);
        colred[warp][j - j0c] = (s0 + s1) + (s2 + s3);
      }
      __syncthreads();
      for (int j = j0c + tid; j < min(N, j0c + kSmallColChunk); j += nt) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += colred[w][j - j0c];
        a.partials[2 + j] = s;
      }
      __syncthreads();
    }
  }
  z = block_sum(z, scratch);
  smu = block_sum(smu, scratch);
  const double inv = 1 / z;  // reciprocal multiply, as inc/utils.h:82-84
  if (writer) {
    double *marg_sigma = a.results + 4, *marg_mu = a.results + 4 + N;
    double es = 0.0;
    for (int j = tid; j < N; j += nt) {  // partials[2 + j]: written before the block-wide syncs above
      const double m = a.partials[2 + j] * inv;
      marg_sigma[j] = m;
      if (a.results_host) a.results_host[4 + j] = m;
      es += m * (double)a.sigma[j];
    }
    for (int i = tid; i < N; i += nt) {
      const double m = a.rowsum[i] * inv;
      marg_mu[i] = m;
      if (a.results_host) a.results_host[4 + N + i] = m;
    }
    es = block_sum(es, scratch);
    if (tid == 0) {
      a.partials[0] = z;
      a.partials[1] = smu;
      const double head[4] = {smu * inv, es, z, inv};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a.results[q] = head[q];
        if (a.results_host) a.results_host[q] = head[q];
      }
      if (!CLUSTER) *a.ticket = 0u;
    }
  }
  const float finv = (float)inv;
  if (CLUSTER) {  // this CTA's own rows
    const int r0 = blockIdx.y * g.tile_rows;
    const int r1 = min(N, r0 + g.tile_rows);
    if (r0 < N) scale_cells_cta<VEC4>(g.post + (size_t)r0 * N, (size_t)(r1 - r0) * N, finv);
  } else {
    scale_cells_cta<VEC4>(g.post, (size_t)N * N, finv);
  }
}

}  // namespace cge
