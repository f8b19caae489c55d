// cge_kernels.cuh -- the sm_100a kernels of the grid-estimation hot path.
//
// Reference path being replaced (paths relative to cpp/grid-algorithm/ of the reference):
//   K1 linspaceKernel (inc/kernels.h:51-65), K2 create3dGridKernel (:68-99), K3 normalPdf
//   (:102-112), K4 computeDensitiesKernel (:115-141), K5 computeLikesKernel (:144-205),
//   K6 thrust reduce+transform (inc/cuda_functions.h:155-158), K7 marginalizeKernel
//   (inc/kernels.h:208-259), K8 computeExpectationsCuda (inc/cuda_functions.h:185-221).
//
// A step is THREE launches chained by programmatic dependent launch (one for grids up to 256^2):
//   1. grid_likelihood_kernel  K1-K5 fused, persistent: one CTA per SM, its warps claiming row
//                              groups from the CTA's pool until the grid is done (see "Work
//                              decomposition").  Each CTA derives the scale statistics and its own
//                              range values and column coefficients itself (no set-up launch, no
//                              tables in HBM), writes the unnormalised cells, fixed-order partial
//                              row sums, and column sums accumulated in exact fixed point.
//   2. fold_publish_kernel     K6/K7 fold of the partial sums -> {Z, sum mu*rowsum, colsum[N]};
//                              publishes the vector to the peer GPUs (release flag).
//   3. finish_kernel           waits for the peers' vectors and adds them in rank order (when
//                              sharded), K7/K8 marginals and expectations, and the in-place
//                              posterior *= 1/Z (HBM-bound).
//
// Data layout in HBM
//   posterior  float32 [rows_local][N], sigma fastest (the reference layout, kernels.h:92)
//   rowpart    float64 [ncw][rows_local]    per (128-column slot, row) partial row sums: the rows a warp
//                                       walks down are contiguous, so L2 merges its 8-byte stores
//                                       into full sectors (the [row][slot] layout cost a read-modify-
//                                       write of every sector: 1.09 GB of DRAM reads at 65536^2)
//   colacc     int64   [N]      column sums in per-column fixed point (exact, order-independent)
//   partials   float64 [N+2]    {Z, sum_i mu_i*rowsum_i, colsum[0..N)}  <- the exchange payload
//   results    float64 [4+2N]   {E[mu], E[sigma], Z, 1/Z}, marg_sigma[N], marg_mu[N]
// Nothing of size N*N*n exists, and no mu / sigma / coefficient table either: every value is a
// function of (index, range) evaluated where it is used with the reference's float32 formula.
#pragma once
#include "cge_device.cuh"

namespace cge {

// Peak of the rescaled likelihood: the largest cell ends the product at 2^kPeakLog2.  f32 spans
// 2^-126..2^127: 87 bits of headroom for partial products that overshoot the final value and
// 166 bits (1e-50 relative to the peak) below it before a cell flushes to zero.
constexpr double kPeakLog2 = 40.0;

constexpr int kRowsProduct = 6;   // register patch rows (mu) per thread of the product policies: two
                                  // triples, the middle row of each on MUFU.EX2 (see ProductNeighbour)
constexpr int kRowsLog = 4;       // ... of the log-space policies
                                  // (a policy's kRows is also the row alignment of every unit of work)
constexpr int kRN = 4;            // columns (sigma) per float4 store; a policy owns
                                  // Policy::kCols = 4 or 8 columns = 1 or 2 such groups
constexpr int kGroupCols = 32 * kRN;  // 128 columns: one float4 per lane across a warp

// column of patch element `cidx` for a thread whose first column is jbase: groups of 4 adjacent
// columns, successive groups 128 columns apart (so every float4 store of a warp is one 512-byte run)
__device__ __forceinline__ int patch_col(int jbase, int cidx) {
  return jbase + (cidx >> 2) * kGroupCols + (cidx & 3);
}
// ... or, for a policy whose thread owns kCols ADJACENT columns (one 32-byte store per row)
template <bool ADJ>
__device__ __forceinline__ int col_at(int jbase, int cidx) {
  return ADJ ? jbase + cidx : patch_col(jbase, cidx);
}
constexpr int kObsChunk = 8192;   // floats per streamed observation chunk (32 KB)
constexpr int kObsSingleMax = 16384;  // up to this many (padded) observations live in smem at once
constexpr int kReorderMax = 1024;     // product mode: observations are re-ordered up to this count
constexpr int kMaxPeers = 16;         // GPUs whose partial vectors one finish kernel can sum

struct ScaleInfo {
  double log2s;  // product mode: log2 of the per-factor rescale s
  double shift;  // log-space mode: added to every cell's log2-likelihood
  double m2;     // log2 of the largest likelihood on the grid (closed form, scale choice only)
  double xbar;
  double ss;
  double smin;
  double mu_c;   // log-space: grid value nearest the sample mean; the stage holds x_k - mu_c
  double s_c;    // log-space: sum_k (x_k - mu_c)^2
  double umax;   // bound on |t_(i+1,j,k) - t_(i,j,k)| over adjacent mu rows, whole grid
  double ustep;  // the same bound per unit |a_j|: umax of a column band = max |a_j| in it * ustep
  double seps;   // relative sigma step |dsigma| / min |sigma| of the whole grid (slope policies' eligibility)
  double wide;   // 1 / 2 when the step ran the 8-column wide / tall product policy (reporting only)
};

struct RangeArgs {
  int N;
  float mu0, mu1, sg0, sg1;
};

// Work decomposition of the likelihood kernel.  A GROUP is Policy::kRows rows of one warp-wide
// column STRIP (32 * kCols columns); groups are numbered strip-major, u = strip * groups + group.
// The grid is persistent: ONE CTA PER SM (16 warps for the default product and log-space kernels).
// Every CTA owns an equal, contiguous share of the groups -- its POOL -- and the worker is the
// WARP, not the CTA: the warps of a CTA take `chunk` consecutive groups at a time from the pool
// through a shared-memory counter (one claim ahead, so the atomic's latency is hidden) and never
// synchronise after the set-up.  Measured reasons (B200, 16384^2 x 50; %globaltimer timelines in
// profiles/r02_cta_timeline_*.log):
//   * the warp scheduler does not share an SM fairly between its warps, so equal static shares per
//     warp leave some warps finished early and the rest running at lower occupancy; a counter in
//     shared memory balances the warps of a CTA at the cost of one ~100-cycle atomic per claim;
//   * the same holds between CTAs: three 8-warp CTAs per SM with a pool each finished at 0.89,
//     1.37 and 1.86 ms (mean warp residency 0.74); one CTA per SM with a single pool keeps every
//     warp busy to the end (all CTAs within 1.1 % of each other, residency 0.98);
//   * claims from per-strip counters in GLOBAL memory (the first design) cost an L2 round trip each,
//     and the warp could not hide it: with one group per claim the kernel took 2.48 ms, and small
//     shards (2048 rows, one of 8 GPUs) ran at 63 % of the full grid's rate;
//   * warps that move in lock step all reach the epilogue (stores, shuffles, float64) together and
//     all return to the pipe-bound loop together; independent warps drift apart and the epilogue
//     of one hides behind the arithmetic of the others;
//   * leaving warps out so that a pool divides evenly among the rest does not pay: throughput
//     falls by ~1.5 % per missing warp (profiles/r02_negative_working_warps_per_pool.jsonl).
//
// Results must not depend on who computed what.  Cells and partial row sums go to fixed places.
// Column sums cross pools, so they are accumulated EXACTLY: every thread's kRows-row float32 sum is
// scaled by a power of two chosen per column -- from the closed-form largest cell of that column,
// colmax_j = 2^(a_j S_min + n c0_j + shift), so that the shard's total stays below 2^62 -- rounded
// to an integer and added in int64: per warp in shared memory, then once per strip visit with
// integer atomics into colacc[N].  Integer addition is associative: the sums are bit-identical
// whatever the order, with a resolution of colmax_j * 2^-(57 - log2(rows/kRows)), i.e. ~2^-45
// relative to the column's own largest cell at 16384 rows (finer than the float64 partial sums it
// replaces for the columns that matter, and scaled per column, so the far tails keep their
// relative precision).
struct LikeArgs {
  // Observations: fused set-up (scale_in == NULL): the caller's n floats, any alignment, device
  // or mapped host memory; every CTA stages them and derives the statistics itself.  Streamed
  // (n_pad > kObsSingleMax): the padded (centred) copy setup_kernel made, with its statistics.
  const float *obs;
  const ScaleInfo *scale_in;
  ScaleInfo *scale_out;  // fused set-up: CTA 0 leaves the statistics here (fold kernel, cge_scale_info)
  RangeArgs range;
  int centred;   // stage x_k - mu_c (log-space tiled kernels)
  int reorder;   // product mode: stage the observations in a range-balanced order (setup_scale)
  // optional prior (NULL = flat, the reference's case): separable factors indexed by the global
  // mu row / sigma column, and/or a full table laid out like this shard of the posterior
  const float *prior_mu;
  const float *prior_sigma;
  const float *prior_full;
  float prior_log2max;  // log2 of the largest |prior factor product| (0 for the flat prior)
  float *post;      // shard base (row `row_begin` at offset 0)
  double *rowpart;  // [ncw][rows_local]
  unsigned long long *colacc;  // [N] fixed-point column sums (zero at launch; the fold kernel clears them)
  int *strip_next;  // streamed only: [ncw] next unclaimed item of each strip (zero at launch; the fold
                    // kernel clears them)
  unsigned int *tickets;  // "last block" tickets of the fold [0] and finish [1] kernels of this step:
                          // zeroed here, so a step that failed half-way cannot poison the next one
  int n, n_pad, N;  // n observations; n_pad staged floats (multiple of 4)
  int row_begin, row_end;
  int groups;          // row groups (Policy::kRows rows each) that cover the shard's rows
  int chunk;           // row groups per claim (at most 32 / kRows: one range value per lane)
  int items_per_strip; // streamed only: ceil(groups / (8 * chunk)) items per strip, claimed per CTA
  int fixed_bits;      // 58 - ceil(log2(addends per column)) (- 1 for 6-row groups): see col_fixed_shift
  int ncw;             // column strips per row: ceil(N / (32 * Policy::kCols))
  int vec4;            // N % 4 == 0 and the shard base is 16-byte aligned: float4 stores
  int keep_l2;         // the shard fits in L2: store with the default policy instead of .cs
  int vec8;            // N % 8 == 0 and the shard base is 32-byte aligned: 32-byte stores (wide policy)
  int wide;            // the 8-column policies of the product kernel: 0 = where eligible (tall, else wide),
                       // -1 = never, -2 = never the tall one, 1 = wide forced, 2 = tall forced
#ifdef CGE_DEBUG_TIMES
  unsigned long long *debug;  // per CTA: {start ns, end ns, smid, set-up done ns}, then 32 slots per CTA: warp end ns
#endif
};

__device__ __forceinline__ double coef_a(double sd) { return -0.5 * 1.4426950408889634074 / (sd * sd); }
__device__ __forceinline__ double coef_c0(double sd) { return -log2(sd * 2.5066282746310005024); }

// Fixed-point exponent of column j: its kRows-row sums are accumulated as rint(x * 2^s).  The cells
// of the column are at most 2^e, e = ceil(log2 colmax_j + log2 prior_max) + 1 (one binade of slack
// for the float32 rounding of the product); a 4-row sum is below 2^(e+2) (a 6-row sum below
// 2^(e+3): the host takes one more bit off fixed_bits), and with at most 2^(58 - fixed_bits)
// addends the int64 total stays below 2^62 for s = fixed_bits - e - 2.  The
// clamp keeps 2^s a float32 normal: columns whose largest cell is below ~2^-70 keep an absolute
// resolution of 2^-120 (their cells flush to zero at 2^-126 anyway).  Evaluated with identical
// arithmetic by the likelihood kernel (to scale) and the fold kernel (to scale back).
__device__ __forceinline__ int col_fixed_shift(double a, double c0, const ScaleInfo &sc, int n,
                                               float prior_log2max, int fixed_bits) {
  // (explicit roundings: the two kernels must get the same bits whatever the compiler contracts)
  const double t = __dadd_rn(__dadd_rn(__fma_rn(a, sc.smin, __dmul_rn((double)n, c0)), sc.shift),
                             (double)prior_log2max);
  const int e = t > -2000.0 ? (t < 2000.0 ? (int)ceil(t) + 1 : 2001) : -2000;  // (NaN -> -2000)
  return max(-120, min(120, fixed_bits - e - 2));
}

// Column coefficients in float64: a_j = -0.5*log2(e)/sigma_j^2, c0_j = -log2(sigma_j*sqrt(2 pi)),
// sigma_j from the reference's float32 range formula.  A float64 divide and log2 per column made a
// strip set-up ~1700 instructions per warp (1.4 % of the 16384^2 kernel, 11 % on a 2048-row shard,
// 22 % at 4096^2), so columns come in GROUPS of G neighbours (4, or 8 for the adjacent layouts):
// the group's first column is evaluated exactly and the others from it,
//   sigma_k = sigma_0 (1 + d):  1/sigma_k = (1/sigma_0) / (1 + d),  log2 sigma_k = log2 sigma_0 + log2(1 + d)
// with 8-term series in d (|d| <= 0.01: truncation below 1e-19; coarser groups take the exact
// path).  Every kernel that needs a column's coefficients (the likelihood kernel, and the fold
// kernels for the fixed-point exponent) calls these same functions with the same G, and every
// operation is an explicit intrinsic, so they all get the same bits.
struct ColumnBase {
  double sd, rinv, l2;  // sigma_0, 1/sigma_0, log2(sigma_0 sqrt(2 pi))
};
__device__ __forceinline__ ColumnBase column_base(const RangeArgs &r, int jb) {
  ColumnBase b;
  b.sd = (double)range_value(min(jb, r.N - 1), r.N, r.sg0, r.sg1);
  b.rinv = __ddiv_rn(1.0, b.sd);
  b.l2 = log2(__dmul_rn(b.sd, 2.5066282746310005024));
  return b;
}
__device__ __forceinline__ void column_from_base(const RangeArgs &r, int j, const ColumnBase &b,
                                                 double &a, double &c0) {
  const double sd = (double)range_value(min(j, r.N - 1), r.N, r.sg0, r.sg1);
  const double d = __fma_rn(sd, b.rinv, -1.0);
  double rinv, l2;
  if (fabs(d) <= 0.01) {
    // 1/(1+d) = 1 - d + d^2 - ... + d^8;  ln(1+d) = d (1 - d/2 + d^2/3 - ... - d^7/8)
    double q = 1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) q = __fma_rn(-d, q, 1.0);
    rinv = __dmul_rn(b.rinv, q);
    double p = -1.0 / 8.0;
    p = __fma_rn(p, d, 1.0 / 7.0);
    p = __fma_rn(p, d, -1.0 / 6.0);
    p = __fma_rn(p, d, 1.0 / 5.0);
    p = __fma_rn(p, d, -1.0 / 4.0);
    p = __fma_rn(p, d, 1.0 / 3.0);
    p = __fma_rn(p, d, -1.0 / 2.0);
    p = __fma_rn(p, d, 1.0);
    l2 = __fma_rn(__dmul_rn(p, d), 1.4426950408889634074, b.l2);
  } else {
    rinv = __ddiv_rn(1.0, sd);
    l2 = log2(__dmul_rn(sd, 2.5066282746310005024));
  }
  a = __dmul_rn(-0.5 * 1.4426950408889634074, __dmul_rn(rinv, rinv));
  c0 = -l2;
}
// one column on its own (fold kernels): G = neighbours per group of the policy that ran
__device__ __forceinline__ void column_coef(const RangeArgs &r, int j, int G, double &a, double &c0) {
  const ColumnBase b = column_base(r, j - j % G);
  column_from_base(r, j, b, a, c0);
}

// per-thread coefficients of a policy's columns (groups of 4 neighbours 128 columns apart, or KC
// adjacent columns)
template <int KC, bool ADJ = false>
struct ColumnCoefs {
  static constexpr int kGroup = ADJ ? KC : 4;
  double a[KC], c0[KC];
  __device__ __forceinline__ void compute(const RangeArgs &r, int j0) {
#pragma unroll
    for (int gb = 0; gb < KC; gb += kGroup) {
      const ColumnBase b = column_base(r, col_at<ADJ>(j0, gb));  // (j0 is a multiple of kGroup)
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        const int j = col_at<ADJ>(j0, gb + k);
        double av, cv;
        column_from_base(r, j, b, av, cv);
        a[gb + k] = j < r.N ? av : 0.0;
        c0[gb + k] = j < r.N ? cv : 0.0;
      }
    }
  }
};

// mean a_j of the two middle columns of a thread's KC neighbouring columns (layouts whose columns
// are j0 .. j0 + KC - 1), each replaced by the last column inside the grid when it lies outside
template <int KC, bool ADJ>
__device__ __forceinline__ double middle_slope(const ColumnCoefs<KC, ADJ> &cf, int j0, int N) {
  double lo = cf.a[0], hi = cf.a[0];
#pragma unroll
  for (int cidx = 1; cidx <= KC / 2; ++cidx) {
    const bool in = j0 + cidx < N;
    if (cidx <= KC / 2 - 1 && in) lo = cf.a[cidx];
    if (in) hi = cf.a[cidx];
  }
  return 0.5 * (lo + hi);
}

// ---------------------------------------------------------------------------------------
// setup_scale: block-wide (every thread of the block calls it, every thread gets the result).
// Stages the observations into `stage` (shared, or global for the streamed path) -- zero-padded
// to n_pad, raw or centred (x_k - mu_c), optionally re-ordered -- and derives the scale statistics.
//
// The rescale only positions the float32 dynamic range; it cancels in the posterior.  It needs
// max_ij log2 L_ij = max_j (a_j * S_min + n*c0_j) with S_min = min_i [SS + n (xbar - mu_i)^2].
// S(mu) is a parabola and f(sigma) = a(sigma) S + n c0(sigma) is unimodal (peak at
// sigma^2 = S/n), so the grid maxima sit next to the continuous ones: a handful of candidate
// grid points (the neighbours of the continuous optimum plus both ends) is exact, O(1).
//
// Re-ordering (product mode, n <= kReorderMax, `sorted` = n floats of shared scratch): the global
// rescale pins the FINAL product of the peak cell at 2^40, but a float32 running product also has
// to stay in range on the way there, and that depends on the order of the factors -- ascending
// input at n = 250 overflows 2^127 on the README grid, input sorted by |x - mean| at n = 300
// flushes live cells through 2^-126 (the reference multiplies in float64 and does not care).  The
// stage therefore holds the observations sorted and then visited with a golden-ratio stride,
// rank(p) = p*s mod n, s ~ 0.618 n coprime to n: every prefix is spread evenly over the sample's
// quantiles, so for every cell the running product follows its own geometric mean to within a few
// binades whatever order the caller supplied (a product does not depend on the order of its
// factors, so this changes rounding only).  `scratch` holds >= 32 doubles of shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ ScaleInfo setup_scale(const RangeArgs &r, const float *__restrict__ obs,
                                                 float *stage, int n, int n_pad, int centred,
                                                 float *sorted, double *scratch) {
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
  const double ss = block_sum(q, scratch);  // (its barriers also make the whole stage visible)
  if (centred)
    for (int k = tid; k < n; k += nt) stage[k] = (float)((double)stage[k] - mu_c);
  if (sorted != nullptr && n >= 3) {
    __syncthreads();  // (the centred values, if any, are visible to every thread)
    // rank by counting (ties broken by index): O(n^2 / threads), no barrier inside
    for (int k = tid; k < n; k += nt) {
      const float x = stage[k];
      int rank = 0;
      for (int m = 0; m < n; ++m) {
        const float y = stage[m];
        rank += (y < x || (y == x && m < k)) ? 1 : 0;
      }
      sorted[min(rank, n - 1)] = x;  // (NaN input collides here; Z is NaN then anyway)
    }
    int stride = (int)(0.6180339887498949 * (double)n + 0.5);
    for (;; --stride) {  // largest stride <= 0.618 n coprime to n (1 always is)
      int a = n, b = stride;
      while (b) { const int t = a % b; a = b; b = t; }
      if (a == 1) break;
    }
    __syncthreads();
    for (int p = tid; p < n; p += nt) stage[p] = sorted[(int)(((long long)p * stride) % n)];
  }

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
  // the float32 rounding of the range values.  ProductNeighbour is only used below kPairedUmax.
  {
    const double mlo = fmin((double)mu0, (double)mu1), mhi = fmax((double)mu0, (double)mu1);
    const double reach = fmax(fmax(fabs(xhi - mlo), fabs(xlo - mlo)), fmax(fabs(xhi - mhi), fabs(xlo - mhi)));
    const double dmu = N > 1 ? (mhi - mlo) / (double)(N - 1) + 2.4e-7 * fmax(fabs(mlo), fabs(mhi)) : 0.0;
    const double sdmin = fmin((double)sg0, (double)sg1);
    out.ustep = dmu * 2.0 * reach;
    out.umax = fabs(coef_a(sdmin)) * out.ustep;
    const double sabs = fmin(fabs((double)sg0), fabs((double)sg1));
    out.wide = 0.0;
    out.seps = (N > 1 && sabs > 0.0 && (double)sg0 * (double)sg1 > 0.0)
                   ? fabs((double)sg1 - (double)sg0) / (double)(N - 1) / sabs
                   : (N > 1 ? 1e300 : 0.0);
  }
  __syncthreads();  // the stage is complete for every thread of the block
  return out;
}

// ---------------------------------------------------------------------------------------
// Arithmetic policies of the fused likelihood kernel.  Each owns the per-thread state for a
// kRows x kCols patch of cells and consumes the observations of one row group.
//   kSmemDoubles  float64 values per thread the policy keeps in shared memory (element k of
//                 thread t at own[k * NT]; `own` is already offset by the thread index)
// ---------------------------------------------------------------------------------------
template <class Policy>
__device__ __forceinline__ void consume_scalar_obs(Policy &pol, const float *xs, int count,
                                                   const float (&mu)[Policy::kRows]);

// float32 product path, every exponential on MUFU.EX2 (the MUFU-roofline case): per pdf-eval one
// FFMA (exponent), one MUFU.EX2, one FMUL (running product); the subtract and square are shared
// by the kRN columns of a row.
//   factor_k = 2^(a_j*(x_k-mu_i)^2 + c_j),  c_j = c0_j + log2 s   (constants FMA-folded)
struct ProductF32 {
  static __device__ __forceinline__ bool eligible(double, double) { return true; }
  static constexpr double kUmax = 1e300;
  static constexpr bool kColSmem = false;
  static constexpr bool kAdjacent = false;
  static constexpr int kRows = kRowsProduct;
  static constexpr int kCols = 4;
  static constexpr int kSmemDoubles = 0;
  static __device__ __forceinline__ void consume(ProductF32 &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    consume_scalar_obs(pol, xs, count, mu);
  }
  float a[kRN], c[kRN];
  float p[kRows][kRN];

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf, double *) {
#pragma unroll
    for (int cidx = 0; cidx < kRN; ++cidx) {
      const bool ok = j0 + cidx < g.N;
      a[cidx] = (float)cf.a[cidx];
      c[cidx] = ok ? (float)(cf.c0[cidx] + sc.log2s) : 0.0f;
    }
  }
  __device__ __forceinline__ void begin_rows() {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kRN; ++cidx) p[r][cidx] = 1.0f;
  }
  __device__ __forceinline__ void step(float x, const float (&mu)[kRows]) {
    float dd[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float d = x - mu[r];  // same f32 rounding as the reference's (x - mu), kernels.h:111
      dd[r] = d * d;
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kRN; ++cidx)
        p[r][cidx] *= ex2_approx(fmaf(dd[r], a[cidx], c[cidx]));
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *) const {
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
  static __device__ __forceinline__ bool eligible(double, double) { return true; }
  static constexpr double kUmax = 1e300;
  static constexpr bool kColSmem = false;
  static constexpr bool kAdjacent = false;
  static constexpr int kRows = 4;
  static constexpr int kCols = 8;
  static constexpr int kSmemDoubles = kCols;  // n*c0_j + shift of the thread's columns
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

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf, double *own) {
#pragma unroll
    for (int cidx = 0; cidx < kCols; ++cidx) {
      a[cidx] = cf.a[cidx];
      own[cidx * NT] = patch_col(j0, cidx) < g.N ? (double)g.n * cf.c0[cidx] + sc.shift : 0.0;
    }
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
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *own) const {
    return ex2_approx((float)(acc[r][cidx] + own[cidx * NT]));
  }
};

// ---------------------------------------------------------------------------------------
// log-space path, float32x2 tiles: the same sum of log2-pdf terms, one FMA per pdf-eval, but on
// the FP32 pipe (packed FFMA2, twice the FP64 pipe's rate) with the float64 work reduced to one
// add per cell and tile of observations.  What makes float32 accumulation accurate enough is
// centring: with mu_c the grid value nearest the sample mean, x' = x - mu_c, mu' = mu_i - mu_c,
//   (x_k - mu_i)^2 = (x_k - mu_c)^2 + e_ik,    e_ik = mu'_i (mu'_i - 2 x'_k)      (exact identity)
//   log2 L_ij = [a_j S_c + n c0_j] + sum_k a_j e_ik,        S_c = sum_k (x_k - mu_c)^2
// The bracket is float64, once per column (S_c from the set-up).  The per-eval terms a_j e_ik are
// small wherever the posterior is not negligible (|mu'| <= ~sqrt(66 / (n |a|)) for cells within
// 1e-20 of the peak), so float32 tile sums (TILE = 256 observations) folded into float64 lose
// ~7e-6 relative on a cell at n = 10^4 -- a plain float32 sum of a_j (x - mu)^2 would lose 1e-4.
// Per (row, pair of observations): one FFMA2 (e = x' * (-2 mu') + mu'^2 of both) shared by the 8
// columns, then 8 FFMA2 -- an all-packed stream, 72 FFMA2 per 4 observations.
// ---------------------------------------------------------------------------------------
template <int TILE = 256>
struct LogSpaceTiled {
  static __device__ __forceinline__ bool eligible(double, double) { return true; }
  static constexpr double kUmax = 1e300;
  static constexpr bool kColSmem = false;
  static constexpr bool kAdjacent = false;
  static constexpr int kRows = 4;
  static constexpr int kCols = 8;
  static constexpr int kSmemDoubles = kCols;  // a_j S_c + n c0_j + shift of the thread's columns
  static constexpr int kTile = TILE;  // observations per float32 tile sum
  static constexpr int kPairs = kCols / 2;
  f32x2 a2[kPairs];
  double acc[kRows][kCols];
  double mu_c_;
  bool vec_;  // always true, but only known at run time (see consume)

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf, double *own) {
    mu_c_ = sc.mu_c;
    vec_ = g.n >= 0;
#pragma unroll
    for (int h = 0; h < kPairs; ++h) a2[h] = pack2((float)cf.a[2 * h], (float)cf.a[2 * h + 1]);
#pragma unroll
    for (int cidx = 0; cidx < kCols; ++cidx)
      own[cidx * NT] = patch_col(j0, cidx) < g.N
                           ? cf.a[cidx] * sc.s_c + (double)g.n * cf.c0[cidx] + sc.shift
                           : -1.0 / 0.0;
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
  // two observations: the per-row terms e of both as one packed FMA (an all-packed instruction
  // stream: 72 FFMA2 per 4 observations, measured 1.6 % faster than 16 FFMA + 64 FFMA2)
  __device__ __forceinline__ void step2(f32x2 xc2, const float (&gm)[kRows], const float (&gb)[kRows],
                                        f32x2 (&t2)[kRows][kPairs]) {
    float e0[kRows], e1[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      unpack2(fma2(xc2, pack2(gm[r], gm[r]), pack2(gb[r], gb[r])), e0[r], e1[r]);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const f32x2 ea = pack2(e0[r], e0[r]);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) t2[r][h] = fma2(ea, a2[h], t2[r][h]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const f32x2 eb = pack2(e1[r], e1[r]);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) t2[r][h] = fma2(eb, a2[h], t2[r][h]);
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
      // The rows are warp-uniform, and ptxas keeps warp-uniform scalars in uniform registers --
      // from where each FMA below (two such operands) needs a MOV first: 12 of the loop's 99
      // instructions at the time.  A select against a lane-dependent value (never taken) makes them ordinary
      // vector registers.
      const float lane_val = __int_as_float((int)threadIdx.x);
      gm[r] = pol.vec_ ? -2.0f * m : lane_val;
      gb[r] = pol.vec_ ? m * m : lane_val;
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
        pol.step2(pack2(x.x, x.y), gm, gb, t2);
        pol.step2(pack2(x.z, x.w), gm, gb, t2);
      }
      pol.fold(t2);
    }
    const int rest = ntiles * kTile;
    if (rest < count) {
      for (int k = rest; k < count; ++k) pol.step(xs[k], gm, gb, t2);
      pol.fold(t2);
    }
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *own) const {
    return ex2_approx((float)(acc[r][cidx] + own[cidx * NT]));  // 2^-inf = 0 outside the grid
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
// float32 product path, hybrid exponential (coarse grids): the same per-pdf-eval work as
// ProductF32 (exponent FMA, 2^t, running product) but
//   * the per-cell arithmetic runs as packed f32x2 (FFMA2/FMUL2): a thread's 4 columns are two
//     register pairs; the per-row (x - mu)^2 stays scalar and enters FFMA2 as a broadcast operand;
//   * of the 12 column pairs a thread updates per observation, 3 take 2^t from the degree-5
//     FMA-pipe polynomial (ex2_poly2) instead of MUFU.EX2.  With MUFU-only code the XU pipe is
//     96 % busy while the FMA pipe idles at 30 % and half the issue slots are free
//     (profiles/r01_ncu_full_mufu_only_16384x50.csv); moving a quarter of the exponentials over
//     balances the two pipes.
// Which cells use which pipe is fixed at compile time, so results stay bit-reproducible.
// ---------------------------------------------------------------------------------------
struct ProductHybrid {
  static __device__ __forceinline__ bool eligible(double, double) { return true; }
  static constexpr double kUmax = 1e300;  // any grid
  static constexpr bool kColSmem = false;
  static constexpr bool kAdjacent = false;
  static constexpr int kRows = kRowsProduct;
  static constexpr int kCols = 4;
  static constexpr int kSmemDoubles = 0;
  static constexpr int kPoly = 3;  // (row, pair) slots per observation on the polynomial
  f32x2 a2[kRN / 2], c2[kRN / 2];
  f32x2 p2[kRows][kRN / 2];

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf, double *) {
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
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kRN / 2; ++h) p2[r][h] = one;
  }
  __device__ __forceinline__ void step(float x, const float (&mu)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float d = x - mu[r];  // same f32 rounding as the reference's (x - mu), kernels.h:111
      const float dd = d * d;
      const f32x2 dd2 = pack2(dd, dd);  // becomes a scalar-broadcast operand of FFMA2
#pragma unroll
      for (int h = 0; h < kRN / 2; ++h) {
        const f32x2 t2 = fma2(dd2, a2[h], c2[h]);
        // the polynomial slots, spread over the rows and both pairs: (r, h) = (0,0), (2,1), (4,0)
        const bool poly = (r == 0 && h == 0) || (r == 2 && h == 1) || (r == 4 && h == 0);
        const f32x2 e2 = poly ? ex2_poly2(t2) : ex2_mufu2(t2);
        p2[r][h] = mul2(p2[r][h], e2);
      }
    }
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductHybrid &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
#pragma unroll 1
    for (int k = 0; k < count; ++k) pol.step(xs[k], mu);
  }
};

// ---------------------------------------------------------------------------------------
// float32 product path, neighbour rows: of every three vertically adjacent cells (mu_(m-1), mu_m,
// mu_(m+1) at the same sigma_j) the middle one takes its factor from MUFU.EX2 and the two others
// derive theirs from it,
//   pdf(x_k; mu_r, sigma_j) = pdf(x_k; mu_m, sigma_j) * 2^u,
//   u = a_j * g_rk,   g_rk = (mu_r - mu_m) * (mu_m + mu_r - 2 x_k)      (exact identity, r = m -+ 1)
// with 2^u - 1 from a degree-2 polynomial on the FMA pipe: on a fine grid |u| is tiny (6e-3 at
// 16384 rows for the README ranges), so no range reduction is needed.  Every pdf-eval still forms
// its own factor and multiplies it into its own running product; what changes is the pipe mix:
//   MUFU row   : FFMA2 (exponent), 2 MUFU.EX2, FMUL2 (product)            per column pair
//   derived row: FFMA2 (a_j k1 + g a_j^2 k2), FMUL2 (* g), FFMA2 (e + e v), FMUL2 (product)
// i.e. per 6 evals 2 MUFU.EX2 and 10 packed FMA-pipe instructions (20 lane-ops).  History of the
// mix on B200 (16 MUFU lanes and 128 FMA lanes per SM and clock):
//   every exponential on MUFU         1 MUFU + 2 lane-ops per eval   XU-bound   (r01: XU 97 %)
//   pairs (one derived row of two)  1/2 MUFU + 3 lane-ops            XU-bound, FMA 84 % of XU time:
//                                    both pipes throttle each other, XU stuck at 79-87 %
//   triples (two derived of three)  1/3 MUFU + 3.33 lane-ops         FMA-bound, XU at 2/3 of it
// (+ the per-row scalars).  The pipe micro-benchmark (tools/pipe_mix.py) puts the triple mix at
// 32 T lane-ops/s with 24 warps per SM (88 % of the FFMA peak) against 28 T for the pair mix.
// Polynomial: 2^u = 1 + u (k1 + k2 u) with k2 = ln(2)^2/2 and the cubic term folded into
// k1 = ln 2 + (ln(2)^3/6)(3/4) U^2 (Chebyshev economisation on [-U, U], U = the strip's umax):
// error <= 0.0139 U^3 = 1.8e-8 at the eligibility bound kPairedUmax, below the float32 rounding
// of the factor itself.  The derived factor is e*(1+v) evaluated as fma(e, v, e): one rounding.
// DEG = 3 adds the cubic term, 2^u - 1 = u (k1 + u (k2 + u k3)), k3 = ln(2)^3/6 and the quartic
// folded into k2 = ln(2)^2/2 + (ln(2)^4/24) U^2: error <= 0.0024 U^4 = 1.5e-8 at its bound
// kPairedUmax3 = 0.05, which takes grids 4.5x coarser (README ranges: N >= 690) for one more
// FFMA2 per derived pair.  Grids too coarse for that run ProductHybrid instead (chosen on the
// device per column strip); which rows are MUFU rows depends on the global row index only
// (row % 3 == 1).  Four observations per loop iteration (one LDS.128), the per-row scalars of the
// two triples packed.
// ---------------------------------------------------------------------------------------
constexpr int kObsLoopUnroll = 1;  // iterations (of 4 observations) per unrolled loop body
constexpr double kPairedUmax = 0.011;
constexpr double kPairedUmax3 = 0.05;

// the observation loop of the neighbour-row policies: the per-row scalars of the group
// (Policy::Rows), then four observations per iteration (one LDS.128)
template <class Policy>
__device__ __forceinline__ void neighbour_consume(Policy &pol, const float *xs, int count,
                                                  const float (&mu)[Policy::kRows]) {
  const typename Policy::Rows rs = pol.rows(mu);
  const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
  const int n4 = count >> 2;
#pragma unroll kObsLoopUnroll
  for (int q = 0; q < n4; ++q) {
    const float4 x = xs4[q];
    pol.step(x.x, rs);
    pol.step(x.y, rs);
    pol.step(x.z, rs);
    pol.step(x.w, rs);
  }
  if (count & 2) {  // the tail as one LDS.64 and at most one scalar load
    const float2 x = *reinterpret_cast<const float2 *>(xs + (n4 << 2));
    pol.step(x.x, rs);
    pol.step(x.y, rs);
  }
  if (count & 1) pol.step(xs[count - 1], rs);
}

template <int DEG = 2>
struct ProductNeighbour {
  static_assert(DEG == 2 || DEG == 3, "polynomial degree of 2^u - 1");
  static constexpr double kUmax = DEG == 2 ? kPairedUmax : kPairedUmax3;  // eligibility bound
  static __device__ __forceinline__ bool eligible(double umax, double) { return umax <= kUmax; }
  static constexpr bool kColSmem = true;
  static constexpr bool kAdjacent = false;
  static constexpr int kRows = kRowsProduct;
  static constexpr int kCols = 4;
  static constexpr int kSmemDoubles = 0;
  static constexpr int kPairs = kCols / 2;
  static constexpr int kTriples = kRows / 3;
  static_assert(kRows == 6, "two triples: the per-row scalars are packed pairwise");
  f32x2 a2[kPairs], c2[kPairs];    // exponent coefficients of the MUFU rows
  f32x2 k1a[kPairs], k2a[kPairs];  // a_j k1, a_j^2 k2 of the derived rows
  f32x2 k3a[DEG == 3 ? kPairs : 1];  // a_j^3 k3
  f32x2 p2[kRows][kPairs];

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols> &cf, double *) {
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
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) p2[r][h] = one;
  }
  // the per-row scalars of a group: mid2 = {mu_1, mu_4} (the MUFU rows); gmL/gbL give
  // g = x*gm + gb of the rows below them (0, 3), gmU/gbU of the rows above (2, 5), each as one
  // packed pair over the triples.  g of row r against its triple's middle row m:
  // (mu_r - mu_m) * (mu_m + mu_r - 2 x)
  struct Rows {
    f32x2 mid2, gmL, gbL, gmU, gbU;
  };
  __device__ __forceinline__ Rows rows(const float (&mu)[kRows]) const {
    float gm[kRows], gb[kRows];
#pragma unroll
    for (int t = 0; t < kTriples; ++t)
#pragma unroll
      for (int s = 0; s < 3; s += 2) {
        const int r = 3 * t + s, m = 3 * t + 1;
        const float delta = __fsub_rn(mu[r], mu[m]);  // exact (adjacent range values)
        gm[r] = -2.0f * delta;
        gb[r] = __fmul_rn(delta, __fadd_rn(mu[m], mu[r]));
      }
    Rows rs;
    rs.mid2 = pack2(mu[1], mu[4]);
    rs.gmL = pack2(gm[0], gm[3]);
    rs.gbL = pack2(gb[0], gb[3]);
    rs.gmU = pack2(gm[2], gm[5]);
    rs.gbU = pack2(gb[2], gb[5]);
    return rs;
  }
  // one observation
  __device__ __forceinline__ void step(float x, const Rows &rs) {
    float ddv[kTriples], glo[kTriples], gup[kTriples];
    {
      const f32x2 x2 = pack2(x, x);
      const f32x2 d2 = sub2(x2, rs.mid2);  // same f32 rounding as the reference's (x - mu)
      unpack2(mul2(d2, d2), ddv[0], ddv[1]);
      unpack2(fma2(x2, rs.gmL, rs.gbL), glo[0], glo[1]);
      unpack2(fma2(x2, rs.gmU, rs.gbU), gup[0], gup[1]);
    }
#pragma unroll
    for (int t = 0; t < kTriples; ++t) {
      const f32x2 dd2 = pack2(ddv[t], ddv[t]);  // (scalar-broadcast operands of the packed instructions)
      const f32x2 gl2 = pack2(glo[t], glo[t]);
      const f32x2 gu2 = pack2(gup[t], gup[t]);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        const f32x2 e2 = ex2_mufu2(fma2(dd2, a2[h], c2[h]));
        p2[3 * t + 1][h] = mul2(p2[3 * t + 1][h], e2);
        {
          const f32x2 q2 = DEG == 3 ? fma2(gl2, k3a[h], k2a[h]) : k2a[h];
          const f32x2 v2 = mul2(gl2, fma2(gl2, q2, k1a[h]));  // 2^u - 1
          const f32x2 f2 = fma2(e2, v2, e2);                   // the lower neighbour's factor
          p2[3 * t][h] = mul2(p2[3 * t][h], f2);
        }
        {
          const f32x2 q2 = DEG == 3 ? fma2(gu2, k3a[h], k2a[h]) : k2a[h];
          const f32x2 v2 = mul2(gu2, fma2(gu2, q2, k1a[h]));
          const f32x2 f2 = fma2(e2, v2, e2);                   // the upper neighbour's factor
          p2[3 * t + 2][h] = mul2(p2[3 * t + 2][h], f2);
        }
      }
    }
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductNeighbour &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    neighbour_consume(pol, xs, count, mu);
  }
};

// ---------------------------------------------------------------------------------------
// float32 product path, neighbour rows with a shared slope (the finest grids): the same triples as
// ProductNeighbour -- the middle row's factor e from MUFU.EX2, the rows above and below derive
// theirs, f = e * 2^u, u = a_j g_rk -- but the polynomial is evaluated once per (row, observation)
// instead of once per cell.  With a_0 the mean |a_j| slope of the thread's four adjacent columns,
//   2^u - 1 = a_j g (k1 + k2 a_j g) = a_j H_rk + a_j (a_j - a_0) k2 g^2,   H_rk = g (k1 + k2 a_0 g),
// and the last term is dropped.  a = -0.72/sigma^2, so a column `s` steps from the one a_0 belongs to
// has |a_j - a_0| / |a_j| <= eps(s) = 2 x (1 + 1.5 x), x = s |dsigma| / sigma_min, and the dropped
// term is at most k2 U^2 eps (1e-9 on the 16384^2 README grid).  The policy is eligible while that
// bound stays below kSlopeErr = 2^-24 -- the float32 rounding unit of the factor itself -- and
// U <= kPairedUmax (cubic term, as ProductNeighbour<2>); coarser strips fall back to
// ProductNeighbour.  Every pdf-eval still forms its own factor and multiplies it into its
// own running product:
//   MUFU row   : FFMA2 (exponent), 2 MUFU.EX2, FMUL2 (product)                  per column pair
//   shared     : FMUL2 (ea = e * a_j)                                           per triple and pair
//   derived row: FFMA2 (f = ea * H + e), FMUL2 (product)
// i.e. per 6 evals 2 MUFU.EX2 and 7 packed instructions instead of 10, plus 6 packed per-row
// scalars per observation (d, d^2 and the two H, each a Horner quadratic in d).  Per thread and
// observation: 34 packed FMA-pipe instructions (68 pipe cycles per warp) against 8 MUFU.EX2 (64 XU
// cycles).
// ---------------------------------------------------------------------------------------
//
// KC = 8 is the WIDE form: a thread owns 8 ADJACENT columns (one 32-byte store per row; the warp's
// strip is 256 columns = two row-partial slots), so the 6 packed row scalars are shared by twice
// the cells -- 62 packed + 16 MUFU.EX2 per 48 evals instead of 68 + 16 -- and the per-group
// epilogue (row sums, claims, range values) is paid once per 48 cells.  a_0 is then up to 3.5
// column steps from a thread's outer columns instead of 1.5.
constexpr double kSlopeErr = 5.9604644775390625e-8;  // 2^-24
template <int KC>
struct ProductSlopeT {
  static_assert(KC == 4 || KC == 8, "4 columns (two groups' worth of one float4) or 8 adjacent ones");
  static constexpr double kUmax = kPairedUmax;
  static constexpr double kSteps = KC == 8 ? 3.5 : 1.5;  // columns between a_0 and the outermost one
  // rstep: the strip's (or grid's) relative sigma step (strip_sigma_step / ScaleInfo::seps); 1 %
  // on top for the float32 rounding of the range values
  static __device__ __forceinline__ bool eligible(double umax, double rstep) {
    const double x = kSteps * rstep * 1.01;
    return umax <= kPairedUmax && 0.2402265069591007 * umax * umax * (2.0 * x * (1.0 + 1.5 * x)) <= kSlopeErr;
  }
  static constexpr bool kColSmem = true;
  static constexpr bool kAdjacent = KC == 8;
  static constexpr int kRows = kRowsProduct;
  static constexpr int kCols = KC;
  static constexpr int kSmemDoubles = 0;
  static constexpr int kPairs = kCols / 2;
  static constexpr int kTriples = kRows / 3;
  static_assert(kRows == 6, "two triples: the per-row scalars are packed pairwise");
  f32x2 a2[kPairs], c2[kPairs];  // exponent coefficients of the MUFU rows (a2 also scales ea)
  float k1, a0k2;                // k1 and kappa = a_0 k2
  f32x2 p2[kRows][kPairs];

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols, kAdjacent> &cf, double *) {
    const double U = sc.umax;
    const double k2 = 0.24022650695910071233;
    k1 = (float)(0.69314718055994530942 + 0.0555041086648215799 * 0.75 * U * U);
    // slope of the thread's columns: the mean of the two middle ones (of the last one inside the
    // grid for a thread that straddles its edge, so it is always a slope of its own columns)
    a0k2 = (float)(middle_slope(cf, j0, g.N) * k2);
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      float av[2], cv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = col_at<kAdjacent>(j0, 2 * h + e);
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
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) p2[r][h] = one;
  }
  // The per-row scalars of a group.  With d = x - mu_m (the MUFU row's own difference, rounded as
  // the reference rounds it) and delta = mu_r - mu_m, the neighbour's exponent offset is
  // g = delta (delta - 2 d) and H = g (k1 + kappa g) is a quadratic in d,
  //   H = [4 kappa delta^2] d^2 - [2 delta (k1 + 2 kappa delta^2)] d + [delta^2 (k1 + kappa delta^2)],
  // two packed FMAs per observation for the rows below (0, 3) and two for the rows above (2, 5),
  // with no cancellation (the linear term dominates).
  struct Rows {
    f32x2 mid2, aL, bL, cL, aU, bU, cU;
  };
  __device__ __forceinline__ Rows rows(const float (&mu)[kRows]) const {
    float qa[kRows], qb[kRows], qc[kRows];
#pragma unroll
    for (int t = 0; t < kTriples; ++t)
#pragma unroll
      for (int s = 0; s < 3; s += 2) {
        const int r = 3 * t + s, m = 3 * t + 1;
        const float delta = __fsub_rn(mu[r], mu[m]);  // exact (adjacent range values)
        const float kd2 = a0k2 * delta * delta;
        qa[r] = 4.0f * kd2;
        qb[r] = -2.0f * delta * fmaf(2.0f, kd2, k1);
        qc[r] = delta * delta * (k1 + kd2);
      }
    Rows rs;
    rs.mid2 = pack2(mu[1], mu[4]);
    rs.aL = pack2(qa[0], qa[3]);
    rs.bL = pack2(qb[0], qb[3]);
    rs.cL = pack2(qc[0], qc[3]);
    rs.aU = pack2(qa[2], qa[5]);
    rs.bU = pack2(qb[2], qb[5]);
    rs.cU = pack2(qc[2], qc[5]);
    return rs;
  }
  // one observation
  __device__ __forceinline__ void step(float x, const Rows &rs) {
    float ddv[kTriples], hlo[kTriples], hup[kTriples];
    {
      const f32x2 d2 = sub2(pack2(x, x), rs.mid2);  // same f32 rounding as the reference's (x - mu)
      unpack2(mul2(d2, d2), ddv[0], ddv[1]);
      unpack2(fma2(fma2(d2, rs.aL, rs.bL), d2, rs.cL), hlo[0], hlo[1]);  // H of rows 0, 3
      unpack2(fma2(fma2(d2, rs.aU, rs.bU), d2, rs.cU), hup[0], hup[1]);  // H of rows 2, 5
    }
#pragma unroll
    for (int t = 0; t < kTriples; ++t) {
      const f32x2 dd2 = pack2(ddv[t], ddv[t]);  // (scalar-broadcast operands of the packed instructions)
      const f32x2 hl2 = pack2(hlo[t], hlo[t]);
      const f32x2 hu2 = pack2(hup[t], hup[t]);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        const f32x2 e2 = ex2_mufu2(fma2(dd2, a2[h], c2[h]));
        p2[3 * t + 1][h] = mul2(p2[3 * t + 1][h], e2);
        const f32x2 ea2 = mul2(e2, a2[h]);
        p2[3 * t][h] = mul2(p2[3 * t][h], fma2(ea2, hl2, e2));          // the lower neighbour's factor
        p2[3 * t + 2][h] = mul2(p2[3 * t + 2][h], fma2(ea2, hu2, e2));  // the upper neighbour's factor
      }
    }
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductSlopeT &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    neighbour_consume(pol, xs, count, mu);
  }
};
using ProductSlope = ProductSlopeT<4>;
using ProductSlopeWide = ProductSlopeT<8>;

// ---------------------------------------------------------------------------------------
// float32 product path, TALL shared-slope policy (the finest grids): 7 rows x 8 adjacent columns
// per thread, ONE MUFU row (the middle one) and six derived rows at distances 1, 2 and 3 -- the
// same arithmetic as ProductSlopeWide with a longer neighbour chain:
//   per (observation, column pair): FFMA2 (exponent), 2 MUFU.EX2, FMUL2 (product), FMUL2 (ea = e a_j),
//                                   then FFMA2 (f = ea H_r + e) + FMUL2 (product) per derived row
//   per pair of observations     : d, d^2 and six Horner quadratics H_r(d) of both, packed
// i.e. 67 packed FMA-pipe instructions and 8 MUFU.EX2 per 56 evals: 0.143 MUFU and 2.39 FP32
// lane-ops per eval against 0.333 / 2.58 for the wide policy -- 150 issue cycles per 56 evals
// instead of 158 per 48.  The exponent offset of a row three away is three times as large, so the
// bounds are those of the wide policy with U -> 3 U: 3 umax <= kPairedUmax (cubic term of the
// degree-2 polynomial) and k2 (3 U)^2 eps(3.5 columns) <= 2^-24 -- README ranges: N >= ~9400.
// Every pdf-eval still forms its own factor and multiplies it into its own running product.
// ---------------------------------------------------------------------------------------
struct ProductSlopeTall {
  static constexpr double kUmax = kPairedUmax / 3.0;
  static __device__ __forceinline__ bool eligible(double umax, double rstep) {
    const double x = 3.5 * rstep * 1.01, u3 = 3.0 * umax;
    return u3 <= kPairedUmax && 0.2402265069591007 * u3 * u3 * (2.0 * x * (1.0 + 1.5 * x)) <= kSlopeErr;
  }
  static constexpr bool kColSmem = true;
  static constexpr bool kAdjacent = true;
  static constexpr int kRows = 7;
  static constexpr int kMid = 3;  // the MUFU row
  static constexpr int kCols = 8;
  static constexpr int kSmemDoubles = 0;
  static constexpr int kPairs = kCols / 2;
  f32x2 a2[kPairs], c2[kPairs];
  float k1, a0k2;
  f32x2 p2[kRows][kPairs];

  template <int NT>
  __device__ __forceinline__ void load_columns(const LikeArgs &g, int j0, const ScaleInfo &sc,
                                               const ColumnCoefs<kCols, true> &cf, double *) {
    const double U = 3.0 * sc.umax;  // the rows furthest from the MUFU row
    const double k2 = 0.24022650695910071233;
    k1 = (float)(0.69314718055994530942 + 0.0555041086648215799 * 0.75 * U * U);
    a0k2 = (float)(middle_slope(cf, j0, g.N) * k2);
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
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
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int h = 0; h < kPairs; ++h) p2[r][h] = one;
  }
  // per-row scalars of a group: H_r(d) = qa d^2 + qb d + qc (ProductSlopeT::rows), d = x - mu_mid
  struct Rows {
    float mid, qa[kRows], qb[kRows], qc[kRows];  // (index kMid unused)
  };
  __device__ __forceinline__ Rows rows(const float (&mu)[kRows]) const {
    Rows rs;
    rs.mid = mu[kMid];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float delta = __fsub_rn(mu[r], mu[kMid]);  // exact (nearby range values)
      const float kd2 = a0k2 * delta * delta;
      rs.qa[r] = 4.0f * kd2;
      rs.qb[r] = -2.0f * delta * fmaf(2.0f, kd2, k1);
      rs.qc[r] = delta * delta * (k1 + kd2);
    }
    return rs;
  }
  __device__ __forceinline__ void step(float x, const Rows &rs) {
    const float d = x - rs.mid;  // same f32 rounding as the reference's (x - mu)
    const float dd = d * d;
    float hv[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (r != kMid) hv[r] = fmaf(fmaf(d, rs.qa[r], rs.qb[r]), d, rs.qc[r]);
    const f32x2 dd2 = pack2(dd, dd);  // (scalar-broadcast operands of the packed instructions)
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      const f32x2 e2 = ex2_mufu2(fma2(dd2, a2[h], c2[h]));
      p2[kMid][h] = mul2(p2[kMid][h], e2);
      const f32x2 ea2 = mul2(e2, a2[h]);
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (r != kMid) p2[r][h] = mul2(p2[r][h], fma2(ea2, pack2(hv[r], hv[r]), e2));
    }
  }
  // two observations: the row scalars of both as packed FMAs -- an all-packed instruction stream
  // (measured 1.173 against 1.229 ms for the scalar form at 16384^2 x 50: mixing scalar and packed
  // FMAs costs issue slots)
  __device__ __forceinline__ void step2(f32x2 x2, const Rows &rs) {
    const f32x2 d2 = sub2(x2, pack2(rs.mid, rs.mid));
    float dd0, dd1;
    unpack2(mul2(d2, d2), dd0, dd1);
    float h0[kRows], h1[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (r != kMid)
        unpack2(fma2(fma2(d2, pack2(rs.qa[r], rs.qa[r]), pack2(rs.qb[r], rs.qb[r])), d2, pack2(rs.qc[r], rs.qc[r])),
                h0[r], h1[r]);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const float dd = o ? dd1 : dd0;
      const f32x2 dd2 = pack2(dd, dd);
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        const f32x2 e2 = ex2_mufu2(fma2(dd2, a2[h], c2[h]));
        p2[kMid][h] = mul2(p2[kMid][h], e2);
        const f32x2 ea2 = mul2(e2, a2[h]);
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (r != kMid) {
            const float hv = o ? h1[r] : h0[r];
            p2[r][h] = mul2(p2[r][h], fma2(ea2, pack2(hv, hv), e2));
          }
      }
    }
  }
  template <int NT>
  __device__ __forceinline__ float value(const LikeArgs &, int r, int cidx, const double *) const {
    float lo, hi;
    unpack2(p2[r][cidx >> 1], lo, hi);
    return (cidx & 1) ? hi : lo;
  }
  static __device__ __forceinline__ void consume(ProductSlopeTall &pol, const float *xs, int count,
                                                 const float (&mu)[kRows]) {
    const Rows rs = pol.rows(mu);
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const int n4 = count >> 2;
#pragma unroll 1
    for (int q = 0; q < n4; ++q) {
      const float4 x = xs4[q];
      pol.step2(pack2(x.x, x.y), rs);
      pol.step2(pack2(x.z, x.w), rs);
    }
    if (count & 2) {
      const float2 x = *reinterpret_cast<const float2 *>(xs + (n4 << 2));
      pol.step2(pack2(x.x, x.y), rs);
    }
    if (count & 1) pol.step(xs[count - 1], rs);
  }
};

// ---------------------------------------------------------------------------------------
// Shared-memory layout of the likelihood kernels (dynamic):
//   [0, 16)    two mbarriers (streamed observations only)
//   [16, 112)  ScaleInfo of this step (written once per CTA by the set-up)
//   [128, 160) claim slots: [0..2] the item a CTA takes next (streamed), [4] the pool counter
//   [160, ...) observation stage: n_pad floats, or two chunks of kObsChunk when streamed
//   then       per-thread slots [k][NT]: kCols int64 column accumulators (as 16-byte pairs), kCols
//              float32 column scales 2^s_j (as 8-byte pairs), the policy's own float64 values, and
//              kRows floats per thread of row-sum transposition buffer (kRows * 128 bytes per
//              warp); during the set-up the same bytes serve as the sort scratch
// ---------------------------------------------------------------------------------------
constexpr int kSmemHeader = 160;
static_assert(16 + sizeof(ScaleInfo) <= 128, "shared-memory header layout");
__host__ __device__ __forceinline__ size_t stage_bytes(int n_pad) {
  return (((size_t)(n_pad > kObsSingleMax ? 2 * kObsChunk : n_pad) * 4) + 15) & ~(size_t)15;
}
// bytes per thread behind the observation stage
template <class Policy>
constexpr int policy_smem_bytes() {
  return Policy::kCols * 8 + Policy::kCols * 4 + Policy::kSmemDoubles * 8 + (Policy::kRows == 7 ? 8 : Policy::kRows) * 4;
}

// thread 0 only: start the bulk copy of chunk `ch` of the streamed observations into stage `st`
__device__ __forceinline__ void issue_chunk(const LikeArgs &g, uint64_t *bars, float *obs_s, int ch,
                                            int st) {
  const int floats = min(kObsChunk, g.n_pad - ch * kObsChunk);
  const uint32_t bytes = (uint32_t)floats * 4u;
  mbar_arrive_expect_tx(&bars[st], bytes);
  bulk_copy_g2s(obs_s + (size_t)st * kObsChunk, g.obs + (size_t)ch * kObsChunk, bytes, &bars[st]);
}

constexpr int kWarps = 8;             // warps per CTA of every likelihood kernel
constexpr int kThreads = kWarps * 32;

// ---------------------------------------------------------------------------------------
// run_strip: everything this warp does in column strip `cw` -- K1-K5 fused.
//   The warp owns the strip's 32*kCols sigma columns (lane -> kCols/4 groups of 4 adjacent
//   columns, one 16-byte store per group and row) and walks row groups kRows rows at a time.  Per
//   row group it emits its partial row sums (fixed-order shuffle tree) and adds its kRows-row
//   column sums into its fixed-point accumulators (shared memory); when the strip is left they go
//   to colacc[] with one integer atomic per column.  No floating-point atomics.
//   Pooled (not STREAMED): [u0, u1) is the claim the warp holds (global group numbers, u0 inside
//   this strip); it works it off, takes further claims from the CTA's pool counter while they
//   start in this strip, and returns in [u0, u1) the first claim (or remainder of one) that does
//   not -- u0 >= pool_end when the pool is exhausted.
//   STREAMED (more than 16384 observations): the CTA moves as one -- u0 is the item of strip `cw`
//   it holds; thread 0 claims for all 8 warps from the strip's global counter; the warps take
//   interleaved groups of the same strip and consume each 32 KB chunk of observations together
//   (the chunks arrive by TMA bulk copy into two shared-memory stages).
// ---------------------------------------------------------------------------------------
template <class Policy, bool STREAMED, int NT>
__device__ __forceinline__ void run_strip(const LikeArgs &g, const ScaleInfo &sc, int cw, int &u0,
                                          int &u1, int pool_end, uint64_t *bars, float *obs_s,
                                          unsigned char *tsm_raw, volatile int *claim,
                                          int &claim_par, int &streams) {
  constexpr int kWarpsCta = NT / 32;
  constexpr int kRows = Policy::kRows;
  constexpr int kCols = Policy::kCols;
  constexpr int kGroups = kCols / kRN;
  constexpr bool kAdj = Policy::kAdjacent;
  // float offset between a thread's successive 4-column groups: adjacent, or 128 columns apart
  constexpr int kGroupStride = kAdj ? kRN : kGroupCols;
  static_assert(kRows == 4 || kRows == 6 || kRows == 7, "the epilogue is written for 4-, 6- and 7-row patches");
  static_assert(!kAdj || (kCols == 8 && kRows >= 6 && !STREAMED), "the adjacent layout is the wide product policies'");
  static_assert(kRows != 7 || kAdj, "7-row patches: the tall policy");
  constexpr int kBufRows = kRows == 7 ? 8 : kRows;  // rows of the transposition buffer
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // first column; group g starts 128*g further (adjacent layout: all kCols columns in a row)
  const int j0 = cw * (32 * kCols) + lane * (kAdj ? kCols : kRN);
  // row groups start at a multiple of kRows below row_begin, so a cell's position inside the
  // register patch (hence which exponential path it takes) depends on its GLOBAL row only: a row
  // shard computes bit-identical unnormalised cells to the single-GPU run
  const int row_base = g.row_begin - g.row_begin % kRows;
  const size_t rows_loc = (size_t)(g.row_end - g.row_begin);  // rowpart is [slot][row]: a warp's rows are contiguous

  // per-thread shared-memory slots, element k of thread t at [k * NT + t]
  // (the column accumulators and scales are kept as pairs: one 16-byte / 8-byte access per pair)
  constexpr int kPairs = kCols / 2;
  longlong2 *colacc_s = reinterpret_cast<longlong2 *>(tsm_raw) + tid;
  float2 *colscale_s = reinterpret_cast<float2 *>(tsm_raw + (size_t)kCols * 8 * NT) + tid;
  double *own = reinterpret_cast<double *>(tsm_raw + (size_t)kCols * 12 * NT) + tid;
  float *rowbuf = reinterpret_cast<float *>(tsm_raw + (size_t)(kCols * 12 + Policy::kSmemDoubles * 8) * NT) +
                  warp * (kBufRows * 32);  // this warp's [kBufRows][32] transposition buffer

  Policy pol;
  {
    ColumnCoefs<kCols, kAdj> cf;
    cf.compute(g.range, j0);
    pol.template load_columns<NT>(g, j0, sc, cf, own);
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      const int s0 = col_fixed_shift(cf.a[2 * h], cf.c0[2 * h], sc, g.n, g.prior_log2max, g.fixed_bits);
      const int s1 = col_fixed_shift(cf.a[2 * h + 1], cf.c0[2 * h + 1], sc, g.n, g.prior_log2max, g.fixed_bits);
      // 2^s, s in [-120, 120]
      colscale_s[h * NT] = make_float2(__int_as_float((s0 + 127) << 23), __int_as_float((s1 + 127) << 23));
      colacc_s[h * NT] = make_longlong2(0, 0);
    }
  }
  const bool col_any = j0 < g.N;  // columns only grow with cidx
  const bool has_prior = g.prior_mu || g.prior_sigma || g.prior_full;

  float mu[kRows];
  int row = 0;
  bool rows_any = false;

  // The range values of a claim's rows (at most 32 / kRows groups: one row per lane) are evaluated
  // once per claim, lane l taking row first_row + l (rows past the end repeat the last value; they
  // are masked in the epilogue); a group then takes its kRows by shuffle: the reference's float32
  // divide is paid once per claim, not once per group.
  float mu_item = 0.0f;
  int item_row0 = 0;
  auto begin_item = [&](int grp0) {
    item_row0 = row_base + grp0 * kRows;
    mu_item = range_value(min(item_row0 + lane, g.N - 1), g.N, g.range.mu0, g.range.mu1);
  };
  auto begin_group = [&](int grp) {
    row = row_base + grp * kRows;
    rows_any = row < g.row_end && row + kRows > g.row_begin;
    const int src = row - item_row0;
#pragma unroll
    for (int r = 0; r < kRows; ++r) mu[r] = __shfl_sync(0xffffffffu, mu_item, src + r);
    pol.begin_rows();
  };

  // Epilogue of one row group: store the kRows x kCols patch, add it into the row / column sums.
  // Within the patch the sums are float32 in a fixed pairwise order; the column sums go to the
  // fixed-point accumulators, the row sums of the warp are reduced in float64 in a fixed order.
  auto end_group = [&]() {
    if (!rows_any) return;  // warp-uniform
    float v[kRows][kCols];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) v[r][cidx] = pol.template value<NT>(g, r, cidx, own);
    const bool interior =
        row >= g.row_begin && row + kRows <= g.row_end && col_at<kAdj>(j0, kCols - 1) < g.N;
    if (!interior) {  // shard / grid edges: cells outside contribute exactly zero
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const bool rv = row + r >= g.row_begin && row + r < g.row_end;
#pragma unroll
        for (int cidx = 0; cidx < kCols; ++cidx)
          if (!(rv && col_at<kAdj>(j0, cidx) < g.N)) v[r][cidx] = 0.0f;
      }
    }
    if (has_prior) {  // likelihood x prior before normalisation (cuda_functions.h:145-148)
      float pm[kRows], ps[kCols];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const bool rv = row + r >= g.row_begin && row + r < g.row_end;
        pm[r] = (g.prior_mu && rv) ? g.prior_mu[row + r] : 1.0f;
      }
#pragma unroll
      for (int cidx = 0; cidx < kCols; ++cidx) {
        const int j = col_at<kAdj>(j0, cidx);
        ps[cidx] = (g.prior_sigma && j < g.N) ? g.prior_sigma[j] : 1.0f;
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int cidx = 0; cidx < kCols; ++cidx) v[r][cidx] *= pm[r] * ps[cidx];
      if (g.prior_full) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (!(row + r >= g.row_begin && row + r < g.row_end)) continue;
          const float *pf = g.prior_full + (size_t)(row + r - g.row_begin) * (size_t)g.N;
#pragma unroll
          for (int cidx = 0; cidx < kCols; ++cidx) {
            const int j = col_at<kAdj>(j0, cidx);
            if (j < g.N) v[r][cidx] *= pf[j];
          }
        }
      }
    }
    float *rowp = g.post + (ptrdiff_t)(row - g.row_begin) * (ptrdiff_t)g.N + j0;
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    bool stored = false;
    if constexpr (kAdj) {
      if (g.vec8 && warp_interior) {
        // adjacent layout: one 32-byte store per lane and row, a 1 KB run per instruction
        if (g.keep_l2) {
#pragma unroll
          for (int r = 0; r < kRows; ++r, rowp += g.N) st_keep_f8(rowp, v[r]);
        } else {
#pragma unroll
          for (int r = 0; r < kRows; ++r, rowp += g.N) st_stream_f8(rowp, v[r]);
        }
        stored = true;
      }
    }
    if (stored) {
    } else if (g.vec4 && warp_interior) {
      // the whole warp is inside the shard and the grid (all but the edge groups): straight-line
      // stores, one 512-byte run per instruction
      if (g.keep_l2) {
#pragma unroll
        for (int r = 0; r < kRows; ++r, rowp += g.N)
#pragma unroll
          for (int grp4 = 0; grp4 < kGroups; ++grp4)
            st_keep_f4(rowp + grp4 * kGroupStride, v[r][4 * grp4], v[r][4 * grp4 + 1],
                       v[r][4 * grp4 + 2], v[r][4 * grp4 + 3]);
      } else {
#pragma unroll
        for (int r = 0; r < kRows; ++r, rowp += g.N)
#pragma unroll
          for (int grp4 = 0; grp4 < kGroups; ++grp4)
            st_stream_f4(rowp + grp4 * kGroupStride, v[r][4 * grp4], v[r][4 * grp4 + 1],
                         v[r][4 * grp4 + 2], v[r][4 * grp4 + 3]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < kRows; ++r, rowp += g.N) {
        const bool rv = interior || (row + r >= g.row_begin && row + r < g.row_end);
#pragma unroll
        for (int grp4 = 0; grp4 < kGroups; ++grp4) {
          const int jg = j0 + grp4 * kGroupStride;
          if (g.vec4) {
            if (rv && jg < g.N)
              st_keep_f4(rowp + grp4 * kGroupStride, v[r][4 * grp4], v[r][4 * grp4 + 1],
                         v[r][4 * grp4 + 2], v[r][4 * grp4 + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < kRN; ++e)
              if (rv && jg + e < g.N) st_keep_f1(rowp + grp4 * kGroupStride + e, v[r][4 * grp4 + e]);
          }
        }
      }
    }
    // column sums over the patch rows (float32, pairwise), scaled by 2^s_j and added as integers
    auto colsum = [&](int cidx) {
      float t = (v[0][cidx] + v[1][cidx]) + (v[2][cidx] + v[3][cidx]);
      if (kRows >= 6) t += v[4][cidx] + v[5][cidx];
      if (kRows == 7) t += v[6][cidx];
      return t;
    };
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      const float c0 = colsum(2 * h), c1 = colsum(2 * h + 1);
      const float2 sc2 = colscale_s[h * NT];
      longlong2 acc = colacc_s[h * NT];
      acc.x += __float2ll_rn(c0 * sc2.x);
      acc.y += __float2ll_rn(c1 * sc2.y);
      colacc_s[h * NT] = acc;
    }
    // row sums over the patch columns (float32, pairwise), then across the warp in float64 in a
    // fixed order: every lane parks its kRows sums in the warp's shared [row][lane] buffer; lane l
    // takes back 4 adjacent lanes' sums of row l / 8 (one LDS.128), adds them, and three shuffles
    // finish rows 0..3 inside groups of 8 lanes; with 6 rows, lane l also takes 2 adjacent lanes'
    // sums of row 4 + l / 16 (one LDS.64) and four shuffles finish rows 4, 5 inside groups of 16
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float t = (v[r][0] + v[r][1]) + (v[r][2] + v[r][3]);
      if (kCols == 8) t += (v[r][4] + v[r][5]) + (v[r][6] + v[r][7]);
      rowbuf[r * 32 + lane] = t;
    }
    if (kRows == 7) rowbuf[7 * 32 + lane] = 0.0f;  // (rows 4..7 go through the 4-row path again)
    __syncwarp();
    const float4 q = *reinterpret_cast<const float4 *>(rowbuf + lane * 4);  // row lane/8, lanes 4*(lane%8)..+3
    float2 w = make_float2(0.0f, 0.0f);
    float4 q7 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (kRows == 6) w = *reinterpret_cast<const float2 *>(rowbuf + 128 + lane * 2);  // row 4 + lane/16
    if (kRows == 7) q7 = *reinterpret_cast<const float4 *>(rowbuf + 128 + lane * 4);  // row 4 + lane/8
    __syncwarp();
    double tot = ((double)q.x + (double)q.y) + ((double)q.z + (double)q.w);
    // (adjacent layout: the warp's 256 columns are two row-partial slots of 128 columns, lanes
    // 0..15 and 16..31: the tree stops one level early and publishes both halves)
    if (!kAdj) tot += __shfl_xor_sync(0xffffffffu, tot, 4);
    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    const int rq = row + (lane >> 3);  // lanes 0, 8, 16, 24 publish rows 0..3 of the patch
    if (!kAdj) {
      if ((lane & 7) == 0 && rq >= g.row_begin && rq < g.row_end)
        g.rowpart[(size_t)cw * rows_loc + (size_t)(rq - g.row_begin)] = tot;
    } else {
      const int slot = 2 * cw + ((lane >> 2) & 1);  // (lanes 0, 4 of every 8)
      if ((lane & 3) == 0 && rq >= g.row_begin && rq < g.row_end && slot < g.ncw)
        g.rowpart[(size_t)slot * rows_loc + (size_t)(rq - g.row_begin)] = tot;
    }
    if constexpr (kRows == 7) {  // rows 4..6 exactly as rows 0..2 (adjacent layout: two slots)
      double tot7 = ((double)q7.x + (double)q7.y) + ((double)q7.z + (double)q7.w);
      tot7 += __shfl_xor_sync(0xffffffffu, tot7, 2);
      tot7 += __shfl_xor_sync(0xffffffffu, tot7, 1);
      const int rq7 = row + 4 + (lane >> 3);
      const int slot = 2 * cw + ((lane >> 2) & 1);
      if ((lane & 3) == 0 && (lane >> 3) < 3 && rq7 >= g.row_begin && rq7 < g.row_end && slot < g.ncw)
        g.rowpart[(size_t)slot * rows_loc + (size_t)(rq7 - g.row_begin)] = tot7;
    }
    if (kRows == 6) {
      double tot2 = (double)w.x + (double)w.y;
      if (!kAdj) tot2 += __shfl_xor_sync(0xffffffffu, tot2, 8);
      tot2 += __shfl_xor_sync(0xffffffffu, tot2, 4);
      tot2 += __shfl_xor_sync(0xffffffffu, tot2, 2);
      tot2 += __shfl_xor_sync(0xffffffffu, tot2, 1);
      const int rq2 = row + 4 + (lane >> 4);  // lanes 0, 16 publish rows 4, 5
      if (!kAdj) {
        if ((lane & 15) == 0 && rq2 >= g.row_begin && rq2 < g.row_end)
          g.rowpart[(size_t)cw * rows_loc + (size_t)(rq2 - g.row_begin)] = tot2;
      } else {
        const int slot = 2 * cw + ((lane >> 3) & 1);  // (lanes 0, 8 of every 16)
        if ((lane & 7) == 0 && rq2 >= g.row_begin && rq2 < g.row_end && slot < g.ncw)
          g.rowpart[(size_t)slot * rows_loc + (size_t)(rq2 - g.row_begin)] = tot2;
      }
    }
  };

  if (!STREAMED) {
    const int strip_first = cw * g.groups, strip_last = strip_first + g.groups;
    int *pool_next = const_cast<int *>(claim) + 4;
    while (true) {
      const int e = min(u1, strip_last);
      const bool whole = u1 <= strip_last;  // the claim ends in this strip: take the next one now and
      int nxt = 0;                          // look at it after this one is done
      if (whole && lane == 0) nxt = atomicAdd(pool_next, g.chunk);
      begin_item(u0 - strip_first);
      for (int grp = u0 - strip_first; grp < e - strip_first; ++grp) {
        begin_group(grp);
        // (no skip for groups or columns outside the grid: they are masked in the epilogue, and a
        // single control path keeps the patch initialisation from being emitted per branch)
        Policy::consume(pol, obs_s, g.n, mu);
        end_group();
      }
      if (!whole) {  // the rest of the claim lies in the next strip
        u0 = strip_last;
        break;
      }
      u0 = __shfl_sync(0xffffffffu, nxt, 0);
      u1 = min(u0 + g.chunk, pool_end);
      if (u0 >= pool_end || u0 >= strip_last) break;
    }
  } else {
    int item = u0;
    const int nchunks = (g.n_pad + kObsChunk - 1) / kObsChunk;
    while (item < g.items_per_strip) {
      const int g0 = item * g.chunk * kWarpsCta, g1 = min(g.groups, g0 + g.chunk * kWarpsCta);
      claim_par ^= 1;
      int next_item = 0;
      if (tid == 0) next_item = atomicAdd(g.strip_next + cw, 1);
      for (int base = g0; base < g1; base += kWarpsCta) {
        // 32 KB chunks through two stages, the next chunk in flight while this one is consumed
        // (stage = parity of the running load count)
        if (tid == 0) issue_chunk(g, bars, obs_s, 0, streams & 1);
        begin_item(base + warp);
        begin_group(base + warp);
        if (base + warp >= g1) rows_any = false;  // (belongs to the next item)
        for (int ch = 0; ch < nchunks; ++ch, ++streams) {
          if (tid == 0 && ch + 1 < nchunks) issue_chunk(g, bars, obs_s, ch + 1, (streams + 1) & 1);
          mbar_wait(&bars[streams & 1], (streams >> 1) & 1);
          const int count = min(kObsChunk, g.n - ch * kObsChunk);
          if (rows_any && col_any)
            Policy::consume(pol, obs_s + (size_t)(streams & 1) * kObsChunk, count, mu);
          __syncthreads();  // every warp is done with this stage before it is refilled
        }
        end_group();
      }
      if (tid == 0) claim[claim_par] = next_item;
      __syncthreads();
      item = claim[claim_par];
    }
  }

  // leave the strip: the exact partial column sums join the shard's, one integer atomic each
  if (col_any) {
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      const longlong2 part = colacc_s[h * NT];
      const int ja = col_at<kAdj>(j0, 2 * h), jb = col_at<kAdj>(j0, 2 * h + 1);
      if (ja < g.N && part.x != 0) atomicAdd(g.colacc + ja, (unsigned long long)part.x);
      if (jb < g.N && part.y != 0) atomicAdd(g.colacc + jb, (unsigned long long)part.y);
    }
  }
}

// |a_j| is monotone along the sigma axis, so a strip's largest exponent step sits at one of its
// ends.  The exponent step scales with |a_j| = 0.72/sigma_j^2, so on a grid with a wide sigma range
// only the strips at small sigma need the costlier arithmetic.  A function of the strip (columns)
// and of the global statistics only: the same for every row range, shard and rank.
__device__ __forceinline__ double strip_umax(const RangeArgs &r, int cw, int strip_cols, double ustep) {
  const int jfirst = cw * strip_cols;
  const int jlast = min(jfirst + strip_cols, r.N) - 1;
  const double s0 = (double)range_value(jfirst, r.N, r.sg0, r.sg1);
  const double s1 = (double)range_value(jlast, r.N, r.sg0, r.sg1);
  return fmax(fabs(coef_a(s0)), fabs(coef_a(s1))) * ustep;
}
// Relative sigma step of the strip, |dsigma| / min |sigma| (the slope policies' eligibility)
__device__ __forceinline__ double strip_sigma_step(const RangeArgs &r, int cw, int strip_cols) {
  const int jfirst = cw * strip_cols;
  const int jlast = min(jfirst + strip_cols, r.N) - 1;
  const double s0 = (double)range_value(jfirst, r.N, r.sg0, r.sg1);
  const double s1 = (double)range_value(jlast, r.N, r.sg0, r.sg1);
  const double ds = r.N > 1 ? fabs((double)r.sg1 - (double)r.sg0) / (double)(r.N - 1) : 0.0;
  const double smin = fmin(fabs(s0), fabs(s1));
  return (smin > 0.0 && s0 * s1 > 0.0) ? ds / smin : 1e300;  // (a strip across sigma = 0: never)
}

// ---------------------------------------------------------------------------------------
// likelihood_cta: everything CTA `cta` of `nctas` does in the likelihood phase.
//   SELECT: the policy is chosen per column strip on the device, the first of (Finest, Fine, Mid,
//   Coarse) whose eligibility bound the strip meets -- by default ProductSlope, ProductNeighbour
//   of degree 2 and 3 (exponent step up to 0.011 / 0.05), ProductHybrid (any grid).  The choice
//   depends on (ranges, N, observations, strip) only -- never on the rows, the shard or the warp --
//   so every rank of a sharded run takes the same path for the same cells and results stay
//   bit-reproducible.  !SELECT: Finest everywhere.
//   PRELOADED: the caller (small_grid_kernel) has staged the observations and written the
//   ScaleInfo slot of shared memory, and synchronised the block.
//   NT: threads per CTA.
// ---------------------------------------------------------------------------------------
template <class Finest, class Fine, class Mid, class Coarse, bool SELECT, bool PRELOADED,
          bool STREAMED = false, int NT = kThreads, bool WIDE = false>
__device__ __forceinline__ void likelihood_cta(const LikeArgs &g, unsigned char *smem_raw,
                                               double *scratch, unsigned cta, unsigned nctas) {
  static_assert(Fine::kCols == Coarse::kCols && Fine::kCols == Mid::kCols && Fine::kCols == Finest::kCols,
                "same tiling");
  static_assert(Fine::kRows == Coarse::kRows && Fine::kRows == Mid::kRows && Fine::kRows == Finest::kRows,
                "same row groups");
  static_assert(!(PRELOADED && STREAMED), "the small-grid kernel stages every observation at once");
  static_assert(!WIDE || (SELECT && !PRELOADED && !STREAMED), "the wide policy belongs to the default product kernel");
  constexpr int kCols = Fine::kCols;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
  ScaleInfo *sc_s = reinterpret_cast<ScaleInfo *>(smem_raw + 16);
  volatile int *claim = reinterpret_cast<volatile int *>(smem_raw + 128);
  float *obs_s = reinterpret_cast<float *>(smem_raw + kSmemHeader);
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned char *after_stage = smem_raw + kSmemHeader + stage_bytes(g.n_pad);

  // WIDE: the whole grid runs ProductSlopeTall or ProductSlopeWide (strips of 256 columns) when
  // the grid-level bounds allow it -- a function of (ranges, N, observations) only, like the
  // per-strip choice below.  0: no, 1: wide (6 rows x 8 columns), 2: tall (7 x 8)
  int wide = 0;
  if (!PRELOADED) {
    if (!STREAMED) {
      // fused set-up: this CTA's own copy of the observations and of the statistics
      const ScaleInfo sc = setup_scale(g.range, g.obs, obs_s, g.n, g.n_pad, g.centred,
                                       g.reorder ? reinterpret_cast<float *>(after_stage) : nullptr,
                                       scratch);
      if (WIDE) {
        if (g.wide > 0) wide = g.wide;  // forced
        else if (g.wide == 0 && ProductSlopeTall::eligible(sc.umax, sc.seps)) wide = 2;
        else if (g.wide != -1 && ProductSlopeWide::eligible(sc.umax, sc.seps)) wide = 1;
      }
      if (tid == 0) {
        *sc_s = sc;
        if (cta == 0) {
          *g.scale_out = sc;
          g.scale_out->wide = (double)wide;
        }
      }
    } else {
      if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init_fence();
        *sc_s = *g.scale_in;
      }
    }
  }
  // this CTA's pool: an equal contiguous share of the groups, strip-major
  const int ncw_run = wide ? (g.N + 32 * ProductSlopeWide::kCols - 1) / (32 * ProductSlopeWide::kCols) : g.ncw;
  // (the tall policy's groups are 7 rows, anchored at global multiples of 7)
  const int groups_run =
      wide == 2 ? (g.row_end - (g.row_begin - g.row_begin % ProductSlopeTall::kRows) + ProductSlopeTall::kRows - 1) /
                      ProductSlopeTall::kRows
                : g.groups;
  const long long units = (long long)ncw_run * groups_run;
  const int pool_begin = (int)(units * cta / nctas), pool_end = (int)(units * (cta + 1) / nctas);
  if (tid == 0) claim[4] = pool_begin;
  __syncthreads();
#ifdef CGE_DEBUG_TIMES
  if (g.debug && tid == 0) {  // set-up done
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.debug[4 * cta + 3] = t;
  }
#endif

  const int strip_cols = 32 * kCols;
  int claim_par = 0, streams = 0;
  auto dispatch = [&](int cw, int &u0, int &u1) {
    ScaleInfo sc = *sc_s;
    sc.umax = strip_umax(g.range, cw, strip_cols, sc.ustep);  // the strip's own bound
    const double eps = SELECT ? strip_sigma_step(g.range, cw, strip_cols) : 0.0;
    if (!SELECT || Finest::eligible(sc.umax, eps))
      run_strip<Finest, STREAMED, NT>(g, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim, claim_par, streams);
    else if (Fine::eligible(sc.umax, eps))
      run_strip<Fine, STREAMED, NT>(g, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim, claim_par, streams);
    else if (Mid::eligible(sc.umax, eps))
      run_strip<Mid, STREAMED, NT>(g, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim, claim_par, streams);
    else
      run_strip<Coarse, STREAMED, NT>(g, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim, claim_par, streams);
  };

  if (!STREAMED) {
    // a claim covers at most 32 rows (run_strip: one range value per lane)
    const int chunk_run = wide == 2 ? min(g.chunk, 32 / ProductSlopeTall::kRows) : g.chunk;
    int u0 = 0;
    if (lane == 0) u0 = atomicAdd(const_cast<int *>(claim) + 4, chunk_run);
    u0 = __shfl_sync(0xffffffffu, u0, 0);
    int u1 = min(u0 + chunk_run, pool_end);
    if constexpr (WIDE) {
      if (wide == 2) {
        LikeArgs gt = g;
        gt.groups = groups_run;
        gt.chunk = chunk_run;
        while (u0 < pool_end) {
          const int cw = u0 / gt.groups;
          ScaleInfo sc = *sc_s;
          sc.umax = strip_umax(g.range, cw, 32 * ProductSlopeTall::kCols, sc.ustep);
          run_strip<ProductSlopeTall, false, NT>(gt, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim,
                                                 claim_par, streams);
        }
        return;
      }
      if (wide == 1) {
        while (u0 < pool_end) {
          const int cw = u0 / g.groups;
          ScaleInfo sc = *sc_s;
          sc.umax = strip_umax(g.range, cw, 32 * ProductSlopeWide::kCols, sc.ustep);
          run_strip<ProductSlopeWide, false, NT>(g, sc, cw, u0, u1, pool_end, bars, obs_s, after_stage, claim,
                                                 claim_par, streams);
        }
        return;
      }
    }
    while (u0 < pool_end) dispatch(u0 / g.groups, u0, u1);
    return;
  }

  // streamed: the CTAs are spread evenly over the strips; a CTA takes its home strip's items with
  // everyone else who lives there, then walks on through the other strips
  int cw = (int)(((long long)cta * g.ncw) / nctas);
  for (int visited = 0; visited < g.ncw; ++visited) {
    if (tid == 0) claim[2] = atomicAdd(g.strip_next + cw, 1);
    __syncthreads();
    int item = claim[2], unused = 0;
    __syncthreads();  // (the slot is rewritten at the next strip)
    if (item < g.items_per_strip) dispatch(cw, item, unused);
    cw = cw + 1 == g.ncw ? 0 : cw + 1;
  }
}

template <class Finest, class Fine, class Mid, class Coarse, bool SELECT, int MINB = 1,
          bool STREAMED = false, int WARPS = kWarps, bool WIDE = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    grid_likelihood_kernel(const LikeArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double scratch[32];
  pdl_trigger();  // the fold kernel may be made resident as this grid's CTAs retire
  pdl_wait();     // the previous step's finish kernel still reads the posterior this one overwrites,
                  // and the caller's kernels may have produced the observations
  if (blockIdx.x == 0 && threadIdx.x < 2 && g.tickets) g.tickets[threadIdx.x] = 0u;
#ifdef CGE_DEBUG_TIMES
  unsigned long long t0 = 0;
  if (g.debug && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
  likelihood_cta<Finest, Fine, Mid, Coarse, SELECT, false, STREAMED, WARPS * 32, WIDE>(g, smem_raw, scratch,
                                                                                     blockIdx.x, gridDim.x);
#ifdef CGE_DEBUG_TIMES
  if (g.debug && (threadIdx.x & 31) == 0) {  // per warp: when it ran out of work
    unsigned long long tw;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw));
    g.debug[4 * gridDim.x + 32 * blockIdx.x + (threadIdx.x >> 5)] = tw;  // (32 slots per CTA)
  }
  if (g.debug && threadIdx.x == 0) {
    unsigned long long t1;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g.debug[4 * blockIdx.x + 0] = t0;
    g.debug[4 * blockIdx.x + 1] = t1;
    g.debug[4 * blockIdx.x + 2] = smid;
  }
#endif
}

// ---------------------------------------------------------------------------------------
// small_grid_kernel: the whole step in ONE launch for grids of a few hundred points per axis,
// where three launches are all latency (the README workload, N = 101, is 510 050 pdf-evals:
// ~2 us of arithmetic).  Every CTA
//   1. stages the observations and derives the scale statistics itself (identical in every CTA);
//   2. runs the same fused likelihood body on its units (unnormalised cells + partial sums);
//   3. CLUSTER: the grid is ONE thread-block cluster (one column band, <= 8 CTAs: every grid up
//      to 256 x 256), so its CTAs are co-scheduled and a hardware cluster barrier follows; after
//      it every CTA folds the row partials itself (same order everywhere, so the same Z bit for
//      bit) and scales ITS OWN rows in parallel; CTA 0 also writes the marginals and expectations.
//      !CLUSTER: the last CTA to take a ticket does K6-K8 and scales the whole (L2-resident)
//      posterior: no co-residency requirement.
// Observations may be read from, and the results block also written to, mapped pinned host
// memory, which removes both small DMA copies (~6 us and ~10 us) from the call's latency.
// Sums are in a fixed order (index order within a thread, then the fixed block tree).
// ---------------------------------------------------------------------------------------
struct SmallArgs {
  LikeArgs like;  // row_begin = 0, row_end = N; one band
  double *partials, *rowsum, *results;
  double *results_host;  // optional mapped pinned copy of `results` (saves the D2H copy)
  unsigned int *ticket;
};

// in place: post[0..count) *= finv, reading through L2 (other CTAs wrote some of it), several
// independent loads in flight per thread; `post` is 16-byte aligned when vec4
__device__ __forceinline__ void scale_cells_cta(float *post, size_t count, float finv, bool vec4) {
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr int kDepth = 8;
  if (vec4) {
    float4 *p4 = reinterpret_cast<float4 *>(post);
    const size_t nvec = count >> 2;  // vec4 implies N % 4 == 0, so count % 4 == 0
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

template <class Finest, class Fine, class Mid, class Coarse, bool SELECT, bool CLUSTER>
__global__ void __launch_bounds__(kThreads)
    small_grid_kernel(const SmallArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double scratch[32];
  __shared__ bool is_last;
  const LikeArgs &g = a.like;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = g.N;
  const unsigned cta = blockIdx.x, nctas = gridDim.x;
  float *obs_s = reinterpret_cast<float *>(smem_raw + kSmemHeader);
  float *sort_s = reinterpret_cast<float *>(smem_raw + kSmemHeader + stage_bytes(g.n_pad));

  {
    const ScaleInfo sc = setup_scale(g.range, g.obs, obs_s, g.n, g.n_pad, g.centred,
                                     g.reorder ? sort_s : nullptr, scratch);
    if (tid == 0) {
      *reinterpret_cast<ScaleInfo *>(smem_raw + 16) = sc;
      if (cta == 0) *g.scale_out = sc;
    }
    __syncthreads();
  }
  likelihood_cta<Finest, Fine, Mid, Coarse, SELECT, true>(g, smem_raw, scratch, cta, nctas);
  __syncthreads();  // the warps of this CTA worked independently: all of them are done

  bool writer;  // the CTA that writes partials / results: the last one, or cluster rank 0
  if (CLUSTER) {
    cluster_barrier();
    writer = cta == 0;
  } else {
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.ticket, 1u) == nctas - 1;
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
    for (int w = 0; w < g.ncw; ++w) s += __ldcg(g.rowpart + (size_t)w * N + i);  // [slot][row], rows_local = N
    if (writer) a.rowsum[i] = s;
    z += s;
    smu += s * (double)range_value(i, N, g.range.mu0, g.range.mu1);
  }
  if (writer) {  // column sums: the exact fixed-point totals scaled back (col_fixed_shift), and the
                 // accumulators and claim counters left at zero for the next call
    const ScaleInfo sc = *reinterpret_cast<const ScaleInfo *>(smem_raw + 16);
    for (int j = tid; j < N; j += nt) {
      double ca, cc;
      column_coef(g.range, j, 4, ca, cc);  // (the single-launch kernel has the 4-neighbour layouts only)
      const int s = col_fixed_shift(ca, cc, sc, g.n, g.prior_log2max, g.fixed_bits);
      a.partials[2 + j] = scalbn((double)(long long)__ldcg(g.colacc + j), -s);
      g.colacc[j] = 0ull;
    }
    for (int b = tid; b < g.ncw; b += nt) g.strip_next[b] = 0;
    __syncthreads();
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
      es += m * (double)range_value(j, N, g.range.sg0, g.range.sg1);
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
  if (CLUSTER) {  // an even share of the rows per CTA (whole rows, so float4 alignment holds)
    const int r0 = (int)((long long)N * cta / nctas), r1 = (int)((long long)N * (cta + 1) / nctas);
    if (r0 < r1) scale_cells_cta(g.post + (size_t)r0 * N, (size_t)(r1 - r0) * N, finv, g.vec4 != 0);
  } else {
    scale_cells_cta(g.post, (size_t)N * N, finv, g.vec4 != 0);
  }
}

}  // namespace cge
