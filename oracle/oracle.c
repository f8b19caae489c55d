/*
 * oracle.c -- CPU restatement of the reference's grid-estimation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path (cuda_grid_estimation_b200/) never
 * links or imports it.
 *
 * Parity pinning: the stage functions below are checked against every
 * known-answer vector the reference's own tests hold for this path
 * (cpp/grid-algorithm/src/tests.cu:62,93-95,126-128,183-184,206-208,239-240 and
 * inc/utils.h:16-30) in tests/test_oracle_kat.py, against the reference's
 * Python statement of the algorithm (notebooks/quick-prototyping.ipynb cell 4,
 * imported in the build container by tests/golden/make_golden.py) and against
 * outputs of the reference CUDA pipeline itself run on a B200
 * (oracle/ref_harness.cu -> tests/golden/ref_cuda_*.npz).
 *
 * Every function cites the reference lines it follows; paths are relative to
 * /root/reference/cpp/grid-algorithm/.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: the range formula and the f32 pdf must round
 * after every operation, like the reference's separate f32 instructions.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* numerics selectors for oracle_grid_posterior */
enum {
  ORACLE_NUM_REF = 0, /* f32 pdf (reference formula), f64 tree product       */
  ORACLE_NUM_F64 = 1, /* everything in f64 from the f32-valued inputs        */
  ORACLE_NUM_LOG = 2  /* f64 sum of log-pdf, log-sum-exp normalisation       */
};

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------------------
 * a1  linspace -- inc/kernels.h:61-64 (launch: inc/cuda_functions.h:47-62).
 * The evaluation order pins the bits: int index widened to f32, times the f32
 * span, divided by f32(size-1), plus start, each rounded to f32.
 * ------------------------------------------------------------------------- */
static inline float range_value(int idx, int size, float start, float end) {
  float span = end - start;
  float num = (float)idx * span;
  float q = num / (float)(size - 1);
  return start + q;
}

ORACLE_API void oracle_linspace(float *out, int size, float start, float end) {
  for (int i = 0; i < size; ++i) out[i] = range_value(i, size, start, end);
}

/* ---------------------------------------------------------------------------
 * a2  3-way meshgrid -- inc/kernels.h:88-98; layout pinned by tests.cu:93-95:
 * flat = i*N*n + j*n + k with mu slowest, sigma middle, observation fastest.
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_meshgrid3(const float *vx, const float *vy,
                                 const float *vz, float *gx, float *gy,
                                 float *gz, int nxy, int nz) {
  for (int i = 0; i < nxy; ++i)
    for (int j = 0; j < nxy; ++j)
      for (int k = 0; k < nz; ++k) {
        size_t f = ((size_t)i * nxy + j) * nz + k;
        gx[f] = vx[i];
        gy[f] = vy[j];
        gz[f] = vz[k];
      }
}

/* ---------------------------------------------------------------------------
 * a3  Normal pdf with the reference's mixed precision -- inc/kernels.h:102-112:
 * z, z^2, -0.5*z^2 and exp are f32 (pow/exp resolve to the f32 overloads in
 * device code); sqrt(2.0f*M_PI) and the divide are f64 (M_PI is a double
 * literal); the result is truncated to f32 on return.
 * ------------------------------------------------------------------------- */
static inline float pdf_ref_f32(float x, float mu, float sigma) {
  float z = (x - mu) / sigma;
  float zz = powf(z, 2.0f);
  float e = expf(-0.5f * zz);
  double denom = (double)sigma * sqrt(2.0f * M_PI);
  return (float)((double)e / denom);
}

ORACLE_API float oracle_normal_pdf(float x, float mu, float sigma) {
  return pdf_ref_f32(x, mu, sigma);
}

/* a4  elementwise densities -- inc/kernels.h:134-140. */
ORACLE_API void oracle_densities(float *dens, const float *gx, const float *gy,
                                 const float *gz, size_t count) {
#pragma omp parallel for schedule(static)
  for (size_t f = 0; f < count; ++f) dens[f] = pdf_ref_f32(gz[f], gx[f], gy[f]);
}

/* ---------------------------------------------------------------------------
 * a8  per-row product, pairwise tree -- inc/kernels.h:190-204.
 * The reference starts at stride blockDim.x/2 = 128 and is therefore only
 * valid for cols <= 256; `tree_first_stride` keeps 128 in that range and
 * grows to the next power of two above it so the oracle stays defined.
 * `work` is one row of f64 copies (the reference widens f32 -> f64 first,
 * inc/wrappers.h:117 + inc/utils.h:46-50).
 * ------------------------------------------------------------------------- */
static inline int tree_first_stride(int len) {
  int s = 128;
  while (2 * s < len) s <<= 1;
  return s;
}

static inline double tree_product(double *work, int len) {
  for (int stride = tree_first_stride(len); stride > 0; stride >>= 1)
    for (int t = 0; t < stride && t + stride < len; ++t) work[t] *= work[t + stride];
  return work[0];
}

static inline double tree_sum_strided(double *work, int len) {
  for (int stride = tree_first_stride(len); stride > 0; stride >>= 1)
    for (int t = 0; t < stride && t + stride < len; ++t) work[t] += work[t + stride];
  return work[0];
}

ORACLE_API void oracle_row_products(double *likes, const float *dens,
                                    size_t rows, int cols) {
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * (size_t)cols);
#pragma omp for schedule(static)
    for (size_t r = 0; r < rows; ++r) {
      for (int k = 0; k < cols; ++k) work[k] = (double)dens[r * cols + k];
      likes[r] = tree_product(work, cols);
    }
    free(work);
  }
}

/* ---------------------------------------------------------------------------
 * a10  posterior under the flat prior -- inc/cuda_functions.h:155-158 and
 * inc/utils.h:82-84: Z = sum(likes); posterior = likes * (1/Z) (reciprocal
 * multiply, not a divide).  Thrust's summation order is unspecified; a plain
 * f64 sum differs from it by O(1e-16) relative.
 * ------------------------------------------------------------------------- */
ORACLE_API double oracle_posterior(const double *likes, double *post,
                                   size_t count) {
  double z = 0.0;
  for (size_t f = 0; f < count; ++f) z += likes[f];
  double inv = 1 / z;
  for (size_t f = 0; f < count; ++f) post[f] = likes[f] * inv;
  return z;
}

/* ---------------------------------------------------------------------------
 * a11  marginal along an axis, pairwise tree -- inc/kernels.h:237-258 with the
 * axis meaning of inc/wrappers.h:169-170: axis=1 folds columns (row sums ->
 * mu marginal), axis=0 folds rows (column sums -> sigma marginal).
 * `mat` is row-major rows x cols and is left untouched (the reference
 * destroys its copy).
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_marginalize(double *marginal, const double *mat, int rows,
                                   int cols, int axis) {
  int outer = axis == 0 ? rows : cols;  /* the folded dimension */
  int inner = axis == 0 ? cols : rows;  /* the surviving dimension */
  double *work = (double *)malloc(sizeof(double) * (size_t)outer);
  for (int s = 0; s < inner; ++s) {
    for (int t = 0; t < outer; ++t)
      work[t] = axis == 0 ? mat[(size_t)t * cols + s] : mat[(size_t)s * cols + t];
    marginal[s] = tree_sum_strided(work, outer);
  }
  free(work);
}

/* a12  expectation -- inc/cuda_functions.h:201-218: f32 range widened to f64,
 * elementwise product, f64 sum. */
ORACLE_API double oracle_expectation(const double *marginal, const float *vec,
                                     int size) {
  double e = 0.0;
  for (int i = 0; i < size; ++i) e += marginal[i] * (double)vec[i];
  return e;
}

/* ---------------------------------------------------------------------------
 * Blocked whole-path oracle: a1..a13 end to end without materialising the
 * N*N*n tensors (src/grid.cu:79-127 is the call sequence it follows).
 *
 *   likes_out   optional, (row_end-row_begin) x N unnormalised values:
 *               likelihoods for REF/F64, natural-log likelihoods for LOG
 *   post_out    optional, N x N posterior (only when the whole grid is asked)
 *   marg_mu/marg_sigma optional, N each; expect2 optional, {E[mu], E[sigma]}
 *
 * Returns 0, or 1 when the normaliser is zero / not finite.
 * ------------------------------------------------------------------------- */
static inline double cell_product_ref(const float *obs, int n, float mu,
                                      float sigma, double *work) {
  for (int k = 0; k < n; ++k) work[k] = (double)pdf_ref_f32(obs[k], mu, sigma);
  return tree_product(work, n);
}

static inline double cell_product_f64(const float *obs, int n, float mu,
                                      float sigma) {
  double m = mu, s = sigma, p = 1.0;
  double norm = 1.0 / (s * sqrt(2.0 * M_PI));
  for (int k = 0; k < n; ++k) {
    double z = ((double)obs[k] - m) / s;
    p *= exp(-0.5 * z * z) * norm;
  }
  return p;
}

static inline double cell_loglik_f64(const float *obs, int n, float mu,
                                     float sigma) {
  double m = mu, s = sigma, acc = 0.0;
  for (int k = 0; k < n; ++k) {
    double z = ((double)obs[k] - m) / s;
    acc += -0.5 * z * z;
  }
  return acc - (double)n * log(s * sqrt(2.0 * M_PI));
}

ORACLE_API int oracle_grid_likes(const float *obs, int n, int N, float mu0,
                                 float mu1, float sg0, float sg1, int numerics,
                                 int row_begin, int row_end, double *likes_out) {
  float *sg = (float *)malloc(sizeof(float) * (size_t)N);
  oracle_linspace(sg, N, sg0, sg1);
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
    for (int i = row_begin; i < row_end; ++i) {
      float mu = range_value(i, N, mu0, mu1);
      double *dst = likes_out + (size_t)(i - row_begin) * N;
      for (int j = 0; j < N; ++j) {
        if (numerics == ORACLE_NUM_REF)
          dst[j] = cell_product_ref(obs, n, mu, sg[j], work);
        else if (numerics == ORACLE_NUM_F64)
          dst[j] = cell_product_f64(obs, n, mu, sg[j]);
        else
          dst[j] = cell_loglik_f64(obs, n, mu, sg[j]);
      }
    }
    free(work);
  }
  free(sg);
  return 0;
}

ORACLE_API int oracle_grid_posterior(const float *obs, int n, int N, float mu0,
                                     float mu1, float sg0, float sg1,
                                     int numerics, double *post_out,
                                     double *marg_mu, double *marg_sigma,
                                     double *expect2) {
  size_t cells = (size_t)N * N;
  double *likes = post_out ? post_out : (double *)malloc(sizeof(double) * cells);
  if (!likes) return 2;
  oracle_grid_likes(obs, n, N, mu0, mu1, sg0, sg1, numerics, 0, N, likes);

  if (numerics == ORACLE_NUM_LOG) {
    double mx = -INFINITY;
    for (size_t f = 0; f < cells; ++f) mx = likes[f] > mx ? likes[f] : mx;
#pragma omp parallel for schedule(static)
    for (size_t f = 0; f < cells; ++f) likes[f] = exp(likes[f] - mx);
  }

  /* normalise: reciprocal multiply, as inc/utils.h:82-84 */
  double z = 0.0;
  for (size_t f = 0; f < cells; ++f) z += likes[f];
  int bad = !(z > 0.0) || !isfinite(z);
  double inv = 1 / z;
#pragma omp parallel for schedule(static)
  for (size_t f = 0; f < cells; ++f) likes[f] *= inv;

  float *mu = (float *)malloc(sizeof(float) * (size_t)N);
  float *sg = (float *)malloc(sizeof(float) * (size_t)N);
  oracle_linspace(mu, N, mu0, mu1);
  oracle_linspace(sg, N, sg0, sg1);
  double *mm = marg_mu ? marg_mu : (double *)malloc(sizeof(double) * (size_t)N);
  double *ms = marg_sigma ? marg_sigma : (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i) { /* row sums = mu marginal (wrappers.h:169) */
    double s = 0.0;
    for (int j = 0; j < N; ++j) s += likes[(size_t)i * N + j];
    mm[i] = s;
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < N; ++j) { /* column sums = sigma marginal (wrappers.h:170) */
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += likes[(size_t)i * N + j];
    ms[j] = s;
  }
  if (expect2) {
    expect2[0] = oracle_expectation(mm, mu, N);
    expect2[1] = oracle_expectation(ms, sg, N);
  }
  if (!marg_mu) free(mm);
  if (!marg_sigma) free(ms);
  free(mu);
  free(sg);
  if (!post_out) free(likes);
  return bad;
}

/* ---------------------------------------------------------------------------
 * Timed CPU baseline helper (bench.py cpu_baseline / --impl reference): the
 * REF-numerics likelihood + running row/column sums over a bounded row range,
 * all host threads, nothing stored per cell except the row of likelihoods.
 * Returns the number of pdf evaluations performed.
 * ------------------------------------------------------------------------- */
ORACLE_API double oracle_bench_rows(const float *obs, int n, int N, float mu0,
                                    float mu1, float sg0, float sg1,
                                    int numerics, int row_begin, int row_end,
                                    double *row_sums, double *col_sums) {
  float *sg = (float *)malloc(sizeof(float) * (size_t)N);
  oracle_linspace(sg, N, sg0, sg1);
  memset(col_sums, 0, sizeof(double) * (size_t)N);
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *cols = (double *)calloc((size_t)N, sizeof(double));
#pragma omp for schedule(dynamic, 1)
    for (int i = row_begin; i < row_end; ++i) {
      float mu = range_value(i, N, mu0, mu1);
      double rs = 0.0;
      for (int j = 0; j < N; ++j) {
        double v;
        if (numerics == ORACLE_NUM_REF)
          v = cell_product_ref(obs, n, mu, sg[j], work);
        else if (numerics == ORACLE_NUM_F64)
          v = cell_product_f64(obs, n, mu, sg[j]);
        else
          v = cell_loglik_f64(obs, n, mu, sg[j]);
        rs += v;
        cols[j] += v;
      }
      row_sums[i - row_begin] = rs;
    }
#pragma omp critical
    for (int j = 0; j < N; ++j) col_sums[j] += cols[j];
    free(work);
    free(cols);
  }
  free(sg);
  return (double)(row_end - row_begin) * (double)N * (double)n;
}
