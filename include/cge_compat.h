/*
 * cge_compat.h -- the reference's C++ host API, re-implemented on top of the libcge.so C ABI.
 *
 * nablabits/cuda-grid-estimation has no library: its host API is a set of free functions defined
 * in headers (cpp/grid-algorithm/inc/{utils,cuda_functions,wrappers}.h) that src/grid.cu and
 * src/tests.cu include.  This header provides the SAME names, argument lists and argument
 * meaning, so a caller written against the reference (including its main() and its tests)
 * compiles against this file instead of those three headers and links with -lcge.
 * Compile with nvcc (the reference's signatures use thrust::device_vector).
 *
 * Every function below cites the reference definition it replaces (paths relative to
 * cpp/grid-algorithm/).  Behavioural notes, where the reference's behaviour is a bug a drop-in
 * should not keep, are marked DEVIATION.
 *
 * The stage functions materialise what their signatures force them to (the N*N*n `densities`
 * buffer); the fused path that does not is gridPosterior() at the bottom / cge_grid_posterior().
 */
#ifndef CGE_COMPAT_H
#define CGE_COMPAT_H

#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <cmath>
#include <cstdio>
#include <iostream>

#include "cge.h"

namespace cge_compat {

/* One lazily created context on the current device, like the reference's implicit use of the
 * default device/stream.  Failure to create it (no sm_100 GPU) is fatal: there is no CPU path. */
inline cge_context *context() {
  // (function-local static: initialised once, thread-safe since C++11)
  static cge_context *ctx = [] {
    cge_context *c = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cge_create(dev, &c) != CGE_OK) {
      std::fprintf(stderr, "libcge: %s\n", cge_last_error());
      std::abort();
    }
    return c;
  }();
  return ctx;
}

/* the reference's library code ignores errors and prints on logic errors (wrappers.h:110-113);
 * the shim prints libcge's message and carries on, which is the closest equivalent */
inline bool ok(int rc) {
  if (rc != CGE_OK) std::printf("ERROR: %s\n", cge_last_error());
  return rc == CGE_OK;
}

}  // namespace cge_compat

/* ---- inc/utils.h ---------------------------------------------------------------------- */

/* inc/utils.h:34-51 -- flat array -> m x n row-pointer matrix with widening (host loop) */
template <class T, class V>
inline void reshapeArray(T *flatArray, V **output, int m, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) output[i][j] = flatArray[i * m + j];
}

/* inc/utils.h:53-63 -- row-pointer matrix in managed memory.  The reference makes rows+1
 * driver allocations; here it is one block for the data and one for the row pointers. */
inline double **allocateMatrix(int rows, int cols) {
  double **output = nullptr;
  double *block = nullptr;
  cudaMallocManaged(&output, (size_t)rows * sizeof(double *));
  cudaMallocManaged(&block, (size_t)rows * cols * sizeof(double));
  for (int i = 0; i < rows; i++) output[i] = block + (size_t)i * cols;
  return output;
}

/* inc/utils.h:65-71 */
inline void freeMatrix(double **matrix, int rows) {
  (void)rows;
  if (!matrix) return;
  cudaFree(matrix[0]);
  cudaFree(matrix);
}

/* inc/utils.h:9-32 checkArrays is a development assertion hard-wired to N=101, n=50 that reads
 * the materialised grids on the host; its four exact values are a known-answer test here
 * (tests/test_oracle_kat.py::test_kat_check_arrays_values), not runtime code. */

/* ---- inc/cuda_functions.h --------------------------------------------------------------- */

/* inc/cuda_functions.h:13-44 -- observation synthesis with the reference's generator (cuRAND
 * XORWOW, seed 42 as in inc/kernels.h:18, one subsequence per observation).  Not part of the
 * hot path; here so the reference's main() links without its own kernels.h. */
inline void generateNormalCuda(unsigned int n, float mu, float sigma, float *observations) {
  cge_compat::ok(cge_generate_normal(cge_compat::context(), observations, (int)n, mu, sigma, 42ULL, 0));
}

/* inc/cuda_functions.h:47-62 */
inline void linspaceCuda(float *array, int size, float start, float end) {
  cge_compat::ok(cge_linspace(cge_compat::context(), array, size, start, end));
}

/* inc/cuda_functions.h:65-92 */
inline void create3dGridCuda(float *vecX, float *vecY, float *vecZ, float *gridX, float *gridY,
                             float *gridZ, int vecXYSize, int vecZSize) {
  cge_compat::ok(cge_meshgrid3(cge_compat::context(), vecX, vecY, vecZ, gridX, gridY, gridZ,
                               vecXYSize, vecZSize));
}

/* inc/cuda_functions.h:95-115 */
inline void computeDensitiesCuda(float *densities, float *gridX, float *gridY, float *gridZ,
                                 int gridSize) {
  cge_compat::ok(cge_densities(cge_compat::context(), densities, gridX, gridY, gridZ,
                               (size_t)gridSize));
}

/* inc/cuda_functions.h:118-136.  DEVIATION: the reference multiplies in place and destroys
 * likesMatrix, and is only correct for cols <= 256; here the matrix is read-only and any
 * width works. */
inline void computeLikesCuda(double *likes, double **likesMatrix, int rows, int cols) {
  cge_compat::ok(cge_row_products_rows(cge_compat::context(), likes, likesMatrix, rows, cols));
}

/* inc/cuda_functions.h:139-159: posterior = likes * (1 / sum(likes)), flat prior */
inline void computePosteriorCuda(thrust::device_vector<double> &likesV,
                                 thrust::device_vector<double> &posteriorV) {
  cge_compat::ok(cge_normalize(cge_compat::context(), thrust::raw_pointer_cast(likesV.data()),
                               thrust::raw_pointer_cast(posteriorV.data()), likesV.size()));
}

/* inc/cuda_functions.h:161-182.  DEVIATION: the reference sums in place (destroying the
 * matrix) and is only correct for sides <= 256. */
inline void marginalizeCuda(double *marginal, double **posteriorMatrix, int rows, int cols,
                            int axis) {
  cge_compat::ok(cge_marginalize_rows(cge_compat::context(), marginal, posteriorMatrix, rows, cols,
                                      axis));
}

/* inc/cuda_functions.h:185-221 */
inline double computeExpectationsCuda(double *marginal, float *vector, int size) {
  double e = std::nan("");
  cge_compat::ok(cge_expectation(cge_compat::context(), marginal, vector, size, &e));
  return e;
}

/* ---- inc/wrappers.h ----------------------------------------------------------------------- */

/* inc/wrappers.h:19-92.  Same arguments (the int range parameters included: the reference
 * truncates its float ranges to int, wrappers.h:20-21).  Fills vecX, vecY with the ranges and
 * `output` with the N*N*n densities, but goes straight from the three vectors: the three
 * N*N*n meshgrid temporaries and the host-side checkArrays are gone. */
inline void computeDensitiesWrapper(float *vecX, float *vecY, float *vecZ, float *output,
                                    int startX, int endX, int startY, int endY, int vecXYSize,
                                    int vecZSize) {
  cge_context *c = cge_compat::context();
  if (!cge_compat::ok(cge_linspace(c, vecX, vecXYSize, (float)startX, (float)endX))) return;
  if (!cge_compat::ok(cge_linspace(c, vecY, vecXYSize, (float)startY, (float)endY))) return;
  cge_compat::ok(cge_densities_grid(c, output, vecX, vecY, vecZ, vecXYSize, vecZSize));
}

/* inc/wrappers.h:95-121.  Same size check and the same message; no row-pointer matrix, no
 * host-side reshape, no 10 202 allocations. */
inline void computeLikesWrapper(float *densities, double *likes, int densitiesSize, int likesSize) {
  int rows = likesSize;
  int cols = likesSize > 0 ? densitiesSize / likesSize : 0;
  if (likesSize <= 0 || densitiesSize != rows * cols) {
    printf("ERROR: likesSize != rows * cols\n");
    return;
  }
  cge_compat::ok(cge_row_products(cge_compat::context(), densities, likes, (size_t)densitiesSize,
                                  (size_t)likesSize));
}

/* inc/wrappers.h:124-185.  Returns a pointer to function-static storage like the reference
 * (the caller must not free it).  DEVIATION: the reference initialises its static array once,
 * so a second call returns the first call's numbers (wrappers.h:173-176); here every call
 * recomputes. */
inline double *computeExpectationsWrapper(thrust::device_vector<double> &posterior,
                                          float *vectorMu, float *vectorSigma) {
  static double expectations[2];
  cge_context *c = cge_compat::context();
  const int side = (int)std::sqrt((double)posterior.size());
  double *margMu = nullptr, *margSigma = nullptr;
  cudaMallocManaged(&margMu, side * sizeof(double));
  cudaMallocManaged(&margSigma, side * sizeof(double));
  const double *post = thrust::raw_pointer_cast(posterior.data());
  cge_compat::ok(cge_marginalize(c, margMu, post, side, side, 1));     /* row sums: mu */
  cge_compat::ok(cge_marginalize(c, margSigma, post, side, side, 0));  /* column sums: sigma */
  expectations[0] = computeExpectationsCuda(margMu, vectorMu, side);
  expectations[1] = computeExpectationsCuda(margSigma, vectorSigma, side);
  cudaFree(margMu);
  cudaFree(margSigma);
  return expectations;
}

/* ---- the fused path ----------------------------------------------------------------------- */

/* src/grid.cu:79-127 in one call: observations and ranges in; ranges, float32 posterior
 * [mu][sigma], float64 marginals and {E[mu], E[sigma]} out.  Any output pointer may be NULL;
 * pointers may be host, device or managed.  Returns libcge's status code. */
inline int gridPosterior(const float *observations, int nObs, float startMu, float endMu,
                         float startSigma, float endSigma, int gridSize, float *posterior,
                         double *marginalMu, double *marginalSigma, double expectations[2],
                         int mode = CGE_MODE_PRODUCT_F32) {
  cge_params p = {};
  p.n_obs = nObs;
  p.grid_size = gridSize;
  p.mu_start = startMu;
  p.mu_end = endMu;
  p.sigma_start = startSigma;
  p.sigma_end = endSigma;
  p.mode = mode;
  return cge_grid_posterior(cge_compat::context(), &p, observations, posterior, marginalMu,
                            marginalSigma, expectations, nullptr);
}

#endif /* CGE_COMPAT_H */
