// test_compat.cu -- the reference's unit tests (cpp/grid-algorithm/src/tests.cu) restated as
// plain checks against include/cge_compat.h, plus the reference main()'s call sequence
// (src/grid.cu:79-127) through the same names.  googletest is not available in this
// environment, so this is a small self-contained runner: exit code 0 = all passed.
// The RNG test (tests.cu:21-46) is out of scope (observations are an input of the hot path).
//
// usage: test_compat [obs.bin]   (obs.bin: float32 observations for the pipeline check; the
//        pipeline prints "PIPELINE E_mu E_sigma" for the pytest wrapper to compare with the oracle)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/cge_compat.h"

static int failures = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);       \
      ++failures;                                                         \
    }                                                                     \
  } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((double)(a) - (double)(b)) <= (tol))

template <class T>
static T *managed(size_t count) {
  T *p = nullptr;
  cudaMallocManaged(&p, count * sizeof(T));
  return p;
}

static void testLinspaces() {  // tests.cu:49-67
  float *v = managed<float>(3);
  linspaceCuda(v, 3, 1.0f, 2.0f);
  const float want[3] = {1.0f, 1.5f, 2.0f};
  CHECK(std::equal(v, v + 3, want));
  cudaFree(v);
}

static void testGenerateGrids() {  // tests.cu:70-108
  float *x = managed<float>(3), *y = managed<float>(3), *z = managed<float>(2);
  float *gx = managed<float>(18), *gy = managed<float>(18), *gz = managed<float>(18);
  linspaceCuda(x, 3, 1, 3);
  linspaceCuda(y, 3, 1, 3);
  linspaceCuda(z, 2, 4, 5);
  create3dGridCuda(x, y, z, gx, gy, gz, 3, 2);
  const float ex[18] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3};
  const float ey[18] = {1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3};
  const float ez[18] = {4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5};
  CHECK(std::equal(gx, gx + 18, ex));
  CHECK(std::equal(gy, gy + 18, ey));
  CHECK(std::equal(gz, gz + 18, ez));
  cudaFree(x); cudaFree(y); cudaFree(z); cudaFree(gx); cudaFree(gy); cudaFree(gz);
}

static void testComputeDensities() {  // tests.cu:111-135
  float *x = managed<float>(3), *y = managed<float>(3), *z = managed<float>(3), *d = managed<float>(3);
  linspaceCuda(x, 3, 20.0f, 22.0f);
  linspaceCuda(y, 3, 1.0f, 3.0f);
  linspaceCuda(z, 3, 20.0f, 21.0f);
  computeDensitiesCuda(d, x, y, z, 3);
  CHECK_NEAR(d[0], 0.39894228, 1e-5);
  CHECK_NEAR(d[1], 0.19333406, 1e-5);
  CHECK_NEAR(d[2], 0.12579441, 1e-5);
  cudaFree(x); cudaFree(y); cudaFree(z); cudaFree(d);
}

static void testReshapeArray() {  // tests.cu:138-160
  float flat[10] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
  double r0[5], r1[5];
  double *out[2] = {r0, r1};
  reshapeArray<float, double>(flat, out, 5, 2);
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 5; j++) CHECK(out[i][j] == flat[j + i * 5]);
}

static void testComputeLikes() {  // tests.cu:163-194
  float *dens = managed<float>(10);
  double *likes = managed<double>(2);
  double **m = allocateMatrix(2, 5);
  linspaceCuda(dens, 10, 1.0f, 10.0f);
  reshapeArray<float, double>(dens, m, 5, 2);
  computeLikesCuda(likes, m, 2, 5);
  CHECK(likes[0] == 120);    // 1*2*3*4*5
  CHECK(likes[1] == 30240);  // 6*7*8*9*10
  // and the wrapper on the flat float densities (wrappers.h:95-121)
  likes[0] = likes[1] = 0;
  computeLikesWrapper(dens, likes, 10, 2);
  CHECK(likes[0] == 120 && likes[1] == 30240);
  freeMatrix(m, 2);
  cudaFree(dens); cudaFree(likes);
}

static void testComputePosterior() {  // tests.cu:196-209
  thrust::device_vector<double> likesV(3), postV(3);
  likesV[0] = 1.0; likesV[1] = 2.0; likesV[2] = 3.0;
  computePosteriorCuda(likesV, postV);
  CHECK_NEAR(postV[0], 0.166667, 1e-5);
  CHECK_NEAR(postV[1], 0.333333, 1e-5);
  CHECK((double)postV[2] == 0.5);
}

static void testComputeExpectations() {  // tests.cu:211-245
  thrust::device_vector<double> post(9, 0.1);
  post[4] = 0.2;
  float *mu = managed<float>(3), *sg = managed<float>(3);
  linspaceCuda(mu, 3, 1.0f, 3.0f);
  linspaceCuda(sg, 3, 4.0f, 6.0f);
  double *e = computeExpectationsWrapper(post, mu, sg);
  CHECK(e[0] == 2.0f);
  CHECK(e[1] == 5.0f);
  // DEVIATION from wrappers.h:173: a second call must give the second call's answer
  post[4] = 0.1; post[8] = 0.2;  // mass moved to (mu=3, sigma=6)
  e = computeExpectationsWrapper(post, mu, sg);
  CHECK_NEAR(e[0], 2.1, 1e-12);
  CHECK_NEAR(e[1], 5.1, 1e-12);
  // marginalizeCuda on the row-pointer layout (wrappers.h:152-170)
  thrust::host_vector<double> h = post;
  double **m = allocateMatrix(3, 3);
  reshapeArray<double, double>(h.data(), m, 3, 3);
  double *mm = managed<double>(3), *ms = managed<double>(3);
  marginalizeCuda(mm, m, 3, 3, 1);
  marginalizeCuda(ms, m, 3, 3, 0);
  CHECK_NEAR(mm[2], 0.4, 1e-15);
  CHECK_NEAR(ms[2], 0.4, 1e-15);
  CHECK_NEAR(mm[0] + mm[1] + mm[2], 1.0, 1e-15);
  freeMatrix(m, 3);
  cudaFree(mm); cudaFree(ms); cudaFree(mu); cudaFree(sg);
}

// src/grid.cu:19-142 with caller-supplied observations: stage by stage through the reference's
// names, then the fused entry point; both must agree.
static void pipeline(const char *path) {
  std::vector<float> host;
  if (FILE *f = std::fopen(path, "rb")) {
    float v;
    while (std::fread(&v, sizeof v, 1, f) == 1) host.push_back(v);
    std::fclose(f);
  }
  if (host.empty()) { std::printf("FAILED cannot read %s\n", path); ++failures; return; }
  const int rvsSize = (int)host.size(), vecSize = 101, gridSize = vecSize * vecSize * rvsSize;
  float *observations = managed<float>(rvsSize);
  std::copy(host.begin(), host.end(), observations);
  float *vectorMu = managed<float>(vecSize), *vectorSigma = managed<float>(vecSize);
  float *densities = managed<float>(gridSize);
  computeDensitiesWrapper(vectorMu, vectorSigma, observations, densities, 18, 22, 1, 3, vecSize, rvsSize);
  CHECK(vectorMu[50] == 20.0f && vectorMu[49] == 19.96f);       // utils.h:16-30
  CHECK(vectorSigma[1] == 1.02f && vectorSigma[0] == 1.0f);
  const int likesSize = vecSize * vecSize;
  double *likes = managed<double>(likesSize);
  computeLikesWrapper(densities, likes, gridSize, likesSize);
  thrust::device_vector<double> likesV(likes, likes + likesSize), posteriorV(likesSize);
  computePosteriorCuda(likesV, posteriorV);
  double *e = computeExpectationsWrapper(posteriorV, vectorMu, vectorSigma);
  std::printf("PIPELINE %.17g %.17g\n", e[0], e[1]);

  double fused[2] = {0, 0};
  std::vector<float> post(likesSize);
  std::vector<double> mm(vecSize), ms(vecSize);
  int rc = gridPosterior(observations, rvsSize, 18, 22, 1, 3, vecSize, post.data(), mm.data(), ms.data(), fused);
  CHECK(rc == CGE_OK);
  std::printf("FUSED %.17g %.17g\n", fused[0], fused[1]);
  CHECK_NEAR(fused[0], e[0], 1e-6);
  CHECK_NEAR(fused[1], e[1], 1e-6);
  thrust::host_vector<double> hp = posteriorV;
  double worst = 0, pmax = *std::max_element(hp.begin(), hp.end());
  for (int f = 0; f < likesSize; ++f)
    if (hp[f] >= 1e-20 * pmax) worst = std::max(worst, std::fabs(post[f] - hp[f]) / hp[f]);
  std::printf("FUSED_VS_STAGED_MAX_REL %.3e\n", worst);
  CHECK(worst <= 1e-4);
  cudaFree(observations); cudaFree(vectorMu); cudaFree(vectorSigma); cudaFree(densities); cudaFree(likes);
}

int main(int argc, char **argv) {
  testLinspaces();
  testGenerateGrids();
  testComputeDensities();
  testReshapeArray();
  testComputeLikes();
  testComputePosterior();
  testComputeExpectations();
  if (argc > 1) pipeline(argv[1]);
  if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("FAILED cuda error\n"); ++failures; }
  std::printf(failures ? "%d CHECK(S) FAILED\n" : "ALL PASSED\n", failures);
  return failures ? 1 : 0;
}
