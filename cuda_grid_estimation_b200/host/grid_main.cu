// grid_main.cu -- command-line equivalent of the reference's bin/grid (src/grid.cu:19-142) on the
// fused libcge path.  The reference "does not accept parameters" (README.md:120); this one takes
// the sizes, ranges and mode it hard-codes, and reads the observations from a file (raw float32)
// or draws them from N(mu, sigma) with a host generator.  (The reference draws them with cuRAND
// XORWOW on the device -- input synthesis, not part of the hot path.)  Output lines match
// src/grid.cu:129-132.
//
//   grid [--grid N] [--obs n] [--mu a b] [--sigma a b] [--true-mu m] [--true-sigma s]
//        [--mode product|log] [--obs-file path] [--seed k] [--curand] [--timing] [--gpus G | --devices a,b,..]
//   --gpus G (G > 1) row-shards the grid over devices 0..G-1 from this one process
//   (cge_create_multi / cge_grid_posterior_multi: one host thread, one stream per GPU, the
//   partial sums exchanged over NVLink peer pointers); the printed lines are the same.
//   --devices lists the devices instead (a device may repeat: its shards then run back to back,
//   which exercises the exchange on a single-GPU box).
//   --curand draws the observations on the device exactly like the reference's
//   generateNormalCuda (cuRAND XORWOW, seed 42): with no other argument the program then
//   reproduces bin/grid's run, observation for observation.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cge.h"

int main(int argc, char **argv) {
  int N = 101, n = 50, mode = CGE_MODE_PRODUCT_F32, seed = 42, gpus = 1;
  float mu = 20.0f, sigma = 2.0f;
  float mu0 = mu - 2, mu1 = mu + 2, sg0 = sigma - 1, sg1 = sigma + 1;  // src/grid.cu:67-70
  bool ranges_set = false, timing = false, use_curand = false;
  const char *obs_file = nullptr;
  std::vector<int> devices;
  for (int i = 1; i < argc; ++i) {
    auto need = [&](int k) { if (i + k >= argc) { std::fprintf(stderr, "missing value for %s\n", argv[i]); std::exit(2); } };
    if (!std::strcmp(argv[i], "--grid")) { need(1); N = std::atoi(argv[++i]); }
    else if (!std::strcmp(argv[i], "--obs")) { need(1); n = std::atoi(argv[++i]); }
    else if (!std::strcmp(argv[i], "--mu")) { need(2); mu0 = std::atof(argv[++i]); mu1 = std::atof(argv[++i]); ranges_set = true; }
    else if (!std::strcmp(argv[i], "--sigma")) { need(2); sg0 = std::atof(argv[++i]); sg1 = std::atof(argv[++i]); ranges_set = true; }
    else if (!std::strcmp(argv[i], "--true-mu")) { need(1); mu = std::atof(argv[++i]); }
    else if (!std::strcmp(argv[i], "--true-sigma")) { need(1); sigma = std::atof(argv[++i]); }
    else if (!std::strcmp(argv[i], "--mode")) { need(1); ++i; mode = !std::strcmp(argv[i], "log") ? CGE_MODE_LOGSPACE : CGE_MODE_PRODUCT_F32; }
    else if (!std::strcmp(argv[i], "--obs-file")) { need(1); obs_file = argv[++i]; }
    else if (!std::strcmp(argv[i], "--seed")) { need(1); seed = std::atoi(argv[++i]); }
    else if (!std::strcmp(argv[i], "--gpus")) { need(1); gpus = std::atoi(argv[++i]); }
    else if (!std::strcmp(argv[i], "--devices")) {
      need(1);
      for (const char *s = argv[++i]; *s;) {
        char *end = nullptr;
        devices.push_back((int)std::strtol(s, &end, 10));
        if (end == s) { std::fprintf(stderr, "bad device list %s\n", argv[i]); return 2; }
        s = *end == ',' ? end + 1 : end;
      }
      gpus = (int)devices.size();
    }
    else if (!std::strcmp(argv[i], "--timing")) timing = true;
    else if (!std::strcmp(argv[i], "--curand")) use_curand = true;
    else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
  }
  if (!ranges_set) { mu0 = mu - 2; mu1 = mu + 2; sg0 = sigma - 1; sg1 = sigma + 1; }

  cge_context *ctx = nullptr;
  if (cge_create(0, &ctx) != CGE_OK) {
    std::fprintf(stderr, "Error: %s\n", cge_last_error());
    return EXIT_FAILURE;
  }
  std::vector<float> obs;
  if (use_curand) {
    float *dev = nullptr;
    cudaMalloc(&dev, n * sizeof(float));
    if (cge_generate_normal(ctx, dev, n, mu, sigma, (unsigned long long)seed, 0) != CGE_OK) {
      std::fprintf(stderr, "Error: %s\n", cge_last_error());
      return EXIT_FAILURE;
    }
    obs.resize(n);
    cudaMemcpy(obs.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    std::printf("observations[0] = %.4f\n", obs[0]);  // the reference checks 18.5689 (grid.cu:41-47)
  } else if (obs_file) {
    FILE *f = std::fopen(obs_file, "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", obs_file); return 2; }
    float v;
    while (std::fread(&v, sizeof v, 1, f) == 1) obs.push_back(v);
    std::fclose(f);
    n = (int)obs.size();
  } else {
    std::mt19937 gen(seed);
    std::normal_distribution<float> dist(mu, sigma);
    obs.resize(n);
    for (auto &x : obs) x = dist(gen);
  }

  cge_params p = {};
  p.n_obs = n; p.grid_size = N; p.mu_start = mu0; p.mu_end = mu1;
  p.sigma_start = sg0; p.sigma_end = sg1; p.mode = mode;
  double e[2];
  cge_timing t;
  int rc;
  cge_multi *group = nullptr;
  if (gpus > 1) {  // one process, `gpus` devices: mu rows sharded in device order
    rc = cge_create_multi(gpus, devices.empty() ? nullptr : devices.data(), &group);
    if (rc == CGE_OK) rc = cge_grid_posterior_multi(group, &p, obs.data(), nullptr, nullptr, nullptr, e, &t);
  } else {
    rc = cge_grid_posterior(ctx, &p, obs.data(), nullptr, nullptr, nullptr, e, &t);
  }
  if (rc != CGE_OK) {
    std::fprintf(stderr, "Error: %s\n", cge_last_error());
    if (group) cge_destroy_multi(group);
    cge_destroy(ctx);
    return EXIT_FAILURE;
  }
  std::cout.precision(10);
  std::cout << "Inferred mu: " << e[0] << "; Actual mu: " << mu << std::endl;
  std::cout << "Inferred sigma: " << e[1] << "; Actual sigma: " << sigma << std::endl;
  if (timing)
    std::printf("device ms: setup %.4f likelihood %.4f reduce %.4f scale %.4f | total incl. copies %.4f\n",
                t.setup_ms, t.likelihood_ms, t.reduce_ms, t.scale_ms, t.total_ms);
  if (group) cge_destroy_multi(group);
  cge_destroy(ctx);
  return 0;
}
