// cge_host.cuh -- host-side pieces shared by the translation units of libcge.so: error
// reporting, the launch plan, and the launchers of the templated kernels.  The likelihood kernel
// instantiations live in their own .cu files (cge_launch_*.cu) so nvcc compiles them in parallel.
#pragma once
#include <cuda_runtime.h>

#include "../../include/cge.h"
#include "cge_kernels.cuh"

namespace cge_host {

// records the message for cge_last_error() (thread-local) and returns `code`
int set_error(int code, const char *fmt, ...);

#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return ::cge_host::set_error(CGE_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                                   cudaGetErrorString(e_), __FILE__, __LINE__);                  \
  } while (0)

#define CKRC(expr)                 \
  do {                             \
    int rc_ = (expr);              \
    if (rc_ != CGE_OK) return rc_; \
  } while (0)

// tile shape and grid of one likelihood launch (made by make_plan in cge_api.cu)
struct Plan {
  bool valid = false;
  int N = 0, n = 0, n_pad = 0, mode = 0, variant = 0;
  int row_begin = 0, row_end = 0, rows_local = 0;
  int wm = 1, wn = 8;
  int bands = 0, row_tiles = 0, tile_rows = 0;
  int ncw = 0, ncolparts = 0;
  int nb_col = 0, nb_row = 0;
  bool vec4 = false;
  bool cluster = false;  // small-grid path: launch the grid as one thread-block cluster
  size_t smem = 0;
};

// the default product policies (cge_launch_product.cu, cge_launch_small.cu)
using DefaultFine = cge::ProductPaired<4, 4, true, true>;     // degree-2 polynomial, umax <= 0.011
using DefaultMid = cge::ProductPaired<4, 4, true, true, 3>;   // degree-3 polynomial, umax <= 0.05
using DefaultCoarse = cge::ProductHybrid<2, 2, false, 1>;
constexpr int kDefaultMinB = 3;  // <= 80 registers: three 256-thread CTAs (24 warps) per SM

// static + dynamic shared memory beyond 48 KB needs the opt-in attribute; the kernels have up to
// ~17 KB static (small_grid_kernel), so opt in above 24 KB of dynamic and skip the call below it
constexpr size_t kDefaultSmemLimit = 24 * 1024;

// dynamic shared memory of a likelihood launch: mbarriers + observation stage (+ column sums)
template <class Policy, int WM, int WN>
inline size_t like_smem(const Plan &pl) {
  return pl.smem + 16 + (Policy::kColSmem ? (size_t)WM * WN * 32 * Policy::kCols * 8 : 0);
}

template <class Policy, int WM, int WN, int MINB = 1>
int launch_like_shape(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  dim3 grid(pl.bands, pl.row_tiles), block(WM * WN * 32);
  const size_t smem = like_smem<Policy, WM, WN>(pl);
  if (pl.vec4) {
    auto k = cge::grid_likelihood_kernel<Policy, WM, WN, true, MINB>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(args);
  } else {
    auto k = cge::grid_likelihood_kernel<Policy, WM, WN, false, MINB>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(args);
  }
  CK(cudaGetLastError());
  return CGE_OK;
}

template <class Policy, int MINB = 1>
int launch_like(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  if (pl.wm == 1) return launch_like_shape<Policy, 1, 8, MINB>(pl, args, st);
  if (pl.wm == 2) return launch_like_shape<Policy, 2, 4, MINB>(pl, args, st);
  if (pl.wm == 4) return launch_like_shape<Policy, 4, 2, MINB>(pl, args, st);
  return launch_like_shape<Policy, 8, 1, MINB>(pl, args, st);
}

// launchers defined in cge_launch_product.cu / cge_launch_logspace.cu / cge_launch_small.cu
int launch_product(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st);
int launch_logspace(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st);
int launch_small_product(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st);
int launch_small_logspace(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st);

}  // namespace cge_host
