// cge_host.cuh -- host-side pieces shared by the translation units of libcge.so: error
// reporting, the launch plan, and the kernel choosers.  The likelihood kernel instantiations live
// in their own .cu files (cge_launch_*.cu) so nvcc compiles them in parallel.
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

// shape of one likelihood launch (made by make_plan in cge_api.cu)
struct Plan {
  bool valid = false;
  int N = 0, n = 0, n_pad = 0, mode = 0, variant = 0, post_f64 = 0;
  int row_begin = 0, row_end = 0, rows_local = 0;
  int kcols = 4;           // columns per thread: a warp's strip is 32 * kcols columns
  int ncw = 0;             // strips per row
  int krows = 4;           // rows per group (Policy::kRows)
  int groups = 0;          // row groups that cover the shard's rows
  int chunk = 1;           // row groups per claim
  int items_per_strip = 0;
  int fixed_bits = 0;      // 58 - ceil(log2(addends per column))
  int nctas = 0;           // P: persistent CTAs (resident slots of the device)
  int nb_col = 0, nb_row = 0, nb_res = 0;
  bool vec4 = false, keep_l2 = false, streamed = false, reorder = false;
  bool cluster = false;    // small-grid path: launch the grid as one thread-block cluster
  size_t smem = 0;         // dynamic shared memory of the likelihood kernel
};

// one instantiation of the likelihood (or small-grid) kernel: every one takes a single by-value
// argument struct, so they all launch through cudaLaunchKernelExC
struct KernelChoice {
  const void *fn = nullptr;
  int threads = 256;
  int smem_bytes = 0;    // per-thread bytes behind the observation stage
  int kcols = 4;
};

// the default product policies, finest first (chosen per column strip on the device)
using DefaultFinest = cge::ProductSlope;       // shared-slope neighbours: the finest grids
using DefaultFine = cge::ProductNeighbour<2>;  // degree-2 polynomial, strip umax <= 0.011
using DefaultMid = cge::ProductNeighbour<3>;   // degree-3 polynomial, strip umax <= 0.05
using DefaultCoarse = cge::ProductHybrid;
constexpr int kDefaultMinB = 3;  // forced 4-column policies: <= 80 registers, three 256-thread CTAs per SM
constexpr int kWideWarps = 16;   // default kernel: one 512-thread CTA per SM -- one pool for all its
                                 // warps, and 128 registers per thread for the 8-column policies

template <class A, class B, class C, class D>
constexpr int smem_bytes4() {
  constexpr int a = cge::policy_smem_bytes<A>(), b = cge::policy_smem_bytes<B>(),
                c = cge::policy_smem_bytes<C>(), d = cge::policy_smem_bytes<D>();
  constexpr int ab = a > b ? a : b, cd = c > d ? c : d;
  return ab > cd ? ab : cd;
}

// choosers defined in cge_launch_product.cu / cge_launch_logspace.cu / cge_launch_small.cu;
// they read pl.variant, pl.streamed (and pl.cluster for the small-grid kernels)
int choose_product(const Plan &pl, KernelChoice *out);
int choose_logspace(const Plan &pl, KernelChoice *out);
int choose_small_product(const Plan &pl, KernelChoice *out);
int choose_small_logspace(const Plan &pl, KernelChoice *out);

}  // namespace cge_host
