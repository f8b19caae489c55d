// cge_launch_small.cu -- instantiations of the single-launch small-grid kernel
#include "cge_host.cuh"

namespace cge_host {

template <class Finest, class Fine, class Mid, class Coarse, bool SELECT>
static KernelChoice pick(const Plan &pl) {
  KernelChoice kc;
  kc.smem_bytes = smem_bytes4<Finest, Fine, Mid, Coarse>();
  kc.kcols = Fine::kCols;
  kc.fn = pl.cluster ? (const void *)cge::small_grid_kernel<Finest, Fine, Mid, Coarse, SELECT, true>
                     : (const void *)cge::small_grid_kernel<Finest, Fine, Mid, Coarse, SELECT, false>;
  return kc;
}

int choose_small_product(const Plan &pl, KernelChoice *out) {
  *out = pick<DefaultFinest, DefaultFine, DefaultMid, DefaultCoarse, true>(pl);
  return CGE_OK;
}

int choose_small_logspace(const Plan &pl, KernelChoice *out) {
  using L = cge::LogSpaceTiled<256>;
  *out = pick<L, L, L, L, false>(pl);
  return CGE_OK;
}

}  // namespace cge_host
