// cge_launch_product.cu -- instantiations of the float32 product-path likelihood kernels
#include "cge_host.cuh"

namespace cge_host {

template <class Finest, class Fine, class Mid, class Coarse, bool SELECT, int MINB, int WARPS = cge::kWarps,
          bool WIDE = false>
static KernelChoice pick() {
  KernelChoice kc;
  kc.smem_bytes = smem_bytes4<Finest, Fine, Mid, Coarse>();
  if (WIDE && cge::policy_smem_bytes<cge::ProductSlopeWide>() > kc.smem_bytes)
    kc.smem_bytes = cge::policy_smem_bytes<cge::ProductSlopeWide>();
  if (WIDE && cge::policy_smem_bytes<cge::ProductSlopeTall>() > kc.smem_bytes)
    kc.smem_bytes = cge::policy_smem_bytes<cge::ProductSlopeTall>();
  kc.kcols = Fine::kCols;
  kc.threads = WARPS * 32;
  kc.fn = (const void *)cge::grid_likelihood_kernel<Finest, Fine, Mid, Coarse, SELECT, MINB, false, WARPS, WIDE>;
  return kc;
}

// product path.  Default: the policy is selected per column strip on the device -- neighbour rows
// (a third of the exponentials from MUFU.EX2, the rows above and below each derived from it on the
// FMA pipe) where the strip's exponent step allows it: ProductSlope on the finest grids, then
// ProductNeighbour with a degree-2 polynomial (<= kPairedUmax) or degree 3 (<= kPairedUmax3), else
// ProductHybrid (3/4 of the exponentials from MUFU.EX2, 1/4 from a degree-5 polynomial with range
// reduction).  CGE_VARIANT_SLOPE / _PAIRED / _PAIRED3 / _HYBRID force one of the four,
// CGE_VARIANT_MUFU_ONLY is the MUFU-roofline case.
int choose_product(const Plan &pl, KernelChoice *out) {
  using F = cge::ProductF32;
  switch (pl.variant) {
    case CGE_VARIANT_MUFU_ONLY:
      *out = pick<F, F, F, F, false, 1>();
      break;
    case CGE_VARIANT_SLOPE:
      *out = pick<DefaultFinest, DefaultFinest, DefaultFinest, DefaultFinest, false, kDefaultMinB>();
      break;
    case CGE_VARIANT_PAIRED:
      *out = pick<DefaultFine, DefaultFine, DefaultFine, DefaultFine, false, kDefaultMinB>();
      break;
    case CGE_VARIANT_PAIRED3:
      *out = pick<DefaultMid, DefaultMid, DefaultMid, DefaultMid, false, kDefaultMinB>();
      break;
    case CGE_VARIANT_HYBRID:
      *out = pick<DefaultCoarse, DefaultCoarse, DefaultCoarse, DefaultCoarse, false, 1>();
      break;
    default:
      // one 16-warp CTA per SM (a single pool for all its warps) that also holds the 8-column
      // tall and wide policies
      *out = pick<DefaultFinest, DefaultFine, DefaultMid, DefaultCoarse, true, 1, kWideWarps, true>();
      break;
  }
  return CGE_OK;
}

}  // namespace cge_host
