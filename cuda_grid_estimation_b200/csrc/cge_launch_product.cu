// cge_launch_product.cu -- instantiations of the float32 product-path likelihood kernels
#include "cge_host.cuh"

namespace cge_host {

template <class Fine, class Mid, class Coarse, bool SELECT, int MINB>
static KernelChoice pick() {
  KernelChoice kc;
  kc.smem_bytes = smem_bytes3<Fine, Mid, Coarse>();
  kc.kcols = Fine::kCols;
  kc.fn = (const void *)cge::grid_likelihood_kernel<Fine, Mid, Coarse, SELECT, MINB, false>;
  return kc;
}

// product path.  Default: the policy is selected per column strip on the device -- ProductNeighbour
// (a third of the exponentials from MUFU.EX2, the rows above and below each derived from it by a
// degree-2 FMA-pipe polynomial) where the strip's exponent step allows it (<= kPairedUmax), the
// same with a degree-3 polynomial up to kPairedUmax3, else ProductHybrid (3/4 of the exponentials
// from MUFU.EX2, 1/4 from a degree-5 polynomial with range reduction).  CGE_VARIANT_PAIRED / _PAIRED3 /
// _HYBRID force one of the three, CGE_VARIANT_MUFU_ONLY is the MUFU-roofline case.
int choose_product(const Plan &pl, KernelChoice *out) {
  switch (pl.variant) {
    case CGE_VARIANT_MUFU_ONLY:
      *out = pick<cge::ProductF32, cge::ProductF32, cge::ProductF32, false, 1>();
      break;
    case CGE_VARIANT_PAIRED:
      *out = pick<DefaultFine, DefaultFine, DefaultFine, false, kDefaultMinB>();
      break;
    case CGE_VARIANT_PAIRED3:
      *out = pick<DefaultMid, DefaultMid, DefaultMid, false, kDefaultMinB>();
      break;
    case CGE_VARIANT_HYBRID:
      *out = pick<DefaultCoarse, DefaultCoarse, DefaultCoarse, false, 1>();
      break;
    default:
      *out = pick<DefaultFine, DefaultMid, DefaultCoarse, true, kDefaultMinB>();
      break;
  }
  return CGE_OK;
}

}  // namespace cge_host
