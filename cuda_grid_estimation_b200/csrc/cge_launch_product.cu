// cge_launch_product.cu -- instantiations and launcher of the float32 product-path likelihood kernels
#include "cge_host.cuh"

namespace cge_host {

template <class Fine, class Mid, class Coarse, int WM, int WN, int MINB = 1>
int launch_select_shape(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  dim3 grid(pl.bands, pl.row_tiles), block(WM * WN * 32);
  const size_t smem = like_smem<Fine, WM, WN>(pl);
  if (pl.vec4) {
    auto k = cge::grid_likelihood_select_kernel<Fine, Mid, Coarse, WM, WN, true, MINB>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(args);
  } else {
    auto k = cge::grid_likelihood_select_kernel<Fine, Mid, Coarse, WM, WN, false, MINB>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(args);
  }
  CK(cudaGetLastError());
  return CGE_OK;
}

template <class Fine, class Mid, class Coarse, int MINB>
int launch_select(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  if (pl.wm == 1) return launch_select_shape<Fine, Mid, Coarse, 1, 8, MINB>(pl, args, st);
  if (pl.wm == 2) return launch_select_shape<Fine, Mid, Coarse, 2, 4, MINB>(pl, args, st);
  if (pl.wm == 4) return launch_select_shape<Fine, Mid, Coarse, 4, 2, MINB>(pl, args, st);
  return launch_select_shape<Fine, Mid, Coarse, 8, 1, MINB>(pl, args, st);
}

// product path.  Default: grid_likelihood_select_kernel -- ProductPaired (half the exponentials
// from MUFU.EX2, the other half derived from the row above by a degree-2 FMA-pipe polynomial)
// when the grid is fine enough (ScaleInfo::umax <= kPairedUmax, decided on the device), the same
// with a degree-3 polynomial up to kPairedUmax3, else ProductHybrid<2, 2, false, 1> (12 of 16
// exponentials from MUFU.EX2, 4 from a degree-5 polynomial with range reduction).
// CGE_VARIANT_PAIRED / _PAIRED3 / _HYBRID force one of the three, CGE_VARIANT_MUFU_ONLY is the
// roofline case.  A few tuning variants of the paired kernel stay built (wide tile shape only) so
// the sweeps in profiles/r01_variant_sweep_*.jsonl can be re-run; the other ~25 variants those
// files record (hybrid splits, loop shapes, register caps) were removed after the sweeps.
int launch_product(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  switch (pl.variant) {
    case CGE_VARIANT_MUFU_ONLY: return launch_like<cge::ProductF32>(pl, args, st);
    case CGE_VARIANT_PAIRED: return launch_like<DefaultFine, kDefaultMinB>(pl, args, st);
    case CGE_VARIANT_PAIRED3: return launch_like<DefaultMid, kDefaultMinB>(pl, args, st);
    case CGE_VARIANT_HYBRID: return launch_like<DefaultCoarse>(pl, args, st);
    default: break;
  }
  if (pl.wm == 1) {
    switch (pl.variant) {
      // 31: one observation per iteration, scalars unpacked, column sums in registers, uncapped
      // (the first paired kernel); 39 / 45: the default at 128 registers (2 CTAs per SM) with
      // the column sums in registers / shared memory; 381: 8 columns per thread
      case 31: return launch_like_shape<cge::ProductPaired<4, 1>, 1, 8>(pl, args, st);
      case 39: return launch_like_shape<cge::ProductPaired<4, 4, true>, 1, 8, 2>(pl, args, st);
      case 45: return launch_like_shape<cge::ProductPaired<4, 4, true, true>, 1, 8, 2>(pl, args, st);
      case 381: return launch_like_shape<cge::ProductPaired<8, 1>, 1, 8>(pl, args, st);
      default: break;
    }
  }
  if (pl.variant == 381)
    return set_error(CGE_ERR_BAD_ARG, "variant %d needs grid_size >= 2048", pl.variant);
  return launch_select<DefaultFine, DefaultMid, DefaultCoarse, kDefaultMinB>(pl, args, st);
}

}  // namespace cge_host
