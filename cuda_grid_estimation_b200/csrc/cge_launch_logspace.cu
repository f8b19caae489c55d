// cge_launch_logspace.cu -- instantiations of the log-space likelihood kernels
#include "cge_host.cuh"

namespace cge_host {

constexpr int kLogWarps = 16;

template <class P, int MINB, bool STREAMED, int WARPS = cge::kWarps>
static KernelChoice pick() {
  KernelChoice kc;
  kc.smem_bytes = cge::policy_smem_bytes<P>();
  kc.kcols = P::kCols;
  kc.threads = WARPS * 32;
  kc.fn = (const void *)cge::grid_likelihood_kernel<P, P, P, P, false, MINB, STREAMED, WARPS>;
  return kc;
}

// default: float32x2 sums of centred terms in tiles of 256 observations folded into float64
// (LogSpaceTiled), capped at 128 registers; CGE_VARIANT_LOG_F64: one DFMA per pdf-eval (the
// first log-space kernel, kept as the FP64-pipe case).  pl.streamed (more than 16384
// observations): the instantiation that streams them through two 32 KB shared-memory stages.
int choose_logspace(const Plan &pl, KernelChoice *out) {
  if (pl.variant == CGE_VARIANT_LOG_F64)
    *out = pl.streamed ? pick<cge::LogSpaceF64, 1, true>() : pick<cge::LogSpaceF64, 1, false>();
  else
    // (<= 128 registers: like the product kernel, one 16-warp CTA per SM whose warps share a
    // single pool -- 107.6 ms against 112.5 for two 8-warp CTAs with a pool each; the streamed form
    // moves as a CTA and keeps two 8-warp CTAs per SM)
    *out = pl.streamed ? pick<cge::LogSpaceTiled<256>, 2, true>()
                       : pick<cge::LogSpaceTiled<256>, 1, false, kLogWarps>();
  return CGE_OK;
}

}  // namespace cge_host
