// cge_launch_logspace.cu -- instantiations and launcher of the log-space likelihood kernels
#include "cge_host.cuh"

namespace cge_host {

// default: float32x2 sums of centred terms in tiles of 256 observations folded into float64
// (LogSpaceTiled), capped at 128 registers; 6: tiles of 64; CGE_VARIANT_LOG_F64: one DFMA per
// pdf-eval (the first log-space kernel, kept as the FP64-pipe case)
int launch_logspace(const Plan &pl, const cge::LikeArgs &args, cudaStream_t st) {
  switch (pl.variant) {
    case CGE_VARIANT_LOG_F64: return launch_like<cge::LogSpaceF64, 1>(pl, args, st);
    case 6: return launch_like<cge::LogSpaceTiled<64>, 2>(pl, args, st);
    default: return launch_like<cge::LogSpaceTiled<256>, 2>(pl, args, st);
  }
}

}  // namespace cge_host
