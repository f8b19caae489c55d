// cge_launch_small.cu -- instantiations and launchers of the single-launch small-grid kernel
#include "cge_host.cuh"

namespace cge_host {

// the whole grid as ONE thread-block cluster (1 x row_tiles x 1 CTAs, row_tiles <= 8)
template <class K>
int launch_as_cluster(K kernel, const Plan &pl, int threads, size_t smem, const cge::SmallArgs &a,
                      cudaStream_t st) {
  if (smem > kDefaultSmemLimit)
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, pl.row_tiles, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = pl.row_tiles;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kernel, a));
  return CGE_OK;
}

template <class Fine, class Mid, class Coarse, int WM, int WN>
int launch_small_shape(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st) {
  dim3 grid(pl.bands, pl.row_tiles), block(WM * WN * 32);
  const size_t smem = like_smem<Fine, WM, WN>(pl);
  if (pl.cluster) {  // one band, <= 8 row tiles: hardware cluster barrier instead of the ticket
    if (pl.vec4)
      return launch_as_cluster(cge::small_grid_kernel<Fine, Mid, Coarse, WM, WN, true, true>, pl,
                               WM * WN * 32, smem, a, st);
    return launch_as_cluster(cge::small_grid_kernel<Fine, Mid, Coarse, WM, WN, false, true>, pl,
                             WM * WN * 32, smem, a, st);
  }
  if (pl.vec4) {
    auto k = cge::small_grid_kernel<Fine, Mid, Coarse, WM, WN, true>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(a);
  } else {
    auto k = cge::small_grid_kernel<Fine, Mid, Coarse, WM, WN, false>;
    if (smem > kDefaultSmemLimit)  // only stages of thousands of observations get there
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(a);
  }
  CK(cudaGetLastError());
  return CGE_OK;
}

template <class Fine, class Mid, class Coarse>
int launch_small(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st) {
  if (pl.wm == 1) return launch_small_shape<Fine, Mid, Coarse, 1, 8>(pl, a, st);
  if (pl.wm == 2) return launch_small_shape<Fine, Mid, Coarse, 2, 4>(pl, a, st);
  if (pl.wm == 4) return launch_small_shape<Fine, Mid, Coarse, 4, 2>(pl, a, st);
  return launch_small_shape<Fine, Mid, Coarse, 8, 1>(pl, a, st);
}

int launch_small_product(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st) {
  return launch_small<DefaultFine, DefaultMid, DefaultCoarse>(pl, a, st);
}

int launch_small_logspace(const Plan &pl, const cge::SmallArgs &a, cudaStream_t st) {
  return launch_small<cge::LogSpaceTiled<256>, cge::LogSpaceTiled<256>, cge::LogSpaceTiled<256>>(pl, a, st);
}

}  // namespace cge_host
