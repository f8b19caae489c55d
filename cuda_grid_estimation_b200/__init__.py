"""cuda_grid_estimation_b200 -- B200-native (sm_100a) grid likelihood -> posterior -> marginals
-> expectations for a Normal model on a (mu, sigma) grid: the one data-parallel hot path of
nablabits/cuda-grid-estimation (cpp/grid-algorithm/src/grid.cu:79-127), behind a C ABI
(include/cge.h -> libcge.so).  This package is the thin Python host side used by the tests and
the benchmark: ctypes marshalling, row sharding, and the torch.distributed plumbing around the
one collective of the path.  All arithmetic runs in the CUDA library; there is no CPU fallback.
"""
from .api import (MODE_LOGSPACE, MODE_PRODUCT_F32, POSTERIOR_F32, POSTERIOR_F64, CgeError, Context,
                  GridResult, MultiContext, Params, Timing, device_tensor, load_library,
                  make_params)
from .sharding import ShardedGridPosterior, finalize_from_partials, pack_partials, shard_rows

__all__ = [
    "MODE_LOGSPACE", "MODE_PRODUCT_F32", "POSTERIOR_F32", "POSTERIOR_F64", "CgeError", "Context",
    "GridResult", "MultiContext", "Params", "Timing",
    "device_tensor", "load_library", "make_params", "ShardedGridPosterior",
    "finalize_from_partials", "pack_partials", "shard_rows",
]
