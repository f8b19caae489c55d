#!/usr/bin/env python
"""GPU: a small tour of every likelihood path for compute-sanitizer (memcheck), sizes kept tiny:
single-launch kernel, three-launch path with each product policy, log-space tiled and float64,
streamed observations, row shards with odd edges, grids that are not multiples of 4 or 128."""
import sys

import numpy as np

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from cuda_grid_estimation_b200 import api  # noqa: E402

rng = np.random.RandomState(42)
obs = rng.normal(20.0, 2.0, 50).astype(np.float32)
big = np.random.default_rng(1).normal(20.0, 2.0, 17000).astype(np.float32)
with cge.Context(0) as ctx:
    for N in (1, 2, 3, 101, 129, 255, 257, 1030):
        for mode in (0, 1):
            r = ctx.grid_posterior(obs, N, (18, 22), (1, 3), mode=mode)
            assert abs(float(r.posterior.astype(np.float64).sum()) - 1) < 1e-6, (N, mode)
    for N, mu_r in ((130, (19.9, 20.1)), (1030, (19.6, 20.4)), (1030, (18, 22)), (520, (18, 22))):
        for variant in (0, api.VARIANT_UNFUSED, api.VARIANT_PAIRED, api.VARIANT_PAIRED3,
                        api.VARIANT_HYBRID, api.VARIANT_MUFU_ONLY, api.VARIANT_SLOPE,
                        api.VARIANT_SLOPE_WIDE, api.VARIANT_SLOPE_TALL):
            r = ctx.grid_posterior(obs, N, mu_r, (1.6, 2.4), variant=variant)
            assert np.isfinite(r.expectations).all()
    for variant in (0, api.VARIANT_LOG_F64):
        r = ctx.grid_posterior(big[:3001], 70, (19.9, 20.1), (1.9, 2.1), mode=1, variant=variant)
        r = ctx.grid_posterior(big, 33, (19.9, 20.1), (1.9, 2.1), mode=1, variant=variant)  # streamed
        assert np.isfinite(r.expectations).all()
    for rows in ((0, 37), (37, 90), (90, 257)):
        for mode in (0, 1):
            r = ctx.grid_posterior(obs, 257, (19.9, 20.1), (1.8, 2.2), mode=mode, rows=rows)
            assert np.isfinite(r.expectations).all()
    pm = np.linspace(0.5, 1.5, 130).astype(np.float32)
    ctx.set_prior(pm, pm, 130)
    ctx.grid_posterior(obs, 130, (18, 22), (1, 3))
    ctx.grid_posterior(obs, 130, (18, 22), (1, 3), variant=api.VARIANT_UNFUSED)
    ctx.set_prior()
    # several shards on one device through the peer exchange (publish flag, poll, local broadcast)
    with cge.MultiContext([0, 0, 0]) as mc:
        for _ in range(3):
            r = mc.grid_posterior(obs, 301, (19.4, 19.7), (1.7, 2.1))
            assert np.isfinite(r.expectations).all()
print("sanitizer tour done")
