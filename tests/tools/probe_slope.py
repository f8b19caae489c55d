#!/usr/bin/env python
"""GPU: accuracy of the shared-slope product policy (forced and as chosen by default) against the
float64 oracle on grids at and beyond its eligibility bound.
usage: python tests/tools/probe_slope.py"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
import oracle  # noqa: E402
from cuda_grid_estimation_b200 import api  # noqa: E402

obs = oracle.readme_observations()
with cge.Context(0) as ctx:
    for N, mu_r, sg_r in ((1026, (19.4, 19.7), (1.7, 2.1)), (700, (19.4, 19.7), (1.7, 2.1)),
                          (4096, (18, 22), (1, 3)), (8192, (18, 22), (1, 3)), (12288, (18, 22), (1, 3))):
        ref = oracle.grid_posterior(obs, N, *mu_r, *sg_r, oracle.NUM_F64)
        Q = ref["posterior"]
        big = Q >= 1e-20 * Q.max()
        row = {"N": N}
        for name, variant in (("default", 0), ("slope", api.VARIANT_SLOPE), ("paired", api.VARIANT_PAIRED),
                              ("wide", api.VARIANT_SLOPE_WIDE), ("tall", api.VARIANT_SLOPE_TALL)):
            res = ctx.grid_posterior(obs, N, mu_r, sg_r, variant=variant)
            info = ctx.scale_info()
            P = res.posterior.astype(np.float64)
            rel = np.abs(P[big] - Q[big]) / Q[big]
            row[name] = {"rel_max": float(rel.max()), "rel_rms": float(np.sqrt((rel ** 2).mean())),
                         "dE": float(np.abs(res.expectations - ref["expectations"]).max())}
            if variant == 0:
                row["path"] = info["product_path"]
                row["umax"] = info["umax"]
                row["slope_eps"] = info["slope_eps"]
                row["bound4"] = api.slope_error_bound(info["umax"], info["slope_eps"], 1.5)
                row["bound8"] = api.slope_error_bound(info["umax"], info["slope_eps"], 3.5)
        print(json.dumps(row), flush=True)
