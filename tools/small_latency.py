#!/usr/bin/env python
"""GPU: end-to-end latency of cge_grid_posterior on small grids (host observations in;
expectations and marginals to the host; posterior left on the device), single-launch path
against the three-launch path."""
import json
import sys
import time

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402
from cuda_grid_estimation_b200 import api  # noqa: E402

obs = observations()
with cge.Context(0) as ctx:
    for N in (101, 256, 512, 1024):
        for mode in (0, 1):
            for variant in (0, api.VARIANT_UNFUSED):
                kw = dict(mode=mode, posterior="device", variant=variant)
                for _ in range(50):
                    ctx.grid_posterior(obs, N, (18, 22), (1, 3), want_timing=False, **kw)
                t0 = time.perf_counter()
                for _ in range(1000):
                    ctx.grid_posterior(obs, N, (18, 22), (1, 3), want_timing=False, **kw)
                dt = (time.perf_counter() - t0) / 1000
                rt = ctx.grid_posterior(obs, N, (18, 22), (1, 3), **kw)
                print(json.dumps(dict(N=N, mode=mode, variant=variant, e2e_us=dt * 1e6,
                                      device_ms=rt.timing, launches=rt.launches,
                                      E=rt.expectations.tolist())), flush=True)
