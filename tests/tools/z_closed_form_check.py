#!/usr/bin/env python
"""GPU experiment: how far is the normaliser of the float32 product kernel (the sum of the cells it
actually wrote) from the closed-form normaliser (float64, sufficient statistics)?  Decides whether
normalising on write with the closed-form value could replace the scale pass.
(lives under tests/: it uses the oracle, which is test infrastructure)"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
import oracle  # noqa: E402


def closed_form_log2z(obs, N, ranges, chunk=1024):
    m, s = -np.inf, 0.0
    for r0 in range(0, N, chunk):
        ll = oracle.closed_form_loglik(obs, N, *ranges, rows=np.arange(r0, min(N, r0 + chunk)))
        mc = ll.max()
        if mc > m:
            s = s * np.exp(m - mc) if np.isfinite(m) else 0.0
            m = mc
        s += np.exp(ll - m).sum()
    return (m + np.log(s)) / np.log(2.0)


with cge.Context(0) as ctx:
    for N, n, seed in ((101, 50, 42), (1024, 50, 42), (4096, 50, 42), (16384, 50, 42), (4096, 20, 1),
                       (4096, 120, 2), (4096, 300, 3), (2048, 50, 5), (8192, 50, 6)):
        obs = oracle.readme_observations(n=n, seed=seed)
        ranges = (18.0, 22.0, 1.0, 3.0)
        for variant in (0, 1):
            res = ctx.grid_posterior(obs, N, ranges[:2], ranges[2:], posterior="device", variant=variant)
            si = ctx.scale_info()
            p = cge.make_params(n, N, ranges[:2], ranges[2:], 0, None, variant)
            z = ctx.read_results(p)[3]
            log2z_true = np.log2(z) - n * si["log2s"]
            log2z_cf = closed_form_log2z(obs, N, ranges)
            print(json.dumps(dict(N=N, n=n, variant=variant, path=si["product_path"],
                                  rel=float((log2z_true - log2z_cf) * np.log(2.0)))), flush=True)
