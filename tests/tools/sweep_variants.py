#!/usr/bin/env python
"""GPU: time the likelihood kernel of every product-path variant (CUDA events inside libcge's
per-step capture) and report accuracy against the oracle on the README workload.
usage: python tests/tools/sweep_variants.py [N] [n_obs] [steps]
(lives under tests/ because it checks accuracy against the oracle, which is test infrastructure)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
NACC = int(os.environ.get("CGE_ACC_N", "4096"))  # accuracy is checked on the README workload at this N
obs = oracle.readme_observations(n=n)
ref = oracle.grid_posterior(oracle.readme_observations(), NACC, 18, 22, 1, 3, oracle.NUM_F64)
dev = torch.device("cuda", 0)
out = []
with cge.Context(0) as ctx:
    obs_t = torch.from_numpy(obs).to(dev)
    post = torch.empty((N, N), dtype=torch.float32, device=dev)
    for variant in [int(v) for v in os.environ.get("CGE_VARIANTS", "1,4,8,3,9,10,11,0").split(",")]:
        p = cge.make_params(n, N, (18, 22), (1, 3), cge.MODE_PRODUCT_F32, None, variant)
        for _ in range(3):
            ctx.likelihood_partials(p, obs_t, post, None)
            ctx.finalize(p, post, None)
        torch.cuda.synchronize()
        ctx.profile_begin(steps)
        for _ in range(steps):
            ctx.likelihood_partials(p, obs_t, post, None)
            ctx.finalize(p, post, None)
        t, cap = ctx.profile_end()
        r = ctx.grid_posterior(oracle.readme_observations(), NACC, (18, 22), (1, 3), variant=variant)
        big = ref["posterior"] >= 1e-20 * ref["posterior"].max()
        rel = np.abs(r.posterior.astype(np.float64)[big] - ref["posterior"][big]) / ref["posterior"][big]
        row = dict(variant=variant, likelihood_ms=t["likelihood_ms"], total_ms=t["total_ms"],
                   reduce_ms=t["reduce_ms"], scale_ms=t["scale_ms"], setup_ms=t["setup_ms"],
                   T_evals_per_s=N * N * n / t["likelihood_ms"] / 1e9,
                   dE=float(np.abs(r.expectations - ref["expectations"]).max()),
                   max_rel=float(rel.max()), rms_rel=float(np.sqrt((rel ** 2).mean())))
        out.append(row)
        print(json.dumps(row), flush=True)
