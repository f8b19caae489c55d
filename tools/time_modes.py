#!/usr/bin/env python
"""GPU: per-phase times for (N, n_obs, mode, variant) tuples given as N:n:mode:variant args."""
import json
import sys

import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

dev = torch.device("cuda", 0)
with cge.Context(0) as ctx:
    for spec in sys.argv[1:]:
        N, n, mode, variant = (int(v) for v in spec.split(":"))
        steps = 5 if n > 1000 else 20
        obs_t = torch.from_numpy(observations(n)).to(dev)
        post = torch.empty((N, N), dtype=torch.float32, device=dev)
        p = cge.make_params(n, N, (18, 22), (1, 3), mode, None, variant)
        for _ in range(2):
            ctx.likelihood_partials(p, obs_t, post, None)
            ctx.finalize(p, post, None)
        torch.cuda.synchronize()
        ctx.profile_begin(steps)
        for _ in range(steps):
            ctx.likelihood_partials(p, obs_t, post, None)
            ctx.finalize(p, post, None)
        t, cap = ctx.profile_end()
        ex = ctx.read_results(p)[0]
        print(json.dumps(dict(spec=spec, T_evals_s=N * N * n / t["likelihood_ms"] / 1e9, E=ex.tolist(),
                              **{k: round(v, 4) for k, v in t.items()})), flush=True)
        del post
