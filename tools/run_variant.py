#!/usr/bin/env python
"""GPU: run a few steps of one likelihood-kernel variant (for ncu).
usage: python tools/run_variant.py N n_obs variant steps [mode]"""
import sys

import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

N, n, variant, steps = (int(a) for a in sys.argv[1:5])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
obs = observations(n)
dev = torch.device("cuda", 0)
with cge.Context(0) as ctx:
    obs_t = torch.from_numpy(obs).to(dev)
    post = torch.empty((N, N), dtype=torch.float32, device=dev)
    p = cge.make_params(n, N, (18, 22), (1, 3), mode, None, variant)
    for _ in range(steps):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    torch.cuda.synchronize()
    print(ctx.read_results(p)[0])
