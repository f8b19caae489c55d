#!/usr/bin/env python
"""GPU: phase times of one row shard (emulates one rank of an N-way split on a single GPU).
usage: python tools/time_shard.py N n_obs rows steps [first_row]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

N, n, rows, steps = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda", 0)
with cge.Context(0) as ctx:
    obs_t = torch.from_numpy(observations(n)).to(dev)
    post = torch.empty((rows, N), dtype=torch.float32, device=dev)
    rb = int(sys.argv[5]) if len(sys.argv) > 5 else (N - rows) // 2 // 4 * 4
    p = cge.make_params(n, N, (18, 22), (1, 3), 0, (rb, rb + rows), 0)
    for _ in range(5):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    torch.cuda.synchronize()
    ctx.profile_begin(steps)
    for _ in range(steps):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    t, cap = ctx.profile_end()
    print(json.dumps(dict(first_row=rb, rows=rows,
                          T_evals_s=rows * N * n / t["likelihood_ms"] / 1e9,
                          **{k: round(v, 4) for k, v in t.items()})), flush=True)
