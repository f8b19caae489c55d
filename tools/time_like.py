#!/usr/bin/env python
"""GPU: per-phase times of the step for one workload (CUDA events recorded by the library).
usage: python tools/time_like.py N n_obs [mode] [variant] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

N, n = int(sys.argv[1]), int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
dev = torch.device("cuda", 0)
with cge.Context(0) as ctx:
    obs_t = torch.from_numpy(observations(n)).to(dev)
    post = torch.empty((N, N), dtype=torch.float32, device=dev)
    p = cge.make_params(n, N, (18, 22), (1, 3), mode, None, variant)
    for _ in range(3):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    torch.cuda.synchronize()
    ctx.profile_begin(steps)
    for _ in range(steps):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    ph, cap = ctx.profile_end()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    e1.record()
    torch.cuda.synchronize()
    ex = ctx.read_results(p)[0]
    evals = float(N) * N * n
    print(json.dumps({"N": N, "n": n, "mode": mode, "variant": variant,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("CGE_")},
                      "like_ms": ph["likelihood_ms"], "fold_ms": ph["reduce_ms"],
                      "finish_ms": ph["scale_ms"], "total_ms": ph["total_ms"],
                      "unprofiled_ms_per_step": e0.elapsed_time(e1) / steps,
                      "like_Tevals": evals / ph["likelihood_ms"] / 1e9,
                      "step_Tevals": evals / (e0.elapsed_time(e1) / steps) / 1e9,
                      "E": ex.tolist()}))
