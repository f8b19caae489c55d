#!/usr/bin/env python
"""GPU (>= 2 devices): cge_grid_posterior_multi on distinct devices against the single-GPU call
(expectations, both marginals), then its end-to-end time per step with host observations in and
host marginals / expectations out.   usage: python tools/multi_e2e_check.py [n_gpus] [N] [steps]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else min(torch.cuda.device_count(), 8)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
MU, SIGMA = (18.0, 22.0), (1.0, 3.0)
obs = observations(50)
out = {"n_gpus": G, "devices_on_box": torch.cuda.device_count()}
with cge.Context(0) as ctx:
    for n_check in (4100, 1031):     # shard edges on no group boundary
        one = ctx.grid_posterior(obs, n_check, MU, SIGMA)
        with cge.MultiContext(list(range(G))) as mc:
            for _ in range(3):       # several steps: both buffers of the exchange
                many = mc.grid_posterior(obs, n_check, MU, SIGMA, posterior=None)
        d_e = float(np.abs(one.expectations - many.expectations).max())
        d_mu = float(np.abs(one.marg_mu - many.marg_mu).max() / one.marg_mu.max())
        d_sg = float(np.abs(one.marg_sigma - many.marg_sigma).max() / one.marg_sigma.max())
        out[f"check_{n_check}"] = {"dE": d_e, "d_marg_mu_rel": d_mu, "d_marg_sigma_rel": d_sg}
        assert d_e <= 1e-9 and d_mu <= 1e-9 and d_sg <= 1e-9, out
with cge.MultiContext(list(range(G))) as mc:
    for _ in range(5):
        mc.grid_posterior(obs, N, MU, SIGMA, posterior=None)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = mc.grid_posterior(obs, N, MU, SIGMA, posterior=None)
    secs = time.perf_counter() - t0
out.update({"N": N, "steps": steps, "e2e_ms_per_step": secs / steps * 1e3,
            "e2e_G_pdf_evals_per_s": float(N) * N * 50 * steps / secs / 1e9,
            "E": r.expectations.tolist()})
print(json.dumps(out))
