#!/usr/bin/env python
"""GPU, diagnostic build (CGE_BUILD_DEFINES=-DCGE_DEBUG_TIMES): start/end time and SM of every CTA
of the persistent likelihood kernel.   usage: python tools/cta_times.py N n_obs [mode]"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402
from tools._obs import observations  # noqa: E402

N, n = int(sys.argv[1]), int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = int(sys.argv[4]) if len(sys.argv) > 4 else N
dev = torch.device("cuda", 0)
with cge.Context(0) as ctx:
    obs_t = torch.from_numpy(observations(n)).to(dev)
    post = torch.empty((rows, N), dtype=torch.float32, device=dev)
    rb = (N - rows) // 2 // 4 * 4
    p = cge.make_params(n, N, (18, 22), (1, 3), mode, (rb, rb + rows))
    for _ in range(3):
        ctx.likelihood_partials(p, obs_t, post, None)
        ctx.finalize(p, post, None)
    torch.cuda.synchronize()
    buf = np.zeros(36 * 4096, np.uint64)
    cnt = C.c_int()
    lib = cge.load_library()
    lib.cge_debug_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    assert lib.cge_debug_times(ctx._h, buf.ctypes.data, 4096, C.byref(cnt)) == 0
    t = buf[:4 * cnt.value].reshape(-1, 4).astype(np.int64)
    t0 = t[:, 0].min()
    wend = buf[4 * cnt.value:36 * cnt.value].reshape(-1, 32).astype(np.int64)
    wend = wend[:, (wend != 0).any(axis=0)] - t0   # (32 slots per CTA; a CTA has 8, 16 or 20 warps)
    kend = wend.max()
    print(json.dumps({"warp_end_us": {"min": float(wend.min()) / 1e3, "p10": float(np.percentile(wend, 10)) / 1e3,
                                      "median": float(np.median(wend)) / 1e3, "p90": float(np.percentile(wend, 90)) / 1e3,
                                      "max": float(kend) / 1e3},
                      "mean_warp_residency": float(wend.mean() / kend),
                      "spread_within_cta_us": float(np.median(wend.max(axis=1) - wend.min(axis=1))) / 1e3}))
    start, end, sm, setup = t[:, 0] - t0, t[:, 1] - t0, t[:, 2], t[:, 3] - t[:, 0]
    dur = end - start
    per_sm = {}
    for s in sm:
        per_sm[int(s)] = per_sm.get(int(s), 0) + 1
    print(json.dumps({
        "ctas": int(cnt.value), "kernel_us": float(end.max()) / 1e3,
        "start_us": {"min": float(start.min()) / 1e3, "median": float(np.median(start)) / 1e3, "max": float(start.max()) / 1e3},
        "end_us": {"min": float(end.min()) / 1e3, "median": float(np.median(end)) / 1e3, "max": float(end.max()) / 1e3,
                   "p10": float(np.percentile(end, 10)) / 1e3, "p90": float(np.percentile(end, 90)) / 1e3},
        "dur_us": {"min": float(dur.min()) / 1e3, "median": float(np.median(dur)) / 1e3, "max": float(dur.max()) / 1e3},
        "setup_us": {"min": float(setup.min()) / 1e3, "median": float(np.median(setup)) / 1e3, "max": float(setup.max()) / 1e3},
        "ctas_per_sm_hist": {str(k): int(v) for k, v in zip(*np.unique(list(per_sm.values()), return_counts=True))},
        "sms_used": len(per_sm)}))
    order = np.argsort(start)
    print("first/last 8 by start:", [(int(i), int(sm[i]), round(start[i] / 1e3, 1), round(end[i] / 1e3, 1)) for i in list(order[:8]) + list(order[-8:])])
