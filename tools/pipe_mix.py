#!/usr/bin/env python
"""GPU: pipe-peak micro-benchmarks incl. the packed/scalar/MUFU mixes (cge_microbench 0..12)."""
import json
import sys

sys.path.insert(0, ".")
import cuda_grid_estimation_b200 as cge  # noqa: E402

NAMES = ["mufu_ex2", "ffma", "dfma", "ffma2", "4ffma2+4ffma", "4ffma2+8ffma", "6ffma2+2mufu",
         "4ffma2+4ffma+2mufu", "6ffma2+2mufu@8warps/SM", "6ffma2+2mufu@12warps/SM",
         "6ffma2+2mufu@16warps/SM", "10ffma2+2mufu", "10ffma2+2mufu@24warps/SM"]
FMA_PER_ITER = [0, 8, 8, 16, 12, 16, 14, 14, 14, 14, 14, 22, 22]
MUFU_PER_ITER = [8, 0, 0, 0, 0, 0, 2, 2, 2, 2, 2, 2, 2]
with cge.Context(0) as ctx:
    sm = ctx.device_info()["sm_count"]
    out = {}
    for k, name in enumerate(NAMES):
        ops = ctx.microbench(k)
        row = {"Gops": ops / 1e9}
        if k >= 6:  # derive the MUFU rate of the mix from its FMA rate
            row["mufu_Gops"] = ops / FMA_PER_ITER[k] * MUFU_PER_ITER[k] / 1e9
        out[name] = row
    print(json.dumps({"sm_count": sm, **out}))
