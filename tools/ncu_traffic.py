#!/usr/bin/env python
"""Record the measured DRAM traffic of the likelihood kernel for bench.py's roofline.traffic:
reads dram__bytes_read.sum + dram__bytes_write.sum of the first grid_likelihood launch in an
.ncu-rep (on the build container, no GPU needed) and stores it under "<workload>|<n_gpus>" in
profiles/r02_dram_traffic.json.
usage: python tools/ncu_traffic.py gpurun_out/x.ncu-rep 16384x50 1"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, workload, world = sys.argv[1], sys.argv[2], int(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if "grid_likelihood" not in r[idx["Kernel Name"]]:
            continue
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[idx[m]]) * SCALE[units[idx[m]]]
        table = json.load(open(OUT)) if os.path.exists(OUT) else {}
        table[f"{workload}|{world}"] = {
            "bytes_per_launch": total,
            "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) *
            {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[idx["gpu__time_duration.sum"]]],
            "kernel": r[idx["Kernel Name"]][:160],
            "source": "ncu --set full, one launch: " + os.path.basename(rep),
        }
        json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
        print(f"{workload}|{world}: {total / 1e9:.4f} GB per launch")
        return
    raise SystemExit("no grid_likelihood launch in " + rep)


if __name__ == "__main__":
    main()
