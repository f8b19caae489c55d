#!/usr/bin/env python
"""Where a likelihood-kernel launch spends its warp time, by source function: joins the source
page of an .ncu-rep (`ncu --set full --import-source on`, read here without a GPU) with the line
table of the built object (nvdisasm -g) and sums the stall samples and executed instructions per
enclosing device function; the inner per-observation loop (tools/sass_inner_loop.py) is its own
row.   usage: python tools/ncu_regions.py gpurun_out/x.ncu-rep profiles/x_regions.csv [top.csv]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import sass_inner_loop  # noqa: E402

CSRC = os.path.join(ROOT, "cuda_grid_estimation_b200", "csrc")
OBJ = os.path.join(ROOT, "cuda_grid_estimation_b200", "_obj", "cge_launch_product.o")
DEFAULT = "Lb1ELi1ELb0ELi16ELb1E"   # the default product kernel's template tail


def line_table():
    """offset -> (file, line) of the default product kernel."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, check=True,
                              capture_output=True, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(text) if l.startswith(".text.") and DEFAULT in l][0]
    table, cur = {}, None
    for l in text[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def enclosing(cache, fname, line):
    if fname not in cache:
        names, path = [], os.path.join(CSRC, fname)
        if os.path.exists(path):
            for i, l in enumerate(open(path).read().splitlines(), 1):
                m = re.match(r"\s*(?:template\s*<.*>\s*)?(?:static\s+)?(?:__device__|__global__).*?\b(\w+)\s*\(", l)
                if m:
                    names.append((i, m.group(1)))
        cache[fname] = names
    best = "(toolkit header)" if not cache[fname] else "?"
    for i, n in cache[fname]:
        if i > line:
            break
        best = n
    return best


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    assert "(bool)1, (int)1, (bool)0, (int)16, (bool)1" in kernel, kernel
    ix = {h: i for i, h in enumerate(hdr)}

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    loops = sass_inner_loop.inner_loops()
    spans = {k: (b[0][0], b[-1][0]) for k, (_, b) in loops.items() if k in ("tall", "wide")}
    table, cache = line_table(), {}
    base = int(data[0][ix["Address"]], 16)
    samples, insts = collections.Counter(), collections.Counter()
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        key = None
        for name, (lo, hi) in spans.items():
            if lo <= off <= hi:
                key = f"[{name} inner loop, 4 observations per trip]"
        if key is None:
            fl = table.get(off)
            key = f"{fl[0]}:{enclosing(cache, *fl)}" if fl else "?"
        samples[key] += num(r[ix["# Samples"]])
        insts[key] += num(r[ix["Instructions Executed"]])
    ts, ti = sum(samples.values()), sum(insts.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["region (enclosing device function of the SASS line table)", "stall samples",
                    "% of samples", "warp instructions executed", "% of instructions"])
        for k, v in samples.most_common():
            w.writerow([k, int(v), f"{100 * v / ts:.2f}", int(insts[k]), f"{100 * insts[k] / ti:.2f}"])
        w.writerow(["total", int(ts), "100.00", int(ti), "100.00"])
    print("wrote", out)
    if len(sys.argv) > 3:   # the instructions with the most samples, with their stall reasons
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        top = sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:40]
        with open(sys.argv[3], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["offset", "SASS", "# samples", "instructions executed", "top stall reasons"])
            for r in top:
                why = sorted(((num(r[ix[s]]), s) for s in stalls), reverse=True)[:3]
                w.writerow([f"{int(r[ix['Address']], 16) - base:#x}", r[ix["Source"]].strip(),
                            r[ix["# Samples"]], r[ix["Instructions Executed"]],
                            " ".join(f"{s}={int(v)}" for v, s in why if v > 0)])
        print("wrote", sys.argv[3])


if __name__ == "__main__":
    main()
