"""Extracts the inner (per-observation) loops of the likelihood kernels from the built library with
cuobjdump and writes them, with an instruction histogram, to profiles/ -- the SASS counts that
bench.py's PIPE_OPS table and DESIGN.md section 3.2 quote.  Runs without a GPU.

    python tools/sass_inner_loop.py            # writes profiles/r02_sass_inner_loop_*.txt

A loop is recognised as the body of a backward branch; the ones kept are identified by what they
issue per trip (MUFU.EX2 and packed f32x2 instructions per 4 observations):
    tall  (ProductSlopeTall, 7 rows x 8 columns): 32 MUFU.EX2, 268 packed
    wide  (ProductSlopeWide, 6 rows x 8 columns): 64 MUFU.EX2, 248 packed
    log-space (LogSpaceTiled<256>): no MUFU, the longest all-FFMA2 body
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda_grid_estimation_b200", "libcge.so")
OUT = os.path.join(ROOT, "profiles")
PACKED = re.compile(r"\b(FFMA2|FMUL2|FADD2)\b")
INS = re.compile(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def functions(lib=LIB):
    """{mangled name: [(address, text)]} of every kernel in the library."""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
            continue
        m = INS.match(line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def loops(ins):
    """Bodies of the backward branches of one function: (first index, last index)."""
    index = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+.*?0x([0-9a-f]+)", text)
        if m:
            target = int(m.group(1), 16)
            if target < a and target in index:
                yield index[target], k


def describe(body):
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in body)
    mufu = sum("MUFU.EX2" in t for _, t in body)
    packed = sum(bool(PACKED.search(t)) for _, t in body)
    return ops, mufu, packed


def write(name, title, fn, body):
    ops, mufu, packed = describe(body)
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        f.write(f"# {title}\n# kernel: {fn}\n")
        f.write(f"# loop body {body[0][0]:#x} .. {body[-1][0]:#x}: {len(body)} instructions per trip "
                f"(4 observations), {mufu} MUFU.EX2, {packed} packed f32x2 (FFMA2/FMUL2/FADD2)\n")
        f.write("# histogram: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
        f.write("# (cuobjdump -sass of cuda_grid_estimation_b200/libcge.so; tools/sass_inner_loop.py)\n")
        for a, t in body:
            f.write(f"/*{a:05x}*/  {t} ;\n")
    print(f"{name}: {len(body)} instructions, {mufu} MUFU.EX2, {packed} packed", file=sys.stderr)


def inner_loops(fns=None):
    """{"tall" | "wide" | "logspace": (kernel name, loop body)} of the default kernels."""
    fns = fns or functions()
    default = [k for k in fns if "grid_likelihood_kernel" in k and "Lb1ELi1ELb0ELi16ELb1E" in k]
    assert len(default) == 1, default   # SELECT, one 16-warp CTA per SM, WIDE
    ins = fns[default[0]]
    found = {}
    for b, e in loops(ins):
        body = ins[b:e + 1]
        _, mufu, packed = describe(body)
        if (mufu, packed) == (32, 268):
            found["tall"] = (default[0], body)
        elif (mufu, packed) == (64, 248):
            found["wide"] = (default[0], body)
    tiled = [k for k in fns if "LogSpaceTiled" in k and "Lb0ELi16E" in k]  # not streamed, 16 warps
    assert len(tiled) == 1, tiled
    ins = fns[tiled[0]]
    best = None
    for b, e in loops(ins):
        body = ins[b:e + 1]
        _, mufu, packed = describe(body)
        if mufu == 0 and packed >= 0.85 * len(body) and (best is None or len(body) > len(best)):
            best = body
    if best is not None:
        found["logspace"] = (tiled[0], best)
    return found


def main():
    found = inner_loops()
    assert set(found) == {"tall", "wide", "logspace"}, sorted(found)
    write("r02_sass_inner_loop_tall_slope8x7.txt",
          "ProductSlopeTall: 7 rows x 8 adjacent columns per thread, 4 observations per trip "
          "(56 evals per observation: 8 MUFU.EX2 + 67 packed)", *found["tall"])
    write("r02_sass_inner_loop_wide_slope8.txt",
          "ProductSlopeWide: 6 rows x 8 adjacent columns per thread, 4 observations per trip "
          "(48 evals per observation: 16 MUFU.EX2 + 62 packed)", *found["wide"])

    write("r02_sass_inner_loop_logspace_tiled.txt",
          "LogSpaceTiled<256>: 4 rows x 8 columns per thread, sums of centred terms as packed f32x2",
          *found["logspace"])


if __name__ == "__main__":
    main()
