"""In-tree build of libcge.so (sm_100a only).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcge.so")
# one object per translation unit, compiled concurrently: the likelihood-kernel instantiations
# (cge_launch_*.cu) dominate the build time
SOURCES = ["cge_api.cu", "cge_extras.cu", "cge_multi.cu", "cge_launch_product.cu",
           "cge_launch_logspace.cu", "cge_launch_small.cu"]
HEADERS = ["cge_kernels.cuh", "cge_kernels_misc.cuh", "cge_device.cuh", "cge_host.cuh",
           os.path.join("..", "..", "include", "cge.h")]
OBJ_DIR = os.path.join(PKG_DIR, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the SASS + line tables of ~40 kernel instantiations: 10.2 MB uncompressed, 4.0 MB compressed
    # (the driver inflates a kernel's image when the module is first loaded, not per launch)
    "-Xfatbin=-compress-all",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
] + os.environ.get("CGE_BUILD_DEFINES", "").split()   # e.g. -DCGE_DEBUG_TIMES (diagnostic builds)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcge.so cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile csrc/*.cu into cuda_grid_estimation_b200/libcge.so; returns its path.
    `out` / `defines`: a differently configured copy for A/B tuning runs (CGE_LIBRARY selects it)."""
    if out is None and not force and not stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = OBJ_DIR if out is None else OBJ_DIR + "_" + os.path.basename(out)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        done = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in done:
            print(log)
    target = LIB_PATH if out is None else out
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + \
           [obj for obj, _ in done]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return target


# ---- host-side programs on top of the C ABI (C++/CUDA, like the reference's own binaries) ----
BIN_DIR = os.path.join(PKG_DIR, "bin")
ROOT = os.path.dirname(PKG_DIR)
HOST_PROGRAMS = {
    # the reference's main() equivalent on the fused path
    "grid": os.path.join(PKG_DIR, "host", "grid_main.cu"),
    # the reference's unit tests restated against include/cge_compat.h
    "test_compat": os.path.join(ROOT, "tests", "cpp", "test_compat.cu"),
}


def build_host_programs(force: bool = False) -> dict:
    """nvcc the host programs against libcge.so (rpath = the package directory)."""
    build()
    os.makedirs(BIN_DIR, exist_ok=True)
    out = {}
    for name, src in HOST_PROGRAMS.items():
        exe = os.path.join(BIN_DIR, name)
        out[name] = exe
        deps = [src, LIB_PATH, os.path.join(ROOT, "include", "cge.h"),
                os.path.join(ROOT, "include", "cge_compat.h")]
        if not force and os.path.exists(exe) and all(
                os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
            continue
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-w",
               "-o", exe, src, "-L" + PKG_DIR, "-lcge", "-Xlinker", "-rpath," + PKG_DIR,
               "-Xlinker", "-rpath,$ORIGIN/.."]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n" + res.stdout + res.stderr)
    return out
