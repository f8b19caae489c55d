"""CPU: the C-ABI library builds, loads, exports every symbol include/cge.h declares, and the
product path fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re
import subprocess

import pytest

import cuda_grid_estimation_b200 as cge
from cuda_grid_estimation_b200 import _build, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cge.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cge_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = _build.build()
    assert os.path.exists(path)
    lib = cge.load_library()
    assert lib.cge_version() == 2


def test_every_declared_symbol_is_exported():
    lib = cge.load_library()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cge.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared, "api.ABI_SYMBOLS out of sync with include/cge.h"


def test_exports_are_plain_c_and_sm100a_only():
    out = subprocess.run(["nm", "-D", "--defined-only", _build.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(_header_symbols()) <= exported
    # nothing mangled leaks out of the boundary
    assert not [s for s in exported if s.startswith("_Z")], "C++ symbols exported"
    sass = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], capture_output=True, text=True)
    if sass.returncode == 0:
        archs = set(re.findall(r"sm_\d+a?", sass.stdout))
        assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    assert C.sizeof(api.Params) == 44
    assert C.sizeof(api.Timing) == 28


def test_no_cpu_fallback_without_device():
    """In the build container there is no GPU: creating a context must fail with a clear error,
    not silently compute on the host.  (On a GPU box this test checks a bad device index.)"""
    import torch
    if torch.cuda.is_available():
        with pytest.raises(cge.CgeError) as ei:
            cge.Context(10_000)
        assert ei.value.code == api.ERR_BAD_ARG
    else:
        with pytest.raises(cge.CgeError) as ei:
            cge.Context(0)
        assert ei.value.code in (api.ERR_NO_DEVICE, api.ERR_CUDA)
        assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "cuda_grid_estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
    # the timing tools neither (the ones that check accuracy against the oracle live in tests/tools/)
    tools = os.path.join(ROOT, "tools")
    for f in os.listdir(tools):
        if f.endswith(".py"):
            src = open(os.path.join(tools, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
    # bench.py: only the CPU legs (cpu_rate_sample, run_reference, and oracle_all_cores, the helper
    # those two share and nothing else calls) may import it, and only lazily
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert not re.search(r"^(import|from)\s+oracle\b", bench, flags=re.M)
    for m in re.finditer(r"^def (\w+)\(.*?(?=^def |\Z)", bench, flags=re.M | re.S):
        if re.search(r"^\s+import oracle\b", m.group(0), flags=re.M):
            assert m.group(1) in ("cpu_rate_sample", "run_reference", "oracle_all_cores"), m.group(1)
        if "oracle_all_cores()" in m.group(0):
            assert m.group(1) in ("cpu_rate_sample", "run_reference", "oracle_all_cores"), m.group(1)


def test_policy_mirror_matches_the_documented_bounds():
    """api.slope_error_bound / api._product_path / api.band_paths restate the device-side choice
    of the product policy (cge::ProductSlopeT::eligible, likelihood_cta); the numbers DESIGN.md
    quotes for the README ranges follow from them (no GPU needed)."""
    import numpy as np

    from cuda_grid_estimation_b200 import api
    info = dict(paired_umax=0.011, paired3_umax=0.05, slope_err_max=2.0 ** -24, wide=0.0)

    def readme(N):   # umax and relative sigma step of the README ranges (mu 18..22, sigma 1..3)
        return 0.008386777734265279 * 4096 / N, (2.0 / (N - 1)) / 1.0

    u, r = readme(4096)
    assert abs(api.slope_error_bound(u, r, 1.5) - 2.503e-8) < 1e-10      # 4 columns
    assert abs(api.slope_error_bound(u, r, 3.5) - 5.850e-8) < 1e-10      # 8 columns: just inside 2^-24
    assert api.slope_error_bound(u, r, 3.5) <= info["slope_err_max"]
    u, r = readme(3300)
    assert api._product_path(info, u, r) == "slope"
    assert api._product_path(info, u, (3.0 / 3299) / 1.0) == "paired"    # sigma in [1, 4]: columns too far apart
    assert api._product_path(info, *readme(1024)) == "paired3"
    assert api._product_path(info, *readme(512)) == "hybrid"
    # the tall policy's bound: 3 umax <= 0.011 and 9 x the 8-column slope term <= 2^-24
    u, r = readme(9600)
    assert 3 * u <= 0.011 and 9 * api.slope_error_bound(u, r, 3.5) <= 2.0 ** -24
    u, r = readme(9000)
    assert 3 * u > 0.011
    # per-band choice on a wide sigma range: costlier arithmetic only where sigma is small
    sigma = np.linspace(0.4, 3.0, 4096).astype(np.float32)
    paths = api.band_paths(dict(info, ustep=u * 9000 / 4096 / 0.7213475204444817), sigma)
    assert len(paths) == 32 and paths[0] != "slope" and paths[-1] == "slope"
    assert api.band_paths(dict(info, wide=2.0, ustep=0.0), sigma) == ["slope8x7"] * 32


def test_argument_errors_need_no_device():
    """The error behaviour of the boundary (include/cge.h: int status + cge_last_error(), never
    exit, never print) for the mistakes that are caught before any CUDA call -- so it can be checked
    in the build container too."""
    import threading

    lib = cge.load_library()
    null = C.c_void_p()
    params = api.Params()
    assert lib.cge_destroy(None) == api.OK
    assert lib.cge_destroy_multi(None) == api.OK
    assert lib.cge_create(0, None) == api.ERR_BAD_ARG
    assert b"NULL" in lib.cge_last_error()
    out = C.c_void_p()
    for n_gpus in (0, -1, 1000):
        assert lib.cge_create_multi(n_gpus, None, C.byref(out)) == api.ERR_BAD_ARG
        assert b"n_gpus" in lib.cge_last_error() and not out.value
    assert lib.cge_grid_posterior(None, C.byref(params), None, None, None, None, None, None) == api.ERR_BAD_ARG
    assert b"context is NULL" in lib.cge_last_error()
    assert lib.cge_grid_posterior_multi(None, C.byref(params), None, None, None, None, None, None) == api.ERR_BAD_ARG
    assert lib.cge_multi_size(None, None) == api.ERR_BAD_ARG
    assert lib.cge_results(None, C.byref(null), None) == api.ERR_BAD_ARG
    # the message belongs to the calling thread (cge.h: one context per host thread)
    seen = {}

    def other():
        seen["before"] = lib.cge_last_error()
        lib.cge_multi_size(None, None)
        seen["after"] = lib.cge_last_error()

    lib.cge_create(0, None)
    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert seen["before"] == b"" and seen["after"] != b""
    assert b"out is NULL" in lib.cge_last_error()
