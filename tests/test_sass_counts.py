"""CPU (-m "not gpu"): the instruction counts bench.py's roofline block quotes are the ones the
shipped library contains.  cuobjdump disassembles libcge.so here (no GPU needed); the loops are
found by tools/sass_inner_loop.py, which also writes profiles/r02_sass_inner_loop_*.txt."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def found():
    from cuda_grid_estimation_b200 import _build
    _build.build()
    from tools import sass_inner_loop
    return {k: sass_inner_loop.describe(body) + (len(body),)
            for k, (_, body) in sass_inner_loop.inner_loops().items()}


def test_inner_loops_match_the_counts_bench_quotes(found):
    import bench
    assert set(found) == {"tall", "wide", "logspace"}
    # one trip = 4 observations; per thread 7 x 8 (tall) or 6 x 8 (wide) cells
    for key, path, cells in (("tall", "slope8x7", 56), ("wide", "slope8", 48)):
        _, mufu, packed, _ = found[key]
        mufu_pe, fma_pe = bench.PIPE_OPS[path]
        assert mufu / (4 * cells) == pytest.approx(mufu_pe, rel=1e-12)
        assert 2 * packed / (4 * cells) == pytest.approx(fma_pe, rel=1e-12)
    # log-space: 4 rows x 8 columns per thread, 1.125 FP32 lane-ops per pdf-eval, nothing else
    # in the loop but the observation load and the loop control
    ops, mufu, packed, length = found["logspace"]
    assert mufu == 0 and 2 * packed / (4 * 32) == pytest.approx(1.125)
    assert length - packed <= 4 and ops["LDS.128"] == 1


def test_loops_are_all_packed(found):
    # everything the product loops issue besides packed f32x2 and MUFU.EX2: one observation load,
    # the loop control and a few moves -- no scalar FFMA/FMUL left in them
    for key in ("tall", "wide"):
        ops, mufu, packed, length = found[key]
        assert length - packed - mufu <= 12, (key, dict(ops))
        assert not any(op.split(".")[0] in ("FFMA", "FMUL", "FADD") for op in ops), dict(ops)


def test_no_register_spills_inside_the_loops(found):
    # the 128-register kernels spill a few values (profiles/r02_ptxas_v.txt); none of it inside
    # the per-observation loops
    for key, (ops, _, _, _) in found.items():
        assert not any(op.split(".")[0] in ("LDL", "STL") for op in ops), (key, dict(ops))
