"""CPU: pin the oracle against outputs of the reference CUDA pipeline ITSELF.  The fixtures
tests/golden/ref_cuda_*.npz were produced on a B200 by oracle/_ref/ref_harness (the unmodified
reference headers driven in src/grid.cu's call order; see tests/golden/make_ref_cuda_golden.py).

The oracle's REF numerics differ from the reference only in whose expf/powf rounds the f32
densities (glibc vs CUDA libm, <= ~1 ulp per factor), so cells agree to a few 1e-6 relative and
expectations to ~1e-9; ranges must agree bit for bit."""
import os

import numpy as np
import pytest

import oracle

CASES = ["readme", "small", "max256"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_cuda_run(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"ref_cuda_{name}.npz"))
    N, obs = int(g["N"]), g["observations"]
    rng = [float(x) for x in g["ranges"]]
    ref = oracle.grid_posterior(obs, N, *rng, oracle.NUM_REF)
    # a1: bit-exact ranges
    assert np.array_equal(ref["mu"], g["mu"]) and np.array_equal(ref["sigma"], g["sigma"])
    # a12/a13: expectations
    assert np.abs(ref["expectations"] - g["expectations"]).max() < 2e-8
    # a11: marginals
    assert np.abs(ref["marg_mu"] - g["marg_mu"]).max() < 5e-8
    assert np.abs(ref["marg_sigma"] - g["marg_sigma"]).max() < 5e-8
    pmax = ref["posterior"].max()
    if "posterior" in g:
        P = g["posterior"].reshape(N, N)
        big = P >= 1e-20 * pmax
        assert (np.abs(ref["posterior"][big] - P[big]) / P[big]).max() < 1e-5
        assert abs(P.sum() - 1) < 1e-12
        # a8/a9: unnormalised likelihoods (f64 tree product over f32 densities)
        L = oracle.grid_likes(obs, N, *rng, oracle.NUM_REF)
        Lg = g["likes"].reshape(N, N)
        assert (np.abs(L[big] - Lg[big]) / Lg[big]).max() < 1e-5
    else:
        rows, P = g["sample_rows"], g["posterior_rows"]
        big = P >= 1e-20 * pmax
        assert (np.abs(ref["posterior"][rows][big] - P[big]) / P[big]).max() < 1e-5


def test_reference_cuda_readme_values(golden_dir):
    """What the reference binary itself computes for the README observations on a B200."""
    g = np.load(os.path.join(golden_dir, "ref_cuda_readme.npz"))
    assert bool(g["used_wrapper"])  # went through computeDensitiesWrapper (wrappers.h:19)
    assert g["expectations"][0] == pytest.approx(19.549052474386311, abs=1e-12)
    assert g["expectations"][1] == pytest.approx(1.9167625055600417, abs=1e-12)
    # SURVEY's emulation predicted 19.549052474669377, 1.916762507106191
    assert g["expectations"] == pytest.approx([19.549052474669377, 1.916762507106191], abs=5e-9)
