"""CPU (-m "not gpu"): size-independent properties of the oracle -- the checker the GPU parity
tests rely on -- on random small problems (hypothesis).  They are the same properties the GPU
tests use at the sizes the oracle cannot reach (tests/test_gpu_parity.py: 16384^2, 65536^2):
normalisation, marginals = sums of the posterior, row shards = rows of the whole grid, invariance
under a common shift, and the three numerics (reference f32 pdf / f64 / log-space) agreeing.
Reference semantics: inc/kernels.h:102-112 (pdf), inc/cuda_functions.h:139-159 (normalise),
inc/wrappers.h:169-170 (marginal axes), inc/cuda_functions.h:201-218 (expectations)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle

RANGES = (18.0, 22.0, 1.0, 3.0)
problem = st.tuples(st.integers(min_value=2, max_value=48),          # grid_size
                    st.integers(min_value=1, max_value=40),          # n_obs
                    st.integers(min_value=0, max_value=2 ** 31 - 1))  # seed


def _obs(n, seed):
    return np.random.default_rng(seed).normal(20.0, 2.0, n).astype(np.float32)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(problem)
def test_posterior_is_normalised_and_marginals_are_its_sums(p):
    N, n, seed = p
    r = oracle.grid_posterior(_obs(n, seed), N, *RANGES, oracle.NUM_REF)
    post = r["posterior"].reshape(N, N).astype(np.float64)
    assert abs(post.sum() - 1.0) <= 1e-12
    # [mu][sigma] layout: the mu marginal sums over sigma (axis 1), the sigma marginal over mu
    np.testing.assert_allclose(r["marg_mu"], post.sum(axis=1), rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(r["marg_sigma"], post.sum(axis=0), rtol=1e-12, atol=1e-300)
    mu = oracle.linspace(N, RANGES[0], RANGES[1]).astype(np.float64)
    sg = oracle.linspace(N, RANGES[2], RANGES[3]).astype(np.float64)
    assert abs(r["expectations"][0] - (r["marg_mu"] * mu).sum()) <= 1e-12
    assert abs(r["expectations"][1] - (r["marg_sigma"] * sg).sum()) <= 1e-12
    assert RANGES[0] - 1e-12 <= r["expectations"][0] <= RANGES[1] + 1e-12
    assert RANGES[2] - 1e-12 <= r["expectations"][1] <= RANGES[3] + 1e-12


@settings(max_examples=40, deadline=None, derandomize=True)
@given(problem, st.integers(min_value=1, max_value=5))
def test_row_shards_are_rows_of_the_whole_grid(p, shards):
    """What the multi-GPU path relies on: a shard's unnormalised cells depend on the global row
    index only (cuda_grid_estimation_b200/sharding.py:shard_rows)."""
    from cuda_grid_estimation_b200.sharding import shard_rows
    N, n, seed = p
    shards = min(shards, N)
    obs = _obs(n, seed)
    whole = oracle.grid_likes(obs, N, *RANGES, oracle.NUM_REF).reshape(N, N)
    edges = []
    for r in range(shards):
        rb, re = shard_rows(N, shards, r)
        edges.append((rb, re))
        part = oracle.grid_likes(obs, N, *RANGES, oracle.NUM_REF, rb, re).reshape(re - rb, N)
        assert np.array_equal(part, whole[rb:re])
    assert edges[0][0] == 0 and edges[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(problem)
def test_the_three_numerics_agree(p):
    N, n, seed = p
    obs = _obs(n, seed)
    ref = oracle.grid_posterior(obs, N, *RANGES, oracle.NUM_REF)
    f64 = oracle.grid_posterior(obs, N, *RANGES, oracle.NUM_F64)
    log = oracle.grid_posterior(obs, N, *RANGES, oracle.NUM_LOG)
    # f32 pdf (reference numerics) against f64: the tolerance of the GPU parity tests
    assert np.abs(ref["expectations"] - f64["expectations"]).max() <= 1e-5
    assert np.abs(log["expectations"] - f64["expectations"]).max() <= 1e-9
    big = f64["posterior"] >= 1e-20 * f64["posterior"].max()
    rel = np.abs(log["posterior"][big] - f64["posterior"][big]) / f64["posterior"][big]
    assert rel.max() <= 1e-9
    rel = np.abs(ref["posterior"][big] - f64["posterior"][big]) / f64["posterior"][big]
    assert rel.max() <= 1e-4


@settings(max_examples=25, deadline=None, derandomize=True)
@given(problem, st.sampled_from([-16.0, -2.0, 4.0, 32.0]))
def test_common_shift_leaves_the_posterior_alone(p, shift):
    """pdf(x; mu, sigma) depends on x - mu only: shifting the observations and the mu range by the
    same (exactly representable) amount gives the same posterior and E[mu] + shift."""
    N, n, seed = p
    obs = np.round(_obs(n, seed) * 64.0) / 64.0        # exact in f32 before and after the shift
    a = oracle.grid_posterior(obs, N, 18.0, 22.0, 1.0, 3.0, oracle.NUM_F64)
    b = oracle.grid_posterior((obs + np.float32(shift)).astype(np.float32), N, 18.0 + shift,
                              22.0 + shift, 1.0, 3.0, oracle.NUM_F64)
    # (the float32 mu values themselves round differently after the shift -- an ulp of 4e-6 at 54
    # against 2e-6 at 22 -- and a cell feels that as n * |x - mu| * ulp / sigma^2)
    big = a["posterior"] >= 1e-20 * a["posterior"].max()
    np.testing.assert_allclose(b["posterior"][big], a["posterior"][big], rtol=5e-3)
    assert abs(b["expectations"][0] - shift - a["expectations"][0]) <= 1e-4
    assert abs(b["expectations"][1] - a["expectations"][1]) <= 1e-4


def test_observation_order_does_not_matter_in_float64():
    obs = oracle.readme_observations()
    a = oracle.grid_posterior(obs, 101, *RANGES, oracle.NUM_LOG)
    b = oracle.grid_posterior(np.sort(obs), 101, *RANGES, oracle.NUM_LOG)
    np.testing.assert_allclose(b["posterior"], a["posterior"], rtol=1e-9, atol=1e-300)
    assert np.abs(a["expectations"] - b["expectations"]).max() <= 1e-12
