"""GPU (-m gpu): the rows SURVEY section 8(f) lists as "next" around the hot path, each to the same
parity bar -- non-flat prior, the reference's cuRAND observation generator (checked against the
reference binary itself when oracle/_ref/grid is present), credible-interval quantiles."""
import os
import re
import subprocess

import numpy as np
import pytest

import cuda_grid_estimation_b200 as cge
import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- 8(f) row 1: non-flat prior ------------------------------------------------------------
def _posterior_with_prior(obs, N, prior):
    """notebook cell 4: posterior = likes * prior; posterior /= posterior.sum()  (float64)"""
    likes = oracle.grid_likes(obs, N, 18, 22, 1, 3, oracle.NUM_F64)
    post = likes * prior
    post /= post.sum()
    mu = oracle.linspace(N, 18, 22).astype(np.float64)
    sg = oracle.linspace(N, 1, 3).astype(np.float64)
    return dict(posterior=post, marg_mu=post.sum(axis=1), marg_sigma=post.sum(axis=0),
                expectations=np.array([np.sum(mu * post.sum(axis=1)), np.sum(sg * post.sum(axis=0))]))


def _check(res, ref):
    assert np.abs(res.expectations - ref["expectations"]).max() <= 1e-5
    P, Q = res.posterior.astype(np.float64), ref["posterior"]
    big = Q >= 1e-20 * Q.max()
    assert (np.abs(P[big] - Q[big]) / Q[big]).max() <= 1e-4
    assert abs(P.sum() - 1) <= 1e-6
    for got, want in ((res.marg_mu, ref["marg_mu"]), (res.marg_sigma, ref["marg_sigma"])):
        ok = (np.abs(got - want) <= 1e-7 * want.max()) | (np.abs(got - want) <= 1e-4 * np.abs(want))
        assert ok.all()


@pytest.mark.parametrize("N", [101, 1024])
@pytest.mark.parametrize("mode", [cge.MODE_PRODUCT_F32, cge.MODE_LOGSPACE])
def test_non_flat_prior(readme_obs, N, mode):
    import torch
    rng = np.random.default_rng(5)
    mu = oracle.linspace(N, 18, 22).astype(np.float64)
    pm = np.exp(-0.5 * ((mu - 21.0) / 0.5) ** 2).astype(np.float32)      # informative prior on mu
    ps = (1.0 / oracle.linspace(N, 1, 3)).astype(np.float32)              # Jeffreys-like 1/sigma
    pf = rng.uniform(0.5, 1.5, size=(N, N)).astype(np.float32)
    with cge.Context(0) as ctx:
        flat = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        # separable
        ctx.set_prior(pm, ps, N)
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        _check(res, _posterior_with_prior(readme_obs, N, pm.astype(np.float64)[:, None] * ps.astype(np.float64)[None, :]))
        assert abs(res.expectations[0] - flat.expectations[0]) > 0.1   # the prior really moved it
        # full table (device), alone and combined
        pf_dev = torch.from_numpy(pf).cuda()
        ctx.set_prior(None, None, N, pf_dev)
        _check(ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode),
               _posterior_with_prior(readme_obs, N, pf.astype(np.float64)))
        ctx.set_prior(pm, None, N, pf_dev)
        _check(ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode),
               _posterior_with_prior(readme_obs, N, pf.astype(np.float64) * pm.astype(np.float64)[:, None]))
        # back to flat: bit-identical to before
        ctx.set_prior()
        again = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        assert again.posterior.tobytes() == flat.posterior.tobytes()
        # size mismatch is an error
        ctx.set_prior(pm, ps, N)
        with pytest.raises(cge.CgeError):
            ctx.grid_posterior(readme_obs, N + 1, (18, 22), (1, 3), mode=mode)
        ctx.set_prior()


# ---- 8(f) row 2: the reference's observation generator ----------------------------------------
@pytest.mark.parametrize("N", [200, 1500])
def test_sequential_update_equals_batch(readme_obs, N):
    """The prior row used the way Bayes updating uses it: the posterior of the first 20
    observations, left on the device, becomes the full prior table for the other 30; the result
    must be the posterior of all 50 (flat prior).  N=200 goes through the single-launch kernel,
    N=1500 through the three-launch path."""
    import torch
    dev = torch.device("cuda", 0)
    with cge.Context(0) as ctx:
        batch = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
        first = torch.empty((N, N), dtype=torch.float32, device=dev)
        ctx.grid_posterior(readme_obs[:20], N, (18, 22), (1, 3), posterior=first)
        ctx.set_prior(None, None, N, first)
        try:
            seq = ctx.grid_posterior(readme_obs[20:], N, (18, 22), (1, 3))
        finally:
            ctx.set_prior()
    assert np.abs(seq.expectations - batch.expectations).max() <= 1e-6
    P, Q = seq.posterior.astype(np.float64), batch.posterior.astype(np.float64)
    big = Q >= 1e-20 * Q.max()
    assert (np.abs(P[big] - Q[big]) / Q[big]).max() <= 1e-4
    assert abs(P.sum() - 1) <= 1e-6


def test_generate_normal_known_answers(ctx):
    import torch
    out = torch.empty(50, dtype=torch.float32, device="cuda")
    ctx.generate_normal(out, 50, 20.0, 2.0, seed=42)
    obs = out.cpu().numpy()
    # src/grid.cu:41-47: "Due to seed, the first element should be 18.5689"
    assert abs(obs[0] - 18.5689) < 1e-4
    assert abs(obs.mean() - 20.0) < 1.0 and 1.0 < obs.std() < 3.0
    # src/tests.cu:21-46: draw from a zero-filled state, mu=0, sigma=1 -> 0.00459315
    z = torch.empty(3, dtype=torch.float32, device="cuda")
    ctx.generate_normal(z, 3, 0.0, 1.0, zero_state=True)
    assert abs(z[0].item() - 0.00459315) < 1e-5
    # reproducible and seed-dependent
    again = torch.empty_like(out)
    ctx.generate_normal(again, 50, 20.0, 2.0, seed=42)
    assert torch.equal(out, again)
    ctx.generate_normal(again, 50, 20.0, 2.0, seed=43)
    assert not torch.equal(out, again)


def test_against_reference_binary_end_to_end(ctx):
    """The UNMODIFIED reference main() (oracle/_ref/grid, built from /root/reference by
    oracle/Makefile) draws its own observations with cuRAND and prints its expectations with 6
    significant digits; libcge with the same generator must print the same numbers."""
    import torch
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "grid")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/grid not built (needs the reference checkout at build time)")
    out = subprocess.run([ref_bin], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"Inferred mu: ([0-9.eE+-]+); Actual mu: 20\s+Inferred sigma: ([0-9.eE+-]+); Actual sigma: 2",
                  out.stdout)
    assert m, out.stdout
    ref_mu, ref_sigma = float(m.group(1)), float(m.group(2))
    obs = torch.empty(50, dtype=torch.float32, device="cuda")
    ctx.generate_normal(obs, 50, 20.0, 2.0, seed=42)
    torch.cuda.synchronize()
    for mode in (cge.MODE_PRODUCT_F32, cge.MODE_LOGSPACE):
        res = ctx.grid_posterior(obs, 101, (18, 22), (1, 3), mode=mode, posterior=None)
        assert float(f"{res.expectations[0]:.6g}") == ref_mu, (res.expectations, out.stdout)
        assert float(f"{res.expectations[1]:.6g}") == ref_sigma, (res.expectations, out.stdout)
    # and the CLI with --curand prints the reference's lines
    from cuda_grid_estimation_b200 import _build
    cli = subprocess.run([os.path.join(_build.BIN_DIR, "grid"), "--curand"], capture_output=True, text=True)
    assert cli.returncode == 0, cli.stdout + cli.stderr
    assert "observations[0] = 18.5689" in cli.stdout
    got = re.search(r"Inferred mu: ([0-9.]+);", cli.stdout)
    assert abs(float(got.group(1)) - ref_mu) < 5e-6 * ref_mu


# ---- 8(f) row 4: credible intervals from a marginal ---------------------------------------------
def test_quantiles_credible_interval(ctx, readme_obs):
    import torch
    for N in (101, 4096, 65536):
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), posterior=None)
        for marg in (res.marg_mu, res.marg_sigma):
            m_dev = torch.from_numpy(marg).cuda()
            cdf_dev = torch.empty(N, dtype=torch.float64, device="cuda")
            probs = [0.05, 0.5, 0.95]
            idx = ctx.quantiles(m_dev, N, probs, cdf_dev)
            cdf = np.cumsum(marg)
            want = [int(np.argmax(cdf > p)) for p in probs]      # notebook cell 5: cdf > p, first hit
            assert idx.tolist() == want
            assert np.abs(cdf_dev.cpu().numpy() - cdf).max() < 1e-12
    # degenerate: all mass in one cell, and a probability that is never exceeded
    one = torch.zeros(10, dtype=torch.float64, device="cuda")
    one[7] = 1.0
    assert ctx.quantiles(one, 10, [0.05, 0.95, 1.0]).tolist() == [7, 7, 10]


def test_reference_main_unchanged_on_libcge():
    """The reference's src/grid.cu, byte for byte, compiled against include/compat/inc/*.h ->
    cge_compat.h -> libcge.so (oracle/Makefile target grid_on_libcge) prints what the reference's
    own build of the same file prints."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "grid")
    ours = os.path.join(ROOT, "oracle", "_ref", "grid_on_libcge")
    if not (os.path.exists(ref_bin) and os.path.exists(ours)):
        pytest.skip("oracle/_ref binaries not built (need the reference checkout at build time)")
    a = subprocess.run([ref_bin], capture_output=True, text=True, timeout=300)
    b = subprocess.run([ours], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, (a.stdout, a.stderr, b.stdout, b.stderr)
    assert "Inferred mu" in a.stdout
    assert a.stdout.strip().splitlines()[-2:] == b.stdout.strip().splitlines()[-2:], (a.stdout, b.stdout)


@pytest.mark.parametrize("variant", [0, 10, 9])   # default (tall, 7 x 8), wide (6 x 8), 4-column slope
def test_non_flat_prior_on_the_neighbour_row_policies(readme_obs, variant):
    """The prior is multiplied in the likelihood kernel's epilogue: the same check on a grid fine
    enough for the 8-column policies (the prior path stores from the unpacked cells, N = 1100 is
    not a multiple of 8 at the last strip)."""
    import torch
    N, mu_r, sg_r = 1100, (19.6, 20.4), (1.6, 2.4)
    mu = oracle.linspace(N, *mu_r).astype(np.float64)
    sg = oracle.linspace(N, *sg_r).astype(np.float64)
    pm = np.exp(-0.5 * ((mu - 20.2) / 0.1) ** 2).astype(np.float32)
    ps = (1.0 / sg).astype(np.float32)
    pf = np.random.default_rng(6).uniform(0.5, 1.5, size=(N, N)).astype(np.float32)
    likes = oracle.grid_likes(readme_obs, N, *mu_r, *sg_r, oracle.NUM_F64)

    def ref(prior):
        post = likes * prior
        post /= post.sum()
        return dict(posterior=post, marg_mu=post.sum(axis=1), marg_sigma=post.sum(axis=0),
                    expectations=np.array([np.sum(mu * post.sum(axis=1)), np.sum(sg * post.sum(axis=0))]))

    with cge.Context(0) as ctx:
        flat = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
        if variant == 0:
            assert ctx.scale_info()["product_path"] == "slope8x7"
        ctx.set_prior(pm, ps, N)
        res = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
        _check(res, ref(pm.astype(np.float64)[:, None] * ps.astype(np.float64)[None, :]))
        assert abs(res.expectations[0] - flat.expectations[0]) > 0.01
        ctx.set_prior(None, ps, N, torch.from_numpy(pf).cuda())
        _check(ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant),
               ref(pf.astype(np.float64) * ps.astype(np.float64)[None, :]))
        ctx.set_prior()
        again = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
        assert again.posterior.tobytes() == flat.posterior.tobytes()
