"""GPU (-m gpu): parity of the CUDA hot path, called through the C ABI, against the CPU oracle,
the committed golden fixtures and size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.md section 2 / SURVEY section 8(d)):
  |dE[mu]|, |dE[sigma]| <= 1e-5 (expected ~1e-8)
  posterior: relative error <= 1e-4 on cells with P >= 1e-20 * P_max; below that floor the
             absolute error is <= 1e-4 * 1e-20 * P_max (cells more than ~50 decades under the
             peak flush to zero in float32)
  marginals: abs <= 1e-7 * max or rel <= 1e-4
  sum(posterior) = 1 +- 1e-6
"""
import os

import numpy as np
import pytest

import cuda_grid_estimation_b200 as cge
import oracle
from cuda_grid_estimation_b200 import api

pytestmark = pytest.mark.gpu

E_TOL = 1e-5
POST_REL = 1e-4
FLOOR = 1e-20


def check_against(res, ref, post_rel=POST_REL, e_tol=E_TOL, what=""):
    d_e = np.abs(res.expectations - ref["expectations"]).max()
    assert d_e <= e_tol, f"{what}: E {res.expectations} vs {ref['expectations']}"
    for key, got in (("marg_mu", res.marg_mu), ("marg_sigma", res.marg_sigma)):
        want = ref[key]
        ok = (np.abs(got - want) <= 1e-7 * want.max()) | (np.abs(got - want) <= 1e-4 * np.abs(want))
        assert ok.all(), f"{what}: {key} max abs {np.abs(got - want).max()}"
    if res.posterior is not None and ref.get("posterior") is not None:
        P, Q = res.posterior.astype(np.float64), ref["posterior"]
        assert P.shape == Q.shape
        big = Q >= FLOOR * Q.max()
        rel = np.abs(P[big] - Q[big]) / Q[big]
        assert rel.max() <= post_rel, f"{what}: posterior max rel err {rel.max():.3e}"
        # below the floor: absolute, at the floor's own scale (flushed tail included)
        assert np.abs(P[~big] - Q[~big]).max(initial=0.0) <= post_rel * FLOOR * Q.max()
        assert abs(P.sum() - 1.0) <= 1e-6
    return d_e


# ---------------------------------------------------------------------------------------
# config 1: README workload
# ---------------------------------------------------------------------------------------
def test_readme_product_vs_oracle(ctx, readme_obs):
    res = ctx.grid_posterior(readme_obs, 101, (18, 22), (1, 3), mode=cge.MODE_PRODUCT_F32)
    ref = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_REF)
    d = check_against(res, ref, what="readme/product vs REF")
    assert d < 1e-6  # expected ~1e-8
    f64 = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    check_against(res, f64, what="readme/product vs F64")
    assert res.launches == 1  # small grids: the whole step is one launch (small_grid_kernel)
    unf = ctx.grid_posterior(readme_obs, 101, (18, 22), (1, 3), variant=api.VARIANT_UNFUSED)
    assert unf.launches == 3  # likelihood (set-up fused), fold + publish, finish


@pytest.mark.parametrize("variant", [api.VARIANT_MUFU_ONLY, api.VARIANT_HYBRID, api.VARIANT_UNFUSED])
def test_product_kernel_variants(ctx, readme_obs, variant):
    """The MUFU-only scalar kernel, the forced hybrid (MUFU + degree-5 FMA polynomial) and the
    default through the three-launch path meet the same tolerances; N=1024 uses the wide tile
    shape, N=101 the narrow one."""
    for N in (101, 1024):
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=variant)
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_REF)
        d = check_against(res, ref, what=f"variant {variant} N={N}")
        assert d < 1e-6
    with pytest.raises(cge.CgeError):
        ctx.grid_posterior(readme_obs, 64, (18, 22), (1, 3), variant=77)


def test_readme_logspace_vs_oracle(ctx, readme_obs):
    res = ctx.grid_posterior(readme_obs, 101, (18, 22), (1, 3), mode=cge.MODE_LOGSPACE)
    f64 = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    d = check_against(res, f64, post_rel=2e-5, what="readme/log vs F64")
    assert d < 1e-7


def test_readme_vs_reference_python_golden(ctx, readme_obs, golden_dir):
    """Reference's own numpy_estimator output (np.linspace ranges differ from the C++ formula by
    <= 1.9e-6, which moves cells by up to ~7e-5 relative; expectations by ~4e-8)."""
    g = np.load(os.path.join(golden_dir, "numpy_estimator_readme.npz"))
    res = ctx.grid_posterior(g["observations"], 101, (18, 22), (1, 3))
    assert np.abs(res.expectations - g["expectations_axis_correct"]).max() <= E_TOL
    Q = g["posterior_sigma_mu"].T
    big = Q >= 1e-12 * Q.max()
    rel = np.abs(res.posterior.astype(np.float64)[big] - Q[big]) / Q[big]
    assert rel.max() <= 2e-4


@pytest.mark.parametrize("name", ["readme", "small", "max256"])
def test_reference_cuda_golden(ctx, golden_dir, name):
    """Outputs of the reference CUDA pipeline itself (oracle/ref_harness.cu run on a B200 by
    tests/golden/make_ref_cuda_golden.py): the reference's result on the same inputs."""
    g = np.load(os.path.join(golden_dir, f"ref_cuda_{name}.npz"))
    N = int(g["N"])
    rng = [float(x) for x in g["ranges"]]
    for mode in (cge.MODE_PRODUCT_F32, cge.MODE_LOGSPACE):
        res = ctx.grid_posterior(g["observations"], N, rng[:2], rng[2:], mode=mode)
        if "posterior" in g:
            ref = dict(expectations=g["expectations"], marg_mu=g["marg_mu"],
                       marg_sigma=g["marg_sigma"], posterior=g["posterior"].reshape(N, N))
            check_against(res, ref, what=f"{name} mode={mode} vs reference CUDA binary")
        else:
            ref = dict(expectations=g["expectations"], marg_mu=g["marg_mu"], marg_sigma=g["marg_sigma"])
            check_against(res, ref, what=f"{name} mode={mode} vs reference CUDA binary")
            rows, Q = g["sample_rows"], g["posterior_rows"]
            P = res.posterior.astype(np.float64)[rows]
            big = Q >= FLOOR * g["marg_mu"].max() / N
            assert (np.abs(P[big] - Q[big]) / Q[big]).max() <= POST_REL


# ---------------------------------------------------------------------------------------
# edge cases: ragged sizes, tiny grids, single observation, grid missing the MLE
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [2, 3, 5, 100, 127, 128, 129, 255, 256, 257, 511, 513, 1000, 1023, 1025])
def test_ragged_grid_sizes(ctx, readme_obs, N):
    for mode, num in ((cge.MODE_PRODUCT_F32, oracle.NUM_REF), (cge.MODE_LOGSPACE, oracle.NUM_F64)):
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, num)
        check_against(res, ref, what=f"N={N} mode={mode}")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 49, 51, 64, 255, 256])
def test_observation_counts(ctx, n):
    obs = oracle.readme_observations(n=n, seed=7)
    for mode, num in ((cge.MODE_PRODUCT_F32, oracle.NUM_REF), (cge.MODE_LOGSPACE, oracle.NUM_F64)):
        res = ctx.grid_posterior(obs, 96, (18, 22), (1, 3), mode=mode)
        ref = oracle.grid_posterior(obs, 96, 18, 22, 1, 3, num)
        check_against(res, ref, what=f"n={n} mode={mode}")


def test_grid_without_mle_and_small_sigma(ctx, readme_obs):
    for mu_r, sg_r in (((25, 30), (1, 3)), ((18, 22), (0.25, 0.75)), ((-5, 5), (4, 9)), ((19, 20), (1.5, 2.5))):
        for mode, num in ((cge.MODE_PRODUCT_F32, oracle.NUM_REF), (cge.MODE_LOGSPACE, oracle.NUM_F64)):
            res = ctx.grid_posterior(readme_obs, 64, mu_r, sg_r, mode=mode)
            ref = oracle.grid_posterior(readme_obs, 64, *mu_r, *sg_r, num)
            check_against(res, ref, what=f"{mu_r} {sg_r} mode={mode}")


def test_single_point_grid(ctx, readme_obs):
    """N=1: the reference's range formula is 0/0; libcge defines the single point as `start`.
    One cell => posterior is exactly 1."""
    res = ctx.grid_posterior(readme_obs, 1, (20, 20), (2, 2))
    assert res.posterior.shape == (1, 1) and res.posterior[0, 0] == pytest.approx(1.0, abs=1e-6)
    assert res.expectations == pytest.approx([20.0, 2.0], abs=1e-5)


def test_determinism_bitwise(ctx, readme_obs):
    runs = [ctx.grid_posterior(readme_obs, 1000, (18, 22), (1, 3)) for _ in range(3)]
    for r in runs[1:]:
        assert r.posterior.tobytes() == runs[0].posterior.tobytes()
        assert r.marg_mu.tobytes() == runs[0].marg_mu.tobytes()
        assert r.marg_sigma.tobytes() == runs[0].marg_sigma.tobytes()
        assert r.expectations.tobytes() == runs[0].expectations.tobytes()


def test_many_observations_product_and_streamed_logspace(ctx):
    # product path with a few hundred observations (beyond the reference's 256 limit)
    obs = oracle.readme_observations(n=300, seed=3)
    res = ctx.grid_posterior(obs, 128, (18, 22), (1, 3), mode=cge.MODE_PRODUCT_F32)
    check_against(res, oracle.grid_posterior(obs, 128, 18, 22, 1, 3, oracle.NUM_LOG), what="n=300 product")
    # one big shared-memory load (n <= 16384) and the double-buffered streamed path (n > 16384)
    for n in (9001, 16384, 16385, 20000, 40003):
        obs = oracle.readme_observations(n=n, seed=11)
        res = ctx.grid_posterior(obs, 64, (19.9, 20.1), (1.9, 2.1), mode=cge.MODE_LOGSPACE)
        ref = oracle.grid_posterior(obs, 64, 19.9, 20.1, 1.9, 2.1, oracle.NUM_LOG)
        check_against(res, ref, what=f"log-space n={n}")


@pytest.mark.parametrize("variant", [0, api.VARIANT_LOG_F64])
def test_logspace_kernel_variants(ctx, variant):
    """The tiled float32x2 log-space kernel (default: tiles of 256 observations)
    and the float64 one-DFMA-per-eval kernel (5) against the float64 oracle: a few hundred
    and 10^4 observations on the README ranges, a grid that does not contain the sample mean (the
    centre clamps to an edge row, so the centred terms are large), a narrow sigma range at small
    sigma (large |a_j|), and n not a multiple of the 64-observation tile."""
    cases = [
        (257, 1000, (18, 22), (1, 3)),
        (1024, 10000, (18, 22), (1, 3)),
        (130, 777, (21, 23), (1, 3)),       # sample mean ~20 is below the grid
        (130, 5000, (19.8, 20.2), (0.5, 0.7)),
        (64, 63, (18, 22), (1, 3)),
        (64, 65, (18, 22), (1, 3)),
        (64, 257, (18, 22), (1, 3)),
    ]
    for N, n, mu_r, sg_r in cases:
        obs = oracle.readme_observations(n=n, seed=5)
        res = ctx.grid_posterior(obs, N, mu_r, sg_r, mode=cge.MODE_LOGSPACE, variant=variant)
        ref = oracle.grid_posterior(obs, N, *mu_r, *sg_r, oracle.NUM_LOG)
        check_against(res, ref, what=f"log-space variant {variant} N={N} n={n} mu={mu_r} sigma={sg_r}")


@pytest.mark.parametrize("N", [1, 2, 3, 5, 64, 101, 128, 129, 200, 255, 256, 257])
def test_single_launch_small_grids(ctx, readme_obs, N):
    """Grids up to 256 points per axis run as one launch (tables, likelihood, fold, results and
    scale in small_grid_kernel).  Same cells as the three-launch path bit for bit before the 1/Z
    scale; sums agree to float64 rounding; both against the oracle; repeatable bit for bit."""
    for mode, num in ((cge.MODE_PRODUCT_F32, oracle.NUM_REF), (cge.MODE_LOGSPACE, oracle.NUM_F64)):
        one = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        five = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode,
                                  variant=api.VARIANT_UNFUSED)
        assert one.launches == (1 if N <= 256 else 3) and five.launches == 3
        assert np.abs(one.expectations - five.expectations).max() < 1e-9
        assert np.abs(one.marg_mu - five.marg_mu).max() <= 1e-9 * five.marg_mu.max()
        assert np.abs(one.marg_sigma - five.marg_sigma).max() <= 1e-9 * five.marg_sigma.max()
        nz = five.posterior > 0
        assert ((one.posterior > 0) == nz).all()
        assert (np.abs(one.posterior - five.posterior)[nz] / five.posterior[nz]).max(initial=0) < 2.5e-7
        if N >= 2:
            check_against(one, oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, num),
                          what=f"single launch N={N} mode={mode}")
        again = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), mode=mode)
        assert again.posterior.tobytes() == one.posterior.tobytes()
        assert again.expectations.tobytes() == one.expectations.tobytes()


def test_single_launch_fine_small_grid_and_prior(ctx, readme_obs):
    """The one-launch path also picks neighbour rows (the shared-slope policy) on a fine small grid,
    and honours a prior."""
    mu_r, sg_r = (19.4, 19.7), (1.7, 2.1)
    one = ctx.grid_posterior(readme_obs, 250, mu_r, sg_r)
    assert one.launches == 1 and ctx.scale_info()["product_path"] == "slope"
    forced = ctx.grid_posterior(readme_obs, 250, mu_r, sg_r, variant=api.VARIANT_SLOPE)
    nz = forced.posterior > 0
    assert (np.abs(one.posterior - forced.posterior)[nz] / forced.posterior[nz]).max() < 2.5e-7
    N = 128
    pm = np.linspace(0.5, 1.5, N).astype(np.float32)
    ps = np.linspace(2.0, 1.0, N).astype(np.float32)
    ctx.set_prior(pm, ps, N)
    try:
        one = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
        five = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=api.VARIANT_UNFUSED)
    finally:
        ctx.set_prior()
    assert one.launches == 1 and five.launches == 3
    assert np.abs(one.expectations - five.expectations).max() < 1e-9
    flat = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    assert np.abs(one.expectations - flat.expectations).max() > 1e-3  # the prior did something


def test_error_behaviour(ctx, readme_obs):
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(readme_obs, 0, (18, 22), (1, 3))
    assert ei.value.code == api.ERR_BAD_ARG
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(readme_obs, 16, (18, 22), (0, 3))
    assert ei.value.code == api.ERR_BAD_ARG
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(readme_obs, 16, (18, 22), (1, 3), rows=(8, 20))
    assert ei.value.code == api.ERR_BAD_ARG
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(readme_obs, 16, (18, 22), (1, 3), mode=9)
    assert ei.value.code == api.ERR_BAD_ARG
    bad = readme_obs.copy()
    bad[3] = np.nan
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(bad, 16, (18, 22), (1, 3))
    assert ei.value.code == api.ERR_DEGENERATE
    # the context survives errors
    res = ctx.grid_posterior(readme_obs, 16, (18, 22), (1, 3))
    assert abs(res.posterior.astype(np.float64).sum() - 1) < 1e-6


# ---------------------------------------------------------------------------------------
# device-pointer phase API + row shards on one GPU (the multi-GPU data path, minus NCCL)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [257, 4100])  # 257: hybrid path; 4100: neighbour rows, odd shard edges
def test_phase_api_with_row_shards(ctx, readme_obs, N):
    import torch
    dev = torch.device("cuda", 0)
    obs_t = torch.from_numpy(readme_obs).to(dev)
    full = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    stream = torch.cuda.current_stream().cuda_stream
    spans = [cge.shard_rows(N, 3, r) for r in range(3)]
    ctxs = [cge.Context(0) for _ in spans]
    try:
        posts, parts = [], []
        for c, (rb, re) in zip(ctxs, spans):
            p = cge.make_params(50, N, (18, 22), (1, 3), cge.MODE_PRODUCT_F32, (rb, re))
            post = torch.empty((re - rb, N), dtype=torch.float32, device=dev)
            c.likelihood_partials(p, obs_t, post, stream)
            ptr, cnt = c.partials_ptr()
            assert cnt == N + 2
            parts.append(cge.device_tensor(ptr, cnt, "float64", 0))
            posts.append((p, post))
        total = torch.stack(parts).sum(dim=0)      # stands in for the all-reduce
        for part in parts:
            part.copy_(total)
        rows = []
        for c, (p, post) in zip(ctxs, posts):
            c.finalize(p, post, stream)
            ex, mm, ms, z = c.read_results(p, stream)
            # cells are bit-identical to the unsharded run (tiles are anchored to global rows);
            # the sums differ only in the order of the float64 adds and in the float32 4-row
            # patch sums that a shard edge cuts in two (one float32 rounding, 6e-8 of a 4-row sum)
            assert np.abs(ex - full.expectations).max() < 1e-8
            assert np.abs(ms - full.marg_sigma).max() < 1e-8 * full.marg_sigma.max()
            assert np.abs(mm[p.row_begin:p.row_end] - full.marg_mu[p.row_begin:p.row_end]).max() \
                < 1e-8 * full.marg_mu.max()
            assert mm[:p.row_begin].sum() == 0 and mm[p.row_end:].sum() == 0
            rows.append(post.cpu().numpy())
        got = np.concatenate(rows, axis=0)
        rel = np.abs(got - full.posterior) / np.maximum(full.posterior, 1e-30)
        assert rel[full.posterior > 0].max() < 2e-7  # same unnormalised cell, 1/Z rounded once
    finally:
        for c in ctxs:
            c.close()


def test_async_fused_entry(ctx, readme_obs):
    """cge_grid_posterior_async: device buffers on the caller's stream, nothing waited for until
    read_results; same bits as the synchronous call (one launch at N=101, three at N=1000)."""
    import torch
    dev = torch.device("cuda", 0)
    obs_t = torch.from_numpy(readme_obs).to(dev)
    side = torch.cuda.Stream()
    for N in (101, 1000):
        ref = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
        p = cge.make_params(50, N, (18, 22), (1, 3))
        post = torch.empty((N, N), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for _ in range(3):  # back to back, no host wait in between
                ctx.grid_posterior_async(p, obs_t, post, side.cuda_stream)
            ex, mm, ms, z = ctx.read_results(p, side.cuda_stream)
        assert post.cpu().numpy().tobytes() == ref.posterior.tobytes()
        assert ex.tobytes() == ref.expectations.tobytes()
        assert ms.tobytes() == ref.marg_sigma.tobytes() and mm.tobytes() == ref.marg_mu.tobytes()
    with pytest.raises(cge.CgeError):  # a shard needs the exchange between the phases
        ctx.grid_posterior_async(cge.make_params(50, 101, (18, 22), (1, 3), rows=(0, 50)), obs_t,
                                 post, None)


def test_device_buffers_and_unaligned_posterior(ctx, readme_obs):
    import torch
    dev = torch.device("cuda", 0)
    N = 128
    obs_t = torch.from_numpy(readme_obs).to(dev)
    backing = torch.zeros(N * N + 1, dtype=torch.float32, device=dev)
    ref = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    torch.cuda.synchronize()  # the synchronous ABI runs on the context's own stream
    for off in (0, 1):  # off=1 -> 4-byte aligned only: scalar store / scalar scale paths
        post = backing[off:off + N * N]
        res = ctx.grid_posterior(obs_t, N, (18, 22), (1, 3), posterior=post)
        got = post.cpu().numpy().reshape(N, N)
        assert got.tobytes() == ref.posterior.tobytes()
        assert res.expectations.tobytes() == ref.expectations.tobytes()


# ---------------------------------------------------------------------------------------
# config 2: 4096 x 4096 x 50, whole posterior against the oracle
# ---------------------------------------------------------------------------------------
def test_4096_product_vs_oracle(ctx, readme_obs):
    N = 4096
    res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_REF)
    check_against(res, ref, what="4096^2 product vs REF")
    cf = oracle.closed_form_posterior(readme_obs, N, 18, 22, 1, 3)
    assert np.abs(res.expectations - cf["expectations"]).max() < 1e-6


# ---------------------------------------------------------------------------------------
# paired-rows product kernel (the default on fine grids): half the factors come from MUFU.EX2,
# the other half are derived from the row above by a degree-2 polynomial
# ---------------------------------------------------------------------------------------
_REF_4096 = {}


def _ref_4096(readme_obs):
    if "f64" not in _REF_4096:
        _REF_4096["f64"] = oracle.grid_posterior(readme_obs, 4096, 18, 22, 1, 3, oracle.NUM_F64)
    return _REF_4096["f64"]


@pytest.mark.parametrize("variant", [0, 9, 10, 3, 8, 4, 1])
def test_paired_rows_variants_4096(ctx, readme_obs, variant):
    """README ranges at N=4096: |u| <= 0.0084, inside the neighbour-row kernels' bound (0.011), and
    the shared-slope term at 5.85e-8 for 8 columns (bound 2^-24 = 5.96e-8).  Every neighbour-row
    variant (and the hybrid one they replace here) against the float64 oracle."""
    res = ctx.grid_posterior(readme_obs, 4096, (18, 22), (1, 3), variant=variant)
    d = check_against(res, _ref_4096(readme_obs), what=f"variant {variant} N=4096 vs F64")
    assert d < 1e-6


def test_default_picks_the_path_on_the_device(ctx, readme_obs):
    """The default kernel chooses its arithmetic from ScaleInfo (ranges, N and the observations,
    all on the device).  mu in [18, 22]: N=512 is too coarse for neighbour rows (|u| <= 0.067 ->
    hybrid), N=1024 takes the degree-3 polynomial (0.034); N=3300 (0.0105) the degree-2 one when
    the sigma columns are far apart (sigma in [1, 4]: shared-slope term 7.3e-8 > 2^-24) and the
    shared-slope policy on sigma in [1, 3] (5.8e-8 for 4 columns, too much for 8); N=4096 runs
    the 8-column wide policy on the whole grid.  Each is bit-identical to the corresponding forced
    variant and differs from the others.  A far-out observation raises |u| and moves N=4096 to
    the degree-3 polynomial."""
    forced = {"hybrid": api.VARIANT_HYBRID, "paired3": api.VARIANT_PAIRED3, "paired": api.VARIANT_PAIRED,
              "slope": api.VARIANT_SLOPE, "slope8": api.VARIANT_SLOPE_WIDE, "slope8x7": api.VARIANT_SLOPE_TALL}
    bw = api.band_columns(0)
    for N, sg_r, want in ((512, (1, 3), "hybrid"), (1024, (1, 3), "paired3"), (3300, (1, 4), "paired"),
                          (3300, (1, 3), "slope"), (4096, (1, 3), "slope8")):
        d = ctx.grid_posterior(readme_obs, N, (18, 22), sg_r, variant=api.VARIANT_UNFUSED)
        info = ctx.scale_info()
        assert info["product_path"] == want, (N, sg_r, info)
        paths = api.band_paths(info, oracle.linspace(N, *sg_r))
        assert paths[0] == want   # the strip at the smallest sigma is the most demanding one
        for name, variant in forced.items():
            f = ctx.grid_posterior(readme_obs, N, (18, 22), sg_r, variant=variant)
            # same arithmetic in a strip <=> same unnormalised cells; the normaliser differs in the
            # last bits when other strips took another path, so compare ratios along a row
            for b, took in enumerate(paths):
                sl = slice(b * bw, min((b + 1) * bw, N))
                row = N // 2 + 1   # an odd row: a derived row of the paired kernels
                ratio = d.posterior[row, sl].astype(np.float64) / f.posterior[row, sl].astype(np.float64)
                if took == name:
                    assert np.ptp(ratio) < 3e-7, (N, b, name, np.ptp(ratio))
                elif b == 0 and N <= 1024:
                    # (far strips and finer grids: the policies agree to rounding)
                    assert np.ptp(ratio) > 3e-7, (N, b, name, took, np.ptp(ratio))
    # The choice is per column band (|u| scales with 0.72 / sigma^2).  A far-out observation
    # (reach 12 instead of ~5: |u| <= 0.017 at sigma = 1) moves only the band at the smallest
    # sigma of the N=4096 grid to the degree-3 polynomial; the next ones take degree 2 and, as
    # sigma grows, the shared-slope policy.
    far = readme_obs.copy()
    far[0] = 30.0
    d = ctx.grid_posterior(far, 4096, (18, 22), (1, 3))
    info = ctx.scale_info()
    assert info["product_path"] == "paired3"
    paths = api.band_paths(info, oracle.linspace(4096, 1, 3))
    assert paths[0] == "paired3" and paths[-1] == "slope" and "hybrid" not in paths
    for name in ("paired", "paired3", "slope"):
        f = ctx.grid_posterior(far, 4096, (18, 22), (1, 3), variant=forced[name])
        for b, want in enumerate(paths):
            same = d.posterior[:, b * bw:(b + 1) * bw].tobytes() == f.posterior[:, b * bw:(b + 1) * bw].tobytes()
            # (normalised cells: the two runs share Z only to rounding, so compare unnormalised ratios)
            ratio = d.posterior[2000, b * bw:(b + 1) * bw].astype(np.float64) / \
                f.posterior[2000, b * bw:(b + 1) * bw].astype(np.float64)
            assert same == (want == name) or (want == name and np.ptp(ratio) < 3e-7), (b, name)
    ref = oracle.grid_posterior(far, 4096, 18, 22, 1, 3, oracle.NUM_F64)
    check_against(d, ref, what="far observation, N=4096")
    # a wide sigma range: hybrid only where sigma is small, neighbour rows in the other bands
    d = ctx.grid_posterior(readme_obs, 4096, (18, 22), (0.4, 3))
    paths = api.band_paths(ctx.scale_info(), oracle.linspace(4096, 0.4, 3))
    assert paths[0] not in ("paired", "slope") and paths[-1] == "slope", paths
    check_against(d, oracle.grid_posterior(readme_obs, 4096, 18, 22, 0.4, 3, oracle.NUM_F64),
                  what="wide sigma range, N=4096")


@pytest.mark.parametrize("N", [2, 3, 64, 130, 301, 700, 1026])
def test_paired_rows_on_small_fine_grids(ctx, readme_obs, N):
    """Narrow ranges make a small grid fine enough for the neighbour-row kernels, so they also run
    in the narrow tile shapes, with odd N (last row has no partner) and N not a multiple of 4."""
    mu_r, sg_r = ((19.4, 19.7) if N > 64 else (19.5, 19.5 + 2e-3 * (N - 1))), (1.7, 2.1)
    d = ctx.grid_posterior(readme_obs, N, mu_r, sg_r)
    info = ctx.scale_info()
    took = info["product_path"]
    assert took in ("slope8x7", "slope8", "slope", "paired")   # (small N: the sigma columns are too far apart for "slope")
    variants = {"slope8x7": api.VARIANT_SLOPE_TALL, "slope8": api.VARIANT_SLOPE_WIDE, "slope": api.VARIANT_SLOPE,
                "paired": api.VARIANT_PAIRED}
    paths = api.band_paths(info, oracle.linspace(N, *sg_r))
    assert set(paths) <= set(variants)
    if len(set(paths)) == 1:
        forced = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variants[took])
        assert d.posterior.tobytes() == forced.posterior.tobytes()
    else:   # the column bands differ (N = 130: the two columns of the second band are eligible for
            # "slope", the first band is not): same arithmetic in a band <=> a constant ratio
        bw = api.band_columns(N)
        for b, name in enumerate(paths):
            f = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variants[name])
            sl = slice(b * bw, min((b + 1) * bw, N))
            ratio = d.posterior[N // 2, sl].astype(np.float64) / f.posterior[N // 2, sl].astype(np.float64)
            assert np.ptp(ratio) < 3e-7, (N, b, name, np.ptp(ratio))
    ref = oracle.grid_posterior(readme_obs, N, *mu_r, *sg_r, oracle.NUM_F64)
    for name, variant in variants.items():
        if name != took:
            other = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
            check_against(other, ref, what=f"narrow ranges N={N}, forced {name}")
    check_against(d, ref, what=f"narrow ranges N={N}")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 49, 51, 64, 200])
def test_paired_rows_observation_counts(ctx, n):
    """Observation counts around the 4-per-iteration loop of the paired kernel (fine grid, wide
    tile shape and the single-launch path)."""
    obs = oracle.readme_observations(n=n, seed=7)
    for N, mu_r in ((200, (19.9, 20.1)), (1100, (19.6, 20.4))):
        sg_r = (1.6, 2.4)
        ref = oracle.grid_posterior(obs, N, *mu_r, *sg_r, oracle.NUM_F64)
        res = ctx.grid_posterior(obs, N, mu_r, sg_r)
        # (the single launch at N = 200 has the 4-column policies only; N = 1100 runs the tall one)
        assert ctx.scale_info()["product_path"] == ("slope" if N == 200 else "slope8x7")
        check_against(res, ref, what=f"default n={n} N={N}")
        for variant in (api.VARIANT_PAIRED, api.VARIANT_SLOPE, api.VARIANT_SLOPE_WIDE, api.VARIANT_SLOPE_TALL):
            res = ctx.grid_posterior(obs, N, mu_r, sg_r, variant=variant)
            check_against(res, ref, what=f"variant {variant} n={n} N={N}")


def test_descending_and_sign_crossing_ranges(ctx):
    """Ranges given high-to-low and a mu range that crosses zero: the row-pair identity uses the
    signed spacing, the eligibility bound its magnitude."""
    rng = np.random.default_rng(3)
    obs = rng.normal(0.1, 1.0, 40).astype(np.float32)
    for N in (150, 2048):
        for mu_r, sg_r in (((0.6, -0.4), (0.7, 1.5)), ((-0.4, 0.6), (1.5, 0.7))):
            for mode, num in ((cge.MODE_PRODUCT_F32, oracle.NUM_F64), (cge.MODE_LOGSPACE, oracle.NUM_LOG)):
                res = ctx.grid_posterior(obs, N, mu_r, sg_r, mode=mode)
                ref = oracle.grid_posterior(obs, N, *mu_r, *sg_r, num)
                check_against(res, ref, what=f"N={N} mu={mu_r} sigma={sg_r} mode={mode}")
    res = ctx.grid_posterior(obs, 2048, (0.6, -0.4), (0.7, 1.5))
    assert ctx.scale_info()["product_path"] in ("slope8x7", "slope8", "slope", "paired", "paired3")


def test_paired_rows_at_the_eligibility_bound(ctx, readme_obs):
    """Forced paired rows on grids around the bound: the polynomial's error grows as |u|^3, so
    the result must still be well inside tolerance at the bound and degrade gracefully past it
    (this is the margin the bound was chosen with)."""
    for N, rel_max in ((3300, 3e-5), (2600, 1e-4)):  # |u| <= 0.0105, 0.0133
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=api.VARIANT_PAIRED)
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_F64)
        check_against(res, ref, post_rel=rel_max, what=f"forced paired N={N}")
    # shared slope: the dropped term is k2 u^2 eps per factor -- 4 columns: 2.5e-8 at N = 4096,
    # 9.7e-8 at N = 2600; 8 columns: 5.85e-8 at N = 4096 (the bound is 2^-24 = 5.96e-8), 2.3e-7 at 2600
    for N, rel_max in ((4096, 3e-5), (2600, 1e-4)):
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_F64)
        for variant in (api.VARIANT_SLOPE, api.VARIANT_SLOPE_WIDE):
            res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=variant)
            check_against(res, ref, post_rel=rel_max, what=f"forced variant {variant} N={N}")
    # tall policy (one exponential per 7 rows): eligible from N ~ 9400 on the README ranges
    # (3|u| <= 0.011, slope term 9 x the wide policy's); forced at N = 7000: 3|u| = 0.0147, 1.05e-7
    for N, rel_max, want in ((9600, 3e-5, "slope8x7"), (7000, 1e-4, "slope8")):
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_F64)
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
        assert ctx.scale_info()["product_path"] == want
        check_against(res, ref, post_rel=3e-5, what=f"default N={N}")
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=api.VARIANT_SLOPE_TALL)
        check_against(res, ref, post_rel=rel_max, what=f"forced tall N={N}")
    for N, rel_max in ((700, 3e-5), (560, 1e-4)):  # degree 3: |u| <= 0.0495, 0.062
        res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), variant=api.VARIANT_PAIRED3)
        ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_F64)
        check_against(res, ref, post_rel=rel_max, what=f"forced paired3 N={N}")


# ---------------------------------------------------------------------------------------
# configs 3 and 5 at full size: size-independent properties + sampled rows
# ---------------------------------------------------------------------------------------
def _closed_form_logz(obs, N, ranges, chunk=1024):
    """(log sum_ij exp(loglik_ij), max_ij loglik_ij) in float64, by row chunks (independent
    closed form)."""
    m = -np.inf
    s = 0.0
    for r0 in range(0, N, chunk):
        ll = oracle.closed_form_loglik(obs, N, *ranges, rows=np.arange(r0, min(N, r0 + chunk)))
        mc = ll.max()
        if mc > m:
            s = s * np.exp(m - mc) if np.isfinite(m) else 0.0
            m = mc
        s += np.exp(ll - m).sum()
    return m + np.log(s), m


def _full_size_properties(ctx, obs, N, mode, sample_rows):
    import torch
    ranges = (18.0, 22.0, 1.0, 3.0)
    res = ctx.grid_posterior(obs, N, ranges[:2], ranges[2:], mode=mode, posterior="device")
    ptr, cnt = ctx.posterior_ptr()
    assert cnt == N * N
    post = cge.device_tensor(ptr, cnt, "float32", 0).view(N, N)
    # sum = 1, marginals are the row/column sums of what was written
    total = post.sum(dtype=torch.float64).item()
    assert abs(total - 1.0) <= 1e-6
    rows = post.sum(dim=1, dtype=torch.float64).cpu().numpy()
    cols = post.sum(dim=0, dtype=torch.float64).cpu().numpy()
    assert np.abs(rows - res.marg_mu).max() <= 1e-7 * res.marg_mu.max()
    assert np.abs(cols - res.marg_sigma).max() <= 1e-7 * res.marg_sigma.max()
    assert abs(res.marg_mu.sum() - 1) < 1e-9 and abs(res.marg_sigma.sum() - 1) < 1e-9
    # expectations against the independent closed form
    mu = oracle.linspace(N, *ranges[:2]).astype(np.float64)
    sg = oracle.linspace(N, *ranges[2:]).astype(np.float64)
    logz, llmax = _closed_form_logz(obs, N, ranges)
    e_mu = e_sg = 0.0
    for r0 in range(0, N, 1024):
        idx = np.arange(r0, min(N, r0 + 1024))
        p = np.exp(oracle.closed_form_loglik(obs, N, *ranges, rows=idx) - logz)
        e_mu += (p.sum(axis=1) * mu[idx]).sum()
        e_sg += (p.sum(axis=0) * sg).sum()
    assert abs(res.expectations[0] - e_mu) <= E_TOL and abs(res.expectations[1] - e_sg) <= E_TOL
    # sampled rows, cell by cell
    for i in sample_rows:
        want = np.exp(oracle.closed_form_loglik(obs, N, *ranges, rows=np.array([i]))[0] - logz)
        got = post[i].cpu().numpy().astype(np.float64)
        pmax = np.exp(llmax - logz)
        big = want >= FLOOR * pmax
        if big.any():
            rel = np.abs(got[big] - want[big]) / want[big]
            assert rel.max() <= POST_REL, f"row {i}: {rel.max():.3e}"
    return res


def test_16384_product_properties(ctx, readme_obs):
    N = 16384
    peak_row = int(round((19.549 - 18) / 4 * (N - 1)))
    _full_size_properties(ctx, readme_obs, N, cge.MODE_PRODUCT_F32,
                          [0, 1, peak_row - 1500, peak_row, peak_row + 1, peak_row + 2000, N - 1])


def test_65536_product_properties(ctx, readme_obs):
    """BASELINE config 4 at full size on one GPU (17.2 GB posterior): sum = 1, marginals are the
    row / column sums of what was written, expectations and sampled rows against an independent
    float64 closed form (sufficient statistics) evaluated with torch on the device."""
    import torch
    N = 65536
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~20 GB of device memory for the posterior and its checks")
    dev = torch.device("cuda", 0)
    ranges = (18.0, 22.0, 1.0, 3.0)
    res = ctx.grid_posterior(readme_obs, N, ranges[:2], ranges[2:], posterior="device")
    assert ctx.scale_info()["product_path"] == "slope8x7"
    ptr, cnt = ctx.posterior_ptr()
    post = cge.device_tensor(ptr, cnt, "float32", 0).view(N, N)
    mu = torch.from_numpy(oracle.linspace(N, *ranges[:2]).astype(np.float64)).to(dev)
    sg = torch.from_numpy(oracle.linspace(N, *ranges[2:]).astype(np.float64)).to(dev)
    x = torch.from_numpy(readme_obs.astype(np.float64)).to(dev)
    n, xbar = x.numel(), x.mean()
    ss = ((x - xbar) ** 2).sum()

    def loglik(rows):  # [len(rows), N] float64
        s_i = ss + n * (xbar - mu[rows]) ** 2
        return -n * torch.log(sg)[None, :] - 0.5 * n * np.log(2 * np.pi) \
            - s_i[:, None] / (2 * sg ** 2)[None, :]

    chunk = 2048
    llmax = max(float(loglik(torch.arange(r0, r0 + chunk, device=dev)).max()) for r0 in range(0, N, chunk))
    z = e_mu = e_sg = total = 0.0
    rows_sum = torch.empty(N, dtype=torch.float64, device=dev)
    cols_sum = torch.zeros(N, dtype=torch.float64, device=dev)
    for r0 in range(0, N, chunk):
        idx = torch.arange(r0, r0 + chunk, device=dev)
        p = torch.exp(loglik(idx) - llmax)
        z += float(p.sum())
        e_mu += float((p.sum(dim=1) * mu[idx]).sum())
        e_sg += float((p.sum(dim=0) * sg).sum())
        blk = post[r0:r0 + chunk].to(torch.float64)
        rows_sum[idx] = blk.sum(dim=1)
        cols_sum += blk.sum(dim=0)
        total += float(blk.sum())
    assert abs(total - 1.0) <= 1e-6
    assert abs(res.expectations[0] - e_mu / z) <= E_TOL and abs(res.expectations[1] - e_sg / z) <= E_TOL
    assert np.abs(rows_sum.cpu().numpy() - res.marg_mu).max() <= 1e-7 * res.marg_mu.max()
    assert np.abs(cols_sum.cpu().numpy() - res.marg_sigma).max() <= 1e-7 * res.marg_sigma.max()
    peak_row = int(round((19.549 - 18) / 4 * (N - 1)))
    for i in (0, 1, peak_row - 6000, peak_row, peak_row + 1, peak_row + 8000, N - 1):
        want = torch.exp(loglik(torch.tensor([i], device=dev))[0] - llmax) / z
        got = post[i].to(torch.float64)
        big = want >= FLOOR / z
        if bool(big.any()):
            rel = ((got[big] - want[big]).abs() / want[big]).max().item()
            assert rel <= POST_REL, f"row {i}: {rel:.3e}"


def test_16384_logspace_10k_observations(ctx):
    N = 16384
    obs = oracle.readme_observations(n=10000)
    peak_row = int(round((float(obs.astype(np.float64).mean()) - 18) / 4 * (N - 1)))
    _full_size_properties(ctx, obs, N, cge.MODE_LOGSPACE,
                          [0, peak_row - 300, peak_row - 40, peak_row, peak_row + 7, peak_row + 250, N - 1])


# ---------------------------------------------------------------------------------------
# stage API through the C ABI: the reference's own known answers (src/tests.cu)
# ---------------------------------------------------------------------------------------
def test_stage_api_known_answers(ctx):
    import torch
    dev = torch.device("cuda", 0)
    f32 = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
    f64 = lambda n: torch.empty(n, dtype=torch.float64, device=dev)
    # Linspaces (tests.cu:49-67) and the checkArrays values (utils.h:16-30)
    v = f32(3)
    ctx.linspace(v, 3, 1.0, 2.0)
    assert v.cpu().tolist() == [1.0, 1.5, 2.0]
    mu, sg = f32(101), f32(101)
    ctx.linspace(mu, 101, 18.0, 22.0)
    ctx.linspace(sg, 101, 1.0, 3.0)
    assert np.array_equal(mu.cpu().numpy(), oracle.linspace(101, 18, 22))
    assert np.array_equal(sg.cpu().numpy(), oracle.linspace(101, 1, 3))
    assert mu[50].item() == 20.0 and mu[49].item() == np.float32(19.96)
    # GenerateGrids (tests.cu:70-108)
    vx, vy, vz = f32(3), f32(3), f32(2)
    ctx.linspace(vx, 3, 1, 3); ctx.linspace(vy, 3, 1, 3); ctx.linspace(vz, 2, 4, 5)
    gx, gy, gz = f32(18), f32(18), f32(18)
    ctx.meshgrid3(vx, vy, vz, gx, gy, gz, 3, 2)
    assert gx.cpu().tolist() == [1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3]
    assert gy.cpu().tolist() == [1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3]
    assert gz.cpu().tolist() == [4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5]
    # ComputeDensities (tests.cu:111-135)
    a, b, c, d = f32(3), f32(3), f32(3), f32(3)
    ctx.linspace(a, 3, 20, 22); ctx.linspace(b, 3, 1, 3); ctx.linspace(c, 3, 20, 21)
    ctx.densities(d, a, b, c, 3)
    assert d.cpu().tolist() == pytest.approx([0.39894228, 0.19333406, 0.12579441], abs=1e-5)
    # ComputeLikes (tests.cu:163-194)
    dens, likes = f32(10), f64(2)
    ctx.linspace(dens, 10, 1.0, 10.0)
    ctx.row_products(dens, likes, 10, 2)
    assert likes.cpu().tolist() == [120.0, 30240.0]
    with pytest.raises(cge.CgeError) as ei:
        ctx.row_products(dens, likes, 10, 3)  # wrappers.h:110-113 "likesSize != rows * cols"
    assert ei.value.code == api.ERR_SIZE
    # ComputePosterior (tests.cu:196-209)
    lv = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=dev)
    pv = f64(3)
    ctx.normalize(lv, pv, 3)
    got = pv.cpu().tolist()
    assert got[0] == pytest.approx(1 / 6, abs=1e-5) and got[1] == pytest.approx(2 / 6, abs=1e-5)
    assert got[2] == 0.5
    # ComputeExpectations (tests.cu:211-245)
    post = torch.full((9,), 0.1, dtype=torch.float64, device=dev)
    post[4] = 0.2
    mm, ms = f64(3), f64(3)
    ctx.marginalize(mm, post, 3, 3, 1)
    ctx.marginalize(ms, post, 3, 3, 0)
    vmu, vsg = f32(3), f32(3)
    ctx.linspace(vmu, 3, 1.0, 3.0); ctx.linspace(vsg, 3, 4.0, 6.0)
    assert ctx.expectation(mm, vmu, 3) == 2.0
    assert ctx.expectation(ms, vsg, 3) == 5.0


def test_stage_api_pipeline_matches_fused(ctx, readme_obs):
    """The reference's stage-by-stage pipeline (src/grid.cu:79-127) through the stage ABI gives
    the same answer as the fused entry point and the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    N, n = 101, 50
    obs = torch.from_numpy(readme_obs).to(dev)
    mu = torch.empty(N, dtype=torch.float32, device=dev)
    sg = torch.empty(N, dtype=torch.float32, device=dev)
    ctx.linspace(mu, N, 18.0, 22.0)
    ctx.linspace(sg, N, 1.0, 3.0)
    g = [torch.empty(N * N * n, dtype=torch.float32, device=dev) for _ in range(4)]
    ctx.meshgrid3(mu, sg, obs, g[0], g[1], g[2], N, n)
    ctx.densities(g[3], g[0], g[1], g[2], N * N * n)
    likes = torch.empty(N * N, dtype=torch.float64, device=dev)
    ctx.row_products(g[3], likes, N * N * n, N * N)
    post = torch.empty_like(likes)
    ctx.normalize(likes, post, N * N)
    mm = torch.empty(N, dtype=torch.float64, device=dev)
    ms = torch.empty(N, dtype=torch.float64, device=dev)
    ctx.marginalize(mm, post, N, N, 1)
    ctx.marginalize(ms, post, N, N, 0)
    e = np.array([ctx.expectation(mm, mu, N), ctx.expectation(ms, sg, N)])
    ref = oracle.grid_posterior(readme_obs, N, 18, 22, 1, 3, oracle.NUM_REF)
    assert np.abs(e - ref["expectations"]).max() < 1e-8
    P = post.cpu().numpy().reshape(N, N)
    big = ref["posterior"] >= FLOOR * ref["posterior"].max()
    assert (np.abs(P[big] - ref["posterior"][big]) / ref["posterior"][big]).max() < 1e-5
    # densities: a3's value to f32 accuracy
    dens_ref = oracle.densities(g[0].cpu().numpy(), g[1].cpu().numpy(), g[2].cpu().numpy())
    dens = g[3].cpu().numpy()
    assert np.abs(dens - dens_ref).max() <= 1e-5
    assert (np.abs(dens - dens_ref) / dens_ref).max() < 1e-5
    fused = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    assert np.abs(fused.expectations - e).max() < 1e-6


# ---------------------------------------------------------------------------------------
# the C++ host side: reference's own tests against include/cge_compat.h, and the grid CLI
# ---------------------------------------------------------------------------------------
def _bin(name):
    from cuda_grid_estimation_b200 import _build
    path = os.path.join(_build.BIN_DIR, name)
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
    return path


def test_cpp_compat_shim_reference_tests(tmp_path, readme_obs):
    """cpp/grid-algorithm/src/tests.cu restated (tests/cpp/test_compat.cu) + src/grid.cu's call
    sequence through the reference's own function names, compiled against the shim."""
    import subprocess
    obs_file = tmp_path / "obs.bin"
    readme_obs.tofile(obs_file)
    res = subprocess.run([_bin("test_compat"), str(obs_file)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "ALL PASSED" in res.stdout
    line = [l for l in res.stdout.splitlines() if l.startswith("PIPELINE")][0].split()
    got = np.array([float(line[1]), float(line[2])])
    ref = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_REF)
    assert np.abs(got - ref["expectations"]).max() < 1e-8


def test_cpp_grid_cli(tmp_path, readme_obs):
    """bin/grid equivalent (src/grid.cu:129-132 output format) on the fused path."""
    import subprocess
    obs_file = tmp_path / "obs.bin"
    readme_obs.tofile(obs_file)
    res = subprocess.run([_bin("grid"), "--obs-file", str(obs_file), "--timing"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.splitlines()
    assert lines[0].startswith("Inferred mu: ") and lines[0].endswith("; Actual mu: 20")
    assert lines[1].startswith("Inferred sigma: ") and lines[1].endswith("; Actual sigma: 2")
    e_mu = float(lines[0].split()[2].rstrip(";"))
    e_sg = float(lines[1].split()[2].rstrip(";"))
    assert abs(e_mu - 19.549052475) < 1e-6 and abs(e_sg - 1.916762506) < 1e-6
    # the reference's default run (no arguments) also works: 50 host-drawn observations
    res = subprocess.run([_bin("grid")], capture_output=True, text=True)
    assert res.returncode == 0 and "Inferred mu" in res.stdout


def test_cpp_grid_cli_sharded_in_process(tmp_path, readme_obs):
    """`grid --devices 0,0,0` (cge_create_multi + cge_grid_posterior_multi from the C++ main(), three
    shards through the peer exchange on one device) prints the single-GPU expectations."""
    import subprocess
    obs_file = tmp_path / "obs.bin"
    readme_obs.tofile(obs_file)

    def run(*extra):
        res = subprocess.run([_bin("grid"), "--obs-file", str(obs_file), "--grid", "1000", *extra],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        lines = res.stdout.splitlines()
        return float(lines[0].split()[2].rstrip(";")), float(lines[1].split()[2].rstrip(";"))

    one, three = run(), run("--devices", "0,0,0")
    assert abs(one[0] - three[0]) < 1e-9 and abs(one[1] - three[1]) < 1e-9
    res = subprocess.run([_bin("grid"), "--gpus", "100000"], capture_output=True, text=True)
    assert res.returncode != 0 and "Error" in res.stderr
