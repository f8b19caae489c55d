"""GPU (-m gpu): what round 2 added to the path, each through the C ABI against the oracle or
against the unsharded call:
  * product mode is independent of the ORDER of the observations (a float32 running
    product overflowed / flushed on sorted input; the kernel now consumes them in a range-balanced
    order of its own);
  * float64 posterior output (cge_params.posterior_dtype) against the reference-CUDA goldens;
  * several shards driven from one process (cge_create_multi), including several shards on ONE
    device -- which runs the peer-exchange kernels (publish flag, poll, rank-ordered peer sum,
    double-buffered vectors) on a single-GPU box;
  * the step is three launches (one for small grids)."""
import os

import numpy as np
import pytest

import cuda_grid_estimation_b200 as cge
import oracle
from cuda_grid_estimation_b200 import api

from test_gpu_parity import check_against

pytestmark = pytest.mark.gpu


def _orders(obs):
    m = obs.mean()
    return {
        "as given": obs,
        "ascending": np.sort(obs),
        "descending": np.sort(obs)[::-1].copy(),
        "|x-mean| descending": obs[np.argsort(-np.abs(obs - m))].copy(),
        "|x-mean| ascending": obs[np.argsort(np.abs(obs - m))].copy(),
    }


@pytest.mark.parametrize("n", [50, 200, 250, 300, 1000])
@pytest.mark.parametrize("N", [101, 1024])
def test_product_mode_is_order_independent(ctx, n, N):
    """The reference multiplies in float64 in any order (inc/kernels.h:190-197).  Every ordering
    of the same sample must give the oracle's answer -- and all orderings the SAME bits, since the
    kernel sorts the sample before it picks its own order."""
    obs = oracle.readme_observations(n=n, seed=7)
    ref = oracle.grid_posterior(np.sort(obs), N, 18, 22, 1, 3, oracle.NUM_LOG)
    first = None
    # float32 factors carry ~2e-6 relative rounding each in the far tail (|exponent| ~ 50), which
    # accumulates as sqrt(n): 1e-4 on a cell holds to n ~ 300, 2e-4 is the bar at n = 1000
    post_rel = 1e-4 if n <= 300 else 2e-4
    for name, o in _orders(obs).items():
        res = ctx.grid_posterior(o, N, (18, 22), (1, 3))
        check_against(res, ref, post_rel=post_rel, what=f"n={n} N={N} order {name}")
        if first is None:
            first = res
        else:
            assert res.posterior.tobytes() == first.posterior.tobytes(), name
            assert res.expectations.tobytes() == first.expectations.tobytes(), name


def test_product_mode_rejects_more_observations_than_the_stage_holds(ctx):
    obs = oracle.readme_observations(n=16385, seed=1)
    with pytest.raises(cge.CgeError) as ei:
        ctx.grid_posterior(obs, 64, (18, 22), (1, 3))
    assert ei.value.code == api.ERR_BAD_ARG and "LOGSPACE" in str(ei.value)


@pytest.mark.parametrize("name", ["readme", "small", "max256"])
@pytest.mark.parametrize("variant", [0, api.VARIANT_UNFUSED])
def test_float64_posterior_vs_reference_cuda(ctx, golden_dir, name, variant):
    """cge_params.posterior_dtype = float64 (the reference's thrust::device_vector<double>,
    inc/cuda_functions.h:139-159): against the posterior the unmodified reference CUDA pipeline
    produced on a B200 (tests/golden/ref_cuda_*.npz)."""
    g = np.load(os.path.join(golden_dir, f"ref_cuda_{name}.npz"))
    N = int(g["N"])
    mu_r = (float(g["ranges"][0]), float(g["ranges"][1]))
    sg_r = (float(g["ranges"][2]), float(g["ranges"][3]))
    obs = g["observations"]
    res = ctx.grid_posterior(obs, N, mu_r, sg_r, variant=variant,
                             posterior_dtype=api.POSTERIOR_F64)
    assert res.posterior.dtype == np.float64
    P = res.posterior
    if "posterior" in g:
        Q = g["posterior"].reshape(N, N)
    else:  # the reference's posterior was kept for a sample of rows only
        P, Q = P[g["sample_rows"]], g["posterior_rows"]
    big = Q >= 1e-20 * g["marg_mu"].max() / N
    # <= 1e-6 relative is the float64 bar for the normalisation; the cells themselves carry the
    # float32 factors' rounding (~1e-5 at 256 observations), as the float32 output does
    assert (np.abs(P[big] - Q[big]) / Q[big]).max() <= 4e-5
    assert np.abs(res.expectations - g["expectations"]).max() <= 1e-6
    P = res.posterior
    # normalised in float64: the cells sum to one far closer than float32 cells could (Z itself is
    # a float64 sum of per-thread float32 partial sums, good to ~1e-9)
    assert abs(P.sum() - 1.0) <= 5e-9
    # and they are the float32 path's cells before its final rounding
    res32 = ctx.grid_posterior(obs, N, mu_r, sg_r, variant=variant)
    nz = P > 1e-30   # (below that the float32 copy is subnormal or flushed)
    assert (np.abs(res32.posterior.astype(np.float64)[nz] - P[nz]) / P[nz]).max() <= 1.3e-7


def test_float64_posterior_large_grid_and_device_buffer(ctx, readme_obs):
    import torch
    N = 2048
    dev = torch.device("cuda", 0)
    post = torch.empty((N, N), dtype=torch.float64, device=dev)
    res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), posterior=post,
                             posterior_dtype=api.POSTERIOR_F64)
    ref = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3))
    P = post.cpu().numpy()
    assert abs(P.sum() - 1.0) <= 5e-9
    nz = P > 1e-30   # (below that the float32 copy is subnormal or flushed)
    assert (np.abs(ref.posterior.astype(np.float64)[nz] - P[nz]) / P[nz]).max() <= 1.3e-7
    assert np.abs(res.expectations - ref.expectations).max() <= 1e-12
    np.testing.assert_allclose(P.sum(axis=1), res.marg_mu, rtol=2e-7, atol=1e-300)
    np.testing.assert_allclose(P.sum(axis=0), res.marg_sigma, rtol=2e-7, atol=1e-300)


@pytest.mark.parametrize("N,n_obs,mode,shards", [(1000, 50, 0, 2), (257, 50, 0, 3), (4100, 50, 0, 2),
                                                 (512, 2000, 1, 2), (130, 50, 0, 4)])
def test_shards_on_one_device_through_the_peer_exchange(ctx, N, n_obs, mode, shards):
    """cge_create_multi with the same device listed several times: the shards run back to back on
    one stream (fold A, fold B, finish A, finish B -- every flag is up before any poll), so the
    peer-exchange kernels run for real on a single GPU.  Cells must be bit-identical to the
    unsharded call, sums agree to ~1e-8 (a shard edge inside a 4-row patch changes one float32
    rounding of that patch's column sums); several calls exercise the step-parity double buffer."""
    obs = oracle.readme_observations(n=n_obs)
    full = ctx.grid_posterior(obs, N, (18, 22), (1, 3), mode=mode, variant=api.VARIANT_UNFUSED)
    with cge.MultiContext([0] * shards) as mc:
        for step in range(4):
            res = mc.grid_posterior(obs, N, (18, 22), (1, 3), mode=mode)
            assert res.posterior.tobytes() == full.posterior.tobytes(), f"step {step}"
            assert np.abs(res.expectations - full.expectations).max() < 1e-9
            assert np.abs(res.marg_sigma - full.marg_sigma).max() < 2e-8 * full.marg_sigma.max()
            assert np.abs(res.marg_mu - full.marg_mu).max() < 1e-9 * full.marg_mu.max()
        rows = [mc.shard(i)[1:3] for i in range(shards)]
        assert rows[0][0] == 0 and rows[-1][1] == N
        assert all(rows[i][1] == rows[i + 1][0] for i in range(shards - 1))
        # a different grid size on the same group (exchange blocks are re-sized)
        res = mc.grid_posterior(obs, N + 37, (18, 22), (1, 3), mode=mode)
        full2 = ctx.grid_posterior(obs, N + 37, (18, 22), (1, 3), mode=mode, variant=api.VARIANT_UNFUSED)
        assert res.posterior.tobytes() == full2.posterior.tobytes()
        # float64 output through the group
        res64 = mc.grid_posterior(obs, N, (18, 22), (1, 3), mode=mode, posterior_dtype=api.POSTERIOR_F64)
        assert abs(res64.posterior.sum() - 1.0) <= 5e-9


def test_multi_single_device_group_and_errors(ctx, readme_obs):
    with cge.MultiContext([0]) as mc:
        res = mc.grid_posterior(readme_obs, 300, (18, 22), (1, 3))
        full = ctx.grid_posterior(readme_obs, 300, (18, 22), (1, 3), variant=api.VARIANT_UNFUSED)
        assert res.posterior.tobytes() == full.posterior.tobytes()
        assert res.expectations.tobytes() == full.expectations.tobytes()
        with pytest.raises(cge.CgeError):
            mc.grid_posterior(readme_obs, 300, (18, 22), (-1, 3))
    with cge.MultiContext([0, 0]) as mc:
        with pytest.raises(cge.CgeError) as ei:   # an empty shard
            mc.grid_posterior(readme_obs, 1, (18, 22), (1, 3))
        assert ei.value.code == api.ERR_BAD_ARG
    with pytest.raises(cge.CgeError):
        cge.MultiContext([0, 10_000])


def test_step_is_three_launches(ctx, readme_obs):
    res = ctx.grid_posterior(readme_obs, 1024, (18, 22), (1, 3))
    assert res.launches == 3
    res = ctx.grid_posterior(readme_obs, 101, (18, 22), (1, 3))
    assert res.launches == 1
    res = ctx.grid_posterior(readme_obs, 101, (18, 22), (1, 3), variant=api.VARIANT_UNFUSED)
    assert res.launches == 3
    big = oracle.readme_observations(n=17000, seed=2)
    res = ctx.grid_posterior(big, 64, (19.9, 20.1), (1.9, 2.1), mode=cge.MODE_LOGSPACE)
    assert res.launches == 4     # + the staging of > 16384 observations


@pytest.mark.parametrize("N,dtype", [(4100, "f32"), (3001, "f64")])
def test_host_posterior_is_copied_in_row_chunks(ctx, readme_obs, N, dtype):
    """A posterior that goes to host memory and is larger than 64 MB is scaled and copied in row
    chunks (the finish kernel scales the first, scale_chunk_kernel the others, the copies run on a
    second stream): same bits as the posterior left on the device, N odd-sized so the chunks are
    not 16-byte aligned in the float32 case."""
    import torch
    f64 = dtype == "f64"
    kw = dict(posterior_dtype=api.POSTERIOR_F64) if f64 else {}
    host = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), **kw)   # posterior -> numpy
    assert host.launches > 3   # likelihood, fold, finish + one scale launch per further chunk
    dev = torch.empty((N, N), dtype=torch.float64 if f64 else torch.float32, device="cuda:0")
    res = ctx.grid_posterior(readme_obs, N, (18, 22), (1, 3), posterior=dev, **kw)
    assert res.launches == 3
    assert dev.cpu().numpy().tobytes() == host.posterior.tobytes()
    assert res.expectations.tobytes() == host.expectations.tobytes()
    assert abs(host.posterior.astype(np.float64).sum() - 1.0) <= 1e-6


@pytest.mark.parametrize("variant,path", [(0, "slope8x7"), (api.VARIANT_SLOPE_WIDE, "slope8"),
                                           (api.VARIANT_SLOPE, "slope")])
@pytest.mark.parametrize("N,shards", [(1100, 3), (1031, 4)])
def test_shards_take_the_neighbour_row_policies(ctx, readme_obs, variant, path, N, shards):
    """Row shards whose edges are multiples of neither 6 nor 7 rows, on a grid fine enough for the
    8-column policies (the benchmark's path): the tall policy's 7-row groups and the others' 6-row
    groups are anchored at GLOBAL rows, so every shard computes bit-identical cells to the
    unsharded call; N = 1031 is not a multiple of 8 either (scalar store path at the last strip).
    float64 output goes through the same kernels."""
    mu_r, sg_r = (19.6, 20.4), (1.6, 2.4)
    full = ctx.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
    assert ctx.scale_info()["product_path"] == path
    check_against(full, oracle.grid_posterior(readme_obs, N, *mu_r, *sg_r, oracle.NUM_F64),
                  what=f"{path} N={N}")
    with cge.MultiContext([0] * shards) as mc:
        for step in range(2):
            res = mc.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant)
            assert res.posterior.tobytes() == full.posterior.tobytes(), f"step {step}"
            assert np.abs(res.expectations - full.expectations).max() < 1e-9
            assert np.abs(res.marg_sigma - full.marg_sigma).max() < 2e-8 * full.marg_sigma.max()
            assert np.abs(res.marg_mu - full.marg_mu).max() < 1e-9 * full.marg_mu.max()
        res64 = mc.grid_posterior(readme_obs, N, mu_r, sg_r, variant=variant,
                                  posterior_dtype=api.POSTERIOR_F64)
        assert abs(res64.posterior.sum() - 1.0) <= 5e-9
        nz = res64.posterior > 1e-30
        assert (np.abs(full.posterior.astype(np.float64)[nz] - res64.posterior[nz]) /
                res64.posterior[nz]).max() <= 1.3e-7
