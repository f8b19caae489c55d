"""CPU: pin the oracle against every known-answer vector the reference's own tests hold for the
hot path (cpp/grid-algorithm/src/tests.cu, inc/utils.h:16-30), against the reference's Python
statement of the algorithm (golden fixture made by tests/golden/make_golden.py) and against the
end-to-end values established in SURVEY section 8(c)."""
import numpy as np
import pytest

import oracle


# ---- src/tests.cu:49-67 Linspaces -------------------------------------------------------
def test_kat_linspace():
    assert np.array_equal(oracle.linspace(3, 1.0, 2.0), np.array([1.0, 1.5, 2.0], np.float32))


# ---- inc/utils.h:16-30 checkArrays: exact f32 values of the README ranges ---------------
def test_kat_check_arrays_values():
    mu = oracle.linspace(101, 18.0, 22.0)
    sg = oracle.linspace(101, 1.0, 3.0)
    assert mu[50] == np.float32(20.0) and mu[49] == np.float32(19.96)
    assert sg[1] == np.float32(1.02) and sg[0] == np.float32(1.0)
    assert np.array_equal(mu, oracle.numpy_linspace(101, 18.0, 22.0))
    assert np.array_equal(sg, oracle.numpy_linspace(101, 1.0, 3.0))


# ---- src/tests.cu:70-108 GenerateGrids: layout mu slowest, sigma middle, obs fastest -----
def test_kat_meshgrid_layout():
    gx, gy, gz = oracle.meshgrid3(oracle.linspace(3, 1, 3), oracle.linspace(3, 1, 3),
                                  oracle.linspace(2, 4, 5))
    assert gx.tolist() == [1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3]
    assert gy.tolist() == [1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 3, 3]
    assert gz.tolist() == [4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5, 4, 5]


# ---- src/tests.cu:111-135 ComputeDensities ----------------------------------------------
def test_kat_densities():
    d = oracle.densities(oracle.linspace(3, 20, 22), oracle.linspace(3, 1, 3),
                         oracle.linspace(3, 20, 21))
    assert d == pytest.approx([0.39894228, 0.19333406, 0.12579441], abs=1e-5)
    assert d.dtype == np.float32


# ---- src/tests.cu:163-194 ComputeLikes --------------------------------------------------
def test_kat_row_products():
    likes = oracle.row_products(oracle.linspace(10, 1.0, 10.0), 2, 5)
    assert likes.tolist() == [120.0, 30240.0]


def test_row_product_tree_order_matches_reference_strides():
    """inc/kernels.h:190-197: strides 128..1 with guard t+stride<cols.  For 50 columns the
    first productive stride is 32; compare with an explicit emulation of the strided update."""
    rng = np.random.default_rng(0)
    dens = rng.uniform(0.5, 1.5, size=(7, 50)).astype(np.float32)
    got = oracle.row_products(dens, 7, 50)
    work = dens.astype(np.float64)
    stride = 128
    while stride > 0:
        t = np.arange(stride)
        t = t[t + stride < 50]
        work[:, t] = work[:, t] * work[:, t + stride]
        stride >>= 1
    assert np.array_equal(got, work[:, 0])


# ---- src/tests.cu:196-209 ComputePosterior ----------------------------------------------
def test_kat_posterior():
    post, z = oracle.posterior(np.array([1.0, 2.0, 3.0]))
    assert z == 6.0
    assert post[0] == pytest.approx(1 / 6, abs=1e-5)
    assert post[1] == pytest.approx(2 / 6, abs=1e-5)
    assert post[2] == 0.5


# ---- src/tests.cu:211-245 ComputeExpectations -------------------------------------------
def test_kat_expectations():
    post = np.full((3, 3), 0.1)
    post[1, 1] = 0.2
    mm = oracle.marginalize(post, axis=1)
    ms = oracle.marginalize(post, axis=0)
    assert mm == pytest.approx([0.3, 0.4, 0.3], abs=1e-15)
    assert ms == pytest.approx([0.3, 0.4, 0.3], abs=1e-15)
    assert oracle.expectation(mm, oracle.linspace(3, 1.0, 3.0)) == 2.0
    assert oracle.expectation(ms, oracle.linspace(3, 4.0, 6.0)) == 5.0


def test_marginalize_axes():
    """inc/wrappers.h:169-170: axis=1 -> row sums (mu marginal), axis=0 -> column sums."""
    m = np.arange(12, dtype=np.float64).reshape(3, 4)
    assert oracle.marginalize(m, 1).tolist() == m.sum(axis=1).tolist()
    assert oracle.marginalize(m, 0).tolist() == m.sum(axis=0).tolist()


# ---- end to end: SURVEY section 8(c) values ---------------------------------------------
def test_readme_observations_recipe(readme_obs):
    assert readme_obs[:5] == pytest.approx([20.99343, 19.72347, 21.295378, 23.04606, 19.531693],
                                           abs=1e-5)
    assert readme_obs.dtype == np.float32 and readme_obs.size == 50


def test_end_to_end_values(readme_obs):
    f64 = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    assert f64["expectations"][0] == pytest.approx(19.54905247652777, abs=1e-11)
    assert f64["expectations"][1] == pytest.approx(1.9167625043057062, abs=1e-11)
    ref = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_REF)
    # SURVEY's emulation of the reference numerics: 19.549052474669377, 1.916762507106191
    # (its f32 exp/pow came from numpy, ours from glibc; they agree to ~1e-9)
    assert ref["expectations"][0] == pytest.approx(19.549052474669377, abs=5e-9)
    assert ref["expectations"][1] == pytest.approx(1.916762507106191, abs=5e-9)
    log = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_LOG)
    assert np.abs(log["expectations"] - f64["expectations"]).max() < 1e-12
    likes = oracle.grid_likes(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    assert likes.max() == pytest.approx(7.01e-45, rel=2e-3)
    assert likes.min() == pytest.approx(5.27e-123, rel=2e-3)
    big = f64["posterior"] > 1e-12 * f64["posterior"].max()
    rel = np.abs(ref["posterior"][big] - f64["posterior"][big]) / f64["posterior"][big]
    assert rel.max() < 1e-5  # SURVEY measured 3.1e-6


def test_numpy_and_closed_form_cross_checks(readme_obs):
    f64 = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    npy = oracle.numpy_estimator_cpp_axes(readme_obs, 101, 18, 22, 1, 3)
    cf = oracle.closed_form_posterior(readme_obs, 101, 18, 22, 1, 3)
    for other in (npy, cf):
        assert np.abs(other["expectations"] - f64["expectations"]).max() < 1e-12
        assert np.abs(other["posterior"] - f64["posterior"]).max() < 1e-15
        assert np.abs(other["marg_mu"] - f64["marg_mu"]).max() < 1e-15
        assert np.abs(other["marg_sigma"] - f64["marg_sigma"]).max() < 1e-15


# ---- the reference's own Python implementation (imported in the build container) --------
def test_against_reference_numpy_estimator(golden_dir, readme_obs):
    g = np.load(f"{golden_dir}/numpy_estimator_readme.npz")
    assert np.array_equal(g["observations"], readme_obs)
    # the notebook prints axis-swapped expectations; both readings are pinned
    assert g["printed"] == pytest.approx([19.83352500577782, 1.7745262286551085], abs=1e-12)
    # numpy_estimator uses np.linspace ranges (differ from the C++ formula by <= 1.9e-6) and
    # 'xy' meshgrid: posterior[sigma_i, mu_j].  scipy's norm.pdf standardises in the inputs'
    # float32 and evaluates the density in float64 -- the same split the C++ has
    # (inc/kernels.h:111).  Same algorithm on its own ranges:
    mu32, sg32, x32 = g["mu_range"], g["sigma_range"], readme_obs
    mu, sg = mu32.astype(np.float64), sg32.astype(np.float64)
    z = ((x32[None, None, :] - mu32[:, None, None]) / sg32[None, :, None]).astype(np.float64)
    likes = (np.exp(-z * z / 2.0) / np.sqrt(2 * np.pi) / sg[None, :, None]).prod(axis=2)
    post = likes / likes.sum()
    ref_post = g["posterior_sigma_mu"].T  # -> [mu, sigma], the C++ layout
    big = ref_post > 1e-20 * ref_post.max()
    assert (np.abs(post[big] - ref_post[big]) / ref_post[big]).max() < 1e-14
    assert g["expectations_axis_correct"] == pytest.approx(
        [np.sum(mu * post.sum(axis=1)), np.sum(sg * post.sum(axis=0))], abs=1e-12)
    # and with the C++ ranges the oracle lands within the range-formula difference
    f64 = oracle.grid_posterior(readme_obs, 101, 18, 22, 1, 3, oracle.NUM_F64)
    assert np.abs(f64["expectations"] - g["expectations_axis_correct"]).max() < 1e-7


def test_oracle_sizes_and_edges():
    obs = oracle.readme_observations()
    # N = 2, 3, not containing the MLE, n = 1
    r = oracle.grid_posterior(obs, 3, 25, 30, 1, 3, oracle.NUM_F64)
    assert not r["degenerate"] and abs(r["posterior"].sum() - 1) < 1e-12
    r1 = oracle.grid_posterior(obs[:1], 2, 18, 22, 1, 3, oracle.NUM_REF)
    assert abs(r1["posterior"].sum() - 1) < 1e-12
    # many observations: the product underflows even in f64, log-space does not
    big = oracle.readme_observations(n=2000)
    assert oracle.grid_posterior(big, 16, 18, 22, 1, 3, oracle.NUM_F64)["degenerate"]
    lg = oracle.grid_posterior(big, 16, 18, 22, 1, 3, oracle.NUM_LOG)
    assert not lg["degenerate"] and abs(lg["posterior"].sum() - 1) < 1e-12
