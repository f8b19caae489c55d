"""CPU oracle for the grid-estimation hot path -- TEST INFRASTRUCTURE ONLY.

Nothing here is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  ``cuda_grid_estimation_b200`` never imports it.

Two restatements of the reference algorithm live here:

* ``oracle.c`` -> ``liboracle.so`` (ctypes, below): the stage functions a1..a12
  with the reference's numerics, plus a blocked whole-path oracle that never
  materialises the N*N*n tensors (float32-pdf/float64-tree-product "REF",
  all-float64 "F64", and float64 log-space "LOG").
* ``numpy_*`` functions (this file): small-case numpy statements used to
  cross-check the C code, including the closed-form sufficient-statistics
  likelihood that is independent of both.

Citations are relative to /root/reference/cpp/grid-algorithm/ unless they
start with ``notebooks/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NUM_REF, NUM_F64, NUM_LOG = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when the reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_linspace.argtypes = [_f32p, C.c_int, C.c_float, C.c_float]
        L.oracle_meshgrid3.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.oracle_normal_pdf.argtypes = [C.c_float, C.c_float, C.c_float]
        L.oracle_normal_pdf.restype = C.c_float
        L.oracle_densities.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t]
        L.oracle_row_products.argtypes = [_f64p, _f32p, C.c_size_t, C.c_int]
        L.oracle_posterior.argtypes = [_f64p, _f64p, C.c_size_t]
        L.oracle_posterior.restype = C.c_double
        L.oracle_marginalize.argtypes = [_f64p, _f64p, C.c_int, C.c_int, C.c_int]
        L.oracle_expectation.argtypes = [_f64p, _f32p, C.c_int]
        L.oracle_expectation.restype = C.c_double
        L.oracle_grid_likes.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_int, C.c_int, C.c_int, _f64p]
        L.oracle_grid_likes.restype = C.c_int
        L.oracle_grid_posterior.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_grid_posterior.restype = C.c_int
        L.oracle_bench_rows.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_int, C.c_int, C.c_int, _f64p, _f64p]
        L.oracle_bench_rows.restype = C.c_double
        _lib = L
    return _lib


# --------------------------------------------------------------------------
# stage functions (thin wrappers over oracle.c)
# --------------------------------------------------------------------------
def linspace(size: int, start: float, end: float) -> np.ndarray:
    """a1 -- inc/kernels.h:61-64."""
    out = np.empty(size, np.float32)
    lib().oracle_linspace(out, size, start, end)
    return out


def meshgrid3(vx, vy, vz):
    """a2 -- inc/kernels.h:88-98, layout of tests.cu:93-95."""
    vx, vy, vz = (np.ascontiguousarray(v, np.float32) for v in (vx, vy, vz))
    n = vx.size * vy.size * vz.size
    gx, gy, gz = (np.empty(n, np.float32) for _ in range(3))
    lib().oracle_meshgrid3(vx, vy, vz, gx, gy, gz, vx.size, vz.size)
    return gx, gy, gz


def normal_pdf(x: float, mu: float, sigma: float) -> float:
    """a3 -- inc/kernels.h:102-112."""
    return float(lib().oracle_normal_pdf(x, mu, sigma))


def densities(gx, gy, gz) -> np.ndarray:
    """a4 -- inc/kernels.h:134-140: pdf(x=gz, mu=gx, sigma=gy)."""
    gx, gy, gz = (np.ascontiguousarray(v, np.float32) for v in (gx, gy, gz))
    out = np.empty(gx.size, np.float32)
    lib().oracle_densities(out, gx, gy, gz, gx.size)
    return out


def row_products(dens, rows: int, cols: int) -> np.ndarray:
    """a8/a9 -- inc/kernels.h:190-204 (f64 pairwise tree over f32 densities)."""
    dens = np.ascontiguousarray(dens, np.float32)
    assert dens.size == rows * cols
    out = np.empty(rows, np.float64)
    lib().oracle_row_products(out, dens, rows, cols)
    return out


def posterior(likes) -> tuple[np.ndarray, float]:
    """a10 -- inc/cuda_functions.h:155-158: likes * (1 / sum(likes))."""
    likes = np.ascontiguousarray(likes, np.float64)
    out = np.empty_like(likes)
    z = lib().oracle_posterior(likes, out, likes.size)
    return out, float(z)


def marginalize(mat, axis: int) -> np.ndarray:
    """a11 -- inc/kernels.h:237-258; axis=1 row sums (mu), axis=0 column sums (sigma)."""
    mat = np.ascontiguousarray(mat, np.float64)
    rows, cols = mat.shape
    out = np.empty(rows if axis == 1 else cols, np.float64)
    lib().oracle_marginalize(out, mat, rows, cols, axis)
    return out


def expectation(marginal, vec) -> float:
    """a12 -- inc/cuda_functions.h:201-218."""
    marginal = np.ascontiguousarray(marginal, np.float64)
    vec = np.ascontiguousarray(vec, np.float32)
    return float(lib().oracle_expectation(marginal, vec, marginal.size))


# --------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------
def grid_likes(obs, N, mu0, mu1, sg0, sg1, numerics=NUM_REF, row_begin=0, row_end=None):
    """Unnormalised likelihoods (natural-log likelihoods for NUM_LOG) for a row range."""
    obs = np.ascontiguousarray(obs, np.float32)
    row_end = N if row_end is None else row_end
    out = np.empty((row_end - row_begin, N), np.float64)
    lib().oracle_grid_likes(obs, obs.size, N, mu0, mu1, sg0, sg1, numerics, row_begin, row_end, out)
    return out


def grid_posterior(obs, N, mu0, mu1, sg0, sg1, numerics=NUM_REF, want_posterior=True):
    """Whole path, src/grid.cu:79-127.  Returns dict(posterior, marg_mu, marg_sigma, expectations)."""
    obs = np.ascontiguousarray(obs, np.float32)
    post = np.empty((N, N), np.float64) if want_posterior else None
    mm = np.empty(N, np.float64)
    ms = np.empty(N, np.float64)
    ex = np.empty(2, np.float64)
    rc = lib().oracle_grid_posterior(
        obs, obs.size, N, mu0, mu1, sg0, sg1, numerics,
        post.ctypes.data if post is not None else None,
        mm.ctypes.data, ms.ctypes.data, ex.ctypes.data)
    if rc == 2:
        raise MemoryError("oracle_grid_posterior")
    return dict(posterior=post, marg_mu=mm, marg_sigma=ms, expectations=ex, degenerate=bool(rc),
                mu=linspace(N, mu0, mu1), sigma=linspace(N, sg0, sg1))


def bench_rows(obs, N, mu0, mu1, sg0, sg1, numerics, row_begin, row_end):
    """Timed-baseline kernel: returns (pdf_evals, row_sums, col_sums)."""
    obs = np.ascontiguousarray(obs, np.float32)
    rs = np.empty(row_end - row_begin, np.float64)
    cs = np.empty(N, np.float64)
    evals = lib().oracle_bench_rows(obs, obs.size, N, mu0, mu1, sg0, sg1, numerics,
                                    row_begin, row_end, rs, cs)
    return evals, rs, cs


# --------------------------------------------------------------------------
# numpy restatements (small cases; cross-checks of the C code)
# --------------------------------------------------------------------------
def numpy_linspace(size, start, end):
    """inc/kernels.h:61-64 in numpy float32, op by op."""
    start, end = np.float32(start), np.float32(end)
    idx = np.arange(size, dtype=np.int32).astype(np.float32)
    return (start + (idx * (end - start)) / np.float32(size - 1)).astype(np.float32)


def numpy_estimator_cpp_axes(obs, N, mu0, mu1, sg0, sg1):
    """The notebook algorithm (notebooks/quick-prototyping.ipynb cell 4) in float64 with the
    C++ axis convention posterior[i, j] = P(mu_i, sigma_j) (inc/kernels.h:92) and the C++
    range formula.  Materialises N*N*n -- small N only."""
    mu = numpy_linspace(N, mu0, mu1).astype(np.float64)
    sg = numpy_linspace(N, sg0, sg1).astype(np.float64)
    x = np.asarray(obs, np.float32).astype(np.float64)
    z = (x[None, None, :] - mu[:, None, None]) / sg[None, :, None]
    dens = np.exp(-0.5 * z * z) / (sg[None, :, None] * np.sqrt(2 * np.pi))
    likes = dens.prod(axis=2)
    post = likes / likes.sum()
    return dict(posterior=post, marg_mu=post.sum(axis=1), marg_sigma=post.sum(axis=0),
                expectations=np.array([np.sum(mu * post.sum(axis=1)), np.sum(sg * post.sum(axis=0))]),
                mu=mu.astype(np.float32), sigma=sg.astype(np.float32))


def closed_form_loglik(obs, N, mu0, mu1, sg0, sg1, rows=None):
    """Independent float64 check via sufficient statistics:
    sum_k (x_k-mu)^2 = SS + n (xbar-mu)^2.  Never used by the GPU path."""
    x = np.asarray(obs, np.float32).astype(np.float64)
    n = x.size
    xbar = x.mean()
    ss = np.sum((x - xbar) ** 2)
    mu = numpy_linspace(N, mu0, mu1).astype(np.float64)
    sg = numpy_linspace(N, sg0, sg1).astype(np.float64)
    if rows is not None:
        mu = mu[rows]
    s = ss + n * (xbar - mu) ** 2
    return -0.5 * s[:, None] / sg[None, :] ** 2 - n * np.log(sg[None, :] * np.sqrt(2 * np.pi))


def closed_form_posterior(obs, N, mu0, mu1, sg0, sg1):
    ll = closed_form_loglik(obs, N, mu0, mu1, sg0, sg1)
    w = np.exp(ll - ll.max())
    post = w / w.sum()
    mu = numpy_linspace(N, mu0, mu1).astype(np.float64)
    sg = numpy_linspace(N, sg0, sg1).astype(np.float64)
    mm, ms = post.sum(axis=1), post.sum(axis=0)
    return dict(posterior=post, marg_mu=mm, marg_sigma=ms,
                expectations=np.array([np.sum(mu * mm), np.sum(sg * ms)]), loglik_max=ll.max())


def readme_observations(n: int = 50, seed: int = 42, mu: float = 20.0, sigma: float = 2.0):
    """README.md:34-36 / notebook cell 3 for the legacy-seeded 50-obs workload;
    SURVEY section 8(d) config 5 (default_rng) for anything that is not n=50."""
    if n == 50:
        rs = np.random.RandomState(seed)
        return rs.normal(mu, sigma, n).astype(np.float32)
    return np.random.default_rng(seed).normal(mu, sigma, n).astype(np.float32)
