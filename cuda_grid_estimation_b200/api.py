"""ctypes binding of libcge.so (include/cge.h) and the Python-side host helpers.

The arithmetic all lives in the CUDA library; this module only marshals pointers.  There is
no CPU fallback: if libcge.so is missing or no sm_100 device is present, calls raise.

Reference mapping (paths relative to cpp/grid-algorithm/ of nablabits/cuda-grid-estimation):
``Context.grid_posterior`` replaces the call sequence src/grid.cu:79-127; the ``linspace`` ...
``expectation`` methods are the stage functions inc/cuda_functions.h:47-221 and
inc/wrappers.h:95-121.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _build

MODE_PRODUCT_F32 = 0
MODE_LOGSPACE = 1
VARIANT_DEFAULT = 0
VARIANT_MUFU_ONLY = 1
VARIANT_PAIRED = 3
VARIANT_HYBRID = 4
VARIANT_LOG_F64 = 5
VARIANT_PAIRED3 = 8
VARIANT_SLOPE = 9
VARIANT_SLOPE_WIDE = 10
VARIANT_SLOPE_TALL = 11
VARIANT_UNFUSED = 7

OK, ERR_BAD_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_DEGENERATE, ERR_PEER, ERR_SIZE = range(7)
POSTERIOR_F32 = 0
POSTERIOR_F64 = 1

# every symbol include/cge.h declares (tests check the library exports each one)
ABI_SYMBOLS = (
    "cge_create", "cge_destroy", "cge_last_error", "cge_version", "cge_device_sm_count",
    "cge_grid_posterior", "cge_likelihood_partials", "cge_partials", "cge_finalize",
    "cge_peer_local_handle", "cge_peer_connect", "cge_peer_disconnect",
    "cge_results", "cge_read_results", "cge_posterior_dev", "cge_last_timing",
    "cge_profile_begin", "cge_profile_end", "cge_launch_count", "cge_linspace", "cge_meshgrid3", "cge_densities", "cge_densities_grid",
    "cge_row_products_rows", "cge_marginalize_rows", "cge_row_products",
    "cge_normalize", "cge_marginalize", "cge_expectation", "cge_microbench",
    "cge_generate_normal", "cge_quantiles", "cge_set_prior", "cge_scale_info",
    "cge_grid_posterior_async",
    "cge_create_multi", "cge_destroy_multi", "cge_multi_size", "cge_grid_posterior_multi",
    "cge_multi_shard", "cge_multi_step_async", "cge_multi_synchronize",
)


class CgeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libcge error {code}: {message}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("n_obs", C.c_int32), ("grid_size", C.c_int32), ("mu_start", C.c_float),
                ("mu_end", C.c_float), ("sigma_start", C.c_float), ("sigma_end", C.c_float),
                ("mode", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("variant", C.c_int32), ("posterior_dtype", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("setup_ms", C.c_float), ("likelihood_ms", C.c_float), ("reduce_ms", C.c_float),
                ("scale_ms", C.c_float), ("total_ms", C.c_float), ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float)]

    def as_dict(self):
        return {k: float(getattr(self, k)) for k, _ in self._fields_}


_lib: Optional[C.CDLL] = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load cuda_grid_estimation_b200/libcge.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CGE_LIBRARY") or _build.LIB_PATH   # (A/B builds for tuning runs)
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(f"{path} is missing; run __graft_entry__.build()")
        _build.build()
    L = C.CDLL(path)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    PP = C.POINTER(Params)
    sig = {
        "cge_create": [i32, C.POINTER(vp)],
        "cge_destroy": [vp],
        "cge_version": [],
        "cge_device_sm_count": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "cge_grid_posterior": [vp, PP, vp, vp, vp, vp, vp, C.POINTER(Timing)],
        "cge_likelihood_partials": [vp, PP, vp, vp, vp],
        "cge_partials": [vp, C.POINTER(vp), C.POINTER(sz)],
        "cge_finalize": [vp, PP, vp, vp],
        "cge_peer_local_handle": [vp, vp, i32],
        "cge_peer_connect": [vp, i32, i32, vp],
        "cge_peer_disconnect": [vp],
        "cge_results": [vp, C.POINTER(vp), C.POINTER(sz)],
        "cge_read_results": [vp, PP, vp, vp, vp, C.POINTER(C.c_double), vp],
        "cge_posterior_dev": [vp, C.POINTER(vp), C.POINTER(sz)],
        "cge_last_timing": [vp, C.POINTER(Timing)],
        "cge_launch_count": [vp, C.POINTER(i32)],
        "cge_profile_begin": [vp, i32],
        "cge_profile_end": [vp, C.POINTER(Timing), C.POINTER(i32)],
        "cge_linspace": [vp, vp, i32, f32, f32],
        "cge_meshgrid3": [vp, vp, vp, vp, vp, vp, vp, i32, i32],
        "cge_densities": [vp, vp, vp, vp, vp, sz],
        "cge_densities_grid": [vp, vp, vp, vp, vp, i32, i32],
        "cge_row_products_rows": [vp, vp, vp, i32, i32],
        "cge_marginalize_rows": [vp, vp, vp, i32, i32, i32],
        "cge_row_products": [vp, vp, vp, sz, sz],
        "cge_normalize": [vp, vp, vp, sz],
        "cge_marginalize": [vp, vp, vp, i32, i32, i32],
        "cge_expectation": [vp, vp, vp, i32, C.POINTER(C.c_double)],
        "cge_microbench": [vp, i32, C.POINTER(C.c_double)],
        "cge_generate_normal": [vp, vp, i32, f32, f32, C.c_ulonglong, i32],
        "cge_quantiles": [vp, vp, i32, vp, i32, vp, vp],
        "cge_set_prior": [vp, vp, vp, i32, vp],
        "cge_scale_info": [vp, vp, i32, vp],
        "cge_grid_posterior_async": [vp, vp, vp, vp, vp],
        "cge_create_multi": [i32, C.POINTER(i32), C.POINTER(vp)],
        "cge_destroy_multi": [vp],
        "cge_multi_size": [vp, C.POINTER(i32)],
        "cge_grid_posterior_multi": [vp, PP, vp, vp, vp, vp, vp, C.POINTER(Timing)],
        "cge_multi_shard": [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(vp)],
        "cge_multi_step_async": [vp, PP, C.POINTER(vp)],
        "cge_multi_synchronize": [vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.cge_last_error.argtypes = []
    L.cge_last_error.restype = C.c_char_p
    _lib = L
    return L


def _check(rc: int):
    if rc != OK:
        raise CgeError(rc, load_library().cge_last_error().decode("utf-8", "replace"))


def make_params(n_obs, grid_size, mu_range, sigma_range, mode=MODE_PRODUCT_F32, rows=None,
                variant=0, posterior_dtype=POSTERIOR_F32) -> Params:
    rb, re = (0, 0) if rows is None else (int(rows[0]), int(rows[1]))
    return Params(int(n_obs), int(grid_size), float(mu_range[0]), float(mu_range[1]),
                  float(sigma_range[0]), float(sigma_range[1]), int(mode), rb, re, int(variant),
                  int(posterior_dtype))


@dataclass
class GridResult:
    expectations: np.ndarray            # [E[mu], E[sigma]]
    marg_mu: np.ndarray                 # float64 [N]
    marg_sigma: np.ndarray              # float64 [N]
    posterior: Optional[np.ndarray]     # float32 [rows, N] or None (left on the device)
    timing: dict
    launches: int


def _ptr(a) -> Optional[int]:
    """Raw address of a numpy array or a torch tensor (None passes through)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())


def _is_cuda_tensor(t) -> bool:
    return (t is not None and not isinstance(t, (str, np.ndarray)) and
            bool(getattr(t, "is_cuda", False)))


class Context:
    """One libcge context = one device, one stream, grow-only scratch.  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.cge_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cge_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- device info / peaks --------------------------------------------------------
    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        _check(self._lib.cge_device_sm_count(self._h, C.byref(sm), C.byref(ma), C.byref(mi)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value))

    def microbench(self, kind: int) -> float:
        """Measured pipe peak, lane-ops/s.  0 MUFU ex2, 1 FFMA, 2 DFMA, 3 FFMA2 (f32x2)."""
        out = C.c_double()
        _check(self._lib.cge_microbench(self._h, int(kind), C.byref(out)))
        return out.value

    # ---- fused path, synchronous, host or device buffers ---------------------------------
    def grid_posterior(self, obs, grid_size, mu_range, sigma_range, mode=MODE_PRODUCT_F32,
                       rows=None, posterior="host", want_timing=True, variant=0,
                       posterior_dtype=POSTERIOR_F32) -> GridResult:
        """The whole hot path in one call (replaces src/grid.cu:79-127).

        obs        numpy float32 array (host) or CUDA torch tensor (device)
        posterior  "host": return it as numpy; "device"/None: leave it in the context's device
                   buffer; or a CUDA torch tensor / numpy array to write into.
        posterior_dtype  POSTERIOR_F32 (default) or POSTERIOR_F64 (the reference's type)

        The call runs on the context's own stream.  CUDA tensors passed in are first ordered
        against torch's current stream (the work that produced them is waited for), and the call
        returns only after its own work is complete, so torch may use the outputs at once.
        """
        if isinstance(obs, np.ndarray):
            obs = np.ascontiguousarray(obs, np.float32)
        n = int(obs.shape[0])
        p = make_params(n, grid_size, mu_range, sigma_range, mode, rows, variant, posterior_dtype)
        nrows = grid_size if rows is None else rows[1] - rows[0]
        np_dtype = np.float64 if posterior_dtype == POSTERIOR_F64 else np.float32
        post_arr = None
        if isinstance(posterior, str) and posterior == "host":
            post_arr = np.empty((nrows, grid_size), np_dtype)
            post_ptr = post_arr.ctypes.data
        elif posterior is None or (isinstance(posterior, str) and posterior == "device"):
            post_ptr = None
        else:
            post_ptr = _ptr(posterior)
            post_arr = posterior if isinstance(posterior, np.ndarray) else None
        if any(_is_cuda_tensor(t) for t in (obs, posterior)):
            import torch
            torch.cuda.current_stream(self.device).synchronize()
        mm = np.empty(grid_size, np.float64)
        ms = np.empty(grid_size, np.float64)
        ex = np.empty(2, np.float64)
        tm = Timing()
        _check(self._lib.cge_grid_posterior(self._h, C.byref(p), _ptr(obs), post_ptr, mm.ctypes.data,
                                            ms.ctypes.data, ex.ctypes.data,
                                            C.byref(tm) if want_timing else None))
        return GridResult(ex, mm, ms, post_arr, tm.as_dict() if want_timing else {},
                          self.launch_count())

    def grid_posterior_async(self, params: Params, obs_dev, posterior_dev, stream=None):
        """Enqueue the whole step on `stream` (device buffers, whole grid); read_results waits."""
        _check(self._lib.cge_grid_posterior_async(self._h, C.byref(params), _ptr(obs_dev),
                                                  _ptr(posterior_dev), stream))

    # ---- phase API (device pointers, asynchronous on `stream`) -------------------------
    def likelihood_partials(self, params: Params, obs_dev, posterior_dev, stream=None):
        _check(self._lib.cge_likelihood_partials(self._h, C.byref(params), _ptr(obs_dev),
                                                 _ptr(posterior_dev), stream))

    def finalize(self, params: Params, posterior_dev, stream=None):
        _check(self._lib.cge_finalize(self._h, C.byref(params), _ptr(posterior_dev), stream))

    # ---- peer exchange over CUDA IPC -------------------------------------------------------
    IPC_HANDLE_BYTES = 64

    def peer_local_handle(self, max_grid_size: int) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        _check(self._lib.cge_peer_local_handle(self._h, buf, int(max_grid_size)))
        return bytes(buf.raw)

    def peer_connect(self, rank: int, world: int, handles: bytes):
        assert len(handles) == world * self.IPC_HANDLE_BYTES
        _check(self._lib.cge_peer_connect(self._h, int(rank), int(world), handles))

    def peer_disconnect(self):
        _check(self._lib.cge_peer_disconnect(self._h))

    def partials_ptr(self):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _check(self._lib.cge_partials(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def results_ptr(self):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _check(self._lib.cge_results(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def posterior_ptr(self):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _check(self._lib.cge_posterior_dev(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def read_results(self, params: Params, stream=None):
        N = params.grid_size
        mm = np.empty(N, np.float64)
        ms = np.empty(N, np.float64)
        ex = np.empty(2, np.float64)
        z = C.c_double()
        _check(self._lib.cge_read_results(self._h, C.byref(params), ex.ctypes.data, mm.ctypes.data,
                                          ms.ctypes.data, C.byref(z), stream))
        return ex, mm, ms, z.value

    def profile_begin(self, max_steps: int):
        _check(self._lib.cge_profile_begin(self._h, int(max_steps)))

    def profile_end(self):
        """Mean per-phase device times (ms) over the captured steps, and how many there were."""
        tm, n = Timing(), C.c_int()
        _check(self._lib.cge_profile_end(self._h, C.byref(tm), C.byref(n)))
        return tm.as_dict(), n.value

    def launch_count(self) -> int:
        n = C.c_int()
        _check(self._lib.cge_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- stage API (device / managed pointers; torch CUDA tensors here) -----------------
    def linspace(self, out_dev, size, start, end):
        _check(self._lib.cge_linspace(self._h, _ptr(out_dev), int(size), float(start), float(end)))

    def meshgrid3(self, vx, vy, vz, gx, gy, gz, nxy, nz):
        _check(self._lib.cge_meshgrid3(self._h, _ptr(vx), _ptr(vy), _ptr(vz), _ptr(gx), _ptr(gy),
                                       _ptr(gz), int(nxy), int(nz)))

    def densities(self, dens, gx, gy, gz, count):
        _check(self._lib.cge_densities(self._h, _ptr(dens), _ptr(gx), _ptr(gy), _ptr(gz), int(count)))

    def row_products(self, dens, likes, densities_size, likes_size):
        _check(self._lib.cge_row_products(self._h, _ptr(dens), _ptr(likes), int(densities_size),
                                          int(likes_size)))

    def normalize(self, likes, post, size):
        _check(self._lib.cge_normalize(self._h, _ptr(likes), _ptr(post), int(size)))

    def marginalize(self, marginal, matrix, rows, cols, axis):
        _check(self._lib.cge_marginalize(self._h, _ptr(marginal), _ptr(matrix), int(rows), int(cols),
                                         int(axis)))

    def expectation(self, marginal, vector, size) -> float:
        out = C.c_double()
        _check(self._lib.cge_expectation(self._h, _ptr(marginal), _ptr(vector), int(size),
                                         C.byref(out)))
        return out.value


    # -- methods appended for the "next" rows (bound late to keep Context readable) --


def _generate_normal(self, out_dev, n, mu, sigma, seed=42, zero_state=False):
    """Reference observation generator (cuRAND XORWOW); out_dev is a CUDA float32 tensor."""
    _check(self._lib.cge_generate_normal(self._h, _ptr(out_dev), int(n), float(mu), float(sigma),
                                         int(seed), 1 if zero_state else 0))


def _quantiles(self, marginal_dev, n, probs, cdf_dev=None):
    """First index whose cumulative mass exceeds each p; marginal_dev is a CUDA float64 tensor."""
    probs = np.ascontiguousarray(probs, np.float64)
    idx = np.empty(probs.size, np.int32)
    _check(self._lib.cge_quantiles(self._h, _ptr(marginal_dev), int(n), probs.ctypes.data,
                                   probs.size, idx.ctypes.data, _ptr(cdf_dev)))
    return idx


def _set_prior(self, prior_mu=None, prior_sigma=None, grid_size=0, prior_full=None):
    """Non-flat prior for the following calls; all None restores the flat prior."""
    keep = [np.ascontiguousarray(a, np.float32) if isinstance(a, (np.ndarray, list, tuple)) else a
            for a in (prior_mu, prior_sigma)]
    _check(self._lib.cge_set_prior(self._h, _ptr(keep[0]), _ptr(keep[1]), int(grid_size),
                                   _ptr(prior_full)))


def _scale_info(self, stream=None) -> dict:
    """Scale statistics of the last likelihood call (see cge_scale_info)."""
    out = np.empty(15, np.float64)
    _check(self._lib.cge_scale_info(self._h, out.ctypes.data, 15, stream))
    keys = ("log2s", "shift", "m2", "xbar", "ss", "smin", "mu_c", "s_c", "umax", "paired_umax",
            "paired3_umax", "ustep", "slope_eps", "slope_err_max", "wide")
    d = dict(zip(keys, out.tolist()))
    # product_path: "slope8" when the whole grid ran the 8-column wide policy; otherwise the
    # arithmetic of the most demanding column band (the one at the smallest sigma) -- bands at
    # larger sigma may run a cheaper policy (see band_paths)
    d["product_path"] = ("slope8x7" if d["wide"] == 2 else "slope8" if d["wide"] else
                         _product_path(d, d["umax"], d["slope_eps"]))
    return d


_K2 = 0.2402265069591007   # ln(2)^2 / 2


def slope_error_bound(umax: float, rstep: float, steps: float = 1.5) -> float:
    """Largest shared-slope term per factor (cge::ProductSlopeT::eligible): k2 U^2 eps, with
    eps = 2x(1 + 1.5x), x = steps * relative sigma step (1.5 columns for the 4-column policy, 3.5
    for the wide one), 1 % on top."""
    x = steps * rstep * 1.01
    return _K2 * umax * umax * 2.0 * x * (1.0 + 1.5 * x)


def _product_path(info: dict, umax: float, rstep: float) -> str:
    """Mirror of the device-side choice (cge::likelihood_cta): the first eligible of slope,
    paired (degree 2), paired3 (degree 3), hybrid."""
    if umax <= info["paired_umax"] and slope_error_bound(umax, rstep) <= info["slope_err_max"]:
        return "slope"
    if umax <= info["paired_umax"]:
        return "paired"
    return "paired3" if umax <= info["paired3_umax"] else "hybrid"


def band_columns(grid_size: int) -> int:
    """Sigma columns per warp column strip of the product kernels (32 lanes x 4 columns): the
    granularity at which the default kernel chooses its arithmetic (cge::strip_umax)."""
    return 128


def band_paths(info: dict, sigma: np.ndarray) -> list:
    """Which product policy each column band takes, from scale_info() and the sigma table:
    mirrors cge::strip_umax (|a_j| = 0.5*log2(e)/sigma_j^2 at the band's two ends) and
    cge::strip_sigma_step (the column step over the band's smallest sigma)."""
    if info.get("wide"):
        return ["slope8x7" if info["wide"] == 2 else "slope8"] * ((len(sigma) + 127) // 128)
    out = []
    n = len(sigma)
    band_cols = band_columns(n)
    ds = abs(float(sigma[-1]) - float(sigma[0])) / (n - 1) if n > 1 else 0.0
    for j0 in range(0, n, band_cols):
        j1 = min(j0 + band_cols, n) - 1
        s0, s1 = float(sigma[j0]), float(sigma[j1])
        a = max(0.5 * 1.4426950408889634 / s0 ** 2, 0.5 * 1.4426950408889634 / s1 ** 2)
        smin = min(abs(s0), abs(s1))
        eps = ds / smin if smin > 0 and s0 * s1 > 0 else 1e300
        out.append(_product_path(info, a * info["ustep"], eps))
    return out


Context.scale_info = _scale_info
Context.set_prior = _set_prior
Context.generate_normal = _generate_normal
Context.quantiles = _quantiles


class _DeviceView:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def device_tensor(ptr: int, count: int, dtype: str, device_index: int):
    """Wrap library-owned device memory as a torch tensor without copying."""
    import torch
    typestr = {"float64": "<f8", "float32": "<f4"}[dtype]
    return torch.as_tensor(_DeviceView(ptr, count, typestr), device=torch.device("cuda", device_index))


class MultiContext:
    """Several GPUs driven from this one process (cge_create_multi): the grid is row-sharded over
    `devices` and every call is one C entry with host buffers."""

    def __init__(self, devices):
        self._lib = load_library()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _check(self._lib.cge_create_multi(len(devices), devs, C.byref(h)))
        self._h = h
        self.devices = [int(d) for d in devices]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cge_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def grid_posterior(self, obs, grid_size, mu_range, sigma_range, mode=MODE_PRODUCT_F32,
                       posterior="host", want_timing=False, variant=0,
                       posterior_dtype=POSTERIOR_F32) -> GridResult:
        """cge_grid_posterior_multi: host observations in; expectations, marginals (and, with
        posterior="host" or a numpy array, the gathered posterior) out.  posterior=None leaves the
        shards on their devices (see shard())."""
        obs = np.ascontiguousarray(obs, np.float32)
        p = make_params(obs.shape[0], grid_size, mu_range, sigma_range, mode, None, variant,
                        posterior_dtype)
        np_dtype = np.float64 if posterior_dtype == POSTERIOR_F64 else np.float32
        post_arr = None
        if isinstance(posterior, str) and posterior == "host":
            post_arr = np.empty((grid_size, grid_size), np_dtype)
        elif isinstance(posterior, np.ndarray):
            post_arr = posterior
        mm = np.empty(grid_size, np.float64)
        ms = np.empty(grid_size, np.float64)
        ex = np.empty(2, np.float64)
        tm = Timing()
        _check(self._lib.cge_grid_posterior_multi(
            self._h, C.byref(p), obs.ctypes.data, None if post_arr is None else post_arr.ctypes.data,
            mm.ctypes.data, ms.ctypes.data, ex.ctypes.data, C.byref(tm) if want_timing else None))
        return GridResult(ex, mm, ms, post_arr, tm.as_dict() if want_timing else {}, 3 * len(self.devices))

    def shard(self, index: int):
        """(device, row_begin, row_end, device pointer) of shard `index` of the last call."""
        dev, rb, re, ptr = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        _check(self._lib.cge_multi_shard(self._h, int(index), C.byref(dev), C.byref(rb), C.byref(re),
                                         C.byref(ptr)))
        return dev.value, rb.value, re.value, ptr.value

    def step_async(self, params: Params, obs_dev_ptrs):
        arr = (C.c_void_p * len(obs_dev_ptrs))(*[int(x) for x in obs_dev_ptrs])
        _check(self._lib.cge_multi_step_async(self._h, C.byref(params), arr))

    def synchronize(self):
        _check(self._lib.cge_multi_synchronize(self._h))
