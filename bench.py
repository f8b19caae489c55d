#!/usr/bin/env python
"""bench.py -- headline benchmark of the grid likelihood -> posterior -> marginals -> expectations
hot path (BASELINE.json: "G pdf-evals/s ... at 1/2/4/8 B200").

    python bench.py --gpus 1 --steps K --warmup W             # our arm (CUDA, C ABI)
    python bench.py --impl reference --steps K --warmup W     # reference arm (CPU, host cores)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one full pass of the hot path over one batch of synthetic observations:
tables + rescale statistics, the fused likelihood kernel, the deterministic fold, (N>1: the sum of
N+2 float64 over ranks, by peer loads inside the results kernel or NCCL), marginals/expectations
and the in-place 1/Z scale.  The default
workload is BASELINE.json configs[2] -- 16384 x 16384 grid x 50 observations, float32 product
path -- the configuration the headline target (>= 60 % of MUFU peak on one GPU) is quoted on;
with --gpus N the same grid is row-sharded (strong scaling, as configs[2] states).
One JSON line goes to stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pdf_evals_per_s"
UNIT = "G pdf-evals/s"
MU_RANGE = (18.0, 22.0)
SIGMA_RANGE = (1.0, 3.0)
WORKLOADS = {
    # name: (grid_size, n_obs, mode)
    "readme": (101, 50, 0),
    "4096x50": (4096, 50, 0),
    "16384x50": (16384, 50, 0),
    "65536x50": (65536, 50, 0),
    "16384x10k-log": (16384, 10000, 1),
}
DEFAULT_WORKLOAD = "16384x50"
# dram__bytes_read.sum + dram__bytes_write.sum of the likelihood kernel, per launch, from the
# committed ncu --set full captures (profiles/r01_ncu_full_*.csv); keyed by (workload, n_gpus)
NCU_TRAFFIC = {
    ("16384x50", 1): 1.0768e9,        # profiles/r01_ncu_full_paired_16384x50.csv
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: everything libraries print there (the NCCL version banner,
# for one) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
_REAL_STDOUT = None


def quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def observations(n: int) -> np.ndarray:
    """Synthetic observations ~ N(20, 2): README.md:34-36 recipe for n=50 (legacy seed 42),
    default_rng(42) otherwise.  Stated here rather than imported
    from oracle/ so the product arm never touches the oracle package."""
    if n == 50:
        return np.random.RandomState(42).normal(20.0, 2.0, n).astype(np.float32)
    return np.random.default_rng(42).normal(20.0, 2.0, n).astype(np.float32)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


# ---------------------------------------------------------------------------------------
# clocks under load (NVML, sampled in a thread during the timed region)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period, self._stop, self._thr, self.ok = period_s, threading.Event(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of clocks
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------
def cpu_rate_sample(N, n, mode, target_seconds, max_rows=None):
    """Times the oracle (reference numerics: f32 pdf, f64 tree product; f64 log-space for the
    log workload) on a contiguous block of mu rows centred on the posterior peak, all host
    threads.  Returns (evals_per_s, seconds, rows, cores)."""
    import oracle
    numerics = oracle.NUM_LOG if mode == 1 else oracle.NUM_REF
    obs = observations(n)
    cores = int(oracle.lib().oracle_num_threads())
    centre = int((19.55 - MU_RANGE[0]) / (MU_RANGE[1] - MU_RANGE[0]) * (N - 1))
    probe_rows = max(1, min(N, max(cores, 4)))
    if N * n * probe_rows > 4e8:
        probe_rows = max(1, int(4e8 // (N * n)))

    def run(rows):
        rb = max(0, min(N - rows, centre - rows // 2))
        t0 = time.perf_counter()
        evals, _, _ = oracle.bench_rows(obs, N, *MU_RANGE, *SIGMA_RANGE, numerics, rb, rb + rows)
        return evals, time.perf_counter() - t0

    run(probe_rows)                       # warm up threads and caches
    ev, dt = run(probe_rows)
    rate = ev / dt
    rows = int(max(probe_rows, min(N if max_rows is None else max_rows,
                                   target_seconds * rate / (N * n))))
    ev, dt = run(rows)
    return ev / dt, dt, rows, cores


def reference_cuda_binary():
    """Context for the reference line, not part of its value: the reference's own CUDA program
    (src/grid.cu compiled unmodified into oracle/_ref/grid by oracle/Makefile) run once on this
    box's GPU.  It has one workload -- README, 101 x 101 x 50 with cuRAND observations -- and is
    not valid beyond 256, so it cannot serve any of the benchmark configurations."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "grid")
    if not os.path.exists(exe):
        return {"available": False, "why": "oracle/_ref/grid not built"}
    try:
        subprocess.run([exe], capture_output=True, text=True, timeout=120)  # context warm-up
        t0 = time.perf_counter()
        res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        wall = time.perf_counter() - t0
    except Exception as e:
        return {"available": False, "why": repr(e)}
    if res.returncode != 0:
        return {"available": False, "why": f"exit {res.returncode} (no GPU?)"}
    return {"available": True, "workload": "README: 101 x 101 grid x 50 cuRAND observations",
            "process_wall_ms": wall * 1e3, "pdf_evals": 101 * 101 * 50,
            "stdout": res.stdout.strip().splitlines()[-2:]}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation.  The reference's only CPU
    code is Python (numpy_estimator) and cannot travel to the GPU box, and its C++ is CUDA-only,
    so this times the oracle's C port of the same algorithm with all host threads
    (cpu_baseline.kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, n, mode = WORKLOADS[args.workload]
    import oracle
    numerics = oracle.NUM_LOG if mode == 1 else oracle.NUM_REF
    obs = observations(n)
    cores = int(oracle.lib().oracle_num_threads())
    # size one step to ~args.ref_step_seconds of CPU work
    rate, _, _, _ = cpu_rate_sample(N, n, mode, 1.0)
    rows = int(max(1, min(N, args.ref_step_seconds * rate / (N * n))))
    centre = int((19.55 - MU_RANGE[0]) / (MU_RANGE[1] - MU_RANGE[0]) * (N - 1))
    rb = max(0, min(N - rows, centre - rows // 2))
    steps, warmup = args.steps, max(args.warmup, 1)
    if args.steps_defaulted:
        steps, warmup = 10, 3
    for _ in range(warmup):
        oracle.bench_rows(obs, N, *MU_RANGE, *SIGMA_RANGE, numerics, rb, rb + rows)
    t0 = time.perf_counter()
    evals = 0.0
    for _ in range(steps):
        e, _, _ = oracle.bench_rows(obs, N, *MU_RANGE, *SIGMA_RANGE, numerics, rb, rb + rows)
        evals += e
    dt = time.perf_counter() - t0
    value = evals / dt / 1e9
    sample = (f"{rows} of {N} mu rows x {N} sigma x {n} obs per step "
              f"({rows * N * n:.3g} pdf-evals), rows centred on the posterior peak")
    ref_cuda = reference_cuda_binary()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if mode == 1 else "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_cuda_binary": ref_cuda,
    }
    emit(line)
    return 0


def workload_config(name, gpus):
    N, n, mode = WORKLOADS[name]
    return {
        "workload": f"{N}x{N} grid x {n} obs, " + ("log-space float64" if mode else "float32 product path"),
        "grid_size": N, "n_obs": n, "mode": "logspace" if mode else "product_f32",
        "mu_range": list(MU_RANGE), "sigma_range": list(SIGMA_RANGE),
        "observations": "N(20,2) seed 42 (README recipe)" if n == 50 else "N(20,2) default_rng(42)",
        "sharding": f"mu rows split over {gpus} GPU(s); one all-reduce of N+2 float64" if gpus > 1
                    else "single GPU",
        "posterior": "float32 [mu][sigma], device-resident",
        "l2": f"posterior written per step is {N * N * 4 / gpus / 1e6:.0f} MB per GPU "
              + ("(> 126 MB L2: no flush needed)" if N * N * 4 / gpus > 126e6 else
                 "(fits L2: a 256 MB L2 flush runs between timed steps, outside the per-step event "
                 "pairs that ms_per_step sums)"),
    }


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import cuda_grid_estimation_b200 as cge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, n, mode = WORKLOADS[args.workload]
    steps, warmup = args.steps, max(args.warmup, 3)
    ctx = cge.Context(local)
    info = ctx.device_info()
    obs_host = torch.from_numpy(observations(n)).pin_memory()
    obs_dev = obs_host.to(dev)
    sh = cge.ShardedGridPosterior(ctx, N, MU_RANGE, SIGMA_RANGE, n, mode, exchange=args.exchange)
    post = sh.alloc_posterior()
    cells_local = sh.rows_local * N
    need_flush = cells_local * 4 <= 126e6
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if need_flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        if flush_buf is not None:
            flush_buf.add_(1.0)
        sh.step(obs_dev, post)

    for _ in range(warmup):
        one_step()
    barrier()

    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # workloads that fit L2 get a flush between steps; it must not be part of the step time, so
    # those steps are bracketed one by one (events on the same stream) and their times summed
    step_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                for _ in range(steps)] if need_flush else None
    ctx.profile_begin(steps)
    sampler.start()
    barrier()
    ev0.record()
    for i in range(steps):
        if step_evs is None:
            sh.step(obs_dev, post)
        else:
            flush_buf.add_(1.0)
            step_evs[i][0].record()
            sh.step(obs_dev, post)
            step_evs[i][1].record()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    if step_evs is not None:
        elapsed_ms = sum(a.elapsed_time(b) for a, b in step_evs)
    phase, captured = ctx.profile_end()
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    launches_per_step = ctx.launch_count()
    ex, mm, ms_, z = sh.results()
    total_evals = float(N) * N * n
    value = total_evals * steps / (elapsed_ms * 1e-3) / 1e9

    # ---- end to end through the public API: host observations in, expectations+marginals out
    e2e_steps = max(3, min(steps, 50))
    h2d = obs_host.numel() * 4
    d2h = 16 + 2 * N * 8 + 8
    if world == 1:
        obs_np = obs_host.numpy()
        for _ in range(3):
            ctx.grid_posterior(obs_np, N, MU_RANGE, SIGMA_RANGE, mode=mode, posterior=post,
                               want_timing=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = ctx.grid_posterior(obs_np, N, MU_RANGE, SIGMA_RANGE, mode=mode, posterior=post,
                                   want_timing=False)
        e2e_s = time.perf_counter() - t0
        e2e_check = r.expectations.tolist()
        # the same call with the posterior ALSO copied to (pinned) host memory every step
        e2e_full = None
        if cells_local * 4 <= 4.5e9:
            pinned = torch.empty((N, N), dtype=torch.float32).pin_memory()
            host_np = pinned.numpy()
            full_steps = max(2, min(e2e_steps, 10))
            ctx.grid_posterior(obs_np, N, MU_RANGE, SIGMA_RANGE, mode=mode, posterior=host_np,
                               want_timing=False)
            t0 = time.perf_counter()
            for _ in range(full_steps):
                ctx.grid_posterior(obs_np, N, MU_RANGE, SIGMA_RANGE, mode=mode, posterior=host_np,
                                   want_timing=False)
            fs = time.perf_counter() - t0
            e2e_full = {"value": total_evals * full_steps / fs / 1e9, "unit": UNIT,
                        "ms_per_step": fs / full_steps * 1e3, "steps": full_steps,
                        "d2h_bytes_per_step": d2h + N * N * 4,
                        "note": "posterior (float32 N*N) copied to pinned host memory inside the "
                                "timed region as well; PCIe-bound"}
            del pinned, host_np
    else:
        e2e_full = None
        def e2e_once():
            obs_dev.copy_(obs_host, non_blocking=True)
            sh.step(obs_dev, post)
            return sh.results()
        for _ in range(3):
            e2e_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r = e2e_once()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e_check = r[0].tolist()
    e2e_value = total_evals * e2e_steps / e2e_s / 1e9

    # ---- rooflines (rank 0's kernels; every rank runs the same shapes)
    hbm_peak, hbm_src, peaks_doc = measured_peaks()
    out = None
    if rank == 0:
        ex2_peak = ctx.microbench(0)
        ffma_peak = ctx.microbench(1)
        dfma_peak = ctx.microbench(2)
        ffma2_peak = ctx.microbench(3)
        like_s = phase["likelihood_ms"] * 1e-3
        evals_per_launch = float(sh.rows_local) * N * n
        ach = evals_per_launch / like_s
        sinfo = ctx.scale_info()
        traffic = NCU_TRAFFIC.get((args.workload, world))
        if mode == 0:
            # executed op counts per pdf-eval of the path that ran (SASS of the inner loop):
            #   paired rows: 8 MUFU.EX2 + 27 packed (2 lane-ops each) per 16 evals
            #   hybrid     : 12 MUFU.EX2 + 8 + 8 packed (exponent, product) + 2 x 8 packed
            #                (polynomial) + 8 scalar per 16 evals
            #   paired, degree 3: 4 more packed per 16 evals
            path = sinfo["product_path"]
            mufu_pe, fma_pe = {"paired": (0.5, 54 / 16), "paired3": (0.5, 62 / 16)}.get(
                path, (0.75, (2 * 32 + 8) / 16))
            roof = {"bound": "mufu", "achieved": ach / 1e9, "peak": ex2_peak / 1e9, "unit": "Gop/s",
                    "frac": ach / ex2_peak, "traffic": traffic,
                    "kernel": "grid_likelihood_select_kernel<ProductPaired deg 2, ProductPaired deg 3, "
                              "ProductHybrid> -> " + path,
                    "algorithmic": "1 ex2 per pdf-eval (SURVEY 8(d)); evals per launch = rows*N*n",
                    "note": "frac > 1 is by design: the kernel takes " + str(mufu_pe) + " of its "
                            "exponentials per eval from MUFU.EX2 and forms the rest on the FMA pipe "
                            "(paired rows: e*2^u with a degree-2 polynomial; hybrid: degree-5 "
                            "polynomial), so neither pipe alone bounds it; `executed` gives the "
                            "utilisation of each pipe by the instructions actually issued, and "
                            "variant 1 (MUFU-only) is the plain MUFU-roofline kernel",
                    "executed": {"mufu_per_eval": mufu_pe, "fma_lane_ops_per_eval": fma_pe,
                                 "mufu_util": mufu_pe * ach / ex2_peak,
                                 "fma_util": fma_pe * ach / ffma_peak},
                    "umax": sinfo["umax"], "paired_umax": sinfo["paired_umax"],
                    "paired3_umax": sinfo["paired3_umax"],
                    "peak_source": "ex2.approx.ftz micro-benchmark in this run (cge_microbench 0)",
                    "peak_nominal": 16 * info["sm_count"] * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9}
        else:
            roof = {"bound": "fp32", "achieved": ach / 1e9, "peak": ffma_peak / 1e9, "unit": "Gop/s",
                    "frac": ach / ffma_peak, "traffic": traffic,
                    "kernel": "grid_likelihood_kernel<LogSpaceTiled<256>>",
                    "algorithmic": "1 FMA per pdf-eval; evals per launch = rows*N*n",
                    "executed": {"fma_lane_ops_per_eval": 1.125,
                                 "fma_util": 1.125 * ach / ffma_peak},
                    "peak_source": "FFMA micro-benchmark in this run (cge_microbench 1)"}
        scale_s = phase["scale_ms"] * 1e-3
        hbm = {"bound": "hbm", "kernel": "scale_posterior_kernel",
               "achieved": 8.0 * cells_local / scale_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": 8.0 * cells_local / scale_s / 1e9 / hbm_peak, "traffic": None,
               "algorithmic": "4 B read + 4 B written per cell", "peak_source": hbm_src}
        fp32 = {"bound": "fp32", "achieved": 4 * ach / 1e9, "peak": ffma_peak / 1e9,
                "unit": "G lane-instr/s", "frac": 4 * ach / ffma_peak,
                "algorithmic": "4 FP32 instr per pdf-eval (SURVEY 8(d)'s count for the product "
                               "path; see roofline.executed for what the kernel issues)"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": elapsed_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64" if mode else "f32",
            "data": "synthetic", "config": dict(workload_config(args.workload, world),
                           exchange=("none" if world == 1 else
                                     "peer loads over NVLink fused into the results kernel (CUDA IPC)"
                                     if sh.exchange == "p2p" else "NCCL all_reduce")),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "cge_grid_posterior (host obs in; E, marginals to host; posterior stays "
                           "on the device like the reference's device_vector)" if world == 1 else
                           "pinned obs -> device, ShardedGridPosterior.step, cge_read_results"},
            "e2e_posterior_to_host": e2e_full,
            "gpu_launches": launches_per_step * steps,
            "launches_per_step": launches_per_step,
            "roofline": roof, "roofline_hbm": hbm, "roofline_fp32": fp32,
            "phases_ms": {k: phase[k] for k in ("setup_ms", "likelihood_ms", "reduce_ms", "scale_ms", "total_ms")},
            "phase_steps_captured": captured,
            "pipe_peaks": {"mufu_ex2_Gops": ex2_peak / 1e9, "ffma_Gops": ffma_peak / 1e9,
                           "dfma_Gops": dfma_peak / 1e9, "ffma2_Glaneops": ffma2_peak / 1e9,
                           "sm_count": info["sm_count"]},
            "check": {"E_mu": float(ex[0]), "E_sigma": float(ex[1]), "Z": float(z), "e2e_E": e2e_check},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                rate, secs, rows, cores = cpu_rate_sample(N, n, mode, args.cpu_seconds)
                out["cpu_baseline"] = {
                    "value": rate / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{rows} of {N} mu rows x {N} sigma x {n} obs "
                              f"({rows * N * n:.3g} pdf-evals, {secs:.1f} s), oracle C port of the "
                              f"reference numerics, OpenMP over rows"}
            except Exception as e:  # the oracle is test infrastructure; never fail the GPU number on it
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                       "sample": f"unavailable: {e!r}"}
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if out is not None:
        emit(out)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", choices=["auto", "p2p", "nccl"], default="auto",
                    help="N>1: how the N+2 float64 partial vector is summed over ranks")
    args = ap.parse_args()
    quiet_stdout()
    args.steps_defaulted = args.steps is None
    if args.steps is None:
        args.steps = 100
    if args.warmup is None:
        args.warmup = 5
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
