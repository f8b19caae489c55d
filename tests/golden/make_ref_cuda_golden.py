"""Run the UNMODIFIED reference CUDA pipeline (oracle/_ref/ref_harness, built in the build
container from /root/reference by oracle/Makefile) on a B200 and store its outputs as golden
fixtures.  Run on the GPU box:

    gpurun -- python tests/golden/make_ref_cuda_golden.py      # writes gpurun_out/golden/*.npz

then copy gpurun_out/golden/ref_cuda_*.npz into tests/golden/ and commit them.  The reference is
valid only for grid_size <= 256 and n_obs <= 256 (inc/kernels.h:190-191,237-238).
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

import oracle  # noqa: E402  (observation recipes only)


def run_case(name, obs, N, ranges, out_dir, keep_full=True):
    obs = np.ascontiguousarray(obs, np.float32)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<ii", N, obs.size))
            f.write(np.asarray(ranges, np.float32).tobytes())
            f.write(obs.tobytes())
        res = subprocess.run([HARNESS, fin, fout], capture_output=True, text=True)
        print(res.stdout[-400:], res.stderr[-400:])
        res.check_returncode()
        raw = open(fout, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    head = take(np.float64, 4)
    mu, sigma = take(np.float32, N), take(np.float32, N)
    likes, post = take(np.float64, N * N), take(np.float64, N * N)
    marg_mu, marg_sigma = take(np.float64, N), take(np.float64, N)
    assert off == len(raw)
    data = dict(N=N, observations=obs, ranges=np.asarray(ranges, np.float32),
                expectations=head[:2], elapsed_ms=head[2], used_wrapper=head[3], mu=mu, sigma=sigma,
                marg_mu=marg_mu, marg_sigma=marg_sigma)
    if keep_full:
        data.update(likes=likes, posterior=post)
    else:  # big case: a few rows only
        rows = np.array([0, N // 4, N // 2 - 3, N // 2, N // 2 + 5, N - 1])
        data.update(sample_rows=rows, posterior_rows=post.reshape(N, N)[rows], likes_rows=likes.reshape(N, N)[rows])
    path = os.path.join(out_dir, f"ref_cuda_{name}.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, "E =", head[:2], "elapsed_ms", head[2])


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    # README workload through the reference's own wrappers (integral ranges, N=101, n=50)
    run_case("readme", oracle.readme_observations(), 101, (18, 22, 1, 3), out_dir)
    # small ragged case with non-integral ranges (L1 call sequence)
    run_case("small", oracle.readme_observations(n=7, seed=5), 33, (18.5, 21.25, 0.75, 2.5), out_dir)
    # the reference's maximum valid size
    run_case("max256", oracle.readme_observations(n=256, seed=9), 256, (19, 21, 1.5, 2.5), out_dir,
             keep_full=False)


if __name__ == "__main__":
    main()
