"""CPU, world_size 2 over gloo: the multi-rank host logic of the path -- row sharding, the
packed partial vector, its all-reduce, and the finalisation -- with the oracle standing in for
the per-rank likelihood kernel.  The device kernels fold_publish_kernel / finish_kernel compute
exactly what pack_partials / finalize_from_partials state."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from cuda_grid_estimation_b200 import finalize_from_partials, pack_partials, shard_rows


def test_shard_rows_partition():
    for N in (2, 3, 101, 4096, 16384, 65536):
        for world in (1, 2, 3, 4, 8):
            if N < world:
                with pytest.raises(ValueError):
                    shard_rows(N, world, 0)
                continue
            spans = [shard_rows(N, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _worker(rank, world, port, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle.lib().oracle_set_num_threads(1)
        obs = oracle.readme_observations()
        rb, re = shard_rows(N, world, rank)
        likes = oracle.grid_likes(obs, N, 18, 22, 1, 3, oracle.NUM_F64, rb, re)
        mu = oracle.linspace(N, 18, 22)
        sigma = oracle.linspace(N, 1, 3)
        part = torch.from_numpy(pack_partials(likes, mu[rb:re]))
        dist.all_reduce(part, op=dist.ReduceOp.SUM)          # the path's one collective
        fin = finalize_from_partials(part.numpy(), likes.sum(axis=1), sigma, rb, re)
        post_local = likes * fin["inv_z"]
        gathered = [None] * world
        dist.all_gather_object(gathered, (rb, re, post_local, fin["marg_mu"][rb:re]))
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), expectations=fin["expectations"],
                     marg_sigma=fin["marg_sigma"],
                     posterior=np.concatenate([g[2] for g in gathered], axis=0),
                     marg_mu=np.concatenate([g[3] for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [101, 64])
def test_two_rank_pipeline_matches_single(tmp_path, N):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    ref = oracle.grid_posterior(oracle.readme_observations(), N, 18, 22, 1, 3, oracle.NUM_F64)
    assert np.abs(got["expectations"] - ref["expectations"]).max() < 1e-12
    assert np.abs(got["marg_sigma"] - ref["marg_sigma"]).max() < 1e-14
    assert np.abs(got["marg_mu"] - ref["marg_mu"]).max() < 1e-14
    assert np.abs(got["posterior"] - ref["posterior"]).max() < 1e-15
    assert abs(got["posterior"].sum() - 1.0) < 1e-12
