"""GPU (-m gpu), needs >= 2 GPUs on the box (skipped on a single-GPU box): the sharded path end to
end -- one process per GPU, NCCL process group, both exchange flavours (peer loads fused into the
results kernel over CUDA IPC, and NCCL all_reduce) -- against the unsharded single-GPU result."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, n_obs, mode, out_dir):
    import torch
    import torch.distributed as dist

    import cuda_grid_estimation_b200 as cge
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        obs = oracle.readme_observations(n=n_obs)
        obs_t = torch.from_numpy(obs).to(dev)
        out = {}
        for exchange in ("p2p", "nccl"):
            ctx = cge.Context(rank)
            sh = cge.ShardedGridPosterior(ctx, N, (18, 22), (1, 3), n_obs, mode, exchange=exchange)
            post = sh.alloc_posterior()
            for _ in range(3):      # several steps: exercises the double-buffered exchange block
                sh.step(obs_t, post)
            ex, mm, ms, z = sh.results()
            out[exchange] = dict(ex=ex, mm=mm, ms=ms, z=z, post=post.cpu().numpy(), rows=sh.rows)
            sh.close()
            ctx.close()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
                 **{f"{k}_{f}": v for k, d in out.items() for f, v in d.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,n_obs,mode", [(1000, 50, 0), (257, 50, 0), (512, 2000, 1)])
def test_two_gpu_sharded_matches_single(tmp_path, N, n_obs, mode):
    import torch
    import torch.multiprocessing as mp

    import cuda_grid_estimation_b200 as cge
    import oracle

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on the box")
    world, port = 2, 29600 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, N, n_obs, mode, str(tmp_path)), nprocs=world, join=True)
    with cge.Context(0) as ctx:
        full = ctx.grid_posterior(oracle.readme_observations(n=n_obs), N, (18, 22), (1, 3), mode=mode)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for exchange in ("p2p", "nccl"):
        for g in ranks:
            assert np.abs(g[f"{exchange}_ex"] - full.expectations).max() < 1e-8
            assert np.abs(g[f"{exchange}_ms"] - full.marg_sigma).max() < 1e-8 * full.marg_sigma.max()
            rb, re = g[f"{exchange}_rows"]
            assert np.abs(g[f"{exchange}_mm"][rb:re] - full.marg_mu[rb:re]).max() < 1e-8 * full.marg_mu.max()
        got = np.concatenate([g[f"{exchange}_post"] for g in ranks], axis=0)
        nz = full.posterior > 0
        assert (np.abs(got - full.posterior)[nz] / full.posterior[nz]).max() < 2e-7
        # every rank holds the same totals bit for bit
        assert ranks[0][f"{exchange}_ex"].tobytes() == ranks[1][f"{exchange}_ex"].tobytes()
        assert ranks[0][f"{exchange}_z"].tobytes() == ranks[1][f"{exchange}_z"].tobytes()
    # the rank-ordered peer sum and the NCCL sum agree (two addends: identical)
    assert ranks[0]["p2p_z"] == ranks[0]["nccl_z"]
