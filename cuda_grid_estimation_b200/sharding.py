"""Row sharding of the (mu, sigma) grid across ranks and the one exchange step of the path.

The grid is independent per cell; rank r owns a contiguous block of mu rows of
posterior[mu][sigma].  The only data that crosses ranks is the packed partial vector
    P = {Z_r, sum_i mu_i*rowsum_i, colsum_r[0..N)}            (N+2 float64)
summed over ranks (one all-reduce); after it every rank knows Z, E[mu], the sigma marginal and
E[sigma], and scales its own rows by 1/Z.  No posterior data ever moves between GPUs.

``pack_partials`` / ``finalize_from_partials`` restate, in numpy, exactly what the device
kernels fold_publish_kernel / finish_kernel compute, so the multi-rank host logic can be
exercised on CPU with the gloo backend (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import numpy as np


def shard_rows(grid_size: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of mu rows: the first ``grid_size % world_size`` ranks get one
    extra row.  Empty shards are not allowed (the C ABI rejects row_begin >= row_end)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    if grid_size < world_size:
        raise ValueError(f"grid_size {grid_size} < world_size {world_size}: a rank would be empty")
    base, rem = divmod(grid_size, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_partials(unnormalised_rows: np.ndarray, mu_rows: np.ndarray) -> np.ndarray:
    """{Z, sum_i mu_i*rowsum_i, colsum[0..N)} of one shard of unnormalised likelihoods."""
    rows = np.asarray(unnormalised_rows, np.float64)
    rowsum = rows.sum(axis=1)
    out = np.empty(rows.shape[1] + 2, np.float64)
    out[0] = rowsum.sum()
    out[1] = np.dot(rowsum, np.asarray(mu_rows, np.float64))
    out[2:] = rows.sum(axis=0)
    return out


def finalize_from_partials(partials: np.ndarray, rowsum_local: np.ndarray, sigma: np.ndarray,
                           row_begin: int, row_end: int):
    """Marginals and expectations from the all-reduced partial vector (finish_kernel)."""
    partials = np.asarray(partials, np.float64)
    N = partials.size - 2
    inv = 1.0 / partials[0]
    marg_sigma = partials[2:] * inv
    marg_mu = np.zeros(N, np.float64)
    marg_mu[row_begin:row_end] = np.asarray(rowsum_local, np.float64) * inv
    e_mu = partials[1] * inv
    e_sigma = float(np.sum(marg_sigma * np.asarray(sigma, np.float64)))
    return dict(expectations=np.array([e_mu, e_sigma]), marg_mu=marg_mu, marg_sigma=marg_sigma,
                z=partials[0], inv_z=inv)


class ShardedGridPosterior:
    """One process per GPU.  Each call enqueues, on torch's current stream:
    likelihood + local fold  ->  exchange of P  ->  results + in-place scale of the local rows.

    exchange="p2p" (default for world_size > 1): the ranks swap CUDA IPC handles once (through
    torch.distributed), after which the results kernel reads the peers' partial vectors over
    NVLink and adds them in rank order itself -- no collective launch in the step.
    exchange="nccl": torch.distributed.all_reduce of P between the two phases.
    With world_size == 1 there is no exchange."""

    def __init__(self, ctx, grid_size, mu_range, sigma_range, n_obs, mode=0, group=None, variant=0,
                 exchange="p2p"):
        import torch
        import torch.distributed as dist
        from .api import make_params
        self.torch, self.dist = torch, dist
        self.ctx = ctx
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.rows = shard_rows(grid_size, self.world, self.rank)
        self.params = make_params(n_obs, grid_size, mu_range, sigma_range, mode, self.rows, variant)
        self.grid_size = grid_size
        self._partials = None
        if exchange not in ("p2p", "nccl", "auto"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.exchange = exchange if self.world > 1 else "none"
        if self.exchange in ("p2p", "auto"):
            self.exchange = self._connect_peers(exchange == "auto")

    def _connect_peers(self, may_fall_back: bool) -> str:
        """CUDA IPC handshake through torch.distributed.  With may_fall_back the ranks agree
        (all-reduce of a success flag) to use the NCCL exchange if any of them could not map a
        peer -- both are GPU paths; there is no CPU path to fall back to."""
        from .api import CgeError
        dist, ctx, torch = self.dist, self.ctx, self.torch
        err = None
        try:
            mine = ctx.peer_local_handle(self.grid_size)
        except CgeError as e:
            mine, err = b"\0" * ctx.IPC_HANDLE_BYTES, e
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine, group=self.group)
        if err is None:
            try:
                ctx.peer_connect(self.rank, self.world, b"".join(gathered))
            except CgeError as e:
                err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32,
                          device=torch.device("cuda", ctx.device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            return "p2p"
        try:
            ctx.peer_disconnect()
        except CgeError:
            pass
        if not may_fall_back:
            raise err if err else RuntimeError("a peer rank could not set up the CUDA IPC exchange")
        return "nccl"

    def close(self):
        """Collective: every rank must call it before its context is destroyed."""
        if self.exchange == "p2p":
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            self.ctx.peer_disconnect()
            self.exchange = "none"

    @property
    def rows_local(self) -> int:
        return self.rows[1] - self.rows[0]

    def alloc_posterior(self):
        return self.torch.empty((self.rows_local, self.grid_size), dtype=self.torch.float32,
                                device=self.torch.device("cuda", self.ctx.device))

    def step(self, obs_dev, posterior_dev):
        """Enqueue one pass of the hot path; returns nothing (results stay on the device)."""
        from .api import device_tensor
        stream = self.torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            # no exchange: the fused asynchronous entry (one launch for small grids, three otherwise)
            self.ctx.grid_posterior_async(self.params, obs_dev, posterior_dev, stream)
            return
        self.ctx.likelihood_partials(self.params, obs_dev, posterior_dev, stream)
        if self.exchange == "nccl":
            ptr, cnt = self.ctx.partials_ptr()
            if self._partials is None or self._partials.data_ptr() != ptr:
                self._partials = device_tensor(ptr, cnt, "float64", self.ctx.device)
            self.dist.all_reduce(self._partials, op=self.dist.ReduceOp.SUM, group=self.group)
        self.ctx.finalize(self.params, posterior_dev, stream)

    def results(self):
        """Blocking read of {E[mu], E[sigma]}, marg_mu (this shard's rows), marg_sigma, Z."""
        stream = self.torch.cuda.current_stream().cuda_stream
        return self.ctx.read_results(self.params, stream)
