"""Multi-GPU driver: one process per GPU (torchrun), rays sharded by global ray id, per-GPU
histograms merged by ONE all-reduce (NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests).

The path has no exchange step inside the bounce loop: ray i's random stream depends
only on (seed, i), so any partition of [0, n) gives bit-identical integer maps after the sum.
torch is plumbing here (device memory, streams, torch.distributed), not the product.
"""
import os

import numpy as np


def shard_range(n_rays, rank, world):
    """Rays [lo, hi) of rank `rank` out of `world`: contiguous, sizes differ by at most one."""
    lo = n_rays * rank // world
    hi = n_rays * (rank + 1) // world
    return lo, hi


def owned_scenes(n_scenes, rank, world):
    """Scene sharding (batched sweeps): scene k belongs to rank k mod world, so that the port-angle series --
    whose cost per ray grows 40x along the series -- is dealt round-robin and every rank gets the same mix."""
    return list(range(rank, n_scenes, world))


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def allreduce_counts(buf, group=None):
    """Sum an int64 tensor over all ranks in place (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def merge_host_counts(counts, stats_vec, group=None):
    """Host-side merge used by the gloo tests and by CPU-only callers: counts uint64 ndarray,
    stats_vec a length-8 uint64 ndarray.  Returns the global sums on every rank."""
    import torch
    t = torch.from_numpy(np.concatenate([counts.reshape(-1).view(np.int64), stats_vec.view(np.int64)]).copy())
    allreduce_counts(t, group)
    out = t.numpy().view(np.uint64)
    return out[:counts.size].reshape(counts.shape).copy(), out[counts.size:].copy()


class ShardedTracer:
    """Flux map over all ranks of the job.  Each call traces `n_rays` GLOBAL rays (ids
    ray_id0 .. ray_id0+n_rays-1), this rank doing its shard_range, and returns the all-reduced
    map.  Device-resident: the kernels add into a torch int64 buffer on this rank's GPU and the
    all-reduce runs on the same stream right behind them."""

    def __init__(self, ctx, scenes, src, mp, seed=4357, device=None, shard="rays"):
        """shard = "rays": every rank traces its slice of the ray ids of EVERY scene (balances a single big scene);
        shard = "scenes": every rank traces ALL rays of the scenes it owns (owned_scenes) -- launches stay large when a
        sweep has many scenes.  Ray ids are global either way, so both give the same integer maps."""
        import torch
        if shard not in ("rays", "scenes"):
            raise ValueError("shard must be 'rays' or 'scenes'")
        self.shard = shard
        from .binding import Scene
        self.torch = torch
        self.ctx, self.src, self.mp, self.seed = ctx, src, mp, seed
        self.scenes = [scenes] if isinstance(scenes, Scene) else list(scenes)
        self.rank, self.world, local = env_rank_world()
        self.device = torch.device("cuda", local if device is None else device)
        self.nb = mp.n_theta * mp.n_phi
        ns = len(self.scenes)
        # [counts (ns*nb) | stats (ns*8)] in ONE buffer -> one all-reduce
        self.buf = torch.zeros(ns * (self.nb + 8), dtype=torch.int64, device=self.device)
        self.host = torch.zeros(ns * (self.nb + 8), dtype=torch.int64).pin_memory()

    def step_device(self, n_rays, ray_id0=0):
        """Asynchronous; returns the device buffer (global sums after the all-reduce)."""
        torch = self.torch
        ns = len(self.scenes)
        self.buf.zero_()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        base = self.buf.data_ptr()
        if self.shard == "scenes" and self.world > 1:
            for k in owned_scenes(ns, self.rank, self.world):       # the other ranks' slots stay zero; the sum fills them
                self.ctx.trace_fluxmap_dev(self.scenes[k], self.src, n_rays, self.mp, base + 8 * k * self.nb,
                                           base + 8 * (ns * self.nb + 8 * k), seed=self.seed, ray_id0=ray_id0, stream=stream)
        else:
            lo, hi = shard_range(n_rays, self.rank, self.world)
            self.ctx.trace_fluxmap_dev(self.scenes, self.src, hi - lo, self.mp, base, base + 8 * ns * self.nb,
                                       seed=self.seed, ray_id0=ray_id0 + lo, stream=stream)
        allreduce_counts(self.buf)
        return self.buf

    def step(self, n_rays, ray_id0=0):
        """Blocking public call: (counts[ns, nb] uint64 ndarray, stats[ns, 8] uint64 ndarray) on the host."""
        buf = self.step_device(n_rays, ray_id0)
        self.host.copy_(buf, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        ns = len(self.scenes)
        a = self.host.numpy().view(np.uint64)
        return a[:ns * self.nb].reshape(ns, self.nb).copy(), a[ns * self.nb:].reshape(ns, 8).copy()
