"""N>1 host logic on CPU: world_size-2 gloo job.  Each rank traces ITS shard of the global ray ids (here with
the CPU oracle standing in for the GPU, which this container does not have), the shards are merged by the
same single all-reduce the GPU path uses, and the result must equal the single-process map bit for bit."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_RAYS, SEED = 30_001, 77


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import pyoracle as O
    from altair_raytracing_b200.distributed import env_rank_world, merge_host_counts, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, _ = env_rank_world()
    lo, hi = shard_range(N_RAYS, r, w)
    counts, st = O.fluxmap(O.scene(), O.source(), hi - lo, O.map_spec(mode=O.MAP_DIRECTION), seed=SEED, ray_id0=lo,
                           prec=O.F32, n_threads=1)
    vec = np.array([st["n_rays"], st["n_exited"], st["n_exit_port"], st["n_absorbed"], st["n_suspended"], st["n_bounces"], 0, 0],
                   dtype=np.uint64)
    g_counts, g_vec = merge_host_counts(counts, vec)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), counts=g_counts, stats=g_vec, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_world_size_2_sharding_is_bit_identical(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full, st = oracle.fluxmap(oracle.scene(), oracle.source(), N_RAYS, oracle.map_spec(mode=oracle.MAP_DIRECTION), seed=SEED,
                              prec=oracle.F32)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert got[0]["lo"] == 0 and got[0]["hi"] == got[1]["lo"] and got[1]["hi"] == N_RAYS
    for g in got:                                   # every rank holds the global sums after the all-reduce
        assert np.array_equal(g["counts"], full)
        assert g["stats"][0] == N_RAYS and g["stats"][5] == st["n_bounces"] and g["stats"][2] == st["n_exit_port"]


def _scene_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import pyoracle as O
    from altair_raytracing_b200.distributed import merge_host_counts, owned_scenes
    dist.init_process_group("gloo", rank=rank, world_size=world)
    thetas = [150.0, 160.0, 170.0]
    mp_ = O.map_spec(mode=O.MAP_DIRECTION)
    nb = mp_.n_theta * mp_.n_phi
    counts = np.zeros((len(thetas), nb), dtype=np.uint64)
    vec = np.zeros(8 * len(thetas), dtype=np.uint64)
    for k in owned_scenes(len(thetas), rank, world):            # whole scenes, all ray ids; the other slots stay zero
        c, st = O.fluxmap(O.scene(theta_max=thetas[k]), O.source(), 6000, mp_, seed=SEED, prec=O.F32, n_threads=1)
        counts[k] = c
        vec[8 * k:8 * k + 6] = [st["n_rays"], st["n_exited"], st["n_exit_port"], st["n_absorbed"], st["n_suspended"], st["n_bounces"]]
    g_counts, g_vec = merge_host_counts(counts, vec)
    np.savez(os.path.join(out_dir, f"scene_rank{rank}.npz"), counts=g_counts, stats=g_vec)
    dist.destroy_process_group()


def test_world_size_2_scene_sharding_is_bit_identical(tmp_path, oracle):
    """Batched sweeps deal whole scenes round-robin (ShardedTracer(shard="scenes")): same single all-reduce, same maps."""
    world = 2
    mp.spawn(_scene_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"scene_rank{r}.npz") for r in range(world)]
    for k, th in enumerate([150.0, 160.0, 170.0]):
        full, st = oracle.fluxmap(oracle.scene(theta_max=th), oracle.source(), 6000, oracle.map_spec(mode=oracle.MAP_DIRECTION),
                                  seed=SEED, prec=oracle.F32)
        for g in got:
            assert np.array_equal(g["counts"][k], full)
            assert g["stats"][8 * k] == 6000 and g["stats"][8 * k + 5] == st["n_bounces"]


def test_owned_scenes_partition():
    sys.path.insert(0, ROOT)
    from altair_raytracing_b200.distributed import owned_scenes
    for n in (0, 1, 5, 160):
        for w in (1, 2, 3, 8):
            owned = [owned_scenes(n, r, w) for r in range(w)]
            assert sorted(k for o in owned for k in o) == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    from altair_raytracing_b200.distributed import shard_range
    for n in (0, 1, 7, 1000, 10 ** 9 + 7):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
