#!/usr/bin/env python
"""BASELINE.json config C5: port-angle sweep, 160 scenes (theta_max = 100 + 0.5 k deg) x 1e8 rays each,
source (-60,0,-75) dir (5,0,0), otherwise the fluxAtObserverFast.C scene.  The reference ran such series one
angle at a time (fluxAtObserverFast.C:1641-1673, fluxAtObserverOptimize.C:892-921); here the scenes are a batch
of one call and the rays of every scene are sharded over the ranks (torchrun) by global ray id.

  python tools/port_angle_sweep.py [--scenes 160] [--rays 100000000] [--out profiles/r01_port_angle_sweep.json]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/port_angle_sweep.py ...
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import altair_raytracing_b200 as A  # noqa: E402
from altair_raytracing_b200.distributed import ShardedTracer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=160)
ap.add_argument("--rays", type=int, default=100_000_000)
ap.add_argument("--out", default="")
ap.add_argument("--shard", choices=["rays", "scenes"], default="rays",
                help="N>1: split every scene's rays over the ranks, or deal whole scenes round-robin (measured on 8 B200: 0.448 s vs 0.454 s)")
ap.add_argument("--contract", default="fast7", choices=["exact", "fast", "fast7"])
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
thetas = [100.0 + 0.5 * k for k in range(a.scenes)]
ctx = A.Context([local])
ctx.set_contract({"fast": A.CONTRACT_FAST, "fast7": A.CONTRACT_FAST7, "exact": A.CONTRACT_EXACT}[a.contract])
tr = ShardedTracer(ctx, [A.scene(theta_max=t) for t in thetas], A.source(), A.map_spec(mode=A.MAP_DIRECTION), device=local, shard=a.shard)
tr.step(min(a.rays, 1_000_000))                      # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
counts, stats = tr.step(a.rays)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=tr.device)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    rows = []
    for t, st, c in zip(thetas, stats, counts):
        f = (1 - math.cos(math.radians(180 - t))) / 2
        rows.append({"theta_max": t, "escape_fraction": int(st[2]) / int(st[0]), "thin_wall_formula": 0.99 * f / (1 - 0.99 * (1 - f)),
                     "bounces_per_ray": int(st[5]) / int(st[0]), "map_sum": int(c.sum())})
    tot_b = int(stats[:, 5].sum()); tot_r = int(stats[:, 0].sum())
    summary = {"workload": "C5 port-angle sweep", "contract": a.contract, "n_gpus": world, "shard": a.shard if world > 1 else "none", "scenes": a.scenes, "rays_per_scene": a.rays, "seconds": ms.item() * 1e-3,
               "rays_per_s": tot_r / (ms.item() * 1e-3), "ray_bounces_per_s": tot_b / (ms.item() * 1e-3), "total_bounces": tot_b}
    print(json.dumps(summary))
    for r in rows[::16]:
        print(r)
    if a.out:
        json.dump({"summary": summary, "rows": rows}, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
