#!/usr/bin/env python
"""BASELINE.json config C4: the physical-detector sweep of integratingSphereDetectorSweep.C (362 disk poses at r = 200 cm,
theta = -45..45 deg step 0.5, phi = 0/180; shell 100.1 -> 105 cm, rho = 1, sigma = 0, box 200).  The reference re-traces
10^5 rays for every pose (integratingSphereDetectorSweep.C:54-70); here the rays are traced ONCE and every exited ray is
tested against all poses (k_disk_hits, FP64).  Prints one JSON line.

  python tools/detector_sweep_bench.py [--rays 100000000] [--out profiles/r01_detector_sweep_1gpu.json]"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altair_raytracing_b200 as A  # noqa: E402


def pose(theta, phi, r):           # addDetectorDisk, integratingSphereDetectorSweep.C:145-172: M = Ry(rotTheta) * Rz(rotPhi)
    t, p = math.radians(theta), math.radians(phi)
    c = np.array([r * math.sin(t) * math.cos(p), r * math.sin(t) * math.sin(p), -r * math.cos(t)])
    d = np.array([0.0, 0.0, -100.0]) - c
    rt, rp = -math.atan2(math.hypot(d[0], d[1]), d[2]), math.atan2(d[1], d[0])
    cz, sz, cy, sy = math.cos(rp), math.sin(rp), math.cos(rt), math.sin(rt)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return c, (ry @ rz).reshape(9)


ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=100_000_000)
ap.add_argument("--out", default="")
ap.add_argument("--contract", default="exact", choices=["exact", "fast", "fast7"])
a = ap.parse_args()
poses = [pose(th, ph, 200.0) for th in np.arange(-45.0, 45.0 + 1e-9, 0.5) for ph in (0.0, 180.0)]
centers = np.array([c for c, _ in poses]); rots = np.array([m for _, m in poses])
sc = A.scene(theta_max=170.0, r_outer=105.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000)
src = A.source((-60.0, 0.0, -80.0), (1.0, 0.0, 0.0))
with A.Context([0]) as ctx:
    ctx.set_contract({"exact": A.CONTRACT_EXACT, "fast": A.CONTRACT_FAST, "fast7": A.CONTRACT_FAST7}[a.contract])
    ctx.detector_sweep(sc, src, 1_000_000, centers, rots)            # warm-up
    t0 = time.perf_counter()
    hits, st = ctx.detector_sweep(sc, src, a.rays, centers, rots)
    dt = time.perf_counter() - t0
on_axis = [float(hits[i]) / a.rays for i, (c, _) in enumerate(poses) if abs(c[0]) < 1e-9 and abs(c[1]) < 1e-9]
line = {"workload": "C4 integratingSphereDetectorSweep: trace once, 362 disk poses", "contract": a.contract, "rays": a.rays, "poses": len(poses), "seconds": dt,
        "rays_per_s": a.rays / dt, "ray_bounces_per_s": st["n_bounces"] / dt, "bounces_per_ray": st["n_bounces"] / a.rays,
        "pose_tests_per_s": st["n_exited"] * len(poses) / dt, "on_axis_hit_fraction": on_axis,
        "reference_equivalent_rays": a.rays * len(poses)}
print(json.dumps(line))
if a.out:
    json.dump(line, open(a.out, "w"), indent=1)
