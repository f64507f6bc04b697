#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altair_raytracing_b200 as A  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
with A.Context([0]) as ctx:
    for contract in (A.CONTRACT_EXACT, A.CONTRACT_FAST, A.CONTRACT_FAST7):
        ctx.set_contract(contract)
        for mode in (A.MAP_DIRECTION, A.MAP_LINE, A.MAP_TRACEONCE_COMPAT):
            c, st = ctx.trace_fluxmap(A.scene(brdf_kind=1), A.source(), n, A.map_spec(mode=mode), seed=1)
            print("contract", contract, "mode", mode, "sum", int(c.sum()), "bounces", st[0]["n_bounces"])
        scenes = [A.scene(theta_max=t) for t in (100.0, 138.0, 150.0, 165.0, 170.0, 178.0)]
        c, st = ctx.trace_fluxmap(scenes, A.source(), n // 4, A.map_spec(mode=A.MAP_DIRECTION), seed=2)
        print("batched", [int(x.sum()) for x in c])
        c, st = ctx.trace_fluxmap(A.scene(count_all_status=1, theta_max=178.0), A.source(), n // 4, A.map_spec(33, 7, 100.0, 30.0, A.MAP_LINE), seed=3)
        print("odd grid / count_all", int(c.sum()))
    ctx.set_contract(A.CONTRACT_EXACT)
    # LINE maps of odd shapes through the row-stationary kernel: narrow / wide detectors (few / many row pairs per cap, more
    # than 32 columns per rectangle at the pole), odd n_theta (a last row pair without a second row), odd n_phi
    for nt, npb, w in ((180, 90, 40.0), (45, 20, 10.0), (181, 91, 60.0), (17, 250, 25.0), (250, 9, 80.0)):
        c, st = ctx.trace_fluxmap(A.scene(), A.source(), n // 2, A.map_spec(nt, npb, 100.0, w, A.MAP_LINE), seed=7)
        print("line", nt, npb, w, int(c.sum()))
    # brdf_kind 3 (post-hoc re-scatter + second trace) and the horizon diagnostic
    kw = dict(theta_max=140.0, world_half=103.0, r_outer=102.5, reflectance=0.95, roughness=0.2, count_all_status=1, brdf_kind=3,
              brdf_param=(1.0, 1.0, 0.0, 0.0))
    rec, st = ctx.trace_records(A.scene(**kw), A.source(), n // 4, seed=8)
    c, st2 = ctx.trace_fluxmap(A.scene(**kw), A.source(), n // 4, A.map_spec(45, 20, 100.0, 10.0, A.MAP_LINE), seed=8)
    print("posthoc", st["n_bounces"], int(c.sum()), "horizon", ctx.count_horizon(A.scene(roughness=0.5), A.source(), n // 4))
    rec, st = ctx.trace_records(A.scene(), A.source(), n // 4, seed=4)
    c = ctx.map_records(A.scene(), A.map_spec(20, 10, 100.0, 40.0, A.MAP_PER_POSITION, rays_per_position=50), rec)
    print("records", st["n_bounces"], "per-position", int(c.sum()))
    pts, npts, status = ctx.trace_paths(A.scene(), A.source(), 200, 64, seed=5)
    print("paths", int(npts.sum()))
    tape = np.random.default_rng(0).random((2000 * 40, 8)).astype(np.float32)
    off = (np.arange(2001, dtype=np.uint64) * np.uint64(40))
    ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (2000, 1))
    for full in (False, True):
        r, b, p = ctx.replay(A.scene(), ray0, tape, off, A.map_spec(mode=A.MAP_DIRECTION), full_azimuth=full)
        print("replay", full, int(p.sum()))
    centers = np.array([[0.0, 0.0, -200.0], [30.0, 0.0, -195.0]]); rots = np.tile(np.eye(3).ravel(), (2, 1))
    h, st = ctx.detector_sweep(A.scene(reflectance=1.0, roughness=0.0, world_half=200.0, max_bounces=10000), A.source((-60, 0, -80)), n // 4, centers, rots)
    print("disks", h)
