#!/usr/bin/env python
"""One small hot-path invocation for ncu / compute-sanitizer (same kernels and launch shapes as bench.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altair_raytracing_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=8_000_000)
ap.add_argument("--map", default="direction", choices=["direction", "line", "compat"])
ap.add_argument("--brdf", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--rho", type=float, default=0.99)
ap.add_argument("--theta", type=float, default=170.0)
ap.add_argument("--limit", type=int, default=50000)
ap.add_argument("--sigma", type=float, default=0.01)
ap.add_argument("--contract", default="fast7", choices=["exact", "fast", "fast7"])
ap.add_argument("--batch", type=int, default=0, help="rays per launch (0: library default 2^26)")
a = ap.parse_args()
mode = {"direction": A.MAP_DIRECTION, "line": A.MAP_LINE, "compat": A.MAP_TRACEONCE_COMPAT}[a.map]
with A.Context([0]) as ctx:
    ctx.set_contract({"fast": A.CONTRACT_FAST, "fast7": A.CONTRACT_FAST7, "exact": A.CONTRACT_EXACT}[a.contract])
    if a.batch:
        ctx.set_batch(a.batch)
    for r in range(a.reps):
        counts, st = ctx.trace_fluxmap(A.scene(brdf_kind=a.brdf, roughness=a.sigma, reflectance=a.rho, theta_max=a.theta, max_bounces=a.limit), A.source(), a.rays, A.map_spec(mode=mode), seed=4357,
                                       ray_id0=r * a.rays)
        s = st[0]
        print(f"rep {r}: rays {s['n_rays']} bounces {s['n_bounces']} port {s['n_exit_port']} trace {s['t_trace_s']*1e3:.2f} ms "
              f"map {s['t_map_s']*1e3:.2f} ms -> {s['n_bounces']/s['t_trace_s']:.3e} bounces/s (trace kernel)")
