#!/usr/bin/env python
"""Replay mismatch fractions against the DOUBLE-precision oracle, per scene and per contract (the north star's replay
criterion: status / port flag / direction bin, allowance 1e-4).  Lives under tests/ because it uses the oracle.
  python tests/tools/replay_noise.py [n_rays]        (ALTB_LIB selects an experiment build)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import altair_raytracing_b200 as A  # noqa: E402
import pyoracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
scenes = {"lambert170": dict(theta_max=170.0), "mirror164": dict(theta_max=164.0, brdf_kind=1), "lambert170_s0": dict(theta_max=170.0, roughness=0.0),
          "mirror170_rho1_s0": dict(theta_max=170.0, brdf_kind=1, roughness=0.0, reflectance=1.0, max_bounces=10000, world_half=200.0),
          "lambert170_rho1": dict(theta_max=170.0, reflectance=1.0, max_bounces=10000)}
om = O.map_spec(mode=O.MAP_DIRECTION)
with A.Context([0]) as ctx:
    for name, kw in scenes.items():
        tape, off = O.make_tape(O.scene(**kw), O.source(), n, seed=2)
        ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (n, 1))
        ref = O.replay(O.scene(**kw), ray0, tape, off, prec=O.F64)
        ref_port = O.port_flags(O.scene(**kw), ref)
        ref_bin = np.array([O.lib().orc_direction_bin(C.byref(om), r["dir"].ctypes.data_as(C.POINTER(C.c_float))) if p else -1
                            for r, p in zip(ref, ref_port)], dtype=np.int32)
        out = []
        for cname, c in (("exact", A.CONTRACT_EXACT), ("fast", A.CONTRACT_FAST)):
            ctx.set_contract(c)
            g_rec, g_bin, g_port = ctx.replay(A.scene(**kw), ray0, tape, off, A.map_spec(mode=A.MAP_DIRECTION))
            bad = (g_rec["status"] != ref["status"]) | (g_port.astype(bool) != ref_port) | (g_bin != ref_bin)
            binflip = (g_rec["status"] == ref["status"]) & (g_port.astype(bool) == ref_port) & (g_bin != ref_bin)
            out.append(f"{cname}: {bad.mean():.2e} (bin flips only {binflip.mean():.2e})")
        print(f"{name:20s} exits {ref_port.mean():.2f}  " + "   ".join(out), flush=True)
