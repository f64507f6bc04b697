#!/usr/bin/env python
"""Replay-mode measurement (lives under tests/ because it uses the oracle, which is test infrastructure): record a tape
with the CPU oracle, replay it on the GPU, check it against the Philox run, print the kernel's tape bandwidth
(ALTB_TIMING=1).  The replay kernel is the one HBM-bound kernel of the path: 32 B of recorded draws per surface hit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["ALTB_TIMING"] = "1"
import altair_raytracing_b200 as A  # noqa: E402
import pyoracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
t0 = time.time()
tape, off = O.make_tape(O.scene(), O.source(), n, seed=4357)
print(f"oracle tape: {len(tape)} records for {n} rays in {time.time() - t0:.1f} s")
ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (n, 1))
with A.Context([0]) as ctx:
    for _ in range(3):
        rec, bins, port = ctx.replay(A.scene(), ray0, tape, off, A.map_spec(mode=A.MAP_DIRECTION))
    ref, _ = ctx.trace_records(A.scene(), A.source(), n, seed=4357)
    assert rec.tobytes() == ref.tobytes(), "replay differs from the Philox run it was recorded from"
    print("replay == Philox run: bit-exact;", int(port.sum()), "rays through the port")
