#!/usr/bin/env python
"""tools/root_dump_tape.C's text dump (ROOT + ROBAST side: every random number ROBAST drew per ray, tagged U / G, the polyline
and the final status) -> the replay inputs of include/altair_b200.h:altb_replay_ex:

    ray0[n][6] f64 (start position, direction), tape[total][8] f32 (u_abs, u_r, u_phi, u_sel, u_psi, g0, g1, 0 per surface
    hit), tape_off[n+1] u64 -- plus what the dump says about each ray (status, final direction, number of polyline points)
    for the comparison.

  python tools/tape_from_root_dump.py tape_dump.txt out.npz [--order abs,psi,g0,r,phi]

--order is the sequence in which ROBAST draws per surface hit (SURVEY.md appendix A.3 assumes absorption test, roughness
azimuth, roughness Gaussian, Lambert radius, Lambert azimuth; reading the real order off a dump is the first use of the
macro).  A hit whose draws run out early (the ray was absorbed after the first draw) is padded with zeros.  Replay the
result with ALTB_REPLAY_FULL_AZIMUTH: these uniforms are not the fixed-point turn fractions of this library's own RNG.
"""
import argparse
import sys

import numpy as np

SLOT = {"abs": (0, "U"), "r": (1, "U"), "phi": (2, "U"), "sel": (3, "U"), "psi": (4, "U"), "g0": (5, "G"), "g1": (6, "G")}


def parse_dump(path):
    """-> list of dicts {start[6], draws [(tag, value)], n_points, status, dir[3]}"""
    rays, cur = [], None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "ray":
                cur = {"id": int(t[1]), "start": [float(x) for x in t[3:9]], "draws": [], "n_points": 0, "status": 0, "dir": [0.0, 0.0, 0.0]}
                rays.append(cur)
            elif t[0] == "draws":
                n = int(t[1])
                cur["draws"] = [(t[2 + 2 * k], float(t[3 + 2 * k])) for k in range(n)]
            elif t[0] == "points":
                cur["n_points"] = int(t[1])
            elif t[0] == "status":
                cur["status"] = int(t[1])
                cur["dir"] = [float(x) for x in t[3:6]]
    return rays


def build_tape(rays, order):
    fields = [SLOT[o] for o in order]
    per_hit = len(fields)
    off = np.zeros(len(rays) + 1, dtype=np.uint64)
    recs = []
    for i, r in enumerate(rays):
        d = r["draws"]
        n_hits = -(-len(d) // per_hit)
        for h in range(n_hits):
            rec = np.zeros(8, dtype=np.float64)
            for k, (slot, tag) in enumerate(fields):
                j = h * per_hit + k
                if j >= len(d):
                    break
                if d[j][0] != tag:
                    raise ValueError(f"ray {r['id']} hit {h}: draw {j} is '{d[j][0]}', the order {order} expects '{tag}' -- "
                                     "ROBAST's draw order differs from --order")
                rec[slot] = d[j][1]
            recs.append(rec)
        off[i + 1] = off[i] + n_hits
    tape = np.array(recs, dtype=np.float64).reshape(-1, 8).astype(np.float32)
    # a uniform that rounds up to 1.0f would index past the turn (u in [0,1) by contract)
    for c in (0, 1, 2, 3, 4):
        tape[:, c] = np.minimum(tape[:, c], np.float32(1.0) - np.float32(2.0 ** -24))
    ray0 = np.array([r["start"] for r in rays], dtype=np.float64).reshape(-1, 6)
    return ray0, tape, off


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dump")
    ap.add_argument("out")
    ap.add_argument("--order", default="abs,psi,g0,r,phi")
    a = ap.parse_args()
    order = a.order.split(",")
    for o in order:
        if o not in SLOT:
            sys.exit(f"unknown field '{o}' in --order (known: {', '.join(SLOT)})")
    rays = parse_dump(a.dump)
    ray0, tape, off = build_tape(rays, order)
    np.savez(a.out, ray0=ray0, tape=tape, tape_off=off, status=np.array([r["status"] for r in rays], dtype=np.uint8),
             final_dir=np.array([r["dir"] for r in rays]), n_points=np.array([r["n_points"] for r in rays], dtype=np.uint32))
    print(f"{len(rays)} rays, {len(tape)} surface hits -> {a.out}")


if __name__ == "__main__":
    main()
