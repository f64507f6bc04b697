#!/usr/bin/env python
"""Builds the golden fixtures in this directory from the reference's own committed outputs.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Everything written here is DATA the reference published (CSV rows, footers, histograms), reduced to
small integer arrays; no reference code is copied.  Citations are relative to /root/reference.
"""
import glob
import json
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_map(path):
    """-> (meta dict from '# key: value' lines, fractions[n_rows] in file order)."""
    meta, fr = {}, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith("#"):
                m = re.match(r"#\s*([^:]+):\s*(.*)", line)
                if m:
                    meta[m.group(1).strip()] = m.group(2).strip()
            elif line and not line.startswith("theta"):
                fr.append(float(line.split(",")[2]))
    return meta, np.array(fr)


def main():
    gold = {"escape_counts": {}, "trace_seconds": {}, "perposition_total_hits": {}}
    # --- trace-once maps (semantics B) and their escape counts: fluxAtObserverFast.C:1374-1382 footers
    for name, theta in (("portAngleSweep_04_02_-60_0_-75_160", 160), ("portAngleSweep_04_03_-60_0_-75_164", 164),
                        ("trace_once_test_04_2-60_0_-75_5", 170)):
        files = sorted(glob.glob(f"{REF}/flux_at_observer/{name}/*.csv"))
        total = np.zeros(16200, dtype=np.int64)
        esc, secs = [], []
        for p in files:
            meta, fr = read_map(p)
            assert len(fr) == 16200 and meta["Exit port angle"].startswith(str(theta))
            total += np.rint(fr * 100000).astype(np.int64)
            esc.append(int(meta["Total rays exiting port"].split()[0]))
            secs.append(float(meta["Ray tracing time"].split()[0]))
        gold["escape_counts"][str(theta)] = esc
        gold["trace_seconds"][str(theta)] = secs
        np.savez_compressed(f"{OUT}/traceonce_{theta}.npz", hits=total.astype(np.uint32), n_rays=100000 * len(files))
    # --- per-position maps (semantics A), 50 000 rays per bin: fluxAtObserverOptimize.C:571-579,667-670
    pp = {"170_dir5_0_0": "results_overnight_03_31-60_0_-75_5/fluxmap_50000rays_180x90_src-60_0_-75.csv",
          "163_dir5_0_0": "results_overnight_04_1-60_0_-75_5/fluxmap_50000rays_180x90_src-60_0_-75.csv"}
    for key, rel in pp.items():
        meta, fr = read_map(f"{REF}/flux_at_observer/{rel}")
        assert len(fr) == 16200
        hits = np.rint(fr * 50000).astype(np.uint16)
        gold["perposition_total_hits"][key] = {"footer": int(meta["Total ray hits"].split()[0]), "sum": int(hits.sum()),
                                               "exit_port_angle": meta["Exit port angle"],
                                               "source_direction": meta["Source direction (x,y,z)"]}
        np.savez_compressed(f"{OUT}/perposition_{key}.npz", hits=hits, rays_per_bin=50000)
    # --- exit-direction goldens: distributionSphereDetectorSweep.C:74-99 outputs
    d = np.loadtxt(f"{REF}/3dRayLog.txt", comments="#")
    assert d.shape == (100000, 3)
    gold["raylog"] = {"n": 100000, "mean_dz": float(d[:, 2].mean()), "max_dz": float(d[:, 2].max()),
                      "dz_hist_100": np.histogram(d[:, 2], bins=100, range=(-1, 1))[0].tolist()}
    a = np.loadtxt(f"{REF}/angular_dist.txt", comments="#")
    gold["angular_dist"] = {"centers": a[:, 0].tolist(), "counts": a[:, 1].astype(int).tolist()}
    # --- nonLambertianFlux.C:371-384 output (45x20, 100 000 rays per bin, 10 cm detector, unseeded)
    nl = np.loadtxt(f"{REF}/flux_at_observer/fluxmap_data.csv", delimiter=",", skiprows=1)
    assert nl.shape == (900, 3)
    np.savez_compressed(f"{OUT}/nonlambertian_45x20.npz", hits=np.rint(nl[:, 2] * 100000).astype(np.uint32), rays_per_bin=100000)
    # --- CSV text format: header + first rows + footer of one trace-once and one per-position file (verbatim lines)
    for tag, rel in (("traceonce", "trace_once_test_04_2-60_0_-75_5/fluxmap_traceonce_100000rays_180x90_src-60_0_-75.csv"),
                     ("perposition", "results_overnight_03_31-60_0_-75_5/fluxmap_50000rays_180x90_src-60_0_-75.csv")):
        lines = open(f"{REF}/flux_at_observer/{rel}").read().splitlines()
        head = [l for l in lines[:25] if l.startswith("#") or l.startswith("theta")]
        first = [l for l in lines if l and l[0].isdigit()][:3]
        last = [l for l in lines if l and l[0].isdigit()][-1:]
        foot = [l for l in lines[-8:] if l.startswith("#")]
        gold[f"csv_{tag}"] = {"header": head, "first_rows": first, "last_row": last, "footer": foot, "n_lines": len(lines)}
    # --- integratingSphereDetectorSweep older outputs (weak goldens, parameters not recoverable)
    ds = np.loadtxt(f"{REF}/detector_sweep.txt", skiprows=1)
    on_axis = ds[np.abs(ds[:, 0]) < 1e-9][:, 2]
    gold["detector_sweep_txt"] = {"rows": int(ds.shape[0]), "theta0_mean_fraction": float(on_axis.mean()), "theta0_rows": int(on_axis.size)}
    with open(f"{OUT}/golden.json", "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
