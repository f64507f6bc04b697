"""CSV text produced by the macro mirror (plain C++, runs on a CPU-only box through the writer hooks) against
the reference's own files: same '#' header keys and values, same row formatting and order, same footer keys;
and it parses with the reference's consumer (flux_at_observer/flux_analysis.py:11-57, restated here)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
PKG = os.path.join(ROOT, "altair-raytracing_b200")


@pytest.fixture(scope="module")
def macros(altb):
    import subprocess
    altb.build_library()
    subprocess.check_call(["make", "-C", os.path.join(PKG, "macros")], stdout=subprocess.DEVNULL)
    C.CDLL(altb.library_path(), mode=C.RTLD_GLOBAL)
    L = C.CDLL(os.path.join(PKG, "libaltair_macros.so"))
    L.altbm_unique_filename.restype = C.c_char_p
    d = C.c_double
    L.altbm_write_traceonce_csv.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, d, d, d, d, d, d, d,
                                            C.c_longlong, d, d, d]
    L.altbm_write_perposition_csv.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, d, d, d, d, d, d, d, d]
    return L


def parse_like_flux_analysis(path):
    """flux_analysis.py:11-57: metadata from '# key: value', data = every other line as theta,phi,fraction."""
    meta, rows = {}, []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                if ":" in line:
                    k, v = line[1:].strip().split(":", 1)
                    meta[k.strip()] = v.strip()
            elif line.strip() and not line.startswith("theta"):
                rows.append([float(x) for x in line.split(",")])
    return meta, np.array(rows)


def _mask_times(line):
    line = re.sub(r"\d{4}-\d\d-\d\d \d\d:\d\d:\d\d", "<T>", line)
    return re.sub(r"(time: )[\d.e+-]+( seconds)", r"\1<S>\2", line)


def test_traceonce_csv_matches_reference_text(macros, tmp_path):
    g = GOLD["csv_traceonce"]
    z = np.load(os.path.join(ROOT, "tests", "golden", "traceonce_170.npz"))
    # the first reference run's own rows: rebuild its counts from the published fractions
    first = [float(r.split(",")[2]) for r in g["first_rows"]]
    counts = np.zeros(16200, dtype=np.uint64)
    counts[:3] = np.rint(np.array(first) * 100000).astype(np.uint64)
    counts[-1] = int(round(float(g["last_row"][0].split(",")[2]) * 100000))
    path = str(tmp_path / "t.csv")
    assert macros.altbm_write_traceonce_csv(path.encode(), counts.ctypes.data_as(C.c_void_p), 100000, 180, 90, 170.0, -60.0, 0.0, -75.0,
                                            5.0, 0.0, 0.0, 42303, 3.37814, 149.994, 306.763) == 0
    lines = open(path).read().splitlines()
    assert len(lines) == g["n_lines"] == 16 + 16200 + 5
    head = [l for l in lines[:25] if l.startswith("#") or l.startswith("theta")]
    assert [_mask_times(l) for l in head] == [_mask_times(l) for l in g["header"]]
    data = [l for l in lines if l and l[0].isdigit()]
    assert data[:3] == g["first_rows"] and data[-1:] == g["last_row"] and len(data) == 16200
    foot = [l for l in lines[-8:] if l.startswith("#")]
    assert [_mask_times(l) for l in foot] == [_mask_times(l) for l in g["footer"]]
    assert foot[1:4] == g["footer"][1:4]          # default-format seconds: 306.763, 3.37814, 149.994
    meta, rows = parse_like_flux_analysis(path)
    assert rows.shape == (16200, 3) and meta["Exit port angle"] == "170 degrees" and meta["Number of rays"] == "100000"
    assert np.allclose(rows[:90, 0], 0.25) and np.allclose(rows[:3, 1], [2, 6, 10])           # theta-major order
    assert z["hits"].shape == (16200,)


def test_perposition_csv_matches_reference_text(macros, tmp_path):
    g = GOLD["csv_perposition"]
    z = np.load(os.path.join(ROOT, "tests", "golden", "perposition_170_dir5_0_0.npz"))
    counts = z["hits"].astype(np.uint64)
    path = str(tmp_path / "p.csv")
    assert macros.altbm_write_perposition_csv(path.encode(), 0, counts.ctypes.data_as(C.c_void_p), 50000, 180, 90, 170.0, -60.0, 0.0,
                                              -75.0, 5.0, 0.0, 0.0, 12523.937080) == 0
    lines = open(path).read().splitlines()
    assert len(lines) == g["n_lines"]
    head = [l for l in lines[:25] if l.startswith("#") or l.startswith("theta")]
    assert [_mask_times(l) for l in head] == [_mask_times(l) for l in g["header"]]
    foot = [l for l in lines[-8:] if l.startswith("#")]
    # fixed-format seconds (the row stream's std::fixed is sticky) and the hit total reproduce the reference footer exactly
    assert foot[1:] == g["footer"][1:]
    data = [l for l in lines if l and l[0].isdigit()]
    assert data[:3] == g["first_rows"] and data[-1:] == g["last_row"]
    # every data row of the reference file is reproduced byte for byte from its counts
    meta, rows = parse_like_flux_analysis(path)
    assert np.array_equal(np.rint(rows[:, 2] * 50000).astype(np.uint64), counts)
    assert meta["Total ray hits"] == "5723365 out of 810000000"


def test_twofold_row_order_and_unique_filename(macros, tmp_path):
    counts = np.arange(18 * 10, dtype=np.uint64)
    path = str(tmp_path / "w.csv")
    assert macros.altbm_write_perposition_csv(path.encode(), 1, counts.ctypes.data_as(C.c_void_p), 1000, 18, 10, 164.0, -60.0, 0.0, -80.0,
                                              5.0, 2.0, 0.0, 1.5) == 0
    meta, rows = parse_like_flux_analysis(path)
    assert rows.shape == (180, 3)
    # fluxAtObserverFast.C:700-720: (theta, phi1), (theta, phi1+180) pairs
    assert np.allclose(rows[:4, 1], [18.0, 198.0, 54.0, 234.0])
    assert np.allclose(rows[:2, 2], [0 / 1000, 5 / 1000])
    assert "Twofold" in open(path).read().splitlines()[0]
    # getUniqueFilename (fluxAtObserverOptimize.C:336-387): never overwrite
    assert macros.altbm_unique_filename(path.encode()).decode() == str(tmp_path / "w_1.csv")
    open(tmp_path / "w_1.csv", "w").close()
    assert macros.altbm_unique_filename(path.encode()).decode() == str(tmp_path / "w_2.csv")
    assert macros.altbm_unique_filename(str(tmp_path / "new.csv").encode()).decode() == str(tmp_path / "new.csv")
