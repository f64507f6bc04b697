"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/altair_b200.h declares;
struct layouts agree between the header, the binding and the (independently declared) oracle structs."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "altair_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(altb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(altb):
    altb.build_library()
    lib = C.CDLL(altb.library_path())
    names = _declared_functions()
    assert len(names) >= 15 and "altb_trace_fluxmap" in names and "altb_replay" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.altb_version.restype = C.c_int
    assert lib.altb_version() == 1


def test_struct_layouts(altb, oracle):
    assert C.sizeof(altb.Scene) == C.sizeof(oracle.Scene) == 6 * 8 + 4 * 4 + 4 * 8 + 8
    assert C.sizeof(altb.Source) == 48 and C.sizeof(altb.MapSpec) == C.sizeof(oracle.MapSpec) == 32
    assert C.sizeof(altb.Stats) == 64
    assert altb.RECORD_DTYPE.itemsize == oracle.RECORD_DTYPE.itemsize == 32
    for (a, _), (b, _) in zip(altb.Scene._fields_, oracle.Scene._fields_):
        assert a == b and getattr(altb.Scene, a).offset == getattr(oracle.Scene, b).offset
    # defaults are the reference's constants (fluxAtObserverFast.C:33-41, 199)
    s = altb.scene()
    assert (s.r_inner, s.r_outer, s.theta_max_deg, s.world_half) == (100.1, 101.0, 170.0, 300.0)
    assert (s.reflectance, s.roughness_rad, s.max_bounces, s.exit_z) == (0.99, 0.01, 50000, -100.0)


def test_no_cpu_fallback(altb):
    """Without a CUDA device the product path fails loudly (on a GPU box the context simply works)."""
    import torch
    if torch.cuda.is_available():
        altb.Context().close()
        return
    with pytest.raises(altb.AltbError) as e:
        altb.Context()
    assert e.value.code == -4 and "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, include/, the macros or tools/ may reference it
    (measurement helpers that need the oracle live under tests/tools/)."""
    bad = []
    for base in ("altair-raytracing_b200", "altair_raytracing_b200", "include", "tools"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".C")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"pyoracle|altair_oracle|orc_[a-z]+\(|oracle/", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_hot_loop_has_no_spills(altb):
    """Build-quality guard (no GPU needed): at 64 registers per thread ptxas is one temporary away from spilling inside
    the bounce bodies of k_trace, which costs 15-20 % (profiles/README.md).  tools/sass_spills.py counts the local-memory
    instructions between the Philox blocks of the unrolled loop; the roughness instances the benchmarks use must stay at
    the handful that belong to the loop's edges."""
    import shutil
    import subprocess
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    altb.build_library()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_spills.py"), altb.library_path()],
                         capture_output=True, text=True, check=True).stdout
    inside = {m.group(1): int(m.group(2)) for m in re.finditer(r"k_trace<(\d,\d,\d)>:.*?: (\d+) in the bounce bodies", out)}
    # <rough, model, sink>: Lambert / CustomMirror with roughness, sinks 0 records, 1 in-kernel direction map, 2 batched scenes
    want = {f"1,{m},{s}" for m in (0, 1) for s in (0, 1, 2)}
    assert set(inside) >= want, out
    assert all(inside[k] <= 4 for k in want), out
    # the fast contracts' instances (sink 3: the LINES sink as well)
    for label in ("fast", "fast7"):
        got = {m.group(1): int(m.group(2)) for m in re.finditer(r"k_trace<(\d,\d,\d)> %s:.*?: (\d+) in the bounce bodies" % label, out)}
        want_f = {f"1,{m},{s}" for m in (0, 1) for s in (0, 1, 2, 3)}
        assert set(got) >= want_f, out
        assert all(got[k] <= 4 for k in want_f), (label, out)
    # static length of one unrolled bounce body of the headline instance (CustomMirror + roughness, in-kernel direction map):
    # the kernel is bound by issue slots, so this number IS the throughput (profiles/README.md: 185 -> 172 in the third
    # session of round 2; exact contract 344)
    body = {m.group(1): int(m.group(2)) for m in re.finditer(r"k_trace<1,1,1>( fast7| fast|):.*?, (\d+) per bounce body", out)}
    assert body.get(" fast7", 999) <= 176 and body.get(" fast", 999) <= 192 and body.get("", 999) <= 350, body
