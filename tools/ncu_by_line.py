#!/usr/bin/env python
"""Join an ncu report's per-instruction execution counts with nvdisasm line info of the profiled library:
dynamic warp-instructions per CUDA source line / per file, for one kernel.

  python tools/ncu_by_line.py gpurun_out/prof.ncu-rep altair-raytracing_b200/libaltair_b200.so k_traceILb1ELi0 [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, mangled = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = {}, ("?", 0), False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        on = mangled in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
per_line, per_file = defaultdict(lambda: [0, 0]), defaultdict(lambda: [0, 0])
per_op = defaultdict(int)          # dynamic warp-instructions per SASS opcode (predicate and modifiers stripped)
tot = tott = 0
for r in rows[2:]:
    try:
        a, e, t = int(r[ia], 16), int(r[ie]), int(r[it])
    except Exception:
        continue
    if base is None:
        base = a
    off = a - base
    loc, ins = lines.get(off, (("?", 0), ""))
    op = re.sub(r"^@!?U?P\d+\s+", "", ins).split(" ")[0].split(".")[0] if ins else "?"
    per_op[op] += e
    per_line[loc][0] += e; per_line[loc][1] += t
    per_file[loc[0]][0] += e; per_file[loc[0]][1] += t
    tot += e; tott += t
print(f"total warp-instructions {tot:,}  avg active threads {tott / max(tot, 1):.1f}")
srcs = {}
def text(fn, ln):
    if fn not in srcs:
        p = [os.path.join(dp, fn) for dp, _, fs in os.walk(os.path.dirname(os.path.dirname(lib))) for f in fs if f == fn]
        srcs[fn] = open(p[0]).read().splitlines() if p else []
    s = srcs[fn]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
for fn, (e, t) in sorted(per_file.items(), key=lambda x: -x[1][0]):
    print(f"  {e / tot * 100:5.1f}%  thr {t / max(e, 1):4.1f}  {fn}")
# per function: every source line belongs to the last function header above it (inlined code keeps its own line info)
def func_of(fn, ln):
    if fn not in srcs:
        text(fn, 1)
    best = "?"
    for i, l in enumerate(srcs.get(fn, [])[:ln], 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__|ALTB_HD)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.strip().startswith("//"):
            best = m.group(1)
    return best
per_func = defaultdict(lambda: [0, 0])
for (fn, ln), (e, t) in per_line.items():
    per_func[(fn, func_of(fn, ln))][0] += e; per_func[(fn, func_of(fn, ln))][1] += t
unit = float(os.environ.get("NCU_BY_LINE_UNITS", "0"))          # e.g. bounces / 32: prints instructions per unit as well
print("by function:")
for (fn, f), (e, t) in sorted(per_func.items(), key=lambda x: -x[1][0])[:40]:
    extra = f"  {e / unit:7.2f} per unit" if unit else ""
    print(f"  {e / tot * 100:5.2f}%  thr {t / max(e, 1):4.1f}{extra}  {fn}:{f}")
print("by opcode:")
for op, e in sorted(per_op.items(), key=lambda x: -x[1])[:32]:
    extra = f"  {e / unit:7.2f} per unit" if unit else ""
    print(f"  {e / tot * 100:5.2f}%{extra}  {op}")
print("top lines:")
for (fn, ln), (e, t) in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
    print(f"  {e / tot * 100:5.2f}%  thr {t / max(e, 1):4.1f}  {fn}:{ln}  {text(fn, ln)}")
