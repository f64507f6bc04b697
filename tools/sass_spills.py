#!/usr/bin/env python
"""Where are k_trace's spills?  Lists, per template instance, the local-memory instructions (STL/LDL) that sit inside
the two unrolled bounce bodies of the hot loop (between the first and the last Philox IMAD.WIDE of the loop) versus
elsewhere (regeneration, slow path).  A spill inside a bounce body costs throughput; check after every kernel change.

  python tools/sass_spills.py [altair-raytracing_b200/libaltair_b200.so]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "altair-raytracing_b200/libaltair_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
cur, funcs = None, {}
for l in sass:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if "k_traceILb" not in name:
        continue
    ph = [a for a, t in ins if "IMAD.WIDE.U32" in t]
    clusters = []
    for a in ph:
        if not clusters or a - clusters[-1][1] > 0x400:
            clusters.append([a, a])
        clusters[-1][1] = a
    big = [c for c in clusters if c[1] - c[0] >= 0x100]          # full 10-round blocks
    hot = (big[0][0], big[1][1] + 0x1800) if len(big) >= 2 else (0, 0)
    # the second body ends where the drain's FP64 work starts
    f64 = [a for a, t in ins if re.match(r"D(ADD|MUL|FMA|SETP)", t) and a > (big[1][1] if len(big) >= 2 else 0)]
    if f64 and len(big) >= 2:
        hot = (big[0][0], f64[0])
    loc = [(a, t) for a, t in ins if re.search(r"\b(STL|LDL)", t)]
    inside = [a for a, t in loc if hot[0] <= a <= hot[1]]
    tag = re.search(r"k_traceILb(\d)ELi(\d)", name)
    print(f"k_trace<{tag.group(1)},{tag.group(2)}>: {len(ins)} instructions, hot loop {hot[0]:#x}..{hot[1]:#x}, "
          f"local-memory instructions: {len(inside)} in the bounce bodies, {len(loc) - len(inside)} elsewhere")
