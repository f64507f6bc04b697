#!/usr/bin/env python
"""Where are k_trace's spills?  Lists, per template instance, the local-memory instructions (STL/LDL) that sit inside
the two unrolled bounce bodies of the hot loop (between the first and the last Philox IMAD.WIDE of the loop) versus
elsewhere (regeneration, slow path), and the static length of one bounce body (SASS instructions from one Philox block of
the unrolled loop to the next).  A spill inside a bounce body costs throughput; check after every kernel change.

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
        if not clusters or a - clusters[-1][1] > 0x60:
            clusters.append([a, a, 0])
        clusters[-1][1] = a
        clusters[-1][2] += 1
    big = [c for c in clusters if c[2] >= 14]                    # full Philox blocks (10 rounds: 20 wide multiplies; 7 rounds: 14)
    # bounce bodies: from the first Philox block of the hot loop to one body length past the last one
    hot = (0, 0)
    body = 0
    if len(big) >= 2:
        body = big[1][0] - big[0][0]
        n_unrolled = 1
        while n_unrolled < len(big) and abs((big[n_unrolled][0] - big[n_unrolled - 1][0]) - body) < 0x300:
            n_unrolled += 1
        hot = (big[0][0], big[n_unrolled - 1][0] + body)
    loc = [(a, t) for a, t in ins if re.search(r"\b(STL|LDL)", t)]
    inside = [a for a, t in loc if hot[0] <= a <= hot[1]]
    tag = re.search(r"k_traceILb(\d)ELi(\d)ELi(\d)ELi(\d)", name)
    print(f"k_trace<{tag.group(1)},{tag.group(2)},{tag.group(3)}>{('', ' fast', ' fast7')[int(tag.group(4))]}: {len(ins)} instructions, hot loop {hot[0]:#x}..{hot[1]:#x}, {body // 16} per bounce body, "
          f"local-memory instructions: {len(inside)} in the bounce bodies, {len(loc) - len(inside)} elsewhere")
