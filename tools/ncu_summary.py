#!/usr/bin/env python
"""Print the metrics profiles/*.md quote from an ncu report as a markdown table:
  python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]
  python tools/ncu_summary.py gpurun_out/x.ncu-rep k_trace --json "<capture command>" <bounces in the launch> > profiles/rNN_k_trace_ncu.json
The JSON form is what bench.py attaches to its line as `roofline.profile_reference` (numbers of a committed capture with the
capture's own configuration -- never re-labelled as measurements of the bench run)."""
import json
import csv
import re
import subprocess
import sys

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__warps_eligible.avg.per_cycle_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum derived__smsp__inst_executed_op_branch_pct
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio""".split()

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
as_json = len(sys.argv) > 3 and sys.argv[3] == "--json"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    if as_json:
        def val(m):
            return float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] not in ("", "n/a") else None
        def byt(m):
            v, u = val(m), units[hdr.index(m)] if m in hdr else ""
            return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        bounces = float(sys.argv[5]) if len(sys.argv) > 5 else None
        inst = val("smsp__inst_executed.sum")
        dur, du = val("gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")]
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(du, 1)
        out_ = {"kernel": name, "capture": sys.argv[4] if len(sys.argv) > 4 else "", "report": rep,
                "duration_ms": dur_ms, "bounces_in_launch": bounces,
                "warp_instructions": inst,
                "warp_instructions_per_32_bounces": inst * 32 / bounces if bounces else None,
                "active_lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
                "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "dram_bytes_per_launch": (byt("dram__bytes_read.sum") or 0) + (byt("dram__bytes_write.sum") or 0),
                "registers_per_thread": val("launch__registers_per_thread"), "grid": val("launch__grid_size"),
                "block": val("launch__block_size")}
        print(json.dumps(out_, indent=1))
        break
    print(f"## {name}\n\n| metric | unit | value |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            print(f"| {w} | {units[hdr.index(w)]} | {r[hdr.index(w)]} |")
    print()
