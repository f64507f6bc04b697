#!/bin/bash
# ncu --set full of ONE k_trace launch (2^28 rays of the C3 scene) per contract: usage tools/gpu_traceprof.sh <tag> [contracts...]
tag=${1:-x}; shift; out=gpurun_out; mkdir -p $out
for c in ${@:-fast7}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -o $out/${tag}_ktrace_$c python tools/profile_case.py --rays 268435456 --reps 1 --contract $c > $out/${tag}_ncu_$c.log 2>&1; echo "ncu $c rc=$?"; tail -1 $out/${tag}_ncu_$c.log
done
