#!/bin/bash
# ncu of the LINE-map kernels: usage tools/gpu_lineprof.sh <tag>
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map line | tail -2
timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map compat | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_line.csv python bench.py --map line --rays 100000000 --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_line.log 2>&1; echo "ncu list line rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_map_line_rect|k_prepare_raw" -c 2 -o $out/${tag}_linemap python tools/profile_case.py --rays 100000000 --reps 1 --map line > $out/${tag}_ncu_linemap.log 2>&1; echo "ncu linemap rc=$?"
