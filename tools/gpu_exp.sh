#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_map_line_rect -c 1 -o $out/${tag}_rect python tools/profile_case.py --rays 30000000 --reps 1 --map line > $out/${tag}_ncu_rect.log 2>&1; echo "ncu rect rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prepare_lines -c 1 -o $out/${tag}_prep python tools/profile_case.py --rays 30000000 --reps 1 --map line > $out/${tag}_ncu_prep.log 2>&1; echo "ncu prep rc=$?"
