#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 300 python bench.py --map line --rays 100000000 --no-cpu 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; print('LINE value %.4g ms/step %.1f trace %.1f map %.2f crc %s' % (j['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch'], j['map_crc']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_line.csv python tools/profile_case.py --rays 100000000 --reps 3 --map line > $out/${tag}_ncu_line.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_compat.csv python tools/profile_case.py --rays 100000000 --reps 3 --map compat > $out/${tag}_ncu_compat.log 2>&1; echo "ncu list rc=$?"
timeout 900 python -m pytest tests -m gpu -q -k "line or fluxmap or map_stage or macro or Detector" 2>&1 | tail -4
