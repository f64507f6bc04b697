#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
ALTB_LINE_RECORDS=1 timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map line | tail -1
timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map line | tail -1
timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map compat | tail -1
timeout 900 python -m pytest tests -m gpu -q -k "line or fluxmap or map_stage or macro or Detector or residual" 2>&1 | tail -4
timeout 300 python bench.py --map line --rays 100000000 --no-cpu 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; print('LINE value %.4g ms/step %.1f trace %.1f map %.2f crc %s' % (j['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch'], j['map_crc']))"
