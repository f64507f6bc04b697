#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 300 python bench.py --map line --rays 100000000 --no-cpu 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; print('LINE value %.4g ms/step %.1f trace %.1f map %.2f crc %s' % (j['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch'], j['map_crc']))"
timeout 300 python bench.py --no-cpu 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; print('C3 fast value %.4g e2e %.4g ms/step %.1f crc %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['map_crc']))"
