#!/bin/bash
# short gpurun call: GPU tests, smoke, the two bench lines.  usage: tools/gpu_check.sh <tag>
tag=${1:-x}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log; echo "pytest ${SECONDS}s"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$? ${SECONDS}s"
timeout 600 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line.json 2>> $out/${tag}_bench.err; echo "bench line rc=$? ${SECONDS}s"
for c in exact fast7; do timeout 300 python tools/detector_sweep_bench.py --contract $c > $out/${tag}_c4_$c.json 2>> $out/${tag}_bench.err; echo "c4 $c rc=$? ${SECONDS}s"; grep -o '"seconds": [0-9.]*' $out/${tag}_c4_$c.json; done
python - <<PY
import json
for f in ('bench','bench_line'):
    try:
        j=json.load(open('$out/${tag}_%s.json'%f)); r=j['roofline']; print(f, 'value %.4g e2e %.4g ms %.1f trace %.1f map %.2f launches %d crc %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch'], j['gpu_launches'], j['map_crc']))
    except Exception as e: print(f, 'failed', e)
PY
