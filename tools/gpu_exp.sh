#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line.json 2>> $out/${tag}_bench.err; echo "bench line rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_line.csv python bench.py --map line --rays 100000000 --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_line.log 2>&1; echo "ncu list line rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_map_line_rect|k_prepare_raw|k_map_line" -c 3 -o $out/${tag}_linemap python tools/profile_case.py --rays 100000000 --reps 1 --map line > $out/${tag}_ncu_linemap.log 2>&1; echo "ncu linemap rc=$?"
timeout 300 python tools/profile_case.py --rays 100000000 --reps 2 --map compat | tail -1
python -c "
import json
for f in ('bench','bench_line'):
    j=json.load(open('$out/${tag}_%s.json'%f)); r=j['roofline']; print(f, 'value %.4g e2e %.4g ms %.1f trace %.1f map %.2f launches %d crc %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch'], j['gpu_launches'], j['map_crc']))
"
