#!/bin/bash
# experiments of the moment (not the round script): usage tools/gpu_exp.sh <tag>
tag=${1:-x}; out=gpurun_out; mkdir -p $out
for v in default altair-raytracing_b200/variants/*.so; do
  [ "$v" = default ] || [ -f "$v" ] || continue
  echo "== $v" >> $out/${tag}_line.log
  if [ "$v" = default ]; then unset ALTB_LIB; else export ALTB_LIB=$v; fi
  timeout 300 python bench.py --map line --rays 100000000 --no-cpu 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; print('value %.4g ms/step %.1f trace %.1f map %.2f' % (j['value'], j['ms_per_step'], r['avg_launch_ms'], r['map_ms_per_launch']))" >> $out/${tag}_line.log
done
unset ALTB_LIB
cat $out/${tag}_line.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_line.csv python bench.py --map line --rays 100000000 --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_line.log 2>&1; echo "ncu list rc=$?"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
