#!/bin/bash
# experiments of the moment (not the round script): usage tools/gpu_exp.sh <tag>
tag=${1:-x}; out=gpurun_out; mkdir -p $out
echo "== replay noise, default lib" > $out/${tag}_noise.log
timeout 300 python tests/tools/replay_noise.py 200000 >> $out/${tag}_noise.log 2>&1
for v in altair-raytracing_b200/variants/*.so; do
  [ -f "$v" ] || continue
  echo "== replay noise, $v" >> $out/${tag}_noise.log
  ALTB_LIB=$v timeout 300 python tests/tools/replay_noise.py 200000 2>&1 | grep -v "^$" >> $out/${tag}_noise.log
  for c in fast; do ALTB_LIB=$v timeout 120 python tools/profile_case.py --rays 200000000 --reps 2 --contract $c 2>&1 | tail -1 >> $out/${tag}_noise.log; done
done
cat $out/${tag}_noise.log
timeout 600 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err; echo "bench line rc=$?"
ALTB_LINE_TILES=1 timeout 600 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line_tiles.json 2>> $out/${tag}_bench.err; echo "bench line (tiles) rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_line.csv python bench.py --map line --rays 100000000 --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_line.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_map_line_rect -c 1 -o $out/${tag}_rect python tools/profile_case.py --rays 30000000 --reps 1 --map line > $out/${tag}_ncu_rect.log 2>&1; echo "ncu rect rc=$?"
timeout 600 python -m pytest tests/test_gpu_fast_contract.py -q -x 2>&1 | tail -5
