#!/bin/bash
# One gpurun call: GPU tests, bench lines, C5 sweep, then ncu (launch list of bench.py + one full capture of k_trace).
# usage: tools/gpu_round.sh <tag> [skip-tests] ; env BENCH_FLAGS (e.g. "--contract exact")
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1800 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -8 $out/${tag}_pytest.log
fi
timeout 600 python bench.py $BENCH_FLAGS > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --contract exact --no-cpu > $out/${tag}_bench_exact.json 2>> $out/${tag}_bench.err; echo "bench exact rc=$?"
timeout 600 python bench.py --map line --rays 100000000 --no-cpu $BENCH_FLAGS > $out/${tag}_bench_line.json 2>> $out/${tag}_bench.err; echo "bench line rc=$?"
timeout 300 python tools/port_angle_sweep.py --out $out/${tag}_c5.json > $out/${tag}_c5.log 2>&1; echo "c5 rc=$?"; head -1 $out/${tag}_c5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu $BENCH_FLAGS > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -o $out/${tag}_ktrace python tools/profile_case.py --rays 60000000 --reps 1 $PROFILE_FLAGS > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -20
for v in altair-raytracing_b200/variants/*.so; do
  [ -f "$v" ] || continue
  echo "variant $v" >> $out/${tag}_variants.log
  for c in fast exact; do ALTB_LIB=$v timeout 120 python tools/profile_case.py --rays 200000000 --reps 2 --contract $c 2>&1 | tail -1 >> $out/${tag}_variants.log; done
done
echo "variant default" >> $out/${tag}_variants.log
for c in fast exact; do timeout 120 python tools/profile_case.py --rays 200000000 --reps 2 --contract $c 2>&1 | tail -1 >> $out/${tag}_variants.log; done
cat $out/${tag}_variants.log
