#!/bin/bash
# One gpurun call: GPU tests, bench lines, C5 sweep, then ncu (launch lists + full captures of the hot kernels).
# usage: tools/gpu_round.sh <tag> [skip-tests|tests] [ref]
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1800 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -4 $out/${tag}_pytest.log
fi
if [ "$3" == "ref" ]; then timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench.err; echo "reference arm rc=$?"; fi
timeout 600 python bench.py > $out/${tag}_bench.json 2>> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --contract fast --no-cpu > $out/${tag}_bench_fast.json 2>> $out/${tag}_bench.err; echo "bench fast rc=$?"
timeout 600 python bench.py --contract exact --no-cpu > $out/${tag}_bench_exact.json 2>> $out/${tag}_bench.err; echo "bench exact rc=$?"
timeout 600 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line.json 2>> $out/${tag}_bench.err; echo "bench line rc=$?"
timeout 300 python tools/port_angle_sweep.py --out $out/${tag}_c5.json > $out/${tag}_c5.log 2>&1; echo "c5 rc=$?"; head -1 $out/${tag}_c5.log
timeout 300 python tools/port_angle_sweep.py --contract exact --out $out/${tag}_c5_exact.json > $out/${tag}_c5_exact.log 2>&1; head -1 $out/${tag}_c5_exact.log
timeout 300 python tools/detector_sweep_bench.py > $out/${tag}_c4.log 2>&1; tail -2 $out/${tag}_c4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_line.csv python bench.py --map line --rays 100000000 --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_line.log 2>&1; echo "ncu list line rc=$?"
# full captures: summarised HERE (markdown table, JSON for bench.py's profile_reference, per-function / per-opcode / per-line
# instruction tables), the 25 MB reports themselves are dropped (gpurun copies at most 64 MiB back)
lib=altair-raytracing_b200/libaltair_b200.so
for c in fast7 exact; do
  cmd="tools/profile_case.py --rays 268435456 --reps 1 --contract $c"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -o $out/${tag}_ktrace_$c python $cmd > $out/${tag}_ncu_$c.log 2>&1; echo "ncu $c rc=$?"
  b=$(grep -o "bounces [0-9]*" $out/${tag}_ncu_$c.log | head -1 | cut -d" " -f2)
  python tools/ncu_summary.py $out/${tag}_ktrace_$c.ncu-rep k_trace > $out/${tag}_ncu_ktrace_$c.md
  python tools/ncu_summary.py $out/${tag}_ktrace_$c.ncu-rep k_trace --json "ncu --set full --clock-control none -k regex:k_trace -c 1 python $cmd" $b > $out/${tag}_k_trace_ncu_$c.json
  m=k_traceILb1ELi1ELi1ELi2E; [ $c == exact ] && m=k_traceILb1ELi1ELi1ELi0E
  NCU_BY_LINE_UNITS=$(python -c "print($b/32)") python tools/ncu_by_line.py $out/${tag}_ktrace_$c.ncu-rep $lib $m 40 > $out/${tag}_byline_ktrace_$c.txt 2>&1
  rm -f $out/${tag}_ktrace_$c.ncu-rep
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_map_line_rect|k_prepare_raw|k_map_line" -c 3 -o $out/${tag}_linemap python tools/profile_case.py --rays 100000000 --reps 1 --map line > $out/${tag}_ncu_linemap.log 2>&1; echo "ncu linemap rc=$?"
python tools/ncu_summary.py $out/${tag}_linemap.ncu-rep > $out/${tag}_ncu_linemap.md
p=$(grep -o "port [0-9]*" $out/${tag}_ncu_linemap.log | head -1 | cut -d" " -f2)
NCU_BY_LINE_UNITS=$p python tools/ncu_by_line.py $out/${tag}_linemap.ncu-rep $lib k_map_line_rect 30 > $out/${tag}_byline_linemap.txt 2>&1
rm -f $out/${tag}_linemap.ncu-rep
du -sh $out
