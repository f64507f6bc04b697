#!/bin/bash
# The last short call of a round: GPU tests first, then the bench line, the ncu launch list and the full k_trace capture,
# then (if time is left) the LINE bench and C5.   usage: tools/gpu_final2.sh <tag>
tag=${1:-x}; out=gpurun_out; mkdir -p $out
SECONDS=0
timeout 200 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; prc=$?; echo "pytest rc=$prc" >> $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log; echo "pytest ${SECONDS}s"
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 120 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$? ${SECONDS}s"
python - <<PY
import json
try:
    j = json.load(open('$out/${tag}_bench.json')); r = j['roofline']; print('bench value %.4g e2e %.4g ms %.1f trace %.1f launches %d crc %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], r['avg_launch_ms'], j['gpu_launches'], j['map_crc']))
except Exception as e: print('bench failed', e)
PY
[ $prc -ne 0 ] && exit 1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$? ${SECONDS}s"
lib=altair-raytracing_b200/libaltair_b200.so
c=fast7; cmd="tools/profile_case.py --rays 268435456 --reps 1 --contract $c"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -o $out/${tag}_ktrace_$c python $cmd > $out/${tag}_ncu_$c.log 2>&1; echo "ncu $c rc=$? ${SECONDS}s"
b=$(grep -o "bounces [0-9]*" $out/${tag}_ncu_$c.log | head -1 | cut -d" " -f2)
python tools/ncu_summary.py $out/${tag}_ktrace_$c.ncu-rep k_trace > $out/${tag}_ncu_ktrace_$c.md
python tools/ncu_summary.py $out/${tag}_ktrace_$c.ncu-rep k_trace --json "ncu --set full --clock-control none -k regex:k_trace -c 1 python $cmd" $b > $out/${tag}_k_trace_ncu_$c.json
NCU_BY_LINE_UNITS=$(python -c "print($b/32)") python tools/ncu_by_line.py $out/${tag}_ktrace_$c.ncu-rep $lib k_traceILb1ELi1ELi1ELi2E 40 > $out/${tag}_byline_ktrace_$c.txt 2>&1
rm -f $out/${tag}_ktrace_$c.ncu-rep
echo "ncu done ${SECONDS}s"
timeout 60 python bench.py --map line --rays 100000000 --no-cpu > $out/${tag}_bench_line.json 2>> $out/${tag}_bench.err; echo "bench line rc=$? ${SECONDS}s"
timeout 60 python tools/port_angle_sweep.py --out $out/${tag}_c5.json > $out/${tag}_c5.log 2>&1; echo "c5 rc=$? ${SECONDS}s"; head -c 300 $out/${tag}_c5.log
