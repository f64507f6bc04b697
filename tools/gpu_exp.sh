#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py 30000 > $out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_case.py 8000 > $out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $out/${tag}_racecheck.log
