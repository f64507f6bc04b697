#!/bin/bash
# multi-GPU checks on one box: usage [SKIP_REF=1] [SKIP_EXACT=1] tools/gpu_multi.sh <N> <tag>   (an 8-GPU box is charged 8 x: the CPU reference arm alone is 12 GPU-minutes there)
N=${1:-2}; tag=${2:-m}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | head -$N
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" 2>&1 | tail -3
[ -z "$SKIP_REF" ] && timeout 600 $TR bench.py --gpus $N --impl reference --steps 5 --warmup 1 > $out/${tag}_ref_n$N.json 2> $out/${tag}_ref_n$N.err; echo "reference arm N=$N rc=$?"; cut -c1-300 $out/${tag}_ref_n$N.json
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err; echo "bench N=$N rc=$?"
[ -z "$SKIP_EXACT" ] && timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --contract exact --no-other > $out/${tag}_bench_exact_n$N.json 2>> $out/${tag}_bench_n$N.err; echo "bench exact N=$N rc=$?"
timeout 600 $TR tools/port_angle_sweep.py --out $out/${tag}_c5_n$N.json > $out/${tag}_c5_n$N.log 2>&1; head -1 $out/${tag}_c5_n$N.log
timeout 600 python - <<'PY' > $out/${tag}_macro_n${N}.log 2>&1
# the C++ macro mirror with threads = N (one process, N GPUs, NCCL inside the C ABI) against threads = 1: byte-identical CSV
import os, sys, filecmp, tempfile
sys.path.insert(0, os.getcwd())
import torch
from altair_raytracing_b200 import macros
n = torch.cuda.device_count()
d = tempfile.mkdtemp()
macros.set_output_dir(d); macros.set("verbose", 0); macros.set("advance_ray_ids", 0); macros.set("traceonce_rays", 2000000)
macros.sweepDetectorTraceOnce(False, "one", 1, -60, 0, -75, 5, 0, 0, 170.0); a = macros.last_csv()
macros.sweepDetectorTraceOnce(False, "many", n, -60, 0, -75, 5, 0, 0, 170.0); b = macros.last_csv()
def body(p): return [l for l in open(p) if not l.startswith("#")]
print("devices", n, "rows identical:", body(a) == body(b), a, b)
PY
cat $out/${tag}_macro_n${N}.log | tail -2
python - <<PY
import json
for f in ("bench","bench_exact"):
    try:
        j=json.load(open("$out/${tag}_%s_n$N.json"%f))
        print(f, "N", j["n_gpus"], j["scaling"], "value %.4g"%j["value"], "e2e %.4g"%j["e2e"]["value"], "ms/step %.1f"%j["ms_per_step"], "crc", j["map_crc"], "other", j.get("other_scaling"), "inproc", j.get("inproc_context"))
    except Exception as e: print(f, "failed", e)
PY
