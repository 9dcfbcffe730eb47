#!/bin/bash
# Kernel experiment visit: parity tests, then the north-star / cfg3 bench for each library variant.
# Usage: bash tools/gpu_exp.sh <tag> [variant tags...]
TAG=${1:-exp}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -n "$SKIPTESTS" ] || timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
for v in default "$@"; do
  if [ "$v" = default ]; then unset YTB200_LIB; else export YTB200_LIB=$PWD/yt_b200/libytb200_$v.so; fi
  for w in ${WL:-northstar_vr_1024c_2048px cfg3_proj_1024c_2048px}; do
    timeout 300 python bench.py --workload $w --no-cpu --steps 5 > $OUT/bench_${v}_$w.json 2> $OUT/bench_${v}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${v}_$w.json").read().strip().splitlines()[-1])
    print("$v $w ms/frame %.3f  value %.1f  e2e_ms %.3f kernel_ms %.3f clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_frame"], d["roofline"]["kernel_ms"], d["clocks"]))
except Exception as e:
    print("$v $w FAILED", e); print(open("$OUT/bench_${v}_$w.err").read()[-1500:])
PY
  done
done
