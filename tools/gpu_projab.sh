#!/bin/bash
# A/B of projection-kernel variants on cfg3.  Usage: bash tools/gpu_projab.sh <tag> <variant> [<variant> ...]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for V in "$@"; do
  for M in 1 0; do
    YTB200_PROJ_MAP=$M YTB200_LIB=$PWD/yt_b200/libytb200_$V.so python bench.py --workload cfg3_proj_1024c_2048px --steps 10 --warmup 3 --no-cpu 2>>$OUT/err.log > $OUT/cfg3_${V}_map$M.json
    python - <<PY
import json
try:
    d = json.load(open("$OUT/cfg3_${V}_map$M.json"))
    print("%-8s map$M  kernel %.3f ms  step %.3f ms  hbm_frac %.3f" % ("$V", d["roofline"]["kernel_ms"], d["ms_per_step"], d["roofline"]["hbm_frac"]))
except Exception as e:
    print("$V map$M failed", e)
PY
  done
done
