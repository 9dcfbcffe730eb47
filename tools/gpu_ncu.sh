#!/bin/bash
# One full ncu capture of the cast kernel of a workload.  Usage: bash tools/gpu_ncu.sh <tag> <workload> [lib variant]
TAG=$1; W=$2; V=${3:-default}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$V" != default ]; then export YTB200_LIB=$PWD/yt_b200/libytb200_$V.so; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cast_kernel -s 2 -c 1 -o $OUT/prof_${W}_$V \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > $OUT/ncu_${W}_$V.log 2>&1
tail -3 $OUT/ncu_${W}_$V.log
