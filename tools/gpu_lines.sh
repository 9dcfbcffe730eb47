#!/bin/bash
# One full ncu capture of a workload's cast kernel, summarised on the box: raw-page summary + instructions / stalls by source line.
# Usage: bash tools/gpu_lines.sh <tag> <workload> [kernel regex] [lib variant]
TAG=$1; W=$2; K=${3:-cast_kernel}; V=${4:-default}
OUT=gpurun_out/$TAG; mkdir -p $OUT
LIB=$PWD/yt_b200/libytb200.so
if [ "$V" != default ]; then export YTB200_LIB=$PWD/yt_b200/libytb200_$V.so; LIB=$YTB200_LIB; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $OUT/prof_${W}_$V \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > $OUT/ncu_${W}_$V.log 2>&1
python tools/ncu_summary.py $OUT/prof_${W}_$V.ncu-rep > $OUT/${W}_${V}_summary.txt
python tools/ncu_lines.py $OUT/prof_${W}_$V.ncu-rep $LIB 80 > $OUT/${W}_${V}_lines.txt
[ -n "$KEEP_REP" ] || rm -f $OUT/prof_${W}_$V.ncu-rep
head -30 $OUT/${W}_${V}_lines.txt
