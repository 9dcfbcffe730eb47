#!/bin/bash
# A/B visit for a kernel variant: parity tests WITH the variant library, then the benches for default and variant.
# Usage: bash tools/gpu_ab.sh <tag> <variant>   (WL="workloads..." to override)
TAG=$1; V=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
YTB200_LIB=$PWD/yt_b200/libytb200_$V.so timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$V.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$V.log
tail -15 $OUT/pytest_gpu_$V.log
SKIPTESTS=1 bash tools/gpu_exp.sh $TAG $V
