#!/bin/bash
# Multi-GPU visit: parity tests + bench lines.  Usage: bash tools/gpu_multi.sh <tag> <N> <workload> [<workload> ...]
TAG=$1; N=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L | wc -l
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -rs 2>&1 | tail -6 > $OUT/mgpu_test.log; cat $OUT/mgpu_test.log
for W in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --workload $W > $OUT/bench_n${N}_$W.out 2> $OUT/bench_n${N}_$W.err
  grep '^{' $OUT/bench_n${N}_$W.out | tail -1 > $OUT/bench_n${N}_$W.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_n${N}_$W.json"))
    print("$W N=$N step %.3f ms  e2e %.3f ms  cast kernel %.3f ms  parity_vs_n1 %s  launches/frame %.1f  %s" % (
        d["ms_per_step"], d["e2e"]["ms_per_frame"], d["roofline"]["kernel_ms"], d.get("parity_vs_n1", {}).get("value"), d["gpu_launches"] / d["steps"], d["config"]["composite"]))
except Exception as e:
    print("$W failed:", e)
    print(open("$OUT/bench_n${N}_$W.err").read()[-1500:])
PY
done
