mkdir -p gpurun_out/r02za
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -22 > gpurun_out/r02za/pytest_gpu.log; cat gpurun_out/r02za/pytest_gpu.log
for v in line noline; do
  if [ "$v" = noline ]; then export YTB200_NO_LINE=1; else unset YTB200_NO_LINE; fi
  for W in northstar_vr_1024c_2048px cfg2_vr_512c_1024px cfg4_amr_vr_15kbricks_persp_2048px; do
    timeout 300 python bench.py --workload $W --no-cpu --steps 10 > gpurun_out/r02za/bench_${v}_$W.json 2> gpurun_out/r02za/bench_${v}_$W.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02za/bench_${v}_$W.json').read().strip().splitlines()[-1]); print('$v $W N=1 step %.3f e2e %.3f kernel %.3f' % (d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02za/bench_${v}_$W.err
  done
done
