mkdir -p gpurun_out/r02zd
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02zd/pytest_gpu.log; cat gpurun_out/r02zd/pytest_gpu.log
for v in bands nobands; do
  if [ "$v" = nobands ]; then export YTB200_NO_BANDS=1; else unset YTB200_NO_BANDS; fi
  timeout 300 python bench.py --workload cfg3_proj_1024c_2048px --no-cpu --steps 10 > gpurun_out/r02zd/bench_${v}_cfg3.json 2> gpurun_out/r02zd/bench_${v}_cfg3.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02zd/bench_${v}_cfg3.json').read().strip().splitlines()[-1]); print('$v cfg3 N=1 step %.3f e2e %.3f kernel %.3f' % (d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02zd/bench_${v}_cfg3.err
done
