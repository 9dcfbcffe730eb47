mkdir -p gpurun_out/r02t
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02t/pytest_gpu.log; cat gpurun_out/r02t/pytest_gpu.log
for W in iproj_1024c_2048px northstar_vr_1024c_2048px; do
  timeout 300 python bench.py --workload $W --no-cpu --steps 5 > gpurun_out/r02t/bench_$W.json 2> gpurun_out/r02t/bench_$W.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02t/bench_$W.json').read().strip().splitlines()[-1]); print('$W N=1 step %.3f e2e %.3f kernel %.3f' % (d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02t/bench_$W.err
done
