mkdir -p gpurun_out/r02ze
export YTB200_LIB=$PWD/yt_b200/libytb200_tex.so
timeout 600 python -m pytest tests/test_fullsize_oracle_gpu.py tests/test_parity_gpu.py -m gpu -q -s -k "north_star or cfg2 or test_fast_path_matches_oracle" 2>&1 | grep "vs reference\|passed\|failed" > gpurun_out/r02ze/pytest_tex.log; cat gpurun_out/r02ze/pytest_tex.log
for v in tex default; do
  if [ "$v" = default ]; then unset YTB200_LIB; fi
  for W in northstar_vr_1024c_2048px cfg2_vr_512c_1024px; do
    timeout 300 python bench.py --workload $W --no-cpu --steps 10 > gpurun_out/r02ze/bench_${v}_$W.json 2> gpurun_out/r02ze/bench_${v}_$W.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02ze/bench_${v}_$W.json').read().strip().splitlines()[-1]); print('$v $W N=1 step %.3f kernel %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02ze/bench_${v}_$W.err
  done
done
