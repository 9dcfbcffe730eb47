mkdir -p gpurun_out/r02zc
for v in base rowsonly both; do
  if [ "$v" = both ]; then unset YTB200_LIB; else export YTB200_LIB=$PWD/yt_b200/libytb200_$v.so; fi
  for W in northstar_vr_1024c_2048px cfg4_amr_vr_15kbricks_persp_2048px; do
    timeout 300 python bench.py --workload $W --no-cpu --steps 10 > gpurun_out/r02zc/bench_${v}_$W.json 2> gpurun_out/r02zc/bench_${v}_$W.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02zc/bench_${v}_$W.json').read().strip().splitlines()[-1]); print('$v $W N=1 step %.3f kernel %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02zc/bench_${v}_$W.err
  done
done
