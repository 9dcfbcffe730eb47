mkdir -p gpurun_out/r02y
for v in default ip3 ip4; do
  if [ "$v" = default ]; then unset YTB200_LIB; else export YTB200_LIB=$PWD/yt_b200/libytb200_$v.so; fi
  timeout 300 python bench.py --workload iproj_1024c_2048px --no-cpu --steps 10 > gpurun_out/r02y/bench_${v}_iproj.json 2> gpurun_out/r02y/bench_${v}_iproj.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02y/bench_${v}_iproj.json').read().strip().splitlines()[-1]); print('$v iproj N=1 step %.3f e2e %.3f kernel %.3f' % (d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms']))" || tail -5 gpurun_out/r02y/bench_${v}_iproj.err
done
unset YTB200_LIB
bash tools/gpu_lines.sh r02y iproj_1024c_2048px cast_kernel
