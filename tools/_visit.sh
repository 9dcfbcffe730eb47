mkdir -p gpurun_out/r02zf
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02zf/pytest_gpu.log; cat gpurun_out/r02zf/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r02zf/bench_default.json 2> gpurun_out/r02zf/bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02zf/bench_default.json').read().strip().splitlines()[-1]); print('default bench: steps %d step %.3f e2e %.3f kernel %.3f launches %d clocks %s' % (d['steps'], d['ms_per_step'], d['e2e']['ms_per_frame'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks']))"
timeout 300 python tools/bench_scene.py > gpurun_out/r02zf/bench_scene.json 2> gpurun_out/r02zf/bench_scene.err; cat gpurun_out/r02zf/bench_scene.json
