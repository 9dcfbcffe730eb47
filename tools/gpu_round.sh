#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list and one full capture of the cast kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_northstar.json 2> $OUT/bench_northstar.err; tail -c 3000 $OUT/bench_northstar.json
timeout 300 python bench.py --workload cfg3_proj_1024c_2048px --no-cpu > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; tail -c 1500 $OUT/bench_cfg3.json
timeout 300 python bench.py --workload cfg2_vr_512c_1024px --no-cpu > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; tail -c 1500 $OUT/bench_cfg2.json
timeout 300 python bench.py --workload cfg4_amr_vr_15kbricks_persp_2048px --no-cpu > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; tail -c 1500 $OUT/bench_cfg4.json
timeout 300 python bench.py --workload cfg4_amr_vr_24kbricks_persp_2048px --no-cpu > $OUT/bench_cfg4_24k.json 2> $OUT/bench_cfg4_24k.err; tail -c 600 $OUT/bench_cfg4_24k.json
timeout 300 python bench.py --workload iproj_1024c_2048px --no-cpu > $OUT/bench_iproj.json 2> $OUT/bench_iproj.err; tail -c 600 $OUT/bench_iproj.json
timeout 300 python tools/bench_ert.py > $OUT/bench_ert.json 2> $OUT/bench_ert.err; cat $OUT/bench_ert.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 1000 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cast_kernel|proj_kernel|count_kernel|extent_kernel|fill_f64|compose_kernel|apply_kernel|f64_to_f32|swap_' -c 400 --csv --log-file $OUT/launches_northstar.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1
# full captures are summarised on the box and deleted there: one report is ~35 MB and gpurun merges back at most 64 MiB
# (set KEEP_REP=1 to bring a report back for the source / SASS pages)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cast_kernel -s 2 -c 1 -o $OUT/prof_vr \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_vr.log 2>&1
python tools/ncu_summary.py $OUT/prof_vr.ncu-rep > $OUT/vr_northstar_cast_kernel.txt 2>> $OUT/ncu_full_vr.log
python tools/ncu_lines.py $OUT/prof_vr.ncu-rep yt_b200/libytb200.so 60 > $OUT/vr_northstar_cast_kernel_by_source_line.txt 2>> $OUT/ncu_full_vr.log
[ -n "$KEEP_REP" ] || rm -f $OUT/prof_vr.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj_kernel -s 2 -c 1 -o $OUT/prof_proj \
    python bench.py --workload cfg3_proj_1024c_2048px --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_proj.log 2>&1
python tools/ncu_summary.py $OUT/prof_proj.ncu-rep > $OUT/proj_cfg3_cast_kernel.txt 2>> $OUT/ncu_full_proj.log
rm -f $OUT/prof_proj.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cast_kernel -s 2 -c 1 -o $OUT/prof_cfg4 \
    python bench.py --workload cfg4_amr_vr_15kbricks_persp_2048px --steps 1 --warmup 3 --no-cpu > $OUT/ncu_full_cfg4.log 2>&1
python tools/ncu_summary.py $OUT/prof_cfg4.ncu-rep > $OUT/vr_cfg4_cast_kernel.txt 2>> $OUT/ncu_full_cfg4.log
python tools/ncu_lines.py $OUT/prof_cfg4.ncu-rep yt_b200/libytb200.so 60 > $OUT/vr_cfg4_cast_kernel_by_source_line.txt 2>> $OUT/ncu_full_cfg4.log
[ -n "$KEEP_REP" ] || rm -f $OUT/prof_cfg4.ncu-rep
ls -la $OUT
