mkdir -p gpurun_out/r02s
timeout 1200 python -m pytest tests/test_fullsize_oracle_gpu.py tests/test_ert_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02s/pytest_gpu.log; cat gpurun_out/r02s/pytest_gpu.log
true
