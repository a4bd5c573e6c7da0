python -m pytest tests/test_gpu_bench.py -x -q > gpurun_out/pytest_bench.log 2>&1; echo rc=$?; tail -8 gpurun_out/pytest_bench.log | cut -c1-400
