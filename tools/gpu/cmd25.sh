python -m pytest tests/test_gpu_parity.py tests/test_gpu_fasta.py -x -q -k "id_list or tiny_records" > gpurun_out/pytest_new.log 2>&1; echo rc=$?; tail -15 gpurun_out/pytest_new.log | cut -c1-300
