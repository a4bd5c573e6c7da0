# round 2, second check: full GPU test suite, smoke, default bench line (with the outgroup and pair-search blocks),
# launch list of the whole default bench, full ncu capture of the outgroup scan and the pair-search kernels
python -m pytest tests -m gpu -q > gpurun_out/r02_n_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_n_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_n_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r02_n_smoke.log
python bench.py > gpurun_out/r02_n_bench_c2.json 2> gpurun_out/r02_n_bench_c2.err; echo bench rc=$?
python tools/og_bench.py 10 5000000 50000 > gpurun_out/r02_n_og.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_og_scan|k_og_encode|k_og_products' --launch-skip 3 --launch-count 3 -o gpurun_out/r02_n_og_prof -f python tools/og_bench.py 10 5000000 50000 > gpurun_out/r02_n_ncu_og.log 2>&1; echo ncu-og rc=$?
