# round 2: full GPU test suite, smoke, default bench line, launch list, one full ncu capture of the hot kernels
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_m_bench_c2.json 2> gpurun_out/r02_m_bench_c2.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_m_ref.json 2> gpurun_out/r02_m_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fasta-leg --no-dropin-leg --no-strong-c4 > gpurun_out/r02_m_ncu.log 2>&1; echo ncu-launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_cand1|k_scan_filtered|k_scan_plain|k_positions|k_aliveN|k_build_at' --launch-skip 6 --launch-count 8 -o gpurun_out/r02_m_prof -f python tools/prof_one.py 2 > gpurun_out/r02_m_ncu_full.log 2>&1; echo ncu-full rc=$?
