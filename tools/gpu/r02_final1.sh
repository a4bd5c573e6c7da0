# round 2, final single-GPU run: full GPU test suite, smoke, bench at configs 2 (default line), 3 and 5, reference arm,
# launch list of the default bench, full ncu captures of the hot kernels of the three stages
python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_final_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench_c2_n1.json 2> gpurun_out/r02_final_bench_c2_n1.err; echo bench rc=$?
for c in 3 5; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg --no-dropin-leg --no-outgroup --no-strong-c4 > gpurun_out/r02_final_bench_c${c}_n1.json 2> gpurun_out/r02_final_bench_c${c}_n1.err; echo bench c$c rc=$?; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err; echo ref rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fasta-leg --no-strong-c4 > gpurun_out/r02_final_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fasta-leg --no-strong-c4 > gpurun_out/r02_final_ncu_launches.log 2>&1; echo ncu-launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_cand1|k_scan_filtered|k_scan_plain|k_pick|k_aliveN|k_build_at|k_finish|k_id_records' --launch-skip 8 --launch-count 9 -o gpurun_out/r02_final_prof -f python tools/prof_one.py 2 > gpurun_out/r02_final_ncu_full.log 2>&1; echo ncu-full rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_og_' --launch-skip 7 --launch-count 5 -o gpurun_out/r02_final_og_prof -f python tools/og_bench.py 10 5000000 50000 > gpurun_out/r02_final_ncu_og.log 2>&1; echo ncu-og rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_pp_' --launch-skip 9 --launch-count 9 -o gpurun_out/r02_final_pp_prof -f python tools/pp_bench.py 10 5000000 > gpurun_out/r02_final_ncu_pp.log 2>&1; echo ncu-pp rc=$?
python - <<PY
import json
for c in (2, 3, 5):
    try:
        d = json.loads(open(f"gpurun_out/r02_final_bench_c{c}_n1.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["phases_ms"], round(d["roofline"]["ms_per_launch"], 3), d["gpu_launches"] / d["steps"],
              d.get("parity", {}).get("ok"), d.get("outgroup", {}).get("ms_per_step"), d.get("outgroup", {}).get("parity", {}).get("ok"), d.get("pairs", {}).get("ms_per_step"), d.get("pairs", {}).get("parity", {}).get("ok"))
    except Exception as e:
        print(c, "ERR", e)
PY
