python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke_final.log
python bench.py > gpurun_out/final_c2_n1.json 2> gpurun_out/final_c2_n1.err; echo bench default rc=$?
for c in 3 4 5; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/final_c${c}_n1.json 2> gpurun_out/final_c${c}_n1.err; echo bench c$c rc=$?; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo ref rc=$?
python - <<PY
import json
for c in (2,3,4,5):
    try:
        d=json.load(open(f"gpurun_out/final_c{c}_n1.json"))
        print(c, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], round(d["roofline"]["frac"],3), round(d["roofline"]["ms_per_launch"],3), d.get("e2e_fasta",{}).get("ms_per_step"), d["clocks"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(c, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:'k_cand1|k_scan_filtered|k_scan_plain|k_positions|k_aliveN|k_encode|k_compact' --launch-skip 6 --launch-count 8 -o gpurun_out/prof_v14 -f python tools/prof_one.py 2 > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches14.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo ncu2 rc=$?
