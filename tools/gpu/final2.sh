python bench.py > gpurun_out/final2_c2_n1.json 2> gpurun_out/final2_c2_n1.err; echo bench default rc=$?
for c in 3 4; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/final2_c${c}_n1.json 2> gpurun_out/final2.err; echo bench c$c rc=$?; done
python - <<PY
import json
for c in (2,3,4):
    d=json.load(open(f"gpurun_out/final2_c{c}_n1.json"))
    print(c, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], round(d["roofline"]["frac"],3), d["gpu_launches"]/d["steps"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches15.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo ncu2 rc=$?
