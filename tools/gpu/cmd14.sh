python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
for c in 2 3 4; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench36_c$c.json 2> gpurun_out/bench36_c$c.err; echo rc=$?; done
python - <<PY
import json
for c in (2,3,4):
    d=json.load(open(f"gpurun_out/bench36_c{c}.json"))
    print(c, round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"]["ms_per_step"])
PY
