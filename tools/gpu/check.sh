python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for c in 2 3; do python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench47_c$c.json 2> gpurun_out/bench47.err; echo rc=$?; done
python - <<PY
import json
for c in (2,3):
    d=json.load(open(f"gpurun_out/bench47_c{c}.json"))
    print(c, round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"]["ms_per_step"], d["gpu_launches"]/d["steps"])
PY
