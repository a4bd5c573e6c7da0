python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench45_c2.json 2> gpurun_out/bench45.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench45_c2.json"))
print(round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"]["ms_per_step"], d["e2e_fasta"]["ms_per_step"])
PY
