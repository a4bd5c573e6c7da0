PFB200_TRACE=1 python tools/trace_run.py 2 2> gpurun_out/trace_c2.log; echo rc=$?; tail -24 gpurun_out/trace_c2.log | grep -v "copy landed: genome [2-8]"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench46_c2.json 2> gpurun_out/bench46.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench46_c2.json"))
print(round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"], d["e2e_fasta"]["ms_per_step"])
PY
