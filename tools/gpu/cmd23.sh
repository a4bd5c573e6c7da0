python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench43_c2.json 2> gpurun_out/bench43.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench43_c2.json"))
print(round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"], d["e2e_fasta"]["ms_per_step"])
PY
python -m pytest tests/test_gpu_parity.py -x -q -k "pipelined or drop_in or picked" > gpurun_out/pytest_p.log 2>&1; echo rc=$?; tail -2 gpurun_out/pytest_p.log
