for m in 0 2; do python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline --no-fasta-leg --scan-mode $m > gpurun_out/bench33_c5_m$m.json 2> gpurun_out/bench33_c5_m$m.err; echo rc=$?; done
python - <<PY
import json
for m in (0,2):
    d=json.load(open(f"gpurun_out/bench33_c5_m{m}.json"))
    print(m, round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
PY
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
