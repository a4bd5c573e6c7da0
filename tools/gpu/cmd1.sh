python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
for c in 2 3; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench21_c$c.json 2> gpurun_out/bench21_c$c.err; echo bench c$c rc=$?; done
python - <<PY
import json
for c in (2,3):
    try:
        d=json.load(open(f"gpurun_out/bench21_c{c}.json"))
        print(c, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], d["counts"], d["roofline"]["ms_per_launch"])
    except Exception as e:
        print(c, "ERR", e)
PY
