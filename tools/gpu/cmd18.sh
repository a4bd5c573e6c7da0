python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
python bench.py --genomes 1 --steps 30 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench40_g1.json 2> gpurun_out/bench40.err; echo rc=$?
for c in 2 5; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench40_c$c.json 2> gpurun_out/bench40.err; echo rc=$?; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench40_*.json")):
    d=json.load(open(f))
    print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["e2e"]["ms_per_step"])
PY
