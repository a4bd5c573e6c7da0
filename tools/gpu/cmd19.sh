for c in 2 3; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench41_c$c.json 2> gpurun_out/bench41.err; echo rc=$?; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench41_*.json")):
    d=json.load(open(f))
    print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],3), d["phases_ms"], d["roofline"]["ms_per_launch"])
PY
