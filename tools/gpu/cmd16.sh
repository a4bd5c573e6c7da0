for m in 1 2; do python bench.py --genomes 1 --steps 30 --warmup 3 --no-cpu-baseline --no-fasta-leg --scan-mode $m > gpurun_out/bench38_g1_m$m.json 2> gpurun_out/bench38_g1_m$m.err; echo rc=$?; done
for L in 2000000 10000000 20000000; do python bench.py --genomes 1 --length $L --steps 30 --warmup 3 --no-cpu-baseline --no-fasta-leg > gpurun_out/bench38_g1_L$L.json 2> gpurun_out/bench38_g1_L$L.err; echo rc=$?; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench38_g1_*.json")):
    d=json.load(open(f))
    print(f.split('/')[-1], round(d["ms_per_step"],3), d["phases_ms"], d["counts"]["n_slots"])
PY
