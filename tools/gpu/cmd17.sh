for L in 2000000 10000000 20000000; do python bench.py --genomes 1 --length $L --steps 30 --warmup 3 --no-cpu-baseline --no-fasta-leg --scan-mode 1 > gpurun_out/bench39_g1_L${L}_m1.json 2> gpurun_out/bench39.err; echo rc=$?; done
python bench.py --config 5 --genomes 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fasta-leg --scan-mode 1 > gpurun_out/bench39_c5g1_m1.json 2> gpurun_out/bench39.err; echo rc=$?
python bench.py --config 5 --genomes 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fasta-leg --scan-mode 2 > gpurun_out/bench39_c5g1_m2.json 2> gpurun_out/bench39.err; echo rc=$?
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench39_*.json")):
    d=json.load(open(f))
    print(f.split('/')[-1], round(d["ms_per_step"],3), d["phases_ms"], d["counts"]["n_slots"])
PY
