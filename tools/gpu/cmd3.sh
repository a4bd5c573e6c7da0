python -m pytest tests/test_gpu_fasta.py -x -q > gpurun_out/pytest_fasta.log 2>&1; echo fasta rc=$?; tail -25 gpurun_out/pytest_fasta.log | cut -c1-400
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fasta.py > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --config 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench32_c2.json 2> gpurun_out/bench32_c2.err; echo bench rc=$?; tail -3 gpurun_out/bench32_c2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench32_c2.json"))
print(round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), "fasta", d.get("e2e_fasta"))
PY
