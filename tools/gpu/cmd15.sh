python -m pytest tests/test_gpu_parity.py tests/test_gpu_fasta.py -x -q -k "drop_in" > gpurun_out/pytest_dropin.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_dropin.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dropin-leg > gpurun_out/bench37_dropin.json 2> gpurun_out/bench37_dropin.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench37_dropin.json"))
print(d.get("e2e_dropin"))
PY
tail -8 gpurun_out/bench37_dropin.err
