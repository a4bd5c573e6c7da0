python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo multi rc=$?; tail -5 gpurun_out/pytest_multi.log | cut -c1-800
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench35_n2.json 2> gpurun_out/bench35_n2.err; echo bench2 rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench35_n2.json"))
print(round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], d.get("exchange",{}).get("ms_per_step"))
PY
