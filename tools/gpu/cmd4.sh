python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo multi rc=$?; tail -5 gpurun_out/pytest_multi.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench22_n2.json 2> gpurun_out/bench22_n2.err; echo bench2 rc=$?; tail -3 gpurun_out/bench22_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config 3 --scaling strong --steps 20 --warmup 3 > gpurun_out/bench22_c3_strong_n2.json 2> gpurun_out/bench22_c3_strong_n2.err; echo bench3 rc=$?; tail -3 gpurun_out/bench22_c3_strong_n2.err
python - <<PY
import json
for f in ("bench22_n2","bench22_c3_strong_n2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], d.get("exchange"))
    except Exception as e:
        print(f, "ERR", e)
PY
