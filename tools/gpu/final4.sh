python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_final.log 2>&1; echo multi rc=$?; tail -3 gpurun_out/pytest_multi_final.log | cut -c1-600
for n in 2 4; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/final_c2_n$n.json 2> gpurun_out/final_c2_n$n.err; echo bench n$n rc=$?; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 4 --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref_n4.json 2> gpurun_out/final_ref_n4.err; echo ref4 rc=$?; head -c 300 gpurun_out/final_ref_n4.json; echo
python - <<PY
import json
for n in (2,4):
    d=json.load(open(f"gpurun_out/final_c2_n{n}.json"))
    print(n, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], d.get("exchange",{}).get("ms_per_step"))
PY
