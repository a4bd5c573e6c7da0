# round 2, 2 GPUs: the multi-GPU parity tests (both transports, both drivers) and the default bench line at N = 2
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_r_pytest_multi.log 2>&1; echo multi rc=$?; tail -4 gpurun_out/r02_r_pytest_multi.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_r_bench_n2.json 2> gpurun_out/r02_r_bench_n2.err; echo bench2 rc=$?
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_r_bench_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["phases_ms"], d.get("exchange", {}).get("ms_per_step"), d.get("parity", {}).get("ok"), d["gpu_launches"] / d["steps"],
      {k: d.get("strong_c4", {}).get(k) for k in ("ms_per_step", "ms_per_step_one_gpu", "speedup_vs_one_gpu", "ms_exchange_rank0")})
PY
