# round 2, 8 GPUs: the default bench line (weak config 2 + strong_c4 block + parity), config 3 strong, timeline of config 4 strong
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_p_bench_n$N.json 2> gpurun_out/r02_p_bench_n$N.err; echo bench rc=$?
$TR bench.py --gpus $N --config 3 --scaling strong --steps 20 --warmup 3 --no-strong-c4 > gpurun_out/r02_p_c3_strong_n$N.json 2> gpurun_out/r02_p_c3_strong_n$N.err; echo c3 rc=$?
PFB200_TRACE=1 $TR tools/trace_mgpu.py 4 strong > gpurun_out/r02_p_trace_c4_n$N.out 2> gpurun_out/r02_p_trace_c4_n$N.txt; echo trace rc=$?
python - <<PY
import json
for f in ("gpurun_out/r02_p_bench_n$N.json", "gpurun_out/r02_p_c3_strong_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), d["phases_ms"], d.get("exchange", {}).get("ms_per_step"), d.get("parity", {}).get("ok"),
              {k: d.get("strong_c4", {}).get(k) for k in ("ms_per_step", "ms_per_step_one_gpu", "speedup_vs_one_gpu", "ms_exchange_rank0")})
    except Exception as e:
        print(f, "ERR", e)
PY
