N=${1:-8}
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo $name rc=$?
}
run bench44_weak_c2_n$N --steps 50 --warmup 5
run bench44_strong_c4_n$N --config 4 --scaling strong --steps 20 --warmup 3
run bench44_strong_c3_n$N --config 3 --scaling strong --steps 20 --warmup 3
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench44_*_n$N.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["phases_ms"], d.get("exchange",{}).get("ms_per_step"), d.get("host_cpus_bound"))
    except Exception as e:
        print(f, "ERR", e)
PY
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -i -E "numa|^CPU\(s\)|model name" >> gpurun_out/topo.txt
