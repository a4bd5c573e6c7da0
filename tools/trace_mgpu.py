"""PFB200_TRACE=1 torchrun ... tools/trace_mgpu.py [config] [strong]: timeline of resident steps on N ranks (weak shape of
bench.py; "strong": the config's genomes dealt over the ranks like bench.py --scaling strong)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from primerforge_b200 import synth
from primerforge_b200.dist import attach_comm, shard_genomes
from primerforge_b200.engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg_i = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synth.CONFIGS[cfg_i]
per = cfg["n_genomes"]
idx = [0] + list(range(1 + (per - 1) * rank, 1 + (per - 1) * (rank + 1)))
if len(sys.argv) > 2 and sys.argv[2] == "strong":
    idx = shard_genomes(per, world, rank)
gens = synth.make_genomes(max(idx) + 1, cfg["length"], seed=synth.BASE_SEED + cfg_i, n_contigs=cfg["n_contigs"], only=idx)
eng = Engine(device=local, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
attach_comm(eng)
for i in idx:
    eng.add_genome(gens[i].contigs)
eng.upload()
for step in range(5):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(f"---- step {step} rank {rank}", file=sys.stderr)
    eng.compute()
print(rank, eng.timing()["ms_total"], eng.timing()["ms_exchange"], file=sys.stderr)
dist.destroy_process_group()
