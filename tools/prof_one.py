"""one warm-up + one profiled pass of the stage on config 2 (device-resident inputs) -- for ncu captures"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primerforge_b200 import synth
from primerforge_b200.engine import Engine

cfg_i = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gens, cfg = synth.make_config(cfg_i)
eng = Engine(device=0, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
for g in gens:
    eng.add_genome(g.contigs)
eng.upload()
for _ in range(2):
    eng.compute()
print(eng.timing())
