"""python tools/profile_scan.py [config] [n_compute]: a few resident steps of one config (the thing to put under ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primerforge_b200 import synth
from primerforge_b200.engine import Engine

cfg_i = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gens, cfg = synth.make_config(cfg_i)
with Engine(device=0, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1) as eng:
    for g in gens:
        eng.add_genome(g.contigs)
    eng.upload()
    for _ in range(n):
        eng.compute()
    t = eng.timing()
    print({k: t[k] for k in ("ms_total", "ms_scan_main", "n_launches", "n_ids", "n_records", "n_picked", "n_reruns")})
