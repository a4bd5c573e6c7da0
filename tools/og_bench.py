"""python tools/og_bench.py [n_outgroup] [length] [n_pairs]: the outgroup screen on BASELINE config 3's outgroup shape"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from primerforge_b200 import synth
from primerforge_b200.outgroup import OutgroupScreen

n_og = int(sys.argv[1]) if len(sys.argv) > 1 else 10
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
n_pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 500_000
first = synth.make_genomes(1, length, seed=synth.BASE_SEED + 3)[0]
outs = synth.make_outgroup(first, n_og)
w, k, pf, pr = synth.make_pairs(first, n_pairs)
with OutgroupScreen(0) as scr:
    scr.set_keys_and_pairs(w, k, pf, pr, 120, 2400)
    for o in outs:
        scr.add_genome(o.contigs)
    t0 = time.time(); scr.run(); t1 = time.time()
    for _ in range(3):
        scr.compute()
    t = scr.timing()
    res = scr.result(with_hits=False)
    print({**t, "wall_run_s": t1 - t0, "n_keys": int(w.size), "n_pairs": n_pairs, "kept": int(res.kept.sum()),
           "Mbp_per_s_device": t["n_bases"] / t["ms_total"] / 1e3})
