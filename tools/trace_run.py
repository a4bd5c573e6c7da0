"""PFB200_TRACE=1 python tools/trace_run.py [config]: timeline of one pipelined run (host buffers -> results)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from primerforge_b200 import synth
from primerforge_b200.engine import Engine

cfg_i = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gens, cfg = synth.make_config(cfg_i)
pinned = []
eng = Engine(device=0, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
for g in gens:
    cat = np.concatenate(g.contigs) if len(g.contigs) > 1 else g.contigs[0]
    t = torch.empty(cat.size, dtype=torch.uint8).pin_memory()
    t.numpy()[:] = cat
    off = np.zeros(len(g.contigs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([c.size for c in g.contigs])
    pinned.append((t, off))
    eng.add_genome_ptr(t.data_ptr(), t.numel(), off)
for i in range(4):
    print(f"---- run {i}", file=sys.stderr)
    t0 = time.perf_counter()
    eng.run()
    t1 = time.perf_counter()
    eng.result_by_id()
    eng.sync()
    t2 = time.perf_counter()
    print(f"host: run {1e3 * (t1 - t0):.3f} ms, fetch {1e3 * (t2 - t1):.3f} ms", file=sys.stderr)
