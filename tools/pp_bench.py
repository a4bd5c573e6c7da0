"""python tools/pp_bench.py [n_genomes] [length]: candidate stage -> pair search on BASELINE config 2's shape (timings + oracle check)"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np
from primerforge_b200 import synth
from primerforge_b200.candidates import _getAllCandidateKmers
from primerforge_b200.pairs import PairSearch, _columns

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
check = len(sys.argv) > 3
gens = synth.make_genomes(G, length, seed=synth.BASE_SEED + 2)
log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
with tempfile.TemporaryDirectory() as d:
    paths = synth.write_fasta(gens, d)
    prm = SimpleNamespace(ingroupFns=paths, format="fasta", minLen=16, maxLen=20, minGc=40.0, maxGc=60.0, minTm=55.0, maxTm=68.0,
                          maxRepeatLen=4, tempTolerance=5.0, numThreads=1, p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6,
                          p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log, pickles={1: "s.p", 2: "c.p"},
                          dumpObj=lambda *a, **k: None, loadObj=lambda *a: None, debug=False)
    lazy = _getAllCandidateKmers(prm, False, thal="off")
t0 = time.time()
word, k, tm, names, per_genome, first_lists = _columns(lazy, gens[0].name)
t1 = time.time()
kw = dict(max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0)
with PairSearch(0) as ps:
    ps.set_candidates(word, k, tm)
    for g in per_genome:
        ps.add_genome(*g)
    for _ in range(3):
        t2 = time.time()
        ps.set_candidates(word, k, tm)
        for g in per_genome:
            ps.add_genome(*g)
        ps.run(**kw); t3 = time.time()
    res = ps.result()
    print({**ps.timing(), "columns_s": t1 - t0, "run_wall_s": t3 - t2, "kept": int(res.kept.sum())})
if check:
    from oracle import pf_oracle as po
    t4 = time.time(); o = po.run_pairs(word, k, tm, per_genome, **kw); t5 = time.time()
    p = res.pairs
    pairs = np.stack([p["fwd"], p["rev"], p["contig"], p["fbin"], p["rbin"], p["pcr_len"]], axis=1).astype(np.uint32)
    shared = res.shared4()
    print("oracle s", t5 - t4, "pairs equal", np.array_equal(pairs, o.pairs), "shared equal", np.array_equal(shared, o.shared))
