# first-run diagnostic: run one golden case, print intermediate counts
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from tests.util import load_golden, golden_cases
from oracle import pf_oracle as po
from primerforge_b200.engine import Engine
for name in golden_cases():
    gold = load_golden(name)
    kw = dict(k_min=gold.k_min, k_max=gold.k_max, min_gc=gold.min_gc, max_gc=gold.max_gc, min_tm=gold.min_tm, max_tm=gold.max_tm, max_repeat=gold.max_repeat, table_size=gold.table_size)
    o = po.run(gold.genomes, po.default_params(**kw))
    for pf in (0, 1):
        eng = Engine(device=0, prefilter=pf, **kw); eng.set_debug_counts(True)
        for g in gold.genomes: eng.add_genome(g.contigs)
        eng.run()
        n, m = eng.counts(); t = eng.timing()
        print(name, "prefilter", pf, "allowed gpu", eng.first_genome_allowed().tolist(), "oracle", o.n_allowed[0].tolist())
        print("   slots", t["n_slots"], "records", n, "picked", m, "oracle shared", o.n_shared, "oracle pass", int(((o.flags&7)==7).sum()), "oracle picked", int(((o.flags&8)!=0).sum()))
        if pf == 0 and n:
            rec = eng.records()
            same = rec.size == o.n_shared and np.array_equal(rec["word"], o.word)
            print("   words equal", same)
            if not same:
                a = set(zip(rec["word"].tolist(), rec["k"].tolist())); b = set(zip(o.word.tolist(), o.k.tolist()))
                print("   only gpu", len(a-b), "only oracle", len(b-a))
        eng.close()
