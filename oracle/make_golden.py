"""Generate tests/golden/*.npz by running the reference's UNMODIFIED
bin/getCandidateKmers.py (from /root/reference, present only in the build container) against
the stand-in dependencies (oracle/standins).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # rewrites every fixture

Each fixture stores the inputs (contigs as ASCII bytes), the stage parameters, the reference's
``sharedKmers`` dict (keys in insertion order, per-genome contig / start / strand) and the
reference's returned candidates (per genome: k-mer, contig, Primer.start, Primer.end, strand,
Tm, gcPer).  `table_size` != 9999991 means the khmer stand-in was forced to a small prime
(PF_STANDIN_TABLE) so that sketch collisions and junction pollution occur on kilobase inputs.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from primerforge_b200 import synth  # noqa: E402
from oracle.run_reference import run_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = {65: 84, 84: 65, 67: 71, 71: 67}


def _split(gens, seed, n_cuts):
    rng = np.random.default_rng(seed)
    for gen in gens:
        a = gen.contigs[0]
        cuts = np.sort(rng.choice(np.arange(1, a.size), size=n_cuts, replace=False))
        gen.contigs = [np.ascontiguousarray(x) for x in np.split(a, cuts)]
        gen.contig_ids = [f"{gen.name[:-6]}_c{i:03d}" for i in range(len(gen.contigs))]
    return gens


def _plant_palindromes(gens, seed, n):
    rng = np.random.default_rng(seed)
    length = gens[0].contigs[0].size
    for _ in range(n):
        k = int(rng.choice([16, 18, 20]))
        half = _ASCII[rng.integers(0, 4, k // 2)]
        pal = np.concatenate([half, np.array([_COMP[int(x)] for x in half[::-1]], dtype=np.uint8)])
        pos = int(rng.integers(50, length - 50))
        for gen in gens:
            gen.contigs[0][pos:pos + k] = pal
    return gens


def case_basic():
    return synth.make_genomes(3, 12000, seed=101, snp_rate=0.002, n_inversions=2), {}


def case_collide_multi():
    return (synth.make_genomes(4, 10000, seed=102, snp_rate=0.003, n_inversions=3, n_contigs=8),
            dict(table_size=20011, thal="pseudo"))


def case_palindrome():
    g = synth.make_genomes(3, 8000, seed=32, snp_rate=0.0, n_inversions=0)
    g = _plant_palindromes(g, 32, 150)
    return _split(g, 32, 60), dict(table_size=30011)


def case_widek():
    return (synth.make_genomes(3, 9000, seed=104, snp_rate=0.002, n_inversions=2, n_contigs=6),
            dict(table_size=65521, thal="pseudo", k_min=12, k_max=30, max_repeat=3))


def case_nbases():
    return (synth.make_genomes(3, 12000, seed=105, snp_rate=0.003, n_inversions=2, n_contigs=3, n_frac=0.001),
            dict(table_size=65521))


def case_tinycontigs():
    g = synth.make_genomes(3, 6000, seed=106, snp_rate=0.002, n_inversions=0)
    for gen in g:
        a = gen.contigs[0]
        gen.contigs = [a[:1000], a[1000:1007], a[1007:1010], a[1010:3000], a[3000:3019], a[3019:]]
        gen.contig_ids = [f"{gen.name[:-6]}_c{i:03d}" for i in range(6)]
    return g, dict(table_size=4093, k_min=10, k_max=32)


def case_tightfilters():
    return (synth.make_genomes(5, 9000, seed=107, snp_rate=0.004, n_inversions=4, n_contigs=2),
            dict(table_size=9999991, thal="pseudo", min_gc=45.0, max_gc=55.0, min_tm=56.0, max_tm=62.0))


CASES = {"basic": case_basic, "collide_multi": case_collide_multi, "palindrome": case_palindrome,
         "widek": case_widek, "nbases": case_nbases, "tinycontigs": case_tinycontigs,
         "tightfilters": case_tightfilters}


def make_one(name):
    gens, opt = CASES[name]()
    table = int(opt.get("table_size", 9999991))
    thal = opt.get("thal", "pass")
    par = dict(k_min=opt.get("k_min", 16), k_max=opt.get("k_max", 20), max_repeat=opt.get("max_repeat", 4),
               min_gc=opt.get("min_gc", 40.0), max_gc=opt.get("max_gc", 60.0), min_tm=opt.get("min_tm", 55.0),
               max_tm=opt.get("max_tm", 68.0), temp_tolerance=5.0)
    with tempfile.TemporaryDirectory() as d:
        paths = synth.write_fasta(gens, d)
        if table != 9999991:
            os.environ["PF_STANDIN_TABLE"] = str(table)
        else:
            os.environ.pop("PF_STANDIN_TABLE", None)
        shared, cands = run_reference(paths, thal_mode=thal, min_len=par["k_min"], max_len=par["k_max"],
                                      min_gc=par["min_gc"], max_gc=par["max_gc"], min_tm=par["min_tm"],
                                      max_tm=par["max_tm"], max_repeats=par["max_repeat"])
    names = [g.name for g in gens]
    keys = list(shared.keys())
    key_idx = {k: i for i, k in enumerate(keys)}
    out = dict(
        genome_names=np.array(names), table_size=np.int64(table), thal=np.array(thal),
        params=np.array([par["k_min"], par["k_max"], par["max_repeat"]], dtype=np.int64),
        fparams=np.array([par["min_gc"], par["max_gc"], par["min_tm"], par["max_tm"], par["temp_tolerance"]]),
        shared_kmers=np.array(keys, dtype="S32"))
    for gi, gen in enumerate(gens):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        out[f"g{gi}_contig_ids"] = np.array(gen.contig_ids)
        out[f"g{gi}_bases"] = np.concatenate(gen.contigs)
        out[f"g{gi}_offsets"] = np.cumsum([0] + [c.size for c in gen.contigs]).astype(np.int64)
        out[f"g{gi}_sh_contig"] = np.array([cidx[shared[k][gen.name][0]] for k in keys], dtype=np.int32)
        out[f"g{gi}_sh_start"] = np.array([shared[k][gen.name][1] for k in keys], dtype=np.int64)
        out[f"g{gi}_sh_strand"] = np.array([shared[k][gen.name][2] == "-" for k in keys], dtype=np.uint8)
        rec = []
        for contig, lst in cands[gen.name].items():
            for p in lst:
                rec.append((key_idx[str(p)], cidx[contig], p.start, p.end, p.strand == "-", p.Tm, p.gcPer))
        rec.sort()
        out[f"g{gi}_cand_kmer"] = np.array([r[0] for r in rec], dtype=np.int64)
        out[f"g{gi}_cand_contig"] = np.array([r[1] for r in rec], dtype=np.int32)
        out[f"g{gi}_cand_start"] = np.array([r[2] for r in rec], dtype=np.int64)
        out[f"g{gi}_cand_end"] = np.array([r[3] for r in rec], dtype=np.int64)
        out[f"g{gi}_cand_strand"] = np.array([r[4] for r in rec], dtype=np.uint8)
        out[f"g{gi}_cand_tm"] = np.array([r[5] for r in rec], dtype=np.float64)
        out[f"g{gi}_cand_gc"] = np.array([r[6] for r in rec], dtype=np.float64)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(keys)} shared, {len(out['g0_cand_kmer'])} candidates -> {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make_one(nm)
