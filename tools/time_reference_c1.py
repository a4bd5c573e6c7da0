"""One-off, clearly labelled timing of the UNMODIFIED reference Python on BASELINE config 1.

    python tools/time_reference_c1.py [--length 1000000] [--threads N] [--out profiles/...json]

Runs /root/reference/bin/getCandidateKmers.py::_getAllCandidateKmers (unmodified) on the three
synthetic 1 Mbp FASTA genomes of config 1, against the stand-ins for khmer / pyahocorasick /
primer3 / Bio under oracle/standins (none of the four is installable in this image), hairpin /
homodimer screening stubbed to "pass".  The number is therefore "reference Python + stand-in
deps" (BASELINE.md section 2), NOT "primerForge as installed": the stand-ins are Python / numpy,
where the real dependencies are C / C++.  TEST INFRASTRUCTURE: only works where /root/reference
exists (the build container).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_digests(shared, cands, gens, paths):
    """the reference's sharedKmers dict and returned candidates in oracle/digest.py's canonical layout"""
    import numpy as np
    from oracle import digest
    names = [os.path.basename(p) for p in paths]
    kmers = [str(x) for x in shared.keys()]
    word, k = digest.strings_to_words(kmers)
    contig, start, strand = [], [], []
    for name, gen in zip(names, gens):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        rows = [shared[x][name] for x in shared]
        contig.append(np.array([cidx[r[0]] for r in rows], dtype=np.int32))
        start.append(np.array([r[1] for r in rows], dtype=np.int64))
        strand.append(np.array([r[2] == "-" for r in rows], dtype=np.uint8))
    out = {"n_shared": len(kmers), "shared": digest.shared_digest(word, k, contig, start, strand), "candidates": []}
    for name, gen in zip(names, gens):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        plist = [(cidx[c], p) for c, lst in cands[name].items() for p in lst]
        w, kk = digest.strings_to_words([str(p.seq) for _c, p in plist])
        out["candidates"].append(digest.candidate_digest(
            w, kk, [c for c, _p in plist], [p.start for _c, p in plist], [p.end for _c, p in plist],
            [p.strand == "-" for _c, p in plist], [p.Tm for _c, p in plist]))
    out["n_candidates"] = sum(len(v) for v in cands[names[0]].values())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_reference_python_c1.json"))
    args = ap.parse_args()
    from primerforge_b200 import synth
    from oracle import run_reference as rr
    gens, cfg = synth.make_config(1, length=args.length)
    with tempfile.TemporaryDirectory() as tmp:
        paths = synth.write_fasta(gens, tmp)
        mod = rr._import_reference("pass")
        sink = {}
        params = rr.make_params(paths, sink=sink, min_len=cfg["k_min"], max_len=cfg["k_max"],
                                max_repeats=cfg["max_repeats"], num_threads=args.threads)
        # the stage's own two phases (README.md:342-346): shared k-mers, then evaluation
        t = {}
        real_dump = params.dumpObj

        def dump(obj, fn, what, prefix=""):
            t[f"dump_{fn}"] = time.perf_counter()
            real_dump(obj, fn, what, prefix)
        params.dumpObj = dump
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            cands = mod._getAllCandidateKmers(params, False)
        t1 = time.perf_counter()
        digests = reference_digests(sink["sharedKmers.p"], cands, gens, paths)
    n_bases = sum(g.n_bases for g in gens)
    phase_a = t["dump_sharedKmers.p"] - t0
    out = {"what": "UNMODIFIED /root/reference/bin/getCandidateKmers.py::_getAllCandidateKmers on BASELINE config 1 "
                   "(3 synthetic genomes, primer_len 16,20, defaults), stand-in khmer / pyahocorasick / primer3 / Bio "
                   "(Python + numpy; oracle/standins), hairpin / homodimer stubbed to pass",
           "label": "reference Python + stand-in deps (NOT primerForge as installed)",
           "genomes": len(gens), "genome_bp": args.length, "num_threads": args.threads, "host_cpus": os.cpu_count(),
           "seconds_total": t1 - t0, "seconds_phase_A_shared_kmers": phase_a, "seconds_phase_B_evaluate": t1 - t0 - phase_a,
           "Mbp_per_s": n_bases / 1e6 / (t1 - t0), "n_shared": len(sink["sharedKmers.p"]),
           "n_candidates_first_genome": sum(len(v) for v in cands[os.path.basename(paths[0])].values()),
           "where": "build container (CPU only)"}
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    if args.length == 1_000_000:
        # the reference's own result at full config-1 size, as digests (oracle/digest.py): tests/ compare the CPU
        # oracle (-m "not gpu") and the CUDA path (-m gpu) with these
        with open(os.path.join(ROOT, "tests", "golden", "c1_reference_digest.json"), "w") as fh:
            json.dump({"made_by": "tools/time_reference_c1.py (unmodified reference Python + oracle/standins)",
                       "config": 1, "genomes": len(gens), "genome_bp": args.length, **digests}, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
