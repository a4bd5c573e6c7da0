"""Run the reference's UNMODIFIED ``bin/getCandidateKmers.py`` in this container.

TEST INFRASTRUCTURE ONLY (never imported by the product path).  The reference
lives at /root/reference (read-only, absent on the GPU box); its four
third-party dependencies are absent from this image, so they are supplied by
the stand-ins under ``oracle/standins`` (semantics documented there).  Used by
``oracle/make_golden.py`` to produce the committed fixtures in tests/golden/.
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("PF_REFERENCE_ROOT", "/root/reference")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


class _NullLog:
    def __init__(self):
        self.debug_lines = []

    def rename(self, *_a, **_k):
        pass

    def info(self, *_a, **_k):
        pass

    def debug(self, msg, *_a, **_k):
        self.debug_lines.append(str(msg))

    def error(self, *_a, **_k):
        pass

    def critical(self, *_a, **_k):
        pass


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bin", "getCandidateKmers.py"))


def _import_reference(thal_mode: str):
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ["PF_STANDIN_THAL"] = thal_mode
    os.environ.setdefault("CI", "1")  # disables the Clock spinner subprocess (bin/Clock.py:220-225)
    for p in (REFERENCE_ROOT, _STANDINS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import primer3  # the stand-in
    primer3.THAL_MODE = thal_mode
    mod = importlib.import_module("bin.getCandidateKmers")
    return mod


def make_params(ingroup_fns, fmt="fasta", min_len=16, max_len=20, min_gc=40.0, max_gc=60.0,
                min_tm=55.0, max_tm=68.0, max_repeats=4, temp_tolerance=5.0, num_threads=1, sink=None, **_unused):
    """a stand-in for ``bin.Parameters.Parameters`` carrying the fields the stage reads
    (SURVEY.md section 8b); ``sink`` receives what the stage pickles"""
    sink = sink if sink is not None else {}

    def dump_obj(obj, fn, what, prefix=""):
        sink[fn] = obj

    def load_obj(fn):
        return sink[fn]

    return SimpleNamespace(
        ingroupFns=list(ingroup_fns), format=fmt, minLen=min_len, maxLen=max_len, minGc=min_gc,
        maxGc=max_gc, minTm=min_tm, maxTm=max_tm, maxRepeatLen=max_repeats,
        tempTolerance=temp_tolerance, numThreads=num_threads, p3_mvConc=50.0, p3_dvConc=1.5,
        p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=_NullLog(),
        pickles={1: "sharedKmers.p", 2: "candidates.p"}, dumpObj=dump_obj, loadObj=load_obj,
        debug=False, keepIntermediateFiles=False)


def run_reference(ingroup_fns, thal_mode="pass", **kw):
    """returns (shared, candidates): the reference's ``sharedKmers`` dict
    (kmer -> {genome -> (contig, start, strand)}, insertion-ordered) and the
    stage's return value (genome -> contig -> list[Primer])"""
    mod = _import_reference(thal_mode)
    sink = {}
    params = make_params(ingroup_fns, sink=sink, **kw)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        cands = mod._getAllCandidateKmers(params, False)
    return sink["sharedKmers.p"], cands


def run_reference_pairs(candidates, ingroup_fns, min_len=16, max_len=20, min_pcr=120, max_pcr=2400, max_tm_diff=5.0,
                        max_bin_size=64):
    """the reference's UNMODIFIED ``bin/getPrimerPairs.py::_getPrimerPairs`` on a candidate dict"""
    _import_reference("pass")
    gpp = importlib.import_module("bin.getPrimerPairs")
    prm = SimpleNamespace(ingroupFns=list(ingroup_fns), maxBinSize=max_bin_size, minLen=min_len, maxLen=max_len,
                          minPcr=min_pcr, maxPcr=max_pcr, maxTmDiff=max_tm_diff, numThreads=1, tempTolerance=5.0,
                          p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30,
                          log=_NullLog())
    return gpp._getPrimerPairs(candidates, prm)


def run_reference_outgroup(outgroup_fns, pairs, disallowed, fmt="fasta"):
    """the reference's UNMODIFIED ``bin/removeOutgroupPrimers.py::_removeOutgroupPrimers`` (modifies ``pairs`` in
    place like the reference does); the outgroup generators are made the way ``bin/main.py:41-60`` makes them.
    Returns the debug log lines; raises what the reference raises."""
    _import_reference("pass")
    rop = importlib.import_module("bin.removeOutgroupPrimers")
    from Bio import SeqIO  # the stand-in
    log = _NullLog()
    prm = SimpleNamespace(disallowedLens=disallowed, log=log)
    outgroup = {os.path.basename(fn): SeqIO.parse(fn, fmt) for fn in outgroup_fns}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        rop._removeOutgroupPrimers(outgroup, pairs, prm)
    return log.debug_lines


def reference_extract_kmer_data(seq: str, strand: str, kmers):
    """``bin/removeOutgroupPrimers.py::__extractKmerData`` (:34-62) on one string; kmers = list of primer strings"""
    _import_reference("pass")
    rop = importlib.import_module("bin.removeOutgroupPrimers")
    from ahocorasick import Automaton  # the stand-in
    aut = Automaton()
    for k in kmers:
        aut.add_word(k, k)
    aut.make_automaton()
    return getattr(rop, "__extractKmerData")(seq, strand, aut)
