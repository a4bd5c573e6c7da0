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
                min_tm=55.0, max_tm=68.0, max_repeats=4, temp_tolerance=5.0, num_threads=1, sink=None):
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
