"""Columnar result of the candidate-k-mer stage behind the reference's container types.

``_getAllCandidateKmers`` must return ``dict[genome] -> dict[contig] -> list[Primer]``
(/root/reference/bin/getCandidateKmers.py:517-530): G x |S| Python objects, which -- once the arithmetic runs on
the GPU -- cost a thousand times more than computing them (0.8 us each against milliseconds for the whole
stage).  The stage therefore returns the SAME shapes, but backed by the columns the GPU produced:

* ``LazyGenome`` is a ``dict`` (contig id -> list) that sorts its genome's candidates by (contig, leftmost
  start) and creates its per-contig lists the first time anything is asked of it;
* ``LazyPrimerList`` is a ``list`` that creates its ``Primer`` objects (``bin.Primer.Primer`` when primerForge is
  importable, else the mirror in primer.py) the first time it is iterated, indexed or compared; ``len()`` is
  answered from the columns.

After that first use both ARE ordinary populated dicts / lists, so every consumer -- ``getPrimerPairs.py:53-60,
344-362, 533-536`` iterate them -- works unchanged.  Pickling goes through ``__reduce__`` as a plain ``dict`` /
``list`` of Primers: ``candidates.p`` has the reference's layout and loads without this package.
"""
from __future__ import annotations

import gc

import numpy as np

from .engine import words_to_strings
from .primer import Seq, primer_class


class Columns:
    """what every view of one run shares: the picked k-mers (columns, in dict order) and their place in every genome"""

    def __init__(self, names, contig_ids, contig_word_starts, word, k, tm, gc_count, pos, thal=None):
        self.names = list(names)
        self.contig_ids = contig_ids                   # per genome: list of contig ids (file order)
        self.contig_word_starts = contig_word_starts   # per genome: None (one contig) or uint32 [n_contigs] word offsets
        self.word = np.ascontiguousarray(word, dtype=np.uint64)
        self.k = np.ascontiguousarray(k, dtype=np.int64)
        self.tm = np.ascontiguousarray(tm, dtype=np.float64)
        self.gc_per = np.asarray(gc_count, dtype=np.float64) / self.k * 100          # Primer.py:100
        self.pos = pos                                 # per genome: uint32 [n] padded coordinate | strand << 31
        self.thal = thal                               # None, or dict record index -> (hairpin, rcHairpin, homodimer, rcHomodimer)
        self._seqs = None
        self.cls = primer_class()

    @property
    def n(self) -> int:
        return int(self.word.size)

    def seqs(self) -> list:
        """Bio.Seq objects of the picked k-mers, made once and shared by every genome's Primers"""
        if self._seqs is None:
            self._seqs = [Seq(s) for s in words_to_strings(self.word, self.k)]
        return self._seqs

    def place(self, g: int):
        """(contig index, leftmost start, minus) of every picked k-mer in genome g"""
        v = self.pos[g]
        lp = (v & np.uint32(0x7FFFFFFF)).astype(np.int64)
        minus = (v >> np.uint32(31)).astype(bool)
        ws = self.contig_word_starts[g]
        if ws is None:
            return np.zeros(lp.size, dtype=np.int64), lp, minus
        # contigs are laid out in the order added, each starting at a multiple of 32 bases
        contig = np.searchsorted(ws, lp >> 5, side="right") - 1
        return contig, lp - 32 * ws[contig].astype(np.int64), minus


def _no_gc(fn):
    """millions of small containers: the cyclic collector's passes over them cost more than creating them"""
    def wrapped(*a, **kw):
        was = gc.isenabled()
        gc.disable()
        try:
            return fn(*a, **kw)
        finally:
            if was:
                gc.enable()
    return wrapped


class LazyPrimerList(list):
    """the candidates of one contig of one genome, sorted by leftmost start (getCandidateKmers.py:598-602)"""

    def __init__(self, cols: Columns, contig_name: str, idx, left, right, minus):
        super().__init__()
        self._lazy = (cols, contig_name, idx, left, right, minus)
        self._view = (cols, idx, left, minus)      # the columns behind the list while nobody has edited it (pairs.py)
        self._n = int(len(idx))

    @_no_gc
    def _fill(self):
        lz = self._lazy
        if lz is None:
            return
        self._lazy = None
        cols, cname, idx, left, right, minus = lz
        seqs = cols.seqs()
        new, cls = cols.cls.__new__, cols.cls
        idx_l = idx.tolist()
        starts = np.where(minus, right, left).tolist()          # '-' records keep start > end (Primer.py:42-45)
        ends = np.where(minus, left, right).tolist()
        strands = np.where(minus, "-", "+").tolist()
        tm = cols.tm[idx].tolist()
        gcp = cols.gc_per[idx].tolist()
        ks = cols.k[idx].tolist()
        thal = cols.thal
        none4 = (None, None, None, None)
        out = []
        append = out.append
        for j, r in enumerate(idx_l):
            p = new(cls)
            th = none4 if thal is None else (thal.get(r) or none4)
            # one instance dict per record and no constructor call: the constructor would recompute GC and Tm,
            # which come from the device
            p.__dict__ = {"seq": seqs[r], "start": starts[j], "end": ends[j], "contig": cname, "strand": strands[j],
                          "Tm": tm[j], "gcPer": gcp[j], "hairpinTm": th[0], "rcHairpin": th[1],
                          "homodimerTm": th[2], "rcHomodimer": th[3], "_Primer__length": ks[j]}
            append(p)
        list.extend(self, out)

    # answered from the columns
    def __len__(self):
        return self._n if self._lazy is not None else list.__len__(self)

    def __bool__(self):
        return len(self) > 0

    # everything else needs the objects
    def __iter__(self):
        self._fill()
        return list.__iter__(self)

    def __getitem__(self, i):
        self._fill()
        return list.__getitem__(self, i)

    def __reversed__(self):
        self._fill()
        return list.__reversed__(self)

    def __contains__(self, x):
        self._fill()
        return list.__contains__(self, x)

    def __eq__(self, other):
        self._fill()
        if isinstance(other, LazyPrimerList):
            other._fill()
        return list.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._fill()
        return list.__repr__(self)

    def __add__(self, other):
        self._fill()
        return list(self) + list(other)

    def __reduce__(self):          # candidates.p: a plain list of Primers (the reference's layout)
        self._fill()
        return (list, (list(list.__iter__(self)),))

    def __copy__(self):
        self._fill()
        return list(list.__iter__(self))

    def __deepcopy__(self, memo):
        import copy
        self._fill()
        return copy.deepcopy(list(list.__iter__(self)), memo)


def _filling(name):
    def method(self, *a, **kw):
        self._fill()
        if name not in ("index", "count", "copy", "__mul__", "__rmul__", "__lt__", "__le__", "__gt__", "__ge__"):
            self._view = None                      # edited: the columns no longer describe the list
        return getattr(list, name)(self, *a, **kw)
    method.__name__ = name
    return method


for _name in ("append", "extend", "insert", "pop", "remove", "clear", "index", "count", "sort", "reverse", "copy",
              "__setitem__", "__delitem__", "__iadd__", "__mul__", "__imul__", "__rmul__", "__lt__", "__le__", "__gt__", "__ge__"):
    setattr(LazyPrimerList, _name, _filling(_name))


class LazyGenome(dict):
    """contig id -> list[Primer] of one genome; the contigs appear in the order of their first candidate"""

    def __init__(self, cols: Columns, g: int):
        super().__init__()
        self._lazy = (cols, g)

    def _fill(self):
        lz = self._lazy
        if lz is None:
            return
        self._lazy = None
        cols, g = lz
        contig, left, minus = cols.place(g)
        k = cols.k
        # (contig, leftmost start, k, strand): the reference sorts each contig's set by min(start, end) and leaves
        # ties in set-iteration order; here ties are ordered by length, then strand
        key = (contig << 40) | (left << 8) | (k << 1) | minus.astype(np.int64)
        order = np.argsort(key, kind="stable")
        contig_sorted = contig[order]
        n = order.size
        bounds = np.flatnonzero(np.r_[True, contig_sorted[1:] != contig_sorted[:-1], True]) if n else np.zeros(1, dtype=np.int64)
        ids = cols.contig_ids[g]
        left_o, minus_o = left[order], minus[order]
        right_o = left_o + k[order] - 1
        for a, b in zip(bounds[:-1].tolist(), bounds[1:].tolist()):
            cname = ids[int(contig_sorted[a])]
            dict.__setitem__(self, cname, LazyPrimerList(cols, cname, order[a:b], left_o[a:b], right_o[a:b], minus_o[a:b]))

    def n_candidates(self) -> int:
        """number of candidates in this genome (every picked k-mer occurs exactly once in every genome)"""
        if self._lazy is not None:
            return self._lazy[0].n
        return sum(len(v) for v in dict.values(self))

    def __reduce__(self):          # a plain dict of plain lists of Primers
        self._fill()
        return (dict, ({c: list(v) for c, v in dict.items(self)},))

    def __eq__(self, other):
        self._fill()
        if isinstance(other, LazyGenome):
            other._fill()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __bool__(self):
        return self.n_candidates() > 0


def _dict_filling(name):
    def method(self, *a, **kw):
        self._fill()
        return getattr(dict, name)(self, *a, **kw)
    method.__name__ = name
    return method


for _name in ("__getitem__", "__iter__", "__len__", "__contains__", "__repr__", "__setitem__", "__delitem__", "__reversed__",
              "__or__", "__ior__", "keys", "values", "items", "get", "pop", "popitem", "setdefault", "update", "copy", "clear"):
    setattr(LazyGenome, _name, _dict_filling(_name))


def build_lazy_output(cols: Columns) -> dict:
    """__buildOutput (getCandidateKmers.py:442-489) + the union / sort of :583-602 as views over the columns"""
    return {name: LazyGenome(cols, g) for g, name in enumerate(cols.names)}
