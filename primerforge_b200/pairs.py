"""Drop-in for primerForge's pair search.

    from primerforge_b200.pairs import _getPrimerPairs
    import bin.main; bin.main._getPrimerPairs = _getPrimerPairs

Same entry point, arguments, return value and errors as /root/reference/bin/getPrimerPairs.py:505-564.  Binning,
the substring reduction of the first genome's bins, the bin pairs, the first suitable primer pair of every bin pair,
the look-up of every pair in every other genome and its bin numbers there run in libpfb200.so on the GPU (pfb_pp_*,
include/pfb200.h); primer3's heterodimer check (:146-178) runs on the host on the proposed pairs; the host builds the
``dict[(Primer, Primer)] -> dict[genome] -> Product`` the interface asks for.  There is no CPU path.

When the candidates are the lazy containers ``_getAllCandidateKmers`` of this package returned, their columns go to
the GPU as they are and no ``Primer`` object is created except the members of the pairs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from .engine import PfbError
from .outgroup import _product_class, strings_to_words

ERR_NO_PAIRS = "could not identify suitable primer pairs from the candidate kmers"         # getPrimerPairs.py:519
ERR_NO_SHARED_PAIRS = "could not identify primer pairs present in every ingroup genome"    # getPrimerPairs.py:520


@dataclass
class PairResult:
    pairs: np.ndarray      # PP_PAIR_DTYPE [n]: what the generator of __getCandidatePrimerPairs yields, in its order
    valid: np.ndarray      # bool [n, G]: the pair has a Product in that genome (genome 0 = the first genome)
    length: np.ndarray     # uint32 [n, G]: its size there
    kept: np.ndarray       # bool [n]: a Product in every genome (getPrimerPairs.py:459-462)
    bin_of: list           # per genome: uint32 [S] bin number of every candidate within its contig
    timing: dict

    def shared4(self) -> np.ndarray:
        """[n, G, 4] uint32: valid, length, fbin, rbin -- the layout of oracle/pf_oracle_pairs.c's result"""
        n, G = self.valid.shape
        out = np.zeros((n, G, 4), dtype=np.uint32)
        out[:, :, 0] = self.valid
        out[:, :, 1] = self.length * self.valid
        for g in range(G):
            out[:, g, 2] = self.bin_of[g][self.pairs["fwd"]] * self.valid[:, g]
            out[:, g, 3] = self.bin_of[g][self.pairs["rev"]] * self.valid[:, g]
        return out


class PairSearch:
    """one pfb_pp handle: candidates, genomes, run, results"""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.pfb_pp_create(C.byref(self._h), int(device))
        if rc != _lib.PFB_OK:
            raise PfbError(rc, (self._lib.pfb_pp_last_error(None) or b"").decode())
        self.n_candidates = 0
        self.n_genomes = 0

    def close(self):
        if self._h:
            self._lib.pfb_pp_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != _lib.PFB_OK:
            raise PfbError(rc, (self._lib.pfb_pp_last_error(self._h) or b"").decode())

    def set_candidates(self, word, k, tm):
        word = np.ascontiguousarray(word, dtype=np.uint64)
        k = np.ascontiguousarray(k, dtype=np.uint8)
        tm = np.ascontiguousarray(tm, dtype=np.float64)
        if not (word.size == k.size == tm.size):
            raise ValueError("candidates: arrays of different lengths")
        self._ck(self._lib.pfb_pp_set_candidates(self._h, word.ctypes.data, k.ctypes.data, tm.ctypes.data, word.size))
        self.n_candidates, self.n_genomes = int(word.size), 0

    def add_genome(self, ids, contig, left, strand):
        """the genome's candidates in the order of its dict of lists"""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        contig = np.ascontiguousarray(contig, dtype=np.uint32)
        left = np.ascontiguousarray(left, dtype=np.uint32)
        strand = np.ascontiguousarray(strand, dtype=np.uint8)
        if not (ids.size == contig.size == left.size == strand.size):
            raise ValueError("genome: arrays of different lengths")
        self._ck(self._lib.pfb_pp_add_genome(self._h, ids.ctypes.data, contig.ctypes.data, left.ctypes.data, strand.ctypes.data, ids.size))
        self.n_genomes += 1

    def run(self, max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0):
        prm = _lib.PfbPpParams(max_bin=int(max_bin), min_primer_len=int(min_primer_len), min_pcr=int(min_pcr), max_pcr=int(max_pcr),
                               max_tm_diff=float(max_tm_diff))
        self._ck(self._lib.pfb_pp_run(self._h, C.byref(prm)))

    def timing(self) -> dict:
        t = _lib.PfbPpTiming()
        self._ck(self._lib.pfb_pp_timings(self._h, C.byref(t)))
        return {name: getattr(t, name) for name, _ in t._fields_}

    def bins(self, genome: int, with_reduced: bool = False):
        b = np.empty(self.n_candidates, dtype=np.uint32)
        r = np.zeros(self.n_candidates, dtype=np.uint8) if with_reduced else None
        self._ck(self._lib.pfb_pp_bins(self._h, int(genome), b.ctypes.data, r.ctypes.data if r is not None else None, b.size))
        return (b, r) if with_reduced else b

    def result(self) -> PairResult:
        n, kept_n, nb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.pfb_pp_counts(self._h, C.byref(n), C.byref(kept_n), C.byref(nb)))
        G = self.n_genomes
        pairs = np.empty(n.value, dtype=_lib.PP_PAIR_DTYPE)
        shared = np.zeros((G, n.value), dtype=np.uint32)          # genome-major, as the library keeps it
        kept = np.zeros(n.value, dtype=np.uint8)
        self._ck(self._lib.pfb_pp_pairs(self._h, pairs.ctypes.data, pairs.size))
        if G and n.value:
            self._ck(self._lib.pfb_pp_shared(self._h, shared.ctypes.data, shared.size))
            self._ck(self._lib.pfb_pp_kept(self._h, kept.ctypes.data, kept.size))
        bins = [self.bins(g) for g in range(G)] if self.n_candidates else [np.zeros(0, np.uint32) for _ in range(G)]
        shared = shared.T                                          # views [n, G]
        return PairResult(pairs=pairs, valid=(shared & np.uint32(_lib.PP_VALID)) != 0, length=shared & np.uint32(0x7FFFFFFF),
                          kept=kept.astype(bool), bin_of=bins, timing=self.timing())


# ------------------------------------------------------------------------------------------------------------------
def _columns(candidate_kmers, first_name):
    """the candidate dict as columns: (word, k, tm, per-genome (ids, contig, left, strand), first genome's lists)"""
    from .lazy import LazyGenome, LazyPrimerList
    names = [first_name] + [n for n in candidate_kmers if n != first_name]
    first_lists = list(candidate_kmers[first_name].values())
    lazy = all(isinstance(candidate_kmers[n], LazyGenome) for n in names) and \
        all(isinstance(v, LazyPrimerList) and v._view is not None for n in names for v in candidate_kmers[n].values())
    if lazy:
        cols = first_lists[0]._view[0] if first_lists else None
        lazy = cols is not None and all(v._view[0] is cols for n in names for v in candidate_kmers[n].values())
    per_genome = []
    if lazy:
        word, k, tm = cols.word, cols.k.astype(np.uint8), cols.tm
        for n in names:
            ids, ct, lf, st = [], [], [], []
            for ci, v in enumerate(candidate_kmers[n].values()):
                _c, idx, left, minus = v._view
                ids.append(np.asarray(idx)); ct.append(np.full(len(idx), ci, dtype=np.uint32)); lf.append(np.asarray(left)); st.append(np.asarray(minus))
            cat = (lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dtype=dt))
            per_genome.append((cat(ids, np.uint32), cat(ct, np.uint32), cat(lf, np.uint32), cat(st, np.uint8)))
        return word, k, tm, names, per_genome, first_lists
    seqs, tm = [], []
    for plist in first_lists:
        for p in plist:
            seqs.append(str(p.seq)); tm.append(p.Tm)
    index = {s: i for i, s in enumerate(seqs)}
    word, k = strings_to_words(seqs)
    for n in names:
        ids, ct, lf, st = [], [], [], []
        for ci, plist in enumerate(candidate_kmers[n].values()):
            for p in plist:
                ids.append(index[str(p.seq)]); ct.append(ci); lf.append(min(p.start, p.end)); st.append(p.strand == "-")
        per_genome.append((np.array(ids, dtype=np.uint32), np.array(ct, dtype=np.uint32), np.array(lf, dtype=np.uint32),
                           np.array(st, dtype=np.uint8)))
    return word, k, np.array(tm, dtype=np.float64), names, per_genome, first_lists


def _dimer_provider(params, dimer):
    """f(fwd_str, rev_str) -> heterodimer Tm, or None (no check: every pair passes with dimerTm 0.0)"""
    if dimer == "off" or (dimer is None and os.environ.get("PFB200_DIMER", "").lower() == "off"):
        return None
    if callable(dimer):
        return dimer
    try:
        import primer3  # type: ignore
    except ImportError as exc:
        raise ImportError("primer3-py is required for the heterodimer check (getPrimerPairs.py:146-178); "
                          "pass dimer='off' to skip it") from exc
    kw = dict(mv_conc=params.p3_mvConc, dv_conc=params.p3_dvConc, dntp_conc=params.p3_dntpConc, dna_conc=params.p3_dnaConc,
              temp_c=params.p3_tempC, max_loop=params.p3_maxLoop)
    return lambda fwd, rev: primer3.calc_heterodimer_tm(fwd, rev, **kw)


def _getPrimerPairs(candidateKmers, params, dimer=None, device: int | None = None):
    """drop-in for bin/getPrimerPairs.py:505 (same arguments and return value)"""
    device = int(os.environ.get("PFB200_DEVICE", "0")) if device is None else device
    first_name = os.path.basename(params.ingroupFns[0])                                   # :523
    word, k, tm, names, per_genome, first_lists = _columns(candidateKmers, first_name)
    with PairSearch(device=device) as ps:
        ps.set_candidates(word, k, tm)
        for ids, ct, lf, st in per_genome:
            ps.add_genome(ids, ct, lf, st)
        ps.run(max_bin=params.maxBinSize, min_primer_len=params.minLen, min_pcr=params.minPcr, max_pcr=params.maxPcr,
               max_tm_diff=params.maxTmDiff)
        res = ps.result()
    contig_names = [list(candidateKmers[n].keys()) for n in names]
    return build_pairs_output(res, k, names, per_genome, first_lists, contig_names, params, _dimer_provider(params, dimer))


def build_pairs_output(res: PairResult, k, names, per_genome, first_lists, contig_names, params, dimer_fn) -> dict:
    """the host half of the drop-in: primer3's heterodimer check on the proposed pairs and the reference's return value
    ``dict[(Primer, Primer)] -> dict[genome] -> Product`` from a PairResult and the candidate columns (no GPU involved: the
    CPU tests run it on the oracle's result)"""
    Product = _product_class()
    # the first genome's Primer objects of the proposed pairs (list position -> object)
    ids0 = per_genome[0][0]
    pos_of = np.empty(ids0.size, dtype=np.int64)
    pos_of[ids0] = np.arange(ids0.size)
    offs = np.cumsum([0] + [len(v) for v in first_lists])

    def primer_at(cand):
        p = int(pos_of[cand])
        ci = int(np.searchsorted(offs, p, side="right") - 1)
        return first_lists[ci][p - int(offs[ci])]

    # heterodimer check on the proposed pairs (:146-178, 330-341): each is judged on its own
    accepted = []
    for q in range(res.pairs.size):
        f, r = int(res.pairs["fwd"][q]), int(res.pairs["rev"][q])
        fwd = primer_at(f)
        rev = primer_at(r).reverseComplement()                                            # :333
        if dimer_fn is None:
            accepted.append((q, fwd, rev, 0.0))
            continue
        dimer_tm = dimer_fn(str(fwd), str(rev))
        if not dimer_tm >= (min(fwd.Tm, rev.Tm) - params.tempTolerance):                  # :176
            accepted.append((q, fwd, rev, dimer_tm))
    if not accepted:                                                                      # :541-544
        params.log.rename(_getPrimerPairs.__name__)
        params.log.error(ERR_NO_PAIRS)
        raise RuntimeError(ERR_NO_PAIRS)

    # the Products (:336, :463, :501-502): per genome the place of both members comes from the columns
    kept = res.kept
    out = {}
    place = []
    for ids, ct, lf, st in per_genome:
        c_of = np.empty(ids.size, dtype=np.int64); l_of = np.empty(ids.size, dtype=np.int64); s_of = np.empty(ids.size, dtype=np.int64)
        c_of[ids] = ct; l_of[ids] = lf; s_of[ids] = st
        place.append((c_of, l_of, s_of))
    kk = np.asarray(k, dtype=np.int64)
    for q, fwd, rev, dimer_tm in accepted:
        if not kept[q]:
            continue
        a, b = int(res.pairs["fwd"][q]), int(res.pairs["rev"][q])
        ka, kb = int(kk[a]), int(kk[b])
        per = {}
        for g, name in enumerate(names):
            c_of, l_of, s_of = place[g]
            la, lb, sa, sb = int(l_of[a]), int(l_of[b]), int(s_of[a]), int(s_of[b])
            f_start, f_end = (la, la + ka - 1) if sa == 0 else (la + ka - 1, la)
            r_minus = 1 - sb                                                              # k2.reverseComplement() (:458)
            r_start, r_end = (lb, lb + kb - 1) if r_minus == 0 else (lb + kb - 1, lb)
            per[name] = Product(contig_names[g][int(c_of[a])], int(res.length[q, g]), int(res.bin_of[g][a]), int(res.bin_of[g][b]), dimer_tm,
                                f_start, f_end, "-" if sa else "+", r_start, r_end, "-" if r_minus else "+")
        out[(fwd, rev)] = per
    if not out:                                                                           # :550-553
        params.log.rename(_getPrimerPairs.__name__)
        params.log.error(ERR_NO_SHARED_PAIRS)
        raise RuntimeError(ERR_NO_SHARED_PAIRS)
    return out
