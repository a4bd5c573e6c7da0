"""Drop-in for primerForge's candidate-k-mer stage.

    from primerforge_b200.candidates import _getAllCandidateKmers
    import bin.main; bin.main._getAllCandidateKmers = _getAllCandidateKmers

Same entry point, arguments, return value, error strings and side effects as
/root/reference/bin/getCandidateKmers.py:517-624.  The arithmetic (sketch counting, strand algebra,
cross-genome intersection, relocation, GC / Tm / repeat filters, per-position grouping) runs in
libpfb200.so on the GPU(s); hairpin / homodimer screening stays primer3 on the host, evaluated lazily
group by group exactly like :376-388.  The returned ``dict[genome][contig] -> list[Primer]`` is backed by
the GPU's result columns and creates its ``Primer`` objects when they are first looked at (lazy.py).

Environment: ``PFB200_GPUS`` = number of GPUs to use (default 1; "all" = every visible one) -- one process
drives them all (pfb_multi_*); ``PFB200_DEVICE`` = the (first) device; ``PFB200_THAL=off`` skips hairpin /
homodimer screening; ``PFB200_HOST_PARSE=1`` reads the files on the host instead of on the device.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _lib, seqio
from .engine import ERR_NO_CANDIDATES, ERR_NO_SHARED, Engine, MultiEngine, PfbError, StageResult, words_to_strings
from .lazy import Columns, build_lazy_output

_GAP = " " * 4
_SHARED, _CAND = 1, 2        # bin/Parameters.py: pickles[_SHARED], pickles[_CAND]


def _thal_provider(params, thal):
    """returns f(seq_str, rc_str) -> (hairpinTm, rcHairpin, homodimerTm, rcHomodimer) or None (no screening)"""
    if thal == "off":
        return None
    if callable(thal):
        return thal
    try:
        import primer3  # type: ignore
    except ImportError as exc:
        raise ImportError("primer3-py is required for hairpin / homodimer screening "
                          "(getCandidateKmers.py:322-370); pass thal='off' to skip it") from exc
    kw = dict(mv_conc=params.p3_mvConc, dv_conc=params.p3_dvConc, dntp_conc=params.p3_dntpConc,
              dna_conc=params.p3_dnaConc, temp_c=params.p3_tempC, max_loop=params.p3_maxLoop)

    def run(seq, rc):
        return (primer3.calc_hairpin_tm(seq, **kw), primer3.calc_hairpin_tm(rc, **kw),
                primer3.calc_homodimer_tm(seq, **kw), primer3.calc_homodimer_tm(rc, **kw))
    return run


_RC = str.maketrans("ACGT", "TGCA")


def screen_groups(records, positions, kmers, thal_fn, limit: float):
    """host part of __evaluateKmersAtOnePosition (:322-388) for records that already passed GC, Tm
    and the repeat filter on the GPU: per genome, per (contig, start) group, in dict order, the first
    k-mer whose four secondary-structure Tms are all < limit is picked.  Returns (picked mask, thal
    values per picked record)."""
    n = records.size
    cache = {}
    picked = np.zeros(n, dtype=bool)

    def passes(r):
        v = cache.get(r)
        if v is None:
            s = kmers[r]
            v = thal_fn(s, s[::-1].translate(_RC))
            cache[r] = v
        # hairpins first, homodimers only if the hairpins pass (:386-387); values are pure functions
        return v[0] < limit and v[1] < limit and v[2] < limit and v[3] < limit

    ok_filter = (records["flags"] & 7) == 7
    idx_all = np.flatnonzero(ok_filter)
    for pos in positions:
        key = (pos["contig"][idx_all].astype(np.int64) << 32) | pos["start"][idx_all].astype(np.int64)
        order = np.argsort(key, kind="stable")          # members of a group stay in dict (rank) order
        ks = key[order]
        starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
        ends = np.r_[starts[1:], ks.size]
        members = idx_all[order]
        for a, b in zip(starts, ends):
            for r in members[a:b]:
                r = int(r)
                if picked[r] or passes(r):
                    picked[r] = True
                    break
    return picked, cache


def _word_starts(lengths) -> np.ndarray | None:
    """word offset of every contig of a genome in the engine's layout (each contig starts at a multiple of 32)"""
    if len(lengths) <= 1:
        return None
    words = (np.asarray(lengths, dtype=np.int64) + 31) // 32
    return np.concatenate([[0], np.cumsum(words)[:-1]]).astype(np.int64)


class _Stage:
    """one run of the stage on 1..N GPUs: engines, which engine holds which genome, and the per-id columns"""

    def __init__(self, n_genomes: int, devices, **params):
        self.multi = MultiEngine(devices, **params) if len(devices) > 1 else None
        self.engines = self.multi.engines if self.multi else [Engine(device=devices[0], **params)]
        self.n_genomes = n_genomes

    def where(self, g: int):
        return self.multi.where(g) if self.multi else (0, g)

    def add(self, *, contigs=None, image=None, fmt="fasta"):
        target = self.multi or self.engines[0]
        if image is not None:
            target.add_fasta(image, fmt=fmt)
        else:
            target.add_genome(contigs)

    def run(self):
        (self.multi or self.engines[0]).run()

    def close(self):
        if self.multi:
            self.multi.close()
        else:
            self.engines[0].close()

    def contig_table(self, g: int):
        e, li = self.where(g)
        return self.engines[e].contig_table(li)

    def columns(self, sel: np.ndarray, names, contig_ids, thal=None) -> Columns:
        """the picked (or otherwise selected) ids as columns; the records come from engine 0, a genome's places
        from the engine that holds it"""
        views = [e.result_by_id() for e in self.engines]
        v0 = views[0]
        pos, wstarts = [], []
        for g in range(self.n_genomes):
            e, li = self.where(g)
            pos.append(views[e].pos[li][sel].copy())
            wstarts.append(_word_starts(self.engines[e].contig_table(li)[1]))
        return Columns(names, contig_ids, wstarts, v0.word[sel], v0.k[sel], v0.tm[sel], v0.gc_count[sel], pos, thal=thal)


def _devices() -> list:
    first = int(os.environ.get("PFB200_DEVICE", "0"))
    want = os.environ.get("PFB200_GPUS", "1").strip().lower()
    if want in ("", "1"):
        return [first]
    import torch
    have = torch.cuda.device_count()
    n = have if want == "all" else int(want)
    if n < 1 or first + n > have:
        raise RuntimeError(f"PFB200_GPUS={want} but {have} CUDA devices are visible (first device {first})")
    return list(range(first, first + n))


def run_stage(genomes, *, k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4,
              table_size=9999991, prefilter=True, device=0, scan_mode=0, picked_only=False, fasta_images=None,
              contig_ids_out=None, file_format="fasta"):
    """genomes: list of (contig id list, list of uint8 ASCII arrays), or -- fasta_images: list of uint8 FASTA /
    GenBank file images (file_format) -- the files themselves, parsed on the device (their record ids are
    appended to contig_ids_out).  Returns a StageResult in the record-oriented layout (tests, hairpin / homodimer
    screening); with picked_only it holds just the picked records."""
    with Engine(device=device, k_min=k_min, k_max=k_max, min_gc=min_gc, max_gc=max_gc, min_tm=min_tm, max_tm=max_tm,
                max_repeat=max_repeat, table_size=table_size, prefilter=int(bool(prefilter))) as eng:
        eng.set_scan_mode(scan_mode)
        if fasta_images is not None:
            for img in fasta_images:
                eng.add_fasta(img, fmt=file_format)
        else:
            for _ids, contigs in genomes:
                eng.add_genome(contigs)
        eng.run()
        if fasta_images is not None and contig_ids_out is not None:
            for g, img in enumerate(fasta_images):
                contig_ids_out.append(seqio.file_ids(img, eng.contig_table(g)[0], file_format))
        n, _ = eng.counts()
        if n == 0:
            raise RuntimeError(ERR_NO_SHARED)
        if not picked_only:
            return eng.result()
        rec, ss, ct = eng.result_picked()
        positions = []
        for g in range(eng.n_genomes):
            pos = np.zeros(rec.size, dtype=_lib.POS_DTYPE)
            pos["start"] = ss[g] & 0x7FFFFFFF
            pos["strand"] = ss[g] >> 31
            if ct is not None:
                pos["contig"] = ct[g]
            positions.append(pos)
        res = StageResult(records=rec.copy(), positions=positions, timing=eng.timing())
        res.timing["n_shared_reported"] = n
        return res


def _shared_dict(stage: _Stage, names, contig_ids):
    """the reference's sharedKmers object (getCandidateKmers.py:190-219): kmer -> {genome -> (contig, start, strand)}
    in dict insertion order -- G x |shared| Python objects: built only for --debug / --keep runs"""
    v0 = stage.engines[0].result_by_id()
    sel = np.flatnonzero(v0.alive)
    cols = stage.columns(sel, names, contig_ids)
    kmers = words_to_strings(cols.word, cols.k)
    places = [cols.place(g) for g in range(len(names))]
    out = {}
    for r, kmer in enumerate(kmers):
        out[kmer] = {}
    for g, name in enumerate(names):
        contig, left, minus = places[g]
        ids = contig_ids[g]
        for r, kmer in enumerate(kmers):
            out[kmer][name] = (ids[int(contig[r])], int(left[r]), "-" if minus[r] else "+")
    return out


def _debug_counts(stage: _Stage, names, log):
    """'<n> shared kmers after processing <genome>' (getCandidateKmers.py:78,133,145): the reference intersects the
    genomes' allowed sets in the order of their names; here every set is already restricted to the first genome's
    k-mers, so the numbers are the reference's from the point where the first genome has been processed"""
    views = [e.result_by_id() for e in stage.engines]
    n = views[0].n_ids
    order = sorted(range(len(names)), key=lambda g: names[g])
    alive = np.ones(n, dtype=bool)
    seen_first = False
    for g in order:
        e, li = stage.where(g)
        alive &= views[e].pos[li] != np.uint32(0xFFFFFFFF)
        seen_first |= g == 0
        note = "" if seen_first else f" (among the kmers of {names[0]})"
        log.debug(f"{int(alive.sum())} shared kmers after processing {names[g]}{note}")


def _getAllCandidateKmers(params, sharedExists: bool, thal="auto", device: int | None = None):
    """gets all the candidate kmer sequences for a given ingroup (GPU implementation).

    Args / Raises / Returns: as /root/reference/bin/getCandidateKmers.py:517-530.  With `sharedExists` the
    reference unpickles ``sharedKmers.p`` instead of recomputing it; here the pickle is opened (so a missing or
    unreadable checkpoint fails like it does there) and the shared k-mers are recomputed -- milliseconds on the
    GPU, against minutes to walk a dict of millions of strings.  ``sharedKmers.p`` itself is written only for
    runs that keep intermediate files or debug (``params.keepIntermediateFiles`` / ``params.debug``): it is a
    multi-GB Python dict that nothing downstream reads.
    """
    log = params.log
    log.rename(_getAllCandidateKmers.__name__)
    debug = bool(getattr(params, "debug", False))
    keep = debug or bool(getattr(params, "keepIntermediateFiles", False)) or os.environ.get("PFB200_DUMP_SHARED", "0") == "1"
    if sharedExists:
        params.loadObj(params.pickles[_SHARED])          # (existence / readability check, like :550-552)
    else:
        msg1 = f"{_GAP}getting shared ingroup kmers that appear once in each genome"
        log.info(msg1)
        print(msg1, end=" ... ", flush=True)
    t0 = time.time()
    devices = _devices() if device is None else [int(device)]
    thal_fn = _thal_provider(params, os.environ.get("PFB200_THAL", thal) if thal == "auto" else thal)

    names = [os.path.basename(fn) for fn in params.ingroupFns]
    contig_ids, images, parsed = [], None, None
    # the files go to the GPU as they are (records found and compacted there); PFB200_HOST_PARSE=1 reads them
    # on the host instead
    fmt = str(params.format).lower()
    fmt = "genbank" if fmt == "gb" else fmt
    if fmt in ("fasta", "genbank") and os.environ.get("PFB200_HOST_PARSE", "0") != "1":
        images = [np.fromfile(fn, dtype=np.uint8) for fn in params.ingroupFns]
    else:
        parsed = [seqio.read_genome(fn, params.format) for fn in params.ingroupFns]
        contig_ids = [ids for ids, _c in parsed]
    # every shared k-mer (not only those that pass GC / Tm / repeats) is needed only for sharedKmers.p / debug counts
    stage = _Stage(len(names), devices, k_min=params.minLen, k_max=params.maxLen, min_gc=params.minGc, max_gc=params.maxGc,
                   min_tm=params.minTm, max_tm=params.maxTm, max_repeat=params.maxRepeatLen, prefilter=0 if keep else 1,
                   table_size=int(os.environ.get("PFB200_TABLE_SIZE", "9999991")))      # (khmer's table; the variable is a test hook)
    try:
        for g in range(len(names)):
            if images is not None:
                stage.add(image=images[g], fmt=fmt)
            else:
                stage.add(contigs=parsed[g][1])
        try:
            stage.run()
        except PfbError as exc:
            if exc.code == _lib.PFB_ETOOLARGE:
                raise RuntimeError(f"primerforge_b200: {exc} -- a genome of more than 2^26 (67 million) bases per context is "
                                   f"beyond this implementation (see INTEGRATION.md, limits)") from exc
            raise
        if images is not None:
            contig_ids = [seqio.file_ids(images[g], stage.contig_table(g)[0], fmt) for g in range(len(names))]
        n_shared, n_picked = stage.engines[0].counts()
        if n_shared == 0:
            log.error(ERR_NO_SHARED)
            raise RuntimeError(ERR_NO_SHARED)
        if debug:
            _debug_counts(stage, names, log)
        if not sharedExists:
            print(f"done {time.time() - t0:.2f}s")
            log.info(f"{_GAP}done {time.time() - t0:.2f}s")
            if keep:
                params.dumpObj(_shared_dict(stage, names, contig_ids), params.pickles[_SHARED], "shared kmers", prefix=_GAP)

        msg2 = f"{_GAP}evaluating {n_shared} kmers"
        log.info(msg2)
        print(msg2, end=" ... ", flush=True)
        t1 = time.time()
        v0 = stage.engines[0].result_by_id()
        thal_values = None
        if thal_fn is None:
            sel = np.flatnonzero(v0.pick)
        else:
            # hairpin / homodimer screening on the host: the groups are re-formed from the k-mers that passed the
            # three GPU filters, in dict order, and the first member that also passes primer3 is the pick
            ok = np.flatnonzero(v0.alive.astype(bool) & ((v0.flags & 7) == 7))
            cols_ok = stage.columns(ok, names, contig_ids)
            rec = np.zeros(ok.size, dtype=_lib.RECORD_DTYPE)
            rec["flags"] = 7
            positions = []
            for g in range(len(names)):
                contig, left, minus = cols_ok.place(g)
                pos = np.zeros(ok.size, dtype=_lib.POS_DTYPE)
                pos["contig"], pos["start"], pos["strand"] = contig, left, minus
                positions.append(pos)
            picked, cache = screen_groups(rec, positions, words_to_strings(cols_ok.word, cols_ok.k), thal_fn,
                                          params.minTm - params.tempTolerance)
            keep_idx = np.flatnonzero(picked)
            sel = ok[keep_idx]
            thal_values = {i: cache[int(r)] for i, r in enumerate(keep_idx)}
        if sel.size == 0:
            log.error(ERR_NO_CANDIDATES)
            raise RuntimeError(ERR_NO_CANDIDATES)
        cols = stage.columns(sel, names, contig_ids, thal=thal_values)
    finally:
        stage.close()
    out = build_lazy_output(cols)
    print(f"done {time.time() - t1:.2f}s")
    num = out[names[0]].n_candidates()
    print(f"{_GAP}identified {num} candidate kmers")
    log.info(f"{_GAP}done {time.time() - t1:.2f}s")
    log.info(f"{_GAP}identified {num} candidate kmers")
    params.dumpObj(out, params.pickles[_CAND], "candidate kmers", prefix=_GAP)
    return out


def install():
    """splice into an importable primerForge: bin.main resolves the name at call time (main.py:12,86)"""
    import bin.main as ref_main  # type: ignore
    ref_main._getAllCandidateKmers = _getAllCandidateKmers
    return ref_main
