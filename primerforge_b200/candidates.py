"""Drop-in for primerForge's candidate-k-mer stage.

    from primerforge_b200.candidates import _getAllCandidateKmers
    import bin.main; bin.main._getAllCandidateKmers = _getAllCandidateKmers

Same entry point, arguments, return value, error strings and side effects as
/root/reference/bin/getCandidateKmers.py:517-624.  The arithmetic (sketch counting, strand algebra,
cross-genome intersection, relocation, GC / Tm / repeat filters, per-position grouping) runs in
libpfb200.so on the GPU; hairpin / homodimer screening stays primer3 on the host, evaluated lazily
group by group exactly like :376-388.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _lib, seqio
from .engine import ERR_NO_CANDIDATES, ERR_NO_SHARED, Engine, PfbError, StageResult, words_to_strings
from .primer import primer_class

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


def build_output(names, contig_ids, records, positions, kmers, thal_values=None):
    """__buildOutput (:442-489) + the per-contig sort (:598-602): genome -> contig -> list[Primer].

    G x |S| Python objects are the expensive part of the whole stage once the arithmetic runs on the GPU, so
    the records are created column-wise: plain Python lists of the fields, one instance dict per record and
    no constructor call (the constructor would recompute GC and Tm, which come from the device)."""
    import gc
    gc_was_enabled = gc.isenabled()
    gc.disable()        # millions of small containers: the cyclic collector's passes over them cost more than creating them
    try:
        return _build_output(names, contig_ids, records, positions, kmers, thal_values)
    finally:
        if gc_was_enabled:
            gc.enable()


def _build_output(names, contig_ids, records, positions, kmers, thal_values):
    cls = primer_class()
    from .primer import Seq
    k = records["k"].astype(np.int64)
    gc_per = records["gc_count"] / records["k"] * 100
    tm = records["tm"]
    seqs = [Seq(s) for s in kmers]
    n = len(seqs)
    new = cls.__new__
    out = {}
    for g, name in enumerate(names):
        pos = positions[g]
        order = np.lexsort((pos["strand"], k, pos["start"], pos["contig"]))
        left = pos["start"][order].astype(np.int64)
        kk = k[order]
        minus = pos["strand"][order].astype(bool)
        right = left + kk - 1
        starts = np.where(minus, right, left).tolist()          # '-' records keep start > end (Primer.py:42-45)
        ends = np.where(minus, left, right).tolist()
        strands = np.where(minus, "-", "+").tolist()
        order_l = order.tolist()
        seq_o = [seqs[r] for r in order_l]
        tm_o = tm[order].tolist()
        gc_o = gc_per[order].tolist()
        k_o = kk.tolist()
        if thal_values is not None:
            th_o = [thal_values.get(r) or (None, None, None, None) for r in order_l]
        else:
            th_o = None
        contig_sorted = pos["contig"][order]
        # contigs are contiguous runs of the sorted order
        bounds = np.flatnonzero(np.r_[True, contig_sorted[1:] != contig_sorted[:-1], True]) if n else np.zeros(1, dtype=np.int64)
        ids = contig_ids[g]
        per_contig = {}
        for a, b in zip(bounds[:-1].tolist(), bounds[1:].tolist()):
            cname = ids[int(contig_sorted[a])]
            plist = []
            append = plist.append
            for i in range(a, b):
                p = new(cls)
                th = th_o[i] if th_o is not None else (None, None, None, None)
                p.__dict__ = {"seq": seq_o[i], "start": starts[i], "end": ends[i], "contig": cname, "strand": strands[i],
                              "Tm": tm_o[i], "gcPer": gc_o[i], "hairpinTm": th[0], "rcHairpin": th[1],
                              "homodimerTm": th[2], "rcHomodimer": th[3], "_Primer__length": k_o[i]}
                append(p)
            per_contig[cname] = plist
        out[name] = per_contig
    return out


def run_stage(genomes, *, k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4,
              table_size=9999991, prefilter=True, device=0, scan_mode=0, picked_only=False, fasta_images=None,
              contig_ids_out=None, file_format="fasta"):
    """genomes: list of (contig id list, list of uint8 ASCII arrays), or -- fasta_images: list of uint8 FASTA /
    GenBank file images (file_format) -- the files themselves, parsed on the device (their record ids are
    appended to contig_ids_out).  Returns a StageResult; with picked_only it holds just the picked records (the compact
    D2H path of pfb_result_picked)."""
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


def _getAllCandidateKmers(params, sharedExists: bool, thal="auto", device: int | None = None):
    """gets all the candidate kmer sequences for a given ingroup (GPU implementation).

    Args / Raises / Returns: as /root/reference/bin/getCandidateKmers.py:517-530.  `sharedExists` is
    accepted for signature compatibility; the shared k-mers are recomputed (milliseconds on the GPU)
    instead of being unpickled.
    """
    log = params.log
    log.rename(_getAllCandidateKmers.__name__)
    msg1 = f"{_GAP}getting shared ingroup kmers that appear once in each genome"
    log.info(msg1)
    print(msg1, end=" ... ", flush=True)
    t0 = time.time()
    device = int(os.environ.get("PFB200_DEVICE", "0")) if device is None else device
    thal_fn = _thal_provider(params, os.environ.get("PFB200_THAL", thal) if thal == "auto" else thal)

    names, contig_ids, genomes, images = [], [], [], None
    # the files go to the GPU as they are (records found and compacted there); PFB200_HOST_PARSE=1 reads them
    # on the host instead
    fmt = str(params.format).lower()
    fmt = "genbank" if fmt == "gb" else fmt
    if fmt in ("fasta", "genbank") and os.environ.get("PFB200_HOST_PARSE", "0") != "1":
        images = [np.fromfile(fn, dtype=np.uint8) for fn in params.ingroupFns]
        names = [os.path.basename(fn) for fn in params.ingroupFns]
    else:
        for fn in params.ingroupFns:
            ids, contigs = seqio.read_genome(fn, params.format)
            names.append(os.path.basename(fn))
            contig_ids.append(ids)
            genomes.append((ids, contigs))
    try:
        res = run_stage(genomes, k_min=params.minLen, k_max=params.maxLen, min_gc=params.minGc, max_gc=params.maxGc,
                        min_tm=params.minTm, max_tm=params.maxTm, max_repeat=params.maxRepeatLen, device=device,
                        picked_only=thal_fn is None, fasta_images=images, contig_ids_out=contig_ids, file_format=fmt)
    except RuntimeError as exc:
        if str(exc) == ERR_NO_SHARED:
            log.error(ERR_NO_SHARED)
        raise
    print(f"done {time.time() - t0:.2f}s")
    log.info(f"{_GAP}done {time.time() - t0:.2f}s")

    msg2 = f"{_GAP}evaluating {res.timing.get('n_shared_reported', res.records.size)} kmers"
    log.info(msg2)
    print(msg2, end=" ... ", flush=True)
    t1 = time.time()
    if thal_fn is None:
        picked = res.picked
        thal_values = None
    else:
        all_kmers = res.kmer_strings()
        picked, thal_values = screen_groups(res.records, res.positions, all_kmers, thal_fn, params.minTm - params.tempTolerance)
    sel = np.flatnonzero(picked)
    if sel.size == 0:
        log.error(ERR_NO_CANDIDATES)
        raise RuntimeError(ERR_NO_CANDIDATES)
    rec = res.records[sel]
    kmers = words_to_strings(rec["word"], rec["k"])
    pos = [p[sel] for p in res.positions]
    if thal_values is not None:
        thal_values = {i: thal_values[int(r)] for i, r in enumerate(sel)}
    out = build_output(names, contig_ids, rec, pos, kmers, thal_values)
    print(f"done {time.time() - t1:.2f}s")
    num = sum(len(v) for v in out[names[0]].values())
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
