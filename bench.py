#!/usr/bin/env python
"""bench.py -- candidate-k-mer stage throughput (ingroup Mbp/s) on N B200s, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl reference]

A step = one full pass of the stage (2-bit encode, genome-1 candidate table, key table, scan of every other
genome, uniqueness rules, ordered emit, filters, grouping) over one batch of synthetic genomes of
BASELINE.json's config (default: configs[1] = 10 x 5 Mbp, primer_len 16-20, default GC/Tm).

  value : device-resident ASCII bases -> device-resident results, CUDA events on the library's
          stream, L2 flushed between steps, max over ranks.
  e2e   : the same stage through the C ABI with HOST (pinned) buffers: H2D of the bases, compute, the per-id
          result columns streaming to pinned host memory while the run goes on.
  N > 1 : launched by torchrun, one rank per GPU; every rank holds genome 0 (it defines the key table) plus its own
          shard of the other genomes; genome 0's own pipeline is shared out by position; the exchanges (candidate
          lists, packed rows, alive and pick bytes) happen inside the library over mapped peer memory or NCCL.
          Weak scaling: 1 + 9 genomes per rank, throughput counts DISTINCT ingroup bases.  Rank 0 also runs the
          whole job on ONE GPU outside the timed region and compares (the `parity` block), and the line carries a
          `strong_c4` block: BASELINE config 4's 200 genomes dealt over the ranks.
  blocks: `parity` (GPU vs oracle on the cpu_baseline sample), `outgroup` (outgroup screen on config 3's outgroup
          shape), `pairs` (pair search on the drop-in leg's candidates), `e2e_fasta`, `e2e_dropin`, `cpu_baseline`.
  --impl reference : the CPU restatement of the reference (oracle/pf_oracle.c, OpenMP, all host
          threads) on a bounded sample of the same workload; the reference itself is pure Python +
          khmer/primer3/pyahocorasick which are not installable in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate-k-mer stage ingroup Mbp/s"
UNIT = "Mbp/s"
SURVEY_BYTES_PER_BASE = {1: 139.0, 2: 99.0, 3: 99.0, 4: 99.0, 5: 248.0}   # SURVEY.md 8(d) worked estimates


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if smax and x > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads() -> int:
    """every host thread the process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 per rank)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_run(cfg_index: int, threads: int = 0, n_genomes: int = 3, length: int | None = None):
    """the CPU restatement (oracle/pf_oracle.c, OpenMP) on a bounded sample of the workload: the FIRST `n_genomes`
    genomes of the config's generator at the config's FULL genome length (so the sketch load -- the share of unique
    k-mers khmer's table drops -- is the workload's own), same k range and filters.  Returns (cpu_baseline dict,
    seconds, oracle result, genomes): bench.py's parity block compares the GPU with this very result."""
    from oracle import pf_oracle as po
    from primerforge_b200 import synth
    cfg = synth.CONFIGS[cfg_index]
    length = length or cfg["length"]
    threads = threads or host_threads()
    gens = synth.make_genomes(n_genomes, length, seed=synth.BASE_SEED + cfg_index, n_contigs=cfg["n_contigs"])
    par = po.default_params(k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], n_threads=threads)
    t0 = time.perf_counter()
    o = po.run(gens, par)
    dt = time.perf_counter() - t0
    nk = cfg["k_max"] - cfg["k_min"] + 1
    return dict(value=n_genomes * length / 1e6 / dt, unit=UNIT, cores=threads, kind="port",
                sample=f"the first {n_genomes} of the config-{cfg_index} genomes at {length} bp each ({n_genomes * length / 1e6:g} Mbp), "
                       f"k {cfg['k_min']}-{cfg['k_max']}, oracle/pf_oracle.c (C restatement of the reference, OpenMP over "
                       f"{n_genomes * nk} (genome, k) tasks on {threads} threads), {dt:.2f} s, {o.n_shared} shared k-mers"), dt, o, gens


def outgroup_block(device: int, flush, with_cpu: bool, n_outgroup: int = 10, length: int = 5_000_000, n_pairs: int = 50_000,
                   steps: int = 10) -> dict:
    """SURVEY.md 8(f)-3, the outgroup screen (bin/removeOutgroupPrimers.py:181-313) on BASELINE config 3's outgroup
    shape: 10 outgroup genomes of 5 Mbp (the first ingroup genome with 5 % substitutions) against the primers of
    50 000 synthetic pairs.  Device-resident steps (CUDA events inside the library, L2 flushed), the same through the C
    ABI from pinned host bases to host results, and -- the gate -- the result compared with oracle/pf_oracle_outgroup.c
    run on the same inputs with every host thread."""
    import torch
    from primerforge_b200 import synth
    from primerforge_b200.engine import words_to_strings
    from primerforge_b200.outgroup import OutgroupScreen
    first = synth.make_genomes(1, length, seed=synth.BASE_SEED + 3)[0]
    outs = synth.make_outgroup(first, n_outgroup)
    w, k, pf, pr = synth.make_pairs(first, n_pairs)
    pinned = []
    for g in outs:
        cat = np.concatenate(g.contigs)
        t = torch.empty(cat.size, dtype=torch.uint8).pin_memory()
        t.numpy()[:] = cat
        pinned.append((t, np.cumsum([0] + [c.size for c in g.contigs]).astype(np.uint64)))
    bases = sum(t.numel() for t, _ in pinned)
    with OutgroupScreen(device=device) as scr:
        scr.set_keys_and_pairs(w, k, pf, pr, 120, 2400)

        def add_all():
            scr.clear_genomes()
            for t, off in pinned:
                scr._ck(scr._lib.pfb_og_add_genome(scr._h, t.data_ptr(), t.numel(), off.ctypes.data, off.size - 1))
        add_all()
        scr.run()
        ms, parts = [], []
        for i in range(3 + steps):
            flush.zero_()
            torch.cuda.synchronize()
            scr.compute()
            if i >= 3:
                tm = scr.timing()
                ms.append(tm["ms_total"]); parts.append((tm["ms_encode"], tm["ms_scan"], tm["ms_bucket"], tm["ms_products"]))
        e2e = []
        for i in range(2 + steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            add_all()
            scr.run()                       # H2D of the bases, scan, products, D2H of hits / products / verdicts
            if i >= 2:
                e2e.append(time.perf_counter() - t0)
        res = scr.result()
        tm = scr.timing()
    step_ms, e2e_ms = float(np.mean(ms)), float(np.mean(e2e) * 1e3)
    enc, scan, bucket, prod = (float(x) for x in np.mean(np.array(parts), axis=0))
    peak, _ = measured_peaks()
    out = {"what": "outgroup screen (bin/removeOutgroupPrimers.py:181-313): every outgroup contig on both strands against the pairs' primers, "
                   "product sizes per pair, pairs with a disallowed size deleted",
           "workload": f"BASELINE config 3's outgroup: {n_outgroup} x {length / 1e6:g} Mbp (first ingroup genome, 5 % substitutions), "
                       f"{n_pairs} synthetic pairs = {int(w.size)} distinct primers of 16-20 bases, disallowed sizes 120-2400",
           "value": bases / 1e6 / (step_ms / 1e3), "unit": "outgroup Mbp/s", "ms_per_step": step_ms, "steps": steps,
           "phases_ms": {"encode": enc, "scan": scan, "bucket": bucket, "products": prod},
           "e2e": {"value": bases / 1e6 / (e2e_ms / 1e3), "unit": "outgroup Mbp/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(bases),
                   "d2h_bytes_per_step": int(16 * tm["n_hits"] + 16 * tm["n_products"] + 4 * n_pairs)},
           "gpu_launches_per_step": int(tm["n_launches"]),
           "counts": {k2: int(tm[k2]) for k2 in ("n_bases", "n_windows", "n_hits", "n_products", "n_probe_occupied", "table_bytes",
                                                 "filter_in_smem", "filter_bits_per_record")} | {"pairs_kept": int(res.kept.sum())},
           "roofline": {"kernel": "k_og_scan_smem", "ms_per_launch": scan, "windows_per_s_G": tm["n_windows"] / (scan / 1e3) / 1e9,
                        "algorithmic_bytes_per_launch": 0.375 * bases, "hbm_frac": 0.375 * bases / (scan / 1e3) / 1e9 / peak,
                        "bound": "instruction issue + shared-memory filter look-ups (one per window and length); HBM only streams the packed "
                                 "bases once (0.375 B per base), so the HBM fraction is small by construction"}}
    if with_cpu:
        from oracle import pf_oracle as po
        t0 = time.perf_counter()
        o = po.run_outgroup(words_to_strings(w, k), pf, pr, 120, 2400, outs)
        dt = time.perf_counter() - t0
        h = res.hits
        mine = np.stack([h["key"], h["genome"], h["contig"], h["start"], h["strand"]], axis=1).astype(np.uint32)
        theirs = o.hits[np.lexsort((o.hits[:, 3], o.hits[:, 0], o.hits[:, 4], o.hits[:, 2], o.hits[:, 1]))]
        pp = res.products
        ok_h = bool(np.array_equal(mine, theirs))
        ok_r = bool(np.array_equal(res.removed_at, o.removed_at))
        ok_p = bool(np.array_equal(np.stack([pp["pair"], pp["genome"], pp["contig"], pp["size"].astype(np.uint32)], axis=1).reshape(-1, 4), o.products))
        out["cpu_baseline"] = {"value": bases / 1e6 / dt, "unit": "outgroup Mbp/s", "cores": host_threads(), "kind": "port",
                               "sample": f"the same {n_outgroup} genomes and pairs, oracle/pf_oracle_outgroup.c (C restatement of the reference's "
                                         f"string scan, OpenMP over the contigs), {dt:.2f} s"}
        out["parity"] = {"against": "oracle/pf_oracle_outgroup.c on the same inputs", "ok": ok_h and ok_r and ok_p, "hits": ok_h,
                         "deleted_pairs": ok_r, "product_sets": ok_p, "n_hits": int(mine.shape[0])}
    return out


def pairs_block(candidates, first_name: str, device: int, flush, with_cpu: bool, steps: int = 10) -> dict:
    """SURVEY.md 8(f)-4, the pair search (bin/getPrimerPairs.py:505-564) on the candidates the drop-in leg has just
    produced (lazy containers: their columns go to the GPU as they are): binning of every genome, the substring
    reduction, bin pairs, the first suitable primer pair of each, the look-up in every other genome.  The heterodimer
    check is primer3's and is not part of the timed arithmetic.  The gate: the same columns through
    oracle/pf_oracle_pairs.c."""
    import torch
    from primerforge_b200.pairs import PairSearch, _columns
    kw = dict(max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0)      # bin/Parameters.py defaults
    word, k, tm, names, per_genome, _lists = _columns(candidates, first_name)
    with PairSearch(device=device) as ps:
        ps.set_candidates(word, k, tm)
        for g in per_genome:
            ps.add_genome(*g)
        ms, parts, wall = [], [], []
        for i in range(2 + steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ps.set_candidates(word, k, tm)      # H2D of the columns ...
            for g in per_genome:
                ps.add_genome(*g)
            ps.run(**kw)                  # ... every kernel, D2H of the pairs and their per-genome product sizes
            dt = time.perf_counter() - t0
            if i >= 2:
                t = ps.timing()
                ms.append(t["ms_total"]); parts.append((t["ms_bin"], t["ms_pairs"], t["ms_shared"])); wall.append(dt)
        res = ps.result()
        t = ps.timing()
    S, G = int(word.size), len(per_genome)
    b, p2, sh = (float(x) for x in np.mean(np.array(parts), axis=0))
    step_ms = float(np.mean(ms))
    out = {"what": "pair search (bin/getPrimerPairs.py:505-564) without primer3's heterodimer check",
           "workload": f"the {S} candidates of the drop-in leg in {G} genomes, maxBinSize 64, product size 120-2400, maxTmDiff 5",
           "value": S * G / 1e6 / (step_ms / 1e3), "unit": "M candidate placements/s", "ms_per_step": step_ms, "steps": steps,
           "phases_ms": {"binning": b, "reduce_binpairs_primerpairs": p2, "other_genomes": sh},
           "e2e": {"ms_per_step": float(np.mean(wall) * 1e3), "h2d_bytes_per_step": int(S * 17 + G * S * 13),
                   "d2h_bytes_per_step": int(t["n_pairs"] * (25 + 4 * G))},
           "gpu_launches_per_step": int(t["n_launches"]),
           "counts": {"n_candidates": S, "n_genomes": G, "n_bins_first_genome": int(t["n_bins"]), "n_pairs": int(t["n_pairs"]),
                      "n_kept": int(t["n_kept"])}}
    if with_cpu:
        from oracle import pf_oracle as po
        t0 = time.perf_counter()
        o = po.run_pairs(word, k, tm, per_genome, **kw)
        dt = time.perf_counter() - t0
        pr = res.pairs
        pairs = np.stack([pr["fwd"], pr["rev"], pr["contig"], pr["fbin"], pr["rbin"], pr["pcr_len"]], axis=1).astype(np.uint32).reshape(-1, 6)
        shared = res.shared4()
        ok_p, ok_s = bool(np.array_equal(pairs, o.pairs)), bool(np.array_equal(shared, o.shared))
        out["cpu_baseline"] = {"value": S * G / 1e6 / dt, "unit": "M candidate placements/s", "cores": 1, "kind": "port",
                               "sample": f"the same columns, oracle/pf_oracle_pairs.c (C restatement of the reference's list walk, one thread), {dt:.2f} s"}
        out["parity"] = {"against": "oracle/pf_oracle_pairs.c on the same columns", "ok": ok_p and ok_s, "pairs_in_order": ok_p,
                         "per_genome_products": ok_s, "n_pairs": int(pairs.shape[0])}
    return out


def parity_vs_oracle(gens, o, cfg, device: int) -> dict:
    """the correctness gate: the CUDA path on the very genomes the CPU sample ran on, compared
    bit for bit with the oracle's result -- the shared k-mers in dict order, flags and picks, Tm, and
    (contig, start, strand) in every genome -- in both output modes (all shared k-mers / prefiltered, the timed one)"""
    from primerforge_b200.engine import Engine
    out = {"against": "oracle/pf_oracle.c on the cpu_baseline sample", "ok": True, "checked": []}
    for prefilter in (0, 1):
        with Engine(device=device, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=prefilter) as eng:
            for g in gens:
                eng.add_genome(g.contigs)
            eng.run()
            res = eng.result()
        rec = res.records
        keep = ((o.flags & 7) == 7) if prefilter else np.ones(o.n_shared, dtype=bool)
        checks = {"n_records": int(rec.size), "n_oracle": int(keep.sum())}
        same = rec.size == int(keep.sum())
        if same:
            checks["words_and_order"] = bool(np.array_equal(rec["word"], o.word[keep]) and np.array_equal(rec["k"], o.k[keep]))
            checks["flags_and_picks"] = bool(np.array_equal(rec["flags"] & 15, o.flags[keep]))
            checks["tm_max_abs_diff"] = float(np.max(np.abs(rec["tm"] - o.tm[keep]), initial=0.0))
            checks["gc"] = bool(np.array_equal(rec["gc_count"].astype(np.int32), o.gc_count[keep]))
            checks["positions"] = all(
                np.array_equal(res.positions[gi]["contig"].astype(np.int32), o.contig[gi][keep]) and
                np.array_equal(res.positions[gi]["start"].astype(np.int64), o.start[gi][keep]) and
                np.array_equal(res.positions[gi]["strand"].astype(np.uint8), o.strand[gi][keep]) for gi in range(len(gens)))
            same = checks["words_and_order"] and checks["flags_and_picks"] and checks["gc"] and checks["positions"] and \
                checks["tm_max_abs_diff"] <= 0.01
        checks["prefilter"] = prefilter
        checks["ok"] = bool(same)
        out["checked"].append(checks)
        out["ok"] = bool(out["ok"] and same)
    return out


def run_reference_arm(args, emit):
    """the reference's CPU implementation of the path on the host cores.  The reference itself is pure Python whose
    four compiled dependencies cannot be installed in this image, so this times the C
    restatement (`kind: port`) with every host thread, each step one bounded sample of the arm's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from primerforge_b200 import synth
    cfg = synth.CONFIGS[args.config]
    vals, last = [], None
    # every step is a bounded sample (the first n genomes at full length); it shrinks when K + W steps of it would
    # take more than ~3 minutes: first to 2 genomes, then in length
    n_g, length = 3, (getattr(args, "length", None) or cfg["length"])
    _cb, dt0, _o, _g = cpu_sample_run(args.config, n_genomes=n_g, length=length)
    budget = 170.0
    total = max(1, args.warmup + args.steps)
    if dt0 * total > budget:
        n_g = 2
        dt0 *= 2.0 / 3.0
    if dt0 * total > budget:
        length = int(max(100_000, length * budget / (dt0 * total)))
    for i in range(args.warmup + args.steps):
        cb, dt, _o, _g = cpu_sample_run(args.config, n_genomes=n_g, length=length)
        if i >= args.warmup:
            vals.append((cb["value"], dt))
        last = cb
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    last["value"] = value
    # the arm's config, plus what one step of THIS line actually ran (a sample, not the whole workload)
    conf = workload_config(args.config, 1)
    conf["reference_step"] = last["sample"]
    conf["same_workload_as_gpu_arm"] = False
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "reference", "config": conf, "cpu_baseline": last,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def workload_config(index: int, n_gpus: int, strong: bool = False) -> dict:
    from primerforge_b200 import synth
    from primerforge_b200.dist import shard_genomes
    cfg = synth.CONFIGS[index]
    if strong and n_gpus > 1:
        total = cfg["n_genomes"]
        per_rank = len(shard_genomes(total, n_gpus, 0))
        shape = f"{total} synthetic {cfg['length'] / 1e6:g} Mbp ingroup genomes in total, genome 0 on every GPU + {per_rank - 1} others per GPU"
    else:
        per_rank = cfg["n_genomes"]
        total = per_rank if n_gpus == 1 else 1 + (per_rank - 1) * n_gpus
        shape = (f"{per_rank} synthetic {cfg['length'] / 1e6:g} Mbp ingroup genomes per GPU "
                 f"({total} distinct genomes in the job)")
    return {"workload": f"BASELINE config {index}: {shape}, primer_len {cfg['k_min']},{cfg['k_max']}, default GC/Tm, "
                        f"max_repeats {cfg['max_repeats']}, {cfg['n_contigs']} contig(s) per genome",
            "genomes_per_gpu": per_rank, "distinct_genomes": total, "genome_bp": cfg["length"], "k_min": cfg["k_min"],
            "k_max": cfg["k_max"], "table_size": 9999991, "prefilter": 1, "l2": "flushed between steps (256 MiB write)",
            "parallelism": f"genome-sharded x{n_gpus}, key table replicated" if n_gpus > 1 else "single GPU"}


def pin_genomes(gens):
    """(pinned uint8 tensor, contig offsets) per genome: the host side of the e2e leg"""
    import torch
    pinned = []
    for g in gens:
        cat = np.concatenate(g.contigs) if len(g.contigs) > 1 else g.contigs[0]
        t = torch.empty(cat.size, dtype=torch.uint8).pin_memory()
        t.numpy()[:] = cat
        off = np.zeros(len(g.contigs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([c.size for c in g.contigs])
        pinned.append((t, off))
    return pinned


def view_digests(view, with_records: bool) -> dict:
    """digests of one context's per-id result columns: which ids are picked, the picked ids' records, and their place
    in every genome of the context (local order)"""
    import hashlib
    sel = np.flatnonzero(view.pick)
    out = {"n_ids": int(view.n_ids), "n_picked": int(sel.size),
           "pick": hashlib.blake2b(np.ascontiguousarray(view.pick).tobytes(), digest_size=16).hexdigest(),
           "alive": hashlib.blake2b(np.ascontiguousarray(view.alive).tobytes(), digest_size=16).hexdigest(),
           "pos": [hashlib.blake2b(np.ascontiguousarray(view.pos[g][sel]).tobytes(), digest_size=16).hexdigest()
                   for g in range(view.pos.shape[0])]}
    if with_records and view.word is not None:
        h = hashlib.blake2b(digest_size=16)
        for col in (view.word, view.tm, view.info):
            h.update(np.ascontiguousarray(col[sel]).tobytes())
        out["records"] = h.hexdigest()
    return out


def main():
    # exactly ONE line may reach stdout (the driver parses it): libraries that print banners during
    # initialisation (NCCL's version line) are diverted to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=None, help="override genomes per GPU (debug)")
    ap.add_argument("--length", type=int, default=None, help="override genome length (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fasta-leg", action="store_true")
    ap.add_argument("--no-strong-c4", action="store_true", help="skip the strong-scaling block (BASELINE config 4: 200 genomes over the GPUs)")
    ap.add_argument("--no-dropin-leg", action="store_true", help="skip the timing of the drop-in entry point from FASTA files on disk")
    ap.add_argument("--no-pairs", action="store_true", help="skip the pair-search block (the drop-in leg's candidates through _getPrimerPairs' arithmetic)")
    ap.add_argument("--no-outgroup", action="store_true", help="skip the outgroup-screen block (BASELINE config 3's 10 outgroup genomes)")
    ap.add_argument("--scan-mode", type=int, default=0, help="0 auto, 1 plain L2 probe, 2 shared-memory filter (debug)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): the config's genome count PER GPU; strong: the config's genome count in total, "
                         "genome 0 on every rank and the others dealt round-robin")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args, emit)
        return

    import torch
    import torch.distributed as dist
    from primerforge_b200 import synth
    from primerforge_b200.dist import attach_comm, shard_genomes
    from primerforge_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    # pinned host buffers should live on the NUMA node next to this rank's GPU: bind the process to the CPUs
    # NVML names for the device before anything is allocated (no-op where NVML refuses)
    affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = torch.cuda.get_device_properties(local).uuid
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        affinity = len(os.sched_getaffinity(0))
    except Exception as exc:      # noqa: BLE001
        print(f"[bench] NUMA binding skipped: {exc}", file=sys.stderr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg = dict(synth.CONFIGS[args.config])
    if args.genomes:
        cfg["n_genomes"] = args.genomes
    if args.length:
        cfg["length"] = args.length
    strong = args.scaling == "strong" and world > 1
    seed = synth.BASE_SEED + args.config
    if strong:
        n_distinct = cfg["n_genomes"]
        idx = shard_genomes(n_distinct, world, rank)
        per_rank = len(shard_genomes(n_distinct, world, 0))          # the largest shard
    else:
        per_rank = cfg["n_genomes"]
        n_distinct = per_rank if world == 1 else 1 + (per_rank - 1) * world
        idx = list(range(per_rank)) if world == 1 else [0] + list(range(1 + (per_rank - 1) * rank, 1 + (per_rank - 1) * (rank + 1)))
    # (the generator draws every genome's random numbers whatever is asked for: genome g is the same sequence on every rank)
    gens = synth.make_genomes(max(idx) + 1, cfg["length"], seed=seed, n_contigs=cfg["n_contigs"], only=idx)
    mine = [gens[i] for i in idx]
    del gens
    distinct_bases = n_distinct * cfg["length"]
    pinned = pin_genomes(mine)          # the e2e leg copies from these every step
    del mine

    ekw = dict(k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
    eng = Engine(device=local, **ekw)
    attach_comm(eng)                    # N > 1: the exchanges between the ranks happen inside the library from here on
    eng.set_scan_mode(args.scan_mode)
    for t, off in pinned:
        eng.add_genome_ptr(t.data_ptr(), t.numel(), off)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg
    eng.upload()
    for _ in range(args.warmup):
        flush.zero_()
        barrier()
        eng.compute()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.35)        # nvidia-smi needs a moment to start printing; the loop below keeps the GPU busy
    dev_ms, launches, scan_ms, scan_main_ms, enc_ms, first_ms, fin_ms, x_ms = [], 0, [], [], [], [], [], []
    dev_counts = {}
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        eng.compute()
        tm = eng.timing()
        dev_ms.append(tm["ms_total"])
        scan_ms.append(tm["ms_scan"])
        scan_main_ms.append(tm["ms_scan_main"])
        enc_ms.append(tm["ms_encode"]); first_ms.append(tm["ms_first"]); fin_ms.append(tm["ms_finalize"])
        x_ms.append(tm["ms_exchange"])
        dev_counts = {k: int(tm[k]) for k in ("n_bases", "n_windows", "n_slots", "n_ids", "n_records", "n_picked")}
        launches += tm["n_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    tm = eng.timing()
    reruns_resident = int(tm["n_reruns"])

    # ---- end-to-end leg (host buffers -> host results through the C ABI): pfb_run queues the H2D copies of the
    # bases, computes behind them and copies the per-id result columns into pinned host memory WHILE it computes;
    # the step ends when the last result byte is on the host
    e2e_s, e2e_parts = [], []
    h2d = sum(t.numel() for t, _ in pinned)
    d2h = 0
    view = None
    for i in range(max(1, args.warmup // 2) + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        eng.run()                 # collective at N > 1
        t1 = time.perf_counter()
        view = eng.result_by_id()
        eng.sync()
        dt = time.perf_counter() - t0
        if i >= max(1, args.warmup // 2):
            e2e_s.append(dt)
            e2e_parts.append((t1 - t0, dt - (t1 - t0), eng.timing()["ms_total"]))
            d2h = view.pitch * ((18 if view.word is not None else 0) + 4 * view.pos.shape[0] + 2)
    clocks = sampler.stop() if rank == 0 else None
    my_digest = view_digests(view, with_records=rank == 0)
    reruns_total = int(eng.timing()["n_reruns"])

    # ---- the gate at N > 1: the job's result (every rank's picked ids, records, places of its genomes) against ONE
    # GPU running all the job's genomes -- the only way a multi-GPU line can vouch for the sharded path
    parity = None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx, my_digest))
        if rank == 0:
            all_g = synth.make_genomes(n_distinct, cfg["length"], seed=seed, n_contigs=cfg["n_contigs"])
            with Engine(device=local, **ekw) as one:
                for g in all_g:
                    one.add_genome(g.contigs)
                del all_g
                one.upload()
                one.compute()
                ref = view_digests(one.result_by_id(), with_records=True)
            ok = True
            for gi_list, dg in gathered:
                ok &= dg["pick"] == ref["pick"] and dg["alive"] == ref["alive"] and dg["n_ids"] == ref["n_ids"]
                ok &= all(dg["pos"][j] == ref["pos"][gi] for j, gi in enumerate(gi_list))
                if "records" in dg:
                    ok &= dg["records"] == ref["records"]
            parity = {"against": f"the {n_distinct} genomes of the job on ONE GPU (rank 0, outside the timed region)",
                      "ok": bool(ok), "n_ids": ref["n_ids"], "n_picked": ref["n_picked"],
                      "checked": "alive and pick bytes of every rank, records on rank 0, places of every genome on the rank that holds it"}
        barrier()

    # ---- end-to-end from the FASTA FILE IMAGES (80-column, pinned host memory): the records are found and
    # compacted on the device (pfb_add_fasta), everything else as above.  Single GPU only.
    fasta_leg = None
    if world == 1 and not args.no_fasta_leg:
        images = []
        for (t, off), gi in zip(pinned, range(len(pinned))):
            parts = []
            arr = t.numpy()
            for c in range(len(off) - 1):
                seq = arr[int(off[c]):int(off[c + 1])]
                parts.append(np.frombuffer(f">g{gi:03d}_c{c:03d} synthetic\n".encode(), dtype=np.uint8))
                full = (seq.size // 80) * 80
                if full:
                    body = np.empty((seq.size // 80, 81), dtype=np.uint8)
                    body[:, :80] = seq[:full].reshape(-1, 80)
                    body[:, 80] = 10
                    parts.append(body.reshape(-1))
                if seq.size > full:
                    parts.append(seq[full:])
                    parts.append(np.frombuffer(b"\n", dtype=np.uint8))
            img = np.concatenate(parts)
            pt = torch.empty(img.size, dtype=torch.uint8).pin_memory()
            pt.numpy()[:] = img
            images.append(pt)
        eng_f = Engine(device=local, **ekw)
        for pt in images:
            eng_f.add_fasta(ptr=pt.data_ptr(), n_bytes=pt.numel())
        fa_s = []
        vf = None
        for i in range(2 + min(args.steps, 50)):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            eng_f.run()
            vf = eng_f.result_by_id()
            eng_f.sync()
            if i >= 2:
                fa_s.append(time.perf_counter() - t0)
        fa_ms = float(np.mean(fa_s) * 1e3)
        fasta_leg = {"value": distinct_bases / 1e6 / (fa_ms / 1e3), "unit": UNIT, "ms_per_step": fa_ms,
                     "h2d_bytes_per_step": int(sum(pt.numel() for pt in images)),
                     "d2h_bytes_per_step": int(vf.pitch * (18 + 4 * vf.pos.shape[0] + 2)),
                     "n_picked": int(np.count_nonzero(vf.pick)), "input": "80-column FASTA file images in pinned host memory, parsed on the device"}
        eng_f.close()
        del images

    # ---- strong scaling of BASELINE config 4 (200 x 5 Mbp over the GPUs of the job; north_star's target is >= 6x at 8):
    # device-resident steps, max over ranks; rank 0 also times the whole of config 4 on its one GPU so that the
    # line carries its own speed-up
    strong_c4 = None
    if not args.no_strong_c4 and not args.genomes and not args.length:
        c4 = synth.CONFIGS[4]
        idx4 = shard_genomes(c4["n_genomes"], world, rank)
        k4 = dict(k_min=c4["k_min"], k_max=c4["k_max"], max_repeat=c4["max_repeats"], prefilter=1)
        steps4 = max(3, min(args.steps, 10))

        def time_c4(genome_idx, with_comm):
            g4 = synth.make_genomes(max(genome_idx) + 1, c4["length"], seed=synth.BASE_SEED + 4, n_contigs=c4["n_contigs"], only=genome_idx)
            e4 = Engine(device=local, **k4)
            if with_comm:
                attach_comm(e4)
            keep = []
            for gi in genome_idx:
                arr = g4[gi].contigs[0]
                g4[gi] = None
                keep.append(arr)
                e4.add_genome(bases=arr, offsets=np.array([0, arr.size], dtype=np.uint64))
            e4.upload()
            ms, xs = [], []
            for i in range(3 + steps4):
                flush.zero_()
                if with_comm:
                    barrier()
                else:
                    torch.cuda.synchronize()
                e4.compute()
                if i >= 3:
                    ms.append(e4.timing()["ms_total"]); xs.append(e4.timing()["ms_exchange"])
            t4 = e4.timing()
            e4.close()
            return float(np.mean(ms)), float(np.mean(xs)), t4

        ms4, x4, t4 = time_c4(idx4, world > 1)
        if world > 1:
            red = torch.tensor([ms4], dtype=torch.float64, device="cuda")
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            ms4 = float(red.item())
        bases4 = c4["n_genomes"] * c4["length"]
        strong_c4 = {"workload": f"BASELINE config 4: {c4['n_genomes']} x {c4['length'] / 1e6:g} Mbp ingroup genomes in total, "
                                 f"genome 0 on every GPU + {len(idx4) - 1} others on rank 0's, primer_len {c4['k_min']},{c4['k_max']}",
                     "n_gpus": world, "steps": steps4, "ms_per_step": ms4, "value": bases4 / 1e6 / (ms4 / 1e3), "unit": UNIT,
                     "ms_exchange_rank0": x4, "sharded_first_genome": int(t4["sharded_first_genome"]),
                     "n_ids": int(t4["n_ids"]), "n_slots": int(t4["n_slots"])}
        if world > 1:
            # bytes a rank sends per direction (ring / tree all-reduce ~ 2 (N-1)/N of the buffer; all-gather (N-1)/N)
            # against the measured NVLink peer bandwidth
            nid, nsl = int(t4["n_ids"]), int(t4["n_slots"])
            f = (world - 1) / world
            wire = 2 * f * (nid + nid) + 2 * f * 4 * nsl + f * 4 * 0.1 * 25e6        # alive + pick, packed rows, candidate lists (~10 % of genome 1's windows)
            strong_c4["exchange"] = {"what": "all-gather of genome 1's candidate lists, SUM all-reduce of its packed rows (4 B per key), "
                                             "MIN all-reduce of the alive bytes, MAX all-reduce of the pick bytes (NCCL inside the library)",
                                     "wire_bytes_per_rank_per_direction": int(wire), "ms": x4,
                                     "nvlink_peak_GBs_per_direction": 770.0,
                                     "frac_of_nvlink": wire / (x4 / 1e3) / 1e9 / 770.0 if x4 > 0 else None}
            if rank == 0:
                ms1, _x, _t = time_c4(list(range(c4["n_genomes"])), False)
                strong_c4["ms_per_step_one_gpu"] = ms1
                strong_c4["speedup_vs_one_gpu"] = ms1 / ms4
            barrier()
        else:
            strong_c4["speedup_vs_one_gpu"] = 1.0

    step_ms = float(np.mean(dev_ms))
    e2e_ms = float(np.mean(e2e_s) * 1e3)
    if world > 1:
        red = torch.tensor([step_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = red.tolist()
        tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        h2d, d2h = (int(x) for x in tot.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = distinct_bases / 1e6 / (step_ms / 1e3)
    e2e_val = distinct_bases / 1e6 / (e2e_ms / 1e3)
    peak, peak_kind = measured_peaks()
    # roofline of the dominant kernel: the ONE k_scan_filtered launch over genomes 2..G (CUDA events around
    # that launch inside the library, averaged over the timed steps)
    scan_bases = tm["scan_main_bases"]
    bpb = SURVEY_BYTES_PER_BASE[args.config]
    ms_scan = float(np.mean(scan_main_ms)) if scan_main_ms else 0.0
    alg_bytes = bpb * scan_bases
    achieved = alg_bytes / (ms_scan / 1e3) / 1e9 if ms_scan > 0 else 0.0
    nk = cfg["k_max"] - cfg["k_min"] + 1
    windows = scan_bases * nk
    traffic, traffic_src, stored = None, None, {}
    try:
        with open(next(p for p in (os.path.join(ROOT, "profiles", n) for n in ("r02_scan_traffic.json", "r01_scan_traffic.json")) if os.path.exists(p))) as fh:
            tj = json.load(fh)
        stored = tj
        if args.config == 2 and world == 1 and not args.genomes and not args.length:
            traffic, traffic_src = tj["dram_bytes_total"], tj["source"] + " -- a STORED figure of that capture, not measured in this run"
    except Exception:
        pass
    stage_bytes = bpb * distinct_bases
    roofline = {"bound": "hbm", "kernel": "k_scan_filtered (genomes 2..G, one launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_base": bpb,
                "bytes_per_base_source": "SURVEY.md 8(d) stage model (count-everything design); see DESIGN.md section 4",
                "ms_per_launch": ms_scan, "share_of_step": ms_scan / step_ms if step_ms else None,
                "stage_level": {"what": "the same byte model over the WHOLE step (all launches, per GPU)",
                                "achieved": stage_bytes / world / (step_ms / 1e3) / 1e9, "frac": stage_bytes / world / (step_ms / 1e3) / 1e9 / peak},
                "actual_limiter": {"kind": "instruction issue / INT32 ALU pipe (hash + filter + queue append per window); the kernel's DRAM traffic is "
                                           "~10 % of the modelled bytes, i.e. ~10 % of HBM peak: HBM is not the binding roof",
                                   "windows_per_s_G": windows / (ms_scan / 1e3) / 1e9 if ms_scan > 0 else 0.0,
                                   "windows_per_clock_per_sm": windows / (ms_scan / 1e3) / 148 / ((clocks or {}).get("sm_mhz") or 1965.0) / 1e6 if ms_scan > 0 else 0.0,
                                   "stored_ncu_capture": {"issue_active_pct": stored.get("issue_active_pct"), "alu_pipe_pct": stored.get("alu_pipe_pct"),
                                                          "thread_instructions_per_window": (stored["warp_instructions"] * 32 / 225e6) if stored.get("warp_instructions") else None,
                                                          "source": stored.get("source")},
                                   "evidence": "profiles/README.md (ncu summaries, SASS opcode histogram), profiles/r01_scan_ablation.md"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(args.config, world, strong) | ({"genomes_per_gpu": per_rank, "genome_bp": cfg["length"]}),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "ms_run_call": float(np.mean([a for a, _, _ in e2e_parts]) * 1e3), "ms_after_run_call": float(np.mean([b for _, b, _ in e2e_parts]) * 1e3),
                    "ms_device_span_in_run": float(np.mean([c for _, _, c in e2e_parts])),
                    "what": "pfb_run from pinned host bases to the per-id result columns (word, Tm, info, place in every genome, alive, pick) "
                            "in pinned host memory; copies in both directions overlap the kernels; bytes are summed over the ranks"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "host_cpus_bound": affinity,
            "phases_ms": {"ms_encode": float(np.mean(enc_ms)), "ms_first": float(np.mean(first_ms)), "ms_scan": float(np.mean(scan_ms)),
                          "ms_finalize": float(np.mean(fin_ms)), "ms_total": step_ms},
            "counts": dev_counts, "reruns": {"resident_leg": reruns_resident, "all_legs": reruns_total,
                                             "what": "runs repeated because a buffer capacity proved too small (the first run of a context sizes them)"},
            "wall_s_timed_loop": wall}
    if world > 1:
        # ring all-reduce: every rank sends and receives 2 (N - 1) / N of the buffer per direction
        ms_x = float(np.mean(x_ms)) if x_ms else 0.0
        f = (world - 1) / world
        nid, nsl = dev_counts.get("n_ids", 0), dev_counts.get("n_slots", 0)
        wire = 2 * f * 2 * nid + 2 * f * 4 * nsl + f * 4 * 0.1 * cfg["length"] * nk
        line["exchange"] = {"what": "inside the library (NCCL on the context's stream): all-gather of genome 1's candidate lists, SUM all-reduce of "
                                    "its packed rows, MIN all-reduce of the alive bytes, MAX all-reduce of the pick bytes",
                            "wire_bytes_per_rank_per_direction": int(wire), "ms_per_step": ms_x,
                            "wire_GBs_per_direction": wire / (ms_x / 1e3) / 1e9 if ms_x > 0 else None,
                            "nvlink_peak_GBs_per_direction": 770.0,
                            "frac_of_nvlink": wire / (ms_x / 1e3) / 1e9 / 770.0 if ms_x > 0 else None,
                            "note": "latency-bound: a few MB per step; no k-mer of genomes 2..G crosses NVLink (DESIGN.md section 5)"}
    if parity is not None:
        line["parity"] = parity
    if strong_c4 is not None:
        line["strong_c4"] = strong_c4
    if fasta_leg is not None:
        line["e2e_fasta"] = fasta_leg
    if not args.no_dropin_leg and world == 1:
        import tempfile
        from types import SimpleNamespace
        from primerforge_b200.candidates import _getAllCandidateKmers
        gens2 = synth.make_genomes(per_rank, cfg["length"], seed=seed, n_contigs=cfg["n_contigs"])
        with tempfile.TemporaryDirectory() as tmpd:
            paths = synth.write_fasta(gens2, tmpd)
            del gens2
            log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
            prm = SimpleNamespace(ingroupFns=paths, format="fasta", minLen=cfg["k_min"], maxLen=cfg["k_max"], minGc=40.0, maxGc=60.0,
                                  minTm=55.0, maxTm=68.0, maxRepeatLen=cfg["max_repeats"], tempTolerance=5.0, numThreads=1,
                                  p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30,
                                  log=log, pickles={1: "sharedKmers.p", 2: "candidates.p"}, dumpObj=lambda *a, **k: None,
                                  loadObj=lambda *a: None, debug=False)
            t0 = time.perf_counter()
            out = _getAllCandidateKmers(prm, False, thal="off", device=local)
            dt = time.perf_counter() - t0
            n_obj = sum(len(v) for per in out.values() for v in per.values())
        if not args.no_pairs:
            line["pairs"] = pairs_block(out, os.path.basename(paths[0]), local, flush, with_cpu=not args.no_cpu_baseline)
        line["e2e_dropin"] = {"value": distinct_bases / 1e6 / dt, "unit": UNIT, "seconds": dt, "primers": int(n_obj),
                              "what": "_getAllCandidateKmers(params, False) from FASTA files on disk to dict[str, dict[str, list[Primer]]] "
                                      "(hairpin / homodimer screening off): the call the reference's main.py:86 makes"}
    if world == 1 and not args.no_outgroup and not args.genomes and not args.length:
        line["outgroup"] = outgroup_block(local, flush, with_cpu=not args.no_cpu_baseline)
    if world == 1 and not args.no_cpu_baseline:
        # the CPU restatement on a bounded sample; the GPU must reproduce that result bit for bit (the gate)
        line["cpu_baseline"], _dt, o_cpu, gens_cpu = cpu_sample_run(args.config, length=args.length)
        line["parity"] = parity_vs_oracle(gens_cpu, o_cpu, cfg, local)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
