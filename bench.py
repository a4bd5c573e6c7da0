#!/usr/bin/env python
"""bench.py -- candidate-k-mer stage throughput (ingroup Mbp/s) on N B200s, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl reference]

A step = one full pass of the stage (2-bit encode, genome-1 sketch, key table, scan of every other
genome, uniqueness rules, ordered emit, filters, grouping) over one batch of synthetic genomes of
BASELINE.json's config (default: configs[1] = 10 x 5 Mbp, primer_len 16-20, default GC/Tm).

  value : device-resident ASCII bases -> device-resident records, CUDA events on the library's
          stream, L2 flushed between steps, max over ranks.
  e2e   : the same stage through the C ABI with HOST (pinned) buffers: H2D of the bases, compute,
          D2H of the records and of the picked candidates' coordinates in every genome.
  N > 1 : launched by torchrun, one rank per GPU; every rank holds genome 0 (it defines the key
          table) plus its own shard of the other genomes; the only exchanges are a MIN all-reduce of
          the per-key alive bytes and a MAX all-reduce of the pick bytes (NCCL).  Weak scaling:
          1 + 9 genomes per rank, throughput counts DISTINCT ingroup bases.
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
    ap.add_argument("--scan-mode", type=int, default=0, help="0 auto, 1 plain L2 probe, 2 shared-memory filter (debug)")
    ap.add_argument("--dropin-leg", action="store_true",
                    help="also time the drop-in entry point _getAllCandidateKmers(params, False) from FASTA files on disk "
                         "to the dict of Primer objects (SURVEY 8d T_e2e; dominated by building G x |S| Python objects)")
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
    from primerforge_b200 import _lib, synth
    from primerforge_b200.dist import allreduce_stage_buffers, run_sharded
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
    if strong:
        from primerforge_b200.dist import shard_genomes
        n_distinct = cfg["n_genomes"]
        idx = shard_genomes(n_distinct, world, rank)
        per_rank = len(shard_genomes(n_distinct, world, 0))          # the largest shard
        gens = synth.make_genomes(max(idx) + 1, cfg["length"], seed=synth.BASE_SEED + args.config, n_contigs=cfg["n_contigs"])
        mine = [gens[i] for i in idx]
    else:
        per_rank = cfg["n_genomes"]
        n_distinct = per_rank if world == 1 else 1 + (per_rank - 1) * world
        # the generator is sequential in the genome index: every rank generates the prefix it needs
        last_needed = per_rank if world == 1 else 1 + (per_rank - 1) * (rank + 1)
        gens = synth.make_genomes(last_needed, cfg["length"], seed=synth.BASE_SEED + args.config, n_contigs=cfg["n_contigs"])
        mine = gens[:per_rank] if world == 1 else [gens[0]] + gens[1 + (per_rank - 1) * rank: 1 + (per_rank - 1) * (rank + 1)]
    del gens
    distinct_bases = n_distinct * cfg["length"]

    # pinned host buffers (the e2e leg copies from these every step)
    pinned = []
    for g in mine:
        cat = np.concatenate(g.contigs) if len(g.contigs) > 1 else g.contigs[0]
        t = torch.empty(cat.size, dtype=torch.uint8).pin_memory()
        t.numpy()[:] = cat
        off = np.zeros(len(g.contigs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([c.size for c in g.contigs])
        pinned.append((t, off))
    del mine

    eng = Engine(device=local, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
    eng.set_scan_mode(args.scan_mode)
    for t, off in pinned:
        eng.add_genome_ptr(t.data_ptr(), t.numel(), off)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    xchg = {"ms": [], "bytes": 0}
    if world > 1:
        ext = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local))
        xev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def compute_once():
        if world == 1:
            eng.compute()
        else:
            # the two reductions are enqueued on the library's stream between the stages (no host waits);
            # CUDA events on that stream time them
            eng.compute_stage(0)
            xev[0].record(ext)
            allreduce_stage_buffers(eng, _lib.PFB_BUF_ALIVE, "min")
            xev[1].record(ext)
            eng.compute_stage(1)
            xev[2].record(ext)
            allreduce_stage_buffers(eng, _lib.PFB_BUF_PICK, "max")
            xev[3].record(ext)
            eng.compute_stage(2)
            xchg["ms"].append(xev[0].elapsed_time(xev[1]) + xev[2].elapsed_time(xev[3]))
            xchg["bytes"] = eng.device_buffer(_lib.PFB_BUF_ALIVE)[1] + eng.device_buffer(_lib.PFB_BUF_PICK)[1]

    def fetch_results():
        rec, ss, ct = eng.result_picked()     # pinned host buffers, plain D2H copies
        return rec, [ss] + ([ct] if ct is not None else [])

    # ---- device-resident leg
    eng.upload()
    for _ in range(args.warmup):
        flush.zero_()
        barrier()
        compute_once()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.35)        # nvidia-smi needs a moment to start printing; the loop below keeps the GPU busy
    dev_ms, launches, scan_ms, scan_main_ms, enc_ms, first_ms, fin_ms = [], 0, [], [], [], [], []
    dev_counts = {}
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        compute_once()
        tm = eng.timing()
        dev_ms.append(tm["ms_total"])
        scan_ms.append(tm["ms_scan"])
        scan_main_ms.append(tm["ms_scan_main"])
        enc_ms.append(tm["ms_encode"]); first_ms.append(tm["ms_first"]); fin_ms.append(tm["ms_finalize"])
        dev_counts = {k: int(tm[k]) for k in ("n_bases", "n_windows", "n_slots", "n_records", "n_picked")}
        launches += tm["n_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    tm = eng.timing()

    # ---- end-to-end leg (host buffers -> host results through the C ABI)
    e2e_s, e2e_parts = [], []
    h2d = sum(t.numel() for t, _ in pinned)
    d2h = 0
    for i in range(max(1, args.warmup // 2) + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            eng.run()                 # pfb_run: H2D copies overlapped with the key-table build and the scan
        else:
            run_sharded(eng, upload=True)
        t1 = time.perf_counter()
        rec, pos = fetch_results()
        eng.sync()
        dt = time.perf_counter() - t0
        if i >= max(1, args.warmup // 2):
            e2e_s.append(dt)
            e2e_parts.append((t1 - t0, dt - (t1 - t0), eng.timing()["ms_total"]))
            d2h = rec.nbytes + sum(p.nbytes for p in pos)
    clocks = sampler.stop() if rank == 0 else None

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
        eng_f = Engine(device=local, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"], prefilter=1)
        for pt in images:
            eng_f.add_fasta(ptr=pt.data_ptr(), n_bytes=pt.numel())
        fa_s = []
        for i in range(2 + min(args.steps, 50)):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            eng_f.run()
            rec_f, ss_f, ct_f = eng_f.result_picked()
            eng_f.sync()
            if i >= 2:
                fa_s.append(time.perf_counter() - t0)
        fa_ms = float(np.mean(fa_s) * 1e3)
        fasta_leg = {"value": distinct_bases / 1e6 / (fa_ms / 1e3), "unit": UNIT, "ms_per_step": fa_ms,
                     "h2d_bytes_per_step": int(sum(pt.numel() for pt in images)),
                     "d2h_bytes_per_step": int(rec_f.nbytes + ss_f.nbytes + (ct_f.nbytes if ct_f is not None else 0)),
                     "n_picked": int(rec_f.size), "input": "80-column FASTA file images in pinned host memory, parsed on the device"}
        eng_f.close()

    step_ms = float(np.mean(dev_ms))
    e2e_ms = float(np.mean(e2e_s) * 1e3)
    if world > 1:
        red = torch.tensor([step_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = red.tolist()
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
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_scan_traffic.json")) as fh:
            tj = json.load(fh)
        if args.config == 2 and world == 1 and not args.genomes and not args.length:
            traffic, traffic_src = tj["dram_bytes_total"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_scan_filtered (genomes 2..G, one launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_base": bpb,
                "bytes_per_base_source": "SURVEY.md 8(d) stage model (count-everything design); see DESIGN.md section 4",
                "ms_per_launch": ms_scan, "share_of_step": ms_scan / step_ms if step_ms else None,
                "actual_limiter": {"kind": "INT32 ALU pipe issue (hash + filter + queue append per window); DRAM traffic is ~6% of the modelled bytes",
                                   "windows_per_s_G": windows / (ms_scan / 1e3) / 1e9 if ms_scan > 0 else 0.0,
                                   "evidence": "profiles/r01_scan_ablation.md, profiles/r01_v14_ncu_summary.json"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(args.config, world, strong) | ({"genomes_per_gpu": per_rank, "genome_bp": cfg["length"]}),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "ms_run_call": float(np.mean([a for a, _, _ in e2e_parts]) * 1e3), "ms_fetch_results": float(np.mean([b for _, b, _ in e2e_parts]) * 1e3),
                    "ms_device_span_in_run": float(np.mean([c for _, _, c in e2e_parts]))},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "host_cpus_bound": affinity,
            "phases_ms": {"ms_encode": float(np.mean(enc_ms)), "ms_first": float(np.mean(first_ms)), "ms_scan": float(np.mean(scan_ms)),
                          "ms_finalize": float(np.mean(fin_ms)), "ms_total": step_ms},
            "counts": dev_counts,
            "wall_s_timed_loop": wall}
    if world > 1 and xchg["ms"]:
        # ring all-reduce: every rank sends and receives 2 (N - 1) / N of the buffer per direction
        ms_x = float(np.mean(xchg["ms"][-args.steps:]))
        wire = 2.0 * (world - 1) / world * xchg["bytes"]
        line["exchange"] = {"what": "MIN all-reduce of the per-key alive bytes + MAX all-reduce of the pick bytes (NCCL, on the library's stream)",
                            "bytes_reduced_per_rank": int(xchg["bytes"]), "ms_per_step": ms_x,
                            "wire_GBs_per_direction": wire / (ms_x / 1e3) / 1e9 if ms_x > 0 else None,
                            "nvlink_peak_GBs_per_direction": 770.0,
                            "frac_of_nvlink": wire / (ms_x / 1e3) / 1e9 / 770.0 if ms_x > 0 else None,
                            "note": "latency-bound: ~2 MB per step; no k-mer crosses NVLink (DESIGN.md section 5)"}
    if fasta_leg is not None:
        line["e2e_fasta"] = fasta_leg
    if args.dropin_leg and world == 1:
        import tempfile
        from types import SimpleNamespace
        from primerforge_b200.candidates import _getAllCandidateKmers
        gens2 = synth.make_genomes(per_rank, cfg["length"], seed=synth.BASE_SEED + args.config, n_contigs=cfg["n_contigs"])
        with tempfile.TemporaryDirectory() as tmpd:
            paths = synth.write_fasta(gens2, tmpd)
            del gens2
            log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
            prm = SimpleNamespace(ingroupFns=paths, format="fasta", minLen=cfg["k_min"], maxLen=cfg["k_max"], minGc=40.0, maxGc=60.0,
                                  minTm=55.0, maxTm=68.0, maxRepeatLen=cfg["max_repeats"], tempTolerance=5.0, numThreads=1,
                                  p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30,
                                  log=log, pickles={1: "sharedKmers.p", 2: "candidates.p"}, dumpObj=lambda *a, **k: None,
                                  loadObj=lambda *a: None)
            t0 = time.perf_counter()
            out = _getAllCandidateKmers(prm, False, thal="off", device=local)
            dt = time.perf_counter() - t0
        n_obj = sum(len(v) for per in out.values() for v in per.values())
        line["e2e_dropin"] = {"value": distinct_bases / 1e6 / dt, "unit": UNIT, "seconds": dt, "primer_objects": int(n_obj),
                              "what": "_getAllCandidateKmers(params, False) from FASTA files on disk to dict[str, dict[str, list[Primer]]], "
                                      "hairpin / homodimer screening off; the time is building the Python Primer objects"}
    if world == 1 and not args.no_cpu_baseline:
        # the CPU restatement on a bounded sample; the GPU must reproduce that result bit for bit (the gate)
        line["cpu_baseline"], _dt, o_cpu, gens_cpu = cpu_sample_run(args.config, length=args.length)
        line["parity"] = parity_vs_oracle(gens_cpu, o_cpu, cfg, local)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
