"""bench.py pieces that do not need a GPU: workload descriptions and the bounded CPU sample."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_workload_config_weak_and_strong():
    w1 = bench.workload_config(2, 1)
    assert w1["genomes_per_gpu"] == 10 and w1["distinct_genomes"] == 10 and w1["parallelism"] == "single GPU"
    w8 = bench.workload_config(2, 8)
    assert w8["genomes_per_gpu"] == 10 and w8["distinct_genomes"] == 1 + 9 * 8          # genome 0 is on every rank
    s8 = bench.workload_config(4, 8, strong=True)
    assert s8["distinct_genomes"] == 200 and s8["genomes_per_gpu"] == 1 + 25           # the largest shard
    assert "200 synthetic" in s8["workload"] and s8["k_min"] == 16 and s8["k_max"] == 20
    c5 = bench.workload_config(5, 1)
    assert c5["k_min"] == 12 and c5["k_max"] == 30 and "50 contig" in c5["workload"]


def test_cpu_sample_uses_all_threads_whatever_omp_says(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")          # what torchrun does to every rank
    cb, dt = bench.cpu_sample_run(2, length=60_000)
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] > 0 and dt > 0
    assert "3 x 60000 bp" in cb["sample"]
