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
    cb, dt, o, gens = bench.cpu_sample_run(2, length=60_000)
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] > 0 and dt > 0
    assert "first 3 of the config-2 genomes at 60000 bp" in cb["sample"]
    assert len(gens) == 3 and gens[0].n_bases == 60_000 and o.n_shared > 0


def test_cpu_sample_defaults_to_the_full_genome_length():
    """the sample keeps the workload's sketch load: same genome length, fewer genomes"""
    import inspect
    src = inspect.getsource(bench.cpu_sample_run)
    assert 'length = length or cfg["length"]' in src


def test_reference_arm_line_says_what_it_ran(capsys):
    from types import SimpleNamespace
    lines = []
    bench.run_reference_arm(SimpleNamespace(config=2, warmup=0, steps=1, gpus=1, length=40_000), lines.append)
    d = lines[0]
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0
    assert d["config"]["same_workload_as_gpu_arm"] is False and "40000 bp" in d["config"]["reference_step"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
