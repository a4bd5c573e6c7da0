"""bench.py keeps its contract: one JSON line on stdout with the keys the driver reads (small shapes here)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--genomes", "3", "--length", "300000", "--steps", "3", "--warmup", "1", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_line_has_the_contract_keys():
    d = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == 3 * 300000 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert d["e2e_fasta"]["n_picked"] == d["counts"]["n_picked"]          # the device FASTA ingest sees the same genomes
    assert "workload" in d["config"]
    # the correctness gate ran inside the harness: GPU == oracle on the cpu_baseline sample, both output modes
    assert d["parity"]["ok"] is True and len(d["parity"]["checked"]) == 2
    assert all(c["words_and_order"] and c["positions"] and c["flags_and_picks"] for c in d["parity"]["checked"])


def test_reference_arm_line():
    d = _run("--impl", "reference")
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["same_workload_as_gpu_arm"] is False and "300000 bp" in d["config"]["reference_step"]
