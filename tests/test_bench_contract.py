"""bench.py on CPU: the roofline's bytes-per-segment accounting and the JSON line of the reference arm
(`--impl reference`: the oracle on the host cores, the one leg of the bench that runs without a GPU)."""
import json
import os
import subprocess
import sys

import pytest

import __graft_entry__ as entry

sys.path.insert(0, entry.ROOT)
import bench  # noqa: E402


def test_bytes_per_segment_follows_the_survey():
    # the survey's examples leave the 64-bit particle id out; bench.py carries it (+16 bytes per history)
    assert bench.bytes_per_segment(2, 4, 10.0) == pytest.approx(28 + (66 + 16) / 10)       # F32 2-D, 10 segments per history
    assert bench.bytes_per_segment(2, 8, 10.0) == pytest.approx(56 + (114 + 16) / 10)      # F64 2-D
    assert bench.bytes_per_segment(1, 4, 1.2) == pytest.approx(24 + (50 + 16) / 1.2)       # Su-Olson F32
    assert bench.bytes_per_segment(2, 4, 98.8) == pytest.approx(28.83, abs=0.01)           # the default bench workload


def test_reference_arm_prints_the_contract_line(built):
    r = subprocess.run([sys.executable, os.path.join(entry.ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1",
                        "--cpu-sample", "150000"], capture_output=True, text=True, timeout=600, cwd=entry.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "tracked particle-segments/sec" and d["unit"] == "segments/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    # the arm describes what IT ran (a bounded sample), not the GPU arm's mesh and population
    assert d["config"]["mesh"] != d["config"]["gpu_arm_mesh"] == [4096, 4096] and d["config"]["nmax_global"] < d["config"]["gpu_arm_nmax_global"]
    assert "bounded CPU sample" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"] and cb["single_core_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e5
