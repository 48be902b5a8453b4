"""bench.py contract on the CPU side: the reference arm (oracle restatement on the host cores) prints one JSON line with
the keys the driver reads, for the NequIP headline workload and for the SEGNN / EGNN workloads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("kind,nodes", [("egnn", 300), ("segnn", 120)])
def test_family_cpu_baseline_runs_on_a_small_cloud(kind, nodes):
    sys.path.insert(0, ROOT)
    import bench
    val, ms, threads, sample = bench.cpu_family_baseline(kind, nodes, 1, warm=0)
    assert val > 0 and ms > 0 and threads >= 1 and f"{nodes} nodes" in sample


def test_reference_arm_line_for_the_egnn_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "egnn",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]
