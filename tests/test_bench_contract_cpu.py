"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys (it times the CPU
oracle, so it runs anywhere), non-zero ranks of a torchrun launch print nothing, and our arm refuses to run without a GPU
instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--ref-threads", "2", "--ref-impl", "port", "--frames-per-step", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("scan-to-map scans/sec") and d["unit"] == "scans/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 2 * 2 * 2 / (d["ms_per_step"] * 2e-3)) < 1e-6 * d["value"]   # threads x steps x frames per step / time
    assert d["config"]["frames_per_step"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_runs_the_reference_node_sources_when_built():
    """With oracle/_ref/libref_*.so present the arm times the reference's OWN scanRegistration / laserOdometry / laserMapping sources
    (kind "reference"); without them it falls back to the restated port and says so."""
    import refnode_py
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--ref-threads", "2", "--frames-per-step", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == ("reference" if refnode_py.available() else "port") and d["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout.count("GPU ") > 0
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_our_arm_has_no_cpu_fallback():
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
