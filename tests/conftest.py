import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_lvo():
    """Import the package whose directory name contains a hyphen."""
    if "lvo_b200" in sys.modules:
        return sys.modules["lvo_b200"]
    spec = importlib.util.spec_from_file_location("lvo_b200", os.path.join(ROOT, "lidar-visual-odometry_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lvo_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def lvo_mod():
    return load_lvo()


@pytest.fixture(scope="session")
def synth():
    from oracle_py import Synth
    return Synth()


def record_metric(name, value):
    """Measured quantities of the parity tests (e.g. how many kNN rows differ from the oracle after the poses have drifted apart by
    rounding) go to gpurun_out/test_metrics.json when that directory exists, so that they can be quoted in profiles/."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "test_metrics.json")
    try:
        cur = json.load(open(p)) if os.path.exists(p) else {}
    except Exception:
        cur = {}
    cur[name] = value
    json.dump(cur, open(p, "w"), indent=1)
