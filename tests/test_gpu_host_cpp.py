"""The C++ host adaptor (host/lvo_handlers.hpp) driven by host/kitti_pipeline.cpp on KITTI-format .bin sweeps must give
bit-identical poses to the same three calls made through the Python binding (same library, deterministic kernels)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "lidar-visual-odometry_b200", "host", "kitti_pipeline")


@pytest.mark.skipif(not os.path.exists(EXE), reason="host/kitti_pipeline not built")
def test_cpp_driver_matches_python_binding(lvo_mod, synth, tmp_path):
    L = lvo_mod
    files, sweeps = [], []
    for k in range(4):
        pts, _ = synth.sweep(64, 0, k)
        pts[:, 3] = 0.25  # reflectance column of the KITTI record; ignored by the path (scanRegistration.cpp:132-133 drops it)
        p = tmp_path / f"{k:06d}.bin"
        pts.astype(np.float32).tofile(p)
        files.append(str(p)); sweeps.append(pts)
    out = tmp_path / "poses.txt"
    r = subprocess.run([EXE, "64", str(out)] + files, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = open(out).read().splitlines()
    qt = np.array([[float(x) for x in l.split()[2:6] + l.split()[7:10]] for l in lines if l.startswith("# q")])
    kitti = np.array([[float(x) for x in l.split()] for l in lines if not l.startswith("#")])
    assert qt.shape == (4, 7) and kitti.shape == (4, 12)
    lvo = L.Lvo(max_map_corner=1 << 18, max_map_surf=1 << 19)
    for k in range(4):
        _, f = lvo.extract_features(L.to_pcl_layout(sweeps[k]))
        _, rel, w = lvo.scan_to_scan(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        _, pose, _ = lvo.scan_to_map(f["less_sharp"], f["less_flat"], f["full"], w)
        assert np.array_equal(pose, qt[k]), (k, pose, qt[k])
        assert np.allclose(kitti[k][[3, 7, 11]], pose[4:], atol=0, rtol=0)
    lvo.close()
