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


@pytest.mark.skipif(not os.path.exists(EXE), reason="host/kitti_pipeline not built")
def test_cpp_driver_distortion_2_and_map_dump(lvo_mod, synth, tmp_path):
    """--distortion 2 through the three separate calls + LaserOdometry::publishedClouds must equal the fused device pipeline
    (lvo_step_batch) with the same option; --map-out is the whole-map cloud of laserMapping.cpp:823-836 after frame 0."""
    L = lvo_mod
    files, sweeps = [], []
    for k in range(3):
        pts, _ = synth.sweep(16, 2, k, moving=True)
        p = tmp_path / f"{k:06d}.bin"
        pts.astype(np.float32).tofile(p)
        files.append(str(p)); sweeps.append(pts)
    out, mp = tmp_path / "poses.txt", tmp_path / "map.bin"
    r = subprocess.run([EXE, "16", str(out), "--distortion", "2", "--map-out", str(mp)] + files, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    qt = np.array([[float(x) for x in l.split()[2:6] + l.split()[7:10]] for l in open(out).read().splitlines() if l.startswith("# q")])
    lvo = L.Lvo(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4, distortion=2, max_map_corner=1 << 18, max_map_surf=1 << 19)
    for k in range(3):
        _, _, pm = lvo.step_batch([L.to_pcl_layout(sweeps[k])])
        assert np.array_equal(pm[0], qt[k]), (k, pm[0], qt[k])
        if k == 0:
            whole = lvo.map_cloud(0, 1)
            dumped = np.fromfile(mp, np.float32).reshape(-1, 4)
            assert len(whole) > 100 and np.array_equal(whole.view(np.uint32), dumped.view(np.uint32))
    lvo.close()
