"""Live comparison of the hand restatement with the reference's own node sources running in oracle/_ref (skipped where
oracle/_ref/libref_*.so has not been built, i.e. where /root/reference is absent and no prebuilt copy travelled)."""
import numpy as np
import pytest

import refnode_py
from oracle_py import Oracle, Synth

pytestmark = pytest.mark.skipif(not refnode_py.available(), reason="oracle/_ref/libref_*.so not built")
FEATS = ("full", "sharp", "less_sharp", "flat", "less_flat")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_glibc_atan_variant(synth):
    """The node exactly as the reference builds it (glibc atan / atan2): same points, same order, same ring ids and picks as the
    restatement; only the fractional (relTime) part of ~0.1 % of the intensities moves by one or two float ulps — DESIGN.md deviation 3."""
    R = refnode_py.RefPipeline(64, 5.0, 0.4, 0.8)
    O = Oracle(64, 5.0, 0.4, 0.8)
    for k in range(2):
        pts, _ = synth.sweep(64, 3, k)
        feats = R.registration(pts)
        f = O.extract(pts)
        for name, a in zip(FEATS, feats):
            b = f[name]
            assert a.shape == b.shape, (k, name)
            assert np.array_equal(_bits(a[:, :3]), _bits(b[:, :3])), (k, name)
            assert np.array_equal(a[:, 3].astype(int), b[:, 3].astype(int)), (k, name)
            assert np.abs(a[:, 3] - b[:, 3]).max() <= 4e-6 and (_bits(a[:, 3]) != _bits(b[:, 3])).mean() < 0.01, (k, name)
    R.close()


def test_nan_points_with_is_dense_false(synth):
    """removeNaNFromPointCloud (scanRegistration.cpp:136) filters only when the message is not flagged dense."""
    R = refnode_py.RefPipeline(64, 5.0, 0.4, 0.8, lvo_atan=True)
    O = Oracle(64, 5.0, 0.4, 0.8)
    pts, _ = synth.sweep(64, 1, 0)
    pts = pts.copy()
    pts[::97, 0] = np.nan
    pts[5::131, 2] = np.inf
    feats = R.registration(pts, dense=False)
    f = O.extract(pts)
    for name, a in zip(FEATS, feats):
        assert a.shape == f[name].shape and np.array_equal(_bits(a), _bits(f[name])), name
    R.close()


def test_long_run_with_detail(synth):
    """12 HDL-64 frames of a sequence not used by the fixtures: features / labels / map cubes bit-exact, LM control flow equal, poses to 1e-11."""
    R = refnode_py.RefPipeline(64, 5.0, 0.4, 0.8, lvo_atan=True)
    O = Oracle(64, 5.0, 0.4, 0.8)
    for k in range(12):
        pts, _ = synth.sweep(64, 5, k)
        r = R.step(pts, detail=(k % 4 == 3))
        st, odo, mp = O.step(pts, keep_log=True)
        assert np.abs(r["odom"] - odo).max() < 1e-11 and np.abs(r["map"] - mp).max() < 1e-11, k
        if "feats" in r:
            f = O.extract(pts)
            for name, a in zip(FEATS, r["feats"]):
                assert a.shape == f[name].shape and np.array_equal(_bits(a), _bits(f[name])), (k, name)
            n = len(f["full"])
            for name in ("curvature", "picked", "label"):
                assert np.array_equal(r[name][5:n - 5], f[name][5:n - 5]), (k, name)
            # cloudSortInd: std::sort (scanRegistration.cpp:288) leaves the order of EQUAL curvatures unspecified; the restatement and the
            # CUDA kernels fix it to (curvature, index) ascending (DESIGN.md deviation 1).  Same curvature sequence, and the positions
            # that differ are ties.
            a, b, cv = r["sort_ind"][5:n - 5], f["sort_ind"][5:n - 5], f["curvature"]
            assert np.array_equal(cv[a], cv[b]), k
            assert (a != b).mean() < 1e-3
            for which, nm in ((0, "corner"), (1, "surf")):
                p, c = O.map_export(which)
                assert np.array_equal(c, r[f"map_{nm}_cube"]) and np.array_equal(_bits(p), _bits(r[f"map_{nm}"])), (k, nm)
            for o in range(10):
                assert O.mapping_log(o)["lm"][:, 9].astype(int).tolist() == r["lm_map"][o][1:, 9].astype(int).tolist(), (k, o)
                assert O.odometry_log(o)["lm"][:, 9].astype(int).tolist() == r["lm_odo"][o][1:, 9].astype(int).tolist(), (k, o)
    R.close()
