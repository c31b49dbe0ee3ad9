"""CPU tests (no GPU) of the DISTORTION 1 restatement (reference src/laserOdometry.cpp:67,154-191,455-459,549-553;
src/lidarFactor.hpp:27-30,79-82) and of the map outputs (src/laserMapping.cpp:806-836)."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle_py import ROOT, Oracle, Synth

REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblvo_oracle_ref.so")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def qmul(a, b):
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def qrot(q, v):
    u = q[:3]
    uv = 2 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


def _random_case(rng, it):
    f = np.zeros(14)
    f[0] = it % 2
    f[1:13] = rng.normal(size=12) * 10
    ang = rng.uniform(0, 0.3) if it % 3 else rng.uniform(0, 3.0)
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    q = np.r_[np.sin(ang / 2) * ax, np.cos(ang / 2)]
    if it % 7 == 0:
        q = -q                      # d < 0 branch of slerp
    if it % 11 == 0:
        q = np.array([0, 0, 0, 1.0])  # absD >= 1 - eps branch (linear weights)
    x = np.r_[q, rng.normal(size=3)]
    s = rng.uniform(0, 1) if it % 5 else 1.0
    return f, s, x


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_interpolated_factors_match_reference_functors_autodiff():
    """LidarEdgeFactor / LidarPlaneFactor with s != 1: analytic residual + local Jacobian vs forward-mode autodiff of the
    reference's own functors (lidarFactor.hpp compiled unmodified) through EigenQuaternionParameterization."""
    A, R = Oracle(), Oracle(reference_build=True)
    rng = np.random.default_rng(0)
    worst = 0.0
    for it in range(3000):
        f, s, x = _random_case(rng, it)
        r1, J1 = A.eval_factor_s(f, s, x)
        r2, J2 = R.eval_factor_s(f, s, x)
        assert r1.shape == r2.shape
        worst = max(worst, np.abs(r1 - r2).max() / max(1.0, np.abs(r2).max()), np.abs(J1 - J2).max() / max(1.0, np.abs(J2).max()))
    assert worst < 1e-12, worst


def test_interpolated_jacobian_by_finite_differences():
    """Independent of the reference build: central differences through EigenQuaternionParameterization::Plus."""
    A = Oracle()
    rng = np.random.default_rng(1)
    for it in range(200):
        f, s, x = _random_case(rng, it)
        f[1:13] /= 5.0
        if it % 11 == 0:
            continue  # derivative of the linear-weights branch is discontinuous at the switch; covered by the autodiff test
        r0, J = A.eval_factor_s(f, s, x)
        h = 1e-6
        for c in range(6):
            d = np.zeros(6)
            d[c] = h
            xs = []
            for sg in (+1, -1):
                dd = sg * d
                nd = np.linalg.norm(dd[:3])
                dq = np.r_[np.sin(nd) / nd * dd[:3], np.cos(nd)] if nd > 0 else np.array([0, 0, 0, 1.0])
                xs.append(np.r_[qmul(dq, x[:4]), x[4:] + dd[3:]])
            rp, _ = A.eval_factor_s(f, s, xs[0])
            rm, _ = A.eval_factor_s(f, s, xs[1])
            fd = (rp - rm) / (2 * h)
            assert np.allclose(fd, J[:, c], atol=2e-5 * max(1.0, np.abs(J).max())), (it, c, fd, J[:, c])


def test_s_equal_one_is_the_undistorted_factor():
    A = Oracle()
    rng = np.random.default_rng(2)
    for it in range(50):
        f, _, x = _random_case(rng, it)
        r1, J1 = A.eval_factor_s(f, 1.0, x)
        r0, J0 = A.eval_factor(f, x)
        assert np.array_equal(r1, r0) and np.array_equal(J1, J0)


def test_transform_to_start_matches_numpy_slerp():
    """TransformToStart with DISTORTION 1 (:154-172) vs a numpy statement of slerp + rotation in double."""
    A = Oracle()
    rng = np.random.default_rng(3)
    n = 2000
    pts = np.c_[rng.normal(0, 20, (n, 3)), rng.integers(0, 64, n) + rng.uniform(0, 0.0999, n)].astype(np.float32)
    ang = 0.05
    ax = np.array([0.2, -0.1, 0.97])
    ax /= np.linalg.norm(ax)
    T = np.r_[np.sin(ang / 2) * ax, np.cos(ang / 2), 1.0, 0.05, -0.02]
    got = A.transform(pts, T, 1)
    theta = np.arccos(abs(T[3]))
    ref = np.empty((n, 3))
    for i in range(n):
        s = float(np.float32(pts[i, 3] - np.float32(int(pts[i, 3])))) / 0.1
        k0, k1 = np.sin((1 - s) * theta) / np.sin(theta), np.sin(s * theta) / np.sin(theta)
        qs = np.r_[k1 * T[:3], k0 + k1 * T[3]]
        ref[i] = qrot(qs, pts[i, :3].astype(np.float64)) + s * T[4:]
    assert np.abs(got[:, :3] - ref).max() < 4e-6            # float rounding of the output at |p| ~ 60 m
    assert np.array_equal(_bits(got[:, 3]), _bits(pts[:, 3]))
    # mode 0 is the plain rigid transform, intensity untouched
    got0 = A.transform(pts, T, 0)
    ref0 = np.array([qrot(T[:4], p[:3].astype(np.float64)) + T[4:] for p in pts])
    assert np.abs(got0[:, :3] - ref0).max() < 4e-6
    # TransformToEnd strips the fractional part of the intensity (:190) and is the inverse rigid motion of the undistorted point
    end = A.transform(pts, T, 1, to_end=True)
    assert np.array_equal(end[:, 3], np.floor(pts[:, 3]))
    qi = np.r_[-T[:3], T[3]]
    ref_end = np.array([qrot(qi, got[i, :3].astype(np.float64) - T[4:]) for i in range(n)])
    assert np.abs(end[:, :3] - ref_end).max() < 4e-6


def test_undistortion_brings_a_moving_sweep_onto_the_rigid_one():
    """Generator + TransformToEnd geometry: a sweep cast from a moving sensor, undistorted with the true motion, lies closer to
    the rigid sweep cast at the end pose than the raw sweep does."""
    s, o = Synth(), Oracle()
    k = 20
    mv, _ = s.sweep(64, 0, k, moving=True)
    _, g0 = s.sweep(64, 0, k)
    r1, g1 = s.sweep(64, 0, k + 1)
    qi = np.r_[-g0[:3], g0[3]]
    T = np.r_[qmul(qi, g1[:4]), qrot(qi, g1[4:] - g0[4:])]
    full = o.extract(mv)["full"][::5]
    end = o.transform(full, T, 1, to_end=True)
    tree = cKDTree(r1[:, :3])
    d_raw, _ = tree.query(full[:, :3])
    d_end, _ = tree.query(end[:, :3])
    assert np.median(d_end) < 0.75 * np.median(d_raw), (np.median(d_end), np.median(d_raw))


@pytest.mark.parametrize("mode", [1, 2])
def test_pipeline_with_distortion_tracks_ground_truth(mode):
    s = Synth()
    o = Oracle(16, 0.3, 0.2, 0.4, distortion=mode)
    for k in range(12):
        pts, gt = s.sweep(16, 2, k, moving=True)
        st, odo, mp = o.step(pts)
        assert np.all(np.isfinite(odo)) and np.all(np.isfinite(mp))
    assert np.linalg.norm(mp[4:] - gt[4:]) < 0.25, (mp, gt)
    last = o.odometry_last(0)
    frac = last[:, 3] - np.floor(last[:, 3])
    if mode == 2:
        assert np.all(frac == 0)     # :190 "Remove distortion time info"
    else:
        assert frac.max() > 0.05     # the block :610-625 is disabled: relTime survives


def test_distortion_zero_is_the_default_pipeline():
    s = Synth()
    a, b = Oracle(16, 0.3, 0.2, 0.4), Oracle(16, 0.3, 0.2, 0.4, distortion=0)
    for k in range(4):
        pts, _ = s.sweep(16, 1, k)
        ra, rb = a.step(pts), b.step(pts)
        assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])


def test_surround_and_whole_map_clouds():
    """laserMapping.cpp:806-836: surround = valid cubes in loop order, corner then surf of each cube; whole map = all cubes."""
    s = Synth()
    o = Oracle(16, 0.3, 0.2, 0.4)
    for k in range(5):
        o.step(s.sweep(16, 1, k)[0])
    pc, cc = o.map_export(0)
    ps, cs = o.map_export(1)
    info = o.mapping_info()["info"]
    center = info[:3]
    want = []
    for i in range(center[0] - 2, center[0] + 3):
        for j in range(center[1] - 2, center[1] + 3):
            for kk in range(center[2] - 1, center[2] + 2):
                ind = i + 21 * j + 441 * kk
                want.append(pc[cc == ind])
                want.append(ps[cs == ind])
    want = np.concatenate(want)
    got = o.map_cloud(0)
    assert got.shape == want.shape and np.array_equal(_bits(got), _bits(want))
    whole = o.map_cloud(1)
    assert len(whole) == len(pc) + len(ps) == info[6] + info[7]
    cubes = np.unique(np.r_[cc, cs])
    want_all = np.concatenate([np.r_[pc[cc == c], ps[cs == c]] for c in cubes])
    assert np.array_equal(_bits(whole), _bits(want_all))
