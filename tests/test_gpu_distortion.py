"""GPU parity of the DISTORTION 1 path (reference src/laserOdometry.cpp:67,154-191,455-459,549-553,610-625) and of the map
outputs (src/laserMapping.cpp:806-836) against the CPU oracle."""
import numpy as np
import pytest

from oracle_py import Oracle

pytestmark = pytest.mark.gpu
POS_TOL, ROT_TOL = 1e-4, 1e-5


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rot_err(qa, qb):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))


def _ulp_close(got, ref, what):
    """Bit-exact up to the last-ulp differences of acos / sin between CUDA's and glibc's double libm feeding a float rounding:
    at most 1 float ulp apart, and in less than 0.1 % of the coordinates."""
    assert got.shape == ref.shape, what
    g, r = got[:, :3].astype(np.float64), ref[:, :3].astype(np.float64)
    ulp = np.spacing(np.abs(ref[:, :3]).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(g - r) <= ulp), what
    assert (_bits(got[:, :3]) != _bits(ref[:, :3])).mean() < 1e-3, what
    assert np.array_equal(_bits(got[:, 3]), _bits(ref[:, 3])), what


def test_transform_to_start_and_to_end(lvo_mod, synth):
    O = Oracle()
    pts, _ = synth.sweep(64, 0, 3, moving=True)
    full = O.extract(pts)["full"]
    T = np.array([0.004, -0.002, 0.02, 0.0, 1.0, 0.03, -0.01])
    T[3] = np.sqrt(1 - np.dot(T[:3], T[:3]))
    for mode in (0, 1):
        lvo = lvo_mod.Lvo(distortion=mode)
        got = lvo.transform_cloud(full, T)
        ref = O.transform(full, T, mode)
        if mode == 0:
            assert np.array_equal(_bits(got), _bits(ref))       # no transcendental on this path: bit-exact
        else:
            _ulp_close(got, ref, "TransformToStart")
            _ulp_close(lvo.transform_cloud(full, T, to_end=True), O.transform(full, T, 1, to_end=True), "TransformToEnd")
        lvo.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_pipeline_with_distortion(lvo_mod, synth, mode):
    """extract -> scan-to-scan -> scan-to-map on device with distortion modes 1 and 2, against the oracle frame by frame."""
    L = lvo_mod
    O = Oracle(16, 0.3, 0.2, 0.4, distortion=mode)
    lvo = L.Lvo(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4, distortion=mode, max_map_corner=1 << 18, max_map_surf=1 << 19)
    for k in range(8):
        pts, _ = synth.sweep(16, 2, k, moving=True)
        st, odo, mp = lvo.step_batch([pts])
        st_o, odo_o, mp_o = O.step(pts, keep_log=True)
        assert np.linalg.norm(odo[0][4:] - odo_o[4:]) < POS_TOL and rot_err(odo[0][:4], odo_o[:4]) < ROT_TOL, (k, odo, odo_o)
        assert np.linalg.norm(mp[0][4:] - mp_o[4:]) < POS_TOL and rot_err(mp[0][:4], mp_o[:4]) < ROT_TOL, (k, mp, mp_o)
        if k == 1:
            # identical state on both sides at outer iteration 0 of the second frame: interpolated factors produce the same LM run
            tr, lg = lvo.probe(L.P_ODO_LM_TRACE), O.odometry_log(0)
            rows = len(lg["lm"])
            assert np.array_equal(tr[0][:rows, 9], lg["lm"][:, 9])
            assert np.allclose(tr[0][:rows, :7], lg["lm"][:, :7], atol=1e-7, rtol=0)
            assert np.allclose(tr[0][:rows, 7], lg["lm"][:, 7], rtol=1e-7, atol=1e-12)
        if mode == 2:
            _ulp_close(lvo.probe(L.P_LESS_SHARP), O.odometry_last(0), f"frame {k}: less-sharp after TransformToEnd")
            _ulp_close(lvo.probe(L.P_LESS_FLAT), O.odometry_last(1), f"frame {k}: less-flat after TransformToEnd")
            _ulp_close(lvo.probe(L.P_FULL), O.odometry_last(2), f"frame {k}: full cloud after TransformToEnd")
        else:
            assert np.array_equal(_bits(lvo.probe(L.P_LESS_FLAT)), _bits(O.odometry_last(1)))
    lvo.close()


@pytest.mark.parametrize("model,cfg,seq,n", [(16, (16, 0.3, 0.2, 0.4), 1, 6), (64, (64, 5.0, 0.4, 0.8), 0, 4)])
def test_surround_and_whole_map_clouds(lvo_mod, synth, model, cfg, seq, n):
    """lvo_map_cloud vs laserMapping.cpp:806-836 restated: the same points in the same order."""
    L = lvo_mod
    O = Oracle(*cfg)
    lvo = L.Lvo(n_scans=cfg[0], minimum_range=cfg[1], line_res=cfg[2], plane_res=cfg[3], max_map_corner=1 << 18, max_map_surf=1 << 19)
    corr = np.array([0, 0, 0, 1, 0, 0, 0], float)
    for k in range(n + 1):
        pts, _ = synth.sweep(model, seq, k)
        f = O.extract(pts)
        _, _, w = O.odometry(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"], keep_log=False)
        if k == n:
            # identical pre-frame state on both sides: the oracle's map and map correction
            pc, cc = O.map_export(0)
            ps, cs = O.map_export(1)
            lvo.map_import(0, pc, cc, ps, cs)
            lvo.set_map_correction(0, corr)
            assert np.array_equal(_bits(lvo.map_cloud(0, 1)), _bits(O.map_cloud(1)))   # import -> whole-map dump: bit-exact
            st_g, pose_g, _ = lvo.scan_to_map(f["less_sharp"], f["less_flat"], f["full"], w)
        st_o, pose_o, corr = O.mapping(f["less_sharp"], f["less_flat"], f["full"], w, keep_log=False)
    assert st_g == st_o and np.linalg.norm(pose_g[4:] - pose_o[4:]) < POS_TOL
    for which in (0, 1):
        got, ref = lvo.map_cloud(0, which), O.map_cloud(which)
        # the frame's insertions use final poses that agree to ~1e-13 m, not to the last bit: same points, same order, float noise
        assert got.shape == ref.shape and len(got) > 1000, (which, got.shape, ref.shape)
        assert np.abs(got - ref).max() < 1e-3, which
        assert (_bits(got) != _bits(ref)).mean() < 0.01, which
    lvo.close()
