"""GPU parity of the camera-lidar depth association (BASELINE config 5; reference src/vloam/Frame.cpp:289-352,
src/vloam/Frontend.cpp:223-301): depth cloud, 3-NN index sets, validity flags and interpolated depths bit-exact."""
import numpy as np
import pytest

from oracle_py import Oracle

pytestmark = pytest.mark.gpu


def keypoints(cam, rng, per_cell=5):
    """5 uniform-random pixels in each of 28 x 8 sub-regions inside borders 20 / 15 (params/KITTI00.yaml:62-66), normalised as
    featureTracking.cpp:181-182 does."""
    xs = np.linspace(20, cam["width"] - 20, 29)
    ys = np.linspace(15, cam["height"] - 15, 9)
    px = []
    for i in range(28):
        for j in range(8):
            px.append(np.stack([rng.uniform(xs[i], xs[i + 1], per_cell), rng.uniform(ys[j], ys[j + 1], per_cell)], 1))
    px = np.concatenate(px)
    return np.stack([(px[:, 0] - cam["cx"]) / cam["fx"], (px[:, 1] - cam["cy"]) / cam["fy"]], 1).astype(np.float32)


@pytest.mark.parametrize("seq,frame", [(0, 0), (2, 4)])
def test_depth_association_bit_exact(lvo_mod, synth, seq, frame):
    L = lvo_mod
    cam = L.KITTI00_CAMERA
    lvo = L.Lvo(max_map_corner=1 << 17, max_map_surf=1 << 17)
    O = Oracle()
    pts, _ = synth.sweep(64, seq, frame)
    uv = keypoints(cam, np.random.default_rng(seq * 10 + frame))
    assert len(uv) == 1120
    dc_o, src_o, d_o, v_o, nn_o = O.depth(pts, cam["extrinsic"], uv)
    dc_g, d_g, v_g, nn_g = lvo.depth_associate(pts, uv)
    assert dc_g.shape == dc_o.shape and np.array_equal(dc_g.view(np.uint32), dc_o.view(np.uint32)), "depth cloud"
    assert np.array_equal(nn_g, nn_o), f"{(nn_g != nn_o).any(axis=1).sum()} 3-NN rows differ"
    assert np.array_equal(v_g, v_o)
    assert np.array_equal(d_g.view(np.uint32), d_o.view(np.uint32))
    assert 0.3 < v_o.mean() < 0.95  # not vacuous: some keypoints get a depth, some are rejected
    lvo.close()


def test_depth_association_edges(lvo_mod, synth):
    L = lvo_mod
    cam = L.KITTI00_CAMERA
    lvo = L.Lvo(max_map_corner=1 << 17, max_map_surf=1 << 17)
    O = Oracle()
    pts, _ = synth.sweep(64, 1, 2)
    # keypoints far outside the image, on the border, duplicated
    uv = np.array([[5.0, 5.0], [-3.0, 0.1], [0.0, 0.0], [0.0, 0.0], [0.84, -0.24], [-0.84, 0.24]], np.float32)
    dc_o, _, d_o, v_o, nn_o = O.depth(pts, cam["extrinsic"], uv)
    dc_g, d_g, v_g, nn_g = lvo.depth_associate(pts, uv)
    assert np.array_equal(nn_g, nn_o) and np.array_equal(v_g, v_o) and np.array_equal(d_g.view(np.uint32), d_o.view(np.uint32))
    # every point behind the camera -> empty depth cloud, nothing valid
    behind = pts[pts[:, 0] < -1.0]
    dc_g, d_g, v_g, nn_g = lvo.depth_associate(behind, uv)
    assert len(dc_g) == 0 and not v_g.any() and (nn_g == -1).all()
    # fewer than three depth points
    few = pts[pts[:, 0] > 5.0][:2]
    dc_g, d_g, v_g, nn_g = lvo.depth_associate(few, uv)
    assert len(dc_g) == 2 and not v_g.any()
    # no keypoints
    dc_g, d_g, v_g, nn_g = lvo.depth_associate(pts, np.zeros((0, 2), np.float32))
    assert len(dc_g) == len(dc_o) and len(d_g) == 0
    lvo.close()
