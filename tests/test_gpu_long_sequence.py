"""Integration: a 100-frame synthetic HDL-64 sequence through the fused device pipeline against the oracle
pipeline, per-frame odometry and mapped poses within 1e-4 m / 1e-5 rad; the window of map cubes and the map sizes follow."""
import numpy as np
import pytest

from oracle_py import Oracle

pytestmark = pytest.mark.gpu


def rot_err(qa, qb):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))


def test_100_frames_against_oracle(lvo_mod, synth):
    L = lvo_mod
    lvo = L.Lvo(lanes=1, max_map_corner=1 << 19, max_map_surf=1 << 20)
    O = Oracle()
    worst_t = worst_r = 0.0
    for k in range(100):
        sw, gt = synth.sweep(64, 5, k)
        st, odo_g, map_g = lvo.step_batch([sw])
        _, odo_o, map_o = O.step(sw)
        et = max(np.linalg.norm(odo_g[0][4:] - odo_o[4:]), np.linalg.norm(map_g[0][4:] - map_o[4:]))
        er = max(rot_err(odo_g[0][:4], odo_o[:4]), rot_err(map_g[0][:4], map_o[:4]))
        worst_t, worst_r = max(worst_t, et), max(worst_r, er)
        assert et < 1e-4 and er < 1e-5, (k, et, er)
        if k % 20 == 19:
            s = lvo.stats(0)
            info = O.mapping_info()["info"]
            assert list(s.center_cube) == list(info[:3]) and abs(s.map_corner_total - info[6]) <= 10 and abs(s.map_surf_total - info[7]) <= 10
    print(f"worst translation deviation {worst_t:.3e} m, worst rotation deviation {worst_r:.3e} rad over 100 frames; "
          f"final position {map_g[0][4:]}, ground truth {gt[4:]}")
    assert np.linalg.norm(map_g[0][4:] - gt[4:]) < 2.0   # and the trajectory is a sane odometry solution (1 % drift budget)
    lvo.close()
