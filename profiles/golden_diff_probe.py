#!/usr/bin/env python
"""How far the GPU's per-outer-iteration counters are from the HDL-64 golden fixture (to size the tolerances of
tests/test_gpu_golden.py::test_four_hdl64_frames_fused_pipeline_against_golden)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_lvo
from oracle_py import Synth
L = load_lvo(); synth = Synth()
g = np.load(os.path.join(ROOT, "tests", "golden", "hdl64_seq0.npz"))
lvo = L.Lvo(max_map_corner=1 << 18, max_map_surf=1 << 19)
for k in range(4):
    st, odo, mp = lvo.step_batch([synth.sweep(64, 0, k)[0]])
    s = lvo.stats()
    line = [k, "ran", s.odo_outer_executed, s.map_outer_executed]
    if k > 0:
        got = np.stack([list(s.odo_corner_corr)[:10], list(s.odo_plane_corr)[:10], list(s.odo_lm_iters)[:10]], 1)
        line += ["odo dcount", int(np.abs(got - g[f"odo{k}_counts"]).max()), "dcost", float(np.max(np.abs(np.array(list(s.odo_final_cost)[:10]) / g[f"odo{k}_cost"] - 1)))]
        got = np.stack([list(s.map_corner_corr)[:10], list(s.map_surf_corr)[:10], list(s.map_lm_iters)[:10]], 1)
        line += ["map dcount", int(np.abs(got - g[f"map{k}_counts"]).max()), "dcost", float(np.max(np.abs(np.array(list(s.map_final_cost)[:10]) / g[f"map{k}_cost"] - 1)))]
        line += ["from_map", (s.map_corner_from_map, s.map_surf_from_map), g[f"map{k}_from_map_n"].tolist()]
    line += ["totals", (s.map_corner_total, s.map_surf_total), g[f"map{k}_totals"].tolist(),
             "dpos", float(np.linalg.norm(mp[0][4:] - g[f"map{k}_pose"][4:]))]
    print(*line)
