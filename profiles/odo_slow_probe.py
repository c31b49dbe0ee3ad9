"""How many scan-to-scan features leave the thread-per-feature association for the warp-per-feature kernel, per outer iteration, and why
(lvo_stats::odo_slow / odo_slow_why).  usage (under gpurun): python profiles/odo_slow_probe.py [frames]"""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("lvo_b200", os.path.join(ROOT, "lidar-visual-odometry_b200", "__init__.py"))
L = importlib.util.module_from_spec(spec); sys.modules["lvo_b200"] = L; spec.loader.exec_module(L)
from oracle_py import Synth
synth = Synth()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
c = L.Lvo(lanes=2, max_map_corner=1 << 18, max_map_surf=1 << 19, debug_probes=0)
c.set_option(L.LVO_OPT_FIXPOINT_SKIP, 0)
for k in range(n):
    c.step_batch([synth.sweep(64, 0, k)[0], synth.sweep(64, 7, k)[0]])
    if k >= n - 3:
        for lane in range(2):
            s = c.stats(lane)
            print(f"frame {k} lane {lane}: sharp {s.n_sharp} flat {s.n_flat} | slow per outer {list(s.odo_slow)[:10]} | why {list(s.odo_slow_why)} | "
                  f"corner corr {list(s.odo_corner_corr)[:3]} plane corr {list(s.odo_plane_corr)[:3]} | map knn full {list(s.map_knn_full)[:10]} | odo certified {list(s.odo_certified)[:10]}")
c.close()
