#!/usr/bin/env python
"""Single-trajectory latency (1 lane, CUDA-graph replay): ms per frame of extract + scan-to-scan + scan-to-map, HDL-64 and VLP-16."""
import importlib.util, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("lvo_b200", os.path.join(ROOT, "lidar-visual-odometry_b200", "__init__.py"))
L = importlib.util.module_from_spec(spec); sys.modules["lvo_b200"] = L; spec.loader.exec_module(L)
import torch
from oracle_py import Synth
synth = Synth()
out = {}
for rings, kw in ((64, dict(minimum_range=5.0)), (16, dict(minimum_range=0.3, line_res=0.2, plane_res=0.4))):
    ctx = L.Lvo(n_scans=rings, lanes=1, max_points=131072, max_map_corner=1 << 18, max_map_surf=1 << 19, **kw)
    s = torch.cuda.Stream(); ctx.set_stream(s.cuda_stream)
    sweeps = [torch.from_numpy(synth.sweep(rings, 0, k)[0]).cuda() for k in range(50)]
    lat = []
    for k, sw in enumerate(sweeps):
        t0 = time.perf_counter()
        ctx.step_batch_dev([sw.data_ptr()], [sw.shape[0]])
        if k >= 10:
            lat.append(1e3 * (time.perf_counter() - t0))
    ctx.close()
    out[f"rings{rings}_ms_per_frame"] = {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90))}
print(json.dumps(out))
