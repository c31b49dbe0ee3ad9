"""Scratch: where does host time go in the pipelined e2e step?"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle_py import Synth
L = bench.load_pkg()
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
synth = Synth()
frames = 8
sw = {(s, f): synth.sweep(64, s, f)[0] for s in range(8) for f in range(frames)}
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in sw.items()}
ctx = L.Lvo(lanes=lanes, max_points=131072, max_map_corner=1 << 18, max_map_surf=1 << 19)
views = [[pinned[(l % 8, f)].numpy() for l in range(lanes)] for f in range(frames)]
for mode in ("plain", "pipelined"):
    for f in range(frames):
        t0 = time.perf_counter()
        if mode == "plain":
            ctx.step_batch(views[f])
        else:
            ctx.step_batch_pipelined(views[f], views[f + 1] if f + 1 < frames else None)
        t1 = time.perf_counter()
        tm = ctx.timings()
        print(mode, f, "wall ms %.2f" % (1e3 * (t1 - t0)), "gpu stages ms %.2f" % (tm.extract_ms + tm.odometry_ms + tm.mapping_ms), flush=True)
# raw copy speed from the same pinned buffers on a side stream
s2 = torch.cuda.Stream()
dst = torch.empty((lanes, 131072, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for l in range(lanes):
        src = pinned[(l % 8, 0)]
        dst[l, :src.shape[0]].copy_(src, non_blocking=True)
t1 = time.perf_counter()
s2.synchronize()
t2 = time.perf_counter()
print("torch pinned copies: enqueue ms %.2f total ms %.2f" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
