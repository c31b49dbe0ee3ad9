#!/usr/bin/env python
"""Experiment: G contexts x (lanes / G) lanes on G streams, driven from G host threads, vs one context with all lanes.
Wall clock over `steps` frames after warm-up (device-resident sweeps)."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from oracle_py import Synth

def main():
    lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    groups = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4").split(",")]
    steps, warm = 12, 4
    L = bench.load_pkg()
    synth = Synth()
    total = steps + warm
    plan = bench.lane_plan(lanes, 8)
    max_off = max(o for _, o in plan)
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(s, f) for s in range(8) for f in range(total + max_off)]
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        res = list(ex.map(lambda j: synth.sweep(64, j[0], j[1])[0], jobs))
    dev = {j: torch.from_numpy(a).cuda() for j, a in zip(jobs, res)}
    for G in groups:
        per = lanes // G
        ctxs = [L.Lvo(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, lanes=per, device=0, max_points=131072, max_map_corner=1 << 18, max_map_surf=1 << 19) for _ in range(G)]
        bar = threading.Barrier(G + 1)
        def work(g):
            c = ctxs[g]
            pl = plan[g * per:(g + 1) * per]
            for k in range(total):
                if k == warm:
                    bar.wait(); bar.wait()
                ptrs = [dev[(s, k + o)].data_ptr() for s, o in pl]
                ns = [dev[(s, k + o)].shape[0] for s, o in pl]
                st, _, _ = c.step_batch_dev(ptrs, ns)
                assert st >= 0
        th = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        for t in th: t.start()
        bar.wait()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bar.wait()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"G={G} lanes/ctx={per}: {lanes * steps / dt:.0f} scans/s, {1e3 * dt / steps:.2f} ms per {lanes}-lane step", flush=True)
        for c in ctxs: c.close()

main()
