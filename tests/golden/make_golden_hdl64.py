#!/usr/bin/env python
"""Generates tests/golden/hdl64_seq0.npz: what the CPU oracle computes from the first four synthetic HDL-64 sweeps of sequence 0
(BASELINE.json configs[0] shape: 64 rings, 120 000 returns per sweep) run through the three stages back to back.

The sweeps themselves are not stored (1.9 MB each): the seeded generator (synth/lvo_synth.cpp) travels with the repo and the
fixture holds the SHA-256 of every input sweep, so a drift of the generator is caught too.  Like vlp16_seq1.npz these vectors
come from the oracle, not from reference binaries (PCL / Ceres are absent here): they pin the oracle and the CUDA path against
regressions and against each other ("parity unpinned").
Run from the repo root:  python tests/golden/make_golden_hdl64.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_py import Oracle, Synth  # noqa: E402

FRAMES = 4


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def main():
    synth = Synth()
    orc = Oracle(64, 5.0, 0.4, 0.8)
    out = {"frames": np.int64(FRAMES)}
    for k in range(FRAMES):
        pts, gt = synth.sweep(64, 0, k)
        out[f"sweep{k}_n"] = np.int64(len(pts))
        out[f"sweep{k}_sha"] = digest(pts)
        out[f"gt{k}"] = gt
        f = orc.extract(pts)
        for name in ("full", "sharp", "less_sharp", "flat", "less_flat"):
            out[f"f{k}_{name}_n"] = np.int64(len(f[name]))
            out[f"f{k}_{name}_sha"] = digest(f[name])
        for name in ("label", "sort_ind", "curvature", "picked", "scan_start", "scan_end"):
            out[f"f{k}_{name}_sha"] = digest(f[name])
        st, rel, w = orc.odometry(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        out[f"odo{k}_status"] = np.int64(st)
        out[f"odo{k}_rel"] = rel
        out[f"odo{k}_world"] = w
        if k > 0:
            logs = [orc.odometry_log(it) for it in range(10)]
            out[f"odo{k}_counts"] = np.stack([lg["counts"] for lg in logs])     # per outer iteration: corner / plane factors, LM iterations
            out[f"odo{k}_cost"] = np.concatenate([lg["cost"] for lg in logs])
        st, pose, corr = orc.mapping(f["less_sharp"], f["less_flat"], f["full"], w)
        out[f"map{k}_status"] = np.int64(st)
        out[f"map{k}_pose"] = pose
        info = orc.mapping_info()
        out[f"map{k}_corner_stack_n"] = np.int64(len(info["corner_stack"]))
        out[f"map{k}_surf_stack_n"] = np.int64(len(info["surf_stack"]))
        out[f"map{k}_corner_stack_sha"] = digest(info["corner_stack"])
        out[f"map{k}_surf_stack_sha"] = digest(info["surf_stack"])
        out[f"map{k}_from_map_n"] = np.array([len(info["corner_from_map"]), len(info["surf_from_map"])], np.int64)
        if st == 0:
            logs = [orc.mapping_log(it) for it in range(10)]
            out[f"map{k}_counts"] = np.stack([lg["counts"] for lg in logs])
            out[f"map{k}_cost"] = np.concatenate([lg["cost"] for lg in logs])
        pc, cc = orc.map_export(0)
        ps, cs = orc.map_export(1)
        out[f"map{k}_totals"] = np.array([len(pc), len(ps)], np.int64)
    path = os.path.join(HERE, "hdl64_seq0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
