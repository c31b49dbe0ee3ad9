#!/usr/bin/env python
"""Generates tests/golden/vlp16_seq1.npz: three synthetic VLP-16 sweeps and what the CPU oracle computes from them.

The reference ships no fixtures or golden vectors for this path, and its PCL/Ceres binaries cannot be
built here, so these vectors are produced by the oracle (oracle/liblvo_oracle.so) — they pin the oracle and the CUDA
path against regressions and against each other, not against reference binaries ("parity unpinned").
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_py import Oracle, Synth  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def main():
    synth = Synth()
    orc = Oracle(16, 0.3, 0.2, 0.4)
    out = {}
    for k in range(3):
        pts, gt = synth.sweep(16, 1, k)
        out[f"sweep{k}"] = pts[:, :3].copy()
        f = orc.extract(pts)
        for name in ("full", "sharp", "less_sharp", "flat", "less_flat"):
            out[f"f{k}_{name}_n"] = np.int64(len(f[name]))
            out[f"f{k}_{name}_sha"] = digest(f[name])
        out[f"f{k}_label_sha"] = digest(f["label"])
        out[f"f{k}_sort_ind_sha"] = digest(f["sort_ind"])
        out[f"f{k}_curvature_sha"] = digest(f["curvature"])
        if k == 0:
            out["f0_sharp"] = f["sharp"]
            out["f0_flat"] = f["flat"]
        st, rel, w = orc.odometry(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        out[f"odo{k}_status"] = np.int64(st)
        out[f"odo{k}_rel"] = rel
        out[f"odo{k}_world"] = w
        if k > 0:
            lg = orc.odometry_log(0)
            out[f"odo{k}_corner_corr0"] = lg["corner_corr"]
            out[f"odo{k}_plane_corr0"] = lg["plane_corr"]
        st, pose, corr = orc.mapping(f["less_sharp"], f["less_flat"], f["full"], w)
        out[f"map{k}_status"] = np.int64(st)
        out[f"map{k}_pose"] = pose
        info = orc.mapping_info()
        out[f"map{k}_corner_stack_sha"] = digest(info["corner_stack"])
        out[f"map{k}_surf_stack_sha"] = digest(info["surf_stack"])
        if st == 0:
            lg = orc.mapping_log(0)
            out[f"map{k}_corner_knn0"] = lg["corner_knn"]
            out[f"map{k}_surf_knn0_sha"] = digest(lg["surf_knn"])
        pc, cc = orc.map_export(0)
        ps, cs = orc.map_export(1)
        out[f"map{k}_totals"] = np.array([len(pc), len(ps)], np.int64)
    # voxel-grid and kNN known answers on a tiny hand-checkable cloud
    tiny = np.array([[0.05, 0.05, 0.05, 1], [0.15, 0.05, 0.05, 3], [0.25, 0.05, 0.05, 5], [-0.05, 0.05, 0.05, 7], [0.05, 0.25, 0.05, 9]], np.float32)
    vg, idx, order = orc.voxel_grid(tiny, 0.2)
    out["tiny"] = tiny
    out["tiny_voxel"] = vg
    out["tiny_voxel_idx"] = idx
    np.savez_compressed(os.path.join(HERE, "vlp16_seq1.npz"), **out)
    print("wrote", os.path.join(HERE, "vlp16_seq1.npz"), os.path.getsize(os.path.join(HERE, "vlp16_seq1.npz")), "bytes")


if __name__ == "__main__":
    main()
