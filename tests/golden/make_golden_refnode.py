#!/usr/bin/env python
"""Generates tests/golden/refnode_hdl64_seq0.npz and refnode_vlp16_seq1.npz: outputs of the REFERENCE'S OWN node sources
(src/scanRegistration.cpp, src/laserOdometry.cpp, src/laserMapping.cpp compiled unmodified into oracle/_ref/libref_*.so against the
shim ROS / PCL / Ceres headers of oracle/shim; see oracle/refnode/refnode.cpp) on the seeded synthetic sweeps.

Runs only where /root/reference exists (this container); the vectors travel with the repo so that the hand restatement
(tests/test_refnode_golden_cpu.py) and the CUDA path (tests/test_gpu_golden.py) can be checked against the reference's text anywhere.
What is NOT the reference's: pcl::VoxelGrid, pcl::KdTreeFLANN and ceres::Solve are restated back ends (PCL / FLANN / Ceres are
absent from this image), and the scanRegistration node is the `lvo_atan` variant (deterministic atan / atan2 of csrc/lvo_math.h
instead of glibc's; DESIGN.md deviation 3) — the glibc variant differs in the last ulp of ~0.1 % of the intensities
(tests/test_refnode_cpu.py::test_glibc_atan_variant).
Run from the repo root:  python tests/golden/make_golden_refnode.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_py import Synth  # noqa: E402
from refnode_py import RefPipeline  # noqa: E402

FEATS = ("full", "sharp", "less_sharp", "flat", "less_flat")


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def lm_summary(records):
    """per ceres::Solve call (= outer iteration): residual blocks, LM iterations (rows after the initial evaluation), final cost, flags"""
    counts = np.array([[int(t[0, 0]), len(t) - 2] for t in records], np.int64)
    cost = np.array([t[-1, 7] for t in records])
    flags = np.full((len(records), 6), -1, np.int64)
    for i, t in enumerate(records):
        flags[i, :len(t) - 1] = t[1:, 9].astype(np.int64)
    poses = np.stack([t[-1, :7] for t in records])
    return counts, cost, flags, poses


def run(model, seq, cfg, frames, skip, path):
    synth = Synth()
    R = RefPipeline(*cfg, skip_frame=skip, lvo_atan=True)
    out = {"frames": np.int64(frames), "skip_frame": np.int64(skip), "config": np.array(cfg, np.float64)}
    for k in range(frames):
        pts, gt = synth.sweep(model, seq, k)
        out[f"sweep{k}_sha"] = digest(pts)
        r = R.step(pts, detail=True)
        for name, a in zip(FEATS, r["feats"]):
            out[f"f{k}_{name}_n"] = np.int64(len(a))
            out[f"f{k}_{name}_sha"] = digest(a)
        # the node's global arrays (scanRegistration.cpp:66-69) are only rewritten on [5, n - 5) each sweep (:256-266); the few entries
        # outside can hold marks left by an earlier, longer sweep and are never read, so they are not part of the fixture
        n = len(r["feats"][0])
        for name in ("curvature", "sort_ind", "picked", "label"):
            out[f"f{k}_{name}_sha"] = digest(r[name][5:n - 5])
        out[f"odo{k}_world"] = r["odom"]
        out[f"hf{k}"] = r["high_freq"]
        out[f"mapped{k}"] = np.int64(r["map"] is not None)
        if r["lm_odo"]:
            c, cost, fl, xs = lm_summary(r["lm_odo"])
            out[f"odo{k}_counts"], out[f"odo{k}_cost"], out[f"odo{k}_flags"], out[f"odo{k}_x"] = c, cost, fl, xs
        if r["map"] is not None:
            out[f"map{k}_pose"] = r["map"]
            out[f"map{k}_registered_sha"] = digest(r["registered"])
            out[f"map{k}_from_map_n"] = np.array([len(r["corner_from_map"]), len(r["surf_from_map"])], np.int64)
            out[f"map{k}_corner_from_map_sha"] = digest(r["corner_from_map"])
            out[f"map{k}_surf_from_map_sha"] = digest(r["surf_from_map"])
            out[f"map{k}_totals"] = np.array([len(r["map_corner"]), len(r["map_surf"])], np.int64)
            out[f"map{k}_corner_cube_sha"] = digest(r["map_corner_cube"])
            out[f"map{k}_surf_cube_sha"] = digest(r["map_surf_cube"])
            out[f"map{k}_corner_sha"] = digest(r["map_corner"])
            out[f"map{k}_surf_sha"] = digest(r["map_surf"])
            out[f"map{k}_cen"] = r["cen"].astype(np.int64)
            if r["lm_map"]:
                c, cost, fl, xs = lm_summary(r["lm_map"])
                out[f"map{k}_counts"], out[f"map{k}_cost"], out[f"map{k}_flags"], out[f"map{k}_x"] = c, cost, fl, xs
    R.close()
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run(64, 0, (64, 5.0, 0.4, 0.8), 5, 1, os.path.join(HERE, "refnode_hdl64_seq0.npz"))
    run(16, 1, (16, 0.3, 0.2, 0.4), 5, 1, os.path.join(HERE, "refnode_vlp16_seq1.npz"))
    run(64, 2, (64, 5.0, 0.4, 0.8), 6, 2, os.path.join(HERE, "refnode_hdl64_seq2_skip2.npz"))
