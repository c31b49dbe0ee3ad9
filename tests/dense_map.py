"""Synthetic dense map of BASELINE.json configs[2] ("HDL-64 laserMapping dense map cube, 0.4 / 0.8 m voxels, ~500 k map pts"): ~100 k corner
+ ~400 k surf points inside the 5 x 5 x 3-cube neighbourhood of the origin (laserMapping.cpp:509-539), laid out so that they SURVIVE the
per-cube re-filter (:788-801): one point per 0.8 m voxel on four horizontal layers (surf), one point per 0.4 m voxel along vertical
poles (corner).  Shared by bench.py (config 3 measurement) and tests/test_gpu_dense_map.py (parity at that size).  numpy only, seeded."""
import numpy as np


def cube_index(p, cen=(10, 10, 5)):
    """laserMapping.cpp:741-750 with laserCloudCen* = cen: int((v + 25) / 50) + cen, minus one on the negative side."""
    v = p[:, :3].astype(np.float64) + 25.0
    c = np.trunc(v / 50.0).astype(np.int64) + np.array(cen)
    c[v < 0] -= 1
    return (c[:, 0] + 21 * c[:, 1] + 441 * c[:, 2]).astype(np.int32)


def make_dense_map(seed=7, half=124.0):
    rng = np.random.default_rng(seed)
    # surf: four layers, one jittered point per 0.8 m voxel column (voxel k covers [0.8 k, 0.8 (k + 1)))
    k = np.arange(int(-half / 0.8), int(half / 0.8))
    gx, gy = np.meshgrid(k, k, indexing="ij")
    layers = []
    for z in (-1.73, 2.3, 6.3, 10.3):
        n = gx.size
        x = (gx.reshape(-1) + 0.5) * 0.8 + rng.uniform(-0.3, 0.3, n)
        y = (gy.reshape(-1) + 0.5) * 0.8 + rng.uniform(-0.3, 0.3, n)
        zz = z + rng.normal(0, 0.02, n)
        layers.append(np.stack([x, y, zz, np.zeros(n)], 1))
    surf = np.concatenate(layers).astype(np.float32)
    # corner: vertical poles, one point per 0.4 m voxel
    nl = 3850
    px = (np.floor(rng.uniform(-half, half, (nl, 2)) / 0.4) + 0.5) * 0.4
    kz = np.arange(-4, 22)
    z = (kz + 0.5) * 0.4
    corner = np.concatenate([np.repeat(px, len(z), 0) + rng.uniform(-0.1, 0.1, (nl * len(z), 2)), np.tile(z, nl)[:, None] + rng.uniform(-0.1, 0.1, (nl * len(z), 1)),
                             np.zeros((nl * len(z), 1))], 1).astype(np.float32)
    return corner, cube_index(corner), surf, cube_index(surf)
