"""GPU parity of the stand-alone operators: voxel downsample (pcl::VoxelGrid) and exact kNN (pcl::KdTreeFLANN)."""
import numpy as np
import pytest

from oracle_py import Oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx(lvo_mod):
    lvo = lvo_mod.Lvo(n_scans=64, minimum_range=5.0, max_map_corner=1 << 18, max_map_surf=1 << 20)
    yield lvo
    lvo.close()


@pytest.mark.parametrize("leaf", [0.2, 0.4, 0.8])
def test_voxel_downsample_bit_exact(ctx, synth, leaf):
    O = Oracle()
    pts, _ = synth.sweep(64, 0, 0)
    f = O.extract(pts)
    for name in ("less_sharp", "less_flat", "full"):
        ref, _, _ = O.voxel_grid(f[name], leaf)
        got = ctx.voxel_downsample(f[name], leaf)
        assert got.shape == ref.shape, (name, leaf, got.shape, ref.shape)
        assert np.array_equal(_bits(got), _bits(ref)), (name, leaf)


def test_voxel_edge_cases(ctx):
    O = Oracle()
    rng = np.random.default_rng(3)
    assert len(ctx.voxel_downsample(np.zeros((0, 4), np.float32), 0.4)) == 0
    one = rng.normal(0, 10, (1, 4)).astype(np.float32)
    assert np.array_equal(_bits(ctx.voxel_downsample(one, 0.4)), _bits(O.voxel_grid(one, 0.4)[0]))
    # many points in one voxel (long run, sequential float accumulation order matters)
    dense = (rng.random((5000, 4)) * 0.3).astype(np.float32)
    assert np.array_equal(_bits(ctx.voxel_downsample(dense, 0.4)), _bits(O.voxel_grid(dense, 0.4)[0]))
    # negative coordinates / floor behaviour
    neg = (rng.random((20000, 4)) * 40 - 20).astype(np.float32)
    assert np.array_equal(_bits(ctx.voxel_downsample(neg, 0.4)), _bits(O.voxel_grid(neg, 0.4)[0]))
    # "leaf size too small": dx*dy*dz > INT_MAX -> the cloud passes through unchanged
    huge = (rng.random((3000, 4)) * 4000 - 2000).astype(np.float32)
    ref = O.voxel_grid(huge, 0.01)[0]
    assert len(ref) == len(huge)
    assert np.array_equal(_bits(ctx.voxel_downsample(huge, 0.01)), _bits(ref))


def test_knn5_bit_exact_indices(ctx, synth):
    O = Oracle()
    pts, _ = synth.sweep(64, 0, 0)
    f = O.extract(pts)
    cloud, _, _ = O.voxel_grid(f["less_flat"], 0.8)
    q, _, _ = O.voxel_grid(f["less_flat"], 0.8)
    q = q + np.array([0.05, -0.03, 0.02, 0], np.float32)
    ref_i, ref_d = O.knn(cloud, q, 5, 1.0, method=0)
    got_i, got_d = ctx.knn(cloud, q, 5, 1.0)
    assert np.array_equal(got_i, ref_i), f"{(got_i != ref_i).any(axis=1).sum()} of {len(q)} rows differ"
    assert np.array_equal(_bits(got_d), _bits(ref_d))
    assert (ref_i[:, 0] >= 0).mean() > 0.3  # the case is not vacuous


def test_knn1_wide_gate(ctx, synth):
    O = Oracle()
    a, _ = synth.sweep(64, 0, 0)
    b, _ = synth.sweep(64, 0, 1)
    fa, fb = O.extract(a), Oracle().extract(b)
    ref_i, ref_d = O.knn(fa["less_flat"], fb["flat"], 1, 25.0, method=0)
    got_i, got_d = ctx.knn(fa["less_flat"], fb["flat"], 1, 25.0)
    assert np.array_equal(got_i, ref_i) and np.array_equal(_bits(got_d), _bits(ref_d))
    ref_i, _ = O.knn(fa["less_sharp"], fb["sharp"], 1, 25.0, method=0)
    got_i, _ = ctx.knn(fa["less_sharp"], fb["sharp"], 1, 25.0)
    assert np.array_equal(got_i, ref_i)


def test_knn_ties_and_edges(ctx):
    """Exact float ties are broken by index; fewer than K points inside the gate -> -1."""
    O = Oracle()
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.5
    cloud = np.concatenate([g, np.zeros((len(g), 1), np.float32)], 1)
    q = cloud[::3].copy()
    q[:, :3] += 0.25  # equidistant from 8 lattice points
    ref_i, ref_d = O.knn(cloud, q, 5, 1.0, method=0)
    got_i, got_d = ctx.knn(cloud, q, 5, 1.0)
    assert np.array_equal(got_i, ref_i) and np.array_equal(_bits(got_d), _bits(ref_d))
    far = q + 100
    got_i, _ = ctx.knn(cloud, far.astype(np.float32), 5, 1.0)
    assert (got_i == -1).all()
    got_i, _ = ctx.knn(cloud[:3], q, 5, 1.0)
    assert (got_i == -1).all()


def test_knn5_throughput_entry_matches_oracle(lvo_mod):
    """lvo_knn5_throughput: several independent (map, query) problems in one launch; every problem checked against the oracle."""
    import torch
    O = Oracle()
    rng = np.random.default_rng(11)
    maps, queries, mc, qc = [], [], [], []
    for p in range(5):
        m = np.concatenate([(rng.random((4000 + 500 * p, 3)) * [30, 30, 3] + [10 * p, -5, 0]).astype(np.float32), np.zeros((4000 + 500 * p, 1), np.float32)], 1)
        q = m[rng.choice(len(m), 700, replace=False)] + np.array([0.1, 0.05, -0.02, 0], np.float32)
        maps.append(m); queries.append(q); mc.append(len(m)); qc.append(len(q))
    dm, dq = torch.from_numpy(np.concatenate(maps)).cuda(), torch.from_numpy(np.concatenate(queries)).cuda()
    ind = torch.empty((len(dq), 5), dtype=torch.int32, device="cuda")
    sq = torch.empty((len(dq), 5), dtype=torch.float32, device="cuda")
    ctx = lvo_mod.Lvo(max_points=1 << 16, max_map_corner=1 << 16, max_map_surf=1 << 16)
    ms = ctx.knn5_throughput(dm.data_ptr(), mc, dq.data_ptr(), qc, 2, ind.data_ptr(), sq.data_ptr())
    assert ms > 0
    got_i, got_d = ind.cpu().numpy(), sq.cpu().numpy()
    o = 0
    for m, q in zip(maps, queries):
        ref_i, ref_d = O.knn(m, q, 5, 1.0, method=0)
        assert np.array_equal(got_i[o:o + len(q)], ref_i) and np.array_equal(_bits(got_d[o:o + len(q)]), _bits(ref_d))
        o += len(q)
    ctx.close()
