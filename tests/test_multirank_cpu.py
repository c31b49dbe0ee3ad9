"""World-size-2 gloo test (CPU) of bench.py's multi-rank plumbing: lanes are sharded across ranks with no data-path
collective; the only exchange is the max-over-ranks of the timing.  Two ranks running the SAME sequence through
independent oracle pipelines must also agree bitwise (the replicas-only contract of SURVEY §8e)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle_py import Oracle, Synth
    plan = bench.lane_plan(6, 4)
    seqs = bench.rank_sequence_base(rank)
    ms = bench.max_over_ranks([10.0 + rank, 20.0 - rank], world, device="cpu")
    s, o = Synth(), Oracle()
    poses = []
    for k in range(3):
        _, odo, mp_ = o.step(s.sweep(64, 0, k)[0])
        poses.append(np.r_[odo, mp_])
    t = torch.from_numpy(np.array(poses))
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    q.put((rank, plan, seqs, ms, [g.numpy() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, plan0, seq0, ms0, g0), (r1, plan1, seq1, ms1, g1) = res
    assert plan0 == plan1 == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]   # lanes of one sequence are adjacent, one frame apart
    assert seq0 != seq1                                   # ranks work on different sequences
    assert ms0 == ms1 == [11.0, 20.0]                     # max over ranks, identical on every rank
    assert np.array_equal(g0[0], g0[1]) and np.array_equal(g0[0], g1[1])  # same sequence -> bitwise identical poses on any rank
