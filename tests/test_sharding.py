"""Host-side multi-GPU logic on CPU: the frame/stream partition and the max-over-ranks reduction
used by bench.py, exercised with a real world_size-2 gloo process group."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitions_are_disjoint_and_complete(fov):
    sh = fov.sharding
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 16, 64):
            owned = [sh.frames_for_rank(n, world, r) for r in range(world)]
            flat = sorted(i for part in owned for i in part)
            assert flat == list(range(n))
            for r, part in enumerate(owned):
                assert all(sh.owner_of(i, world) == r for i in part)
    # 64 streams over 8 GPUs: 8 streams per GPU (BASELINE.json configs[4])
    assert [len(sh.streams_for_rank(64, 8, r)) for r in range(8)] == [8] * 8
    with pytest.raises(ValueError):
        sh.frames_for_rank(4, 2, 2)


def test_aggregate_is_units_over_slowest_rank(fov):
    assert fov.sharding.aggregate_throughput([8, 8], [1.0, 2.0]) == 8.0


def _worker(rank, world, port, q):
    import importlib

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sh = importlib.import_module("foveated-360-video_b200").sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    try:
        frames = sh.frames_for_rank(21, world, rank)
        local_seconds = 0.5 + rank  # rank 1 is the slow one
        slow = sh.reduce_max_seconds(local_seconds, dist)
        total = sh.reduce_sum_int(len(frames), dist)
        dist.barrier()
        q.put((rank, frames, slow, total))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_reduction():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, slow0, tot0), (r1, f1, slow1, tot1) = results
    assert sorted(f0 + f1) == list(range(21)) and not set(f0) & set(f1)
    assert slow0 == slow1 == 1.5  # both ranks agree on the slowest time
    assert tot0 == tot1 == 21
