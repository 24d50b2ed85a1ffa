"""world_size-2 gloo test of the multi-rank host logic (sharding + timing aggregation)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_b200.shard import aggregate_throughput, shard_range


def test_shard_range_partitions_the_batch():
    for gb in (512, 64, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]      # BASELINE.json configs[2]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    q.put((rank, aggregate_throughput(hi - lo, 0.5 * (rank + 1))))
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_throughput_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        ips, total, tmax = res[r]
        assert total == 10 and tmax == 1.0 and ips == 10.0     # SUM of images / MAX of elapsed


def test_numa_binding_helper_never_raises_without_a_gpu():
    """bind_to_gpu_numa_node is best effort: on a box without NVML / CUDA it returns None and leaves the affinity alone."""
    import os
    from yolo_b200.shard import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), list)
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before
