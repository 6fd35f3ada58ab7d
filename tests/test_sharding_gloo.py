"""CPU suite, part 3: the N>1 plumbing on world_size-2 gloo -- unit sharding is a partition and the
whole-job throughput is (sum of units) / (max of times), exactly what bench.py reports at N GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cloud_transformers_b200.sharding import aggregate_throughput, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(33, rank, world)
    units, ms = aggregate_throughput(units_local=(hi - lo) * 16 * 2048, ms_local=10.0 + 5.0 * rank)
    q.put((rank, lo, hi, units, ms))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (1, 7, 32, 33, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_aggregate_without_process_group():
    assert aggregate_throughput(100, 2.5) == (100.0, 2.5)


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, u0, t0), (r1, lo1, hi1, u1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)
    assert u0 == u1 == 33 * 16 * 2048          # sum over ranks
    assert t0 == t1 == 15.0                    # max over ranks
