"""CPU test of the N>1 plumbing (world size 2, gloo): streams shard across ranks with no data-path
collective; only the timing/token aggregation crosses ranks (max of time, sum of tokens)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = bench.RankReducer(dist, "cpu")
    assert bench.dist_env() == (rank, rank, world)
    # rank r transcribed 8 streams x 108 tokens x 3 steps in (1.0 + 0.5 r) s
    value, total, t = bench.aggregate_throughput(red, 8 * 108 * 3, 1.0 + 0.5 * rank)
    red.barrier()
    audio = bench.make_audio(2, rank)          # every rank gets its own streams
    out.put((rank, value, total, t, float(audio[0, :100].sum())))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, value, total, t, _ in res:
        assert total == 2 * 8 * 108 * 3
        assert t == pytest.approx(1.5)                       # slowest rank
        assert value == pytest.approx(total / 1.5)
    assert res[0][4] != res[1][4]                            # different audio per rank


def test_single_rank_identity():
    import bench
    red = bench.RankReducer(None)
    assert red.max(2.5) == 2.5 and red.sum(3.0) == 3.0
    assert bench.aggregate_throughput(red, 100, 2.0) == (50.0, 100.0, 2.0)
