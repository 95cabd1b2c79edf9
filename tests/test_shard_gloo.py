"""world_size-2 gloo test of the multi-GPU plumbing (openh264_b200/shard.py): streams are dealt to ranks without
overlap or gap, and the job figures are MAX(time) / SUM(pictures) over ranks.  No GPU needed."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, total, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from openh264_b200 import shard
    r, _, w = shard.init("gloo")
    mine = shard.streams_for_rank(total, r, w)
    shard.barrier()
    secs, pics = shard.job_totals(0.5 + 0.25 * r, 10 * len(mine))       # rank 1 is the slow one
    gathered = [None] * w
    dist.all_gather_object(gathered, mine)
    q.put((r, mine, secs, pics, gathered))
    dist.destroy_process_group()


def test_two_ranks_partition_and_totals():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, total, 29533, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_streams = sorted(s for _, mine, _, _, _ in res for s in mine)
    assert all_streams == list(range(total))                               # every stream exactly once
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for _, _, secs, pics, gathered in res:
        assert secs == 0.75 and pics == 10 * total                         # MAX over ranks, SUM over ranks
        assert gathered == [[0, 2, 4, 6], [1, 3, 5]]


def test_partition_properties():
    sys.path.insert(0, ROOT)
    from openh264_b200.shard import streams_for_rank
    for world in (1, 2, 4, 8):
        for total in (0, 1, 8, 33, 64):
            parts = [streams_for_rank(total, r, world) for r in range(world)]
            assert sorted(s for p in parts for s in p) == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
