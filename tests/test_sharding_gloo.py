"""world_size-2 gloo test of the only multi-GPU logic the path has: canvases are sharded by
contiguous index blocks, no collective on the data path; the bench's max-over-ranks timing
reduction is the single all_reduce."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_transformation_b200.batch import shard_range
    from image_transformation_b200 import synth

    lo, hi = shard_range(n_items, rank, world)
    sizes = {1: (300, 200), 2: (640, 480)}
    # every rank derives its own canvases from the global canvas index: no exchange needed
    mine = [synth.canvas_placements(sizes, (1920, 1080), i, n_objects=5) for i in range(lo, hi)]
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py: max over ranks of the device time
    cnt = torch.tensor([hi - lo])
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    q.put((rank, lo, hi, mine[0][0]["box"], float(t), int(cnt)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 1024])
def test_shard_by_canvas_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, box0, t0, c0), (r1, lo1, hi1, box1, t1, c1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == n_items  # contiguous, disjoint, complete
    assert abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    assert t0 == t1 == 2.0 and c0 == c1 == n_items
    sys.path.insert(0, ROOT)
    from image_transformation_b200 import synth

    sizes = {1: (300, 200), 2: (640, 480)}
    assert box1 == synth.canvas_placements(sizes, (1920, 1080), lo1, n_objects=5)[0]["box"]
