"""N>1 path on CPU: world_size-2 gloo.  Sites shard by contiguous range, there is no data-path collective; the only
exchange is the max-over-ranks timing reduce and the ordered concatenation of results on the host."""
import os
import socket
import sys

import numpy as np
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


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bcftools_b200 import synth
    from oracle import pyoracle
    # the whole job is 64 sites; each rank takes a contiguous range and its own seed offset does NOT matter here:
    # shards are cut from ONE batch so that the concatenation can be compared with the single-process result
    params, batch, tab = synth.make_batch("C1", 64)
    lo, hi = 64 * rank // world, 64 * (rank + 1) // world
    res, secs = pyoracle.call("port", params, batch.subset(range(lo, hi)), tab)
    t = torch.tensor([0.01 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                      # bench.py's max-over-ranks timing
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, res.ret.tolist(), res.gt.tolist()))
    if rank == 0:
        out.put((float(t.item()), gathered))
    dist.destroy_process_group()


def test_site_range_sharding_concatenates_to_the_single_process_result():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    tmax, gathered = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
    assert abs(tmax - 0.02) < 1e-12
    sys.path.insert(0, ROOT)
    from bcftools_b200 import synth
    from oracle import pyoracle
    params, batch, tab = synth.make_batch("C1", 64)
    full, _ = pyoracle.call("port", params, batch, tab)
    ret = sum((g[2] for g in sorted(gathered)), [])
    gt = sum((g[3] for g in sorted(gathered)), [])
    assert ret == full.ret.tolist() and gt == full.gt.tolist()
    assert [g[:2] for g in sorted(gathered)] == [(0, 32), (32, 64)]
