"""N>1 path.  Sites shard by contiguous range (mcb_job_partition, include/mcall_job.h), there is no data-path collective;
the only exchange is the max-over-ranks timing reduce and the ordered concatenation of results on the host.

  * CPU, world_size-2 gloo: both ranks take their range from the PRODUCT library's partition call (host arithmetic, loads
    without a GPU) the way bench.py's ranks do; the CPU oracle stands in for the kernels, which cannot run here.
  * GPU: mcb_job_call_host -- one job over several contexts (the same device listed twice on a one-GPU box, every visible
    device otherwise) must equal the single-context result bit for bit, in input order, compacted output included."""
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
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 24)
    first = mcall.partition(params.nsmpl, batch.nals, world)       # product code: the ranges mcb_job_call_host uses
    lo, hi = int(first[rank]), int(first[rank + 1])
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
    from bcftools_b200 import mcall
    params, batch, tab = synth.make_batch("C3", 24)
    full, _ = pyoracle.call("port", params, batch, tab)
    ret = sum((g[2] for g in sorted(gathered)), [])
    gt = sum((g[3] for g in sorted(gathered)), [])
    assert ret == full.ret.tolist() and gt == full.gt.tolist()
    first = mcall.partition(params.nsmpl, batch.nals, world)
    assert [g[:2] for g in sorted(gathered)] == [(int(first[0]), int(first[1])), (int(first[1]), int(first[2]))]
    # the ranges are contiguous, cover the job and carry about the same PL volume (not the same site count)
    vol = params.nsmpl * batch.ngt
    assert first[0] == 0 and first[-1] == batch.nsites and (np.diff(first) > 0).all()
    shares = [vol[first[k]:first[k + 1]].sum() for k in range(world)]
    assert max(shares) - min(shares) <= 2 * vol.max()


def test_partition_edge_cases():
    from bcftools_b200 import mcall
    nals = np.array([2, 5, 2, 2, 3, 2, 4, 2], np.uint8)
    for parts in (1, 2, 3, 8, 11):
        f = mcall.partition(100, nals, parts)
        assert f[0] == 0 and f[-1] == len(nals) and (np.diff(f) >= 0).all(), (parts, f)
    assert mcall.partition(7, np.zeros(0, np.uint8), 3).tolist() == [0, 0, 0, 0]


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("compact,typed", [(False, False), (True, False), (True, True)])
def test_job_over_several_contexts_equals_one_context(compact, typed):
    import torch
    from bcftools_b200 import abi, mcall, synth
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    params, batch, tab = synth.make_batch("C3", 96, flag=abi.CALL_VARONLY)
    with mcall.MCaller(params, ploidy_tab=tab, options={"slab_bytes": 2 << 20}) as mc:
        one = mc.call_host(batch, compact=compact, typed=typed)
    for devs, pack in ((devices, 0), ([0, 0, 0], 0), ([0, 0, 0], 1)):     # pack=1: the gaps between the ranges' blocks closed on the host
        with mcall.MJob(params, devs, ploidy_tab=tab, options={"slab_bytes": 2 << 20, "pack": pack}) as job:
            got = job.call_host(batch, compact=compact, typed=typed)
        assert got.first_site[0] == 0 and got.first_site[-1] == batch.nsites and len(got.first_site) == len(devs) + 1
        for name in ("ret", "als_new", "als_map", "ac", "an", "site_flags"):
            assert (getattr(got, name) == getattr(one, name)).all(), name
        assert (got.qual.view(np.uint32) == one.qual.view(np.uint32)).all()
        if typed:
            got.widen(); one.widen()
        assert (got.gt == one.gt).all() and (got.gq == one.gq).all()
        for i in range(batch.nsites):
            if one.ret[i] > 0 and not (one.site_flags[i] & abi.SITE_PL_DROPPED):
                assert (got.site_pl(i) == one.site_pl(i)).all(), i
        if compact:     # range after range: the blocks of range k lie behind those of range k-1
            offs = got.pl_off_out
            for k in range(len(devs) - 1):
                a = offs[got.first_site[k]:got.first_site[k + 1]]
                b = offs[got.first_site[k + 1]:got.first_site[k + 2]]
                if (a >= 0).any() and (b >= 0).any():
                    assert a[a >= 0].max() < b[b >= 0].min()

