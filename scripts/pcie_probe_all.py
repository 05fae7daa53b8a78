"""Aggregate host<->device copy rate of the box with every GPU copying at once (torchrun, one rank per GPU): the ceiling of the
N-GPU end-to-end path.  Each rank moves the bytes of one e2e step (365 MB up, 454 MB down, pinned, both directions at once)
`reps` times between two barriers; rank 0 prints per-rank and aggregate GB/s, for all ranks together and for rank 0 alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/pcie_probe_all.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:    # the CPUs NVML reports as local to this GPU (what bench.py binds to)
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
except Exception:
    cpus = set()
UP, DOWN, reps = 365 << 20, 454 << 20, 5
h_in, h_out = torch.empty(UP, dtype=torch.uint8).pin_memory(), torch.empty(DOWN, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(UP, dtype=torch.uint8, device="cuda"), torch.empty(DOWN, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def timed(active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


step(); torch.cuda.synchronize()
t_all = timed(True)
t_one = timed(rank == 0)
if rank == 0:
    per = (UP + DOWN) * reps / 1e9
    print(json.dumps(dict(n_gpus=world, bytes_per_rank_per_step=UP + DOWN, all_ranks_GBs=per * world / t_all, per_rank_GBs_when_all_copy=per / t_all,
                          rank0_alone_GBs=per / t_one, cpus_bound=len(cpus), host_cpus=os.cpu_count())))
if world > 1:
    dist.destroy_process_group()
