#!/usr/bin/env python
"""Raw device->host ceiling for the e2e leg: every rank copies the bench's per-step output bytes (4 arrays of
125 x 1250 x 1250 int64 = 6.25 GB) from HBM into page-locked host memory with one cudaMemcpyAsync per array
(torch Tensor.copy_(non_blocking=True)), and the affinities (0.586 GB) the other way; CUDA-event timed, max over ranks.

    python tools/pcie_ceiling.py                       # one rank
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:  # noqa: BLE001
        pass
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = (125, 1250, 1250)
dev = [torch.zeros(shape, dtype=torch.int64, device="cuda") for _ in range(4)]
host = [torch.empty(shape, dtype=torch.int64, pin_memory=True) for _ in range(4)]
a_dev = torch.zeros((3,) + shape, dtype=torch.uint8, device="cuda")
a_host = torch.empty((3,) + shape, dtype=torch.uint8, pin_memory=True)


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    return best


def d2h():
    for d, h in zip(dev, host):
        h.copy_(d, non_blocking=True)


def h2d():
    a_dev.copy_(a_host, non_blocking=True)


def both():
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2):
        a_dev.copy_(a_host, non_blocking=True)
    d2h()
    torch.cuda.current_stream().wait_stream(s2)


d2h_bytes = 4 * dev[0].numel() * 8
h2d_bytes = a_dev.numel()
ms_d2h, ms_h2d, ms_both = timed(d2h), timed(h2d), timed(both)
if rank == 0:
    print(json.dumps({"n_ranks": world, "d2h_bytes_per_rank": d2h_bytes, "h2d_bytes_per_rank": h2d_bytes,
                      "d2h_ms": ms_d2h, "d2h_gbs_per_rank": d2h_bytes / ms_d2h / 1e6,
                      "h2d_ms": ms_h2d, "h2d_gbs_per_rank": h2d_bytes / ms_h2d / 1e6,
                      "both_directions_ms": ms_both,
                      "note": "one cudaMemcpyAsync per array (torch copy_ non_blocking) into page-locked memory; max over ranks, best of 3"}))
if world > 1:
    dist.destroy_process_group()
