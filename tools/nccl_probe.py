"""NCCL sanity numbers of a box: all-gather and ring send/recv of 64 MiB per rank (torchrun --nproc-per-node N tools/nccl_probe.py)"""
import os

import torch
import torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x=torch.empty(64*1024*1024//4, dtype=torch.int32, device="cuda"); out=torch.empty(world*x.numel(), dtype=torch.int32, device="cuda")
for _ in range(3): dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize(); dist.barrier()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10): dist.all_gather_into_tensor(out, x)
ev[1].record(); torch.cuda.synchronize()
ms=ev[0].elapsed_time(ev[1])/10
# p2p send/recv
y=torch.empty_like(x)
dist.barrier(); torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    ops=[dist.P2POp(dist.isend, x, (rank+1)%world), dist.P2POp(dist.irecv, y, (rank-1)%world)]
    for r in dist.batch_isend_irecv(ops): r.wait()
ev[1].record(); torch.cuda.synchronize()
ms2=ev[0].elapsed_time(ev[1])/10
if rank==0:
    print(f"all_gather 64 MiB/rank x{world}: {ms:.3f} ms -> {(world-1)*64/1024/(ms/1e3):.1f} GiB/s received per rank; ring send/recv 64 MiB: {ms2:.3f} ms -> {64/1024/(ms2/1e3):.1f} GiB/s", flush=True)
    print("can_access_peer 0->1:", torch.cuda.can_device_access_peer(0,1), flush=True)
dist.destroy_process_group()
