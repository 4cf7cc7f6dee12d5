#!/usr/bin/env python
"""Headline benchmark: voxels/s, affinities -> segmentation (BASELINE.json metric), blockwise ws path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic CREMI-sized uint8 affinities 3x(125,1250,1250) per GPU,
block_shape (25,250,250), context (3,31,31) (the reference's //8 rule), ws defaults, 3 thresholds.
N > 1: weak scaling — the volume is N slabs of 125 planes stacked in z, one slab per rank (one process per
GPU); ranks exchange fragment halos and RAG edges (bootstrapper_b200/sharded.py).
A step = one pass of the whole path over one volume: stage 1 fragments, stage 2 RAG + agglomeration scores,
stage 3 thresholded CC + relabel for every threshold.

`value`   device-timed (CUDA events, max over ranks), affinities already resident in HBM.
`e2e`     the same through the public API with HOST buffers (ShardedSegmenter.run_host, streaming form): every
          step uploads its affinities from pinned host memory and downloads fragments + all segmentations into
          pinned host memory, all copies inside the timed region; the downloads of one volume overlap the upload
          and compute of the next (two buffer sets), the region ends when the last download is through.
`--impl reference` times the CPU oracle (a port of the reference path; the reference itself cannot be
          installed here) with one process per block on all host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (125, 1250, 1250)
BLOCK = (25, 250, 250)
CONTEXT = (3, 31, 31)
# --config 5 (not the driver's default): BASELINE configs[4], 2048^3 over 8 GPUs = a (256, 2048, 2048) slab per rank
SHAPE5, BLOCK5, CONTEXT5 = (256, 2048, 2048), (256, 256, 256), (32, 32, 32)
# --config 4: BASELINE configs[3], 1024^3 with 3-D seeded fragments + seed_eps over 4 GPUs = a (256, 1024, 1024) slab per rank
SHAPE4, BLOCK4, CONTEXT4 = (256, 1024, 1024), (128, 128, 128), (16, 16, 16)
PARAMS4 = {"fragments_in_xy": False, "seed_eps": 0.01}
THRESHOLDS = [0.2, 0.35, 0.5]
BYTES_PER_VOXEL = 3 * 1 + 8 + 8 * len(THRESHOLDS)      # u8 affs in, u64 fragments + T u64 segmentations out
CPU_SAMPLE = (100, 1000, 1000)                         # 64 blocks of the same geometry (~15 s of CPU work on 16 cores)
METRIC = "voxels/sec affs->segmentation (blockwise ws, fragments + RAG + agglomeration at 3 thresholds)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sample_affs(shape, seed):
    """the synthetic input of the CPU leg: the device generator (bit-identical to the numpy one, ~1000x faster) when a
    GPU is present -- input generation is outside every timed region"""
    try:
        import torch
        if torch.cuda.is_available():
            from bootstrapper_b200 import native
            return native.synth_affs(shape, seed=seed).cpu().numpy()
    except Exception:  # noqa: BLE001
        pass
    from bootstrapper_b200.synth import synth_affs
    return synth_affs(shape, seed=seed)


def cpu_oracle_rate(sample_shape, workers=None, seed=0):
    """voxels/s of the CPU oracle (one process per block) on a bounded sample of the workload"""
    from oracle.parallel import waterz_pipeline_parallel
    affs = sample_affs(sample_shape, seed)
    tm = {}
    waterz_pipeline_parallel(affs, {"thresholds": THRESHOLDS}, block_size=BLOCK, context=CONTEXT, timings=tm, workers=workers)
    return float(np.prod(sample_shape)) / tm["total"], tm


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample = (50, 500, 500) if args.quick else CPU_SAMPLE
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_oracle_rate((25, 250, 250), cores)
    rates, ms = [], []
    for _ in range(max(1, args.steps)):
        r, tm = cpu_oracle_rate(sample, cores)
        rates.append(r)
        ms.append(tm["total"] * 1e3)
    v = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "CREMI-sized synthetic uint8 affinities 3x(125,1250,1250), block (25,250,250), context (3,31,31), "
                               "ws defaults, thresholds [0.2,0.35,0.5]", "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                         "sample": f"sub-volume {sample} of the workload ({int(np.prod(sample) / np.prod(BLOCK))} blocks, same "
                                   "block geometry), one process per block"},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bootstrapper_b200 import native
    from bootstrapper_b200.sharded import ShardedSegmenter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # pinned host buffers of the host-buffer leg should live on the NUMA node next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    slab, block, context = {5: (SHAPE5, BLOCK5, CONTEXT5), 4: (SHAPE4, BLOCK4, CONTEXT4)}.get(args.config, (SHAPE, BLOCK, CONTEXT))
    shape = (slab[0] * world, slab[1], slab[2]) if not args.quick else (50 * world, 500, 500)
    params = {"thresholds": THRESHOLDS}
    if args.config == 4:
        params.update(PARAMS4)
    seg = ShardedSegmenter(shape, block, context, params, rank=rank, world=world, device=dev)
    affs = seg.synth_local_affs(seed=0)                    # this rank's slab + z halo, generated on the device
    torch.cuda.synchronize()
    V_total = float(np.prod(shape))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    native.set_profiling(True)
    # the segmentations land in caller-provided buffers (as the C ABI has it), allocated once
    out = [torch.empty(seg.own_shape, dtype=torch.int64, device=dev) for _ in THRESHOLDS]
    for _ in range(args.warmup):
        seg.run(affs, out=out)
    barrier()
    l0 = native.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_acc = {}
    barrier()
    ev0.record()
    for _ in range(args.steps):
        seg.run(affs, out=out)
        for k, v in seg.last_profile.items():
            prof_acc[k] = prof_acc.get(k, 0.0) + v
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = native.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = V_total / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers (pinned), copies inside the timed region
    if args.config != 2:   # (69 GB of outputs per rank for config 5) the host-buffer leg is measured on the default workload only
        args.no_e2e = True
    if args.no_e2e:
        e2e_ms, h2d, d2h, e2e_stream = None, 0, 0, False
    else:
        e2e_ms, h2d, d2h, e2e_stream = run_e2e(args, seg, affs, barrier, dev, world, out)

    if rank == 0:
        report(args, seg, shape, slab, block, context, world, ms_step, value, V_total, prof_acc, launches, clocks, e2e_ms, h2d, d2h, e2e_stream)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, seg, affs, barrier, dev, world, out):
    """host-buffer leg through ShardedSegmenter.run_host in its streaming form: every step uploads the step's
    affinities from pinned memory and downloads fragments + all segmentations into pinned memory; the downloads of
    step k overlap the upload and compute of step k + 1 (two buffer sets); the timed region ends when the last
    download is through."""
    import torch
    import torch.distributed as dist
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_affs = torch.empty(affs.shape, dtype=affs.dtype, pin_memory=True)
    host_affs.copy_(affs)
    own_shape = seg.own_shape
    n_out = 1 + len(THRESHOLDS)
    host_sets = [[torch.empty(own_shape, dtype=torch.int64, pin_memory=True) for _ in range(n_out)]]
    dev_sets = [out]
    # the second buffer set doubles the page-locked host memory (6.25 GB per set and rank for the default workload): keep the
    # whole job at or below the 50 GB that one set per rank takes on 8 GPUs
    if world * 2 * n_out * int(np.prod(own_shape)) * 8 <= 50e9:
        try:
            host_sets.append([torch.empty(own_shape, dtype=torch.int64, pin_memory=True) for _ in range(n_out)])
            dev_sets.append([torch.empty_like(o) for o in out])
        except RuntimeError:       # not enough pinned / device memory for the second set: one volume in flight
            host_sets, dev_sets = host_sets[:1], dev_sets[:1]
    streaming = len(host_sets) == 2
    e2e_steps = max(2, min(args.steps, 8))
    for k in range(4):                                     # warm-up (touches both buffer sets, fills the allocator caches)
        seg.run_host(host_affs, host_sets[k % len(host_sets)], out=dev_sets[k % len(dev_sets)], wait=not streaming)
    seg.drain()
    barrier()
    ev0.record()
    for k in range(e2e_steps):
        seg.run_host(host_affs, host_sets[k % len(host_sets)], out=dev_sets[k % len(dev_sets)], wait=not streaming)
    seg.drain()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    h2d = host_affs.numel() * host_affs.element_size()
    d2h = sum(o.numel() * 8 for o in host_sets[0])
    return e2e_ms, h2d, d2h, streaming


def report(args, seg, shape, slab, block, context, world, ms_step, value, V_total, prof_acc, launches, clocks, e2e_ms, h2d, d2h, e2e_stream):
    peak, peak_src = measured_peak()
    steps = args.steps
    prof = {k: v / steps for k, v in prof_acc.items()}
    dom = max(prof, key=prof.get)
    # algorithmic bytes of one launch of the dominant kernel = 35 B/voxel x the voxels this rank's launch covers
    alg_bytes = BYTES_PER_VOXEL * float(np.prod(seg.own_shape))
    achieved = alg_bytes / (prof[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.config == 2 and not args.quick:   # the captures are of the default workload
        traffic = json.load(open(tp)).get(dom)
    cpu = None
    if not args.no_cpu and args.config == 2 and world == 1:   # the CPU leg samples the default workload, on rank 0 at N=1 only
        sample = (50, 500, 500) if args.quick else CPU_SAMPLE
        r, tm = cpu_oracle_rate(sample)
        cpu = {"value": r, "unit": "voxels/s", "cores": tm["workers"], "kind": "port",
               "sample": f"sub-volume {sample} of the workload ({tm['blocks']} blocks, same block geometry), "
                         f"one process per block, {tm['total']:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{'CREMI-sized ' if args.config == 2 else ''}synthetic uint8 affinities 3x{shape} ({world} z-slab(s) of {slab}), "
                               f"block {block}, context {context}, "
                               f"{'ws defaults' if args.config != 4 else 'fragments_in_xy=false, seed_eps=0.01'}, thresholds [0.2,0.35,0.5]",
                   "l2": f"inputs larger than L2 ({3 * int(np.prod(slab)) / 1e6:.0f} MB affinities, "
                         f"{32 * int(np.prod(slab)) / 1e9:.1f} GB outputs per step per GPU)",
                   "parity": "bit-exact vs oracle (seed_tie=index, stats_mode=canonical), tests/test_gpu_parity.py"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": prof[dom],
                     "whole_path_frac": (BYTES_PER_VOXEL * V_total / world / (ms_step * 1e-3) / 1e9) / peak},
        "stage_ms": {k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
        "cpu_baseline": cpu,
        "e2e": None if e2e_ms is None else {"value": V_total / (e2e_ms * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": h2d,
                                            "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                                            "mode": ("streaming: downloads of step k overlap upload + compute of step k+1, two buffer sets"
                                                     if e2e_stream else "one volume in flight")},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="small volume (smoke / CI), not a valid bench number")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="2 = BASELINE configs[1] (default, the metric's workload); 4 = 1024^3 3-D seeded + seed_eps over 4 GPUs (a 256x1024x1024 "
                         "slab per rank); 5 = 2048^3 over 8 GPUs (a 256x2048x2048 slab per rank)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
