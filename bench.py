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
PARITY_NOTE = ("bit-exact vs the CPU oracle (fragments incl. ids, nodes, RAG edges, f32 merge scores, LUTs, segmentations): "
               "tests/test_gpu_fullsize.py runs this workload at full size; parity unpinned for the third-party cores the oracle "
               "restates, declared deviations D1-D3 (DESIGN.md 4; D1 census profiles/r02_d1_census.json)")


def workload_string(config, shape, world, slab, block, context):
    return (f"{'CREMI-sized ' if config == 2 else ''}synthetic uint8 affinities 3x{tuple(shape)} ({world} z-slab(s) of {tuple(slab)}), "
            f"block {tuple(block)}, context {tuple(context)}, "
            f"{'ws defaults' if config != 4 else 'fragments_in_xy=false, seed_eps=0.01'}, thresholds [0.2,0.35,0.5]")


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


def sample_affs(shape, seed, in_process=True):
    """the synthetic input of a CPU leg (generation is outside every timed region).  in_process=False (the reference arm):
    a CHILD process runs the device generator and leaves the array under /dev/shm, so that the process that times the CPU
    implementation never maps libbsnative.so; without a GPU the numpy generator (bit-identical) runs on all cores."""
    if in_process:
        import torch
        from bootstrapper_b200 import native
        return native.synth_affs(shape, seed=seed).cpu().numpy()
    cache = f"/dev/shm/bs_bench_affs_{'x'.join(str(v) for v in shape)}_s{seed}.npy"
    if not os.path.exists(cache):
        code = ("import sys, numpy as np, torch\n"
                f"sys.path.insert(0, {ROOT!r})\n"
                "assert torch.cuda.is_available()\n"
                "from bootstrapper_b200 import native\n"
                f"np.save({cache!r} + '.tmp.npy', native.synth_affs({tuple(shape)!r}, seed={seed}).cpu().numpy())\n")
        rc = subprocess.run([sys.executable, "-c", code], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
        if rc == 0:
            os.replace(cache + ".tmp.npy", cache)
        else:
            import multiprocessing as mp
            from bootstrapper_b200.synth import synth_affs
            out = np.empty((3,) + tuple(shape), np.uint8)
            jobs = [(z, min(5, shape[0] - z)) for z in range(0, shape[0], 5)]
            with mp.get_context("fork").Pool(os.cpu_count()) as pool:
                for z, a in zip(jobs, pool.starmap(synth_affs, [((n, shape[1], shape[2]), seed, np.uint8, (z0, 0, 0), tuple(shape))
                                                               for z0, n in jobs])):
                    out[:, z[0]:z[0] + z[1]] = a
            np.save(cache, out)
            return out
    return np.load(cache)


def cpu_oracle_run(affs, workers=None):
    """the CPU oracle (one process per block) on `affs`; returns (voxels/s, timings, result)"""
    from oracle.parallel import waterz_pipeline_parallel
    tm = {}
    ref = waterz_pipeline_parallel(affs, {"thresholds": THRESHOLDS}, block_size=BLOCK, context=CONTEXT, timings=tm, workers=workers)
    return float(np.prod(affs.shape[1:])) / tm["total"], tm, ref


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (kind 'port'; the reference itself
    is pure Python over third-party packages that are not in the image), one process per block on all host
    cores, on the SAME volume as the GPU arm's N=1 workload.  Each step is one pass over the whole volume; the number of
    timed passes is bounded so that the run ends within a few minutes (steps_run says how many)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    shape = (50, 500, 500) if args.quick else SHAPE
    affs = sample_affs(shape, 0, in_process=False)
    if args.warmup > 0:
        cpu_oracle_run(np.ascontiguousarray(affs[:, :25, :250, :250]), cores)
    rates, ms = [], []
    t_start = time.time()
    for _ in range(max(1, args.steps)):
        r, tm, _ref = cpu_oracle_run(affs, cores)
        rates.append(r)
        ms.append(tm["total"] * 1e3)
        if time.time() - t_start + tm["total"] > 150.0:
            break
    v = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "steps_run": len(rates), "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_string(2, shape, 1, shape, BLOCK, CONTEXT), "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                         "sample": f"the whole volume {tuple(shape)} ({int(np.prod(shape) / np.prod(BLOCK))} blocks), one process per block, "
                                   f"{len(rates)} timed pass(es) of {np.mean(ms) / 1e3:.1f} s"},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def compare_with_oracle(r, ref):
    """GPU result dict (segment_blockwise / ShardedSegmenter.run layout) vs an oracle result: what matches, bit for bit"""
    f = r["fragments"].cpu().numpy().view(np.uint64)
    ok_f = bool(np.array_equal(f, ref["fragments"]))
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    got = np.stack([eu.view(np.uint64), ev.view(np.uint64)], 1)
    order = np.lexsort((got[:, 1], got[:, 0]))
    got, gs = got[order], es[order]
    keys = sorted(ref["rag"].edges)
    want = np.array(keys, dtype=np.uint64).reshape(-1, 2)
    ok_e = bool(np.array_equal(got, want))
    ok_s = False
    if ok_e:
        ws = np.array([np.nan if ref["rag"].edges[k] is None else ref["rag"].edges[k] for k in keys], dtype=np.float64)
        nan = np.isnan(ws)
        ok_s = bool(np.array_equal(np.isnan(gs), nan) and np.all(np.abs(gs[~nan] - ws[~nan]) <= 1e-6 * np.abs(ws[~nan])))
    ok_g = all(bool(np.array_equal(seg.cpu().numpy().view(np.uint64), ref["segs"][thr]["seg"])) for thr, seg in r["segs"].items())
    return {"fragments": ok_f, "edges": ok_e, "scores_1e-6": ok_s, "segs": ok_g,
            "n_fragments": int(len(ref["rag"].node_pos)), "n_edges": int(len(keys))}


# --config 3: BASELINE configs[2], mutex-watershed fragments from a 9-offset long-range neighbourhood at 512^3 on one GPU
MWS_NBH = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -9, 0], [0, 0, -9], [-3, 0, 0], [0, -27, 0], [0, 0, -27]]   # segment.py:24-34
MWS_BIAS = [-0.4] * 3 + [-0.7] * 6
MWS_STRIDES = [[1, 1, 1]] * 3 + [[2, 9, 9]] * 3 + [[3, 27, 27]] * 3


def mws_affs9(shape, seed=0):
    """nine-channel synthetic affinities on the device: the three nearest-neighbour channels of the block-addressable
    generator, and for every long-range offset the minimum of the nearest-neighbour affinity along the offset's path"""
    import torch
    from bootstrapper_b200 import native
    a = native.synth_affs(shape, seed=seed)
    out = [a[0], a[1], a[2]]
    for off in MWS_NBH[3:]:
        axis = [i for i, o in enumerate(off) if o][0]
        acc = a[axis].clone()
        for k in range(1, -off[axis]):
            sh = torch.zeros_like(acc)
            dst = [slice(None)] * 3
            src = [slice(None)] * 3
            dst[axis], src[axis] = slice(k, None), slice(0, -k)
            sh[tuple(dst)] = a[axis][tuple(src)]
            acc = torch.minimum(acc, sh)
        out.append(acc)
    return torch.stack(out).contiguous()


CONFIGS = {2: (SHAPE, BLOCK, CONTEXT, {}), 4: (SHAPE4, BLOCK4, CONTEXT4, PARAMS4), 5: (SHAPE5, BLOCK5, CONTEXT5, {})}


class Job:
    """process-wide state of one bench run: ranks, device, collectives"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # pinned host buffers of the host-buffer leg should live on the NUMA node next to this rank's GPU
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local_rank))
            except Exception:  # noqa: BLE001
                pass
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def time_workload(job, config, steps, warmup, quick=False, clocks=False):
    """device-timed steps of one configuration on all ranks (CUDA events on the launching stream, barrier + synchronize
    on both sides, max over ranks).  Returns the segmenter, its resident inputs / outputs and the measurements."""
    import torch
    from bootstrapper_b200 import native
    from bootstrapper_b200.sharded import ShardedSegmenter
    slab, block, context, extra = CONFIGS[config]
    if config == 4 and job.world in (1, 2, 4):
        slab = (1024 // max(job.world, 2) if job.world > 1 else 256, slab[1], slab[2])   # the 1024^3 volume over 2 / 4 ranks
    shape = (slab[0] * job.world, slab[1], slab[2]) if not quick else (50 * job.world, 500, 500)
    params = dict({"thresholds": THRESHOLDS}, **extra)
    seg = ShardedSegmenter(shape, block, context, params, rank=job.rank, world=job.world, device=job.dev)
    affs = seg.synth_local_affs(seed=0)                    # this rank's slab + z halo, generated on the device
    torch.cuda.synchronize()
    native.set_profiling(True)
    # the segmentations land in caller-provided buffers (as the C ABI has it), allocated once
    out = [torch.empty(seg.own_shape, dtype=torch.int64, device=job.dev) for _ in THRESHOLDS]
    for _ in range(warmup):
        seg.run(affs, out=out)
    job.barrier()
    l0 = native.launch_count()
    sampler = ClockSampler(job.local_rank) if clocks and job.rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_acc = {}
    job.barrier()
    ev0.record()
    for _ in range(steps):
        seg.run(affs, out=out)
        for k, v in seg.last_profile.items():
            prof_acc[k] = prof_acc.get(k, 0.0) + v
    ev1.record()
    job.barrier()
    ms_step = job.max_over_ranks(ev0.elapsed_time(ev1)) / steps
    res = dict(config=config, shape=shape, slab=slab if not quick else (50, 500, 500), block=block, context=context, seg=seg, affs=affs,
               out=out, ms_step=ms_step, value=float(np.prod(shape)) / (ms_step * 1e-3),
               stage_ms={k: v / steps for k, v in prof_acc.items()}, launches=int(native.launch_count() - l0),
               clocks=sampler.stop() if sampler else None)
    return res


def run_pipelined(job, res, steps, nthreads=2):
    """whole-job throughput with `nthreads` independent volumes in flight on one GPU: each host thread drives its own segmenter
    (own plan, own output buffers, own stream; the library's scratch is per host thread) over the same resident input, so the
    bandwidth-bound passes of one volume fill the issue slots the latency-bound flood of the other leaves idle.  Host clock
    around device-synchronised begin / end; every thread runs `steps` volumes after one warm-up volume."""
    import threading
    import time as _time
    import torch
    from bootstrapper_b200.sharded import ShardedSegmenter
    seg0 = res["seg"]
    params = dict({"thresholds": THRESHOLDS}, **CONFIGS[res["config"]][3])
    segs = [seg0] + [ShardedSegmenter(res["shape"], res["block"], res["context"], params, rank=0, world=1, device=job.dev) for _ in range(nthreads - 1)]
    outs = [res["out"]] + [[torch.empty_like(o) for o in res["out"]] for _ in range(nthreads - 1)]
    affs = res["affs"]
    errs = []
    gate = threading.Barrier(nthreads + 1)

    def work(k):
        try:
            st = torch.cuda.Stream(device=job.dev)
            with torch.cuda.stream(st):
                segs[k].run(affs, out=outs[k])
                st.synchronize()
                gate.wait()            # warm-up done everywhere
                gate.wait()            # clock started
                for _ in range(steps):
                    segs[k].run(affs, out=outs[k])
                st.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
            gate.abort()

    th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    for t in th:
        t.start()
    try:
        gate.wait()
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        gate.wait()
    except threading.BrokenBarrierError:
        pass
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = _time.perf_counter() - t0
    if errs:
        return {"error": errs[0]}
    same = all(bool(torch.equal(a, b)) for o in outs[1:] for a, b in zip(o, outs[0]))
    ms = dt * 1e3 / (steps * nthreads)
    return {"volumes_in_flight": nthreads, "ms_per_volume": ms, "value": float(np.prod(res["shape"])) / (ms * 1e-3), "unit": "voxels/s",
            "volumes": steps * nthreads, "results_identical_across_threads": same,
            "what": "device-resident throughput with two independent volumes in flight (two host threads, two plans, two streams); "
                    "`value` above is one volume at a time"}


def roofline_of(res, traffic_ok):
    peak, peak_src = measured_peak()
    prof = res["stage_ms"]
    dom = max(prof, key=prof.get)
    # algorithmic bytes of one launch of the dominant kernel = 35 B/voxel x the voxels this rank's launch covers
    alg_bytes = BYTES_PER_VOXEL * float(np.prod(res["seg"].own_shape))
    achieved = alg_bytes / (prof[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if traffic_ok and os.path.exists(tp):   # the captures are of the default workload
        traffic = json.load(open(tp)).get(dom)
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": prof[dom],
            "whole_path_frac": (alg_bytes / (res["ms_step"] * 1e-3) / 1e9) / peak}


def run_e2e(job, res, steps):
    """host-buffer leg through ShardedSegmenter.run_host in its streaming form: every step uploads the step's affinities
    from pinned memory and downloads fragments + all segmentations into pinned memory (a bounded ring of z-chunk staging
    buffers, < 2 GB page-locked per rank, drained by a downloader thread); the downloads of step k overlap the upload and
    compute of step k + 1 (two device buffer sets); the timed region ends when the last chunk of the last step has landed."""
    import torch
    from bootstrapper_b200.sharded import HostRing
    seg, affs, out = res["seg"], res["affs"], res["out"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_affs = torch.empty(affs.shape, dtype=affs.dtype, pin_memory=True)
    host_affs.copy_(affs)
    n_out = 1 + len(THRESHOLDS)
    ring = HostRing(seg.own_shape, device=job.dev)
    dev_sets = [out, [torch.empty_like(o) for o in out]]
    e2e_steps = max(2, min(steps, 8))
    for k in range(3):                                     # warm-up (touches both buffer sets, fills the allocator caches)
        seg.run_host(host_affs, None, out=dev_sets[k % 2], wait=False, ring=ring)
    seg.drain()
    job.barrier()
    ev0.record()
    for k in range(e2e_steps):
        seg.run_host(host_affs, None, out=dev_sets[k % 2], wait=False, ring=ring)
    seg.drain()
    ev1.record()
    job.barrier()
    e2e_ms = job.max_over_ranks(ev0.elapsed_time(ev1)) / e2e_steps
    h2d = host_affs.numel() * host_affs.element_size()
    d2h = n_out * int(np.prod(seg.own_shape)) * 8
    info = dict(ms=e2e_ms, h2d=h2d, d2h=d2h, pinned_bytes=ring.pinned_bytes + h2d, chunks=ring.chunks_done,
                mode=f"streaming: downloads of step k overlap upload + compute of step k+1; results land in a ring of {ring.n_slots} "
                     f"page-locked z-chunk buffers ({ring.pinned_bytes / 1e9:.2f} GB per rank) drained by a downloader thread")
    ring.close()
    return info


def run_e2e_compact(job, res, steps):
    """the host-buffer leg in the compact result form (include/bsnative.h): per step the affinities go up, and ONE int32 plane
    of dense fragment numbers + the node-id table + a LUT row per threshold come down (4 bytes per voxel instead of 32);
    the host decoder (bs_expand_compact, all host threads of this rank) is timed separately, outside the streamed region."""
    import torch
    from bootstrapper_b200 import native
    seg, affs = res["seg"], res["affs"]
    host_affs = torch.empty(affs.shape, dtype=affs.dtype, pin_memory=True)
    host_affs.copy_(affs)
    T = len(THRESHOLDS)
    cap = int(np.prod(seg.vol_shape)) // 256 + 4096          # node-table capacity (the run raises if it is too small)
    sets = [dict(dense=torch.empty(seg.own_shape, dtype=torch.int32, pin_memory=True), nodes=torch.empty(cap, dtype=torch.int64, pin_memory=True),
                 luts=[torch.empty(cap, dtype=torch.int64, pin_memory=True) for _ in range(T)]) for _ in range(2)]
    e_steps = max(2, min(steps, 8))
    info = None
    for k in range(3):
        info = seg.run_host_compact(host_affs, sets[k % 2], wait=False)
    seg.drain()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.barrier()
    ev0.record()
    for k in range(e_steps):
        info = seg.run_host_compact(host_affs, sets[k % 2], wait=False)
    seg.drain()
    ev1.record()
    job.barrier()
    ms = job.max_over_ranks(ev0.elapsed_time(ev1)) / e_steps
    n = info["n_nodes"]
    d2h = int(np.prod(seg.own_shape)) * 4 + n * 8 * (1 + T)
    # the decoder, on this rank's share of the host cores
    threads = max(1, (os.cpu_count() or 1) // job.world)
    hs = sets[(e_steps - 1) % 2]
    outs = [torch.zeros(seg.own_shape, dtype=torch.int64) for _ in range(1 + T)]    # touched: no first-use page faults in the timing
    job.barrier()
    t0 = time.time()
    native.expand_compact(hs["dense"], hs["nodes"][:n].contiguous(), [l[:n].contiguous() for l in hs["luts"]], outs[0], outs[1:], threads=threads)
    job.barrier()
    expand_ms = job.max_over_ranks((time.time() - t0) * 1e3)
    return dict(ms=ms, d2h=d2h, h2d=host_affs.numel(), expand_ms=expand_ms, threads=threads, n_nodes=n)


def run_e2e_expanded(job, res, steps):
    """host-buffer leg that delivers the SAME uint64 arrays in host memory as run_e2e, by the compact route: per step the
    affinities go up, 4 bytes per voxel + tables come down, and a decoder thread (bs_expand_compact on this rank's share of the
    host cores) rebuilds fragments + all segmentations in host memory while the device works on the next volume.  Host clock
    from a synchronised start to the last decoded volume."""
    import torch
    from bootstrapper_b200.sharded import HostExpander
    seg, affs = res["seg"], res["affs"]
    host_affs = torch.empty(affs.shape, dtype=affs.dtype, pin_memory=True)
    host_affs.copy_(affs)
    T = len(THRESHOLDS)
    cap = int(np.prod(seg.vol_shape)) // 256 + 4096
    sets = [dict(dense=torch.empty(seg.own_shape, dtype=torch.int32, pin_memory=True), nodes=torch.empty(cap, dtype=torch.int64, pin_memory=True),
                 luts=[torch.empty(cap, dtype=torch.int64, pin_memory=True) for _ in range(T)]) for _ in range(2)]
    outs = [torch.zeros(seg.own_shape, dtype=torch.int64) for _ in range(1 + T)]
    # one core of the rank's share stays with the thread that drives the device
    threads = max(1, (os.cpu_count() or 1) // job.world - 1)
    exp = HostExpander(threads)
    e_steps = max(2, min(steps, 8))
    info = None

    def step(k):
        exp.acquire()
        i = seg.run_host_compact(host_affs, sets[k % 2], wait=False)
        exp.submit(i["done"], sets[k % 2], i["n_nodes"], outs)
        return i
    for k in range(3):
        info = step(k)
    exp.flush()
    seg.drain()
    job.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e_steps):
        info = step(k + 1)
    exp.flush()
    seg.drain()
    ms = job.max_over_ranks((time.perf_counter() - t0) * 1e3) / e_steps
    job.barrier()
    exp.close()
    # the decoded arrays are the device results
    ok = all(bool(torch.equal(o, d.cpu())) for o, d in zip(outs[1:], res["out"]))      # the timed run's device segmentations, same input
    n = info["n_nodes"]
    return dict(ms=ms, d2h=int(np.prod(seg.own_shape)) * 4 + n * 8 * (1 + T), h2d=host_affs.numel(), threads=threads, checked=ok)


def multi_gpu_parity(job):
    """N > 1: a small volume through the sharded path on all ranks, then rank 0 runs the same volume alone (single-GPU
    path, bit-exact vs the oracle in the tests) and compares its own slab + the global graph.  Outside every timed region."""
    import torch
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from bootstrapper_b200.sharded import ShardedSegmenter
    shape, block, context = (20 * job.world, 500, 500), (10, 250, 250), (2, 31, 31)
    params = {"thresholds": THRESHOLDS}
    seg = ShardedSegmenter(shape, block, context, params, rank=job.rank, world=job.world, device=job.dev)
    r = seg.run(seg.synth_local_affs(seed=3))
    job.barrier()
    if job.rank != 0:
        return None
    full = segment_blockwise(native.synth_affs(shape, seed=3, device=job.dev), params, block, context)
    g = seg.geo
    own = slice(g["z0"], g["z1"])
    ok_f = bool(torch.equal(r["own_fragments"], full["fragments"][own]))
    def canon(e):
        u, v, sc = [t.cpu().numpy() for t in e]
        o = np.lexsort((v, u))
        return u[o], v[o], sc[o].view(np.int32)
    a, b = canon(r["edges"]), canon(full["edges"])
    ok_e = all(x.shape == y.shape and bool(np.array_equal(x, y)) for x, y in zip(a, b))
    ok_s = all(bool(torch.equal(r["segs"][t], full["segs"][t][own])) for t in THRESHOLDS)
    return {"volume": f"3x{shape}, block {block}, context {context}, {job.world} slabs vs one GPU", "fragments": ok_f, "edges": ok_e,
            "segs": ok_s}


def run_ours(args):
    import torch
    from bootstrapper_b200 import native
    job = Job()
    res = time_workload(job, args.config, args.steps, args.warmup, quick=args.quick, clocks=True)
    pipelined = None
    if job.world == 1 and args.config == 2 and not args.quick and not args.no_extra:
        pipelined = run_pipelined(job, res, max(3, args.steps))
        torch.cuda.empty_cache()
    # ---- end to end through the public API with host buffers (pinned), copies inside the timed region
    e2e = None
    if args.config == 2 and not args.no_e2e:   # (69 GB of outputs per rank for config 5) measured on the default workload only
        e2e = run_e2e(job, res, args.steps)
    e2e_c = None
    if args.config == 2 and not args.no_e2e:
        e2e_c = run_e2e_compact(job, res, args.steps)
    e2e_x = None
    if args.config == 2 and not args.no_e2e:
        e2e_x = run_e2e_expanded(job, res, args.steps)
    line = None
    if job.rank == 0:
        line = report(args, job, res, e2e)
        if e2e_x is not None:
            V_total = float(np.prod(res["shape"]))
            xline = {
                "value": V_total / (e2e_x["ms"] * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_x["ms"], "h2d_bytes_per_step": e2e_x["h2d"],
                "d2h_bytes_per_step": e2e_x["d2h"], "host_decode_threads_per_rank": e2e_x["threads"], "decoded_equals_device_result": e2e_x["checked"],
                "what": "the same uint64 fragments + segmentations in host memory as `e2e`, delivered by the compact route: 4 bytes per voxel "
                        "+ tables cross the bus, a decoder thread (bs_expand_compact) rebuilds the arrays in host memory while the device works "
                        "on the next volume; host clock from a synchronised start to the last decoded volume, decode INSIDE the timed region"}
            # `e2e` = the faster of the two routes that end with the full uint64 arrays in host memory; the other one is kept beside it
            if line["e2e"] is not None and xline["ms_per_step"] < line["e2e"]["ms_per_step"]:
                line["e2e_streamed"] = line["e2e"]
                line["e2e"] = dict(xline, mode="compact transfer + overlapped host decode (ShardedSegmenter.run_host_compact + HostExpander); "
                                               "e2e_streamed is the plain download of the uint64 arrays")
            else:
                line["e2e_expanded"] = xline
        if pipelined is not None:
            line["pipelined"] = pipelined
        if e2e_c is not None:
            V_total = float(np.prod(res["shape"]))
            line["e2e_compact"] = {
                "value": V_total / (e2e_c["ms"] * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_c["ms"], "h2d_bytes_per_step": e2e_c["h2d"],
                "d2h_bytes_per_step": e2e_c["d2h"], "host_decode_ms_per_volume": e2e_c["expand_ms"], "host_decode_threads_per_rank": e2e_c["threads"],
                "what": "the same streamed host-buffer leg with the results in the library's compact form: one int32 plane of dense fragment "
                        "numbers + node-id table + one LUT row per threshold (frags[i] = node_ids[d[i]-1], seg_t[i] = lut_t[d[i]-1]); "
                        "host_decode_ms = bs_expand_compact rebuilding the four uint64 arrays on the host, all ranks at once, timed "
                        "separately (not inside ms_per_step)"}
    # ---- parity, outside the timed regions: CPU oracle vs GPU on the cpu_baseline sample (N=1), sharded vs single GPU (N>1)
    parity = {}
    if job.world > 1 and not args.no_parity:
        mg = multi_gpu_parity(job)
        if mg is not None:
            parity["multi_gpu"] = mg
    if job.world == 1 and not args.no_cpu and args.config == 2:
        from bootstrapper_b200.post.pipeline import segment_blockwise
        sample = (50, 500, 500) if args.quick else CPU_SAMPLE
        saffs = native.synth_affs(sample, seed=0, device=job.dev)
        rate, tm, ref = cpu_oracle_run(saffs.cpu().numpy())
        got = segment_blockwise(saffs, {"thresholds": THRESHOLDS}, BLOCK, CONTEXT)
        parity["cpu_oracle"] = dict(compare_with_oracle(got, ref), sample=f"sub-volume {sample} ({tm['blocks']} blocks)")
        line["cpu_baseline"] = {"value": rate, "unit": "voxels/s", "cores": tm["workers"], "kind": "port",
                                "sample": f"sub-volume {sample} of the workload ({tm['blocks']} blocks, same block geometry), "
                                          f"one process per block, {tm['total']:.1f} s"}
        del got, saffs, ref
    # ---- the BASELINE multi-GPU target configurations, driver-run: config 5 on 8 GPUs, config 4 on 4 / 2
    extra = {}
    want = {8: [5], 4: [4], 2: [4]}.get(job.world, []) if (args.config == 2 and not args.quick and not args.no_extra) else []
    if want:
        del res["seg"], res["affs"], res["out"]
        res = None
        torch.cuda.empty_cache()
        native.release_scratch()
    for c in want:
        r2 = time_workload(job, c, 3, 2)
        if job.rank == 0:
            rf = roofline_of(r2, False)
            extra[f"config{c}"] = {"workload": workload_string(c, r2["shape"], job.world, r2["slab"], r2["block"], r2["context"]),
                                   "ms_per_step": r2["ms_step"], "value": r2["value"], "unit": "voxels/s", "steps": 3, "warmup": 2,
                                   "whole_path_frac": rf["whole_path_frac"], "dominant": rf["kernel"],
                                   "stage_ms": {k: round(v, 3) for k, v in sorted(r2["stage_ms"].items(), key=lambda kv: -kv[1])},
                                   "gpu_launches": r2["launches"]}
        del r2
        torch.cuda.empty_cache()
        native.release_scratch()
    if job.world == 1 and args.config == 2 and not args.quick and not args.no_extra:
        # BASELINE configs[2]: mutex-watershed fragments from the 9-offset default neighbourhood at 512^3 on one GPU
        if res is not None:
            del res["seg"], res["affs"], res["out"]
            res = None
        torch.cuda.empty_cache()
        native.release_scratch()
        kw = dict(strides=MWS_STRIDES, noise_eps=0.001, noise_seed=0)
        a9 = mws_affs9((512, 512, 512), seed=0)
        native.mws_agglom(a9, MWS_NBH, MWS_BIAS, **kw)          # warm-up at full size: the first call grows the memory pool by ~20 GB
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        frags9, _, cnt9 = native.mws_agglom(a9, MWS_NBH, MWS_BIAS, **kw)
        ev[1].record()
        torch.cuda.synchronize()
        ms9 = ev[0].elapsed_time(ev[1])
        extra["config3"] = {"workload": "mutex-watershed fragments (bs segment --mws, simple_mutex), 9x(512,512,512) uint8 affinities, offsets / biases / "
                                        "strides = the reference defaults (segment.py:24-51), seeded noise 0.001, one GPU",
                            "ms_per_step": ms9, "value": 512.0 ** 3 / (ms9 * 1e-3), "unit": "voxels/s", "steps": 1, "warmup": 1,
                            "counters": {k: int(v) for k, v in cnt9.items()},
                            "parity": "bit-identical to the sequential restatement of mwatershed.agglom at oracle-sized cases "
                                      "(tests/test_gpu_mws.py; declared tie rule D4, parity unpinned)"}
        del frags9
        # the same volume through the blockwise pipeline with SURVEY 8(d)'s config-3 geometry: 128^3 blocks, context 16
        from bootstrapper_b200.post.pipeline import segment_mws_blockwise
        torch.cuda.empty_cache()
        p9 = dict(aff_neighborhood=MWS_NBH, bias=MWS_BIAS, strides=MWS_STRIDES, noise_eps=0.001, noise_seed=0)
        torch.cuda.synchronize()
        ev[0].record()
        rb = segment_mws_blockwise(a9, p9, (128, 128, 128), (16, 16, 16), profile=True)
        ev[1].record()
        torch.cuda.synchronize()
        msb = ev[0].elapsed_time(ev[1])
        extra["config3_blockwise"] = {
            "workload": "blockwise mws pipeline (bs segment --mws -b: ExtractFrags -> AffAgglom -> GraphMWS -> Relabel), 9x(512,512,512) uint8 "
                        "affinities, 64 blocks of (128,128,128), context (16,16,16), reference default offsets / biases / strides, seeded noise "
                        "0.001, global_bias [1.0, -0.5], one GPU",
            "ms_per_step": msb, "value": 512.0 ** 3 / (msb * 1e-3), "unit": "voxels/s", "steps": 1, "warmup": 0,
            "stage_s": {k: round(v, 3) for k, v in rb["stage_s"].items()},
            "fragments": int(rb["nodes"][0].numel()), "edges": int(rb["edges"][0].numel()),
            "counters": {"extract_frags": [{k: int(v) for k, v in c.items()} for c in rb["counters"]["extract_frags"]],
                         "graph_mws": {k: int(v) for k, v in rb["counters"]["graph_mws"].items()}},
            "parity": "bit-identical to the oracle's restatement of the four volara tasks at oracle-sized cases (tests/test_gpu_mws.py; parity unpinned)"}
        del a9, rb
        torch.cuda.empty_cache()
        native.release_scratch()
    if job.rank == 0:
        line["parity_checked"] = parity or None
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if job.world > 1:
        job.dist.destroy_process_group()


def report(args, job, res, e2e):
    world = job.world
    V_total = float(np.prod(res["shape"]))
    line = {
        "metric": METRIC, "value": res["value"], "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_string(args.config, res["shape"], world, res["slab"], res["block"], res["context"]),
                   "l2": f"inputs larger than L2 ({3 * int(np.prod(res['slab'])) / 1e6:.0f} MB affinities, "
                         f"{32 * int(np.prod(res['slab'])) / 1e9:.1f} GB outputs per step per GPU)",
                   "parity": PARITY_NOTE},
        "roofline": roofline_of(res, args.config == 2 and not args.quick),
        "stage_ms": {k: round(v, 3) for k, v in sorted(res["stage_ms"].items(), key=lambda kv: -kv[1])},
        "cpu_baseline": None,
        "e2e": None if e2e is None else {"value": V_total / (e2e["ms"] * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": e2e["h2d"],
                                         "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["ms"], "mode": e2e["mode"],
                                         "pinned_bytes_per_rank": e2e["pinned_bytes"]},
        "gpu_launches": res["launches"], "clocks": res["clocks"],
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="small volume (smoke / CI), not a valid bench number")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and its parity comparison)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra BASELINE configurations (config 3 at 1 GPU, config 5 at 8 GPUs, config 4 at 4 / 2)")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity comparison")
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
