"""Full-size run of the CUDA pipeline with per-stage device timings; optional full-size parity
against the (multi-process) oracle.  Run on a GPU box:
    python tests/gpu_scale.py --shape 125 1250 1250 --block 25 250 250 --context 3 31 31 [--check]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native  # noqa: E402
from bootstrapper_b200.post.pipeline import segment_blockwise  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[125, 1250, 1250])
    ap.add_argument("--block", type=int, nargs=3, default=[25, 250, 250])
    ap.add_argument("--context", type=int, nargs=3, default=[3, 31, 31])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--xy", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shape, block, ctx = tuple(a.shape), tuple(a.block), tuple(a.context)
    t0 = time.time()
    affs = native.synth_affs(shape, seed=a.seed, device=dev)
    torch.cuda.synchronize()
    print("synth %.2fs" % (time.time() - t0), flush=True)
    params = dict(fragments_in_xy=bool(a.xy))
    native.set_profiling(True)
    V = int(np.prod(shape))
    for rep in range(a.reps):
        torch.cuda.reset_peak_memory_stats()
        torch.cuda.synchronize()
        t0 = time.time()
        plan, p = None, None
        from bootstrapper_b200.post.pipeline import make_plan
        plan, p = make_plan(affs, params, block, ctx)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        frags = plan.fragments(affs)
        prof1 = native.get_profile()
        e[1].record()
        ids, pos, sz = plan.nodes(dev)
        plan.agglomerate(affs, frags)
        prof2 = native.get_profile()
        e[2].record()
        eu, ev, es = plan.edges(dev)
        segs = {}
        for thr in p["thresholds"]:
            comp = native.connected_components(ids, eu, ev, es, float(thr))
            segs[thr] = native.relabel(frags, ids, comp)
        e[3].record()
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
        print("rep %d: wall %.3fs  stage1 %.1f ms  stage2 %.1f ms  stage3 %.1f ms  -> %.3f Gvox/s  nodes %d edges %d" % (
            rep, wall, ms[0], ms[1], ms[2], V / wall / 1e9, ids.numel(), eu.numel()), flush=True)
        print("   s1:", {k: round(v, 2) for k, v in prof1.items()})
        print("   s2:", {k: round(v, 2) for k, v in prof2.items()})
        free, total = torch.cuda.mem_get_info()
        print("   mem: torch peak %.2f GB, device used now %.2f GB" % (torch.cuda.max_memory_allocated() / 1e9, (total - free) / 1e9))
    if a.check:
        from oracle.parallel import waterz_pipeline_parallel
        tm = {}
        haffs = affs.cpu().numpy()
        ref = waterz_pipeline_parallel(haffs, params, block_size=block, context=ctx, timings=tm)
        print("oracle:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in tm.items()}, "-> %.4f Gvox/s" % (V / tm["total"] / 1e9))
        f = frags.cpu().numpy().view(np.uint64)
        print("fragments mismatches:", int((f != ref["fragments"]).sum()))
        got = dict(zip(zip(eu.cpu().numpy().view(np.uint64).tolist(), ev.cpu().numpy().view(np.uint64).tolist()), es.cpu().numpy().tolist()))
        want = ref["rag"].edges
        print("edge sets equal:", set(got) == set(want), len(got), len(want))
        bad = 0
        for k, s in want.items():
            g = got.get(k)
            if g is None:
                bad += 1
            elif s is None:
                bad += not np.isnan(g)
            else:
                bad += not (abs(g - s) <= 1e-6 * abs(s))
        print("edge score mismatches:", bad)
        for thr in p["thresholds"]:
            print("seg", thr, "mismatches:", int((segs[thr].cpu().numpy().view(np.uint64) != ref["segs"][thr]["seg"]).sum()))


if __name__ == "__main__":
    main()
