"""One rank's share of BASELINE config 5 (2048^3 on 8 GPUs = a (256, 2048, 2048) slab per GPU: 64 blocks of 256^3,
context 32) on a single GPU: time, memory and self-consistency properties (the oracle is far too slow at this size).
    python tests/gpu_slab5.py [--z 256] [--yx 2048]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native  # noqa: E402
from bootstrapper_b200.sharded import ShardedSegmenter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--z", type=int, default=256)
    ap.add_argument("--yx", type=int, default=2048)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--params", default="{}", help='ws_params as JSON, e.g. \'{"fragments_in_xy": false, "seed_eps": 0.01}\' (config 4)')
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shape = (a.z, a.yx, a.yx)
    block, ctx = (a.block,) * 3, (a.block // 8,) * 3
    seg = ShardedSegmenter(shape, block, ctx, json.loads(a.params), rank=0, world=1, device=dev)
    t0 = time.time()
    affs = seg.synth_local_affs(seed=0)
    torch.cuda.synchronize()
    print("synth %.2fs  %.2f Gvox" % (time.time() - t0, np.prod(shape) / 1e9), flush=True)
    native.set_profiling(True)
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.time()
        r = seg.run(affs)
        torch.cuda.synchronize()
        dt = time.time() - t0
        free, total = torch.cuda.mem_get_info()
        print("rep %d: %.3f s -> %.3f Gvox/s; nodes %d edges %d; device mem used %.1f GB" % (
            rep, dt, np.prod(shape) / dt / 1e9, r["nodes"].numel(), r["edges"][0].numel(), (total - free) / 1e9), flush=True)
        print("   ", {k: round(v, 1) for k, v in sorted(seg.last_profile.items(), key=lambda kv: -kv[1])}, flush=True)
    # properties: segment ids are fragment ids, coarser thresholds only merge, background preserved
    f = r["own_fragments"]
    prev = f
    for thr in sorted(r["segs"]):
        sg = r["segs"][thr]
        assert bool(((sg == 0) == (f == 0)).all()), "background changed"
        pairs = torch.unique(torch.stack([prev.flatten()[::97], sg.flatten()[::97]], 1), dim=0)
        assert pairs[:, 0].unique().numel() == pairs.shape[0], "a finer segment maps to two coarser ones"
        prev = sg
    print("properties OK")


if __name__ == "__main__":
    main()
