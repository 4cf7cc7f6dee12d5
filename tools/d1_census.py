#!/usr/bin/env python
"""Census of declared deviation D1 at BASELINE config-2 scale, on the CPU oracle only.

D1: skimage pushes all seeds with age 0, so ties among equal-valued seeds follow the binary heap's layout history
(oracle seed_tie="heap", the faithful restatement); the CUDA path lets seeds enter their level's FIFO in ascending
raveled index (seed_tie="index").  This script runs the whole blockwise pipeline both ways on the synthetic config-2
volume 3x(125,1250,1250) u8, block (25,250,250), context (3,31,31), and counts, label-permutation invariant:
  * voxels whose fragment differs (voxels outside the best-overlap partner of their fragment),
  * fragments / RAG edges that exist on one side only,
  * per threshold: voxels whose segment differs, and segments without an identical voxel set on the other side.

    python tools/d1_census.py [--shape 125 1250 1250] [--out profiles/r02_d1_census.json]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BLOCK, CONTEXT, THRESHOLDS = (25, 250, 250), (3, 31, 31), [0.2, 0.35, 0.5]


def _gen(args):
    from bootstrapper_b200.synth import synth_affs
    z0, nz, shape, seed = args
    return z0, synth_affs((nz, shape[1], shape[2]), seed=seed, offset=(z0, 0, 0), vol_shape=shape)


def make_volume(shape, seed, cache):
    if cache and os.path.exists(cache):
        a = np.load(cache, mmap_mode="r")
        if a.shape == (3,) + tuple(shape):
            return np.ascontiguousarray(a)
    out = np.empty((3,) + tuple(shape), np.uint8)
    jobs = [(z, min(5, shape[0] - z), shape, seed) for z in range(0, shape[0], 5)]
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        for z0, a in pool.imap_unordered(_gen, jobs):
            out[:, z0:z0 + a.shape[1]] = a
    if cache:
        np.save(cache, out)
    return out


def partition_diff(a, b):
    """voxels of `a`'s non-zero classes that fall outside the best-overlap class of `b` (and the zero/non-zero mismatch);
    number of a-classes with an identical voxel set in b."""
    a = a.ravel()
    b = b.ravel()
    zero_mismatch = int(np.count_nonzero((a == 0) != (b == 0)))
    m = (a != 0) & (b != 0)
    ua, ia = np.unique(a[m], return_inverse=True)
    ub, ib = np.unique(b[m], return_inverse=True)
    pair = ia.astype(np.int64) * len(ub) + ib
    up, cnt = np.unique(pair, return_counts=True)
    pa, pb = up // len(ub), up % len(ub)
    best = np.zeros(len(ua), np.int64)
    np.maximum.at(best, pa, cnt)
    moved = int(cnt.sum() - best.sum())
    # identical classes: a pair whose count equals both class sizes (sizes over the whole arrays)
    sa = np.bincount(ia, minlength=len(ua))
    sb = np.bincount(ib, minlength=len(ub))
    # class sizes including voxels where the other side is zero
    ta = dict(zip(*np.unique(a[a != 0], return_counts=True)))
    tb = dict(zip(*np.unique(b[b != 0], return_counts=True)))
    same = 0
    for i, j, c in zip(pa, pb, cnt):
        if c == sa[i] == sb[j] and ta[ua[i]] == c and tb[ub[j]] == c:
            same += 1
    return dict(voxels_differ=moved + zero_mismatch, classes_a=len(ta), classes_b=len(tb), classes_identical=same)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[125, 1250, 1250])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cache", default="/dev/shm/bs_census_affs.npy")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_d1_census.json"))
    args = ap.parse_args()
    shape = tuple(args.shape)
    from oracle.parallel import waterz_pipeline_parallel
    t0 = time.time()
    affs = make_volume(shape, args.seed, args.cache)
    print("volume", affs.shape, "%.0f s" % (time.time() - t0), flush=True)
    res = {}
    for tie in ("heap", "index"):
        t0 = time.time()
        res[tie] = waterz_pipeline_parallel(affs, {"thresholds": THRESHOLDS}, block_size=BLOCK, context=CONTEXT, seed_tie=tie,
                                            stats_mode="canonical")
        print(tie, "%.0f s" % (time.time() - t0), flush=True)
    V = int(np.prod(shape))
    fa, fb = res["heap"]["fragments"], res["index"]["fragments"]
    ea, eb = set(res["heap"]["rag"].edges), set(res["index"]["rag"].edges)
    out = {
        "what": "D1 census: oracle seed_tie='heap' (faithful skimage restatement) vs 'index' (what the CUDA path computes), "
                "everything else equal (stats_mode='canonical')",
        "workload": f"synthetic uint8 affinities 3x{shape} seed {args.seed}, block {BLOCK}, context {CONTEXT}, ws defaults, thresholds {THRESHOLDS}",
        "voxels": V,
        "fragments": partition_diff(fa, fb),
        "fragment_arrays_identical": bool(np.array_equal(fa, fb)),
        "rag_edges": {"heap": len(ea), "index": len(eb), "only_heap": len(ea - eb), "only_index": len(eb - ea)},
        "segments": {},
    }
    out["fragments"]["voxel_fraction"] = out["fragments"]["voxels_differ"] / V
    for thr in THRESHOLDS:
        d = partition_diff(res["heap"]["segs"][thr]["seg"], res["index"]["segs"][thr]["seg"])
        d["voxel_fraction"] = d["voxels_differ"] / V
        d["segment_fraction_not_identical"] = 1.0 - d["classes_identical"] / max(1, d["classes_a"])
        out["segments"][str(thr)] = d
    print(json.dumps(out, indent=1))
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
