"""GPU-side debugging aid (not a test): blockwise mws pipeline vs oracle, stage by stage"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_mws import BW_CASES, _affs9  # noqa: E402
from bootstrapper_b200.post.pipeline import segment_mws_blockwise  # noqa: E402
from oracle import mws as om  # noqa: E402

for ci, (shape, block, ctx, params, dtype, use_mask) in enumerate(BW_CASES):
    nbh = params["aff_neighborhood"]
    full = _affs9(shape, seed=11, dtype=dtype)
    affs = np.ascontiguousarray(full[:len(nbh)])
    if shape == (12, 48, 48):
        affs[:, 0:8, 0:30, 0:30] = 0
    mask = None
    if use_mask:
        mask = np.ones(shape, np.uint8)
        mask[:, 8:16, 4:30] = 0
    ref = om.volara_pipeline(affs, params, shape if block is None else block, (0, 0, 0) if block is None else ctx, mask=mask,
                             noise_seed=params.get("noise_seed", 0))
    r = segment_mws_blockwise(torch.from_numpy(affs).cuda(), params, block, ctx, mask=None if mask is None else torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    f = r["fragments"].cpu().numpy().view(np.uint64)
    d = f != ref["fragments"]
    print(f"case {ci} {dtype.__name__} frag voxels differing: {int(d.sum())} of {d.size}; gpu frags {len(np.unique(f))} ref {len(np.unique(ref['fragments']))}"
          f"; zero gpu {int((f == 0).sum())} ref {int((ref['fragments'] == 0).sum())}", flush=True)
    if d.any():
        idx = np.argwhere(d)[:5]
        for z, y, x in idx:
            print("   ", (z, y, x), int(f[z, y, x]), int(ref["fragments"][z, y, x]))
        # same partition?
        pairs = np.unique(np.stack([f.ravel(), ref["fragments"].ravel()], 1), axis=0)
        print("    partition-equal:", len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1])))
        continue
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    keys = sorted(ref["rag"].edges)
    want = np.array(keys, dtype=np.uint64).reshape(-1, 2)
    got = np.stack([eu.view(np.uint64), ev.view(np.uint64)], 1)
    print(f"    edges gpu {len(got)} ref {len(want)} equal sets: {np.array_equal(got, want)}")
    if np.array_equal(got, want):
        ws = np.array([ref["rag"].edges[k] for k in keys], dtype=np.float32)
        print("    scores equal:", np.array_equal(es, ws), "max abs diff", float(np.abs(es - ws).max()) if len(ws) else 0)
    print("    lut equal:", np.array_equal(r["lut"][1].cpu().numpy().view(np.uint64), ref["lut"][1]),
          "seg equal:", np.array_equal(r["seg"].cpu().numpy().view(np.uint64), ref["seg"]))
