"""GPU-side debugging aid for the fused stage-1 front end (not a test): runs one small TMA-eligible volume through the
front-end variants and reports where they differ.  Usage: python tests/gpu_front_debug.py [front_version ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native  # noqa: E402
from bootstrapper_b200.post.pipeline import segment_blockwise  # noqa: E402
from bootstrapper_b200.synth import synth_affs  # noqa: E402

versions = [int(v) for v in sys.argv[1:]] or [1, 2, 3]
affs = torch.from_numpy(synth_affs((4, 128, 160), seed=11)).cuda()
out = {}
for fv in versions:
    native.set_front_version(fv)
    r = segment_blockwise(affs, {}, (2, 64, 80), (1, 8, 8))
    torch.cuda.synchronize()
    out[fv] = r["fragments"].cpu().numpy()
    print("front version", fv, "fragments", len(np.unique(out[fv])) - 1, flush=True)
for fv in versions[1:]:
    d = out[fv] != out[versions[0]]
    print("version", fv, "vs", versions[0], ":", int(d.sum()), "voxels differ", flush=True)
