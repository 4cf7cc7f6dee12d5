"""GPU-side measurement (not a test): the CUDA mutex watershed on growing cubes of BASELINE config-3 geometry (nine offsets,
default strides / biases, seeded noise).  Usage: python tests/gpu_mws_scale.py [edge ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import mws_affs9, MWS_NBH, MWS_BIAS, MWS_STRIDES  # noqa: E402
from bootstrapper_b200 import native  # noqa: E402

for edge in [int(v) for v in sys.argv[1:]] or [128, 256]:
    shape = (edge, edge, edge)
    affs = mws_affs9(shape, seed=0)
    torch.cuda.synchronize()
    t0 = time.time()
    frags, _, cnt = native.mws_agglom(affs, MWS_NBH, MWS_BIAS, strides=MWS_STRIDES, noise_eps=0.001, noise_seed=0)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({"shape": shape, "seconds": dt, "voxels_per_s": edge ** 3 / dt, "fragments": int(torch.unique(frags).numel()), **cnt}), flush=True)
