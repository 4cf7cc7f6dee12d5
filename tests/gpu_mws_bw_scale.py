"""GPU-side measurement (not a test): the blockwise mws pipeline (ExtractFrags -> AffAgglom -> GraphMWS -> Relabel) on cubes of
BASELINE config-3 geometry (nine offsets, default strides / biases, seeded noise, 128^3 blocks, context 16).
Usage: python tests/gpu_mws_bw_scale.py [edge ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import mws_affs9, MWS_NBH, MWS_BIAS, MWS_STRIDES  # noqa: E402
from bootstrapper_b200.post.pipeline import segment_mws_blockwise  # noqa: E402

params = dict(aff_neighborhood=MWS_NBH, bias=MWS_BIAS, strides=MWS_STRIDES, noise_eps=0.001, noise_seed=0)
for edge in [int(v) for v in sys.argv[1:]] or [256]:
    shape = (edge, edge, edge)
    affs = mws_affs9(shape, seed=0)
    torch.cuda.synchronize()
    t0 = time.time()
    r = segment_mws_blockwise(affs, params, (128, 128, 128), (16, 16, 16), profile=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    c = r["counters"]
    print(json.dumps({"shape": shape, "seconds": dt, "voxels_per_s": edge ** 3 / dt, "fragments": int(r["nodes"][0].numel()),
                      "edges": int(r["edges"][0].numel()), "segments": int(torch.unique(r["lut"][1]).numel()),
                      "stage_s": {k: round(v, 3) for k, v in r["stage_s"].items()}, "extract_frags": c["extract_frags"], "graph_mws": c["graph_mws"]}), flush=True)
    del r, affs
    torch.cuda.empty_cache()
