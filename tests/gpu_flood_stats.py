"""flood step counters of one stage-1 batch (debug aid): total steps, steps cut by a higher level, longest tile
    python tests/gpu_flood_stats.py Z Y X bz by bx [json ws_params]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native
from bootstrapper_b200.post.pipeline import make_plan
shape = tuple(int(v) for v in sys.argv[1:4])
blk = tuple(int(v) for v in sys.argv[4:7])
params = json.loads(sys.argv[7]) if len(sys.argv) > 7 else {}
affs = native.synth_affs(shape, seed=0)
native.set_debug(True); native.set_profiling(True)
plan, p = make_plan(affs, params, blk, tuple(max(1, b // 8) for b in blk))
frags = plan.fragments(affs)
torch.cuda.synchronize()
print({k: round(v, 2) for k, v in native.get_profile().items()})
st = plan.debug_fetch("flood_stats", np.uint32)
print("steps %d  interrupted %d  max steps per tile %d  flooded voxels %d" % (st[0], st[1], st[2], int((frags != 0).sum())))
