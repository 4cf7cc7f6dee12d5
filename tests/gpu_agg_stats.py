"""per-block agglomeration counters at bench scale (debug aid)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native
from bootstrapper_b200.post.pipeline import make_plan
shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (50, 750, 750)
affs = native.synth_affs(shape, seed=0)
native.set_debug(True); native.set_profiling(True)
if os.environ.get("BS_AGG_GLOBAL"): native.set_agglom_version(3)
blk = tuple(int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (25, 250, 250)
plan, p = make_plan(affs, {}, blk, tuple(max(1, b // 8) for b in blk))
frags = plan.fragments(affs)
plan.agglomerate(affs, frags)
torch.cuda.synchronize()
print(native.get_profile())
c = plan.debug_fetch("s2_counters", np.uint32).reshape(-1, 6)
eb = plan.debug_fetch("s2_ebase", np.uint32)
nm = plan.debug_fetch("s2_nmerges", np.uint32)
print("block: E merges pops stale dead iters chunksteps appends")
for i in range(len(nm)):
    print(i, eb[i + 1] - eb[i], nm[i], *c[i])
print("mean", np.mean(np.diff(eb)), nm.mean(), c.mean(0))
