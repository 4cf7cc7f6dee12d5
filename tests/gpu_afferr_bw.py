"""bs_aff_errors on a config-2 sized volume: time and algorithmic HBM bandwidth (8 B seg + 4 C pred in, 4 + 1 out, + 4 C
for the optional seg_affs) against the measured copy peak.   python tests/gpu_afferr_bw.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bootstrapper_b200 import native
shape, nh = (125, 1250, 1250), [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]
g = torch.Generator(device="cuda").manual_seed(0)
seg = torch.randint(1, 5000, (shape[0], shape[1] // 25, shape[2] // 25), device="cuda", generator=g).repeat_interleave(25, 1).repeat_interleave(25, 2).contiguous()
peak = 6516.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
n = int(np.prod(shape))
for dt, with_affs in ((torch.float32, True), (torch.float32, False), (torch.uint8, False)):
    pred = (torch.rand((3,) + shape, device="cuda", generator=g) * (255 if dt == torch.uint8 else 1)).to(dt)
    for _ in range(3):
        native.aff_errors(seg, pred, nh, return_seg_affs=with_affs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        native.aff_errors(seg, pred, nh, return_seg_affs=with_affs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byt = n * (8 + 3 * pred.element_size() + 5 + (12 if with_affs else 0))
    print(f"pred {str(dt):14s} seg_affs {with_affs!s:5s}: {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s algorithmic = {byt / ms / 1e6 / peak:.1%} of {peak:.0f} GB/s  ({n / ms / 1e6:.1f} Gvox/s)")
# per-label statistics of the same volume (bs refine): 8 bytes read per voxel
for _ in range(2):
    native.label_stats(seg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ids, sizes, zlo, zhi = native.label_stats(seg)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"label_stats: {ids.numel()} ids, {ms:.3f} ms  {8 * n / ms / 1e6:.0f} GB/s = {8 * n / ms / 1e6 / peak:.1%} of peak")
