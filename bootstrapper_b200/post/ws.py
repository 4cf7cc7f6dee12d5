"""Seeded watershed on boundary distance — the reference's post/ws.py plug point
(watershed_from_affinities :38-112), computed by libbsnative on the GPU.

Accepts a numpy array or a torch tensor (moved to the current CUDA device); there is no CPU fallback.
Fragment ids are consecutive from 1 in raster order of each fragment's first voxel (z-major), which is a
relabelling of the reference's ids (those also count seed plateaus that the mask removes, post/ws.py:19-24);
the partition is identical.
"""
import numpy as np
import torch

from .. import native


def watershed_from_affinities(affs, max_affinity_value=1.0, fragments_in_xy=False, return_seeds=False,
                              min_seed_distance=10):
    if return_seeds:
        raise NotImplementedError("return_seeds=True is not implemented in the CUDA path")
    as_numpy = isinstance(affs, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(affs)) if as_numpy else affs
    if t.dtype == torch.uint8:
        if max_affinity_value != 255:
            raise ValueError("uint8 affinities require max_affinity_value=255")
    elif t.dtype == torch.float64:
        # the kernels read uint8 / float32; the affinities enter post/ws.py only through the boundary mask (:64,77 / :100),
        # so float64 input gets its mask from the reference's own float64 expression (elementwise IEEE arithmetic, evaluated
        # on the device) and goes down as a uint8 stand-in with exactly that mask (255 where set: 510 > 255, 765 > 382)
        if t.shape[0] < 2:
            raise ValueError("need at least the y and x affinity channels")
        t = t.cuda()
        if fragments_in_xy:
            mask = 0.5 * (t[-1] + t[-2]) > 0.5 * max_affinity_value
        else:
            acc = t[0]
            for c in range(1, t.shape[0]):
                acc = acc + t[c]                                  # np.mean(affs, axis=0): planes added in order, one division
            mask = acc / float(t.shape[0]) > 0.5 * max_affinity_value
        t = (mask.to(torch.uint8) * 255).unsqueeze(0).expand(3, -1, -1, -1)
    else:
        if max_affinity_value != 1.0:
            raise ValueError("float affinities require max_affinity_value=1.0")
    if t.shape[0] < 2:
        raise ValueError("need at least the y and x affinity channels")
    if t.shape[0] == 2:   # 2-channel input: prepend an empty z channel (post/watershed.py:305-308)
        t = torch.cat([torch.zeros_like(t[:1]), t], 0)
    t = t[:3].contiguous().cuda()
    frags, n = native.watershed_from_affinities(t, fragments_in_xy, min_seed_distance)
    if as_numpy:
        return frags.cpu().numpy().view(np.uint64), n
    return frags, n
