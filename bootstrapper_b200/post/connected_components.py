"""`bs segment --cc` — drop-in for the reference's post/connected_components.py: `cc_affs` (:15-127),
`cc_blockwise` (:8-9, not implemented upstream either) and `cc_segmentation` (:130-137).  Same config keys,
dataset names and zarr attrs; the thresholding, the component labelling (post/cc.py) and remove_small_objects
run in libbsnative on the GPU.
"""
from pathlib import Path

import numpy as np
import torch

from .. import native
from ..zarrio import open_ds, prepare_ds
from .naming import build_name, dump_params


def cc_blockwise(config):
    raise NotImplementedError("Blockwise connected components not implemented yet")


def cc_in_memory(affs, threshold=0.5, remove_debris=0, mask=None, sigma=None, noise_eps=None):
    """affs: CUDA tensor (C >= 3, Z, Y, X) uint8 or float32.  Returns (fragments, segmentation) as int64 tensors
    holding the reference's uint32 ids."""
    if noise_eps is not None:
        raise NotImplementedError("cc parameter noise_eps (an unseeded RNG in the reference) is not implemented in the CUDA path")
    if not affs.is_cuda:
        raise native.BsError("cc_in_memory needs the affinities on a CUDA device (no CPU fallback)")
    affs = affs.contiguous()
    if sigma is not None:   # affs_data += gaussian_filter(affs_data, (0, *sigma)) - affs_data, after the mask (:62-77)
        affs = native.shift_affinities(affs, mask=mask, sigma=sigma)
        mask = None
    frags, seg, _ = native.cc_affs(affs, threshold, remove_debris, mask)
    return frags, seg


def cc_affs(config):
    affs_ds = config["affs_dataset"]
    frags_ds_prefix = config["fragments_dataset"]
    seg_ds_prefix = config["seg_dataset_prefix"]
    mask_ds = config.get("mask_dataset")
    roi_offset, roi_shape = config.get("roi_offset"), config.get("roi_shape")
    threshold = config.get("threshold", 0.5)
    sigma, noise_eps = config.get("sigma"), config.get("noise_eps")
    remove_debris = config.get("remove_debris", 0)

    affs = open_ds(affs_ds)
    offset, shape = (tuple(roi_offset), tuple(roi_shape)) if roi_offset is not None else affs.roi
    vs = affs.voxel_size
    start = [int((o - ao) // v) for o, ao, v in zip(offset, affs.offset, vs)]
    stop = [s0 + int(sh // v) for s0, sh, v in zip(start, shape, vs)]
    affs_data = affs.read((0,) + tuple(start), (3,) + tuple(stop))
    affs_t = torch.from_numpy(np.ascontiguousarray(affs_data)).to("cuda")
    if affs_t.dtype not in (torch.uint8, torch.float32):
        affs_t = affs_t.to(torch.float32)
    mask_t = None
    if mask_ds is not None:
        mask = open_ds(mask_ds)
        mstart = [int((o - mo) // v) for o, mo, v in zip(offset, mask.offset, mask.voxel_size)]
        mstop = [s0 + (b - a) for s0, a, b in zip(mstart, start, stop)]
        mask_t = (torch.from_numpy(np.ascontiguousarray(mask.read(tuple(mstart), tuple(mstop)))) > 0).to(torch.uint8).to("cuda")

    frag_params = {"threshold": threshold, "sigma": sigma, "noise_eps": noise_eps}
    frags_t, seg_t = cc_in_memory(affs_t, threshold, remove_debris, mask_t, sigma, noise_eps)
    names = affs.axis_names[1:] if affs.axis_names else None

    frags_ds_name = str(Path(frags_ds_prefix) / build_name(frag_params))
    frags = prepare_ds(frags_ds_name, tuple(frags_t.shape), tuple(offset), vs, np.uint64, axis_names=names, units=affs.units)
    frags.write(frags_t.cpu().numpy().view(np.uint64))
    dump_params(frags_ds_name, {"method": "cc", "blockwise": False, **frag_params})

    seg_params = {**frag_params, "remove_debris": remove_debris}
    seg_ds_name = str(Path(seg_ds_prefix) / build_name(seg_params))
    seg = prepare_ds(seg_ds_name, tuple(seg_t.shape), tuple(offset), vs, np.uint64, axis_names=names, units=affs.units)
    seg.write(seg_t.cpu().numpy().view(np.uint64))
    dump_params(seg_ds_name, {"method": "cc", "blockwise": False, **seg_params})


def cc_segmentation(config):
    if config.get("blockwise", False):
        cc_blockwise(config)
    else:
        cc_affs(config)
