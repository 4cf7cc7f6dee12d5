"""mws drivers — drop-in for the reference's post/watershed_mutex.py: `simple_mutex` (single shot, :177-291) and the
dispatch `mutex_watershed_segmentation` (:294-303).  Same config keys, dataset names and zarr attrs; the mutex watershed
runs in libbsnative on the GPU.  The blockwise `volara_pipeline` (:8-174: ExtractFrags -> AffAgglom -> GraphMWS -> Relabel)
is not built and raises.
"""
import os

import numpy as np
import torch

from ..zarrio import open_ds, prepare_ds
from .mws import mwatershed_from_affinities
from .naming import build_name, dump_params


def volara_pipeline(config):
    raise NotImplementedError("blockwise mws (volara ExtractFrags / AffAgglom / GraphMWS) is not part of the CUDA path; "
                              "run `bs segment --mws` without -b (simple_mutex)")


def simple_mutex(config):
    affs_ds = config["affs_dataset"]
    frags_ds_prefix = config["fragments_dataset"]
    seg_ds_prefix = config["seg_dataset_prefix"]
    mask_ds = config.get("mask_dataset", None)
    roi_offset, roi_shape = config.get("roi_offset", None), config.get("roi_shape", None)
    neighborhood, bias = config.get("aff_neighborhood", None), config.get("bias", None)
    sigma, noise_eps = config.get("sigma", None), config.get("noise_eps", None)
    strides, randomized_strides = config.get("strides", None), config.get("randomized_strides", False)
    remove_debris = config.get("remove_debris", 0)

    affs = open_ds(affs_ds)
    if neighborhood is None:
        raise ValueError("Affinities neighborrhood must be provided")
    if bias is None:
        raise ValueError("Affinities bias must be provided")
    assert len(neighborhood) == affs.shape[0], "Number of offsets must match number of affinities channels"
    assert len(neighborhood) == len(bias), "Numbes of biases must match number of affinities channels"

    offset, shape = (tuple(roi_offset), tuple(roi_shape)) if roi_offset is not None else affs.roi
    vs = affs.voxel_size
    start = [int((o - ao) // v) for o, ao, v in zip(offset, affs.offset, vs)]
    stop = [s0 + int(sh // v) for s0, sh, v in zip(start, shape, vs)]
    affs_data = affs.read((0,) + tuple(start), (affs.shape[0],) + tuple(stop))
    affs_t = torch.from_numpy(np.ascontiguousarray(affs_data)).cuda()
    if affs_t.dtype not in (torch.uint8, torch.float32):
        affs_t = affs_t.to(torch.float32)
    mask_t = None
    if mask_ds is not None:
        mask = open_ds(mask_ds)
        if tuple(mask.voxel_size) != tuple(vs):
            raise ValueError("mask voxel size differs from the affinities'")
        mstart = [int((o - mo) // v) for o, mo, v in zip(offset, mask.offset, vs)]
        m = mask.to_ndarray(mstart, [b - a for a, b in zip(start, stop)], fill_value=0)
        mask_t = torch.from_numpy(np.ascontiguousarray((m > 0).astype(np.uint8))).cuda()

    frags, seg, _ = mwatershed_from_affinities(affs_t, neighborhood, bias, sigma, noise_eps, strides, randomized_strides, mask=mask_t,
                                               noise_seed=int(config.get("noise_seed", 0)), remove_debris=remove_debris, return_counters=True)

    frag_params = {"sigma": sigma, "noise_eps": noise_eps, "bias": bias, "strides": strides, "randomized_strides": randomized_strides}
    frags_ds_name = os.path.join(frags_ds_prefix, build_name(frag_params))
    axis_names = affs.axis_names[1:] if affs.axis_names else None
    out = prepare_ds(frags_ds_name, tuple(frags.shape), tuple(offset), vs, np.uint64, axis_names=axis_names, units=affs.units)
    out.write(frags.cpu().numpy().view(np.uint64))
    dump_params(frags_ds_name, {"method": "mws", "blockwise": False, **frag_params})

    seg_params = {**frag_params, "remove_debris": remove_debris}
    seg_ds_name = os.path.join(seg_ds_prefix, build_name(seg_params))
    out = prepare_ds(seg_ds_name, tuple(frags.shape), tuple(offset), vs, np.uint64, axis_names=axis_names, units=affs.units)
    out.write((seg if remove_debris and remove_debris > 0 else frags).cpu().numpy().view(np.uint64))
    dump_params(seg_ds_name, {"method": "mws", "blockwise": False, **seg_params})


def mutex_watershed_segmentation(config):
    blockwise = config.get("blockwise", False)
    if blockwise:
        if config.get("block_shape") == "roi":
            config["blockwise"] = False
        volara_pipeline(config)
    else:
        simple_mutex(config)
