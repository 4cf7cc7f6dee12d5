"""mws drivers — drop-in for the reference's post/watershed_mutex.py: `simple_mutex` (single shot, :177-291) and the
dispatch `mutex_watershed_segmentation` (:294-303).  Same config keys, dataset names and zarr attrs; the mutex watershed
runs in libbsnative on the GPU.  The blockwise `volara_pipeline` (:8-174) runs the arithmetic of the four volara tasks
(ExtractFrags -> AffAgglom -> GraphMWS -> Relabel) for all blocks at once through `pipeline.segment_mws_blockwise` and writes
the same datasets, RAG database (edge attribute zyx_aff) and LUT.
"""
import os

import numpy as np
import torch

from ..zarrio import open_ds, prepare_ds
from .mws import mwatershed_from_affinities
from .naming import build_name, dump_params


def volara_plan(config, affs_shape, affs_chunk_shape, affs_roi):
    """the orchestration of post/watershed_mutex.py:22-173 without the arrays: parameter defaults, block size / context rule,
    the arguments of the four volara tasks and the dataset / LUT names (pinned against the reference function by
    tests/golden/volara_pipeline_glue.json)"""
    from pathlib import Path
    neighborhood, bias = config.get("aff_neighborhood"), config.get("bias")
    global_bias = tuple(config.get("global_bias", [1.0, -0.5]))
    filter_fragments, sigma, noise_eps = config.get("filter_fragments"), config.get("sigma"), config.get("noise_eps")
    strides, randomized_strides = config.get("strides"), config.get("randomized_strides", False)
    remove_debris, min_seed_distance = config.get("remove_debris", 0), config.get("min_seed_distance")
    roi_offset, roi_shape = config.get("roi_offset"), config.get("roi_shape")
    blockwise = config.get("blockwise", False)
    num_workers = config.get("num_workers", 1) if blockwise else 1
    block_shape, context = config.get("block_shape"), config.get("context")
    if neighborhood is None:
        raise ValueError("Affinities neighborhood must be provided")
    if bias is None:
        raise ValueError("Affinities bias must be provided")
    assert len(neighborhood) == len(bias), "Number of biases must match number of affinities channels"
    roi = (tuple(roi_offset), tuple(roi_shape)) if roi_offset is not None else (tuple(affs_roi[0]), tuple(affs_roi[1]))
    if blockwise:
        block_size = tuple(block_shape) if block_shape else tuple(affs_chunk_shape[1:])
        ctx = tuple(context) if context else tuple(max(1, s // 8) for s in block_size)
    else:
        block_size, ctx = tuple(affs_shape[1:]), (0,) * len(affs_roi[0])
    frag_params = {"min_seed_distance": min_seed_distance, "sigma": sigma, "noise_eps": noise_eps, "bias": bias, "strides": strides,
                   "randomized_strides": randomized_strides, "filter_fragments": filter_fragments, "remove_debris": remove_debris}
    seg_params = {"global_bias": list(global_bias), **frag_params}
    return dict(
        blockwise=blockwise, num_workers=num_workers, block_size=block_size, context=ctx, roi=roi, frag_params=frag_params, seg_params=seg_params,
        frags_ds_name=str(Path(config["fragments_dataset"]) / build_name(frag_params)),
        lut_name=str(Path(config["lut_dir"]) / build_name(seg_params)),
        seg_name=str(Path(config["seg_dataset_prefix"]) / build_name(seg_params)),
        extract_frags=dict(bias=bias, sigma=sigma, noise_eps=noise_eps, filter_fragments=filter_fragments, remove_debris=remove_debris,
                           strides=strides, randomized_strides=randomized_strides, min_seed_distance=min_seed_distance),
        scores={"zyx_aff": neighborhood}, weights={"zyx_aff": global_bias}, edge_attrs={"zyx_aff": "float"})


def volara_pipeline(config):
    """post/watershed_mutex.py:8-174"""
    from ..graphdb import LUT, open_db
    from .naming import dump_lut_params
    from .pipeline import segment_mws_blockwise

    affs_dataset = config["affs_dataset"]
    db_config = config["db"]
    mask_dataset = config.get("mask_dataset")
    lut_dir = config["lut_dir"]
    affs = open_ds(affs_dataset)
    vs = affs.voxel_size
    pl = volara_plan(config, affs.shape, affs.chunk_shape, affs.roi)
    blockwise, block_size, ctx = pl["blockwise"], pl["block_size"], pl["context"]
    offset, shape = pl["roi"]
    frag_params, seg_params = pl["frag_params"], pl["seg_params"]
    frags_ds_name, lut_name, seg_name = pl["frags_ds_name"], pl["lut_name"], pl["seg_name"]
    neighborhood, bias, global_bias = config.get("aff_neighborhood"), config.get("bias"), pl["weights"]["zyx_aff"]
    ef = pl["extract_frags"]
    filter_fragments, sigma, noise_eps, strides = ef["filter_fragments"], ef["sigma"], ef["noise_eps"], ef["strides"]
    randomized_strides, remove_debris = ef["randomized_strides"], ef["remove_debris"]

    affs_t = torch.from_numpy(np.ascontiguousarray(affs.read())).cuda()
    if affs_t.dtype not in (torch.uint8, torch.float32):
        affs_t = affs_t.to(torch.float32)
    mask_t = None
    if mask_dataset is not None:
        mask = open_ds(mask_dataset)
        if tuple(mask.voxel_size) != tuple(vs):
            raise ValueError("mask voxel size differs from the affinities'")
        mstart = [int((ao - mo) // v) for ao, mo, v in zip(affs.offset, mask.offset, vs)]
        m = mask.to_ndarray(mstart, list(affs.shape[1:]), fill_value=0)
        mask_t = torch.from_numpy(np.ascontiguousarray((m > 0).astype(np.uint8))).cuda()
    start = tuple(int((o - ao) // v) for o, ao, v in zip(offset, affs.offset, vs))
    size = tuple(int(s // v) for s, v in zip(shape, vs))
    absolute = tuple(int(o) // int(v) for o, v in zip(offset, vs))      # daisy block ids: world offset / voxel size (U10)
    r = segment_mws_blockwise(affs_t, dict(aff_neighborhood=neighborhood, bias=bias, global_bias=global_bias, filter_fragments=filter_fragments,
                                           sigma=sigma, noise_eps=noise_eps, strides=strides, randomized_strides=randomized_strides,
                                           remove_debris=remove_debris, noise_seed=int(config.get("noise_seed", 0))),
                              block_size, ctx, mask=mask_t, roi=(start, size), block_index_offset=absolute)

    axis_names = affs.axis_names[1:] if affs.axis_names else None
    out = prepare_ds(frags_ds_name, size, tuple(offset), vs, np.uint64, chunk_shape=block_size, axis_names=axis_names, units=affs.units)
    out.write(r["fragments"].cpu().numpy().view(np.uint64))
    dump_params(frags_ds_name, {"method": "mws", "blockwise": blockwise, **frag_params})

    db = open_db(db_config, edge_attrs=pl["edge_attrs"])
    db.drop()
    db.init()
    nid, npos, nsz = [t.cpu().numpy() for t in r["nodes"]]
    world = npos.astype(np.int64) * np.array(vs, dtype=np.int64) + np.array(affs.offset, dtype=np.int64)
    db.write_nodes(nid.view(np.uint64), world, nsz)
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    db.write_edges(eu.view(np.uint64), ev.view(np.uint64), es)

    os.makedirs(lut_dir, exist_ok=True)
    lut = np.stack([r["lut"][0].cpu().numpy().view(np.uint64), r["lut"][1].cpu().numpy().view(np.uint64)])
    LUT(path=lut_name).save(lut, edges=np.stack([eu.view(np.uint64), ev.view(np.uint64)], 1))
    dump_lut_params(lut_name, {"method": "mws", "blockwise": blockwise, **seg_params})

    out = prepare_ds(seg_name, size, tuple(offset), vs, np.uint64, chunk_shape=block_size, axis_names=axis_names, units=affs.units)
    out.write(r["seg"].cpu().numpy().view(np.uint64))
    dump_params(seg_name, {"method": "mws", "blockwise": blockwise, **seg_params})


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
