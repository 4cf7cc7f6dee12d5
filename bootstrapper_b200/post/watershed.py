"""ws drivers — drop-in for the reference's post/watershed.py: `waterz_pipeline` (blockwise: fragments ->
per-block waterz scoring -> global thresholded CC -> LUT -> relabel, :8-203), `simple_watershed`
(single shot, :206-354) and `watershed_segmentation` (:357-366).  Same config keys, dataset names,
zarr attrs, LUT files; the numerics run in libbsnative on the GPU.
"""
import logging
import os
from pathlib import Path

import numpy as np
import torch

from .. import native
from ..blockwise import run_volara_task
from ..datasets import Labels, Raw
from ..graphdb import LUT, open_db
from ..zarrio import open_ds, prepare_ds
from .blockwise.watershed_frags import WatershedFrags
from .blockwise.waterz_agglom import WATERZ_MERGE_FUNCTIONS, WaterzAgglom
from .naming import build_name, dump_lut_params, dump_params

logger = logging.getLogger(__name__)


def waterz_pipeline(config):
    affs_dataset = config["affs_dataset"]
    fragments_dataset_prefix = config["fragments_dataset"]
    seg_dataset_prefix = config["seg_dataset_prefix"]
    lut_dir = config.get("lut_dir") or seg_dataset_prefix.replace("segmentations", "luts")
    db_config = config["db"]
    mask_dataset = config.get("mask_dataset")

    frag_params = dict(
        fragments_in_xy=config.get("fragments_in_xy", True), min_seed_distance=config.get("min_seed_distance", 10),
        seed_eps=config.get("seed_eps"), epsilon_agglomerate=config.get("epsilon_agglomerate", 0.0),
        sigma=config.get("sigma"), noise_eps=config.get("noise_eps"), bias=config.get("bias"),
        filter_fragments=config.get("filter_fragments", 0.0), remove_debris=config.get("remove_debris", 0))
    thresholds = config.get("thresholds", [0.2, 0.35, 0.5])
    merge_function = config.get("merge_function", "mean")
    waterz_merge_function = WATERZ_MERGE_FUNCTIONS[merge_function]

    roi_offset, roi_shape = config.get("roi_offset"), config.get("roi_shape")
    blockwise = config.get("blockwise", False)
    num_workers = config.get("num_workers", 1) if blockwise else 1
    block_shape, context = config.get("block_shape"), config.get("context")

    affs = open_ds(affs_dataset)
    total_roi = (tuple(roi_offset), tuple(roi_shape)) if roi_offset is not None else affs.roi
    if blockwise:
        block_size = tuple(block_shape) if block_shape else tuple(affs.chunk_shape[1:])
        ctx = tuple(context) if context else tuple(max(1, s // 8) for s in block_size)
    else:
        block_size = tuple(affs.shape[1:])
        ctx = (0,) * 3

    frags_ds_name = str(Path(fragments_dataset_prefix) / build_name(frag_params))
    affinities = Raw(store=affs_dataset)
    mask_data = Raw(store=mask_dataset) if mask_dataset else None
    db = open_db(db_config)
    fragments = Labels(store=frags_ds_name)
    os.makedirs(lut_dir, exist_ok=True)

    # stage 1: fragments via seeded watershed
    frags_task = WatershedFrags(db=db, affs_data=affinities, frags_data=fragments, mask_data=mask_data,
                                block_size=block_size, context=ctx, num_workers=num_workers, roi=total_roi,
                                noise_seed=config.get("noise_seed", 0), **frag_params)
    run_volara_task(frags_task, blockwise)
    dump_params(frags_ds_name, {"method": "ws", "blockwise": blockwise, **frag_params})

    # stage 2: score RAG edges
    agglom_task = WaterzAgglom(db=db, affs_data=affinities, frags_data=fragments, block_size=block_size, context=ctx,
                               num_workers=num_workers, roi=total_roi, merge_function=waterz_merge_function)
    agglom_task._affs_dev = frags_task._affs_dev        # keep the affinities resident between the stages
    agglom_task._frags_dev = frags_task._frags_dev
    run_volara_task(agglom_task, blockwise)

    # stage 3: thresholded connected components -> LUT -> relabel
    nodes, edges, scores = db.read_graph(total_roi)          # post/watershed.py:156: read_graph(total_roi)
    if nodes.size == 0:
        logger.warning("empty RAG; no fragments to agglomerate")
        return
    keep = ~np.isnan(scores)                              # merge_score NULL: never merged (post/watershed.py:164-166)
    dev = "cuda"
    nodes_t = torch.from_numpy(nodes.view(np.int64)).to(dev)
    eu = torch.from_numpy(np.ascontiguousarray(edges[keep, 0]).view(np.int64)).to(dev)
    ev = torch.from_numpy(np.ascontiguousarray(edges[keep, 1]).view(np.int64)).to(dev)
    es = torch.from_numpy(scores[keep]).to(dev)
    frags_dev = agglom_task._frags()
    frag_arr = fragments.array("r")
    for threshold in thresholds:
        if eu.numel() == 0:
            comp = nodes_t.clone()
        else:
            comp = native.connected_components(nodes_t, eu, ev, es, float(threshold))
        params = {"merge_function": merge_function, "threshold": threshold, **frag_params}
        name = build_name(params)
        recorded = {"method": "ws", "blockwise": blockwise, **params}
        lut = LUT(path=str(Path(lut_dir) / name))
        lut.save(np.array([nodes, comp.cpu().numpy().view(np.uint64)]))
        dump_lut_params(str(Path(lut_dir) / name), recorded)
        seg_store = str(Path(seg_dataset_prefix) / name)
        seg = native.relabel(frags_dev, nodes_t, comp)
        out = prepare_ds(seg_store, frag_arr.shape, frag_arr.offset, frag_arr.voxel_size, np.uint64,
                         chunk_shape=block_size, axis_names=frag_arr.axis_names, units=frag_arr.units,
                         types=frag_arr.types)
        out.write(seg.cpu().numpy().view(np.uint64))
        dump_params(seg_store, recorded)


def simple_watershed(config):
    """post/watershed.py:206-354: single-shot fragments + waterz over the sorted thresholds, no blocks, no DB."""
    from .pipeline import segment_simple
    affs_ds = config["affs_dataset"]
    frags_ds_prefix = config["fragments_dataset"]
    seg_ds_prefix = config["seg_dataset_prefix"]
    mask_ds = config.get("mask_dataset")
    roi_offset, roi_shape = config.get("roi_offset"), config.get("roi_shape")
    thresholds = config.get("thresholds", [0.2, 0.35, 0.5])
    merge_function = config.get("merge_function", "mean")
    frag_params = dict(fragments_in_xy=config.get("fragments_in_xy", True),
                       min_seed_distance=config.get("min_seed_distance", 10), sigma=config.get("sigma"),
                       noise_eps=config.get("noise_eps"), bias=config.get("bias"))

    affs = open_ds(affs_ds)
    offset, shape = (tuple(roi_offset), tuple(roi_shape)) if roi_offset is not None else affs.roi
    vs = affs.voxel_size
    start = [int((o - ao) // v) for o, ao, v in zip(offset, affs.offset, vs)]
    stop = [s0 + int(sh // v) for s0, sh, v in zip(start, shape, vs)]
    affs_data = affs.read((0,) + tuple(start), (min(3, affs.shape[0]),) + tuple(stop))
    dev = "cuda"
    affs_t = torch.from_numpy(np.ascontiguousarray(affs_data)).to(dev)
    if affs_t.dtype not in (torch.uint8, torch.float32):
        affs_t = affs_t.to(torch.float32)
    mask_t = None
    if mask_ds is not None:
        mask = open_ds(mask_ds)
        mstart = [int((o - mo) // v) for o, mo, v in zip(offset, mask.offset, mask.voxel_size)]
        mstop = [s0 + (b - a) for s0, a, b in zip(mstart, start, stop)]
        mask_t = torch.from_numpy(np.ascontiguousarray(mask.read(tuple(mstart), tuple(mstop)))).to(dev)

    r = segment_simple(affs_t, dict(thresholds=thresholds, merge_function=merge_function, **frag_params), mask=mask_t)

    frags_ds_name = str(Path(frags_ds_prefix) / build_name(frag_params))
    frags = prepare_ds(frags_ds_name, tuple(r["fragments"].shape), tuple(offset), vs, np.uint64,
                       axis_names=affs.axis_names[1:] if affs.axis_names else None, units=affs.units)
    frags.write(r["fragments"].cpu().numpy().view(np.uint64))
    dump_params(frags_ds_name, {"method": "ws", "blockwise": False, **frag_params})
    for threshold in thresholds:
        params = {"merge_function": merge_function, "threshold": threshold, **frag_params}
        seg_ds_name = str(Path(seg_ds_prefix) / build_name(params))
        seg = prepare_ds(seg_ds_name, tuple(r["fragments"].shape), tuple(offset), vs, np.uint64,
                         axis_names=affs.axis_names[1:] if affs.axis_names else None, units=affs.units)
        seg.write(r["segs"][threshold].cpu().numpy().view(np.uint64))
        dump_params(seg_ds_name, {"method": "ws", "blockwise": False, **params})


def watershed_segmentation(config):
    blockwise = config.get("blockwise", False)
    if blockwise:
        if config.get("block_shape") == "roi":
            config["blockwise"] = False
        waterz_pipeline(config)
    else:
        simple_watershed(config)
