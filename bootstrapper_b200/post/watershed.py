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
                                block_size=block_size, context=ctx, num_workers=num_workers, roi=total_roi, **frag_params)
    run_volara_task(frags_task, blockwise)
    dump_params(frags_ds_name, {"method": "ws", "blockwise": blockwise, **frag_params})

    # stage 2: score RAG edges
    agglom_task = WaterzAgglom(db=db, affs_data=affinities, frags_data=fragments, block_size=block_size, context=ctx,
                               num_workers=num_workers, roi=total_roi, merge_function=waterz_merge_function)
    agglom_task._affs_dev = frags_task._affs_dev        # keep the affinities resident between the stages
    agglom_task._frags_dev = frags_task._frags_dev
    run_volara_task(agglom_task, blockwise)

    # stage 3: thresholded connected components -> LUT -> relabel
    nodes, edges, scores = db.read_graph()
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
    raise NotImplementedError(
        "the single-shot ws path (waterz with the non-discretised queue, post/watershed.py:206-354) is not built yet; "
        "use blockwise=true (block_shape='roi' gives one block)")


def watershed_segmentation(config):
    blockwise = config.get("blockwise", False)
    if blockwise:
        if config.get("block_shape") == "roi":
            config["blockwise"] = False
        waterz_pipeline(config)
    else:
        simple_watershed(config)
