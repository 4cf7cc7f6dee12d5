"""ORACLE: the blockwise pipeline with one process per block (mirrors daisy's worker processes,
reference blockwise.py:44-50 / configs.py:591-593 num_workers).  Test / cpu_baseline infrastructure only.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import blockwise as ob

_G = {}


def _stage1(i):
    blk = _G["blocks"][i]
    frags = np.zeros(blk.write_shape, dtype=np.uint64)
    rag = ob.Rag()
    # write into a block-local array: shift the roi offset to the block's write offset
    ob.watershed_in_block(blk, _G["affs"], frags, rag, _G["p"], blk.write_offset, _G["block_size"], _G["mask"],
                          _G["seed_tie"], _G["stats_mode"])
    return i, frags, rag.node_pos, rag.node_size


def _stage2(i):
    blk = _G["blocks"][i]
    rag = ob.Rag(node_pos=_G["node_pos"])
    ob.agglomerate_in_block(blk, _G["affs"], _G["frags"], rag, _G["roi_offset"], _G["stats_mode"], _G["keep_cheaper"])
    return i, rag.edges


def waterz_pipeline_parallel(affs, params=None, block_size=None, context=None, roi=None, mask=None, seed_tie="index",
                             stats_mode="canonical", keep_cheaper=True, workers=None, block_subset=None, timings=None,
                             index_offset=None):
    """Same result as oracle.blockwise.waterz_pipeline, blocks distributed over a fork pool."""
    p = dict(ob.WS_DEFAULTS)
    p.update(params or {})
    vol_shape = affs.shape[1:]
    roi_offset, roi_shape = roi if roi is not None else ((0, 0, 0), vol_shape)
    if block_size is None:
        block_size, context = tuple(vol_shape), (0, 0, 0)
    elif context is None:
        context = tuple(max(1, s // 8) for s in block_size)
    blocks = ob.enumerate_blocks(roi_offset, roi_shape, block_size, context, index_offset)
    if block_subset is not None:
        blocks = [blocks[i] for i in block_subset]
    workers = workers or os.cpu_count()
    _G.update(affs=affs, p=p, block_size=block_size, mask=mask, seed_tie=seed_tie, stats_mode=stats_mode,
              keep_cheaper=keep_cheaper, blocks=blocks, roi_offset=roi_offset)
    frags = np.zeros(roi_shape, dtype=np.uint64)
    rag = ob.Rag()
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(workers) as pool:
        for i, f, npos, nsize in pool.imap_unordered(_stage1, range(len(blocks))):
            b = blocks[i]
            sl = tuple(slice(w - o, w - o + s) for w, o, s in zip(b.write_offset, roi_offset, b.write_shape))
            frags[sl] = f
            rag.node_pos.update(npos)
            rag.node_size.update(nsize)
    t1 = time.time()
    _G.update(frags=frags, node_pos=rag.node_pos)
    with ctx.Pool(workers) as pool:
        for i, edges in pool.imap_unordered(_stage2, range(len(blocks))):
            rag.edges.update(edges)
    t2 = time.time()
    segs = ob.global_segmentation(frags, rag, p["thresholds"])
    t3 = time.time()
    if timings is not None:
        timings.update(fragments=t1 - t0, agglomerate=t2 - t1, segment=t3 - t2, total=t3 - t0, workers=workers,
                       blocks=len(blocks))
    return dict(fragments=frags, rag=rag, segs=segs, blocks=blocks, params=p)
