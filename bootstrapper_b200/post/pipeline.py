"""In-memory core of the blockwise ws pipeline: device tensors in, device tensors out.

This is what `post/watershed.py:waterz_pipeline` (the file-based drop-in) and `bench.py` drive.
Mirrors the three stages + barriers of the reference's waterz_pipeline (post/watershed.py:8-203):
WatershedFrags -> WaterzAgglom -> thresholded connected components -> LUT -> Relabel.
Everything numeric runs in libbsnative (CUDA, sm_100a); torch only owns the memory and the stream.
"""
import numpy as np
import torch

from .. import native

WS_DEFAULTS = dict(  # reference segment.py:11-23
    fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0,
    filter_fragments=0.1, remove_debris=64, thresholds=[0.2, 0.35, 0.5], merge_function="mean",
    sigma=None, noise_eps=None, bias=None, noise_seed=0)

UNSUPPORTED = ()
# noise_eps: the reference draws an UNSEEDED np.random.randn per block (watershed_frags.py:119-120); the CUDA path uses a seeded
# counter-based generator instead (extra parameter noise_seed, default 0; include/bsnative.h) that the oracle mirrors


def resolve_ws_params(params):
    p = dict(WS_DEFAULTS)
    p.update(params or {})
    for k in UNSUPPORTED:
        if p.get(k) is not None:
            raise NotImplementedError(f"ws parameter {k!r} is not implemented in the CUDA path yet")
    if p["epsilon_agglomerate"]:
        raise NotImplementedError("epsilon_agglomerate > 0 is not implemented in the CUDA path yet")
    if p["merge_function"] != "mean":
        raise NotImplementedError("blockwise agglomeration supports merge_function='mean' only "
                                  "(as the reference: post/blockwise/waterz_agglom.py:24-36)")
    return p


SIMPLE_MERGE_FUNCTIONS = (  # post/watershed.py:232-244
    "mean", "hist_quant_10", "hist_quant_10_initmax", "hist_quant_25", "hist_quant_25_initmax", "hist_quant_50",
    "hist_quant_50_initmax", "hist_quant_75", "hist_quant_75_initmax", "hist_quant_90", "hist_quant_90_initmax")


def segment_simple(affs, params=None, mask=None):
    """In-memory core of `simple_watershed` (post/watershed.py:206-354): fragments of the whole ROI, then waterz
    with the default (non-discretised) queue, one segmentation per threshold.  affs: CUDA tensor (C, Z, Y, X)
    uint8 (normalised /255 as the reference does) or float32; only [:3] is read; a 2-channel input gets an empty
    z channel.  Returns dict(fragments, n, segs {thr: tensor}, counters)."""
    if not affs.is_cuda:
        raise native.BsError("segment_simple needs the affinities on a CUDA device (no CPU fallback)")
    p = dict(WS_DEFAULTS)
    p.update(params or {})
    if p.get("noise_eps") is not None:
        raise NotImplementedError("ws parameter 'noise_eps' (an unseeded RNG in the reference) is not implemented in the CUDA path")
    if p["merge_function"] not in SIMPLE_MERGE_FUNCTIONS:
        raise KeyError(p["merge_function"])
    if p["merge_function"] != "mean":
        raise NotImplementedError("only merge_function='mean' (OneMinus<MeanAffinity>) is implemented in the CUDA path")
    affs = affs[:3]
    if affs.shape[0] == 2:   # post/watershed.py:305-308
        affs = torch.cat([torch.zeros_like(affs[:1]), affs], 0)
    affs = affs.contiguous()
    if p.get("sigma") is not None or p.get("bias") is not None:
        # affs_data += shift, in float32 as the reference computes it (post/watershed.py:284-303); the mask is applied
        # before the shift, and everything downstream (fragments, waterz) sees the shifted affinities
        affs = native.shift_affinities(affs, mask=None if mask is None else (mask > 0).to(torch.uint8).contiguous(),
                                       sigma=p.get("sigma"), bias=p.get("bias"))
    elif mask is not None:   # affs_data *= (mask > 0)   (post/watershed.py:271-272)
        affs = (affs * (mask > 0).to(affs.dtype)).contiguous()
    vol_shape = tuple(affs.shape[1:])
    plan = native.Plan(vol_shape, vol_shape, (0, 0, 0), native._aff_dtype(affs), n_channels=3,
                       fragments_in_xy=p["fragments_in_xy"], min_seed_distance=p["min_seed_distance"],
                       filter_fragments=0.0, remove_debris=0)
    frags = plan.fragments(affs)
    outs, thr, counters = plan.waterz_segment(affs, frags, p["thresholds"])
    segs = {}
    for t in p["thresholds"]:
        segs[t] = outs[thr.index(float(np.float32(t)))]
    return dict(fragments=frags, n=plan.num_nodes(), segs=segs, counters=counters, params=p, plan=plan)


def default_context(block_size):
    """post/watershed.py:79-83"""
    return tuple(max(1, int(s) // 8) for s in block_size)


def make_plan(affs, params, block_size, context=None, roi=None, **kw):
    p = resolve_ws_params(params)
    vol_shape = tuple(affs.shape[1:])
    if block_size is None:                      # blockwise False / block_shape == "roi"
        block_size, context = vol_shape, (0, 0, 0)
    elif context is None:
        context = default_context(block_size)
    roi_offset, roi_shape = roi if roi is not None else ((0, 0, 0), vol_shape)
    return native.Plan(vol_shape, block_size, context, native._aff_dtype(affs), roi_offset=roi_offset,
                       roi_shape=roi_shape, n_channels=affs.shape[0], fragments_in_xy=p["fragments_in_xy"],
                       min_seed_distance=p["min_seed_distance"], filter_fragments=p["filter_fragments"],
                       remove_debris=p["remove_debris"], bias=p["bias"], seed_eps=p["seed_eps"], sigma=p["sigma"],
                       noise_eps=p["noise_eps"], noise_seed=p.get("noise_seed", 0) or 0, **kw), p


def segment_blockwise(affs, params=None, block_size=None, context=None, roi=None, mask=None, plan=None,
                      out=None):
    """affs: CUDA tensor (C, Z, Y, X) uint8 or float32.  Returns a dict of CUDA tensors:
    fragments (roi_shape, int64 bit pattern of uint64), nodes (ids, positions, sizes),
    edges (u, v, merge_score with NaN = NULL), luts {thr: components}, segs {thr: tensor}."""
    if not affs.is_cuda:
        raise native.BsError("segment_blockwise needs the affinities on a CUDA device (no CPU fallback)")
    if plan is None:
        plan, p = make_plan(affs, params, block_size, context, roi)
    else:
        p = resolve_ws_params(params)
    dev = affs.device
    frags = plan.fragments(affs, frags_out=None if out is None else out.get("fragments"), mask=mask)
    node_ids, node_pos, node_size = plan.nodes(dev)
    plan.agglomerate(affs, frags)
    eu, ev, es = plan.edges(dev)
    thrs = list(p["thresholds"])
    cmap = plan.components(node_ids, eu, ev, es, thrs)
    comps = [cmap[thr] for thr in thrs]
    luts = dict(zip(thrs, comps))
    segs = {}
    for i in range(0, len(thrs), 8):           # Relabel: up to 8 thresholds per pass over the fragments
        outs = None if out is None else out["segs"][i:i + 8]
        for thr, sg in zip(thrs[i:i + 8], plan.relabel(frags, comps[i:i + 8], outs)):
            segs[thr] = sg
    return dict(fragments=frags, nodes=(node_ids, node_pos, node_size), edges=(eu, ev, es), luts=luts, segs=segs,
                plan=plan, params=p)
