"""In-memory core of the blockwise ws pipeline: device tensors in, device tensors out.

This is what `post/watershed.py:waterz_pipeline` (the file-based drop-in) and `bench.py` drive.
Mirrors the three stages + barriers of the reference's waterz_pipeline (post/watershed.py:8-203):
WatershedFrags -> WaterzAgglom -> thresholded connected components -> LUT -> Relabel.
Everything numeric runs in libbsnative (CUDA, sm_100a); torch only owns the memory and the stream.
"""
import torch

from .. import native

WS_DEFAULTS = dict(  # reference segment.py:11-23
    fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0,
    filter_fragments=0.1, remove_debris=64, thresholds=[0.2, 0.35, 0.5], merge_function="mean",
    sigma=None, noise_eps=None, bias=None)

UNSUPPORTED = ("seed_eps", "sigma", "noise_eps", "bias")


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
                       remove_debris=p["remove_debris"], **kw), p


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
    comps = [native.connected_components(node_ids, eu, ev, es, float(thr)) for thr in thrs]
    luts = dict(zip(thrs, comps))
    segs = {}
    for i in range(0, len(thrs), 8):           # Relabel: up to 8 thresholds per pass over the fragments
        outs = None if out is None else out["segs"][i:i + 8]
        for thr, sg in zip(thrs[i:i + 8], plan.relabel(frags, comps[i:i + 8], outs)):
            segs[thr] = sg
    return dict(fragments=frags, nodes=(node_ids, node_pos, node_size), edges=(eu, ev, es), luts=luts, segs=segs,
                plan=plan, params=p)
