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
    if p["epsilon_agglomerate"] and p.get("noise_eps") is not None:
        raise NotImplementedError("epsilon_agglomerate > 0 together with noise_eps is not implemented in the CUDA path")
    if p["merge_function"] != "mean":
        raise NotImplementedError("blockwise agglomeration supports merge_function='mean' only "
                                  "(as the reference: post/blockwise/waterz_agglom.py:24-36)")
    return p


SIMPLE_MERGE_FUNCTIONS = (  # post/watershed.py:232-244
    "mean", "hist_quant_10", "hist_quant_10_initmax", "hist_quant_25", "hist_quant_25_initmax", "hist_quant_50",
    "hist_quant_50_initmax", "hist_quant_75", "hist_quant_75_initmax", "hist_quant_90", "hist_quant_90_initmax")


def merge_function_code(name):
    """post/watershed.py:232-244: "mean" -> (0, False); "hist_quant_Q[_initmax]" -> (Q, init_with_max), the template
    arguments of OneMinus<HistogramQuantileAffinity<RegionGraphType, Q, ScoreValue, 256, init_with_max>>"""
    if name not in SIMPLE_MERGE_FUNCTIONS:
        raise KeyError(name)
    if name == "mean":
        return 0, False
    parts = name.split("_")
    return int(parts[2]), len(parts) == 4


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
    quantile, initmax = merge_function_code(p["merge_function"])
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
    outs, thr, counters = plan.waterz_segment(affs, frags, p["thresholds"], quantile=quantile, init_with_max=initmax)
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


def _empty_block_check_needed(p):
    """The reference returns from a block whose raw affinities are all < 1e-3 before it does anything (watershed_frags.py:
    201-202).  Without that test such a block has no boundary mask and hence no fragments either -- unless the shift can lift
    mean(affs + shift) above 0.5: a bias, or noise (the seeded generator is bounded by 3.47 sigma)."""
    bias = p.get("bias")
    bmax = 0.0 if bias is None else max(bias) if isinstance(bias, (list, tuple)) else float(bias)
    return bmax + 3.47 * float(p.get("noise_eps") or 0.0) + 3e-3 > 0.5


def empty_blocks(plan, affs):
    """per block of the plan: raw affinities of the read ROI (all channels, before mask and scaling) all < 1e-3"""
    ids, wo, ws = plan.block_info()
    ctx = [int(plan.cfg.context[d]) for d in range(3)]
    vol = tuple(affs.shape[1:])
    peaks = []
    for i in range(len(ids)):
        lo = [max(int(wo[i][d]) - ctx[d], 0) for d in range(3)]
        hi = [min(int(wo[i][d]) + int(ws[i][d]) + ctx[d], vol[d]) for d in range(3)]
        peaks.append(affs[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].amax().to(torch.float32))
    peak = torch.stack(peaks).cpu().numpy()
    return (peak == 0) if affs.dtype == torch.uint8 else (peak < 1e-3)


def fragments_all_blocks(plan, affs, p, frags_out=None, mask=None):
    """stage 1 for every block of `plan` (bs_stage1_fragments), leaving out the blocks the reference skips as empty when that
    can change the result (see _empty_block_check_needed).  Returns the fragments tensor."""
    n, _ = plan.num_blocks()
    if not _empty_block_check_needed(p):
        return plan.fragments(affs, frags_out=frags_out, mask=mask)
    empty = empty_blocks(plan, affs)
    if frags_out is None:
        frags_out = torch.zeros(plan.roi_shape, dtype=torch.int64, device=affs.device)
    elif empty.any():
        frags_out.zero_()
    plan.set_owned(np.nonzero(~empty)[0])
    if not empty.all():
        plan.fragments(affs, frags_out=frags_out, mask=mask)
    counts = plan.block_counts()
    counts[empty] = 0
    plan.set_owned(np.arange(n))
    plan.set_block_counts(counts)
    return frags_out


def epsilon_fragments(plan, affs, p, frags_out, mask=None):
    """WatershedFrags with epsilon_agglomerate > 0 (watershed_frags.py:158-176, 182-183) for all blocks of `plan`:
    per block, the watershed fragments of its read ROI are merged by waterz (mean affinity, BinQueue<256>, float32
    affinities as the reference passes them) up to the threshold before they are filtered, cropped and relabelled.

    Orchestration over the library's own kernels: the read ROIs of all blocks of one shape are stacked into an auxiliary
    volume whose blocks have no context (zero fill outside the array, mask applied: exactly the arrays the reference's
    get_fragments sees); an auxiliary plan makes their fragments and scores their RAG with mergeUntil(eps)
    (bs_stage2_agglomerate_until); the connected components of the scored edges are the merged fragments; the back half of
    stage 1 then runs on those labels (bs_stage1_from_labels).  Returns (node ids, positions, sizes) of all blocks."""
    dev = affs.device
    eps = float(p["epsilon_agglomerate"])
    ids, wo, ws = plan.block_info()
    ctx = [int(plan.cfg.context[d]) for d in range(3)]
    vol = tuple(affs.shape[1:])
    groups = {}
    for i in range(len(ids)):
        groups.setdefault(tuple(int(ws[i][d]) + 2 * ctx[d] for d in range(3)), []).append(i)
    counts = np.zeros(len(ids), np.int64)
    nodes_all = []
    a3 = affs[:3]
    for rs, members in sorted(groups.items()):
        rz, ry, rx = rs
        nb = len(members)
        fake = torch.zeros((3, nb * rz, ry, rx), dtype=a3.dtype, device=dev)
        for k, bi in enumerate(members):
            ro = [int(wo[bi][d]) - ctx[d] for d in range(3)]
            lo = [max(ro[d], 0) for d in range(3)]
            hi = [min(ro[d] + rs[d], vol[d]) for d in range(3)]
            if any(h <= l for l, h in zip(lo, hi)):
                continue
            src = (slice(None),) + tuple(slice(l, h) for l, h in zip(lo, hi))
            dz, dy, dx = (lo[d] - ro[d] for d in range(3))
            piece = a3[src]
            if mask is not None:   # affs_data *= mask_data (watershed_frags.py:207-213)
                piece = piece * (mask[src[1:]] > 0).to(piece.dtype)
            fake[:, k * rz + dz:k * rz + dz + (hi[0] - lo[0]), dy:dy + (hi[1] - lo[1]), dx:dx + (hi[2] - lo[2])] = piece
        common = dict(fragments_in_xy=p["fragments_in_xy"], min_seed_distance=p["min_seed_distance"], filter_fragments=0.0,
                      remove_debris=0)
        aux = native.Plan((nb * rz, ry, rx), rs, (0, 0, 0), native._aff_dtype(fake), bias=p["bias"], seed_eps=p["seed_eps"],
                          sigma=p["sigma"], **common)
        frags_a = aux.fragments(fake)
        nodes_a = aux.nodes(dev)[0]
        if fake.dtype == torch.uint8:
            # the reference hands waterz affs_data[:3].astype(float32) of the float64-normalised array (watershed_frags.py:163)
            fake32 = (fake.to(torch.float64) / 255.0).to(torch.float32)
            aux2 = native.Plan((nb * rz, ry, rx), rs, (0, 0, 0), native.BS_DTYPE_F32, **common)
            aux2.set_block_counts(aux.block_counts())
        else:
            fake32, aux2 = fake, aux
        aux2.agglomerate_until(fake32, frags_a, eps)
        eu, ev, es = aux2.edges(dev)
        ok = ~torch.isnan(es)
        comp = native.connected_components(nodes_a, eu[ok].contiguous(), ev[ok].contiguous(), es[ok].contiguous(), 2.0) \
            if nodes_a.numel() else nodes_a
        merged = native.relabel(frags_a, nodes_a, comp) if nodes_a.numel() else frags_a
        labels = aux.dense_fragments(merged)              # 1 + rank of the merged fragment's id among the auxiliary nodes
        plan.set_owned(members)
        plan.fragments_from_labels(affs, labels, nodes_a.numel(), frags_out)
        c = plan.block_counts()
        counts[members] = c[members]
        if plan.num_nodes():
            nodes_all.append(plan.nodes(dev))
    plan.set_owned(np.arange(len(ids)))
    plan.set_block_counts(counts)
    if not nodes_all:
        return (torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev),
                torch.zeros(0, dtype=torch.int32, device=dev))
    nid = torch.cat([n[0] for n in nodes_all])
    order = torch.argsort(nid)
    return nid[order], torch.cat([n[1] for n in nodes_all])[order], torch.cat([n[2] for n in nodes_all])[order]


def _stack_read_rois(plan, affs, mask, members, rs, wo, ctx):
    """the read ROIs of the blocks `members` (all of read shape rs) stacked along z: zero fill outside the array, mask applied
    (affs_data *= mask_data) -- the arrays the reference's per-block task body sees.  Also returns, per block, whether the
    reference would skip it (`if affs_data.max() < 1e-3: return`, checked on the raw values before mask and scaling)."""
    dev = affs.device
    vol = tuple(affs.shape[1:])
    rz, ry, rx = rs
    nb = len(members)
    fake = torch.zeros((affs.shape[0], nb * rz, ry, rx), dtype=affs.dtype, device=dev)
    mstack = None if mask is None else torch.zeros((nb * rz, ry, rx), dtype=torch.bool, device=dev)
    for k, bi in enumerate(members):
        ro = [int(wo[bi][d]) - ctx[d] for d in range(3)]
        lo = [max(ro[d], 0) for d in range(3)]
        hi = [min(ro[d] + rs[d], vol[d]) for d in range(3)]
        if any(h <= l for l, h in zip(lo, hi)):
            continue
        src = tuple(slice(l, h) for l, h in zip(lo, hi))
        dz, dy, dx = (lo[d] - ro[d] for d in range(3))
        dst = (slice(k * rz + dz, k * rz + dz + (hi[0] - lo[0])), slice(dy, dy + (hi[1] - lo[1])), slice(dx, dx + (hi[2] - lo[2])))
        fake[(slice(None),) + dst] = affs[(slice(None),) + src]
        if mask is not None:
            mstack[dst] = mask[src] > 0
    peak = fake.view(affs.shape[0], nb, rz, ry, rx).amax(dim=(0, 2, 3, 4))
    empty = (peak == 0) if affs.dtype == torch.uint8 else (peak < 1e-3)
    if mstack is not None:
        fake *= mstack.to(fake.dtype)
    return fake, empty.cpu().numpy()


MWS_DEFAULTS = dict(aff_neighborhood=None, bias=None, global_bias=[1.0, -0.5], filter_fragments=None, sigma=None, noise_eps=None,
                    strides=None, randomized_strides=False, remove_debris=0, min_seed_distance=None, noise_seed=0)


def segment_mws_blockwise(affs, params, block_size=None, context=None, mask=None, roi=None, block_index_offset=None,
                          agglom_chunk_blocks=None, profile=False, mws_chunk_blocks=None):
    """In-memory core of the blockwise mws pipeline (post/watershed_mutex.py:8-174: ExtractFrags -> AffAgglom -> GraphMWS ->
    Relabel, the four volara tasks) on a CUDA tensor (C, Z, Y, X) uint8 / float32.
      ExtractFrags: per block, mutex-watershed fragments of the read ROI from all C channels (bs_mws_agglom_blocks: every
                    block of one read shape in one call), then filter_fragments / remove_debris / crop / label / id bump /
                    nodes exactly as WatershedFrags' back half (bs_stage1_from_labels);
      AffAgglom:    mean affinity over all offsets between touching fragments (bs_aff_agglom);
      GraphMWS:     global mutex watershed on the fragment graph, w = weight * zyx_aff + bias (bs_graph_mws);
      Relabel:      bs_stage3_relabel.
    Returns dict(fragments, nodes, edges (u, v, zyx_aff), lut (nodes, clusters), seg)."""
    if not affs.is_cuda:
        raise native.BsError("segment_mws_blockwise needs the affinities on a CUDA device (no CPU fallback)")
    from ..synth import block_seed
    p = dict(MWS_DEFAULTS)
    p.update(params or {})
    nbh, bias = p["aff_neighborhood"], p["bias"]
    if nbh is None:
        raise ValueError("Affinities neighborhood must be provided")
    if bias is None:
        raise ValueError("Affinities bias must be provided")
    assert len(nbh) == len(bias), "Number of biases must match number of affinities channels"
    assert len(nbh) == affs.shape[0], "Number of offsets must match number of affinities channels"
    if p["sigma"] is not None:
        raise NotImplementedError("mws parameter 'sigma' is not implemented in the CUDA path")
    if p["randomized_strides"]:
        raise NotImplementedError("randomized_strides=True draws an unseeded random subset of the stride lattice in mwatershed: "
                                  "not reproducible, not implemented; set randomized_strides = false")
    dev = affs.device
    affs = affs.contiguous()
    vol = tuple(affs.shape[1:])
    import time as _time
    stage_s = {}

    def _tick(name, t0):
        if profile:
            torch.cuda.synchronize()
            stage_s[name] = stage_s.get(name, 0.0) + _time.perf_counter() - t0
        return _time.perf_counter()
    if block_size is None:
        block_size, context = vol, (0, 0, 0)
    elif context is None:
        context = default_context(block_size)
    roi_offset, roi_shape = roi if roi is not None else ((0, 0, 0), vol)
    plan = native.Plan(vol, block_size, context, native._aff_dtype(affs), roi_offset=roi_offset, roi_shape=roi_shape,
                       block_index_offset=block_index_offset, n_channels=affs.shape[0], fragments_in_xy=False,
                       filter_fragments=float(p["filter_fragments"] or 0.0), remove_debris=int(p["remove_debris"] or 0))
    ids, wo, ws = plan.block_info()
    ctx = [int(plan.cfg.context[d]) for d in range(3)]
    groups = {}
    for i in range(len(ids)):
        groups.setdefault(tuple(int(ws[i][d]) + 2 * ctx[d] for d in range(3)), []).append(i)
    frags = torch.zeros(plan.roi_shape, dtype=torch.int64, device=dev)
    counts = np.zeros(len(ids), np.int64)
    nodes_all, mws_counters = [], []
    # one mutex watershed per call over as many stacked read ROIs as its 31-bit voxel / 32-bit edge indices hold
    Cn = int(affs.shape[0])
    limit_vox = min((1 << 31) - 1, ((1 << 32) - 2) // Cn)
    work = []
    for rs, members in sorted(groups.items()):
        per_call = max(1, limit_vox // int(np.prod(rs)))
        if mws_chunk_blocks:
            per_call = min(per_call, int(mws_chunk_blocks))
        work += [(rs, members[c0:c0 + per_call]) for c0 in range(0, len(members), per_call)]
    for rs, members in work:
        t0 = _tick("-", 0.0)
        fake, empty = _stack_read_rois(plan, affs, mask, members, rs, wo, ctx)
        t0 = _tick("stack_read_rois", t0)
        seeds = [block_seed(p["noise_seed"], int(ids[bi])) for bi in members] if p["noise_eps"] else None
        labels, cnt = native.mws_agglom_blocks(fake, len(members), nbh, bias, strides=p["strides"], noise_eps=p["noise_eps"], block_seeds=seeds)
        del fake
        t0 = _tick("extract_frags.mws", t0)
        for k in np.nonzero(empty)[0]:      # the reference returns before it writes anything for such a block
            labels[int(k) * rs[0]:(int(k) + 1) * rs[0]] = 0
        mws_counters.append(cnt)
        plan.set_owned(members)
        plan.fragments_from_labels(affs, labels, cnt["n_labels"], frags, mask=mask)
        del labels
        t0 = _tick("extract_frags.back_half", t0)
        c = plan.block_counts()
        counts[members] = c[members]
        if plan.num_nodes():
            nodes_all.append(plan.nodes(dev))
    plan.set_owned(np.arange(len(ids)))
    plan.set_block_counts(counts)
    if nodes_all:
        nid = torch.cat([n[0] for n in nodes_all])
        order = torch.argsort(nid)
        nodes = (nid[order], torch.cat([n[1] for n in nodes_all])[order], torch.cat([n[2] for n in nodes_all])[order])
    else:
        nodes = (torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev),
                 torch.zeros(0, dtype=torch.int32, device=dev))
    t0 = _tick("-", 0.0)
    # AffAgglom in chunks of blocks (bounded hash-table memory); every edge is written by exactly one block
    read_vox = max(int(np.prod([int(ws[i][d]) + 2 * ctx[d] for d in range(3)])) for i in range(len(ids)))
    per_chunk = int(agglom_chunk_blocks) if agglom_chunk_blocks else max(1, (1 << 27) // max(read_vox, 1))
    parts = []
    for c0 in range(0, len(ids), per_chunk):
        plan.set_owned(np.arange(c0, min(c0 + per_chunk, len(ids))))
        plan.aff_agglom(affs, frags, nbh)
        parts.append(plan.edges(dev))
    plan.set_owned(np.arange(len(ids)))
    eu, ev, es = (torch.cat([p_[k] for p_ in parts]) for k in range(3))
    if len(parts) > 1 and eu.numel():
        nid0 = nodes[0]
        key = torch.searchsorted(nid0, eu) * nid0.numel() + torch.searchsorted(nid0, ev)      # (u, v) order of the dense numbers
        order = torch.argsort(key)
        eu, ev, es = eu[order].contiguous(), ev[order].contiguous(), es[order].contiguous()
    t0 = _tick("aff_agglom", t0)
    weight, gbias = (float(v) for v in tuple(p["global_bias"]))
    clusters, gcnt = native.graph_mws(nodes[0], eu, ev, es, weight, gbias)
    t0 = _tick("graph_mws", t0)
    seg = plan.relabel(frags, [clusters])[0] if nodes[0].numel() else torch.zeros_like(frags)
    t0 = _tick("relabel", t0)
    stage_s.pop("-", None)
    return dict(fragments=frags, nodes=nodes, edges=(eu, ev, es), lut=(nodes[0], clusters), seg=seg, plan=plan, params=p,
                counters=dict(extract_frags=mws_counters, graph_mws=gcnt), stage_s=stage_s)


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
    if p["epsilon_agglomerate"] and p["epsilon_agglomerate"] > 0:
        if roi is not None and tuple(roi[0]) != (0, 0, 0):
            raise NotImplementedError("epsilon_agglomerate with a ROI offset is not implemented")
        frags = out.get("fragments") if out is not None else torch.zeros(plan.roi_shape, dtype=torch.int64, device=dev)
        node_ids, node_pos, node_size = epsilon_fragments(plan, affs, p, frags, mask=mask)
    else:
        frags = fragments_all_blocks(plan, affs, p, frags_out=None if out is None else out.get("fragments"), mask=mask)
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
