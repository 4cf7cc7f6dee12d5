"""Parity of the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bit-exact bar for everything integer (fragments incl. ids, nodes, RAG edge sets, LUTs, segmentations);
edge scores: identical float32 (tolerance 1e-6 relative per BASELINE.json north_star, asserted below).
Oracle switches: seed_tie="index" (DESIGN.md D1), stats_mode="canonical" (D2).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-6


def _run_gpu(affs_np, params, block, ctx, roi=None, mask_np=None):
    from bootstrapper_b200.post.pipeline import segment_blockwise
    dev = torch.device("cuda:0")
    affs = torch.from_numpy(affs_np).to(dev)
    mask = torch.from_numpy(mask_np).to(dev) if mask_np is not None else None
    r = segment_blockwise(affs, params, block, ctx, roi=roi, mask=mask)
    torch.cuda.synchronize()
    return r


def _check(r, ref):
    f = r["fragments"].cpu().numpy().view(np.uint64)
    assert np.array_equal(f, ref["fragments"]), "fragment ids differ"
    nid, npos, nsz = [t.cpu().numpy() for t in r["nodes"]]
    rn = np.array(sorted(ref["rag"].node_pos), dtype=np.uint64)
    assert np.array_equal(nid.view(np.uint64), rn)
    if len(rn):
        assert np.array_equal(npos, np.array([ref["rag"].node_pos[int(i)] for i in rn]))
        assert np.array_equal(nsz, np.array([ref["rag"].node_size[int(i)] for i in rn]))
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    got = dict(zip(zip(eu.view(np.uint64).tolist(), ev.view(np.uint64).tolist()), es.tolist()))
    want = ref["rag"].edges
    assert set(got) == set(want), "RAG edge sets differ"
    for k, s in want.items():
        if s is None:
            assert np.isnan(got[k])
        else:
            assert abs(got[k] - s) <= SCORE_RTOL * abs(s), (k, got[k], s)
    for thr, seg in r["segs"].items():
        assert np.array_equal(seg.cpu().numpy().view(np.uint64), ref["segs"][thr]["seg"]), f"segmentation {thr}"
        assert np.array_equal(r["luts"][thr].cpu().numpy().view(np.uint64), ref["segs"][thr]["lut"][1])


def _oracle(affs, params, block, ctx, roi=None, mask=None):
    from oracle.blockwise import waterz_pipeline
    return waterz_pipeline(affs, params, block_size=block, context=ctx, roi=roi, mask=mask, seed_tie="index",
                           stats_mode="canonical")


CASES = [
    # shape, block, context, params, dtype
    ((20, 160, 160), (10, 80, 80), (2, 10, 10), {}, np.uint8),
    ((24, 130, 170), (10, 64, 64), (1, 8, 8), {}, np.uint8),                                  # ragged: trailing blocks shrink
    ((16, 96, 96), (8, 48, 48), (1, 6, 6), {"fragments_in_xy": False}, np.uint8),             # 3-D seeded mode
    ((14, 45, 31), (7, 23, 16), (1, 3, 3), {"fragments_in_xy": False, "min_seed_distance": 4}, np.uint8),   # odd tile sizes
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"filter_fragments": 0.0, "remove_debris": 0}, np.uint8),
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"min_seed_distance": 5, "thresholds": [0.1, 0.9]}, np.uint8),
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {}, np.float32),
    ((12, 100, 100), (6, 50, 50), (1, 6, 6), {"fragments_in_xy": False}, np.float32),
    # optional shifts (watershed_frags.py:118-139): bias, seed_eps (BASELINE config 4 uses 3-D + seed_eps = 0.01)
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"bias": [-0.05, -0.1, -0.1]}, np.uint8),
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"bias": 0.08, "fragments_in_xy": False}, np.float32),
    ((16, 96, 96), (8, 48, 48), (2, 6, 6), {"fragments_in_xy": False, "seed_eps": 0.01}, np.uint8),
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"seed_eps": 0.02}, np.uint8),
    ((12, 100, 100), (6, 50, 50), (2, 6, 6), {"seed_eps": 0.01, "bias": [-0.02, -0.03, -0.03], "fragments_in_xy": False}, np.float32),
    # sigma: scipy's gaussian_filter replayed bit for bit (float64 for uint8 input, float32 for float32 input)
    ((12, 120, 120), (6, 60, 60), (2, 8, 8), {"sigma": [1, 2, 2]}, np.uint8),
    ((12, 100, 100), (6, 50, 50), (2, 6, 6), {"sigma": [0, 1.5, 0.8], "fragments_in_xy": False}, np.float32),
    ((12, 100, 100), (6, 50, 50), (1, 6, 6), {"sigma": [3, 1, 1], "bias": [-0.05, -0.05, -0.05], "seed_eps": 0.01}, np.uint8),
    # noise_eps: a seeded generator (noise_seed) stands in for the reference's unseeded randn, drawn per block read ROI
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {"noise_eps": 0.05, "noise_seed": 3}, np.uint8),
    ((12, 100, 100), (6, 50, 50), (2, 6, 6), {"noise_eps": 0.02, "sigma": [0, 1, 1], "bias": 0.03, "fragments_in_xy": False}, np.float32),
    # epsilon_agglomerate (watershed_frags.py:158-176): waterz merges of the block's fragments before filter / crop / relabel
    ((12, 120, 120), (6, 60, 60), (2, 8, 8), {"epsilon_agglomerate": 0.1}, np.uint8),
    ((16, 96, 96), (8, 48, 48), (2, 6, 6), {"epsilon_agglomerate": 0.15, "fragments_in_xy": False}, np.float32),
    ((14, 130, 110), (6, 64, 64), (1, 8, 8), {"epsilon_agglomerate": 0.05, "filter_fragments": 0.0, "remove_debris": 0}, np.uint8),   # ragged
]


@pytest.mark.parametrize("shape,block,ctx,params,dtype", CASES)
def test_pipeline_matches_oracle(shape, block, ctx, params, dtype):
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs(shape, seed=1, dtype=dtype)
    _check(_run_gpu(affs, params, block, ctx), _oracle(affs, params, block, ctx))


@pytest.mark.parametrize("dtype,params", [(np.uint8, {"bias": [0.6, 0.6, 0.6]}), (np.float32, {"bias": 0.3, "noise_eps": 0.07, "fragments_in_xy": False})])
def test_empty_blocks_are_skipped_like_the_reference(dtype, params):
    """watershed_frags.py:201-202: a block whose raw affinities are all < 1e-3 is left untouched even when a bias / noise would
    lift its (shifted) affinities over the boundary threshold"""
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((12, 80, 80), seed=6, dtype=dtype)
    affs[:, 0:7, 0:45, 0:45] = 0               # the read ROI of block (0, 0, 0)
    if dtype == np.float32:
        affs[:, 0:7, 0:45, 0:45] = 5e-4        # below the reference's 1e-3
    block, ctx = (6, 40, 40), (1, 5, 5)
    ref = _oracle(affs, params, block, ctx)
    assert not ref["fragments"][:6, :40, :40].any() and ref["fragments"].any()
    _check(_run_gpu(affs, params, block, ctx), ref)


def test_single_block_roi_mode():
    """block_shape == "roi": one block, no context (post/watershed.py:84-86, :361-364)"""
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((8, 150, 150), seed=3)
    _check(_run_gpu(affs, {}, None, None), _oracle(affs, {}, None, None))


def test_roi_inside_volume():
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((20, 140, 150), seed=4)
    roi = ((4, 20, 30), (12, 100, 100))
    _check(_run_gpu(affs, {}, (6, 50, 50), (1, 6, 6), roi=roi), _oracle(affs, {}, (6, 50, 50), (1, 6, 6), roi=roi))


def test_mask():
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((10, 120, 120), seed=5)
    mask = np.ones((10, 120, 120), np.uint8)
    mask[:, 40:70, 30:90] = 0
    mask[3:5] = 0
    _check(_run_gpu(affs, {}, (5, 60, 60), (1, 8, 8), mask_np=mask), _oracle(affs, {}, (5, 60, 60), (1, 8, 8), mask=mask))


def test_epsilon_agglomerate_with_mask_and_shifts():
    """epsilon_agglomerate on masked affinities (affs_data *= mask before the fragments and before waterz) with bias and
    seed_eps shifts of the watershed"""
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((10, 120, 120), seed=6)
    mask = np.ones((10, 120, 120), np.uint8)
    mask[:, 40:70, 30:90] = 0
    mask[3:5] = 0
    p = {"epsilon_agglomerate": 0.12, "bias": [-0.02, -0.05, -0.05], "seed_eps": 0.01}
    _check(_run_gpu(affs, p, (5, 60, 60), (1, 8, 8), mask_np=mask), _oracle(affs, p, (5, 60, 60), (1, 8, 8), mask=mask))


def test_empty_and_saturated_inputs():
    zeros = np.zeros((3, 8, 64, 64), np.uint8)
    r = _run_gpu(zeros, {}, (4, 32, 32), (1, 4, 4))
    assert int(r["fragments"].abs().sum()) == 0 and r["nodes"][0].numel() == 0 and r["edges"][0].numel() == 0
    # all-foreground slices exercise scipy's "background at (-1, 0)" rule for the EDT
    ones = np.full((3, 6, 48, 48), 255, np.uint8)
    ones[:, :, 0, :] = 0          # only the first row of every slice is background in the volume
    p = {"filter_fragments": 0.0, "remove_debris": 0}
    _check(_run_gpu(ones, p, (3, 24, 24), (1, 4, 4)), _oracle(ones, p, (3, 24, 24), (1, 4, 4)))
    full = np.full((3, 4, 40, 40), 255, np.uint8)
    _check(_run_gpu(full, p, None, None), _oracle(full, p, None, None))


def test_edt_and_flood_intermediates():
    """stage-1 intermediates straight from the device scratch: exact squared EDT vs scipy, flood vs the
    restated skimage priority flood (index rule) per tile."""
    from scipy.ndimage import distance_transform_edt
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    from oracle.ws import watershed_from_boundary_distance
    affs = synth_affs((6, 150, 130), seed=7)
    native.set_debug(True)
    try:
        plan = native.Plan(affs.shape[1:], affs.shape[1:], (0, 0, 0), native.BS_DTYPE_U8, filter_fragments=0, remove_debris=0)
        plan.fragments(torch.from_numpy(affs).cuda())
        Z, Y, X = affs.shape[1:]
        stride = (Y * X + 31) & ~31                               # tile bases are 32-aligned
        d2 = plan.debug_fetch("d2", np.uint32).reshape(Z, stride)[:, :Y * X].reshape(Z, Y, X)
        flood = plan.debug_fetch("flood", np.uint32).reshape(Z, stride)[:, :Y * X].reshape(Z, Y, X)
    finally:
        native.set_debug(False)
    a = affs.astype(np.float64) / 255
    for z in range(affs.shape[1]):
        mask = 0.5 * (a[2, z] + a[1, z]) > 0.5
        dist = distance_transform_edt(mask)
        assert np.array_equal(d2[z], np.rint(dist * dist).astype(np.uint32))
        ref, _ = watershed_from_boundary_distance(dist, mask, seed_tie="index")
        got = np.where(flood[z] >= 0x80000000, 0, flood[z]).astype(np.int64)
        pairs = np.unique(np.stack([got.ravel(), ref.ravel().astype(np.int64)], 1), axis=0)
        assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


FRONT_CASES = [
    # shape, block, context, params, dtype, with volume mask
    ((12, 120, 120), (6, 60, 60), (1, 8, 8), {}, np.uint8, False),
    ((12, 128, 160), (6, 64, 80), (1, 8, 8), {}, np.uint8, False),                 # rows 16-byte aligned: the TMA mask kernel
    ((9, 131, 173), (5, 64, 64), (1, 8, 8), {"min_seed_distance": 4}, np.uint8, False),   # ragged blocks, odd sizes
    ((8, 96, 112), (4, 48, 56), (1, 6, 6), {"min_seed_distance": 13, "filter_fragments": 0.0, "remove_debris": 0}, np.uint8, True),
    ((10, 100, 100), (5, 50, 50), (1, 6, 6), {}, np.float32, False),
    ((4, 400, 330), (2, 400, 330), (0, 0, 0), {}, np.uint8, False),               # one large tile per slice (>= 2^17 pixels: unfused only)
    ((4, 300, 320), (2, 300, 320), (0, 0, 0), {}, np.uint8, False),               # one tile per slice, near the shared-memory limit
    ((2, 320, 320), (1, 256, 256), (0, 32, 32), {}, np.uint8, False),             # BASELINE config-5 tiles: 320 x 320, the smallest scratch area
]


@pytest.mark.parametrize("shape,block,ctx,params,dtype,use_mask", FRONT_CASES)
def test_front_versions_agree(shape, block, ctx, params, dtype, use_mask):
    """the fused stage-1 front end (mask bits -> on-chip EDT / seeds / levels -> flood on packed records; vector-load and
    TMA variants) gives the unfused chain's fragments, nodes, edges and segmentations bit for bit -- and the oracle's"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs(shape, seed=11, dtype=dtype)
    mask = None
    if use_mask:
        mask = np.ones(shape, np.uint8)
        mask[:, 30:50, 20:70] = 0
        mask[2] = 0
    res = {}
    for fv in (1, 0, 2, 3):
        try:
            native.set_front_version(fv)
            res[fv] = _run_gpu(affs, params, block, ctx, mask_np=mask)
        finally:
            native.set_front_version(0)
    _check(res[1], _oracle(affs, params, block, ctx, mask=mask))
    for fv in (0, 2, 3):
        assert torch.equal(res[fv]["fragments"], res[1]["fragments"]), f"front version {fv}: fragments"
        for a, b in zip(res[fv]["nodes"], res[1]["nodes"]):
            assert torch.equal(a, b)
        assert all(torch.equal(a, b) for a, b in zip(res[fv]["edges"][:2], res[1]["edges"][:2]))
        assert torch.equal(res[fv]["edges"][2].view(torch.int32), res[1]["edges"][2].view(torch.int32))
        for thr in res[1]["segs"]:
            assert torch.equal(res[fv]["segs"][thr], res[1]["segs"][thr])


def test_front_saturated_tiles():
    """all-foreground tiles (scipy's background-at-(-1, 0) rule, squared distances beyond the fused path's 16-bit planes)
    and empty tiles take the same results through the automatic front-end choice as through the unfused chain"""
    from bootstrapper_b200 import native
    p = {"filter_fragments": 0.0, "remove_debris": 0}
    full = np.full((3, 4, 200, 200), 255, np.uint8)           # d2 up to 200^2 + 200^2: overflows the fused tables -> falls back
    full[:, 1] = 0                                            # an empty slice
    full[:, 2, :, 100:] = 0                                   # half a slice: d2 up to 100^2 fits
    out = {}
    for fv in (1, 0):
        try:
            native.set_front_version(fv)
            out[fv] = _run_gpu(full, p, (2, 200, 200), (0, 0, 0))
        finally:
            native.set_front_version(0)
    assert torch.equal(out[0]["fragments"], out[1]["fragments"])
    _check(out[0], _oracle(full, p, (2, 200, 200), (0, 0, 0)))
    half = full[:, 2:3].copy()
    r = _run_gpu(half, p, (1, 200, 200), (0, 0, 0))          # every tile fits: the fused path itself
    _check(r, _oracle(half, p, (1, 200, 200), (0, 0, 0)))


@pytest.mark.parametrize("shape,block,ctx,params,differs", [
    ((10, 250, 250), (5, 125, 125), (1, 16, 16), {}, True),        # large enough for heap-history seed ties to show
    ((9, 131, 173), (5, 64, 64), (1, 8, 8), {"min_seed_distance": 4}, False),
    ((16, 96, 96), (8, 48, 48), (1, 6, 6), {"fragments_in_xy": False}, True),
])
def test_faithful_flood_matches_heap_oracle(shape, block, ctx, params, differs):
    """bs_set_flood_version(6): skimage's binary heap replayed literally on the device -- the whole pipeline then equals
    the oracle's FAITHFUL mode (seed_tie="heap"), i.e. without declared deviation D1"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import waterz_pipeline
    affs = synth_affs(shape, seed=23)
    try:
        native.set_flood_version(6)
        r = _run_gpu(affs, params, block, ctx)
    finally:
        native.set_flood_version(0)
    ref = waterz_pipeline(affs, params, block_size=block, context=ctx, seed_tie="heap", stats_mode="canonical")
    _check(r, ref)
    if differs:   # the heap order does differ from the index rule on this input (else the case would prove nothing)
        ref_index = waterz_pipeline(affs, params, block_size=block, context=ctx, seed_tie="index", stats_mode="canonical")
        assert not np.array_equal(ref["fragments"], ref_index["fragments"])


def test_flood_versions_agree():
    """the flood kernels (1: global-memory v1, 2: v2 with the bitmap in shared memory, 3: v2 with the bitmap in global
    memory, 4: 3 + level tails in shared memory, 0: automatic choice) are the same function"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from bootstrapper_b200.synth import synth_affs
    affs = torch.from_numpy(synth_affs((6, 200, 200), seed=11)).cuda()
    out = []
    for v in (1, 2, 3, 4, 0):
        native.set_flood_version(v)
        try:
            out.append(segment_blockwise(affs, {}, (3, 100, 100), (1, 12, 12))["fragments"].clone())
        finally:
            native.set_flood_version(0)
    assert all(torch.equal(out[0], o) for o in out[1:]) and int((out[0] != 0).sum()) > 0


def test_flood_block_wide_matches_warp_flood():
    """3-D read ROIs (and slices beyond 2^17 pixels) take the CTA-per-tile flood (v3); version 1 forces the one-warp
    kernel: same fragments, also with steps cut by pixels queued above the current level (seed_eps shifts)"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise, segment_simple
    from bootstrapper_b200.synth import synth_affs
    affs = torch.from_numpy(synth_affs((40, 120, 120), seed=5)).cuda()
    for params in ({"fragments_in_xy": False}, {"fragments_in_xy": False, "seed_eps": 0.01}):
        out = []
        for v in (1, 0):
            native.set_flood_version(v)
            try:
                out.append(segment_blockwise(affs, params, (20, 60, 60), (4, 8, 8))["fragments"].clone())
            finally:
                native.set_flood_version(0)
        assert torch.equal(out[0], out[1]) and int((out[0] != 0).sum()) > 0
    big = torch.from_numpy(synth_affs((2, 400, 400), seed=6)).cuda()     # 160000 pixels per slice > 2^17
    out = []
    for v in (1, 0):
        native.set_flood_version(v)
        try:
            out.append(segment_simple(big, {})["fragments"].clone())
        finally:
            native.set_flood_version(0)
    assert torch.equal(out[0], out[1]) and int((out[0] != 0).sum()) > 0


def test_watershed_from_affinities_plug():
    """post/ws.py:38 signature; partition equality (ids are a relabelling, see the module docstring)"""
    from bootstrapper_b200.post.ws import watershed_from_affinities
    from bootstrapper_b200.synth import synth_affs
    from oracle.ws import watershed_from_affinities as ref_ws
    affs = synth_affs((5, 140, 140), seed=8)
    for xy in (True, False):
        got, n = watershed_from_affinities(affs, max_affinity_value=255, fragments_in_xy=xy, min_seed_distance=10)
        ref, _ = ref_ws(affs.astype(np.float64) / 255, fragments_in_xy=xy, min_seed_distance=10, seed_tie="index")
        assert got.dtype == np.uint64 and n == len(np.unique(got)) - 1
        pairs = np.unique(np.stack([got.ravel().astype(np.int64), ref.ravel().astype(np.int64)], 1), axis=0)
        assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))
        # float64 affinities (what the reference's own callers hand to the plug): the mask comes from the reference's float64
        # expression, values that sit exactly on / next to the 0.5 threshold included
        a64 = affs.astype(np.float64) / 255
        a64[1, :, ::7, ::5] = 0.5
        a64[2, :, ::7, ::5] = np.nextafter(0.5, 1.0)
        got64, n64 = watershed_from_affinities(a64, fragments_in_xy=xy, min_seed_distance=10)
        ref64, _ = ref_ws(a64, fragments_in_xy=xy, min_seed_distance=10, seed_tie="index")
        pairs = np.unique(np.stack([got64.ravel().astype(np.int64), ref64.ravel().astype(np.int64)], 1), axis=0)
        assert len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1])) and n64 == len(np.unique(got64)) - 1


def test_connected_components_and_relabel_kernels():
    from bootstrapper_b200 import native
    from oracle.native import connected_components as ref_cc
    rng = np.random.default_rng(0)
    nodes = np.sort(rng.choice(np.arange(1, 10 ** 6), 5000, replace=False)).astype(np.uint64)
    e = rng.choice(nodes, (12000, 2))
    scores = rng.random(12000).astype(np.float32)
    scores[::7] = np.nan
    dev = "cuda"
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int64) if x.dtype == np.uint64 else x).to(dev)  # noqa: E731
    for thr in (0.0, 0.3, 0.3000001, 1.0):
        comp = native.connected_components(t(nodes), t(e[:, 0].copy()), t(e[:, 1].copy()), t(scores), thr)
        keep = ~np.isnan(scores)
        ref = ref_cc(nodes, e[keep], scores[keep], thr)
        assert np.array_equal(comp.cpu().numpy().view(np.uint64), ref)
    frags = rng.choice(np.concatenate([[0, 999999999], nodes]), (40, 50, 60)).astype(np.uint64)
    seg = native.relabel(t(frags), t(nodes), comp).cpu().numpy().view(np.uint64)
    lut = dict(zip(nodes.tolist(), comp.cpu().numpy().view(np.uint64).tolist()))
    ref = np.vectorize(lambda v: lut.get(v, v), otypes=[np.uint64])(frags)
    assert np.array_equal(seg, ref)
    # idempotence: component ids are fixed points of the LUT
    assert np.array_equal(native.relabel(t(seg), t(nodes), comp).cpu().numpy().view(np.uint64), seg)


def test_full_size_properties():
    """CREMI-sized slab (BASELINE config 2 geometry, reduced z) through size-independent properties."""
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    affs = native.synth_affs((50, 1250, 1250), seed=0)
    r = segment_blockwise(affs, {}, (25, 250, 250), (3, 31, 31))
    torch.cuda.synchronize()
    frags = r["fragments"]
    ids, pos, sizes = r["nodes"]
    assert bool((ids[1:] > ids[:-1]).all())                                        # ascending unique ids
    assert int(sizes.sum()) == int((frags != 0).sum())                              # node sizes partition the foreground
    uniq = torch.unique(frags)
    assert torch.equal(uniq[uniq != 0], ids)
    nvox = 25 * 250 * 250
    bid = torch.div(ids, nvox, rounding_mode="floor")
    plan_ids, wo, ws = r["plan"].block_info()
    assert set(bid.unique().tolist()) <= set(plan_ids.tolist())
    eu, ev, es = r["edges"]
    assert bool((eu < ev).all()) and torch.unique(torch.stack([eu, ev]), dim=1).shape[1] == eu.numel()
    ok = ~torch.isnan(es)
    assert bool(((es[ok] >= 0) & (es[ok] < 1)).all())
    prev = None
    for thr in sorted(r["segs"]):
        seg = r["segs"][thr]
        assert torch.equal(seg == 0, frags == 0)
        n = torch.unique(seg).numel()
        assert prev is None or n <= prev                                            # coarser with the threshold
        prev = n
        # every segment id is the smallest fragment id it contains
        assert bool((seg <= frags).all())
    # the kernel variants chosen by batch size (flood with the bitmap in shared / global memory, level tables in shared
    # memory; agglomeration sequential / parallel merges) are the same function at this size too
    for fv, av in ((2, 0), (4, 4), (3, 3)):
        try:
            native.set_flood_version(fv)
            native.set_agglom_version(av)
            r2 = segment_blockwise(affs, {}, (25, 250, 250), (3, 31, 31))
        finally:
            native.set_flood_version(0)
            native.set_agglom_version(0)
        assert torch.equal(r2["fragments"], frags)
        assert all(torch.equal(a, b) for a, b in zip(r2["edges"][:2], r["edges"][:2]))
        assert torch.equal(r2["edges"][2].view(torch.int32), es.view(torch.int32))      # NaN-safe bit comparison
        for thr in r["segs"]:
            assert torch.equal(r2["segs"][thr], r["segs"][thr])


# ---------------------------------------------------------------- single-shot path (BASELINE config 1)
def _same_partition(a, b):
    a = a.ravel().astype(np.int64)
    b = b.ravel().astype(np.int64)
    if ((a == 0) != (b == 0)).any():
        return False
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) and len(np.unique(pairs[:, 1])) == len(pairs)


SIMPLE_CASES = [
    ((10, 128, 128), {"sigma": [1, 2, 2], "bias": [-0.05, -0.1, -0.1]}, np.uint8),      # shifted affinities feed waterz too
    ((8, 96, 96), {"sigma": [0, 1, 1.5]}, np.float32),
    ((10, 128, 128), {}, np.float32),
    ((10, 128, 128), {}, np.uint8),
    ((8, 96, 96), {"fragments_in_xy": False, "thresholds": [0.5, 0.1, 0.3]}, np.float32),    # unsorted thresholds
    ((6, 140, 90), {"min_seed_distance": 6, "thresholds": [0.05, 0.95]}, np.uint8),
    # histogram-quantile scoring functions (post/watershed.py:232-244)
    ((10, 128, 128), {"merge_function": "hist_quant_50"}, np.uint8),
    ((10, 128, 128), {"merge_function": "hist_quant_75_initmax", "thresholds": [0.3, 0.5, 0.8]}, np.float32),
    ((8, 96, 96), {"merge_function": "hist_quant_10", "fragments_in_xy": False, "thresholds": [0.6, 0.9]}, np.float32),
    ((8, 96, 96), {"merge_function": "hist_quant_90", "bias": [-0.3, -0.4, -0.4], "thresholds": [0.4, 0.99]}, np.uint8),   # bins clamp at 0
    ((6, 140, 90), {"merge_function": "hist_quant_25_initmax", "sigma": [0, 1, 1], "thresholds": [0.7]}, np.uint8),
]


@pytest.mark.parametrize("shape,params,dtype", SIMPLE_CASES)
def test_simple_watershed_matches_oracle(shape, params, dtype):
    """simple_watershed (post/watershed.py:206-354): waterz with the default queue, (score, edge id) order.
    Fragment ids of the reference count masked-out seed plateaus, the CUDA path numbers fragments in raster
    order: partitions are compared label-permutation-invariantly, segment ids through the fragment map."""
    from bootstrapper_b200.post.pipeline import segment_simple
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import simple_watershed
    affs = synth_affs(shape, seed=3, dtype=dtype)
    ref = simple_watershed(affs, params, seed_tie="index", stats_mode="canonical")
    r = segment_simple(torch.from_numpy(affs).to("cuda:0"), params)
    torch.cuda.synchronize()
    f = r["fragments"].cpu().numpy().view(np.uint64)
    assert _same_partition(f, ref["fragments"]), "fragment partitions differ"
    assert set(r["segs"]) == set(ref["params"]["thresholds"])
    for thr in ref["params"]["thresholds"]:
        got = r["segs"][thr].cpu().numpy().view(np.uint64)
        assert _same_partition(got, ref["segs"][thr]), f"segmentation at {thr} differs"
        # a segment is a union of fragments: the segmentation must be constant on every fragment
        assert len(np.unique(np.stack([f.ravel(), got.ravel()], 1), axis=0)) == len(np.unique(f))


def test_simple_watershed_cremi_crop():
    """BASELINE config 1: 3x(50,512,512) float32, single block (slices of 512^2 take the global-memory flood)."""
    from bootstrapper_b200.post.pipeline import segment_simple
    from bootstrapper_b200 import native
    from oracle.blockwise import simple_watershed
    shape = (50, 512, 512)
    affs_t = native.synth_affs(shape, seed=0, dtype=torch.float32)
    affs = affs_t.cpu().numpy()
    ref = simple_watershed(affs, {}, seed_tie="index", stats_mode="canonical")
    r = segment_simple(affs_t, {})
    torch.cuda.synchronize()
    assert _same_partition(r["fragments"].cpu().numpy().view(np.uint64), ref["fragments"])
    for thr in ref["params"]["thresholds"]:
        assert _same_partition(r["segs"][thr].cpu().numpy().view(np.uint64), ref["segs"][thr]), thr


# ---------------------------------------------------------------- cc method
def test_cc_affs_golden_and_oracle():
    """bs_cc_affs against outputs of the reference's own post/cc.py (golden) and, at a larger size, the oracle."""
    import os
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    from oracle import cc as occ
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gdir, "cc_flood.npz"))
    for ci in range(3):
        affs = torch.from_numpy(g[f"c{ci}_hard"].astype(np.float32)).cuda()
        frags, seg, n = native.cc_affs(affs, 0.5)
        assert np.array_equal(frags.cpu().numpy(), g[f"c{ci}_seg"].astype(np.int64))
        assert n == int(g[f"c{ci}_seg"].max())
    g = np.load(os.path.join(gdir, "cc_affs.npz"))
    for ci in range(3):
        mask = torch.from_numpy(g[f"c{ci}_mask"]).cuda() if g[f"c{ci}_mask"].size else None
        frags, seg, n = native.cc_affs(torch.from_numpy(g[f"c{ci}_affs"]).cuda(), float(g[f"c{ci}_thr"]), 0, mask)
        assert np.array_equal(frags.cpu().numpy(), g[f"c{ci}_seg"].astype(np.int64))
    for dtype, thr, rd in [(np.uint8, 0.5, 64), (np.float32, 0.9, 5), (np.uint8, 0.0, 0)]:
        affs = synth_affs((12, 150, 170), seed=4, dtype=dtype)
        rf, rs = occ.cc_affs(affs, thr, rd)
        frags, seg, n = native.cc_affs(torch.from_numpy(affs).cuda(), thr, rd)
        assert np.array_equal(frags.cpu().numpy(), rf.astype(np.int64))
        assert np.array_equal(seg.cpu().numpy(), rs.astype(np.int64))
    # sigma shift (connected_components.py:73-77) through the in-memory driver
    from bootstrapper_b200.post.connected_components import cc_in_memory
    affs = synth_affs((10, 90, 110), seed=8, dtype=np.uint8)
    mask = (np.random.default_rng(3).random((10, 90, 110)) < 0.9).astype(np.uint8)
    rf, rs = occ.cc_affs(affs, 0.55, 20, mask, sigma=[1, 1.5, 1.5])
    frags, seg = cc_in_memory(torch.from_numpy(affs).cuda(), 0.55, 20, torch.from_numpy(mask).cuda(), sigma=[1, 1.5, 1.5])
    assert np.array_equal(frags.cpu().numpy(), rf.astype(np.int64)) and np.array_equal(seg.cpu().numpy(), rs.astype(np.int64))


def test_shift_affinities_matches_scipy():
    """bs_shift_affinities against numpy + scipy.ndimage.gaussian_filter, bit for bit (float32 semantics of the single-shot paths)"""
    from scipy.ndimage import gaussian_filter
    from bootstrapper_b200 import native
    rng = np.random.default_rng(0)
    for dtype in (np.uint8, np.float32):
        a = (rng.random((3, 9, 40, 37)) * (255 if dtype == np.uint8 else 1)).astype(dtype)
        mask = (rng.random((9, 40, 37)) < 0.8).astype(np.uint8)
        for sigma, bias in [([1, 2, 2], None), ([0, 0.7, 3.3], [-0.1, 0.05, 0.2]), (None, 0.25), ([4, 0, 0], None)]:
            data = a.astype(np.float32) / 255.0 if dtype == np.uint8 else a.astype(np.float32)
            data = data * (mask > 0).astype(np.uint8)
            shift = np.zeros_like(data)
            if sigma is not None:
                shift += gaussian_filter(data, sigma=(0, *sigma)) - data
            if bias is not None:
                b = [bias] * 3 if isinstance(bias, float) else bias
                shift += np.array([b]).reshape((-1, 1, 1, 1))
            want = data + shift
            got = native.shift_affinities(torch.from_numpy(a).cuda(), torch.from_numpy(mask).cuda(), sigma, bias).cpu().numpy()
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dtype, sigma, bias)


def test_agglomeration_kernels_agree():
    """every form of the agglomeration kernel (default: parallel merges in shared memory) gives the same graph"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((20, 160, 160), seed=5)
    block, ctx = (10, 80, 80), (2, 10, 10)
    ref = _oracle(affs, {}, block, ctx)
    for version in (3, 4):   # parallel merges on global slabs: hybrid (shared-memory union-find / bins) and plain
        try:
            native.set_agglom_version(version)
            r = _run_gpu(affs, {}, block, ctx)
        finally:
            native.set_agglom_version(0)
        _check(r, ref)


# ---------------------------------------------------------------- affinity self-consistency error
@pytest.mark.parametrize("shape,dtype,use_mask,nhood", [
    ((10, 90, 70), np.float32, False, [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]),
    ((10, 90, 70), np.uint8, True, [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -8, 0], [0, 0, -8], [0, 5, -3], [1, 0, 0], [0, 0, 4]]),
    ((7, 45, 31), np.uint8, True, [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, -3, 2]]),       # odd voxel count: the scalar kernels
    ((7, 45, 31), np.float32, False, [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]),
])
def test_aff_errors_matches_oracle(shape, dtype, use_mask, nhood):
    """AddAffErrors.process (gp/add_aff_errors.py:128-183): float32 error map and masks bit for bit"""
    from bootstrapper_b200.eval import add_aff_errors
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from bootstrapper_b200.synth import synth_affs
    from oracle import aff_errors as oa
    affs = synth_affs(shape, seed=4)
    r = segment_blockwise(torch.from_numpy(affs).cuda(), {}, tuple(-(-v // 2) for v in shape), (1, 5, 5))
    seg = r["segs"][0.35].cpu().numpy().view(np.uint64)
    rng = np.random.default_rng(9)
    pred = rng.integers(0, 256, (len(nhood),) + seg.shape, dtype=np.uint8)
    if dtype == np.float32:
        pred = (pred.astype(np.float32) / np.float32(255)).astype(np.float32)
    mask = (rng.random(seg.shape) < 0.85).astype(np.uint8) if use_mask else None
    ref_affs, ref_err, ref_mask = oa.aff_errors(seg, pred, nhood, mask, (0.1, 1.0))
    got = add_aff_errors(r["segs"][0.35], torch.from_numpy(pred).cuda(), nhood,
                         None if mask is None else torch.from_numpy(mask).cuda(), (0.1, 1.0))
    assert np.array_equal(got["seg_affs"].cpu().numpy(), ref_affs)
    assert np.array_equal(got["error_map"].cpu().numpy().view(np.uint32), ref_err.view(np.uint32))
    assert np.array_equal(got["error_mask"].cpu().numpy(), ref_mask)
    assert np.array_equal(got["error_map_u8"].cpu().numpy(), oa.error_map_u8(ref_err))
    assert ref_mask.any() and not ref_mask.all()
    # a perfect prediction: zero error everywhere, empty mask (the max == 0 branch)
    z = add_aff_errors(r["segs"][0.35], torch.from_numpy(ref_affs).cuda(), nhood)
    assert not z["error_map"].any() and not z["error_mask"].any()


def test_run_host_streaming_matches_blocking():
    """ShardedSegmenter.run_host(wait=False): volumes streamed through two buffer sets give the blocking call's bytes"""
    from bootstrapper_b200.sharded import ShardedSegmenter
    from bootstrapper_b200.synth import synth_affs
    shape, block, ctx = (12, 120, 120), (6, 60, 60), (1, 8, 8)
    seg = ShardedSegmenter(shape, block, ctx, {"thresholds": [0.2, 0.5]}, device=torch.device("cuda"))
    vols = [torch.from_numpy(synth_affs(shape, seed=s)).pin_memory() for s in (1, 2, 3, 4, 5)]
    ref = []
    for v in vols:
        ho = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(3)]
        seg.run_host(v, ho)
        ref.append([h.clone() for h in ho])
    host_sets = [[torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(3)] for _ in range(2)]
    dev_sets = [[torch.empty(shape, dtype=torch.int64, device="cuda") for _ in range(2)] for _ in range(2)]
    got = []
    for k, v in enumerate(vols):
        if k >= 2:
            # the set is about to be reused: its volume (k - 2) must be read first
            seg.drain()
            got.append([h.clone() for h in host_sets[k % 2]])
        seg.run_host(v, host_sets[k % 2], out=dev_sets[k % 2], wait=False)
    seg.drain()
    got.append([h.clone() for h in host_sets[len(vols) % 2]])
    got.append([h.clone() for h in host_sets[(len(vols) + 1) % 2]])
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert int((ref[0][0] != 0).sum()) > 0 and not torch.equal(ref[0][1], ref[1][1])


def test_run_host_ring_matches_blocking():
    """ShardedSegmenter.run_host(ring=HostRing): results streamed through a small ring of page-locked z-chunk buffers (a
    few planes per chunk, fewer slots than chunks per volume) reach the sink complete and in the blocking call's bytes"""
    from bootstrapper_b200.sharded import HostRing, ShardedSegmenter
    from bootstrapper_b200.synth import synth_affs
    shape, block, ctx = (12, 120, 120), (6, 60, 60), (1, 8, 8)
    thrs = [0.2, 0.5]
    seg = ShardedSegmenter(shape, block, ctx, {"thresholds": thrs}, device=torch.device("cuda"))
    vols = [torch.from_numpy(synth_affs(shape, seed=s)).pin_memory() for s in (1, 2, 3, 4)]
    ref = []
    for v in vols:
        ho = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(3)]
        seg.run_host(v, ho)
        ref.append([h.clone() for h in ho])
    got = {}
    counter = {"vol": {}}

    def sink(name, z0, z1, view):
        # chunks of one array arrive in z order, arrays of one volume in submission order, volumes in order
        n = counter["vol"].get(name, 0)
        arr = got.setdefault((name, n), torch.zeros(shape, dtype=torch.int64))
        arr[z0:z1] = view
        if z1 == shape[0]:
            counter["vol"][name] = n + 1

    ring = HostRing(shape, torch.device("cuda"), chunk_bytes=5 * 120 * 120 * 8, n_slots=4, sink=sink)
    assert ring.planes == 5 and ring.pinned_bytes == 4 * 5 * 120 * 120 * 8
    dev_sets = [[torch.empty(shape, dtype=torch.int64, device="cuda") for _ in thrs] for _ in range(2)]
    for k, v in enumerate(vols):
        seg.run_host(v, None, out=dev_sets[k % 2], wait=False, ring=ring)
    seg.drain()
    ring.close()
    for k in range(len(vols)):
        assert torch.equal(got[(("fragments", None), k)], ref[k][0])
        for i, t in enumerate(thrs):
            assert torch.equal(got[(("seg", t), k)], ref[k][1 + i])
    assert ring.chunks_done == len(vols) * 3 * 3


def test_run_host_compact_expands_to_the_blocking_result():
    """ShardedSegmenter.run_host_compact: one int32 plane + node table + LUT rows cross the bus; the host decoder
    (bs_expand_compact) rebuilds exactly the uint64 arrays run_host delivers"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.sharded import ShardedSegmenter
    from bootstrapper_b200.synth import synth_affs
    shape, block, ctx = (12, 120, 120), (6, 60, 60), (1, 8, 8)
    thrs = [0.2, 0.5]
    seg = ShardedSegmenter(shape, block, ctx, {"thresholds": thrs}, device=torch.device("cuda"))
    cap = 1 << 16
    for s_ in (1, 2):
        v = torch.from_numpy(synth_affs(shape, seed=s_)).pin_memory()
        ho = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(3)]
        seg.run_host(v, ho)
        comp = dict(dense=torch.empty(shape, dtype=torch.int32).pin_memory(), nodes=torch.empty(cap, dtype=torch.int64).pin_memory(),
                    luts=[torch.empty(cap, dtype=torch.int64).pin_memory() for _ in thrs])
        info = seg.run_host_compact(v, comp)
        n = info["n_nodes"]
        assert n > 0 and int(comp["dense"].max()) == n
        frags, segs = native.expand_compact(comp["dense"], comp["nodes"][:n].contiguous(), [l[:n].contiguous() for l in comp["luts"]], threads=3)
        assert torch.equal(frags, ho[0])
        for a, b in zip(segs, ho[1:]):
            assert torch.equal(a, b)


def test_host_expander_pipeline_delivers_every_volume():
    """the expanded host path: compact transfers of several volumes in flight (two buffer sets), a decoder thread rebuilding the
    uint64 arrays in host memory while the device works on -- every decoded volume equals the blocking result"""
    from bootstrapper_b200.sharded import HostExpander, ShardedSegmenter
    from bootstrapper_b200.synth import synth_affs
    shape, block, ctx = (12, 120, 120), (6, 60, 60), (1, 8, 8)
    thrs = [0.2, 0.5]
    seg = ShardedSegmenter(shape, block, ctx, {"thresholds": thrs}, device=torch.device("cuda"))
    cap = 1 << 16
    vols = [torch.from_numpy(synth_affs(shape, seed=s_)).pin_memory() for s_ in (1, 2, 3, 4, 5)]
    want = []
    for v in vols:
        ho = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(3)]
        seg.run_host(v, ho)
        want.append([h.clone() for h in ho])
    sets = [dict(dense=torch.empty(shape, dtype=torch.int32).pin_memory(), nodes=torch.empty(cap, dtype=torch.int64).pin_memory(),
                 luts=[torch.empty(cap, dtype=torch.int64).pin_memory() for _ in thrs]) for _ in range(2)]
    outs = [[torch.zeros(shape, dtype=torch.int64) for _ in range(3)] for _ in vols]
    exp = HostExpander(threads=3)
    try:
        for k, v in enumerate(vols):
            exp.acquire()
            info = seg.run_host_compact(v, sets[k % 2], wait=False)
            exp.submit(info["done"], sets[k % 2], info["n_nodes"], outs[k])
        exp.flush()
        seg.drain()
    finally:
        exp.close()
    for got, ref in zip(outs, want):
        for a, b in zip(got, ref):
            assert torch.equal(a, b)


# ---------------------------------------------------------------- `bs refine` filters
def test_refine_filters_match_oracle():
    """per-id table (sizes, z-extents) and the four filters of refine.py against the numpy restatement"""
    from bootstrapper_b200 import native, refine
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from bootstrapper_b200.synth import synth_affs
    from oracle import refine as orf
    affs = synth_affs((14, 150, 130), seed=8)
    r = segment_blockwise(torch.from_numpy(affs).cuda(), {}, (7, 75, 65), (1, 9, 9))
    for seg_t in (r["fragments"], r["segs"][0.5]):
        seg = seg_t.cpu().numpy().view(np.uint64)
        uniq, sizes = refine.global_sizes(seg_t)
        ou, os_ = orf.global_sizes(seg)
        assert np.array_equal(uniq, ou) and np.array_equal(sizes, os_) and uniq.size > 10
        # a table that starts too small grows until every id fits
        ids_small = native.label_stats(seg_t.contiguous(), capacity=8)[0]
        assert np.array_equal(ids_small.cpu().numpy().view(np.uint64), ou)
        got, rem = refine.size_filter(seg_t, min_size=int(np.median(sizes)), max_size=int(np.percentile(sizes, 90)))
        ref, orem = orf.size_filter(seg, min_size=int(np.median(sizes)), max_size=int(np.percentile(sizes, 90)))
        assert np.array_equal(rem, orem) and 0 < rem.size < uniq.size
        assert np.array_equal(got.cpu().numpy().view(np.uint64), ref)
        got, rem, st = refine.outlier_filter(seg_t, num_std=1.0, min_size=5)
        ref, orem = orf.outlier_filter(seg, num_std=1.0, min_size=5)
        assert np.array_equal(rem, orem) and rem.size > 0
        assert np.array_equal(got.cpu().numpy().view(np.uint64), ref)
        got, rem = refine.z_filter(seg_t, min_z=2)
        ref, orem = orf.z_filter(seg, min_z=2)
        assert np.array_equal(np.sort(rem), np.sort(orem))
        assert np.array_equal(got.cpu().numpy().view(np.uint64), ref)
        a, b, c, d = (int(v) for v in uniq[[0, 3, 5, 7]])
        got = refine.remap(seg_t, remove_ids=[a], merge_groups=[[b, c, d]])
        assert np.array_equal(got.cpu().numpy().view(np.uint64), orf.remap(seg, {a: 0, b: b, c: b, d: b}))
    with pytest.raises(ValueError):
        refine.remap(r["fragments"], remove_ids=[5], merge_groups=[[5, 6]])
    empty = torch.zeros((3, 8, 8), dtype=torch.int64, device="cuda")
    assert refine.global_sizes(empty)[0].size == 0


def test_two_host_threads_do_not_share_scratch():
    """the library's scratch arena / profiler / launch counter are per host thread: two threads driving two plans on two
    streams at the same time get the results of the serial runs (process-global scratch)"""
    import threading
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    shape, block, ctx = (24, 260, 260), (12, 130, 130), (2, 16, 16)
    affs = [native.synth_affs(shape, seed=s) for s in (1, 2)]
    serial = [segment_blockwise(a, {}, block, ctx) for a in affs]
    torch.cuda.synchronize()
    got, errs = [None, None], []

    def work(k):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(4):
                    r = segment_blockwise(affs[k], {}, block, ctx)
                st.synchronize()
            got[k] = r
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in range(2):
        assert torch.equal(got[k]["fragments"], serial[k]["fragments"])
        for a, b in zip(got[k]["edges"], serial[k]["edges"]):
            assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)
        for thr in serial[k]["segs"]:
            assert torch.equal(got[k]["segs"][thr], serial[k]["segs"][thr])
