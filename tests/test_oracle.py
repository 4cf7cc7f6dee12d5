"""The oracle itself: pinned pieces against fixtures generated from the reference's own files, and
hand-checkable micro cases for the restated third-party routines (parity unpinned, SURVEY U-list)."""
import heapq
import os

import numpy as np
import pytest

from oracle import blockwise as ob
from oracle.merge_tree import MergeTree
from oracle.native import Waterz, connected_components, sk_label, sk_watershed

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_merge_tree_pinned_against_reference():
    d = np.load(os.path.join(GOLD, "merge_tree.npz"))
    n_cases = len([k for k in d.files if k.endswith("_leaves")])
    assert n_cases >= 4
    for ci in range(n_cases):
        mt = MergeTree(d[f"c{ci}_leaves"])
        for a, b, c, s in d[f"c{ci}_hist"]:
            mt.merge(int(a), int(b), int(c), s)
        out = mt.find_merges(d[f"c{ci}_us"], d[f"c{ci}_vs"])
        ref = d[f"c{ci}_out"]
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(out[~np.isnan(out)], ref[~np.isnan(ref)])


def test_cantor_numbers():
    # funlib.math.cantor_number: 1-D identity, 2-D Cantor pairing, 3-D simplex pairing
    assert [ob.cantor_number((i,)) for i in range(4)] == [0, 1, 2, 3]
    assert [ob.cantor_number(c) for c in [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]] == [0, 2, 1, 5, 4, 3]
    ids = {ob.cantor_number((i, j, k)) for i in range(6) for j in range(6) for k in range(6)}
    assert len(ids) == 216 and ob.cantor_number((0, 0, 0)) == 0


# ---- skimage.segmentation.watershed restated
def _ws1d(image, markers, mask=None, tie="heap"):
    image = np.asarray(image, np.float64)[None]
    markers = np.asarray(markers, np.int64)[None]
    mask = np.ones_like(markers, np.uint8) if mask is None else np.asarray(mask, np.uint8)[None]
    return sk_watershed(image, markers, mask, seed_tie=tie)[0].tolist()


@pytest.mark.parametrize("tie", ["heap", "index"])
def test_watershed_micro_cases(tie):
    assert _ws1d([0, 1, 2, 1, 0], [1, 0, 0, 0, 2], tie=tie) == [1, 1, 1, 2, 2]          # FIFO: lower age claims the ridge
    assert _ws1d([0, 0, 0, 0], [1, 0, 0, 2], tie=tie) == [1, 1, 2, 2]                    # plateau split by age
    assert _ws1d([0, 5, 1, 5, 0], [1, 0, 0, 0, 2], tie=tie) == [1, 1, 1, 2, 2]           # un-seeded pit, value not clamped
    assert _ws1d([0, 0, 0, 0, 0], [1, 0, 0, 0, 0], mask=[1, 1, 0, 1, 1], tie=tie) == [1, 1, 0, 0, 0]   # mask blocks
    assert _ws1d([0, 0, 0], [7, 0, 0], mask=[0, 1, 1], tie=tie) == [0, 0, 0]             # markers * mask
    out = sk_watershed(np.zeros((3, 3)), np.array([[0, 1, 0], [2, 0, 0], [0, 0, 0]]), np.ones((3, 3), np.uint8), seed_tie=tie)
    # neighbour order (-row, -col, +col, +row): seed 1 pops first and takes the centre
    assert out.tolist() == [[1, 1, 1], [2, 1, 1], [2, 1, 1]]


def _flood_python(image, markers, mask):
    """independent restatement with heapq; valid for the 'index' rule where (value, age) is a strict order"""
    shape = image.shape
    out = (markers * mask).astype(np.int64)
    seeds = np.flatnonzero(out.ravel())
    heap = [(image.ravel()[i], k - len(seeds), int(i)) for k, i in enumerate(seeds)]
    heapq.heapify(heap)
    age = 1
    nd = image.ndim
    strides = [int(np.prod(shape[d + 1:])) for d in range(nd)]
    offs = [(-1, d) for d in range(nd)] + [(1, d) for d in reversed(range(nd))]
    flat, m, im = out.ravel(), mask.ravel(), image.ravel()
    while heap:
        _, _, i = heapq.heappop(heap)
        c = np.unravel_index(i, shape)
        for sgn, d in offs:
            if not 0 <= c[d] + sgn < shape[d]:
                continue
            j = i + sgn * strides[d]
            if not m[j] or flat[j]:
                continue
            age += 1
            flat[j] = flat[i]
            heapq.heappush(heap, (im[j], age, j))
    return out


@pytest.mark.parametrize("shape", [(17, 23), (5, 9, 11)])
def test_watershed_index_rule_matches_independent_restatement(shape):
    rng = np.random.default_rng(3)
    for _ in range(5):
        image = rng.integers(0, 4, shape).astype(np.float64)       # few levels -> ties everywhere
        markers = (rng.random(shape) < 0.03) * rng.integers(1, 50, shape)
        mask = (rng.random(shape) < 0.9).astype(np.uint8)
        assert np.array_equal(sk_watershed(image, markers, mask, seed_tie="index"), _flood_python(image, markers, mask))


def test_watershed_heap_vs_index_census():
    """D1: the two seed tie rules differ only on contested voxels between equal-valued seeds."""
    from bootstrapper_b200.synth import synth_affs
    from oracle.ws import watershed_from_affinities
    affs = synth_affs((4, 160, 160), seed=5).astype(np.float64) / 255.0
    a, _ = watershed_from_affinities(affs, fragments_in_xy=True, seed_tie="heap")
    b, _ = watershed_from_affinities(affs, fragments_in_xy=True, seed_tie="index")
    assert np.array_equal(a == 0, b == 0)
    frac = float((a != b).mean())
    assert frac < 0.02, frac


def test_label_full_connectivity_raster_ids():
    x = np.array([[5, 0, 5], [0, 5, 0], [7, 0, 0], [0, 0, 5]])
    out, n = sk_label(x)
    assert n == 3 and out.tolist() == [[1, 0, 1], [0, 1, 0], [2, 0, 0], [0, 0, 3]]
    x3 = np.zeros((2, 2, 2), np.int64)
    x3[0, 0, 0] = x3[1, 1, 1] = 4                                   # 26-connectivity joins the corner pair
    assert sk_label(x3)[1] == 1


# ---- waterz restated
def _vol(frags, ax=None, ay=None, az=None):
    frags = np.asarray(frags, np.uint64)
    affs = np.zeros((3,) + frags.shape, np.uint8)
    for c, a in enumerate((az, ay, ax)):
        if a is not None:
            affs[c] = np.asarray(a, np.uint8)
    return affs, frags


def test_waterz_chain_and_thresholds():
    affs, frags = _vol([[[1, 2, 3, 3]]], ax=[[[0, 200, 100, 255]]])
    wz = Waterz(affs, frags, 256, "canonical")
    a, b, c, s = wz.merge_until(0.5)
    assert (a.tolist(), b.tolist(), c.tolist()) == ([1], [2], [1]) and s[0] == np.float32(1 - np.float32(200 / 255))
    assert wz.segmentation().ravel().tolist() == [1, 1, 3, 3]
    a, b, c, s = wz.merge_until(1.0)
    assert (a.tolist(), b.tolist()) == ([1], [3]) and abs(s[0] - (1 - 100 / 255)) < 1e-6
    assert wz.counters()["stale"] == 1                                 # (2,3) was re-scored once after the first merge


def test_waterz_shared_neighbour_keeps_cheaper_edge_and_rescoring():
    affs, frags = _vol([[[1, 2], [3, 3]]], ax=[[[0, 250], [0, 0]]], ay=[[[0, 0], [100, 200]]])
    wz = Waterz(affs, frags, 256, "canonical")
    wz.merge_until(0.0)
    u, v, s, _, cnt = wz.region_graph()
    assert sorted(zip(u.tolist(), v.tolist())) == [(1, 2), (1, 3), (2, 3)]
    assert [(int(a), int(b)) for a, b in zip(u, v)] == [(1, 2), (1, 3), (2, 3)]     # creation order: raster, then z,y,x
    a, b, c, s = wz.merge_until(1.0)
    assert list(zip(a.tolist(), b.tolist())) == [(1, 2), (1, 3)]
    assert abs(s[1] - (1 - 150 / 255)) < 1e-6                           # merged edge: mean of both contacts
    mt = MergeTree(np.array([0, 1, 2, 3], np.uint64))
    for ai, bi, ci, si in zip(a, b, c, s):
        mt.merge(ai, bi, ci, si)
    out = mt.find_merges([1, 1, 2], [2, 3, 3])
    assert out[0] == s[0] and out[1] == s[1] and out[2] == s[1]


def test_waterz_zero_affinity_edges_never_merge():
    affs, frags = _vol([[[1, 2, 3]]], ax=[[[0, 0, 255]]])
    wz = Waterz(affs, frags, 256, "canonical")
    a, b, _, s = wz.merge_until(1.0)
    assert list(zip(a.tolist(), b.tolist())) == [(2, 3)]               # score(1,2) == 1.0 is not < 1.0


def test_connected_components_threshold_is_inclusive():
    nodes = np.array([3, 5, 9, 11], np.uint64)
    edges = np.array([[3, 5], [5, 9], [9, 11]], np.uint64)
    scores = np.array([0.2, 0.35, 0.5], np.float32)
    assert connected_components(nodes, edges, scores, 0.35).tolist() == [3, 3, 3, 11]
    assert connected_components(nodes, edges, scores, 0.34999).tolist() == [3, 3, 9, 11]


def test_blockwise_oracle_is_block_order_independent():
    from bootstrapper_b200.synth import synth_affs
    from oracle.parallel import waterz_pipeline_parallel
    affs = synth_affs((12, 100, 100), seed=2)
    kw = dict(block_size=(6, 50, 50), context=(1, 6, 6), seed_tie="index", stats_mode="canonical")
    a = ob.waterz_pipeline(affs, **kw)
    b = waterz_pipeline_parallel(affs, workers=2, **kw)
    assert np.array_equal(a["fragments"], b["fragments"]) and a["rag"].edges == b["rag"].edges
    assert len(a["rag"].node_pos) > 20 and len(a["rag"].edges) > 20


def test_cc_restatement_matches_reference_goldens():
    """oracle/cc.py against outputs of the reference's own post/cc.py (tests/golden/make_golden.py)."""
    from oracle import cc as occ
    g = np.load(os.path.join(GOLD, "cc_flood.npz"))
    for ci in range(3):
        assert np.array_equal(occ.compute_connected_component_segmentation(g[f"c{ci}_hard"]), g[f"c{ci}_seg"])
    g = np.load(os.path.join(GOLD, "cc_affs.npz"))
    for ci in range(3):
        mask = g[f"c{ci}_mask"] if g[f"c{ci}_mask"].size else None
        frags, _ = occ.cc_affs(g[f"c{ci}_affs"], float(g[f"c{ci}_thr"]), 0, mask)
        assert np.array_equal(frags, g[f"c{ci}_seg"])


def test_aff_errors_restatement_matches_reference_goldens():
    """create_diff / create_mask against AddAffErrors._create_diff / _create_mask executed from the reference file;
    seg_to_affgraph (gunpowder, unpinned) on hand-checkable cases"""
    from oracle import aff_errors as oa
    g = np.load(os.path.join(GOLD, "aff_errors.npz"))
    for ci in range(3):
        m = g[f"m{ci}"] if f"m{ci}" in g.files else None
        diff = oa.create_diff(g[f"a{ci}"].copy(), g[f"b{ci}"].copy(), None if m is None else m.copy())
        assert diff.dtype == np.float32 and np.array_equal(diff, g[f"diff{ci}"])
        assert np.array_equal(oa.create_mask(diff, tuple(g[f"thr{ci}"])), g[f"mask{ci}"])
    assert not g["diff2"].any()                                   # the max == 0 branch
    seg = np.array([[[1, 1, 2], [0, 1, 2], [3, 3, 0]]], dtype=np.uint64)
    aff = oa.seg_to_affgraph(seg, [[0, -1, 0], [0, 0, -1], [0, 0, 2]])
    assert aff.shape == (3, 1, 3, 3)
    assert aff[0, 0].tolist() == [[0, 0, 0], [0, 1, 1], [0, 0, 0]]     # p and p - e_y: same id, both > 0; row 0 has no neighbour
    assert aff[1, 0].tolist() == [[0, 1, 0], [0, 0, 0], [0, 1, 0]]
    assert aff[2, 0].tolist() == [[0, 0, 0], [0, 0, 0], [0, 0, 0]]     # (0,0)->(0,2): 1 vs 2; (2,0)->(2,2): 3 vs background


def test_refine_restatement_on_a_hand_checked_volume():
    """oracle/refine.py (bs refine filters, refine.py:98-307) on a volume small enough to check by eye"""
    from oracle import refine as orf
    seg = np.zeros((4, 3, 4), dtype=np.uint64)
    seg[0, 0, :] = 7          # 4 voxels, one plane
    seg[1:4, 1, 0:2] = 9      # 6 voxels, three planes
    seg[2, 2, 3] = 5          # 1 voxel
    seg[0:2, 2, 0] = 11       # 2 voxels, two planes
    uniq, sizes = orf.global_sizes(seg)
    assert uniq.tolist() == [5, 7, 9, 11] and sizes.tolist() == [1, 4, 6, 2]
    out, rem = orf.size_filter(seg, min_size=2, max_size=4)
    assert rem.tolist() == [5, 9] and sorted(np.unique(out).tolist()) == [0, 7, 11]
    out, rem = orf.z_filter(seg, min_z=1)
    assert sorted(rem.tolist()) == [5, 7] and sorted(np.unique(out).tolist()) == [0, 9, 11]
    # sizes 1,4,6,2: mean 3.25, population std 1.92 -> a 1-sigma cut keeps [1.33, 5.17]
    out, rem = orf.outlier_filter(seg, num_std=1.0)
    assert rem.tolist() == [5, 9]
    out = orf.remap(seg, {7: 0, 9: 9, 11: 9})
    assert sorted(np.unique(out).tolist()) == [0, 5, 9] and int((out == 9).sum()) == 8


def test_flood_of_a_mask_component_does_not_depend_on_the_others(tie="index"):
    """the priority flood only moves through mask pixels, so the (level, age) order restricted to one connected
    component of the mask is that component's own order: flooding a component alone (same image, same seeds, the
    other components masked out) gives exactly its part of the whole-slice result.  (Ground for splitting a tile's
    flood by component)  This holds for the index rule of the CUDA path (D1); under skimage's own
    tie behaviour the seeds all carry age 0 and their order follows the heap's layout history, which the other
    components' seeds take part in."""
    from scipy.ndimage import distance_transform_edt, label, maximum_filter
    from bootstrapper_b200.synth import synth_affs
    affs = synth_affs((2, 150, 170), seed=13).astype(np.float64) / 255
    for z in range(2):
        mask = 0.5 * (affs[1, z] + affs[2, z]) > 0.5
        dist = distance_transform_edt(mask)
        seeds, n = label(maximum_filter(dist, 10) == dist)
        whole = sk_watershed(dist.max() - dist, seeds, mask, seed_tie=tie)
        comps, nc = label(mask)
        assert nc > 10
        for c in range(1, nc + 1, max(1, nc // 12)):
            sub = comps == c
            alone = sk_watershed(dist.max() - dist, seeds, sub, seed_tie=tie)
            assert np.array_equal(alone[sub], whole[sub]) and not alone[~sub].any()


def test_filter_avg_fragments_pinned_against_reference():
    """oracle.blockwise.filter_avg_fragments against the reference's own method body (watershed_frags.py:148-156)"""
    g = np.load(os.path.join(GOLD, "filter_fragments.npz"))
    for ci in range(2):
        frags = g[f"frags{ci}"].copy()
        ob.filter_avg_fragments(g[f"affs{ci}"], frags, float(g[f"thr{ci}"]))
        assert np.array_equal(frags, g[f"out{ci}"])


def test_compute_fragments_shift_pinned_against_reference(monkeypatch):
    """the array oracle.blockwise.compute_fragments hands to the watershed (sigma / bias / seed_eps shifts) against the
    one the reference's own method body produces (watershed_frags.py:115-146, tests/golden/make_golden.py)"""
    import json
    g = np.load(os.path.join(GOLD, "compute_fragments_shift.npz"))
    cases = json.load(open(os.path.join(GOLD, "compute_fragments_shift.json")))
    seen = {}

    def recorder(affs, fragments_in_xy=False, min_seed_distance=10, seed_tie="heap"):
        seen["affs"] = affs
        return np.zeros(affs.shape[1:], dtype=np.uint64), 0

    monkeypatch.setattr(ob, "watershed_from_affinities", recorder)
    for ci, case in enumerate(cases):
        p = dict(case, noise_eps=None)
        ob.compute_fragments(g[f"in{ci}"].copy(), p)
        ref = g[f"shifted{ci}"]
        assert seen["affs"].dtype == ref.dtype and np.array_equal(seen["affs"], ref)


def test_ws_glue_pinned_against_reference():
    """oracle.ws.watershed_from_affinities against the reference's post/ws.py run on the same flood (ws.py:8-112):
    fragments, seeds and max id, 2-D per-slice and 3-D mode"""
    from oracle.ws import watershed_from_affinities
    g = np.load(os.path.join(GOLD, "ws_glue.npz"))
    for ci in range(3):
        maxv, xy, msd, max_id = g[f"meta{ci}"]
        frags, mid, seeds = watershed_from_affinities(g[f"affs{ci}"], max_affinity_value=float(maxv), fragments_in_xy=bool(xy),
                                                      return_seeds=True, min_seed_distance=int(msd), seed_tie="heap")
        assert mid == int(max_id)
        assert np.array_equal(frags, g[f"frags{ci}"]) and np.array_equal(seeds, g[f"seeds{ci}"])


def test_agglomerate_glue_pinned_against_reference():
    """oracle.blockwise.agglomerate_in_block against the reference's own method body run around the reference's
    MergeTree and the same restated waterz (waterz_agglom.py:106-170): every initial RAG edge and its merge score"""
    g = np.load(os.path.join(GOLD, "agglomerate_glue.npz"))
    for ci in range(2):
        affs, frags = g[f"affs{ci}"], g[f"frags{ci}"]
        shape = frags.shape
        blk = ob.Block(index=(0, 0, 0), block_id=0, read_offset=(0, 0, 0), read_shape=shape, write_offset=(0, 0, 0), write_shape=shape)
        dbg = ob.agglomerate_in_block(blk, affs, frags, ob.Rag(), (0, 0, 0), stats_mode="faithful", return_debug=True)
        us, vs, _ = dbg["initial"]
        got = {(min(int(u), int(v)), max(int(u), int(v))): s for u, v, s in zip(us, vs, dbg["lca"])}
        ref = {(min(int(u), int(v)), max(int(u), int(v))): s for u, v, s in zip(g[f"u{ci}"], g[f"v{ci}"], g[f"score{ci}"])}
        assert got.keys() == ref.keys() and len(ref) > 20
        for k, s in ref.items():
            assert (np.isnan(s) and np.isnan(got[k])) or float(got[k]) == float(s), (k, got[k], s)


def test_watershed_in_block_glue_pinned_against_reference():
    """oracle.blockwise.watershed_in_block over all blocks of a small volume against the reference's own method bodies
    (watershed_frags.py:178-246) run on the same restated skimage pieces: fragment array, node ids / positions / sizes;
    uint8 and float32 input, 1- and 255-valued masks, 2-D and 3-D mode"""
    g = np.load(os.path.join(GOLD, "watershed_in_block_glue.npz"))
    for ci in range(3):
        meta = g[f"meta{ci}"]
        xy, bs, ctx = bool(meta[0]), tuple(int(v) for v in meta[1:4]), tuple(int(v) for v in meta[4:7])
        affs = g[f"affs{ci}"]
        mask = g[f"mask{ci}"] if f"mask{ci}" in g.files else None
        shape = affs.shape[1:]
        p = dict(ob.WS_DEFAULTS, fragments_in_xy=xy, filter_fragments=0.1, remove_debris=16, min_seed_distance=10)
        frags = np.zeros(shape, dtype=np.uint64)
        rag = ob.Rag()
        for b in ob.enumerate_blocks((0, 0, 0), shape, bs, ctx):
            ob.watershed_in_block(b, affs, frags, rag, p, (0, 0, 0), bs, mask=mask, seed_tie="heap", stats_mode="faithful")
        assert np.array_equal(frags, g[f"frags{ci}"])
        ids = g[f"ids{ci}"]
        assert sorted(rag.node_pos) == [int(i) for i in ids]
        assert np.array_equal(np.array([rag.node_pos[int(i)] for i in ids], dtype=np.int64), g[f"pos{ci}"])
        assert np.array_equal(np.array([rag.node_size[int(i)] for i in ids], dtype=np.int64), g[f"size{ci}"])


def test_simple_watershed_glue_pinned_against_reference():
    """oracle.blockwise.simple_watershed against the reference's own `simple_watershed` (post/watershed.py:206-354)
    executed on in-memory datasets around the same restated skimage / waterz pieces: fragments and one segmentation
    per threshold; uint8 / float32 input, mask, sigma + bias, 3-D mode"""
    import json
    g = np.load(os.path.join(GOLD, "simple_watershed_glue.npz"))
    meta = json.load(open(os.path.join(GOLD, "simple_watershed_glue.json")))
    for ci, case in enumerate(meta["cases"]):
        mask = g[f"mask{ci}"] if case["mask"] else None
        r = ob.simple_watershed(g[f"affs{ci}"], dict(case["cfg"]), mask=mask, seed_tie="heap", stats_mode="faithful")
        names = meta["names"][str(ci)]
        assert names[0].startswith("frags/") and all(n.startswith("segs/") for n in names[1:])
        assert np.array_equal(r["fragments"], g[f"out{ci}_0"])
        thrs = sorted(r["segs"])
        assert len(thrs) == len(names) - 1
        for k, thr in enumerate(thrs):
            assert f"--t{thr:g}--" in names[1 + k]
            assert np.array_equal(r["segs"][thr], g[f"out{ci}_{1 + k}"])


def test_waterz_pipeline_glue_pinned_against_reference():
    """oracle.blockwise.waterz_pipeline against the reference's own `waterz_pipeline` (post/watershed.py:8-203) executed
    with task stand-ins: block size / `// 8` context defaults, fragments, and stage 3 (LUT + segmentation per threshold)"""
    import json
    g = np.load(os.path.join(GOLD, "waterz_pipeline_glue.npz"))
    meta = json.load(open(os.path.join(GOLD, "waterz_pipeline_glue.json")))
    r = ob.waterz_pipeline(g["affs"], dict(meta["cfg"]), block_size=tuple(meta["block_size"]), context=None,
                           seed_tie="heap", stats_mode="faithful")
    assert [tuple(meta["context"])] == [tuple(max(1, s // 8) for s in meta["block_size"])]
    assert np.array_equal(r["fragments"], g["frags"])
    thrs = sorted(r["segs"])
    assert len(thrs) == len(meta["names"])
    for k, thr in enumerate(thrs):
        assert f"--t{thr:g}--" in meta["names"][k]
        assert np.array_equal(r["segs"][thr]["seg"], g[f"seg{k}"])
        ref_lut, lut = g[f"lut{k}"], r["segs"][thr]["lut"]
        assert dict(zip(ref_lut[0].tolist(), ref_lut[1].tolist())) == dict(zip(lut[0].tolist(), lut[1].tolist()))


def test_cc_affs_function_pinned_against_reference():
    """oracle.cc.cc_affs against the reference's own `cc_affs` (post/connected_components.py:12-119) executed on in-memory
    datasets with the reference's cc.py: fragments and the remove_debris segmentation; mask, sigma"""
    import json
    from oracle import cc as occ
    g = np.load(os.path.join(GOLD, "cc_affs_func.npz"))
    meta = json.load(open(os.path.join(GOLD, "cc_affs_func.json")))
    for ci, case in enumerate(meta["cases"]):
        cfg = case["cfg"]
        mask = g[f"mask{ci}"] if case["mask"] else None
        frags, seg = occ.cc_affs(g[f"affs{ci}"], cfg["threshold"], cfg.get("remove_debris", 0), mask, cfg.get("sigma"))
        assert np.array_equal(frags, g[f"frags{ci}"]) and np.array_equal(seg, g[f"seg{ci}"])
        assert frags.any() and (cfg.get("remove_debris", 0) == 0 or not np.array_equal(frags, seg))


def test_refine_filters_pinned_against_reference():
    """oracle/refine.py against the reference's own refine.py functions (`_global_sizes` over z tiles, outlier / size / z
    filters, remap table) executed on an in-memory array (tests/golden/make_golden.py::golden_refine_filters)"""
    from oracle import refine as orf
    g = np.load(os.path.join(GOLD, "refine_filters.npz"))
    seg = g["seg"]
    uniq, sizes = orf.global_sizes(seg)
    assert np.array_equal(uniq, g["uniq"]) and np.array_equal(sizes, g["sizes"])
    assert np.array_equal(np.sort(orf.outlier_filter(seg, 1.0, 20)[1]), g["outlier"])
    assert np.array_equal(np.sort(orf.size_filter(seg, 60, 900)[1]), g["size"])
    assert np.array_equal(np.sort(orf.z_filter(seg, 2)[1]), g["z"])
    # the remap table the host mirror builds (bootstrapper_b200/refine.py::remap follows refine.py:281-300)
    remove = {int(x) for x in g["remap_remove"]}
    merge = {}
    for grp in g["remap_groups"]:
        ids = [int(x) for x in grp if x]
        for mid in ids:
            merge[mid] = ids[0]
    mapping = {i: 0 for i in remove} | merge
    assert sorted(mapping) == [int(k) for k in g["remap_keys"]]
    assert [mapping[int(k)] for k in g["remap_keys"]] == [int(v) for v in g["remap_vals"]]


# ---------------------------------------------------------------- mutex watershed (A12)
def test_mws_oracle_hand_cases():
    """mwatershed.agglom restated: attractive / repulsive / mutex-blocked edges on a 4-voxel row, the declared tie rule D4
    (equal |w|: ascending (channel, voxel)), strides, NaN edges"""
    from oracle import native as on
    # 0-1 attractive 0.9, 1-2 repulsive -0.8, 2-3 attractive 0.7, long range 0-2 attractive 0.6 (blocked by the 1|2 mutex)
    a = np.zeros((2, 1, 1, 4))
    a[0, 0, 0, 1], a[0, 0, 0, 2], a[0, 0, 0, 3] = 0.9, -0.8, 0.7
    a[1, 0, 0, 2], a[1, 0, 0, 3] = 0.6, 0.05
    c = {}
    assert on.mws_agglom(a, [[0, 0, -1], [0, 0, -2]], counters=c).ravel().tolist() == [1, 1, 3, 3]
    assert c["merges"] == 2 and c["mutexes"] == 1 and c["blocked"] == 2
    # the same with the long-range attraction stronger than the repulsion: 0-2 merge first, then 1|2 finds one cluster
    a[1, 0, 0, 2] = 0.95
    assert on.mws_agglom(a, [[0, 0, -1], [0, 0, -2]]).ravel().tolist() == [1, 1, 1, 1]
    # tie rule: 1-2 attractive 0.5 (channel 0) and 1-2... a repulsive -0.5 on channel 1 between 0 and 2 ties with the
    # attractive 0.5 edges of channel 0: channel 0 goes first, so 0-1-2 merge before the mutex arrives
    t = np.zeros((2, 1, 1, 3))
    t[0, 0, 0, 1], t[0, 0, 0, 2] = 0.5, 0.5
    t[1, 0, 0, 2] = -0.5
    assert on.mws_agglom(t, [[0, 0, -1], [0, 0, -2]]).ravel().tolist() == [1, 1, 1]
    assert on.mws_agglom(t[::-1].copy(), [[0, 0, -2], [0, 0, -1]]).ravel().tolist() == [1, 1, 3]   # mutex first now
    # strides: the offset -1 edges exist only at even x
    s = np.full((1, 1, 1, 6), 0.9)
    assert on.mws_agglom(s, [[0, 0, -1]], strides=[[1, 1, 2]]).ravel().tolist() == [1, 2, 2, 4, 4, 6]
    # NaN edges are skipped, zero weight is repulsive
    s[0, 0, 0, 3] = np.nan
    s[0, 0, 0, 4] = 0.0
    assert on.mws_agglom(s, [[0, 0, -1]]).ravel().tolist() == [1, 1, 1, 4, 5, 5]


def test_mws_glue_pinned_against_reference():
    """oracle.mws.mwatershed_from_affinities == the reference's own post/mws.py run with the oracle's agglom
    (tests/golden/mws_glue.npz, make_golden.golden_mws_glue): sigma shift, bias, strides"""
    from oracle import mws as om
    g = np.load(os.path.join(GOLD, "mws_glue.npz"))
    nbh = g["nbh"].tolist()
    for ci in range(3):
        sigma = g[f"sigma{ci}"]
        strides = g[f"strides{ci}"]
        got = om.mwatershed_from_affinities(g[f"affs{ci}"].copy(), nbh, g[f"bias{ci}"].tolist(), sigma=None if sigma[0] < 0 else sigma.tolist(),
                                            strides=None if strides.size == 0 else strides.tolist())
        assert got.dtype == np.uint64 and np.array_equal(got, g[f"frags{ci}"])


def test_mws_seeded_noise_is_standard_normal_like():
    from oracle import mws as om
    n = om.seeded_noise((3, 20, 40, 40), seed=5)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01 and n.min() > -3.5 and n.max() < 3.5
    assert not np.array_equal(n[0], n[1]) and np.array_equal(n, om.seeded_noise((3, 20, 40, 40), seed=5))


def test_histogram_quantile_scores_against_brute_force():
    """the oracle's HistogramQuantileAffinity scores of the initial region graph == a numpy brute force over the contact
    affinities of every fragment pair (bin = min(int(a * 256), 255), pivot = Q * n // 100 + 1, (bin + 0.5) / 256); initmax
    keeps the largest bin only"""
    from oracle.native import Waterz
    rng = np.random.default_rng(4)
    frags = rng.integers(1, 12, (4, 9, 10)).astype(np.uint64)
    frags = np.repeat(np.repeat(frags, 2, 1), 2, 2)
    affs = rng.random((3,) + frags.shape, dtype=np.float32)
    pairs = {}
    for d in range(3):
        a = np.moveaxis(frags, d, 0)
        lo, hi, av = a[:-1], a[1:], np.moveaxis(affs[d], d, 0)[1:]
        m = lo != hi
        for u, v, x in zip(lo[m], hi[m], av[m]):
            pairs.setdefault((min(u, v), max(u, v)), []).append(min(int(np.float32(x) * np.float32(256)), 255))
    for q, initmax in ((50, False), (10, False), (90, True), (25, False)):
        wz = Waterz(affs, frags, 0, "faithful", True, quantile=q, initmax=initmax)
        wz.merge_until(1e-6)                       # scores every edge, merges nothing (scores >= 1/512)
        u, v, sc = wz.region_graph()[:3]
        assert len(u) == len(pairs)
        for a, b, s_ in zip(u, v, sc):
            bins = sorted(pairs[(a, b)])
            if initmax:
                bins = bins[-1:]
            pivot = q * len(bins) // 100 + 1
            want = np.float32(1.0 - (bins[pivot - 1] + 0.5) / 256)
            assert s_ == want, (a, b, s_, want)


def test_mws_oracle_against_an_independent_brute_force():
    """the C++ restatement of mwatershed.agglom against a second, deliberately naive implementation (explicit edge list, python
    sort with the declared key, clusters as sets, mutexes as a set of frozenset pairs rebuilt on every union): random
    volumes, offsets (incl. positive and diagonal ones), strides, ties from quantised weights, NaNs"""
    from oracle import native as on
    rng = np.random.default_rng(12)
    for trial in range(12):
        shape = tuple(int(v) for v in rng.integers(2, 6, 3))
        C_ = int(rng.integers(1, 5))
        offsets = []
        while len(offsets) < C_:
            o = [int(v) for v in rng.integers(-2, 3, 3)]
            if any(o) and o not in offsets:
                offsets.append(o)
        strides = [[int(v) for v in rng.integers(1, 3, 3)] for _ in range(C_)] if trial % 2 else None
        w = rng.normal(0, 1, (C_,) + shape)
        if trial % 3 == 0:
            w = np.round(w * 2) / 2                      # many equal |w|, exact zeros
        if trial % 4 == 0:
            w[rng.random(w.shape) < 0.05] = np.nan
        # --- brute force
        V = int(np.prod(shape))
        idx = np.arange(V).reshape(shape)
        edges = []
        for c, off in enumerate(offsets):
            for p in np.ndindex(*shape):
                q = tuple(a + b for a, b in zip(p, off))
                if any(v < 0 or v >= n for v, n in zip(q, shape)):
                    continue
                if strides is not None and any(a % s for a, s in zip(p, strides[c])):
                    continue
                x = w[(c,) + p]
                if x != x:
                    continue
                edges.append((-abs(x), c, int(idx[p]), int(idx[q]), bool(x > 0)))
        edges.sort(key=lambda e: e[:3])
        cluster = {i: {i} for i in range(V)}
        of = list(range(V))
        mutex = set()
        for _, _, a, b, attractive in edges:
            ca, cb = of[a], of[b]
            if ca == cb:
                continue
            if attractive:
                if frozenset((ca, cb)) in mutex:
                    continue
                keep, gone = min(ca, cb), max(ca, cb)
                for v in cluster[gone]:
                    of[v] = keep
                cluster[keep] |= cluster.pop(gone)
                mutex = {frozenset(keep if x == gone else x for x in m) for m in mutex}
            else:
                mutex.add(frozenset((ca, cb)))
        want = np.array([min(cluster[of[i]]) + 1 for i in range(V)], dtype=np.uint64).reshape(shape)
        got = on.mws_agglom(w, offsets, strides)
        assert np.array_equal(got, want), (trial, shape, offsets, strides)


def test_graph_mws_cluster_against_naive_sets():
    """oracle.mws.mws_cluster (union-find + per-root mutex sets, smaller set moved) against clusters-as-sets with a global set
    of mutex pairs, on random signed graphs"""
    from oracle import mws as om
    rng = np.random.default_rng(21)
    for n, m in ((6, 12), (40, 160), (120, 300)):
        edges = [(bool(rng.random() < 0.55), int(a), int(b)) for a, b in rng.integers(0, n, (m, 2)) if a != b]
        of = list(range(n))
        members = {i: {i} for i in range(n)}
        mutex = set()
        for attractive, a, b in edges:
            ca, cb = of[a], of[b]
            if ca == cb:
                continue
            if attractive:
                if frozenset((ca, cb)) in mutex:
                    continue
                keep, gone = min(ca, cb), max(ca, cb)
                for v in members[gone]:
                    of[v] = keep
                members[keep] |= members.pop(gone)
                mutex = {frozenset(keep if x == gone else x for x in p) for p in mutex}
            else:
                mutex.add(frozenset((ca, cb)))
        want = [min(members[of[i]]) for i in range(n)]
        assert om.mws_cluster(n, edges).tolist() == want


def test_aff_agglom_restatement_against_voxel_loops():
    """oracle.mws.aff_agglom_in_block (slices + np.unique) against plain loops over (offset, voxel): the same pairs, integer
    sums, counts and means, and only edges whose smaller node sits in the block's write ROI are written"""
    from oracle import mws as om
    import oracle.blockwise as ob
    rng = np.random.default_rng(8)
    shape = (6, 9, 8)
    frags = rng.integers(0, 7, shape).astype(np.uint64)
    frags = np.where(rng.random(shape) < 0.15, 0, frags).astype(np.uint64)
    nbh = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, 3, 0], [0, -2, 2]]
    for dtype in (np.uint8, np.float32):
        affs = rng.integers(0, 256, (len(nbh),) + shape).astype(np.uint8)
        if dtype == np.float32:
            affs = (affs.astype(np.float32) / np.float32(255))
        blk = ob.Block(index=(0, 0, 0), block_id=0, write_offset=(1, 2, 1), write_shape=(4, 5, 6), read_offset=(0, 0, 0), read_shape=shape)
        rag = ob.Rag()
        for f in range(1, 7):           # nodes 1..3 positioned inside the write ROI, 4..6 outside
            rag.node_pos[f] = (2, 3, 2) if f <= 3 else (0, 0, 0)
        om.aff_agglom_in_block(blk, affs, frags, rag, (0, 0, 0), nbh, None)
        want = {}
        for c, off in enumerate(nbh):
            for p in np.ndindex(*shape):
                q = tuple(a + b for a, b in zip(p, off))
                if any(v < 0 or v >= n for v, n in zip(q, shape)):
                    continue
                f1, f2 = int(frags[p]), int(frags[q])
                if f1 == 0 or f2 == 0 or f1 == f2:
                    continue
                x = affs[(c,) + p]
                v = int(x) if dtype == np.uint8 else int(np.rint(np.ldexp(np.float64(x), 38)))
                e = want.setdefault((min(f1, f2), max(f1, f2)), [0, 0])
                e[0] += v
                e[1] += 1
        want = {k: v for k, v in want.items() if k[0] <= 3}
        assert set(rag.edges) == set(want) and len(want) > 5
        for k, (sm, cn) in want.items():
            mean = np.float32(np.float64(sm) / 255.0 / cn) if dtype == np.uint8 else np.float32(np.ldexp(np.float64(sm), -38) / cn)
            assert np.float32(rag.edges[k]) == mean
