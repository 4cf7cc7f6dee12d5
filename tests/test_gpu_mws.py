"""`bs segment --mws` (BASELINE config 3): the CUDA mutex watershed (bs_mws_agglom) against the oracle's sequential
restatement of mwatershed.agglom, bit for bit (same labelling: 1 + smallest voxel of the cluster).  Parity unpinned for the
third-party core; declared tie rule D4 on both sides."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NBH3 = [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]
NBH9 = NBH3 + [[-2, 0, 0], [0, -9, 0], [0, 0, -9], [-3, 0, 0], [0, -27, 0], [0, 0, -27]]     # segment.py:24-34 defaults
BIAS9 = [-0.4] * 3 + [-0.7] * 6
STRIDES9 = [[1, 1, 1]] * 3 + [[2, 9, 9]] * 3 + [[3, 27, 27]] * 3


def _affs9(shape, seed, dtype=np.uint8):
    """nine channels from the 3-channel generator: long-range channels = min of the nearest-neighbour affinity along the
    offset's path (what a long-range affinity means), cheap to build on the CPU"""
    from bootstrapper_b200.synth import synth_affs
    a = synth_affs(shape, seed=seed).astype(np.float64) / 255.0
    out = [a[0], a[1], a[2]]
    for off in NBH9[3:]:
        axis = [i for i, o in enumerate(off) if o][0]
        n = -off[axis]
        acc = a[axis].copy()
        for k in range(1, n):
            sl_dst = [slice(None)] * 3
            sl_src = [slice(None)] * 3
            sl_dst[axis], sl_src[axis] = slice(k, None), slice(0, -k)
            sh = np.zeros_like(acc)
            sh[tuple(sl_dst)] = a[axis][tuple(sl_src)]
            acc = np.minimum(acc, sh)
        out.append(acc)
    a9 = np.stack(out)
    return np.rint(a9 * 255).astype(np.uint8) if dtype == np.uint8 else a9.astype(np.float32)


CASES = [
    # shape, neighborhood, bias, strides, noise_eps, dtype, mask
    ((4, 24, 24), NBH3, [-0.5] * 3, None, None, np.uint8, False),                       # ties everywhere (uint8, no noise)
    ((6, 40, 40), NBH3, [-0.4, -0.5, -0.45], None, 0.001, np.uint8, False),
    ((6, 32, 36), NBH3, [-0.5] * 3, None, 0.002, np.float32, True),
    ((8, 64, 64), NBH9, BIAS9, STRIDES9, 0.001, np.uint8, False),                       # the `bs segment --mws` defaults
    ((8, 64, 64), NBH9, BIAS9, None, 0.001, np.uint8, False),                           # dense long-range edges
    ((5, 30, 33), NBH9[:6], [-0.3] * 3 + [-0.8] * 3, [[1, 1, 1]] * 3 + [[1, 2, 3]] * 3, None, np.float32, False),
]


@pytest.mark.parametrize("shape,nbh,bias,strides,noise_eps,dtype,use_mask", CASES)
def test_mws_matches_oracle(shape, nbh, bias, strides, noise_eps, dtype, use_mask):
    from bootstrapper_b200.post.mws import mwatershed_from_affinities
    from oracle import mws as om
    affs = _affs9(shape, seed=31, dtype=dtype)[:len(nbh)]
    mask = None
    if use_mask:
        mask = np.ones(shape, np.uint8)
        mask[:, 10:20, 5:25] = 0
    params = dict(aff_neighborhood=nbh, bias=bias, strides=strides, noise_eps=noise_eps, remove_debris=5)
    ref = om.simple_mutex(affs, params, mask=mask, noise_seed=3)
    frags, seg, cnt = mwatershed_from_affinities(torch.from_numpy(affs).cuda(), nbh, bias, noise_eps=noise_eps, strides=strides,
                                                 mask=None if mask is None else torch.from_numpy(mask).cuda(), noise_seed=3,
                                                 remove_debris=5, return_counters=True)
    assert np.array_equal(frags.cpu().numpy().view(np.uint64), ref["fragments"]), "fragments differ from the oracle"
    assert np.array_equal(seg.cpu().numpy().view(np.uint64), ref["seg"]), "remove_debris result differs"
    assert cnt["merges"] == int(np.prod(shape)) - len(np.unique(ref["fragments"])) and cnt["rounds"] > 0


def test_mws_block_128_matches_oracle():
    """a 128 x 128 x 32 volume with the default 9-offset neighbourhood, strides and noise: ~1.6 M edges"""
    from bootstrapper_b200.post.mws import mwatershed_from_affinities
    from oracle import mws as om
    shape = (32, 128, 128)
    affs = _affs9(shape, seed=40)
    ref = om.simple_mutex(affs, dict(aff_neighborhood=NBH9, bias=BIAS9, strides=STRIDES9, noise_eps=0.001), noise_seed=1)
    frags, _, cnt = mwatershed_from_affinities(torch.from_numpy(affs).cuda(), NBH9, BIAS9, noise_eps=0.001, strides=STRIDES9, noise_seed=1,
                                               return_counters=True)
    assert np.array_equal(frags.cpu().numpy().view(np.uint64), ref["fragments"])
    assert cnt["edges"] > 1_500_000 and cnt["rounds"] > 0


def test_mws_rejects_unreproducible_options():
    from bootstrapper_b200.post.mws import mwatershed_from_affinities
    a = torch.zeros((3, 2, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(NotImplementedError):
        mwatershed_from_affinities(a, NBH3, [-0.5] * 3, randomized_strides=True)
    with pytest.raises(NotImplementedError):
        mwatershed_from_affinities(a, NBH3, [-0.5] * 3, sigma=[1, 1, 1])


def test_simple_mutex_files(tmp_path):
    """non-blockwise `bs segment --mws`: dataset names, attrs and contents (post/watershed_mutex.py:177-291)"""
    import toml
    from bootstrapper_b200 import segment, zarrio
    from oracle import mws as om
    shape = (6, 48, 48)
    affs = _affs9(shape, seed=44)[:6]
    nbh = NBH9[:6]
    vs, off = (40, 4, 4), (80, 0, 8)
    store = str(tmp_path / "v.zarr")
    a = zarrio.prepare_ds(os.path.join(store, "affs"), affs.shape, off, vs, np.uint8, chunk_shape=(6, 3, 24, 24),
                          axis_names=["c^", "z", "y", "x"], units=["nm"] * 3, compressor={"id": "zlib", "level": 1})
    a.write(affs)
    params = dict(aff_neighborhood=nbh, bias=[-0.4] * 3 + [-0.7] * 3, noise_eps=0.001, strides=[[1, 1, 1]] * 3 + [[1, 3, 3]] * 3,
                  randomized_strides=False, remove_debris=8)
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), mws_params=params)
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    segment.run_segmentation(str(p), "mws")
    ref = om.simple_mutex(affs, params, noise_seed=0)
    from bootstrapper_b200.post.naming import build_name
    fname = build_name({k: params[k] for k in ("noise_eps", "bias", "strides", "randomized_strides")})
    assert fname.startswith("eps0.001--b-0.4_-0.4_-0.4_-0.7_-0.7_-0.7--st") and fname.endswith("--rs0")
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", fname))
    assert fr.offset == off and fr.voxel_size == vs and fr.dtype == np.uint64 and fr.attrs["bs_params"]["method"] == "mws"
    assert np.array_equal(fr.read(), ref["fragments"])
    seg = zarrio.open_ds(os.path.join(store, "post/segmentations", fname + "--rd8"))
    assert np.array_equal(seg.read(), ref["seg"]) and seg.attrs["bs_params"]["remove_debris"] == 8


# ---------------------------------------------------------------- blockwise pipeline (volara tasks)
NBH6 = NBH3 + [[-2, 0, 0], [0, -5, 0], [0, 0, -5]]
BW_CASES = [
    # shape, block, context, params, dtype, mask
    ((12, 48, 48), (6, 24, 24), (2, 6, 6), dict(aff_neighborhood=NBH6, bias=[-0.4] * 3 + [-0.7] * 3, noise_eps=0.001), np.uint8, False),
    ((13, 50, 45), (6, 24, 24), (2, 6, 6), dict(aff_neighborhood=NBH6, bias=[-0.4] * 3 + [-0.7] * 3, noise_eps=0.001,
                                                strides=[[1, 1, 1]] * 3 + [[1, 2, 2]] * 3, filter_fragments=0.3, remove_debris=3,
                                                global_bias=[1.0, -0.4]), np.uint8, False),       # ragged blocks
    ((10, 40, 40), (5, 20, 20), (1, 5, 5), dict(aff_neighborhood=NBH3, bias=[-0.5] * 3, noise_eps=0.002, noise_seed=7, filter_fragments=0.3), np.float32, True),
    ((10, 40, 40), (5, 20, 20), (1, 5, 5), dict(aff_neighborhood=NBH3, bias=[-0.5] * 3, noise_eps=0.002, filter_fragments=0.45, remove_debris=2), np.uint8, True),
    ((8, 36, 36), None, None, dict(aff_neighborhood=NBH6, bias=[-0.3] * 3 + [-0.8] * 3, noise_eps=0.001), np.uint8, False),   # one block
]


@pytest.mark.parametrize("shape,block,ctx,params,dtype,use_mask", BW_CASES)
def test_mws_blockwise_pipeline_matches_oracle(shape, block, ctx, params, dtype, use_mask):
    """post/watershed_mutex.py:8-174 (ExtractFrags -> AffAgglom -> GraphMWS -> Relabel) on the device vs the oracle's restatement
    of the four volara tasks: fragments, nodes, edges with their mean affinities, LUT and segmentation, all bit for bit"""
    from bootstrapper_b200.post.pipeline import segment_mws_blockwise
    from oracle import mws as om
    nbh = params["aff_neighborhood"]
    affs = _affs9(shape, seed=11, dtype=dtype)[:len(nbh)] if len(nbh) <= 3 else None
    if affs is None:
        full = _affs9(shape, seed=11, dtype=dtype)
        affs = np.ascontiguousarray(np.stack([full[0], full[1], full[2], full[3], full[4], full[5]]))
    mask = None
    if use_mask:
        mask = np.ones(shape, np.uint8)
        mask[:, 8:16, 4:30] = 0
    if shape == (12, 48, 48):
        affs[:, 0:8, 0:30, 0:30] = 0          # the whole read ROI of block (0, 0, 0) is empty: the reference skips that block
    seed = params.get("noise_seed", 0)
    ref = om.volara_pipeline(affs, params, shape if block is None else block, (0, 0, 0) if block is None else ctx, mask=mask, noise_seed=seed)
    r = segment_mws_blockwise(torch.from_numpy(affs).cuda(), params, block, ctx, mask=None if mask is None else torch.from_numpy(mask).cuda(),
                              agglom_chunk_blocks=3 if shape == (13, 50, 45) else None,          # one case through the chunked AffAgglom
                              mws_chunk_blocks=2 if shape == (13, 50, 45) else None)             # ... and through several mws calls per shape group
    torch.cuda.synchronize()
    f = r["fragments"].cpu().numpy().view(np.uint64)
    assert np.array_equal(f, ref["fragments"]), "fragments differ"
    rag = ref["rag"]
    nid, npos, nsz = [t.cpu().numpy() for t in r["nodes"]]
    rn = np.array(sorted(rag.node_pos), dtype=np.uint64)
    assert np.array_equal(nid.view(np.uint64), rn)
    assert np.array_equal(npos, np.array([rag.node_pos[int(i)] for i in rn]).reshape(-1, 3))
    assert np.array_equal(nsz, np.array([rag.node_size[int(i)] for i in rn]))
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    keys = sorted(rag.edges)
    assert np.array_equal(np.stack([eu.view(np.uint64), ev.view(np.uint64)], 1), np.array(keys, dtype=np.uint64).reshape(-1, 2)), "edge sets differ"
    assert np.array_equal(es, np.array([rag.edges[k] for k in keys], dtype=np.float32)), "zyx_aff differs"
    assert np.array_equal(r["lut"][0].cpu().numpy().view(np.uint64), ref["lut"][0])
    assert np.array_equal(r["lut"][1].cpu().numpy().view(np.uint64), ref["lut"][1]), "LUT differs"
    assert np.array_equal(r["seg"].cpu().numpy().view(np.uint64), ref["seg"]), "segmentation differs"
    assert len(rn) > 20 and len(keys) > len(rn) // 2


def test_graph_mws_matches_sequential_cluster():
    """bs_graph_mws against the oracle's sequential mwatershed.cluster restatement on random signed graphs (ties included)"""
    from bootstrapper_b200 import native
    from oracle import mws as om
    rng = np.random.default_rng(5)
    for n, m, levels in ((50, 200, 7), (2000, 9000, 0), (1, 0, 0), (300, 2500, 3)):
        nodes = np.sort(rng.choice(10 ** 6, n, replace=False)).astype(np.uint64) + 1
        pairs = set()
        while len(pairs) < m:
            a, b = rng.integers(0, n, 2)
            if a != b:
                pairs.add((min(a, b), max(a, b)))
        pairs = sorted(pairs)
        sc = rng.random(len(pairs)).astype(np.float32)
        if levels:
            sc = (np.floor(sc * levels) / levels).astype(np.float32)       # many equal |w|
        w = 1.0 * sc.astype(np.float64) - 0.5
        order = sorted(range(len(pairs)), key=lambda i: -abs(w[i]))
        want = om.mws_cluster(n, [(w[i] > 0, pairs[i][0], pairs[i][1]) for i in order])
        u = torch.from_numpy(nodes[[p[0] for p in pairs]].astype(np.int64) if pairs else np.zeros(0, np.int64)).cuda()
        v = torch.from_numpy(nodes[[p[1] for p in pairs]].astype(np.int64) if pairs else np.zeros(0, np.int64)).cuda()
        got, cnt = native.graph_mws(torch.from_numpy(nodes.astype(np.int64)).cuda(), u, v, torch.from_numpy(sc).cuda(), 1.0, -0.5)
        assert np.array_equal(got.cpu().numpy().view(np.uint64), nodes[want]), (n, m, levels, cnt)


def test_mws_blockwise_files(tmp_path):
    """`bs segment --mws -b` (post/watershed_mutex.py:8-174) end to end on files: a dataset at a world offset of several blocks,
    datasets / RAG (edge attribute zyx_aff) / LUT named and laid out as the reference does, contents = the oracle's"""
    import json
    import toml
    from bootstrapper_b200 import segment, zarrio
    from bootstrapper_b200.graphdb import LUT, open_db
    from bootstrapper_b200.post.naming import build_name
    from oracle import mws as om
    import oracle.blockwise as ob
    shape = (12, 48, 48)
    affs = _affs9(shape, seed=45)[:6]
    nbh = NBH9[:3] + [[-2, 0, 0], [0, -5, 0], [0, 0, -5]]
    vs, off = (40, 4, 4), (40 * 6 * 3, 4 * 24 * 2, 4 * 24 * 5)       # offset = (3, 2, 5) blocks
    store = str(tmp_path / "v.zarr")
    a = zarrio.prepare_ds(os.path.join(store, "affs"), affs.shape, off, vs, np.uint8, chunk_shape=(6, 6, 24, 24),
                          axis_names=["c^", "z", "y", "x"], units=["nm"] * 3, compressor={"id": "zlib", "level": 1})
    a.write(affs)
    params = dict(aff_neighborhood=nbh, bias=[-0.4] * 3 + [-0.7] * 3, noise_eps=0.001, strides=[[1, 1, 1]] * 6, randomized_strides=False,
                  remove_debris=2, filter_fragments=0.2, global_bias=[1.0, -0.45])
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), mws_params=params)
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    db_file = str(tmp_path / "rag.sqlite")
    segment.run_segmentation(str(p), "mws", blockwise=True, block_shape=[6, 24, 24], context=[2, 6, 6], db=dict(db_file=db_file))
    # the oracle numbers blocks from the absolute block index (3, 2, 5)
    orig = ob.enumerate_blocks
    try:
        ob.enumerate_blocks = lambda ro, rs, bs, ctx, index_offset=None: orig(ro, rs, bs, ctx, (18, 48, 120))
        ref = om.volara_pipeline(affs, params, (6, 24, 24), (2, 6, 6), noise_seed=0)
    finally:
        ob.enumerate_blocks = orig
    fparams = {k: params.get(k) for k in ("noise_eps", "bias", "strides", "randomized_strides", "filter_fragments", "remove_debris")}
    fname = build_name({"min_seed_distance": None, "sigma": None, **fparams})
    sname = build_name({"global_bias": [1.0, -0.45], "min_seed_distance": None, "sigma": None, **fparams})
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", fname))
    assert fr.offset == off and fr.voxel_size == vs and fr.attrs["bs_params"]["method"] == "mws" and fr.attrs["bs_params"]["blockwise"] is True
    assert np.array_equal(fr.read(), ref["fragments"])
    seg = zarrio.open_ds(os.path.join(store, "post/segmentations", sname))
    assert np.array_equal(seg.read(), ref["seg"])
    lut_dir = os.path.join(store, "post/luts")
    assert np.array_equal(LUT(os.path.join(lut_dir, sname)).load(), ref["lut"])
    assert json.load(open(os.path.join(lut_dir, sname + ".json")))["global_bias"] == [1.0, -0.45]
    nodes, edges, scores = open_db(dict(db_file=db_file), edge_attrs={"zyx_aff": "float"}).read_graph()
    assert np.array_equal(nodes, np.array(sorted(ref["rag"].node_pos), dtype=np.uint64))
    keys = sorted(ref["rag"].edges)
    order = np.lexsort((edges[:, 1], edges[:, 0]))
    assert np.array_equal(edges[order], np.array(keys, dtype=np.uint64).reshape(-1, 2))
    assert np.array_equal(scores[order], np.array([ref["rag"].edges[k] for k in keys], dtype=np.float32))
