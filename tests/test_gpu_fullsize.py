"""BASELINE configurations at their REAL sizes: CUDA path (through the C ABI) vs the CPU oracle, bit for bit.

The oracle runs one process per block on all host cores (oracle/parallel.py, ~30 s for config 2 on 16 cores); inputs
come from the device generator (bit-identical to bootstrapper_b200/synth.py, pinned by test_synth_device_matches_numpy).
Oracle switches as everywhere: seed_tie="index" (DESIGN.md D1), stats_mode="canonical" (D2) -- parity unpinned for the
third-party cores the oracle restates.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _both(shape, block, ctx, params, seed=0):
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from oracle.parallel import waterz_pipeline_parallel
    affs = native.synth_affs(shape, seed=seed)
    r = segment_blockwise(affs, params, block, ctx)
    torch.cuda.synchronize()
    ref = waterz_pipeline_parallel(affs.cpu().numpy(), params, block_size=block, context=ctx, seed_tie="index",
                                   stats_mode="canonical")
    return r, ref


def check_vectorised(r, ref):
    """the assertions of test_gpu_parity._check, vectorised for ~10^6 nodes / edges"""
    f = r["fragments"].cpu().numpy().view(np.uint64)
    assert np.array_equal(f, ref["fragments"]), "fragment ids differ"
    nid, npos, nsz = [t.cpu().numpy() for t in r["nodes"]]
    rag = ref["rag"]
    rn = np.array(sorted(rag.node_pos), dtype=np.uint64)
    assert np.array_equal(nid.view(np.uint64), rn), "node ids differ"
    assert np.array_equal(npos, np.array([rag.node_pos[int(i)] for i in rn]).reshape(-1, 3)), "node positions differ"
    assert np.array_equal(nsz, np.array([rag.node_size[int(i)] for i in rn])), "node sizes differ"
    eu, ev, es = [t.cpu().numpy() for t in r["edges"]]
    got = np.stack([eu.view(np.uint64), ev.view(np.uint64)], 1)
    order = np.lexsort((got[:, 1], got[:, 0]))
    got, gs = got[order], es[order]
    keys = sorted(rag.edges)
    want = np.array(keys, dtype=np.uint64).reshape(-1, 2)
    assert np.array_equal(got, want), "RAG edge sets differ"
    ws = np.array([np.nan if rag.edges[k] is None else rag.edges[k] for k in keys], dtype=np.float64)
    nan = np.isnan(ws)
    assert np.array_equal(np.isnan(gs), nan), "NULL merge scores differ"
    assert np.all(np.abs(gs[~nan].astype(np.float64) - ws[~nan]) <= 1e-6 * np.abs(ws[~nan])), "merge scores differ (1e-6 rel)"
    for thr, seg in r["segs"].items():
        assert np.array_equal(seg.cpu().numpy().view(np.uint64), ref["segs"][thr]["seg"]), f"segmentation {thr} differs"
        assert np.array_equal(r["luts"][thr].cpu().numpy().view(np.uint64), ref["segs"][thr]["lut"][1]), f"LUT {thr} differs"
    return dict(fragments=int(len(rn)), edges=int(len(keys)))


def test_config2_full_size_matches_oracle():
    """BASELINE configs[1]: 3x(125,1250,1250) u8, block (25,250,250), context (3,31,31) -- the bench workload itself."""
    r, ref = _both((125, 1250, 1250), (25, 250, 250), (3, 31, 31), {})
    n = check_vectorised(r, ref)
    assert n["fragments"] > 100_000 and n["edges"] > n["fragments"]


def test_config2_full_size_faithful_flood_matches_heap_oracle():
    """the same workload through the FAITHFUL flood (bs_set_flood_version(6): skimage's binary heap replayed literally, seed
    ties as the reference resolves them) against the oracle's faithful mode seed_tie="heap": no declared deviation D1"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.post.pipeline import segment_blockwise
    from oracle.parallel import waterz_pipeline_parallel
    shape, block, ctx = (125, 1250, 1250), (25, 250, 250), (3, 31, 31)
    affs = native.synth_affs(shape, seed=0)
    try:
        native.set_flood_version(6)
        r = segment_blockwise(affs, {}, block, ctx)
        torch.cuda.synchronize()
    finally:
        native.set_flood_version(0)
    ref = waterz_pipeline_parallel(affs.cpu().numpy(), {}, block_size=block, context=ctx, seed_tie="heap", stats_mode="canonical")
    check_vectorised(r, ref)


def test_config4_block_layer_matches_oracle():
    """BASELINE configs[3] geometry: 3-D seeded fragments + seed_eps, 128^3 blocks, context 16 -- one layer of 16 blocks."""
    r, ref = _both((128, 512, 512), (128, 128, 128), (16, 16, 16), {"fragments_in_xy": False, "seed_eps": 0.01})
    check_vectorised(r, ref)


def test_config5_blocks_match_oracle():
    """BASELINE configs[4] geometry: 256^3 blocks, context 32 (block graphs beyond shared memory) -- two neighbouring
    blocks, so that halo fragments and a cross-block edge layer are part of the case."""
    r, ref = _both((256, 256, 512), (256, 256, 256), (32, 32, 32), {})
    check_vectorised(r, ref)


def test_synth_device_matches_numpy():
    """the device generator the full-size cases use is the numpy generator, bit for bit"""
    from bootstrapper_b200 import native
    from bootstrapper_b200.synth import synth_affs
    for dtype, tdt in ((np.uint8, torch.uint8), (np.float32, torch.float32)):
        a = synth_affs((7, 70, 90), seed=5, dtype=dtype, offset=(3, 10, 20), vol_shape=(40, 200, 200))
        b = native.synth_affs((7, 70, 90), seed=5, dtype=tdt, offset=(3, 10, 20))
        assert np.array_equal(a, b.cpu().numpy())
