"""File-level drop-in: `bs segment`-style TOML -> zarr fragments / segmentations, LUTs, SQLite RAG."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_waterz_pipeline_files(tmp_path):
    import toml
    from bootstrapper_b200 import segment, zarrio
    from bootstrapper_b200.graphdb import LUT, open_db
    from bootstrapper_b200.post.naming import build_name
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import waterz_pipeline as ref_pipeline
    affs = synth_affs((12, 120, 120), seed=9)
    vs, off = (40, 4, 4), (80, 16, 16)
    store = str(tmp_path / "v.zarr")
    a = zarrio.prepare_ds(os.path.join(store, "affs"), affs.shape, off, vs, np.uint8, chunk_shape=(3, 6, 60, 60),
                          axis_names=["c^", "z", "y", "x"], units=["nm"] * 3, compressor={"id": "zlib", "level": 1})
    a.write(affs)
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), blockwise=True, num_workers=2,
               block_shape=[6, 60, 60], context="1 8 8", db=dict(db_file=str(tmp_path / "rag.sqlite")))
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    ref = ref_pipeline(affs, {}, block_size=(6, 60, 60), context=(1, 8, 8), seed_tie="index", stats_mode="canonical")

    for mode_kwargs in ({}, {"num_workers": 1}):
        segment.run_segmentation(str(p), "ws", **mode_kwargs)
        fname = build_name(dict(segment.DEFAULTS["ws"], thresholds=None, merge_function=None))
        assert fname == "xy--msd10--ea0--ff0.1--rd64"
        fr = zarrio.open_ds(os.path.join(store, "post/fragments", fname))
        assert fr.offset == off and fr.voxel_size == vs and fr.dtype == np.uint64 and fr.chunks == (6, 60, 60)
        assert np.array_equal(fr.read(), ref["fragments"])
        assert fr.attrs["bs_params"]["method"] == "ws" and fr.attrs["bs_params"]["blockwise"] is True
        nodes, edges, scores = open_db(cfg["db"]).read_graph()
        assert np.array_equal(nodes, np.array(sorted(ref["rag"].node_pos), np.uint64))
        assert {tuple(e) for e in edges.tolist()} == set(ref["rag"].edges)
        for thr in segment.DEFAULTS["ws"]["thresholds"]:
            name = build_name(dict(segment.DEFAULTS["ws"], thresholds=None, threshold=thr))
            seg = zarrio.open_ds(os.path.join(store, "post/segmentations", name))
            assert np.array_equal(seg.read(), ref["segs"][thr]["seg"])
            lut = LUT(os.path.join(store, "post/luts", name)).load()
            assert np.array_equal(lut, ref["segs"][thr]["lut"])
            assert json.load(open(os.path.join(store, "post/luts", name + ".json")))["threshold"] == thr

    # serial path (daisy SerialServer analogue): one process_block call per block gives the same files
    from bootstrapper_b200.post.watershed import waterz_pipeline
    c2 = segment.get_seg_config(str(p), "ws")
    c2["blockwise"] = True
    import bootstrapper_b200.post.watershed as W
    orig = W.run_volara_task
    W.run_volara_task = lambda task, mp: orig(task, False)
    try:
        waterz_pipeline(c2)
    finally:
        W.run_volara_task = orig
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", "xy--msd10--ea0--ff0.1--rd64"))
    assert np.array_equal(fr.read(), ref["fragments"])
    nodes, edges, scores = open_db(cfg["db"]).read_graph()
    assert {tuple(e) for e in edges.tolist()} == set(ref["rag"].edges)


def _write_affs(tmp_path, affs, vs, off):
    from bootstrapper_b200 import zarrio
    store = str(tmp_path / "v.zarr")
    a = zarrio.prepare_ds(os.path.join(store, "affs"), affs.shape, off, vs, affs.dtype, chunk_shape=(3, 6, 60, 60),
                          axis_names=["c^", "z", "y", "x"], units=["nm"] * 3, compressor={"id": "zlib", "level": 1})
    a.write(affs)
    return store


def test_simple_watershed_files(tmp_path):
    """non-blockwise `bs segment --ws`: dataset names, attrs and contents of the single-shot path
    (post/watershed.py:206-354; names per post/naming.py)."""
    import toml
    from bootstrapper_b200 import segment, zarrio
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import simple_watershed as ref_simple
    from test_gpu_parity import _same_partition
    affs = synth_affs((10, 120, 120), seed=12)
    vs, off = (40, 4, 4), (0, 8, 8)
    store = _write_affs(tmp_path, affs, vs, off)
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), ws_params=dict(thresholds=[0.3, 0.6]))
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    segment.run_segmentation(str(p), "ws")
    ref = ref_simple(affs, dict(thresholds=[0.3, 0.6]), seed_tie="index", stats_mode="canonical")
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", "xy--msd10"))
    assert fr.offset == off and fr.voxel_size == vs and fr.dtype == np.uint64
    assert fr.attrs["bs_params"]["method"] == "ws" and fr.attrs["bs_params"]["blockwise"] is False
    assert _same_partition(fr.read(), ref["fragments"])
    for thr in (0.3, 0.6):
        seg = zarrio.open_ds(os.path.join(store, "post/segmentations", f"mfmean--t{thr:g}--xy--msd10"))
        assert _same_partition(seg.read(), ref["segs"][thr])
        assert seg.attrs["bs_params"]["threshold"] == thr


def test_cc_files(tmp_path):
    """`bs segment --cc` (post/connected_components.py:15-127): fragments + debris-filtered segmentation datasets."""
    import toml
    from bootstrapper_b200 import segment, zarrio
    from bootstrapper_b200.synth import synth_affs
    from oracle import cc as occ
    affs = synth_affs((10, 120, 120), seed=13)
    vs, off = (40, 4, 4), (40, 0, 4)
    store = _write_affs(tmp_path, affs, vs, off)
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), cc_params=dict(threshold=0.6, remove_debris=30))
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    segment.run_segmentation(str(p), "cc")
    rf, rs = occ.cc_affs(affs, 0.6, 30)
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", "t0.6"))
    assert fr.offset == off and fr.attrs["bs_params"]["method"] == "cc"
    assert np.array_equal(fr.read(), rf.astype(np.uint64))
    seg = zarrio.open_ds(os.path.join(store, "post/segmentations", "t0.6--rd30"))
    assert np.array_equal(seg.read(), rs.astype(np.uint64))
    with pytest.raises(ValueError):      # segment.py:115-116: "Blockwise connected components is not supported!"
        segment.run_segmentation(str(p), "cc", blockwise=True)
