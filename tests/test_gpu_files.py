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


def test_waterz_pipeline_files_offset_roi_mask(tmp_path):
    """a dataset whose world offset is several blocks from the origin, a sub-ROI, and a mask dataset that is smaller than
    and shifted against the affinities: block ids come from the ABSOLUTE block index, so every fragment id in
    the zarr, the RAG and the LUTs carries it; the mask is cropped / zero-padded in world units
    (watershed_frags.py:207-213 mask.to_ndarray(block.read_roi, fill_value=0))."""
    import toml
    from bootstrapper_b200 import segment, zarrio
    from bootstrapper_b200.graphdb import open_db
    from bootstrapper_b200.post.naming import build_name
    from bootstrapper_b200.synth import synth_affs
    from oracle.blockwise import cantor_number, waterz_pipeline as ref_pipeline
    affs = synth_affs((18, 180, 180), seed=21)
    vs = (40, 4, 4)
    off = (40 * 6 * 3 + 80, 4 * 60 * 5, 4 * 60 * 2 + 16)          # 3 / 5 / 2 blocks (+ a fraction) from the origin
    store = str(tmp_path / "v.zarr")
    a = zarrio.prepare_ds(os.path.join(store, "affs"), affs.shape, off, vs, np.uint8, chunk_shape=(3, 6, 60, 60),
                          axis_names=["c^", "z", "y", "x"], units=["nm"] * 3, compressor={"id": "zlib", "level": 1})
    a.write(affs)
    # mask: starts 4 / 20 / 10 voxels inside the affinity array and ends before its end; values 0 / 255
    mshape, mstart = (10, 140, 150), (4, 20, 10)
    rng = np.random.default_rng(3)
    mask = (rng.random(mshape) > 0.02).astype(np.uint8) * 255
    mask[:, 60:70, :] = 0
    moff = tuple(o + s * v for o, s, v in zip(off, mstart, vs))
    m = zarrio.prepare_ds(os.path.join(store, "mask"), mshape, moff, vs, np.uint8, chunk_shape=(5, 70, 75))
    m.write(mask)
    roi_vox_off, roi_vox_shape = (6, 0, 60), (12, 180, 120)     # sub-ROI in voxels of the array
    roi_offset = [o + r * v for o, r, v in zip(off, roi_vox_off, vs)]
    roi_shape = [s * v for s, v in zip(roi_vox_shape, vs)]
    cfg = dict(affs_dataset=os.path.join(store, "affs"), fragments_dataset=os.path.join(store, "post/fragments"),
               seg_dataset_prefix=os.path.join(store, "post/segmentations"), mask_dataset=os.path.join(store, "mask"),
               roi_offset=roi_offset, roi_shape=roi_shape, blockwise=True, num_workers=1, block_shape=[6, 60, 60],
               context=[1, 8, 8], db=dict(db_file=str(tmp_path / "rag.sqlite")))
    p = tmp_path / "seg.toml"
    p.write_text(toml.dumps(cfg))
    full_mask = np.zeros(affs.shape[1:], np.uint8)
    full_mask[tuple(slice(s, s + n) for s, n in zip(mstart, mshape))] = mask > 0
    absolute = tuple(o // v + r for o, v, r in zip(off, vs, roi_vox_off))
    ref = ref_pipeline(affs, {}, block_size=(6, 60, 60), context=(1, 8, 8), roi=(roi_vox_off, roi_vox_shape), mask=full_mask,
                       seed_tie="index", stats_mode="canonical", index_offset=absolute)
    # the smallest block id is the cantor number of the absolute index of the ROI's first block, far from 0
    first = cantor_number(tuple(a0 // b for a0, b in zip(absolute, (6, 60, 60))))
    assert first > 100 and min(b.block_id for b in ref["blocks"]) == first
    segment.run_segmentation(str(p), "ws")
    fname = build_name(dict(segment.DEFAULTS["ws"], thresholds=None, merge_function=None))
    fr = zarrio.open_ds(os.path.join(store, "post/fragments", fname))
    assert tuple(fr.offset) == tuple(roi_offset) and fr.shape == roi_vox_shape
    got = fr.read()
    assert got.max() >= first * 6 * 60 * 60
    assert np.array_equal(got, ref["fragments"])
    nodes, edges, scores = open_db(cfg["db"]).read_graph()
    assert np.array_equal(nodes, np.array(sorted(ref["rag"].node_pos), np.uint64))
    assert {tuple(e) for e in edges.tolist()} == set(ref["rag"].edges)
    for thr in segment.DEFAULTS["ws"]["thresholds"]:
        name = build_name(dict(segment.DEFAULTS["ws"], thresholds=None, threshold=thr))
        seg = zarrio.open_ds(os.path.join(store, "post/segmentations", name))
        assert np.array_equal(seg.read(), ref["segs"][thr]["seg"])
    # a mask with a different voxel size is refused (no silent misalignment): the task fails, run_volara_task reports it as
    # the reference's check_task_states does (blockwise.py:15-22)
    zarrio.prepare_ds(os.path.join(store, "mask2"), mshape, moff, (40, 8, 8), np.uint8, chunk_shape=(5, 70, 75)).write(mask)
    cfg["mask_dataset"] = os.path.join(store, "mask2")
    p.write_text(toml.dumps(cfg))
    with pytest.raises(RuntimeError):
        segment.run_segmentation(str(p), "ws")
