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
    with pytest.raises(NotImplementedError):      # blockwise mws (volara pipeline) is not built
        segment.run_segmentation(str(p), "mws", blockwise=True, db=dict(db_file=str(tmp_path / "x.sqlite")))
