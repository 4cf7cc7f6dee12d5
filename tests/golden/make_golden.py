"""Generate the committed golden fixtures by EXECUTING the reference's own files.

Run in the build container only (needs /root/reference):
    NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden.py
Writes tests/golden/*.npz / *.json.  Nothing at test time reads /root/reference.

Sources executed (unmodified, loaded by path):
  bootstrapper/post/merge_tree.py   -> merge_tree.npz   (MergeTree.merge / find_merges)
  bootstrapper/post/cc.py           -> cc_flood.npz, cc_affs.npz (compute_connected_component_segmentation)
  bootstrapper/gp/add_aff_errors.py -> aff_errors.npz  (_create_diff / _create_mask; gunpowder + skimage imports stubbed)
  bootstrapper/post/naming.py       -> naming.json      (build_name; `import zarr` stubbed)
  bootstrapper/segment.py           -> seg_config.json  (DEFAULTS, get_seg_config)
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/bootstrapper"
OUT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_merge_tree():
    mt_mod = load("ref_merge_tree", f"{REF}/post/merge_tree.py")
    rng = np.random.default_rng(7)
    cases = {}
    for ci, n in enumerate([2, 5, 40, 300]):
        leaves = np.sort(rng.choice(np.arange(1, 10 * n + 50), size=n, replace=False)).astype(np.uint64)
        leaf_arg = np.concatenate([[0], leaves]).astype(np.uint64)      # relabel_map incl. 0
        tree = mt_mod.MergeTree(leaf_arg)
        comp = {int(l): int(l) for l in leaves}          # leaf -> current representative
        alive = [int(l) for l in leaves]
        hist = []
        n_merges = int(n * 0.8)
        for _ in range(n_merges):
            if len(alive) < 2:
                break
            i, j = rng.choice(len(alive), size=2, replace=False)
            a, b = alive[i], alive[j]
            score = float(np.float32(rng.random()))
            tree.merge(a, b, a, score)
            hist.append((a, b, a, score))
            alive.remove(b)
        q = 4 * n + 5
        us = rng.choice(np.concatenate([leaves, [10 ** 6 + 1]]), size=q)
        vs = rng.choice(np.concatenate([leaves, [10 ** 6 + 2]]), size=q)
        out = tree.find_merges([int(u) for u in us], [int(v) for v in vs])
        cases[f"c{ci}_leaves"] = leaf_arg
        cases[f"c{ci}_hist"] = np.array(hist, dtype=np.float64).reshape(-1, 4)
        cases[f"c{ci}_us"] = us.astype(np.uint64)
        cases[f"c{ci}_vs"] = vs.astype(np.uint64)
        cases[f"c{ci}_out"] = out
    np.savez_compressed(f"{OUT}/merge_tree.npz", **cases)


def golden_cc():
    cc_mod = load("ref_cc", f"{REF}/post/cc.py")
    rng = np.random.default_rng(11)
    cases = {}
    for ci, (shape, p) in enumerate([((4, 6, 7), 0.5), ((6, 12, 12), 0.35), ((3, 20, 20), 0.7)]):
        hard = rng.random((3,) + shape) < p
        seg = cc_mod.compute_connected_component_segmentation(hard)
        cases[f"c{ci}_hard"] = hard
        cases[f"c{ci}_seg"] = seg
    np.savez_compressed(f"{OUT}/cc_flood.npz", **cases)


def golden_cc_affs():
    """cc_affs front end (post/connected_components.py:52-56,66,81) restated in numpy + the reference's cc.py."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    cc_mod = load("ref_cc", f"{REF}/post/cc.py")
    rng = np.random.default_rng(5)
    cases = {}
    specs = [("u8", (8, 48, 48), 0.5, False), ("u8", (6, 40, 56), 0.8, True), ("f32", (5, 32, 32), 0.3, True)]
    for ci, (kind, shape, thr, with_mask) in enumerate(specs):
        if kind == "u8":
            affs = synth_affs(shape, seed=20 + ci, dtype=np.uint8)
            data = affs.astype(np.float32) / 255.0
        else:
            affs = rng.random((3,) + shape).astype(np.float32)
            data = affs.astype(np.float32)
        mask = None
        if with_mask:
            mask = (rng.random(shape) < 0.9).astype(np.uint8)
            data *= (mask > 0).astype(np.uint8)
        hard = data > thr
        seg = cc_mod.compute_connected_component_segmentation(hard)
        cases[f"c{ci}_affs"] = affs
        cases[f"c{ci}_thr"] = np.float64(thr)
        cases[f"c{ci}_mask"] = mask if mask is not None else np.zeros(0, np.uint8)
        cases[f"c{ci}_seg"] = seg
    np.savez_compressed(f"{OUT}/cc_affs.npz", **cases)


NAMING_CASES = [
    dict(fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0, sigma=None,
         noise_eps=None, bias=None, filter_fragments=0.1, remove_debris=64),
    dict(merge_function="mean", threshold=0.35, fragments_in_xy=True, min_seed_distance=10, seed_eps=None,
         epsilon_agglomerate=0.0, sigma=None, noise_eps=None, bias=None, filter_fragments=0.1, remove_debris=64),
    dict(fragments_in_xy=True, min_seed_distance=10, sigma=None, noise_eps=None, bias=None),
    dict(fragments_in_xy=False, min_seed_distance=7, seed_eps=0.01, epsilon_agglomerate=0.05,
         sigma=[1, 2, 2], noise_eps=0.001, bias=[-0.1, -0.2, -0.2], filter_fragments=0.0, remove_debris=0),
    dict(merge_function="hist_quant_75", threshold=0.5, fragments_in_xy=True, min_seed_distance=10,
         bias=-0.5, sigma=[2, 2, 2]),
    dict(global_bias=[1.0, -0.5], noise_eps=0.001, bias=[-0.4, -0.4, -0.4, -0.7, -0.7, -0.7, -0.7, -0.7, -0.7],
         strides=[[1, 1, 1]] * 3 + [[2, 9, 9]] * 3 + [[3, 27, 27]] * 3, randomized_strides=True,
         filter_fragments=0.1, remove_debris=64, sigma=None),
    dict(threshold=1e-05, remove_debris=100000, randomized_strides=False),
]


def golden_aff_errors():
    """AddAffErrors._create_diff / _create_mask (gp/add_aff_errors.py:163-183) executed from the reference file; the
    gunpowder / skimage imports of the module are stubbed (the two functions are plain numpy)."""
    gp = types.ModuleType("gunpowder")
    for n in ("BatchFilter", "Array", "BatchRequest", "Batch", "Coordinate"):
        setattr(gp, n, type(n, (), {}))
    nodes = types.ModuleType("gunpowder.nodes")
    addaff = types.ModuleType("gunpowder.nodes.add_affinities")
    addaff.seg_to_affgraph = None
    morph = types.ModuleType("skimage.morphology")
    morph.ball = morph.disk = None
    sk = types.ModuleType("skimage")
    for name, mod in (("gunpowder", gp), ("gunpowder.nodes", nodes), ("gunpowder.nodes.add_affinities", addaff),
                      ("skimage", sk), ("skimage.morphology", morph)):
        sys.modules.setdefault(name, mod)
    ref = load("ref_add_aff_errors", f"{REF}/gp/add_aff_errors.py").AddAffErrors
    rng = np.random.default_rng(23)
    out = {}
    for ci, (C, shape, use_mask, thr) in enumerate([(3, (5, 17, 19), False, (0.1, 1.0)), (9, (4, 12, 11), True, (0.05, 0.7)),
                                                    (3, (3, 8, 8), True, (0.1, 1.0))]):
        a = (rng.random((C,) + shape) < 0.7).astype(np.float32)
        b = rng.random((C,) + shape).astype(np.float32)
        if ci == 2:
            b = a.copy()                                               # zero error everywhere: the max == 0 branch
        m = (rng.random(shape) < 0.8).astype(np.uint8) if use_mask else None
        diff = ref._create_diff(None, a.copy(), b.copy(), None if m is None else m.copy())
        msk = ref._create_mask(None, diff, thr)
        out[f"a{ci}"], out[f"b{ci}"], out[f"diff{ci}"], out[f"mask{ci}"] = a, b, diff, msk
        out[f"thr{ci}"] = np.array(thr)
        if m is not None:
            out[f"m{ci}"] = m
    np.savez_compressed(os.path.join(OUT, "aff_errors.npz"), **out)
    print("aff_errors.npz", len(out))


def golden_naming():
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))       # naming.py only uses zarr in dump_params
    nm = load("ref_naming", f"{REF}/post/naming.py")
    out = [dict(params=c, name=nm.build_name(c)) for c in NAMING_CASES]
    with open(f"{OUT}/naming.json", "w") as f:
        json.dump(out, f, indent=1)


CONFIG_TOMLS = {
    "plain": """
affs_dataset = "/d/v.zarr/affs"
fragments_dataset = "/d/v.zarr/post/fragments"
seg_dataset_prefix = "/d/v.zarr/post/segmentations"
""",
    "blockwise": """
affs_dataset = "/d/v.zarr/affs"
fragments_dataset = "/d/v.zarr/post/fragments"
seg_dataset_prefix = "/d/v.zarr/post/segmentations"
blockwise = true
num_workers = 4
block_shape = "25 250 250"
context = [3, 31, 31]
roi_offset = "0,0,0"
roi_shape = "5000 5000 5000"
[db]
db_file = "/d/rag.sqlite"
[ws_params]
thresholds = [0.1, 0.9]
remove_debris = 10
""",
}
CONFIG_CALLS = [
    ("plain", "ws", {}),
    ("plain", "ws", {"param": ("thresholds=[0.3]", "fragments_in_xy=False", "merge_function=hist_quant_50")}),
    ("plain", "mws", {}),
    ("plain", "cc", {"param": ("threshold=0.7",)}),
    ("blockwise", "ws", {}),
    ("blockwise", "ws", {"block_context": "2 8 8", "num_workers": 9, "param": ("seed_eps=0.01",)}),
    ("blockwise", "ws", {"block_shape": "roi"}),
]
CONFIG_ERRORS = [
    ("plain", "ws", {"param": ("bogus=1",)}),
    ("blockwise", "cc", {}),
    ("plain", "ws", {"blockwise": True}),
]


def golden_config():
    for name in ("waterz", "funlib", "funlib.segment"):
        sys.modules.setdefault(name, types.ModuleType(name))
    seg = load("ref_segment", f"{REF}/segment.py")
    out = dict(defaults=seg.DEFAULTS, calls=[], errors=[])
    with tempfile.TemporaryDirectory() as td:
        paths = {}
        for k, text in CONFIG_TOMLS.items():
            paths[k] = os.path.join(td, f"{k}.toml")
            with open(paths[k], "w") as f:
                f.write(text)
        for cfg, method, kwargs in CONFIG_CALLS:
            res = seg.get_seg_config(paths[cfg], method, **kwargs)
            out["calls"].append(dict(toml=cfg, method=method, kwargs=kwargs, result=res))
        for cfg, method, kwargs in CONFIG_ERRORS:
            try:
                seg.get_seg_config(paths[cfg], method, **kwargs)
                err = None
            except Exception as e:  # noqa: BLE001
                err = [type(e).__name__, str(e)]
            out["errors"].append(dict(toml=cfg, method=method, kwargs=kwargs, error=err))
    out["tomls"] = CONFIG_TOMLS
    with open(f"{OUT}/seg_config.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    golden_merge_tree()
    golden_cc()
    golden_cc_affs()
    golden_aff_errors()
    golden_naming()
    golden_config()
    print("golden fixtures written to", OUT)
