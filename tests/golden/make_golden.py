"""Generate the committed golden fixtures by EXECUTING the reference's own files.

Run in the build container only (needs /root/reference):
    NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden.py
Writes tests/golden/*.npz / *.json.  Nothing at test time reads /root/reference.

Sources executed (unmodified, loaded by path):
  bootstrapper/post/merge_tree.py   -> merge_tree.npz   (MergeTree.merge / find_merges)
  bootstrapper/post/cc.py           -> cc_flood.npz, cc_affs.npz (compute_connected_component_segmentation)
  bootstrapper/gp/add_aff_errors.py -> aff_errors.npz  (_create_diff / _create_mask; gunpowder + skimage imports stubbed)
  bootstrapper/post/blockwise/watershed_frags.py -> filter_fragments.npz (filter_avg_fragments, method body via ast)
  bootstrapper/post/blockwise/watershed_frags.py -> compute_fragments_shift.npz (compute_fragments with the watershed call recorded)
  bootstrapper/post/ws.py           -> ws_glue.npz      (whole file; skimage's watershed replaced by the oracle's restatement)
  bootstrapper/post/blockwise/waterz_agglom.py -> agglomerate_glue.npz (agglomerate_in_block around the reference's MergeTree;
                                       waterz / funlib relabel replaced by the oracle's restatements)
  bootstrapper/post/blockwise/watershed_frags.py -> watershed_in_block_glue.npz (get_fragments / watershed_in_block over all blocks)
  bootstrapper/post/watershed.py    -> simple_watershed_glue.npz (simple_watershed with in-memory datasets)
  bootstrapper/post/watershed.py    -> waterz_pipeline_glue.npz (waterz_pipeline; task stand-ins run the oracle's per-block stages)
  bootstrapper/post/connected_components.py -> cc_affs_func.npz (cc_affs with in-memory datasets)
  bootstrapper/refine.py            -> refine_filters.npz (_global_sizes and the outlier / size / z filters, remap)
  bootstrapper/blockwise.py         -> task_states.json (check_task_states messages)
  bootstrapper/post/blockwise/*.py  -> task_fields.json (field names / defaults / methods of the two task classes)
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
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod                 # numba's on-disk cache re-imports the defining module by name
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


def golden_filter_fragments():
    """WatershedFrags.filter_avg_fragments (post/blockwise/watershed_frags.py:148-156), the method body executed as it
    stands in the reference file (extracted by ast: the module itself needs volara / funlib / skimage to import);
    volara.tmp.replace_values is stubbed by its documented in-place value map."""
    import ast
    from scipy.ndimage import mean as ndi_mean
    src = open(f"{REF}/post/blockwise/watershed_frags.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "filter_avg_fragments")
    mod = ast.Module(body=[fn], type_ignores=[])

    def replace_values(arr, old, new):
        lut = dict(zip(old.tolist(), new.tolist()))
        flat = arr.reshape(-1)
        for i, v in enumerate(flat.tolist()):
            if v in lut:
                flat[i] = lut[v]

    ns = {"np": np, "ndi_mean": ndi_mean, "replace_values": replace_values}
    exec(compile(mod, "watershed_frags.filter_avg_fragments", "exec"), ns)
    rng = np.random.default_rng(31)
    out = {}
    for ci, (shape, dtype, thr) in enumerate([((4, 24, 20), np.float64, 0.5), ((3, 16, 16), np.float32, 0.35)]):
        frags = np.zeros(shape, dtype=np.uint64)
        ids = rng.permutation(np.arange(1, 40))
        for k, (z, y, x) in enumerate(np.ndindex(shape[0], shape[1] // 4, shape[2] // 4)):
            frags[z, 4 * y:4 * y + 4, 4 * x:4 * x + 4] = ids[k % len(ids)] if rng.random() < 0.9 else 0
        level = rng.random(int(frags.max()) + 1)
        affs = np.clip(level[frags.astype(np.int64)][None] + 0.15 * rng.standard_normal((3,) + shape), 0, 1).astype(dtype)
        got = frags.copy()
        ns["filter_avg_fragments"](None, affs, got, thr)
        out[f"affs{ci}"], out[f"frags{ci}"], out[f"out{ci}"], out[f"thr{ci}"] = affs, frags, got, np.array(thr)
        assert 0 < np.unique(got).size < np.unique(frags).size
    np.savez_compressed(os.path.join(OUT, "filter_fragments.npz"), **out)
    print("filter_fragments.npz", len(out))


SHIFT_CASES = [
    dict(sigma=None, bias=[-0.05, -0.1, -0.1], seed_eps=None, min_seed_distance=10, fragments_in_xy=True),
    dict(sigma=[1, 2, 2], bias=None, seed_eps=None, min_seed_distance=10, fragments_in_xy=True),
    dict(sigma=None, bias=0.07, seed_eps=0.01, min_seed_distance=6, fragments_in_xy=False),
    dict(sigma=[0, 1.5, 0.8], bias=[-0.02, -0.03, -0.03], seed_eps=0.02, min_seed_distance=10, fragments_in_xy=False),
]


def golden_compute_fragments_shift():
    """WatershedFrags.compute_fragments (post/blockwise/watershed_frags.py:115-146), the method body executed as it
    stands in the reference file (extracted by ast), with `watershed_from_affinities` replaced by a recorder: pins the
    array the reference hands to the watershed (sigma / bias / seed_eps shifts; scipy is the real scipy)."""
    import ast
    from scipy.ndimage import distance_transform_edt, gaussian_filter, label, maximum_filter
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    src = open(f"{REF}/post/blockwise/watershed_frags.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "compute_fragments")
    seen = {}

    def recorder(affs, fragments_in_xy=False, min_seed_distance=10):
        seen["affs"], seen["xy"], seen["msd"] = affs, fragments_in_xy, min_seed_distance
        return np.zeros(affs.shape[1:], dtype=np.uint64), 0

    ns = {"np": np, "gaussian_filter": gaussian_filter, "distance_transform_edt": distance_transform_edt,
          "maximum_filter": maximum_filter, "label": label, "watershed_from_affinities": recorder}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "watershed_frags.compute_fragments", "exec"), ns)
    out = {}
    for ci, case in enumerate(SHIFT_CASES):
        dtype = np.float64 if ci % 2 == 0 else np.float32          # uint8 input arrives as float64 / 255, float32 as is
        affs = (synth_affs((8, 40, 36), seed=40 + ci).astype(np.float64) / 255).astype(dtype)
        me = types.SimpleNamespace(noise_eps=None, **case)
        ns["compute_fragments"](me, affs.copy())
        assert seen["xy"] == case["fragments_in_xy"] and seen["msd"] == case["min_seed_distance"]
        out[f"in{ci}"], out[f"shifted{ci}"] = affs, seen["affs"]
    np.savez_compressed(os.path.join(OUT, "compute_fragments_shift.npz"), **out)
    with open(os.path.join(OUT, "compute_fragments_shift.json"), "w") as f:
        json.dump(SHIFT_CASES, f, indent=1)
    print("compute_fragments_shift.npz", len(out))


def golden_ws_glue():
    """post/ws.py executed unmodified, with skimage.segmentation.watershed (absent here) replaced by the oracle's
    restatement of it: pins everything AROUND the flood -- mean affinities, the > 0.5 * max threshold, scipy's EDT /
    maximum filter / label, the per-slice id offsets and the seed output (ws.py:8-112) -- given that flood."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle.native import sk_watershed
    seg = types.ModuleType("skimage.segmentation")
    seg.watershed = lambda image, markers, mask=None: sk_watershed(image, markers, mask, seed_tie="heap")
    sys.modules["skimage"] = sys.modules.get("skimage") or types.ModuleType("skimage")
    sys.modules["skimage.segmentation"] = seg
    ws = load("ref_ws", f"{REF}/post/ws.py")
    out = {}
    for ci, (shape, dtype, maxv, xy, msd) in enumerate([((5, 60, 50), np.float64, 255.0, True, 10), ((6, 40, 44), np.float32, 1.0, False, 6),
                                                         ((3, 64, 64), np.float32, 1.0, True, 10)]):
        a8 = synth_affs(shape, seed=50 + ci)
        affs = a8.astype(dtype) if maxv == 255.0 else (a8.astype(np.float32) / np.float32(255)).astype(dtype)
        frags, max_id, seeds = ws.watershed_from_affinities(affs, max_affinity_value=maxv, fragments_in_xy=xy, return_seeds=True,
                                                            min_seed_distance=msd)
        out[f"affs{ci}"], out[f"frags{ci}"], out[f"seeds{ci}"] = affs, frags, seeds
        out[f"meta{ci}"] = np.array([maxv, int(xy), msd, max_id], dtype=np.float64)
        assert frags.dtype == np.uint64 and frags.any()
    np.savez_compressed(os.path.join(OUT, "ws_glue.npz"), **out)
    print("ws_glue.npz", len(out))


def golden_mws_glue():
    """post/mws.py executed unmodified, with the `mwatershed` package (absent here) replaced by the oracle's restatement of
    mwatershed.agglom: pins the glue of mwatershed_from_affinities (mws.py:12-59) -- the gaussian shift, the bias broadcast,
    the float64 cast, the offsets / strides hand-over -- given that mutex watershed.  noise_eps is None (unseeded upstream)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle.native import mws_agglom
    calls = []

    def agglom(affs, offsets, strides=None, randomized_strides=False):
        assert affs.dtype == np.float64 and not randomized_strides
        calls.append((offsets, strides))
        return mws_agglom(affs, offsets, strides)

    mod = types.ModuleType("mwatershed")
    mod.agglom = agglom
    sys.modules["mwatershed"] = mod
    if "scipy.ndimage.filters" not in sys.modules:
        try:
            import scipy.ndimage.filters  # noqa: F401
        except Exception:  # noqa: BLE001  (removed namespace: the reference imports gaussian_filter from it)
            import scipy.ndimage as ndi
            f = types.ModuleType("scipy.ndimage.filters")
            f.gaussian_filter = ndi.gaussian_filter
            sys.modules["scipy.ndimage.filters"] = f
    ref = load("ref_mws", f"{REF}/post/mws.py")
    nbh = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -5, 0], [0, 0, -5]]
    out = {}
    cases = [((6, 40, 36), [-0.4] * 3 + [-0.7] * 3, None, None),
             ((5, 32, 32), [-0.5, -0.4, -0.45, -0.7, -0.6, -0.7], [1, 2, 2], [[1, 1, 1]] * 3 + [[2, 3, 3]] * 3),
             ((4, 30, 28), [-0.3] * 3 + [-0.8] * 3, [0, 1.5, 1.5], None)]
    for ci, (shape, bias, sigma, strides) in enumerate(cases):
        a8 = synth_affs(shape, seed=70 + ci)
        # six channels: the three nearest-neighbour ones twice (the long-range copies get their own weights below)
        affs = np.concatenate([a8, a8[:, ::-1, :, :]], 0).astype(np.float64) / 255.0
        frags = ref.mwatershed_from_affinities(affs.copy(), nbh, bias, sigma=sigma, noise_eps=None, strides=strides, randomized_strides=False)
        out[f"affs{ci}"], out[f"frags{ci}"] = affs, frags
        out[f"bias{ci}"] = np.array(bias, dtype=np.float64)
        out[f"sigma{ci}"] = np.array(sigma if sigma is not None else [-1, -1, -1], dtype=np.float64)
        out[f"strides{ci}"] = np.array(strides if strides is not None else [], dtype=np.int64)
        assert frags.dtype == np.uint64 and len(np.unique(frags)) > 3
    out["nbh"] = np.array(nbh, dtype=np.int64)
    assert len(calls) == len(cases)
    np.savez_compressed(os.path.join(OUT, "mws_glue.npz"), **out)
    print("mws_glue.npz", len(out))


def golden_agglomerate_glue():
    """WaterzAgglom.agglomerate_in_block (post/blockwise/waterz_agglom.py:106-170), the method body executed as it
    stands in the reference file (extracted by ast) around the reference's own MergeTree; `waterz.agglomerate` and
    funlib's `relabel` (absent) are the oracle's restatements, arrays / graph store are small in-memory stand-ins.
    Pins the glue: float32 normalisation, dense relabel + backwards map, initial RAG at threshold 0, history -> merge
    tree, merge_score per initial edge (NaN -> None)."""
    import ast
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle import blockwise as ob
    from oracle.native import Waterz
    mt_mod = load("ref_merge_tree2", f"{REF}/post/merge_tree.py")

    def agglomerate(affs, thresholds, fragments, scoring_function, discretize_queue, return_merge_history, return_region_graph):
        assert scoring_function == "OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>" and discretize_queue == 256
        wz = Waterz(affs, fragments, 256, "faithful", True)
        for thr in thresholds:
            a, b, c, sc = wz.merge_until(float(thr))
            u, v, s, _, _ = wz.region_graph()
            yield (wz.segmentation(),
                   [dict(a=int(x), b=int(y), c=int(z), score=float(w)) for x, y, z, w in zip(a, b, c, sc)],
                   [dict(u=int(x), v=int(y), score=float(w)) for x, y, w in zip(u, v, s)])

    waterz = types.ModuleType("waterz")
    waterz.agglomerate = agglomerate
    arrays = types.ModuleType("funlib.segment.arrays")
    arrays.relabel = lambda a, return_backwards_map=True: ob.funlib_relabel(a)
    for name, mod in (("waterz", waterz), ("funlib", types.ModuleType("funlib")), ("funlib.segment", types.ModuleType("funlib.segment")),
                      ("funlib.segment.arrays", arrays)):
        sys.modules[name] = mod
    src = open(f"{REF}/post/blockwise/waterz_agglom.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "agglomerate_in_block")
    ns = {"np": np, "MergeTree": mt_mod.MergeTree}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "waterz_agglom.agglomerate_in_block", "exec"), ns)

    class Arr:
        def __init__(self, a):
            self.a = a

        def to_ndarray(self, roi, fill_value=0):
            return ob.to_ndarray(self.a, roi[0], roi[1], fill_value)

    class Graph:
        def __init__(self):
            self.e = {}

        def add_edge(self, u, v, **data):
            self.e[(u, v)] = dict(data)

        def edges(self, data=True):
            return [(u, v, d) for (u, v), d in self.e.items()]

    class Provider:
        def __getitem__(self, roi):
            return Graph()

        def write_graph(self, rag, roi, write_nodes=False):
            self.written = [(u, v, d["merge_score"]) for u, v, d in rag.edges(data=True)]

    out = {}
    for ci, (shape, dtype) in enumerate([((6, 48, 40), np.uint8), ((5, 40, 40), np.float32)]):
        a8 = synth_affs(shape, seed=60 + ci)
        affs = a8 if dtype == np.uint8 else (a8.astype(np.float32) / np.float32(255))
        frags, _ = __import__("oracle.ws", fromlist=["x"]).watershed_from_affinities(
            a8.astype(np.float64) / 255, fragments_in_xy=True, seed_tie="index")
        frags[frags > 0] += np.uint64(1000 * (ci + 1))                  # ids far from dense: the backwards map matters
        roi = ((0, 0, 0), shape)
        prov = Provider()
        me = types.SimpleNamespace(merge_function="OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>")
        ns["agglomerate_in_block"](me, types.SimpleNamespace(read_roi=roi, write_roi=roi), Arr(affs), Arr(frags), prov)
        e = prov.written
        out[f"affs{ci}"], out[f"frags{ci}"] = affs, frags
        out[f"u{ci}"] = np.array([x[0] for x in e], dtype=np.uint64)
        out[f"v{ci}"] = np.array([x[1] for x in e], dtype=np.uint64)
        out[f"score{ci}"] = np.array([np.nan if x[2] is None else x[2] for x in e], dtype=np.float64)
        assert len(e) > 20 and np.isfinite(out[f"score{ci}"]).any()
    np.savez_compressed(os.path.join(OUT, "agglomerate_glue.npz"), **out)
    print("agglomerate_glue.npz", len(out))


def golden_watershed_in_block_glue():
    """WatershedFrags.get_fragments / watershed_in_block (post/blockwise/watershed_frags.py:178-246), the method bodies
    executed as they stand in the reference file (extracted by ast) together with compute_fragments /
    filter_avg_fragments and the reference's post/ws.py; skimage's watershed / label / remove_small_objects are the
    oracle's restatements, funlib's Array / Roi / Coordinate and the graph store small in-memory stand-ins
    (Coordinate truncates floats like funlib.geometry does [3P-recall]).  Pins the glue of stage 1: normalisation,
    the empty-block early-out, mask handling, crop to the write ROI, relabel + block id offset, node positions / sizes."""
    import ast
    from scipy.ndimage import center_of_mass, distance_transform_edt, gaussian_filter, label, maximum_filter
    from scipy.ndimage import mean as ndi_mean
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle import blockwise as ob
    from oracle.native import sk_label, sk_watershed
    seg = types.ModuleType("skimage.segmentation")
    seg.watershed = lambda image, markers, mask=None: sk_watershed(image, markers, mask, seed_tie="heap")
    sys.modules["skimage"] = sys.modules.get("skimage") or types.ModuleType("skimage")
    sys.modules["skimage.segmentation"] = seg
    ws = load("ref_ws2", f"{REF}/post/ws.py")

    class Coord(tuple):
        def __new__(cls, v):
            return super().__new__(cls, (int(x) for x in v))

        def __mul__(self, o):
            return Coord(a * b for a, b in zip(self, o))

        __rmul__ = __mul__

        def __add__(self, o):
            return Coord(a + b for a, b in zip(self, o))

        __radd__ = __add__

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape, self.dims = Coord(offset), Coord(shape), len(shape)

    class Vol:
        def __init__(self, a, voxel_size=(1, 1, 1)):
            self.a, self.dtype, self.voxel_size = a, a.dtype, Coord(voxel_size)

        def to_ndarray(self, roi, fill_value=0):
            return ob.to_ndarray(self.a, roi.offset, roi.shape, fill_value)

        def __setitem__(self, roi, data):
            self.a[tuple(slice(o, o + s) for o, s in zip(roi.offset, roi.shape))] = data

    class Array:                                   # funlib.persistence.arrays.Array(data, offset=, voxel_size=)
        def __init__(self, data, offset, voxel_size):
            self.data, self.offset = data, offset

        def to_ndarray(self, roi):
            return self.data[tuple(slice(o - b, o - b + s) for o, b, s in zip(roi.offset, self.offset, roi.shape))]

    class Graph:
        def __init__(self):
            self.nodes = {}

        def add_node(self, i, **data):
            self.nodes[i] = data

    class Provider:
        def __init__(self):
            self.nodes = {}

        def __getitem__(self, roi):
            return Graph()

        def write_graph(self, rag, roi):
            self.nodes.update(rag.nodes)

    src = open(f"{REF}/post/blockwise/watershed_frags.py").read()
    tree = ast.parse(src)
    wanted = ("compute_fragments", "filter_avg_fragments", "get_fragments", "watershed_in_block")
    fns = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in wanted]
    cls = ast.ClassDef(name="Frags", bases=[], keywords=[], body=fns, decorator_list=[])
    mod = ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[]))
    ns = {"np": np, "gaussian_filter": gaussian_filter, "distance_transform_edt": distance_transform_edt,
          "maximum_filter": maximum_filter, "label": label, "ndi_mean": ndi_mean, "center_of_mass": center_of_mass,
          "watershed_from_affinities": ws.watershed_from_affinities, "Array": Array, "Coordinate": Coord,
          "relabel": lambda x, return_num=True: sk_label(x),
          "remove_small_objects": lambda x, min_size: ob.remove_small_objects(x, min_size),
          "replace_values": lambda arr, old, new: arr.__setitem__(np.isin(arr, old), 0),
          "logger": types.SimpleNamespace(info=lambda *a, **k: None)}
    exec(compile(mod, "watershed_frags.Frags", "exec"), ns)
    out = {}
    shape, bs, ctx = (8, 64, 56), (4, 32, 28), (1, 6, 6)
    for ci, (dtype, use_mask, xy) in enumerate([(np.uint8, False, True), (np.float32, True, True), (np.uint8, True, False)]):
        a8 = synth_affs(shape, seed=70 + ci)
        affs = a8 if dtype == np.uint8 else (a8.astype(np.float32) / np.float32(255))
        mask = None
        if use_mask:
            mask = np.ones(shape, dtype=np.uint8) * (255 if ci == 1 else 1)
            mask[:, :10, :12] = 0
        frags = np.zeros(shape, dtype=np.uint64)
        prov = Provider()
        me = ns["Frags"]()
        for k, v in dict(noise_eps=None, sigma=None, bias=None, seed_eps=None, min_seed_distance=10, fragments_in_xy=xy,
                         epsilon_agglomerate=0, filter_fragments=0.1, remove_debris=16, num_voxels_in_block=int(np.prod(bs)),
                         voxel_size=Coord((1, 1, 1))).items():
            setattr(me, k, v)
        blocks = ob.enumerate_blocks((0, 0, 0), shape, bs, ctx)
        for b in blocks:
            blk = types.SimpleNamespace(read_roi=Roi(b.read_offset, b.read_shape), write_roi=Roi(b.write_offset, b.write_shape),
                                        block_id=("task", b.block_id))
            me.watershed_in_block(blk, Vol(affs), Vol(frags), prov, mask=None if mask is None else Vol(mask))
        ids = np.array(sorted(prov.nodes), dtype=np.uint64)
        out[f"affs{ci}"], out[f"frags{ci}"], out[f"ids{ci}"] = affs, frags, ids
        out[f"pos{ci}"] = np.array([prov.nodes[int(i)]["position"] for i in ids], dtype=np.int64)
        out[f"size{ci}"] = np.array([prov.nodes[int(i)]["size"] for i in ids], dtype=np.int64)
        if mask is not None:
            out[f"mask{ci}"] = mask
        out[f"meta{ci}"] = np.array([int(xy)] + list(bs) + list(ctx), dtype=np.int64)
        assert ids.size > 10 and np.array_equal(np.unique(frags[frags > 0]), ids)
    np.savez_compressed(os.path.join(OUT, "watershed_in_block_glue.npz"), **out)
    print("watershed_in_block_glue.npz", len(out))


def golden_simple_watershed_glue():
    """post/watershed.py `simple_watershed` (:206-354), the function executed as it stands in the reference file
    (extracted by ast) with the reference's own naming.py and ws.py; funlib's open_ds / prepare_ds / Roi are in-memory
    stand-ins, skimage's watershed and `waterz.agglomerate` (default queue) the oracle's restatements.  Pins the glue
    of the single-shot path: float32 normalisation, mask, sigma / bias shift, fragments, one segmentation per threshold,
    dataset names."""
    import ast
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle.native import Waterz, sk_watershed
    seg = types.ModuleType("skimage.segmentation")
    seg.watershed = lambda image, markers, mask=None: sk_watershed(image, markers, mask, seed_tie="heap")
    sys.modules["skimage"] = sys.modules.get("skimage") or types.ModuleType("skimage")
    sys.modules["skimage.segmentation"] = seg
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))
    store, written = {}, {}

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape = tuple(offset), tuple(shape)

        def sl(self):
            return tuple(slice(o, o + s) for o, s in zip(self.offset, self.shape))

    class DS:
        def __init__(self, a, name=None):
            self.a, self.name, self.shape, self.dtype = a, name, a.shape, a.dtype
            self.roi = Roi((0, 0, 0), a.shape[-3:])
            self.voxel_size, self.axis_names, self.units = (1, 1, 1), ["c^", "z", "y", "x"][-a.ndim:], ["nm"] * 3

        def __getitem__(self, roi):
            return self.a[(Ellipsis,) + roi.sl()]

        def __setitem__(self, roi, data):
            self.a[roi.sl()] = data
            written[self.name] = self.a

    persistence = types.ModuleType("funlib.persistence")
    persistence.open_ds = lambda path: DS(store[path])
    persistence.prepare_ds = lambda name, shape, offset, voxel_size, axis_names, dtype, units: DS(np.zeros(shape, dtype=dtype), name)
    geometry = types.ModuleType("funlib.geometry")
    geometry.Roi = Roi

    def agglomerate(affs, thresholds, fragments, scoring_function):
        import re
        m = re.fullmatch(r"OneMinus<HistogramQuantileAffinity<RegionGraphType, (\d+), ScoreValue, 256, (true|false)>>", scoring_function)
        assert m or scoring_function == "OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>"
        wz = Waterz(affs, fragments, 0, "faithful", True, quantile=int(m.group(1)) if m else 0, initmax=bool(m) and m.group(2) == "true")
        for thr in sorted(thresholds):
            wz.merge_until(float(thr))
            yield wz.segmentation()

    waterz = types.ModuleType("waterz")
    waterz.agglomerate = agglomerate
    pkg = types.ModuleType("refpost")
    pkg.__path__ = []
    for name, mod in (("funlib", types.ModuleType("funlib")), ("funlib.persistence", persistence), ("funlib.geometry", geometry),
                      ("waterz", waterz), ("refpost", pkg)):
        sys.modules[name] = mod
    naming = load("refpost.naming", f"{REF}/post/naming.py")
    naming.dump_params = lambda *a, **k: None
    sys.modules["refpost.naming"] = naming
    sys.modules["refpost.ws"] = load("refpost.ws", f"{REF}/post/ws.py")
    src = open(f"{REF}/post/watershed.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "simple_watershed")
    ns = {"__package__": "refpost", "__name__": "refpost.watershed"}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "watershed.simple_watershed", "exec"), ns)
    out, names = {}, {}
    cases = [dict(dtype="uint8", mask=False, cfg={}),
             dict(dtype="float32", mask=True, cfg={"sigma": [0, 1.5, 1.0], "bias": [-0.05, -0.1, -0.1], "thresholds": [0.1, 0.3, 0.6]}),
             dict(dtype="uint8", mask=False, cfg={"fragments_in_xy": False, "min_seed_distance": 6, "bias": [-0.03, -0.03, -0.03]}),
             dict(dtype="uint8", mask=False, cfg={"merge_function": "hist_quant_75_initmax", "thresholds": [0.3, 0.6]}),
             dict(dtype="float32", mask=True, cfg={"merge_function": "hist_quant_25", "bias": [-0.1, -0.2, -0.2], "thresholds": [0.5]})]
    for ci, case in enumerate(cases):
        shape = (5, 56, 48)
        a8 = synth_affs(shape, seed=80 + ci)
        store["affs"] = a8 if case["dtype"] == "uint8" else (a8.astype(np.float32) / np.float32(255))
        cfg = dict(affs_dataset="affs", fragments_dataset="frags", seg_dataset_prefix="segs", **case["cfg"])
        if case["mask"]:
            m = np.ones(shape, dtype=np.uint8)
            m[:, 40:, :20] = 0
            store["mask"] = m
            cfg["mask_dataset"] = "mask"
            out[f"mask{ci}"] = m
        written.clear()
        ns["simple_watershed"](cfg)
        names[str(ci)] = sorted(written)
        out[f"affs{ci}"] = store["affs"]
        for k, name in enumerate(sorted(written)):
            out[f"out{ci}_{k}"] = written[name].copy()
        assert len(written) == 1 + len(cfg.get("thresholds", [0.2, 0.35, 0.5]))
    np.savez_compressed(os.path.join(OUT, "simple_watershed_glue.npz"), **out)
    with open(os.path.join(OUT, "simple_watershed_glue.json"), "w") as f:
        json.dump(dict(cases=cases, names=names), f, indent=1)
    print("simple_watershed_glue.npz", len(out), names["1"])


def golden_waterz_pipeline_glue():
    """post/watershed.py `waterz_pipeline` (:8-203), the function executed as it stands in the reference file (extracted
    by ast) with the reference's own naming.py.  The volara tasks are stand-ins that record their arguments and run the
    oracle's (separately pinned) per-block stages; the graph store, LUT and datasets are in memory; funlib's
    `connected_components` is the oracle's restatement.  Pins the orchestration: parameter defaults, block size and the
    `// 8` context rule, task arguments, and stage 3 -- unscored edges skipped, float32 scores, one component array and
    LUT per threshold, relabel -- plus the dataset / LUT names."""
    import ast
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle import blockwise as ob
    from oracle.native import connected_components
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))
    store, luts, segs, tasks = {}, {}, {}, []

    class Coordinate(tuple):
        def __new__(cls, v):
            return super().__new__(cls, (int(x) for x in v))

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape, self.dims = Coordinate(offset), Coordinate(shape), len(shape)

    class DS:
        def __init__(self, a):
            self.a, self.shape, self.roi, self.chunk_shape = a, a.shape, Roi((0, 0, 0), a.shape[1:]), (a.shape[0], 4, 32, 28)

    class Holder:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    rag = ob.Rag()

    class Graph:
        nodes = property(lambda self: list(rag.node_pos))

        def edges(self, data=True):
            return [(u, v, {"merge_score": sc}) for (u, v), sc in rag.edges.items()]

    class DB(Holder):
        def open(self, mode):
            return self

        def read_graph(self, roi, edge_attrs=None):
            return Graph()

    class LUT(Holder):
        def save(self, arr):
            luts[os.path.basename(self.path)] = np.array(arr)

    class Task(Holder):
        pass

    def make(name):
        return type(name, (Task,), {})

    WatershedFrags, WaterzAgglom, Relabel = make("WatershedFrags"), make("WaterzAgglom"), make("Relabel")

    def run_volara_task(task, blockwise):
        tasks.append((type(task).__name__, dict(task.__dict__), blockwise))
        affs = store[task.affs_data.store] if hasattr(task, "affs_data") else None
        if isinstance(task, Relabel):
            lut = luts[os.path.basename(task.lut.path)]
            frags = store["frags:" + task.frags_data.store]
            seg = frags.copy()
            idx = np.searchsorted(np.sort(lut[0]), frags)
            order = np.argsort(lut[0])
            idx[idx >= lut.shape[1]] = 0
            hit = lut[0][order][idx] == frags
            seg[hit] = lut[1][order][idx[hit]]
            segs[os.path.basename(task.seg_data.store)] = seg
            return
        off, shape = task.roi
        blocks = ob.enumerate_blocks(tuple(off), tuple(shape), tuple(task.block_size), tuple(task.context))
        if isinstance(task, WatershedFrags):
            p = dict(ob.WS_DEFAULTS, **{k: getattr(task, k) for k in ("fragments_in_xy", "min_seed_distance", "seed_eps", "epsilon_agglomerate",
                                                                       "sigma", "noise_eps", "bias", "filter_fragments", "remove_debris")})
            frags = store.setdefault("frags:" + task.frags_data.store, np.zeros(tuple(shape), dtype=np.uint64))
            for b in blocks:
                ob.watershed_in_block(b, affs, frags, rag, p, tuple(off), tuple(task.block_size), None, "heap", "faithful")
        else:
            frags = store["frags:" + task.frags_data.store]
            for b in blocks:
                ob.agglomerate_in_block(b, affs, frags, rag, tuple(off), "faithful", True)

    src_agg = open(f"{REF}/post/blockwise/waterz_agglom.py").read()
    mf = next(n for n in ast.walk(ast.parse(src_agg)) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "WATERZ_MERGE_FUNCTIONS")
    mods = {
        "funlib": types.ModuleType("funlib"), "funlib.geometry": Holder(Coordinate=Coordinate, Roi=Roi),
        "funlib.persistence": Holder(open_ds=lambda path: DS(store[path])),
        "funlib.segment": types.ModuleType("funlib.segment"), "funlib.segment.graphs": types.ModuleType("funlib.segment.graphs"),
        "funlib.segment.graphs.impl": Holder(connected_components=connected_components),
        "volara": types.ModuleType("volara"), "volara.blockwise": Holder(Relabel=Relabel),
        "volara.datasets": Holder(Labels=Holder, Raw=Holder), "volara.dbs": Holder(SQLite=DB, PostgreSQL=DB),
        "volara.lut": Holder(LUT=LUT), "volara.logging": Holder(set_log_basedir=lambda p: None),
        "refpkg": types.ModuleType("refpkg"), "refpkg.post": types.ModuleType("refpkg.post"),
        "refpkg.post.blockwise": types.ModuleType("refpkg.post.blockwise"),
        "refpkg.post.blockwise.watershed_frags": Holder(WatershedFrags=WatershedFrags),
        "refpkg.post.blockwise.waterz_agglom": Holder(WaterzAgglom=WaterzAgglom, WATERZ_MERGE_FUNCTIONS=ast.literal_eval(mf.value)),
        "refpkg.blockwise": Holder(run_volara_task=run_volara_task),
    }
    for n in ("refpkg", "refpkg.post", "refpkg.post.blockwise"):
        mods[n].__path__ = []
    sys.modules.update(mods)
    naming = load("refpkg.post.naming", f"{REF}/post/naming.py")
    naming.dump_params = naming.dump_lut_params = lambda *a, **k: None
    sys.modules["refpkg.post.naming"] = naming
    src = open(f"{REF}/post/watershed.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "waterz_pipeline")
    ns = {"__package__": "refpkg.post", "__name__": "refpkg.post.watershed", "logger": types.SimpleNamespace(warning=lambda *a: None)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "watershed.waterz_pipeline", "exec"), ns)
    shape = (8, 64, 56)
    store["affs.zarr/affs"] = synth_affs(shape, seed=90)
    tmp = tempfile.mkdtemp()
    cfg = dict(affs_dataset="affs.zarr/affs", fragments_dataset="out.zarr/frags", seg_dataset_prefix="out.zarr/segs",
               lut_dir=os.path.join(tmp, "luts"), db={"db_file": os.path.join(tmp, "rag.db")}, blockwise=True, num_workers=3,
               filter_fragments=0.1, remove_debris=16, thresholds=[0.2, 0.35, 0.5])          # block_shape / context left to the defaults
    ns["waterz_pipeline"](cfg)
    kinds = [t[0] for t in tasks]
    assert kinds == ["WatershedFrags", "WaterzAgglom", "Relabel", "Relabel", "Relabel"], kinds
    t0 = tasks[0][1]
    assert tuple(t0["block_size"]) == (4, 32, 28) and tuple(t0["context"]) == (1, 4, 3) and t0["num_workers"] == 3
    out = {"affs": store["affs.zarr/affs"], "frags": next(v for k, v in store.items() if k.startswith("frags:"))}
    meta = dict(block_size=list(t0["block_size"]), context=list(t0["context"]), frags_name=next(k for k in store if k.startswith("frags:"))[6:],
                names=sorted(luts), cfg={k: v for k, v in cfg.items() if k in ("filter_fragments", "remove_debris", "thresholds")})
    assert sorted(luts) == sorted(segs)
    for k, name in enumerate(sorted(luts)):
        out[f"lut{k}"], out[f"seg{k}"] = luts[name], segs[name]
    np.savez_compressed(os.path.join(OUT, "waterz_pipeline_glue.npz"), **out)
    with open(os.path.join(OUT, "waterz_pipeline_glue.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("waterz_pipeline_glue.npz", len(out), meta["frags_name"], meta["names"])


def golden_volara_pipeline_glue():
    """post/watershed_mutex.py `volara_pipeline` (:8-174), the function executed as it stands in the reference file (extracted
    by ast) with the reference's own naming.py; the volara tasks / datasets / LUT / DB are stand-ins that record their
    arguments.  Pins the orchestration of the blockwise mws pipeline: parameter defaults, block size and the `// 8` context
    rule, the arguments every task receives (ExtractFrags, AffAgglom scores, GraphMWS weights, Relabel) and the dataset / LUT
    names."""
    import ast
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))
    tasks = []

    class Coordinate(tuple):
        def __new__(cls, v):
            return super().__new__(cls, (int(x) for x in v))

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape, self.dims = Coordinate(offset), Coordinate(shape), len(shape)

    class DS:
        shape, chunk_shape = (6, 20, 96, 80), (6, 5, 48, 40)
        roi = Roi((0, 0, 0), (20, 96, 80))

    class Holder:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    def make(name):
        return type(name, (Holder,), {})

    ExtractFrags, AffAgglom, GraphMWS, Relabel = make("ExtractFrags"), make("AffAgglom"), make("GraphMWS"), make("Relabel")

    def plain(v):
        if isinstance(v, Holder):
            return {k: plain(x) for k, x in v.__dict__.items()}
        if isinstance(v, Roi):
            return [list(v.offset), list(v.shape)]
        if isinstance(v, dict):
            return {k: plain(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [plain(x) for x in v]
        return v

    def run_volara_task(task, blockwise=None, multiprocessing=None):
        tasks.append(dict(task=type(task).__name__, args=plain(task), blockwise=blockwise, multiprocessing=multiprocessing))

    mods = {
        "funlib": types.ModuleType("funlib"), "funlib.geometry": Holder(Coordinate=Coordinate, Roi=Roi),
        "funlib.persistence": Holder(open_ds=lambda path: DS()),
        "volara": types.ModuleType("volara"),
        "volara.blockwise": Holder(ExtractFrags=ExtractFrags, AffAgglom=AffAgglom, GraphMWS=GraphMWS, Relabel=Relabel),
        "volara.datasets": Holder(Affs=make("Affs"), Labels=make("Labels"), Raw=make("Raw")),
        "volara.dbs": Holder(SQLite=make("SQLite"), PostgreSQL=make("PostgreSQL")),
        "volara.lut": Holder(LUT=make("LUT")), "volara.logging": Holder(set_log_basedir=lambda p: None),
        "refmws": types.ModuleType("refmws"), "refmws.post": types.ModuleType("refmws.post"),
        "refmws.blockwise": Holder(run_volara_task=run_volara_task),
    }
    for n in ("refmws", "refmws.post"):
        mods[n].__path__ = []
    sys.modules.update(mods)
    naming = load("refmws.post.naming", f"{REF}/post/naming.py")
    naming.dump_params = naming.dump_lut_params = lambda *a, **k: None
    sys.modules["refmws.post.naming"] = naming
    src = open(f"{REF}/post/watershed_mutex.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "volara_pipeline")
    ns = {"__package__": "refmws.post", "__name__": "refmws.post.watershed_mutex"}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "watershed_mutex.volara_pipeline", "exec"), ns)
    nbh = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -5, 0], [0, 0, -5]]
    tmp = tempfile.mkdtemp()
    base = dict(affs_dataset="v.zarr/affs", fragments_dataset="v.zarr/post/fragments", seg_dataset_prefix="v.zarr/post/segmentations",
                lut_dir=os.path.join(tmp, "luts"), db={"db_file": "rag.db"}, aff_neighborhood=nbh, bias=[-0.4] * 3 + [-0.7] * 3)
    cases = [dict(blockwise=True, num_workers=4, noise_eps=0.001, remove_debris=8),
             dict(blockwise=True, block_shape=[10, 48, 40], context=[2, 6, 6], strides=[[1, 1, 1]] * 3 + [[1, 2, 2]] * 3, filter_fragments=0.2,
                  global_bias=[0.9, -0.4], roi_offset=[0, 0, 0], roi_shape=[10, 96, 80]),
             dict(blockwise=False, noise_eps=0.002)]
    out = []
    for extra in cases:
        tasks.clear()
        cfg = dict(base, **extra)
        ns["volara_pipeline"](cfg)
        for t in tasks:
            for k in ("lut",):
                if isinstance(t["args"].get(k), dict) and "path" in t["args"][k]:
                    t["args"][k]["path"] = os.path.relpath(t["args"][k]["path"], tmp)
        out.append(dict(cfg={k: v for k, v in extra.items()}, tasks=json.loads(json.dumps(tasks))))
    with open(os.path.join(OUT, "volara_pipeline_glue.json"), "w") as f:
        json.dump(dict(base={k: v for k, v in base.items() if k != "lut_dir"}, affs=dict(shape=list(DS.shape), chunk_shape=list(DS.chunk_shape)),
                       cases=out), f, indent=1)
    print("volara_pipeline_glue.json", [[t["task"] for t in c["tasks"]] for c in out])


def golden_cc_affs_func():
    """post/connected_components.py `cc_affs` (:12-119), the function executed as it stands in the reference file
    (extracted by ast) with the reference's own cc.py and naming.py; funlib's open_ds / prepare_ds / Roi are in-memory
    stand-ins, skimage's remove_small_objects the oracle's restatement.  Covers mask, sigma and remove_debris."""
    import ast
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle import cc as occ
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))
    store, written = {}, {}

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape = tuple(offset), tuple(shape)

        def sl(self):
            return tuple(slice(o, o + s) for o, s in zip(self.offset, self.shape))

    class DS:
        def __init__(self, a, name=None):
            self.a, self.name, self.shape = a, name, a.shape
            self.roi = Roi((0, 0, 0), a.shape[-3:])
            self.voxel_size, self.axis_names, self.units = (1, 1, 1), ["c^", "z", "y", "x"][-a.ndim:], ["nm"] * 3

        def __getitem__(self, roi):
            return self.a[(Ellipsis,) + roi.sl()]

        def __setitem__(self, roi, data):
            self.a[roi.sl()] = data
            written[self.name] = self.a

    holder = lambda **kw: types.SimpleNamespace(**kw)  # noqa: E731
    pkg = types.ModuleType("refcc")
    pkg.__path__ = []
    sys.modules.update({
        "funlib": types.ModuleType("funlib"), "funlib.geometry": holder(Roi=Roi),
        "funlib.persistence": holder(open_ds=lambda path: DS(store[path]),
                                     prepare_ds=lambda name, shape, offset, voxel_size, axis_names, dtype, units: DS(np.zeros(shape, dtype=dtype), name)),
        "skimage": sys.modules.get("skimage") or types.ModuleType("skimage"),
        "skimage.morphology": holder(remove_small_objects=lambda x, min_size: occ.remove_small_objects(x, min_size)),
        "refcc": pkg})
    naming = load("refcc.naming", f"{REF}/post/naming.py")
    naming.dump_params = lambda *a, **k: None
    sys.modules["refcc.naming"] = naming
    sys.modules["refcc.cc"] = load("ref_cc", f"{REF}/post/cc.py")          # one module name per file (numba cache)
    src = open(f"{REF}/post/connected_components.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "cc_affs")
    ns = {"__package__": "refcc", "__name__": "refcc.connected_components"}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "connected_components.cc_affs", "exec"), ns)
    out, names = {}, {}
    cases = [dict(dtype="uint8", mask=False, cfg={"threshold": 0.5}),
             dict(dtype="uint8", mask=True, cfg={"threshold": 0.7, "sigma": [0, 1.0, 1.5], "remove_debris": 40}),
             dict(dtype="float32", mask=True, cfg={"threshold": 0.4, "remove_debris": 12})]
    for ci, case in enumerate(cases):
        shape = (5, 44, 40)
        a8 = synth_affs(shape, seed=95 + ci)
        store["affs"] = a8 if case["dtype"] == "uint8" else (a8.astype(np.float32) / np.float32(255))
        cfg = dict(affs_dataset="affs", fragments_dataset="frags", seg_dataset_prefix="segs", **case["cfg"])
        if case["mask"]:
            m = np.ones(shape, dtype=np.uint8)
            m[:, :9, 30:] = 0
            store["mask"] = m
            cfg["mask_dataset"] = "mask"
            out[f"mask{ci}"] = m
        written.clear()
        ns["cc_affs"](cfg)
        names[str(ci)] = sorted(written)
        assert len(written) == 2 and names[str(ci)][0].startswith("frags/")
        out[f"affs{ci}"], out[f"frags{ci}"], out[f"seg{ci}"] = store["affs"], written[names[str(ci)][0]].copy(), written[names[str(ci)][1]].copy()
    np.savez_compressed(os.path.join(OUT, "cc_affs_func.npz"), **out)
    with open(os.path.join(OUT, "cc_affs_func.json"), "w") as f:
        json.dump(dict(cases=cases, names=names), f, indent=1)
    print("cc_affs_func.npz", len(out), names["1"])


def golden_refine_filters():
    """refine.py `_scan_tiles`, `_global_sizes`, `outlier_filter`, `size_filter`, `z_filter`, `remap` (:79-108, :147-307),
    the functions executed as they stand in the reference file (extracted by ast, click decorators dropped); the zarr
    array is an in-memory stand-in, fastremap.unique is numpy's, the blockwise rewrite is replaced by a recorder of the
    ids to remove / the id mapping.  Pins the decision arithmetic of the filters (sizes over z tiles, mean / std cut,
    ranges, z extents, remap table)."""
    import ast
    import contextlib
    import io
    import click
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from bootstrapper_b200.synth import synth_affs
    from oracle.ws import watershed_from_affinities

    class Coordinate(tuple):
        def __new__(cls, *v):
            v = v[0] if len(v) == 1 else v
            return super().__new__(cls, (int(x) for x in v))

        def __mul__(self, o):
            return Coordinate(a * b for a, b in zip(self, o))

        def __add__(self, o):
            return Coordinate(a + b for a, b in zip(self, o))

        __radd__ = __add__

    class Roi:
        def __init__(self, offset, shape):
            self.offset, self.shape = Coordinate(offset), Coordinate(shape)

    class DS:
        def __init__(self, a):
            self.a, self.shape, self.dtype = a, a.shape, a.dtype
            self.roi, self.voxel_size, self.chunk_shape = Roi((0, 0, 0), a.shape), Coordinate(1, 1, 1), (2, 16, 16)

        def to_ndarray(self, roi):
            return self.a[tuple(slice(o, o + s) for o, s in zip(roi.offset, roi.shape))]

    seen = {}

    def finish(in_ds, in_array, out_array, remove_ids, num_workers, dry_run, suffix, name):
        seen[name] = np.array(remove_ids)

    def run_blockwise(name, in_ds, out_ds, process_block, num_workers, **kw):
        seen[name] = process_block.args[2]

    fastremap = types.SimpleNamespace(unique=lambda a, return_counts=False: np.unique(a, return_counts=return_counts))
    src = open(f"{REF}/refine.py").read()
    wanted = ("_scan_tiles", "_global_sizes", "outlier_filter", "size_filter", "z_filter", "remap", "_remap_block")
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    for f in fns:
        f.decorator_list = []
    vol = {}
    ns = {"np": np, "click": click, "Roi": Roi, "Coordinate": Coordinate, "fastremap": fastremap, "partial": __import__("functools").partial,
          "open_ds": lambda path: DS(vol["seg"]), "_finish_filter": finish, "_run_blockwise": run_blockwise,
          "_default_out": lambda a, sfx: a + "_" + sfx, "_prepare_like": lambda in_ds, out: None}
    exec(compile(ast.fix_missing_locations(ast.Module(body=fns, type_ignores=[])), "refine", "exec"), ns)
    a8 = synth_affs((9, 72, 64), seed=99)
    frags, _ = watershed_from_affinities(a8.astype(np.float64) / 255, fragments_in_xy=False, seed_tie="index", min_seed_distance=6)
    frags[frags > 0] += np.uint64(5000)
    frags[4:, :, :20][frags[4:, :, :20] % 3 == 0] = 0          # uneven sizes and z extents
    vol["seg"] = frags
    out = {"seg": frags}
    with contextlib.redirect_stdout(io.StringIO()):
        ns["outlier_filter"]("seg", None, 1.0, 20, 1, False)
        ns["size_filter"]("seg", None, 60, 900, 1, False)
        ns["z_filter"]("seg", None, 2, 1, False)
        ids = np.unique(frags[frags > 0])
        rm, grp = [int(ids[1])], [[int(ids[2]), int(ids[5]), int(ids[7])], [int(ids[9]), int(ids[3])]]
        ns["remap"]("seg", None, ",".join(map(str, rm)), tuple(",".join(map(str, g)) for g in grp), 1)
    out["outlier"], out["size"], out["z"] = np.sort(seen["OutlierFilter"]), np.sort(seen["SizeFilter"]), np.sort(seen["ZFilter"])
    mapping = seen["Remap"]
    out["remap_keys"] = np.array(sorted(mapping), dtype=np.uint64)
    out["remap_vals"] = np.array([mapping[k] for k in sorted(mapping)], dtype=np.uint64)
    out["remap_remove"], out["remap_groups"] = np.array(rm, dtype=np.uint64), np.array([g + [0] * (3 - len(g)) for g in grp], dtype=np.uint64)
    uniq, sizes = ns["_global_sizes"](DS(frags))
    out["uniq"], out["sizes"] = uniq, sizes
    assert 0 < out["outlier"].size < uniq.size and 0 < out["size"].size < uniq.size and 0 < out["z"].size < uniq.size
    np.savez_compressed(os.path.join(OUT, "refine_filters.npz"), **out)
    print("refine_filters.npz", {k: v.shape for k, v in out.items() if k != "seg"})


TASK_STATE_CASES = [
    {"a": (4, 0, 0)},
    {"a": (4, 0, 0), "b": (4, 2, 0)},
    {"frags": (125, 3, 1), "agglom": (125, 0, 0), "relabel": (64, 0, 7)},
]


def golden_task_states():
    """blockwise.py `check_task_states` (:11-22), the function executed as it stands (extracted by ast: the module
    imports daisy): the RuntimeError message for failed / orphaned blocks."""
    import ast
    src = open(f"{REF}/blockwise.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "check_task_states")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "blockwise.check_task_states", "exec"), ns)
    out = []
    for case in TASK_STATE_CASES:
        states = {k: types.SimpleNamespace(total_block_count=t, failed_count=f, orphaned_count=o) for k, (t, f, o) in case.items()}
        try:
            ns["check_task_states"](states)
            msg = None
        except RuntimeError as e:
            msg = str(e)
        out.append(dict(states=case, message=msg))
    with open(os.path.join(OUT, "task_states.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("task_states.json", [o["message"] for o in out])


def golden_task_fields():
    """field names, defaults and task_type of the two volara task classes (post/blockwise/watershed_frags.py:30-62,
    waterz_agglom.py:39-75), read from the class bodies by ast (the modules need volara / pydantic models to import)."""
    import ast
    out = {}
    for fname, cname in (("watershed_frags.py", "WatershedFrags"), ("waterz_agglom.py", "WaterzAgglom")):
        tree = ast.parse(open(f"{REF}/post/blockwise/{fname}").read())
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cname)
        fields = {}
        for n in cls.body:
            if isinstance(n, ast.AnnAssign) and isinstance(n.target, ast.Name):
                try:
                    default = ast.literal_eval(n.value) if n.value is not None else "<required>"
                except ValueError:
                    default = "<expr>"
                fields[n.target.id] = default
        props = [n.name for n in cls.body if isinstance(n, ast.FunctionDef)]
        out[cname] = dict(fields=fields, methods=props)
    with open(os.path.join(OUT, "task_fields.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("task_fields.json", {k: len(v["fields"]) for k, v in out.items()})


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
    golden_filter_fragments()
    golden_compute_fragments_shift()
    golden_ws_glue()
    golden_mws_glue()
    golden_agglomerate_glue()
    golden_watershed_in_block_glue()
    golden_simple_watershed_glue()
    golden_waterz_pipeline_glue()
    golden_volara_pipeline_glue()
    golden_cc_affs_func()
    golden_refine_filters()
    golden_task_states()
    golden_task_fields()
    golden_naming()
    golden_config()
    print("golden fixtures written to", OUT)
