"""ORACLE: the blockwise `bs segment --ws -b` pipeline, restated on in-memory arrays.

Follows (file:line in /root/reference/bootstrapper):
  post/watershed.py:8-203                       waterz_pipeline (3 stages + barriers)
  post/blockwise/watershed_frags.py:115-246     WatershedFrags (stage 1)
  post/blockwise/waterz_agglom.py:106-170       WaterzAgglom   (stage 2)
  post/merge_tree.py                            MergeTree      (restated in merge_tree.py)
  post/watershed.py:206-354                     simple_watershed (single shot)
Third-party behaviour restated here (parity unpinned, SURVEY U-list):
  daisy block enumeration / block ids (U10), funlib.persistence zero-fill reads and edge
  ownership (U9), funlib.segment relabel / connected_components (U7, U8), volara
  replace_values / LUT / Relabel (U11), skimage remove_small_objects (U3).
All coordinates are in voxels (voxel_size only scales node positions).
"""
from dataclasses import dataclass, field

import numpy as np
from scipy.ndimage import center_of_mass, distance_transform_edt, gaussian_filter, label, maximum_filter
from scipy.ndimage import mean as ndi_mean

from .merge_tree import MergeTree
from .native import Waterz, connected_components, sk_label
from .ws import watershed_from_affinities

WS_DEFAULTS = dict(  # segment.py:11-23
    fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0,
    filter_fragments=0.1, remove_debris=64, thresholds=[0.2, 0.35, 0.5], merge_function="mean",
    sigma=None, noise_eps=None, bias=None)


# --------------------------------------------------------------------------- geometry
def pyramide_volume(dims, edge_length):
    if edge_length == 0:
        return 0
    v = 1
    for d in range(dims):
        v *= edge_length + d
    for d in range(dims):
        v //= d + 1
    return v


def cantor_number(coordinate):
    """funlib.math.cantor_number (used by daisy for block ids, U10)."""
    coordinate = tuple(int(c) for c in coordinate)
    if len(coordinate) == 1:
        return coordinate[0]
    return pyramide_volume(len(coordinate), sum(coordinate)) + cantor_number(coordinate[:-1])


@dataclass
class Block:
    index: tuple          # grid index of the block
    block_id: int         # cantor number (block.block_id[1])
    write_offset: tuple
    write_shape: tuple
    read_offset: tuple
    read_shape: tuple


def enumerate_blocks(roi_offset, roi_shape, block_size, context, index_offset=None):
    """daisy blocks of a volara BlockwiseTask with fit='shrink' (U10):
    write ROIs tile the task's write_roi from its offset in steps of block_size, the
    trailing blocks are clipped; read_roi = write_roi grown by context.
    Block ids number the ABSOLUTE block index write_roi.offset / write_roi.shape (floor division, U10): index_offset is
    the task ROI's absolute offset in voxels (dataset offset / voxel_size + roi_offset; default roi_offset)."""
    roi_offset = tuple(int(v) for v in roi_offset)
    index_offset = roi_offset if index_offset is None else tuple(int(v) for v in index_offset)
    roi_shape = tuple(int(v) for v in roi_shape)
    nb = [-(-s // b) for s, b in zip(roi_shape, block_size)]
    blocks = []
    for i in range(nb[0]):
        for j in range(nb[1]):
            for k in range(nb[2]):
                idx = (i, j, k)
                wo = tuple(o + n * b for o, n, b in zip(roi_offset, idx, block_size))
                ws = tuple(min(b, o + s - w) for b, o, s, w in zip(block_size, roi_offset, roi_shape, wo))
                ro = tuple(w - c for w, c in zip(wo, context))
                rs = tuple(s + 2 * c for s, c in zip(ws, context))
                aidx = tuple((io + n * b) // b for io, n, b in zip(index_offset, idx, block_size))
                blocks.append(Block(idx, cantor_number(aidx), wo, ws, ro, rs))
    return blocks


def to_ndarray(vol, offset, shape, fill_value=0):
    """funlib.persistence Array.to_ndarray(roi, fill_value=0): zero-fill outside the array.
    vol has spatial dims last (…, Z, Y, X), array offset 0."""
    lead = vol.shape[:-3]
    out = np.full(lead + tuple(shape), fill_value, dtype=vol.dtype)
    src, dst = [], []
    for o, s, n in zip(offset, shape, vol.shape[-3:]):
        lo, hi = max(o, 0), min(o + s, n)
        if hi <= lo:
            return out
        src.append(slice(lo, hi))
        dst.append(slice(lo - o, hi - o))
    out[(Ellipsis,) + tuple(dst)] = vol[(Ellipsis,) + tuple(src)]
    return out


# --------------------------------------------------------------------------- stage 1
def seeded_noise_block(shape, seed, block_id):
    """float64 (3, Z, Y, X): the seeded stand-in for the reference's unseeded np.random.randn(*affs_data.shape) of one block
    (watershed_frags.py:119-120): unit-variance sum of four 16-bit uniforms from a SplitMix64 hash of (seed, block id,
    channel, raveled voxel of the 4-D array), the generator csrc/stage1.cu implements bit for bit"""
    from bootstrapper_b200.synth import _hash
    n = int(np.prod(shape))
    nv = int(np.prod(shape[1:]))
    idx = np.arange(n, dtype=np.int64)
    h = _hash(seed, np.full(n, int(block_id), dtype=np.int64), idx // nv, idx, np.full(n, 13, dtype=np.int64))
    s = ((h & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + (h >> np.uint64(48)).astype(np.int64))
    return ((s - 131070).astype(np.float64) * (np.sqrt(3.0) / 65536.0)).reshape(shape)


def compute_fragments(affs_data, p, seed_tie="heap", block_id=0):
    """watershed_frags.py:115-146; noise_eps draws from the seeded generator (p["noise_seed"], block id) instead of the
    reference's unseeded np.random.randn."""
    affs_data = affs_data[:3]
    shift = np.zeros_like(affs_data)
    if p.get("noise_eps") is not None:
        shift += seeded_noise_block(affs_data.shape, p.get("noise_seed", 0) or 0, block_id) * p["noise_eps"]
    if p.get("sigma") is not None:
        shift += gaussian_filter(affs_data, sigma=(0, *p["sigma"])) - affs_data
    if p.get("bias") is not None:
        bias = p["bias"]
        bias = list(bias) if isinstance(bias, (list, tuple)) else [bias] * affs_data.shape[0]
        shift += np.array([bias]).reshape((-1, *((1,) * (len(affs_data.shape) - 1))))
    if p.get("seed_eps") is not None:
        boundary_mask = np.mean(affs_data, axis=0) > 0.5
        boundary_distances = distance_transform_edt(boundary_mask)
        max_filtered = maximum_filter(boundary_distances, p["min_seed_distance"])
        seeds, _ = label(max_filtered == boundary_distances)
        seeds[~boundary_mask] = 0
        shift -= p["seed_eps"] * distance_transform_edt(seeds == 0)
    fragments_data, _ = watershed_from_affinities(
        affs_data + shift, fragments_in_xy=p["fragments_in_xy"],
        min_seed_distance=p["min_seed_distance"], seed_tie=seed_tie)
    return fragments_data


def filter_avg_fragments(affs, fragments_data, filter_value):
    """watershed_frags.py:148-156 (volara.tmp.replace_values == in-place map)."""
    average_affs = np.mean(affs[0:3], axis=0)
    fragment_ids = np.unique(fragments_data)
    means = ndi_mean(average_affs, fragments_data, fragment_ids)
    filtered = np.array([f for f, m in zip(fragment_ids, means) if m < filter_value],
                        dtype=fragments_data.dtype)
    if filtered.size:
        fragments_data[np.isin(fragments_data, filtered)] = 0


def epsilon_agglomerate_fragments(affs_data, fragments_data, eps, stats_mode):
    """watershed_frags.py:158-177: waterz mean agglomeration, BinQueue<256>."""
    affs = np.ascontiguousarray(affs_data[:3].astype(np.float32))
    wz = Waterz(affs, fragments_data, 256, stats_mode)
    wz.merge_until(eps)
    fragments_data[:] = wz.segmentation()
    return fragments_data


def remove_small_objects(x, min_size):
    """skimage.morphology.remove_small_objects on an int label image (U3): value counts."""
    sizes = np.bincount(x.ravel())
    too_small = sizes < min_size
    out = x.copy()
    out[too_small[x]] = 0
    return out


def get_fragments(affs_data, p, seed_tie="heap", stats_mode="faithful", block_id=0):
    """watershed_frags.py:179-194."""
    fragments_data = compute_fragments(affs_data, p, seed_tie, block_id)
    if p["epsilon_agglomerate"] > 0:
        fragments_data = epsilon_agglomerate_fragments(affs_data, fragments_data,
                                                       p["epsilon_agglomerate"], stats_mode)
    if p["filter_fragments"] > 0:
        filter_avg_fragments(affs_data, fragments_data, p["filter_fragments"])
    if p["remove_debris"] > 0:
        dtype = fragments_data.dtype
        fragments_data = remove_small_objects(fragments_data.astype(np.int64),
                                              p["remove_debris"]).astype(dtype)
    return fragments_data


@dataclass
class Rag:
    """In-memory stand-in for the volara / funlib.persistence graph DB."""
    node_pos: dict = field(default_factory=dict)     # id -> (z, y, x) voxel position (world / voxel_size)
    node_size: dict = field(default_factory=dict)    # id -> voxel count
    edges: dict = field(default_factory=dict)        # (u, v) u<v -> merge_score (float or None)


def watershed_in_block(block, affs, frags_out, rag, p, roi_offset, block_size, mask=None,
                       seed_tie="heap", stats_mode="faithful", fragmenter=None):
    """watershed_frags.py:196-246.  frags_out is the task-ROI-sized uint64 array.
    fragmenter(affs_data, block) replaces get_fragments (volara ExtractFrags shares this skeleton, oracle/mws.py)."""
    affs_data = to_ndarray(affs, block.read_offset, block.read_shape, 0)
    if affs.dtype == np.uint8:
        max_affinity_value = 255.0
        affs_data = affs_data.astype(np.float64)
    else:
        max_affinity_value = 1.0
    if affs_data.max() < 1e-3:
        return
    affs_data /= max_affinity_value
    if mask is not None:
        mask_data = to_ndarray(mask, block.read_offset, block.read_shape, 0)
        if mask_data.ndim == 4:
            mask_data = (np.min(mask_data, axis=0) > 0).astype(np.uint8)
        if np.max(mask_data) == 255:
            mask_data = (mask_data > 0).astype(np.uint8)
        affs_data *= mask_data
    if fragmenter is not None:
        fragments_data = fragmenter(affs_data, block)
    else:
        fragments_data = get_fragments(affs_data, p, seed_tie, stats_mode, block.block_id)
    # crop to the write roi
    c0 = [w - r for w, r in zip(block.write_offset, block.read_offset)]
    sl = tuple(slice(c, c + s) for c, s in zip(c0, block.write_shape))
    fragments_data = fragments_data[sl]
    fragments_data, max_id = sk_label(fragments_data)
    fragments_data = fragments_data.astype(np.uint64)
    nvox = int(np.prod(block_size))
    assert max_id < nvox
    fragments_data[fragments_data > 0] += np.uint64(block.block_id * nvox)
    wsl = tuple(slice(w - o, w - o + s) for w, o, s in zip(block.write_offset, roi_offset, block.write_shape))
    frags_out[wsl] = fragments_data
    if fragments_data.max() == 0:
        return
    fragment_ids, counts = np.unique(fragments_data, return_counts=True)
    keep = fragment_ids > 0
    fragment_ids, counts = fragment_ids[keep], counts[keep]
    centers = center_of_mass(np.ones_like(fragments_data), fragments_data, list(fragment_ids))
    for fid, center, count in zip(fragment_ids, centers, counts):
        # position = write_roi.offset + voxel_size * Coordinate(center)  (Coordinate truncates)
        rag.node_pos[int(fid)] = tuple(int(w) + int(c) for w, c in zip(block.write_offset, center))
        rag.node_size[int(fid)] = int(count)


# --------------------------------------------------------------------------- stage 2
def funlib_relabel(a):
    """funlib.segment.arrays.relabel(a, return_backwards_map=True) (U8)."""
    old = np.unique(a)
    old = old[old != 0]
    back = np.concatenate([np.zeros(1, np.uint64), old.astype(np.uint64)])
    dense = np.searchsorted(back[1:], a).astype(np.uint64) + 1
    dense[a == 0] = 0
    return dense, len(old), back


def agglomerate_in_block(block, affs, frags, rag, roi_offset, stats_mode="faithful",
                         keep_cheaper=True, return_debug=False):
    """waterz_agglom.py:106-170.  frags is the task-ROI-sized fragment array (offset
    roi_offset); reads outside it are zero-filled like a funlib Array."""
    affs_data = to_ndarray(affs, block.read_offset, block.read_shape, 0)[:3]
    fo = [r - o for r, o in zip(block.read_offset, roi_offset)]
    frags_data = to_ndarray(frags, fo, block.read_shape, 0)
    frags_relabelled, _, relabel_map = funlib_relabel(frags_data)
    if affs_data.dtype != np.uint8:
        affs_data = affs_data.astype(np.float32)
    # (uint8 is normalised to float32 / 255 inside the restated waterz)
    wz = Waterz(affs_data, frags_relabelled, 256, stats_mode, keep_cheaper)
    wz.merge_until(0.0)
    u0, v0, s0, _, _ = wz.region_graph()           # initial RAG (threshold 0)
    a, b, c, sc = wz.merge_until(1.0)              # full merge history
    mt = MergeTree(relabel_map)
    for ai, bi, ci, si in zip(a, b, c, sc):
        mt.merge(relabel_map[ai], relabel_map[bi], relabel_map[ci], si)
    us = relabel_map[u0.astype(np.int64)]
    vs = relabel_map[v0.astype(np.int64)]
    scores = mt.find_merges(us, vs) if len(us) else np.zeros(0)
    # write_graph(rag, block.write_roi, write_nodes=False): an undirected edge is persisted
    # iff node min(u, v) has a stored position inside write_roi (U9)
    wlo = np.array(block.write_offset)
    whi = wlo + np.array(block.write_shape)
    written = 0
    for ui, vi, si in zip(us, vs, scores):
        lo, hi = (int(ui), int(vi)) if ui < vi else (int(vi), int(ui))
        pos = rag.node_pos.get(lo)
        if pos is None or not (np.all(np.array(pos) >= wlo) and np.all(np.array(pos) < whi)):
            continue
        rag.edges[(lo, hi)] = None if np.isnan(si) else float(si)
        written += 1
    if return_debug:
        return dict(history=(relabel_map[a.astype(np.int64)], relabel_map[b.astype(np.int64)], sc),
                    initial=(us, vs, s0), lca=scores, counters=wz.counters(), written=written)


# --------------------------------------------------------------------------- stage 3
def global_segmentation(frags, rag, thresholds):
    """post/watershed.py:156-203: thresholded CC -> LUT -> relabel."""
    nodes = np.array(sorted(rag.node_pos.keys()), dtype=np.uint64)
    out = {}
    if nodes.size == 0:
        return out
    items = [(u, v, s) for (u, v), s in rag.edges.items() if s is not None]
    edges = (np.array([(u, v) for u, v, _ in items], dtype=np.uint64)
             if items else np.zeros((0, 2), np.uint64))
    scores = np.array([s for _, _, s in items], dtype=np.float32)
    for thr in thresholds:
        if edges.shape[0] == 0:
            components = nodes.copy()
        else:
            components = connected_components(nodes, edges, scores, thr)
        lut = np.array([nodes, components])
        seg = frags.copy()
        idx = np.searchsorted(nodes, frags)
        idx[idx >= nodes.size] = 0
        hit = nodes[idx] == frags
        seg[hit] = components[idx[hit]]
        out[thr] = dict(lut=lut, seg=seg)
    return out


def waterz_pipeline(affs, params=None, block_size=None, context=None, roi=None, mask=None,
                    seed_tie="heap", stats_mode="faithful", keep_cheaper=True, index_offset=None):
    """post/watershed.py:8-203 on in-memory arrays.  affs: (C,Z,Y,X) uint8 or float.
    index_offset: absolute offset of the ROI in voxels (block ids, U10); default roi_offset."""
    p = dict(WS_DEFAULTS)
    p.update(params or {})
    vol_shape = affs.shape[1:]
    roi_offset, roi_shape = roi if roi is not None else ((0, 0, 0), vol_shape)
    if block_size is None:        # blockwise False / block_shape == "roi"   (watershed.py:84-86)
        block_size = tuple(vol_shape)
        context = (0, 0, 0)
    elif context is None:         # watershed.py:79-83
        context = tuple(max(1, s // 8) for s in block_size)
    blocks = enumerate_blocks(roi_offset, roi_shape, block_size, context, index_offset)
    frags = np.zeros(roi_shape, dtype=np.uint64)
    rag = Rag()
    for blk in blocks:
        watershed_in_block(blk, affs, frags, rag, p, roi_offset, block_size, mask, seed_tie, stats_mode)
    for blk in blocks:
        agglomerate_in_block(blk, affs, frags, rag, roi_offset, stats_mode, keep_cheaper)
    segs = global_segmentation(frags, rag, p["thresholds"])
    return dict(fragments=frags, rag=rag, segs=segs, blocks=blocks, params=p)


# --------------------------------------------------------------------------- single shot
def merge_function_code(name):
    """post/watershed.py:232-244: 'mean' -> (0, False); 'hist_quant_Q[_initmax]' -> (Q, initmax)"""
    if name == "mean":
        return 0, False
    parts = name.split("_")
    assert parts[:2] == ["hist", "quant"] and int(parts[2]) in (10, 25, 50, 75, 90), name
    return int(parts[2]), len(parts) == 4 and parts[3] == "initmax"


def simple_watershed(affs, params=None, mask=None, seed_tie="heap", stats_mode="faithful",
                     keep_cheaper=True):
    """post/watershed.py:206-354 on in-memory arrays: float32 normalise, fragments,
    waterz with the default (non-discretised) queue over the sorted thresholds."""
    p = dict(WS_DEFAULTS)
    p.update(params or {})
    affs_data = affs[:3]
    raw_u8 = affs_data if affs_data.dtype == np.uint8 else None
    if affs_data.dtype == np.uint8:
        affs_data = affs_data.astype(np.float32) / 255.0
    else:
        affs_data = affs_data.astype(np.float32)
    if mask is not None:
        affs_data = affs_data * (mask > 0).astype(np.uint8)
        raw_u8 = None
    assert p.get("noise_eps") is None
    if any([p.get("sigma"), p.get("bias")]):
        raw_u8 = None
        shift = np.zeros_like(affs_data)
        if p.get("sigma") is not None:
            shift += gaussian_filter(affs_data, sigma=(0, *p["sigma"])) - affs_data
        if p.get("bias") is not None:
            bias = p["bias"]
            bias = [bias] * 3 if isinstance(bias, float) else list(bias)
            shift += np.array([bias]).reshape((-1, 1, 1, 1))
        affs_data += shift
    fragments_data, n = watershed_from_affinities(
        affs_data, fragments_in_xy=p["fragments_in_xy"], return_seeds=False,
        min_seed_distance=p["min_seed_distance"], seed_tie=seed_tie)
    thresholds = sorted(p["thresholds"])
    quantile, initmax = merge_function_code(p["merge_function"])
    wz = Waterz(raw_u8 if (raw_u8 is not None and stats_mode == "canonical") else affs_data,
                fragments_data, 0, stats_mode, keep_cheaper, quantile=quantile, initmax=initmax)
    segs = {}
    for thr in thresholds:
        wz.merge_until(thr)
        segs[thr] = wz.segmentation()
    return dict(fragments=fragments_data, n=n, segs=segs, params=p)
