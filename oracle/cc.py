"""ORACLE (test infrastructure, never on the product path).

CPU restatement of `bs segment --cc` (post/connected_components.py:15-127 + post/cc.py:7-74): thresholded
affinities, components of the "+e_d" graph, uint32 ids in raster order of each component's first voxel.
The reference's serial flood is restated with scipy.sparse.csgraph; pinned against tests/golden/cc_flood.npz and
cc_affs.npz, which were produced by executing the reference's own post/cc.py (tests/golden/make_golden.py).
remove_small_objects (skimage, absent here) follows SURVEY U3: value counts, `count < min_size` removed -- unpinned.
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def hard_affs(affs, threshold, mask=None, sigma=None):
    """connected_components.py:52-81: float32 normalise, mask multiply, optional gaussian shift, compare."""
    from scipy.ndimage import gaussian_filter
    data = affs[:3]
    data = data.astype(np.float32) / 255.0 if data.dtype == np.uint8 else data.astype(np.float32)
    if mask is not None:
        data = data * (mask > 0).astype(np.uint8)
    if sigma is not None:
        shift = np.zeros_like(data)
        shift += gaussian_filter(data, sigma=(0, *sigma)) - data
        data = data + shift
    return data > threshold


def compute_connected_component_segmentation(hard):
    """post/cc.py:7-74 without the serial flood."""
    shape = hard.shape[1:]
    n = int(np.prod(shape))
    idx = np.arange(n).reshape(shape)
    rows, cols = [], []
    for d in range(3):
        sl_a = [slice(None)] * 3
        sl_b = [slice(None)] * 3
        sl_a[d] = slice(0, shape[d] - 1)
        sl_b[d] = slice(1, shape[d])
        h = hard[d][tuple(sl_a)]
        rows.append(idx[tuple(sl_a)][h])
        cols.append(idx[tuple(sl_b)][h])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    g = coo_matrix((np.ones(len(rows), np.uint8), (rows, cols)), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    labelled = hard.any(axis=0).ravel()
    labelled[cols] = True                       # reached through a lower neighbour's affinity
    first = np.full(comp.max() + 1, n, np.int64)
    np.minimum.at(first, comp[labelled], np.flatnonzero(labelled))
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    seg = np.zeros(n, np.uint32)
    seg[labelled] = rank[comp[labelled]] + 1    # components without a labelled voxel sort last and are unused
    return seg.reshape(shape)


def remove_small_objects(x, min_size):
    sizes = np.bincount(x.ravel())
    out = x.copy()
    out[(sizes < min_size)[x]] = 0
    return out


def cc_affs(affs, threshold=0.5, remove_debris=0, mask=None, sigma=None):
    frags = compute_connected_component_segmentation(hard_affs(affs, threshold, mask, sigma))
    seg = frags
    if remove_debris > 0:
        seg = remove_small_objects(frags.astype(np.int64), remove_debris).astype(frags.dtype)
    return frags, seg
