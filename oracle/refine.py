"""CPU restatement of the `bs refine` filters -- TEST INFRASTRUCTURE ONLY.

Follows bootstrapper/refine.py: `_global_sizes` :98-108, `outlier_filter` :147-172, `size_filter` :190-213, `z_filter`
:229-258, `_mask_block` :111-116, `remap` / `_remap_block` :265-307.  The reference calls fastremap (unique / mask /
remap; absent here, unpinned dependency) on zarr tiles; their documented behaviour is restated with numpy
[3P-recall].  The decision arithmetic is PINNED: `_global_sizes` (over z tiles), the three filters and the remap table
are checked against the reference's own functions executed on an in-memory array
(tests/golden/refine_filters.npz); only fastremap's mask / remap of the rewrite stay recalled."""
import numpy as np


def global_sizes(seg):
    uniq, counts = np.unique(seg, return_counts=True)
    fg = uniq != 0
    return uniq[fg], counts[fg].astype(np.int64)


def mask_ids(seg, remove_ids):
    out = seg.copy()
    out[np.isin(out, remove_ids)] = 0
    return out


def outlier_filter(seg, num_std, min_size=0):
    uniq, sizes = global_sizes(seg)
    stat_sizes = sizes[sizes >= min_size]
    mean, std = float(stat_sizes.mean()), float(stat_sizes.std())
    lo, hi = mean - num_std * std, mean + num_std * std
    remove_ids = uniq[(sizes < lo) | (sizes > hi)]
    return mask_ids(seg, remove_ids), remove_ids


def size_filter(seg, min_size=0, max_size=None):
    uniq, sizes = global_sizes(seg)
    remove = np.zeros(uniq.size, dtype=bool)
    if min_size > 0:
        remove |= sizes < min_size
    if max_size:
        remove |= sizes > max_size
    return mask_ids(seg, uniq[remove]), uniq[remove]


def z_filter(seg, min_z=1):
    zmin, zmax = {}, {}
    for gz in range(seg.shape[0]):
        present = np.unique(seg[gz])
        for lbl in present[present != 0].tolist():
            if lbl not in zmin:
                zmin[lbl] = gz
                zmax[lbl] = gz
            else:
                zmin[lbl] = min(zmin[lbl], gz)
                zmax[lbl] = max(zmax[lbl], gz)
    ids = np.fromiter(zmin.keys(), dtype=seg.dtype, count=len(zmin))
    spans = np.array([zmax[int(i)] - zmin[int(i)] + 1 for i in ids], dtype=np.int64)
    remove_ids = ids[spans <= min_z]
    return mask_ids(seg, remove_ids), remove_ids


def remap(seg, mapping):
    """fastremap.remap(data, mapping, preserve_missing_labels=True)"""
    out = seg.copy()
    for k, v in mapping.items():
        out[seg == k] = v
    return out
