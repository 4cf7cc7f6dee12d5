"""`bs refine` filters on device label volumes: the compute of bootstrapper/refine.py's
`outlier_filter` (:147-172), `size_filter` (:190-213), `z_filter` (:229-258) and `remap` (:281-307).

The reference scans the zarr array tile by tile with fastremap, decides on the host which ids to drop, then rewrites
the array blockwise (`fastremap.mask` / `fastremap.remap`).  Here one kernel builds the per-id table (voxel count,
first / last z plane; `bs_label_stats`), the decisions are the reference's own numpy arithmetic on that table, and the
rewrite is the LUT relabel kernel (`bs_relabel`: ids absent from the LUT are unchanged).  The zarr / click plumbing is
not reproduced; every function takes and returns CUDA tensors."""
import numpy as np
import torch

from . import native


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


def global_sizes(seg):
    """refine.py:98-108 `_global_sizes`: (ids ascending, voxel counts int64) of the non-zero labels"""
    ids, sizes, _, _ = native.label_stats(seg.contiguous())
    return _u64(ids), sizes.cpu().numpy()


def z_extents(seg):
    """refine.py:236-255: (ids, z-extent = last plane - first plane + 1) of the non-zero labels"""
    ids, _, zlo, zhi = native.label_stats(seg.contiguous())
    return _u64(ids), (zhi.cpu().numpy().astype(np.int64) - zlo.cpu().numpy().astype(np.int64) + 1)


def mask_ids(seg, remove_ids):
    """refine.py:111-116 `_mask_block` (fastremap.mask): listed ids -> 0"""
    remove_ids = np.unique(np.asarray(remove_ids, dtype=np.uint64))
    if remove_ids.size == 0:
        return seg.clone()
    keys = torch.from_numpy(remove_ids.view(np.int64)).to(seg.device)
    return native.relabel(seg.contiguous(), keys, torch.zeros_like(keys))


def outlier_filter(seg, num_std, min_size=0):
    """refine.py:147-172: two-sided sigma cut on the object sizes (statistics over objects >= min_size).
    Returns (filtered volume, removed ids, dict(mean, std, lo, hi))."""
    uniq, sizes = global_sizes(seg)
    if uniq.size == 0:
        raise ValueError("no foreground objects in volume")
    stat_sizes = sizes[sizes >= min_size]
    if stat_sizes.size == 0:
        raise ValueError(f"no objects with size >= min_size ({min_size})")
    mean, std = float(stat_sizes.mean()), float(stat_sizes.std())
    lo, hi = mean - num_std * std, mean + num_std * std
    remove_ids = uniq[(sizes < lo) | (sizes > hi)]
    return mask_ids(seg, remove_ids), remove_ids, dict(mean=mean, std=std, lo=lo, hi=hi)


def size_filter(seg, min_size=0, max_size=None):
    """refine.py:190-213: drop objects outside [min_size, max_size] (0 / None = no cap)"""
    uniq, sizes = global_sizes(seg)
    if uniq.size == 0:
        raise ValueError("no foreground objects in volume")
    remove = np.zeros(uniq.size, dtype=bool)
    if min_size > 0:
        remove |= sizes < min_size
    if max_size:
        remove |= sizes > max_size
    remove_ids = uniq[remove]
    return mask_ids(seg, remove_ids), remove_ids


def z_filter(seg, min_z=1):
    """refine.py:229-258: drop objects whose z-extent is min_z planes or fewer"""
    ids, spans = z_extents(seg)
    remove_ids = ids[spans <= min_z]
    return mask_ids(seg, remove_ids), remove_ids


def remap(seg, remove_ids=(), merge_groups=()):
    """refine.py:281-307: remove ids (-> 0) and / or merge each group into its first id; other ids are unchanged"""
    remove = {int(x) for x in remove_ids}
    merge = {}
    for group in merge_groups:
        ids = [int(x) for x in group]
        for mid in ids:
            merge[mid] = ids[0]
    conflict = remove & set(merge)
    if conflict:
        raise ValueError(f"ids given to both remove_ids and merge_groups: {sorted(conflict)}")
    mapping = {i: 0 for i in remove} | merge
    if not mapping:
        raise ValueError("nothing to do: pass remove_ids and/or merge_groups")
    keys = np.array(sorted(mapping), dtype=np.uint64)
    vals = np.array([mapping[int(k)] for k in keys], dtype=np.uint64)
    return native.relabel(seg.contiguous(), torch.from_numpy(keys.view(np.int64)).to(seg.device),
                          torch.from_numpy(vals.view(np.int64)).to(seg.device))
