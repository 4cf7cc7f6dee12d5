"""ORACLE: seeded watershed on boundary distance — follows post/ws.py line by line.

scipy.ndimage is executed directly (installed here); skimage.watershed is the C
restatement in oracle/csrc/skimage_restated.c (parity unpinned, see oracle/__init__).
"""
import numpy as np
from scipy.ndimage import distance_transform_edt, label, maximum_filter

from .native import sk_watershed


def watershed_from_boundary_distance(boundary_distances, boundary_mask, return_seeds=False,
                                     id_offset=0, min_seed_distance=10, seed_tie="heap"):
    """post/ws.py:8-35."""
    max_filtered = maximum_filter(boundary_distances, min_seed_distance)      # ws.py:16
    maxima = max_filtered == boundary_distances                               # ws.py:17
    seeds, n = label(maxima)                                                  # ws.py:19
    if n == 0:                                                                # ws.py:21-22
        return np.zeros(boundary_distances.shape, dtype=np.uint64), id_offset
    seeds[seeds != 0] += id_offset                                            # ws.py:24
    fragments = sk_watershed(boundary_distances.max() - boundary_distances,   # ws.py:26-28
                             seeds, boundary_mask, seed_tie=seed_tie)
    ret = (fragments.astype(np.uint64), n + id_offset)
    if return_seeds:
        ret = ret + (seeds.astype(np.uint64),)
    return ret


def watershed_from_affinities(affs, max_affinity_value=1.0, fragments_in_xy=False,
                              return_seeds=False, min_seed_distance=10, seed_tie="heap"):
    """post/ws.py:38-112.  Returns (fragments, max_id[, seeds])."""
    if fragments_in_xy:
        mean_affs = 0.5 * (affs[-1] + affs[-2])                               # ws.py:64
        depth = mean_affs.shape[0]
        fragments = np.zeros(mean_affs.shape, dtype=np.uint64)
        if return_seeds:
            seeds = np.zeros(mean_affs.shape, dtype=np.uint64)
        id_offset = 0
        for z in range(depth):                                                # ws.py:75-92
            boundary_mask = mean_affs[z] > 0.5 * max_affinity_value
            boundary_distances = distance_transform_edt(boundary_mask)
            ret = watershed_from_boundary_distance(
                boundary_distances, boundary_mask, return_seeds=return_seeds,
                id_offset=id_offset, min_seed_distance=min_seed_distance, seed_tie=seed_tie)
            fragments[z] = ret[0]
            if return_seeds:
                seeds[z] = ret[2]
            id_offset = ret[1]
        ret = (fragments, id_offset)
        if return_seeds:
            ret += (seeds,)
    else:
        boundary_mask = np.mean(affs, axis=0) > 0.5 * max_affinity_value      # ws.py:100
        boundary_distances = distance_transform_edt(boundary_mask)
        ret = watershed_from_boundary_distance(
            boundary_distances, boundary_mask, return_seeds,
            min_seed_distance=min_seed_distance, seed_tie=seed_tie)
    return ret
