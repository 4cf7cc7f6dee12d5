"""ORACLE (test infrastructure): the mutex-watershed path of `bs segment --mws`, restated.

* mwatershed_from_affinities: post/mws.py:12-59 line by line (shift = noise + (gauss - affs) + bias; mws.agglom on
  float64).  The reference's noise is an UNSEEDED np.random.randn: parity is impossible there, so `noise_seed` selects a
  seeded counter-based stand-in (sum of four 16-bit uniforms from a SplitMix64 hash of (seed, channel, voxel), unit
  variance) that the CUDA path (csrc/mws.cu) implements bit for bit.  randomized_strides draws an unseeded random subset of
  the stride lattice upstream: not restated, must be False.
* simple_mutex: post/watershed_mutex.py:177-291 on in-memory arrays (normalise, mask, fragments, remove_debris).
* mwatershed.agglom itself: oracle/csrc/mws_restated.cpp (PARITY UNPINNED, declared tie rule D4).
"""
import numpy as np
from scipy.ndimage import gaussian_filter

from . import native as on
from bootstrapper_b200.synth import _hash

NOISE_K = np.sqrt(3.0) / 65536.0


def seeded_noise(shape, seed):
    """float64 array (C, Z, Y, X): the seeded stand-in for np.random.randn(*shape) (see module docstring)"""
    C = shape[0]
    V = int(np.prod(shape[1:]))
    p = np.arange(V, dtype=np.int64)
    out = np.empty((C, V), dtype=np.float64)
    for c in range(C):
        h = _hash(seed, np.full(V, c, dtype=np.int64), p, np.full(V, 11, dtype=np.int64))
        s = ((h & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + (h >> np.uint64(48)).astype(np.int64))
        out[c] = (s - 131070).astype(np.float64) * NOISE_K
    return out.reshape(shape)


def mwatershed_from_affinities(affs, neighborhood, bias, sigma=None, noise_eps=None, strides=None, randomized_strides=False,
                               noise_seed=0, counters=None):
    """post/mws.py:12-59"""
    assert not randomized_strides, "randomized_strides is an unseeded RNG upstream"
    if sigma is not None:
        sigma = (0, *sigma)
    shift = np.zeros_like(affs)
    if noise_eps is not None:
        shift += seeded_noise(affs.shape, noise_seed) * noise_eps
    if sigma is not None:
        shift += gaussian_filter(affs, sigma=sigma) - affs
    shift += np.array([bias]).reshape((-1, *((1,) * (len(affs.shape) - 1))))
    return on.mws_agglom((affs + shift).astype(np.float64), neighborhood, strides, counters=counters)


def remove_small_objects(x, min_size):
    """skimage.morphology.remove_small_objects on a label array (U3): labels with fewer than min_size voxels -> 0"""
    out = x.copy()
    ids, counts = np.unique(x, return_counts=True)
    small = ids[(counts < min_size) & (ids != 0)]
    out[np.isin(x, small)] = 0
    return out


def simple_mutex(affs, params, mask=None, noise_seed=0):
    """post/watershed_mutex.py:177-291 on in-memory arrays: affs (C, Z, Y, X) uint8 or float.
    Returns dict(fragments, seg)."""
    neighborhood, bias = params["aff_neighborhood"], params["bias"]
    assert len(neighborhood) == affs.shape[0] == len(bias)
    if affs.dtype == np.uint8:
        affs_data = affs.astype(np.float64) / 255.0
    else:
        affs_data = affs.astype(np.float64)
    if mask is not None:
        affs_data *= (mask > 0).astype(np.uint8)
    frags = mwatershed_from_affinities(affs_data, neighborhood, bias, params.get("sigma"), params.get("noise_eps"),
                                       params.get("strides"), params.get("randomized_strides", False), noise_seed=noise_seed)
    seg = frags
    rd = params.get("remove_debris", 0)
    if rd and rd > 0:
        seg = remove_small_objects(frags.astype(np.int64), rd).astype(frags.dtype)
    return dict(fragments=frags, seg=seg)
