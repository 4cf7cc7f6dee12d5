"""post/mws.py plug point — `mwatershed_from_affinities` on a CUDA tensor (reference post/mws.py:12-59).

The mutex watershed runs in libbsnative (bs_mws_agglom, csrc/mws.cu): weights = affs + shift in float64, edges by
descending |w| with the declared tie rule D4 (ascending (channel, raveled voxel)).  Differences from the reference call,
all loud:
  * noise_eps: the reference draws UNSEEDED np.random.randn noise; here a seeded counter-based generator (noise_seed)
    stands in, mirrored bit for bit by the oracle;
  * sigma (gaussian shift) and randomized_strides=True (an unseeded random subset of the stride lattice) raise
    NotImplementedError.
There is no CPU fallback.
"""
import torch

from .. import native


def mwatershed_from_affinities(affs, neighborhood, bias, sigma=None, noise_eps=None, strides=None, randomized_strides=False,
                               mask=None, noise_seed=0, remove_debris=0, return_counters=False):
    """affs: CUDA tensor (C, Z, Y, X) uint8 (normalised / 255 as simple_mutex does) or float32, already holding what the
    reference passes as float64 `affs`.  Returns the uint64 fragments as a torch.int64 tensor (Z, Y, X)."""
    if not affs.is_cuda:
        raise native.BsError("mwatershed_from_affinities needs the affinities on a CUDA device (no CPU fallback)")
    if sigma is not None:
        raise NotImplementedError("mws parameter 'sigma' is not implemented in the CUDA path")
    if randomized_strides:
        raise NotImplementedError("randomized_strides=True draws an unseeded random subset of the stride lattice in mwatershed: "
                                  "not reproducible, not implemented; set randomized_strides = false")
    if affs.dtype == torch.float64:
        raise NotImplementedError("float64 affinities are not supported; pass uint8 or float32")
    frags, seg, counters = native.mws_agglom(affs.contiguous(), neighborhood, bias, strides=strides, mask=mask, noise_eps=noise_eps,
                                             noise_seed=noise_seed, remove_debris=remove_debris)
    if return_counters:
        return frags, seg, counters
    return frags
