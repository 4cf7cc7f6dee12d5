"""Affinity self-consistency error of a segmentation: the compute of the reference's gunpowder node
`AddAffErrors` (bootstrapper/gp/add_aff_errors.py:12-183) as used by `bs evaluate` (eval/compute_errors.py:179-213),
on device arrays.  The gunpowder plumbing around it (requests, padding, zarr sources) is not reproduced; the node's
`process` body is: affinities of the segmentation on the neighbourhood -> error map -> error mask."""
import torch

from .. import native


def add_aff_errors(segmentation, pred_affs, neighborhood, labels_mask=None, thresholds=(0.1, 1.0)):
    """segmentation: CUDA tensor (Z,Y,X) of ids; pred_affs: CUDA tensor (C,Z,Y,X) float32, or uint8 (normalised as
    gp.Normalize does, compute_errors.py:146); neighborhood: C offsets (z,y,x); labels_mask: optional (Z,Y,X) uint8.
    Returns dict(seg_affs float32 (C,Z,Y,X), error_map float32, error_mask uint8, error_map_u8) -- error_map_u8 is what
    `bs evaluate` writes (IntensityScaleShift(255, 0) + AsType(uint8), compute_errors.py:203-204)."""
    if not segmentation.is_cuda:
        raise native.BsError("add_aff_errors needs CUDA tensors (no CPU fallback)")
    seg = segmentation.contiguous()
    if seg.dtype != torch.int64:
        seg = seg.to(torch.int64)
    seg_affs, err, emask = native.aff_errors(seg, pred_affs.contiguous(), neighborhood,
                                             None if labels_mask is None else labels_mask.contiguous().to(torch.uint8), thresholds)
    return dict(seg_affs=seg_affs, error_map=err, error_mask=emask, error_map_u8=(err * 255).to(torch.uint8))
