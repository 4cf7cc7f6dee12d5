"""Shared geometry of the two blockwise ws tasks (volara BlockwiseTask as the reference uses it)."""
import numpy as np
import torch

from ... import native
from ...datasets import Block, Roi


class BlockwiseTask:
    task_type = "task"
    fit = "shrink"
    read_write_conflict = False

    # subclasses set: affs_data, frags_data, block_size, context, roi (offset, shape) world units or None,
    # num_workers, plus the ws parameters
    def _affs_array(self):
        return self.affs_data.array("r")

    @property
    def voxel_size(self):
        return self._affs_array().voxel_size

    @property
    def write_roi(self):
        a = self._affs_array()
        off, shp = a.roi
        if self.roi is not None:
            lo = tuple(max(o, r) for o, r in zip(off, self.roi[0]))
            hi = tuple(min(o + s, r + t) for o, s, r, t in zip(off, shp, self.roi[0], self.roi[1]))
            off, shp = lo, tuple(h - l for l, h in zip(lo, hi))
        return Roi(tuple(off), tuple(shp))

    @property
    def write_size(self):
        return tuple(b * v for b, v in zip(self.block_size, self.voxel_size))

    @property
    def context_size(self):
        return tuple(c * v for c, v in zip(self.context, self.voxel_size))

    @property
    def num_voxels_in_block(self):
        return int(np.prod(self.block_size))

    def drop(self):
        self.drop_artifacts()

    # ---- device residency + plan
    def _voxel_roi(self):
        a = self._affs_array()
        vs = a.voxel_size
        w = self.write_roi
        off = tuple((o - ao) // v for o, ao, v in zip(w.offset, a.offset, vs))
        shp = tuple(s // v for s, v in zip(w.shape, vs))
        return off, shp

    def _load_affs(self):
        if getattr(self, "_affs_dev", None) is None:
            a = self._affs_array()
            data = a.read()
            if data.shape[0] < 3:
                raise ValueError("the ws path needs at least 3 affinity channels")
            self._affs_dev = torch.from_numpy(np.ascontiguousarray(data)).cuda()
        return self._affs_dev

    def _make_plan(self, **kw):
        affs = self._load_affs()
        off, shp = self._voxel_roi()
        # daisy block ids come from the ABSOLUTE write-ROI offset: world offset / voxel size
        absolute = tuple(int(o) // int(v) for o, v in zip(self.write_roi.offset, self.voxel_size))
        return native.Plan(tuple(affs.shape[1:]), tuple(self.block_size), tuple(self.context), native._aff_dtype(affs),
                           roi_offset=off, roi_shape=shp, n_channels=affs.shape[0], block_index_offset=absolute, **kw)

    def blocks(self):
        """daisy blocks (ascending block id), world-unit ROIs."""
        plan = self._plan()
        ids, wo, ws = plan.block_info()
        a = self._affs_array()
        vs, ao = a.voxel_size, a.offset
        out = []
        for i in range(len(ids)):
            w_off = tuple(int(o) * v + a0 for o, v, a0 in zip(wo[i], vs, ao))
            w_shp = tuple(int(s) * v for s, v in zip(ws[i], vs))
            ctx = self.context_size
            r_off = tuple(o - c for o, c in zip(w_off, ctx))
            r_shp = tuple(s + 2 * c for s, c in zip(w_shp, ctx))
            out.append(Block(Roi(r_off, r_shp), Roi(w_off, w_shp), (self.task_name, int(ids[i])), i))
        return out
