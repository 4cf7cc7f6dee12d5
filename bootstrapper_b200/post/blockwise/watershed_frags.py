"""WatershedFrags — drop-in for the reference's blockwise fragments task
(post/blockwise/watershed_frags.py:31-258): same fields, properties and process_block contract;
the per-block numerics run in libbsnative (bs_stage1_fragments), all blocks in one batched pass.
"""
from contextlib import contextmanager

import numpy as np
import torch

from ..pipeline import resolve_ws_params
from .base import BlockwiseTask


class WatershedFrags(BlockwiseTask):
    task_type = "watershed-frags"
    _out_array_dtype = np.dtype(np.uint64)

    def __init__(self, db, affs_data, frags_data, block_size, context, mask_data=None, num_workers=1, roi=None,
                 fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0, sigma=None,
                 noise_eps=None, bias=None, filter_fragments=0.0, remove_debris=0, noise_seed=0):
        self.db, self.affs_data, self.frags_data, self.mask_data = db, affs_data, frags_data, mask_data
        self.block_size, self.context = tuple(int(v) for v in block_size), tuple(int(v) for v in context)
        self.num_workers, self.roi = num_workers, roi
        self.params = resolve_ws_params(dict(
            fragments_in_xy=fragments_in_xy, min_seed_distance=min_seed_distance, seed_eps=seed_eps,
            epsilon_agglomerate=epsilon_agglomerate, sigma=sigma, noise_eps=noise_eps, noise_seed=noise_seed, bias=bias,
            filter_fragments=filter_fragments, remove_debris=remove_debris))
        self._plan_obj = None
        self._frags_dev = None
        self._mask_cache = None

    @property
    def task_name(self):
        return f"{self.frags_data.name}-{self.task_type}"

    @property
    def output_datasets(self):
        return [self.frags_data]

    def drop_artifacts(self):
        self.frags_data.drop()
        self.db.drop()

    def init(self):
        self.db.init()
        self.init_out_array()

    def init_out_array(self):
        a = self._affs_array()
        vs = self.voxel_size
        w = self.write_roi
        self.frags_data.prepare(tuple(s // v for s, v in zip(w.shape, vs)), self.block_size, w.offset, vs, units=a.units,
                                axis_names=a.axis_names[1:] if a.axis_names else None,
                                types=a.types[1:] if a.types else None, dtype=self._out_array_dtype)

    def _plan(self):
        if self._plan_obj is None:
            p = self.params
            self._plan_obj = self._make_plan(fragments_in_xy=p["fragments_in_xy"], min_seed_distance=p["min_seed_distance"],
                                             filter_fragments=p["filter_fragments"], remove_debris=p["remove_debris"],
                                             bias=p["bias"], seed_eps=p["seed_eps"], sigma=p["sigma"],
                                             noise_eps=p["noise_eps"], noise_seed=p.get("noise_seed", 0) or 0)
        return self._plan_obj

    def _mask_dev(self):
        """the mask on the affinity array's voxel grid (watershed_frags.py:207-213 reads mask.to_ndarray(block.read_roi,
        fill_value=0)): cropped / zero-padded in world units, so a smaller or shifted mask dataset lines up; a different
        voxel size is refused.  The reference multiplies the affinities by the mask's raw values unless they are 0 / 255 or
        the mask has channels; masks holding other values than 0, 1, 255 are refused rather than silently binarised."""
        if self.mask_data is None:
            return None
        if getattr(self, "_mask_cache", None) is not None:
            return self._mask_cache
        ma, a = self.mask_data.array("r"), self._affs_array()
        vs = tuple(a.voxel_size)
        if tuple(ma.voxel_size) != vs:
            raise ValueError(f"mask voxel size {tuple(ma.voxel_size)} differs from the affinities' {vs}")
        if any((ao - mo) % v for ao, mo, v in zip(a.offset, ma.offset, vs)):
            raise ValueError("mask offset is not aligned with the affinity voxel grid")
        start = tuple((ao - mo) // v for ao, mo, v in zip(a.offset, ma.offset, vs))
        m = ma.to_ndarray(start, a.spatial_shape, fill_value=0)
        if m.ndim == 4:                                   # watershed_frags.py:209-210
            m = (m.min(axis=0) > 0).astype(np.uint8)
        else:
            vals = np.unique(m)
            if not np.isin(vals, (0, 1, 255)).all():
                raise NotImplementedError("mask values other than 0 / 1 / 255 scale the affinities in the reference "
                                          "(watershed_frags.py:211-213); binarise the mask first")
            if vals.size and vals.max() == 255 and (vals == 1).any():
                raise NotImplementedError("a mask mixing the values 1 and 255 is binarised per block in the reference; "
                                          "binarise the mask first")
            m = (m > 0).astype(np.uint8)
        self._mask_cache = torch.from_numpy(np.ascontiguousarray(m)).cuda()
        return self._mask_cache

    def _frags(self):
        if self._frags_dev is None:
            self._frags_dev = torch.zeros(self._plan().roi_shape, dtype=torch.int64, device="cuda")
        return self._frags_dev

    def _store(self, plan, block_indices=None):
        """write fragments (write ROIs of the processed blocks) and their RAG nodes"""
        frags = self._frags()
        out = self.frags_data.array("r+")
        ids, wo, ws = plan.block_info()
        off, _ = self._voxel_roi()
        sel = range(len(ids)) if block_indices is None else block_indices
        if block_indices is None:
            out.write(frags.cpu().numpy().view(np.uint64))
        else:
            for i in sel:
                lo = [int(o) - r for o, r in zip(wo[i], off)]
                sl = tuple(slice(l, l + int(s)) for l, s in zip(lo, ws[i]))
                out.write(frags[sl].cpu().numpy().view(np.uint64), start=lo)
        nid, npos, nsz = [t.cpu().numpy() for t in plan.nodes("cuda")]
        if len(nid):
            a = self._affs_array()
            world = npos.astype(np.int64) * np.array(self.voxel_size) + np.array(a.offset)
            self.db.write_nodes(nid.view(np.uint64), world, nsz.astype(np.int64))

    def run_all(self):
        plan = self._plan()
        n, _ = plan.num_blocks()
        plan.set_owned(np.arange(n))
        if self.params["epsilon_agglomerate"] and self.params["epsilon_agglomerate"] > 0:
            # watershed_frags.py:158-176, 182-183: waterz merges of every block's fragments before filter / crop / relabel
            from ..pipeline import epsilon_fragments
            if tuple(self._voxel_roi()[0]) != (0, 0, 0):
                raise NotImplementedError("epsilon_agglomerate with a ROI offset is not implemented")
            nid, npos, nsz = epsilon_fragments(plan, self._load_affs(), self.params, self._frags(), mask=self._mask_dev())
            self.frags_data.array("r+").write(self._frags().cpu().numpy().view(np.uint64))
            if nid.numel():
                a = self._affs_array()
                world = npos.cpu().numpy().astype(np.int64) * np.array(self.voxel_size) + np.array(a.offset)
                self.db.write_nodes(nid.cpu().numpy().view(np.uint64), world, nsz.cpu().numpy().astype(np.int64))
            return
        from ..pipeline import fragments_all_blocks
        fragments_all_blocks(plan, self._load_affs(), self.params, frags_out=self._frags(), mask=self._mask_dev())
        self._store(plan)

    @contextmanager
    def process_block_func(self):
        plan = self._plan()
        affs, mask = self._load_affs(), self._mask_dev()

        def process_block(block):
            if self.params["epsilon_agglomerate"] and self.params["epsilon_agglomerate"] > 0:
                raise NotImplementedError("epsilon_agglomerate runs through run_all() (all blocks in one batch), not block by block")
            plan.set_owned([block.plan_index])
            plan.fragments(affs, frags_out=self._frags(), mask=mask)
            self._store(plan, [block.plan_index])

        yield process_block
