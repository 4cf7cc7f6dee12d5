"""WaterzAgglom — drop-in for the reference's blockwise RAG-scoring task
(post/blockwise/waterz_agglom.py:39-181); numerics in libbsnative (bs_stage2_agglomerate)."""
from contextlib import contextmanager

import numpy as np
import torch

from .base import BlockwiseTask

# only "mean" is enabled, as in the reference (waterz_agglom.py:24-36)
WATERZ_MERGE_FUNCTIONS = {"mean": "OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>"}


class WaterzAgglom(BlockwiseTask):
    task_type = "waterz-agglom"

    def __init__(self, db, affs_data, frags_data, block_size, context, num_workers=1, roi=None,
                 merge_function=WATERZ_MERGE_FUNCTIONS["mean"]):
        if merge_function != WATERZ_MERGE_FUNCTIONS["mean"]:
            raise NotImplementedError(f"scoring function {merge_function!r} is not enabled (reference: mean only)")
        self.db, self.affs_data, self.frags_data = db, affs_data, frags_data
        self.block_size, self.context = tuple(int(v) for v in block_size), tuple(int(v) for v in context)
        self.num_workers, self.roi, self.merge_function = num_workers, roi, merge_function
        self._plan_obj = None
        self._frags_dev = None

    @property
    def task_name(self):
        return f"{self.db.id}-{self.task_type}"

    @property
    def output_datasets(self):
        return []

    def drop_artifacts(self):
        self.db.drop_edges()

    def init(self):
        self.db.init()

    def _plan(self):
        if self._plan_obj is None:
            self._plan_obj = self._make_plan()
        return self._plan_obj

    def _frags(self):
        if self._frags_dev is None:
            f = self.frags_data.array("r").read()
            self._frags_dev = torch.from_numpy(np.ascontiguousarray(f).view(np.int64)).cuda()
        return self._frags_dev

    def _install_counts(self, plan):
        """fragment ids are 1..n per block (+ block_id * prod(block_size)): n = max local id in the block"""
        frags = self._frags()
        ids, wo, ws = plan.block_info()
        off, _ = self._voxel_roi()
        nvox = self.num_voxels_in_block
        counts = np.zeros(len(ids), np.int64)
        for i in range(len(ids)):
            lo = [int(o) - r for o, r in zip(wo[i], off)]
            sl = tuple(slice(l, l + int(s)) for l, s in zip(lo, ws[i]))
            m = int(frags[sl].max().item())
            counts[i] = m - int(ids[i]) * nvox if m > 0 else 0
        plan.set_block_counts(counts)

    def _store(self, plan):
        u, v, s = [t.cpu().numpy() for t in plan.edges("cuda")]
        if len(u):
            self.db.write_edges(u.view(np.uint64), v.view(np.uint64), s)

    def run_all(self):
        plan = self._plan()
        n, _ = plan.num_blocks()
        self._install_counts(plan)
        plan.set_owned(np.arange(n))
        plan.agglomerate(self._load_affs(), self._frags())
        self._store(plan)

    @contextmanager
    def process_block_func(self):
        plan = self._plan()
        self._install_counts(plan)
        affs, frags = self._load_affs(), self._frags()

        def process_block(block):
            plan.set_owned([block.plan_index])
            plan.agglomerate(affs, frags)
            self._store(plan)

        yield process_block
