"""Stand-ins for volara.datasets.Raw / Labels (a zarr store path + helpers) and daisy's Block, as far as
the ws tasks use them (post/blockwise/watershed_frags.py:39-113, waterz_agglom.py:49-104)."""
import os
import shutil
from dataclasses import dataclass

import numpy as np

from . import zarrio


@dataclass
class Dataset:
    store: str

    @property
    def name(self):
        return os.path.basename(os.path.normpath(self.store))

    def array(self, mode="r"):
        return zarrio.open_ds(self.store, mode)

    def drop(self):
        if os.path.exists(self.store):
            shutil.rmtree(self.store)

    def prepare(self, shape, chunk_shape, offset, voxel_size, units=None, axis_names=None, types=None,
                dtype=np.uint64):
        return zarrio.prepare_ds(self.store, shape, offset, voxel_size, dtype, chunk_shape=chunk_shape,
                                 axis_names=axis_names, units=units, types=types)


class Raw(Dataset):
    pass


class Labels(Dataset):
    pass


@dataclass
class Roi:
    offset: tuple
    shape: tuple

    @property
    def begin(self):
        return self.offset

    @property
    def end(self):
        return tuple(o + s for o, s in zip(self.offset, self.shape))


@dataclass
class Block:
    """daisy.Block as the tasks see it: world-unit ROIs and (task_name, cantor number)."""
    read_roi: Roi
    write_roi: Roi
    block_id: tuple
    plan_index: int = -1
