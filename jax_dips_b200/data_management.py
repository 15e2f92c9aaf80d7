"""Batching and the multi-resolution schedule (jax_dips/data/data_management.py:79-186, 189-248,
319-326).

The reference materialises `(n_gpus, n_batches, batch, 3)` point arrays.  The CUDA path works from
the grid's 1-D coordinate arrays, so a batch is only a contiguous RANGE of the z-fastest flattened
point list; `DatasetDict` here computes those ranges with the reference's arithmetic.

Ragged sizes (the reference's own LPBE example: 32^3 points in batches of 3996, lpbe.yaml:76-84): the
reference fills the short last batch of every device with `batch_size - last` RANDOM points drawn from
jax PRNGKey(0) (:70-76, :135-150), which cannot be reproduced without jax.  Here the short batch is
trained on its real points only (its loss is the mean over those points); a device whose block ends
early gets empty ranges so that every device still takes the same number of steps.  `padded` tells.
"""
from __future__ import annotations

import math
from typing import List, Tuple


class DatasetDict:
    def __init__(self, x_data=None, batch_size: int = 131072, num_gpus: int = 1, num_points: int = None):
        self._len = int(num_points if num_points is not None else len(x_data))
        self.batch_size = int(batch_size)
        self.num_gpus = int(num_gpus)
        self._len_per_gpu = int(math.ceil(self._len / self.num_gpus))          # :87
        if self.batch_size > self._len_per_gpu:                                  # :90-91
            self.batch_size = self._len_per_gpu
        self.last_batch_size = self._len_per_gpu % self.batch_size               # :109
        self.extra_batch_per_gpu = 1 if self.last_batch_size != 0 else 0         # :93
        self.num_batches_per_gpu = self._len_per_gpu // self.batch_size          # :94
        # where the reference pads with random points
        self.padded = bool(self.extra_batch_per_gpu or self._len % self.num_gpus)

    @property
    def num_batches(self) -> int:
        return self.num_batches_per_gpu + self.extra_batch_per_gpu

    def batch_range(self, gpu: int, batch_id: int) -> Tuple[int, int]:
        """flattened-point range of batch `batch_id` on device `gpu` (:121-130, :152-165); may be short (the last
        batch) or empty (a device whose block ends before this batch)"""
        begin = gpu * self._len_per_gpu + batch_id * self.batch_size
        size = self.batch_size if batch_id < self.num_batches_per_gpu else self.last_batch_size
        begin = min(begin, self._len)
        return begin, min(begin + size, self._len)

    def ranges(self, gpu: int = 0) -> List[Tuple[int, int]]:
        return [self.batch_range(gpu, b) for b in range(self.num_batches)]


class TrainData:
    """data_management.py:189-248: train points = grid nodes; sequential zoom schedule :319-326."""

    def __init__(self, gstate, lvl_set_fn=None, refine=False, refine_lod=False, refine_normals=False,
                 v_cycle_period: int = 4, rest_at_level: int = 10):
        if refine or refine_lod or refine_normals:
            raise NotImplementedError("point refinement (kaolin) is off on this path (trainer.py:159-161)")
        self.gstate = gstate
        self.lvl_set_fn = lvl_set_fn
        self.v_cycle_period, self.rest_at_level = v_cycle_period, rest_at_level
        self.alt_res = False
        self.base_level = gstate.base_level()

    @property
    def train_points(self):
        return self.gstate.R

    def alternate_res_sequentially(self, num_epochs: int, epoch: int, train_dx=None, train_dy=None, train_dz=None):
        """:319-326  cell size = grid spacing * 0.5**(epoch // (num_epochs // 4))"""
        if num_epochs // 4 == 0:
            raise ZeroDivisionError("num_epochs must be >= 4 on the single-GPU path (data_management.py:322)")
        zoom_lvl = epoch // (num_epochs // 4)
        return self.zoom_cell(zoom_lvl)

    def zoom_level(self, num_epochs: int, epoch: int) -> int:
        if num_epochs // 4 == 0:
            raise ZeroDivisionError("num_epochs must be >= 4 on the single-GPU path (data_management.py:322)")
        return epoch // (num_epochs // 4)

    def zoom_cell(self, zoom_lvl: int):
        import torch
        f = torch.tensor(0.5 ** zoom_lvl, dtype=torch.float32)
        return (float(self.gstate.dx * f), float(self.gstate.dy * f), float(self.gstate.dz * f))
