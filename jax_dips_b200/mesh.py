"""Uniform grid state (mirrors jax_dips/domain/mesh.py:39-208).

`construct(3)` returns `(init_mesh_fn, coord_at)` exactly like the reference; `init_mesh_fn(x, y, z)`
builds a `GridState` whose flattened point list `R` is z-fastest
(`column_stack(meshgrid(x, y, z, indexing="ij").flatten())`, mesh.py:121-153).  Arrays are torch
tensors (float32); `R` and the six boundary-face lists are materialised lazily because a 512^3 `R`
is 1.6 GB and the CUDA path never needs it (the kernels work from the 1-D coordinate arrays).
"""
from __future__ import annotations

import math
from typing import Callable, Tuple

import numpy as np
import torch


def _as_f32(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.detach().to(torch.float32).contiguous()
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous()


class GridState:
    """A struct containing the state of the grid (mesh.py:39-89)."""

    def __init__(self, x, y, z):
        self.x, self.y, self.z = _as_f32(x), _as_f32(y), _as_f32(z)
        self.dx = self.x[1] - self.x[0]
        self.dy = self.y[1] - self.y[0]
        self.dz = self.z[1] - self.z[0]
        self._R = None

    # -- lazily built point lists ---------------------------------------------------------
    @property
    def R(self) -> torch.Tensor:
        if self._R is None:
            X, Y, Z = torch.meshgrid(self.x, self.y, self.z, indexing="ij")
            self._R = torch.stack((X.reshape(-1), Y.reshape(-1), Z.reshape(-1)), dim=1)
        return self._R

    def _face(self, axis: int, idx: int) -> torch.Tensor:
        nx, ny, nz = self.shape()
        return self.R.reshape(nx, ny, nz, 3).select(axis, idx).reshape(-1, 3)

    @property
    def R_xmin_boundary(self): return self._face(0, 0)
    @property
    def R_xmax_boundary(self): return self._face(0, -1)
    @property
    def R_ymin_boundary(self): return self._face(1, 0)
    @property
    def R_ymax_boundary(self): return self._face(1, -1)
    @property
    def R_zmin_boundary(self): return self._face(2, 0)
    @property
    def R_zmax_boundary(self): return self._face(2, -1)

    # -- mesh.py:51-89 ---------------------------------------------------------------------
    def shape(self) -> Tuple[int, int, int]:
        return (self.x.shape[0], self.y.shape[0], self.z.shape[0])

    def xmin(self): return self.x.min()
    def xmax(self): return self.x.max()
    def ymin(self): return self.y.min()
    def ymax(self): return self.y.max()
    def zmin(self): return self.z.min()
    def zmax(self): return self.z.max()

    def base_level(self) -> int:
        n = self.x.shape[0] * self.y.shape[0] * self.z.shape[0]
        return int(math.log2(n ** (1.0 / 3.0)))

    def num_points(self) -> int:
        nx, ny, nz = self.shape()
        return nx * ny * nz


def construct(dimension: int):
    """mesh.py:92-176.  Only the 3-D mesher exists on this path."""
    if dimension != 3:
        raise NotImplementedError("the NBM Poisson path is three-dimensional")

    def init_fn_3d(x, y, z) -> GridState:
        return GridState(x, y, z)

    def point3d_at(gstate: GridState, idx):
        i, j, k = idx
        return [gstate.x[i], gstate.y[j], gstate.z[k]]

    return init_fn_3d, point3d_at


def linspace_grid(lo, hi, n) -> GridState:
    """The way the reference drivers build their grids (tests/test_poisson.py:112-125):
    float32 `linspace` per axis."""
    ax = [torch.linspace(float(lo[a]), float(hi[a]), int(n[a]), dtype=torch.float64).to(torch.float32)
          for a in range(3)]
    return GridState(*ax)
