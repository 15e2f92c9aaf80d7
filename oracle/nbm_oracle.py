"""CPU ORACLE (test infrastructure, never shipped, never measured as the product).

A literal, vectorised torch-CPU restatement of the JAX-DIPS neural-bootstrapping (NBM)
training step for the interfacial Poisson problem.  Every function cites the reference
file:line it follows (paths relative to /root/reference).  It deliberately keeps the
reference's redundant formulation (197 network evaluations per point, both branches of every
`where` evaluated) so that it states WHAT the reference computes, not how the CUDA product
computes it.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product (`jax_dips_b200/`) never does.

Pinning status
--------------
* geometry: pinned by the reference's own known-answer test (sphere r=0.5 in [-2,2]^3 at
  128^3: area = pi +- 0.02, volume = pi/6 +- 0.02; tests/test_geometric_integrations.py:181-182)
  -> tests/test_oracle_pinning.py::test_reference_kat_sphere_area_and_volume
* geometry / regression / residual rows: pinned against the reference's OWN source files
  executed in the build container through a minimal numpy stand-in for the jax API
  (oracle/jax_shim + oracle/make_golden.py -> tests/golden/*.npz).
* loss / gradient: the reference holds no golden vector and `jax.value_and_grad` cannot run without jax, but the
  loss itself and its central differences along 11 parameter directions ARE computed by the reference's own source
  files in x64 mode (oracle/make_golden_grad.py -> tests/golden/grad_*.npz); the oracle's loss agrees to 8e-8 and
  its autograd gradient (projected on those directions) to 4e-7.  Pinned.
* optimizer chain: follows optax 0.1.5's published semantics; learned preconditioner: restated flax
  `nn.Dense`/tanh/sigmoid.  "parity unpinned" for those two items (neither optax nor flax is installed).

dtype: every function works in the dtype of its inputs.  float32 mirrors what the reference
prints (`jax_enable_x64 = False`); float64 gives the exact-arithmetic value of the same
algorithm (level-set samples are always taken through `phi_fn`, which the caller keeps in
float32 so that all sign decisions are identical in both modes).
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# A.1 sign helpers  (geometry/geometric_integrations_per_point.py:32-49, discretization.py:239-247)
# ----------------------------------------------------------------------------------------
def sign_pm_fn(a: Tensor) -> Tensor:
    return torch.sign(torch.sign(a) - 0.5)


def sign_p_fn(a: Tensor) -> Tensor:
    return torch.floor(0.5 * torch.sign(a) + 0.75)


def sign_m_fn(a: Tensor) -> Tensor:
    return torch.ceil(0.5 * torch.sign(a) - 0.75) * (-1.0)


def perturb_level_set_fn(phi_fn: Callable[[Tensor], Tensor]) -> Callable[[Tensor], Tensor]:
    """geometry/level_set.py:34-48"""
    EPS = 1.0e-10

    def perturbed(R: Tensor) -> Tensor:
        lvl = phi_fn(R)
        return lvl + sign_pm_fn(lvl) * EPS

    return perturbed


def nan_to_num(x: Tensor) -> Tensor:
    """jnp.nan_to_num: nan -> 0, +-inf -> +-finfo.max"""
    fi = torch.finfo(x.dtype)
    return torch.nan_to_num(x, nan=0.0, posinf=fi.max, neginf=fi.min)


# ----------------------------------------------------------------------------------------
# a2: uniform grid, z fastest   (domain/mesh.py:121-153)
# ----------------------------------------------------------------------------------------
class OracleGrid:
    def __init__(self, x: Tensor, y: Tensor, z: Tensor):
        self.x, self.y, self.z = x, y, z
        self.dx = x[1] - x[0]
        self.dy = y[1] - y[0]
        self.dz = z[1] - z[0]
        X, Y, Z = torch.meshgrid(x, y, z, indexing="ij")
        self.R = torch.stack((X.reshape(-1), Y.reshape(-1), Z.reshape(-1)), dim=1)

    def shape(self):
        return (self.x.shape[0], self.y.shape[0], self.z.shape[0])

    def xmin(self): return self.x.min()
    def xmax(self): return self.x.max()
    def ymin(self): return self.y.min()
    def ymax(self): return self.y.max()
    def zmin(self): return self.z.min()
    def zmax(self): return self.z.max()


def make_grid(lo: Sequence[float], hi: Sequence[float], n: Sequence[int], dtype=torch.float32) -> OracleGrid:
    """linspace grid exactly as the reference's drivers build it (tests/test_poisson.py:112-125):
    jnp.linspace in float32."""
    ax = [torch.linspace(lo[a], hi[a], n[a], dtype=torch.float64).to(dtype) for a in range(3)]
    return OracleGrid(*ax)


# ----------------------------------------------------------------------------------------
# a14: level-set interpolants on the lvl grid   (domain/interpolate.py)
# ----------------------------------------------------------------------------------------
def add_ghost_layer_3d(x: Tensor, y: Tensor, z: Tensor, c: Tensor):
    """interpolate.py:762-816 : one linearly extrapolated layer, x then y then z."""
    nx, ny, nz = c.shape
    g = torch.zeros((nx + 2, ny + 2, nz + 2), dtype=c.dtype)
    g[1:-1, 1:-1, 1:-1] = c

    def ext(a):
        out = torch.zeros(a.shape[0] + 2, dtype=a.dtype)
        out[1:-1] = a
        out[0] = a[0] - (a[1] - a[0])
        out[-1] = a[-1] + (a[-1] - a[-2])
        return out

    xx, yy, zz = ext(x), ext(y), ext(z)
    g[0, 1:-1, 1:-1] = 2 * c[0] - c[1]
    g[-1, 1:-1, 1:-1] = 2 * c[-1] - c[-2]
    g[:, 0, :] = 2 * g[:, 1, :] - g[:, 2, :]
    g[:, -1, :] = 2 * g[:, -2, :] - g[:, -3, :]
    g[:, :, 0] = 2 * g[:, :, 1] - g[:, :, 2]
    g[:, :, -1] = 2 * g[:, :, -2] - g[:, :, -3]
    return xx, yy, zz, g


def _cell_index(p: Tensor, a: Tensor, d: Tensor) -> Tensor:
    """interpolate.py:946-956 : trunc((p - a0)/d), clamp high to n-2, then `<=1 -> 2`."""
    i = ((p - a[0]) / d).to(torch.int32).to(torch.int64)  # astype(int32) truncates toward zero
    n = a.shape[0]
    i = torch.where(i >= n - 1, torch.full_like(i, n - 2), i)
    i = torch.where(i <= 1, torch.full_like(i, 2), i)
    return i


def multilinear_interpolation(c: Tensor, grid: OracleGrid) -> Callable[[Tensor], Tensor]:
    """interpolate.py:906-1021 (trilinear on the ghosted grid)."""
    nx, ny, nz = grid.shape()
    x, y, z, cube = add_ghost_layer_3d(grid.x, grid.y, grid.z, c.reshape(nx, ny, nz))
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]

    def interp(R: Tensor) -> Tensor:
        R = R.to(cube.dtype)
        xp, yp, zp = R[:, 0], R[:, 1], R[:, 2]
        i, j, k = _cell_index(xp, x, dx), _cell_index(yp, y, dy), _cell_index(zp, z, dz)
        c000 = cube[i, j, k]; c100 = cube[i + 1, j, k]
        c010 = cube[i, j + 1, k]; c110 = cube[i + 1, j + 1, k]
        c001 = cube[i, j, k + 1]; c101 = cube[i + 1, j, k + 1]
        c011 = cube[i, j + 1, k + 1]; c111 = cube[i + 1, j + 1, k + 1]
        xd = (xp - x[i]) / (x[i + 1] - x[i])
        yd = (yp - y[j]) / (y[j + 1] - y[j])
        zd = (zp - z[k]) / (z[k + 1] - z[k])
        c00 = c000 * (1.0 - xd) + c100 * xd
        c01 = c001 * (1.0 - xd) + c101 * xd
        c10 = c010 * (1.0 - xd) + c110 * xd
        c11 = c011 * (1.0 - xd) + c111 * xd
        c0 = c00 * (1.0 - yd) + c10 * yd
        c1 = c01 * (1.0 - yd) + c11 * yd
        return c0 * (1.0 - zd) + c1 * zd

    return interp


def nonoscillatory_quadratic_interpolation(c: Tensor, grid: OracleGrid) -> Callable[[Tensor], Tensor]:
    """interpolate.py:388-569 (per-point variant used by examples/dragon/solve_dragon.py:177).
    Out-of-range gathers (i+2 at the last cell) clamp to the edge like XLA's gather."""
    nx, ny, nz = grid.shape()
    x, y, z, cube = add_ghost_layer_3d(grid.x, grid.y, grid.z, c.reshape(nx, ny, nz))
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]
    GX, GY, GZ = cube.shape

    def at(i, j, k):
        return cube[i.clamp(0, GX - 1), j.clamp(0, GY - 1), k.clamp(0, GZ - 1)]

    def interp(R: Tensor) -> Tensor:
        R = R.to(cube.dtype)
        xp, yp, zp = R[:, 0], R[:, 1], R[:, 2]
        i, j, k = _cell_index(xp, x, dx), _cell_index(yp, y, dy), _cell_index(zp, z, dz)
        xd = (xp - x[i]) / (x[i + 1] - x[i])
        yd = (yp - y[j]) / (y[j + 1] - y[j])
        zd = (zp - z[k]) / (z[k + 1] - z[k])
        c000 = at(i, j, k); c100 = at(i + 1, j, k)
        c010 = at(i, j + 1, k); c110 = at(i + 1, j + 1, k)
        c001 = at(i, j, k + 1); c101 = at(i + 1, j, k + 1)
        c011 = at(i, j + 1, k + 1); c111 = at(i + 1, j + 1, k + 1)
        c00 = c000 * (1.0 - xd) + c100 * xd
        c01 = c001 * (1.0 - xd) + c101 * xd
        c10 = c010 * (1.0 - xd) + c110 * xd
        c11 = c011 * (1.0 - xd) + c111 * xd
        c0 = c00 * (1.0 - yd) + c10 * yd
        c1 = c01 * (1.0 - yd) + c11 * yd
        val = c0 * (1.0 - zd) + c1 * zd
        mins = []
        for axis in range(3):
            cand = []
            for (a, b, cc) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 0), (1, 1, 1)):
                ii, jj, kk = i + a, j + b, k + cc
                e = [0, 0, 0]; e[axis] = 1
                d2 = at(ii + e[0], jj + e[1], kk + e[2]) - 2 * at(ii, jj, kk) + at(ii - e[0], jj - e[1], kk - e[2])
                cand.append(d2.abs())
            mins.append(torch.stack(cand, 0).min(dim=0).values)
        val = (val - mins[0] * 0.5 * xd * (1.0 - xd) - mins[1] * 0.5 * yd * (1.0 - yd)
               - mins[2] * 0.5 * zd * (1.0 - zd))
        return val

    return interp


# ----------------------------------------------------------------------------------------
# a12/a13: cut-cell pieces    (geometry/geometric_integrations_per_point.py:53-367)
# ----------------------------------------------------------------------------------------
_CORNER_SIGNS = [  # :212-224, corner order 000,100,101,001,010,110,011,111
    [-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1],
    [-1, 1, -1], [1, 1, -1], [-1, 1, 1], [1, 1, 1],
]
# :312-316  S1..S5 as indices into the corner list above
_TETS = [[0, 1, 4, 3], [5, 1, 4, 7], [2, 1, 7, 3], [6, 7, 4, 3], [7, 1, 4, 3]]


def cell_corners(points: Tensor, dx, dy, dz) -> Tensor:
    d = torch.stack([torch.as_tensor(v, dtype=points.dtype) for v in (dx, dy, dz)])
    dX = torch.tensor(_CORNER_SIGNS, dtype=points.dtype) * d * 0.5  # (8,3)
    return dX[None, :, :] + points[:, None, :]


def is_cell_crossed(points: Tensor, dx, dy, dz, phi_fn) -> Tensor:
    """:203-263  -> -1 / 0 / +1 (float, like jnp.sign)."""
    n = points.shape[0]
    corners = cell_corners(points, dx, dy, dz)
    phis = phi_fn(corners.reshape(-1, 3)).reshape(n, 8).to(points.dtype)
    eta = sign_m_fn(phis).sum(dim=1).to(torch.int64)
    return torch.where(eta * (eta - 8) == 0, torch.sign(phis[:, 0]), torch.zeros_like(phis[:, 0]))


def _cut(phi_s: Tensor, S_s: Tensor, a: int, b: int) -> Tensor:
    """(phi_a S_b - phi_b S_a)/(phi_a - phi_b) on sorted arrays (:70 etc.)."""
    pa, pb = phi_s[..., a, None], phi_s[..., b, None]
    return (pa * S_s[..., b, :] - pb * S_s[..., a, :]) / (pa - pb)


def _sort_tet(S: Tensor, phi: Tensor):
    idx = torch.argsort(phi, dim=-1, stable=True)
    phi_s = torch.gather(phi, -1, idx)
    S_s = torch.gather(S, -2, idx[..., None].expand(*idx.shape, 3))
    return S_s, phi_s


def intersect_gamma(S: Tensor, phi: Tensor, eta: Tensor) -> Tensor:
    """get_vertices_S_intersect_Gamma :53-107.  S (...,4,3), phi (...,4), eta (...) -> (...,2,3,3)."""
    zeros_tri = torch.zeros(S.shape[:-2] + (3, 3), dtype=S.dtype)

    def eta1(S, phi):
        Ss, ps = _sort_tet(S, phi)
        tri0 = torch.stack((_cut(ps, Ss, 0, 1), _cut(ps, Ss, 0, 2), _cut(ps, Ss, 0, 3)), dim=-2)
        return torch.stack((tri0, zeros_tri), dim=-3)

    def eta2(S, phi):
        Ss, ps = _sort_tet(S, phi)
        Q0, Q1, Q2, Q5 = _cut(ps, Ss, 0, 2), _cut(ps, Ss, 0, 3), _cut(ps, Ss, 1, 3), _cut(ps, Ss, 1, 2)
        return torch.stack((torch.stack((Q0, Q1, Q2), dim=-2), torch.stack((Q0, Q5, Q2), dim=-2)), dim=-3)

    e = eta[..., None, None, None]
    r1, r2, r3 = eta1(S, phi), eta2(S, phi), eta1(S, -1.0 * phi)
    out23 = torch.where(e == 2, r2, r3)
    out123 = torch.where(e == 1, r1, out23)
    return torch.where(e * (e - 4) == 0, torch.zeros_like(r1), out123)


def intersect_omega_m(S: Tensor, phi: Tensor, eta: Tensor) -> Tensor:
    """get_vertices_S_intersect_Omega_m :111-198 -> (...,3,4,3)."""
    Ss, ps = _sort_tet(S, phi)
    Z = torch.zeros_like(S)

    # eta = 1
    r1 = torch.stack((torch.stack((Ss[..., 0, :], _cut(ps, Ss, 0, 1), _cut(ps, Ss, 0, 2), _cut(ps, Ss, 0, 3)), -2), Z, Z), -3)
    # eta = 2
    Q0, Q1 = Ss[..., 0, :], Ss[..., 1, :]
    Q2, Q3, Q4, Q5 = _cut(ps, Ss, 0, 2), _cut(ps, Ss, 1, 3), _cut(ps, Ss, 1, 2), _cut(ps, Ss, 0, 3)
    r2 = torch.stack((torch.stack((Q0, Q1, Q2, Q3), -2), torch.stack((Q4, Q1, Q2, Q3), -2),
                      torch.stack((Q0, Q5, Q2, Q3), -2)), -3)
    # eta = 3
    Q0, Q1, Q2 = Ss[..., 0, :], Ss[..., 1, :], Ss[..., 2, :]
    Q3, Q4, Q5 = _cut(ps, Ss, 1, 3), _cut(ps, Ss, 0, 3), _cut(ps, Ss, 2, 3)
    r3 = torch.stack((torch.stack((Q0, Q1, Q2, Q3), -2), torch.stack((Q0, Q4, Q2, Q3), -2),
                      torch.stack((Q5, Q4, Q2, Q3), -2)), -3)
    # eta = 4: the simplex itself (unsorted), eta = 0: nothing
    r4 = torch.stack((S, Z, Z), -3)
    e = eta[..., None, None, None]
    out23 = torch.where(e == 2, r2, r3)
    out123 = torch.where(e == 1, r1, out23)
    out04 = torch.where(e == 0, torch.zeros_like(r4), r4)
    return torch.where(e * (e - 4) == 0, out04, out123)


def cell_pieces(points: Tensor, dx, dy, dz, phi_fn):
    """get_vertices_of_cell_intersection_with_interface_at_point_ :265-362.
    Returns gamma (n,5,2,3,3), omega (n,5,3,4,3), S (n,5,4,3)."""
    n = points.shape[0]
    corners = cell_corners(points, dx, dy, dz)
    phis = phi_fn(corners.reshape(-1, 3)).reshape(n, 8).to(points.dtype)
    tets = torch.tensor(_TETS, dtype=torch.int64)
    S = corners[:, tets, :]          # (n,5,4,3)
    phi_S = phis[:, tets]            # (n,5,4)
    eta = sign_m_fn(phi_S).sum(dim=-1).to(torch.int64)
    return intersect_gamma(S, phi_S, eta), intersect_omega_m(S, phi_S, eta), S


def _det2(G):
    return G[..., 0, 0] * G[..., 1, 1] - G[..., 0, 1] * G[..., 1, 0]


def _det3(G):
    # closed form used by jnp.linalg.det for 3x3 (jax/_src/numpy/linalg.py, jax 0.4.13)
    return (G[..., 0, 0] * G[..., 1, 1] * G[..., 2, 2] + G[..., 0, 1] * G[..., 1, 2] * G[..., 2, 0]
            + G[..., 0, 2] * G[..., 1, 0] * G[..., 2, 1] - G[..., 0, 2] * G[..., 1, 1] * G[..., 2, 0]
            - G[..., 0, 0] * G[..., 1, 2] * G[..., 2, 1] - G[..., 0, 1] * G[..., 1, 0] * G[..., 2, 2])


def vol_fn(A: Tensor) -> Tensor:
    """:370-374  A (...,4,3)"""
    E = A[..., 1:, :] - A[..., :1, :]
    G = E @ E.transpose(-1, -2)
    return (1.0 / 6.0) * torch.sqrt(torch.abs(nan_to_num(_det3(G))))


def area_fn(A: Tensor) -> Tensor:
    """:377-380  A (...,3,3)"""
    E = A[..., 1:, :] - A[..., :1, :]
    G = E @ E.transpose(-1, -2)
    return 0.5 * torch.sqrt(torch.abs(nan_to_num(_det2(G))))


# ----------------------------------------------------------------------------------------
# a11: interface integral   (:383-421)
# ----------------------------------------------------------------------------------------
def integrate_over_interface(points: Tensor, dx, dy, dz, phi_fn, u_fn) -> Tensor:
    n = points.shape[0]
    gamma, _, _ = cell_pieces(points, dx, dy, dz, phi_fn)            # (n,5,2,3,3)
    areas = area_fn(gamma)                                           # (n,5,2)
    vals = u_fn(gamma.reshape(-1, 3)).reshape(n, 5, 2, 3).to(points.dtype).mean(dim=-1)
    integral = torch.zeros(n, dtype=points.dtype)
    for t in range(5):
        for j in range(2):
            integral = integral + areas[:, t, j] * vals[:, t, j]
    flag = is_cell_crossed(points, dx, dy, dz, phi_fn)
    return torch.where(flag == 0, integral, torch.zeros_like(integral))


# ----------------------------------------------------------------------------------------
# a10: face coefficients / volumes   (:464-1009)
# ----------------------------------------------------------------------------------------
_FACE_TETS = {  # :841-852 (indices into S1..S5)
    "xm": (0, 3), "xp": (1, 2), "ym": (0, 2), "yp": (1, 3), "zm": (0, 1), "zp": (2, 3),
}
_FIDUCIAL = 200.0  # :582


def _face_area_minus(omega_a: Tensor, omega_b: Tensor, axis: int, face: Tensor, atol: Tensor) -> Tensor:
    """extract_area_minus_*_face + compute_area_from_partitions (:584-839)."""
    total = torch.zeros(omega_a.shape[0], dtype=omega_a.dtype)
    for part in (omega_a, omega_b):                                   # (n,3,4,3)
        coord = part[..., axis]
        on_face = (coord - face[:, None, None]).abs() <= (atol + 1e-5 * face.abs())[:, None, None]  # jnp.isclose
        masked = torch.where(on_face[..., None], part, torch.full_like(part, _FIDUCIAL))
        cnt = (masked[..., 0] - _FIDUCIAL != 0).sum(dim=-1)          # count_nonzero(simplex[:,0]-fiducial)
        idx = torch.argsort(masked[..., 0], dim=-1, stable=True)     # argsort on the x column (always column 0)
        srt = torch.gather(masked, -2, idx[..., None].expand(*idx.shape, 3))[..., :3, :]
        a = area_fn(srt)
        a = torch.where(cnt == 3, a, torch.zeros_like(a))
        total = total + (a[:, 0] + a[:, 1] + a[:, 2])
    return total


def cell_faces_areas_values(points: Tensor, dx, dy, dz, phi_fn, mu_m_fn, mu_p_fn) -> Tensor:
    """compute_face_centroids_values_plus_minus_at_point (:998-1007) -> (n,26)."""
    dt = points.dtype
    n = points.shape[0]
    dx, dy, dz = (torch.as_tensor(v, dtype=dt) for v in (dx, dy, dz))
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    vol = dx * dy * dz
    area_x, area_y, area_z = dy * dz, dx * dz, dx * dy
    dXf = 0.5 * torch.stack((torch.stack((-dx, 0 * dx, 0 * dx)), torch.stack((dx, 0 * dx, 0 * dx)),
                             torch.stack((0 * dy, -dy, 0 * dy)), torch.stack((0 * dy, dy, 0 * dy)),
                             torch.stack((0 * dz, 0 * dz, -dz)), torch.stack((0 * dz, 0 * dz, dz))))
    Rf = dXf[None] + points[:, None, :]                               # (n,6,3)
    mu_m_faces = mu_m_fn(Rf.reshape(-1, 3)).reshape(n, 6).to(dt)
    mu_p_faces = mu_p_fn(Rf.reshape(-1, 3)).reshape(n, 6).to(dt)
    flag = is_cell_crossed(points, dx, dy, dz, phi_fn)

    # ---- crossed branch (compute_interface_faces :514-904)
    _, omega, _ = cell_pieces(points, dx, dy, dz, phi_fn)             # (n,5,3,4,3)
    vols = vol_fn(omega)                                              # (n,5,3)
    vol_m = torch.zeros(n, dtype=dt)
    for t in range(5):
        for j in range(3):
            vol_m = vol_m + vols[:, t, j]
    pos = lambda v: torch.where(v < 0, torch.zeros_like(v), v)
    vol_p = pos(vol - vol_m)
    faces = {"xm": (0, x - 0.5 * dx, dx), "xp": (0, x + 0.5 * dx, dx), "ym": (1, y - 0.5 * dy, dy),
             "yp": (1, y + 0.5 * dy, dy), "zm": (2, z - 0.5 * dz, dz), "zp": (2, z + 0.5 * dz, dz)}
    am = {}
    for name, (axis, fc, dd) in faces.items():
        ta, tb = _FACE_TETS[name]
        am[name] = _face_area_minus(omega[:, ta], omega[:, tb], axis, fc, 1e-10 * dd)
    nominal = {"xm": area_x, "xp": area_x, "ym": area_y, "yp": area_y, "zm": area_z, "zp": area_z}
    ap = {k: pos(nominal[k] - am[k]) for k in am}
    order = ["xm", "xp", "ym", "yp", "zm", "zp"]
    dd = {"xm": dx, "xp": dx, "ym": dy, "yp": dy, "zm": dz, "zp": dz}
    cols_i: List[Tensor] = []
    for f, name in enumerate(order):
        cols_i.append(am[name] * mu_m_faces[:, f] / dd[name])
        cols_i.append(ap[name] * mu_p_faces[:, f] / dd[name])
    cols_i += [vol_m, vol_p]
    for name in order:
        cols_i += [am[name], ap[name]]
    crossed = torch.stack(cols_i, dim=1)

    # ---- uncrossed branch (compute_domain_faces :906-996)
    m_mask, p_mask = sign_m_fn(flag), sign_p_fn(flag)
    mm = mu_m_faces * m_mask[:, None]
    mp = mu_p_faces * p_mask[:, None]
    cols_d: List[Tensor] = []
    for f, name in enumerate(order):
        cols_d.append(nominal[name] * mm[:, f] / dd[name])
        cols_d.append(nominal[name] * mp[:, f] / dd[name])
    cols_d += [vol * m_mask, vol * p_mask]
    for name in order:
        cols_d += [nominal[name] * m_mask, nominal[name] * p_mask]
    domain = torch.stack(cols_d, dim=1)
    return torch.where((flag == 0)[:, None], crossed, domain)


# ----------------------------------------------------------------------------------------
# a15: DoubleMLP    (nn/mlp/MLP.py:93-139)
# ----------------------------------------------------------------------------------------
class NetShape:
    """model_dict['mlp'] : (hidden_layers_p, hidden_dim_p, hidden_layers_m, hidden_dim_m); tanh."""

    def __init__(self, layers_p=2, dim_p=10, layers_m=1, dim_m=1):
        self.Lp, self.Hp, self.Lm, self.Hm = layers_p, dim_p, layers_m, dim_m

    @staticmethod
    def _count(L, H):
        return 3 * H + H + (L - 1) * (H * H + H) + H + 1

    @property
    def n_p(self): return self._count(self.Lp, self.Hp)
    @property
    def n_m(self): return self._count(self.Lm, self.Hm)
    @property
    def n_params(self): return self.n_p + self.n_m


def _unpack(flat: Tensor, L: int, H: int, off: int):
    layers = []
    fan_in = 3
    for _ in range(L):
        W = flat[off: off + fan_in * H].reshape(fan_in, H); off += fan_in * H
        b = flat[off: off + H]; off += H
        layers.append((W, b)); fan_in = H
    W = flat[off: off + fan_in].reshape(fan_in, 1); off += fan_in
    b = flat[off: off + 1]; off += 1
    layers.append((W, b))
    return layers, off


def mlp_eval(flat: Tensor, L: int, H: int, off: int, R: Tensor) -> Tensor:
    """hk.Linear: y = x W + b, W (in,out); tanh on hidden layers (MLP.py:111-116,128-139)."""
    layers, _ = _unpack(flat, L, H, off)
    h = R
    for (W, b) in layers[:-1]:
        h = torch.tanh(h @ W + b)
    W, b = layers[-1]
    return (h @ W + b).reshape(-1)


def init_params(shape: NetShape, seed: int = 42, dtype=torch.float32) -> Tensor:
    """Flat parameter vector in the C-ABI order: p-net first, then m-net, each
    W(in,out) row-major then b.  Hidden layers TruncNormal(sigma=0.1) (MLP.py:65,70), output layer
    haiku default TruncNormal(1/sqrt(fan_in)), biases 0.  The reference draws from PRNGKey(42)
    (threefry, trainer.py:222) which cannot be reproduced without jax: init is an INPUT here."""
    g = torch.Generator().manual_seed(seed)

    def trunc(n, std):
        out = torch.empty(n, dtype=torch.float64)
        torch.nn.init.trunc_normal_(out, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return out * std

    parts = []
    for (L, H) in ((shape.Lp, shape.Hp), (shape.Lm, shape.Hm)):
        fan_in = 3
        for _ in range(L):
            parts += [trunc(fan_in * H, 0.1), torch.zeros(H, dtype=torch.float64)]
            fan_in = H
        parts += [trunc(fan_in, 1.0 / math.sqrt(fan_in)), torch.zeros(1, dtype=torch.float64)]
    return torch.cat(parts).to(dtype)


# ----------------------------------------------------------------------------------------
# learned preconditioner   (nn/preconditioner.py:10-35; trainer.py:229-243, 846-847;
# discretization.py:339, 418-419)
# ----------------------------------------------------------------------------------------
class PrecondShape:
    """model_dict['preconditioner']: tanh MLP 26 -> layer_widths -> 1, P = 0.5 + scaling_coeff * sigmoid(.)
    (flax nn.Dense: y = x @ kernel + bias, kernel (in, out))."""
    N_IN = 26   # length of coeffs_ (geometric_integrations_per_point.py:873-903)

    def __init__(self, layer_widths=(8, 4), scaling_coeff: float = 1.0):
        self.Ds, self.scale = tuple(int(d) for d in layer_widths), float(scaling_coeff)

    @property
    def n_params(self):
        n, fan_in = 0, self.N_IN
        for d in self.Ds:
            n += fan_in * d + d
            fan_in = d
        return n + fan_in + 1


def precond_eval(flat: Tensor, shape: PrecondShape, coeffs: Tensor) -> Tensor:
    """Preconditioner.__call__ (nn/preconditioner.py:22-35) on (n, 26) inputs; `flat` holds Dense_0.kernel
    (row-major (in,out)), Dense_0.bias, Dense_1.kernel, ... in that order."""
    off, fan_in, h = 0, shape.N_IN, coeffs.to(flat.dtype)
    for d in shape.Ds:
        W = flat[off: off + fan_in * d].reshape(fan_in, d); off += fan_in * d
        b = flat[off: off + d]; off += d
        h = torch.tanh(h @ W + b)
        fan_in = d
    W = flat[off: off + fan_in].reshape(fan_in, 1); off += fan_in
    b = flat[off: off + 1]
    return 0.5 + shape.scale * torch.sigmoid((h @ W + b).reshape(-1))


def init_precond_params(shape: PrecondShape, seed: int = 42, dtype=torch.float32) -> Tensor:
    """glorot_uniform kernels (nn/preconditioner.py:20), zero biases (flax default).  As for the solution
    network the reference's PRNGKey(42) stream cannot be reproduced without jax: init is an input."""
    g = torch.Generator().manual_seed(seed + 1000)
    parts, fan_in = [], shape.N_IN
    for d in list(shape.Ds) + [1]:
        lim = math.sqrt(6.0 / (fan_in + d))
        parts += [(torch.rand(fan_in * d, generator=g, dtype=torch.float64) * 2 - 1) * lim, torch.zeros(d, dtype=torch.float64)]
        fan_in = d
    return torch.cat(parts).to(dtype)


# ----------------------------------------------------------------------------------------
# problem bundle: batched callables (what trainer.setup builds by vmap, trainer.py:995-1005)
# ----------------------------------------------------------------------------------------
class OracleProblem:
    """Batched callables (n,3)->(n,), the lvl-grid box bounds, the network shape."""

    def __init__(self, phi_fn, mu_m_fn, mu_p_fn, k_m_fn, k_p_fn, f_m_fn, f_p_fn, alpha_fn, beta_fn,
                 dir_bc_fn, bounds, shape: NetShape = None, nonlinear_op_m=None, nonlinear_op_p=None,
                 precond: "PrecondShape" = None):
        self.phi_fn = phi_fn
        self.mu_m_fn, self.mu_p_fn = mu_m_fn, mu_p_fn
        self.k_m_fn, self.k_p_fn = k_m_fn, k_p_fn
        self.f_m_fn, self.f_p_fn = f_m_fn, f_p_fn
        self.alpha_fn, self.beta_fn = alpha_fn, beta_fn
        self.dir_bc_fn = dir_bc_fn
        self.bounds = bounds  # (xmin,xmax,ymin,ymax,zmin,zmax) from lvl_gstate (discretization.py:60-65)
        self.shape = shape or NetShape()
        self.nonlinear_op_m = nonlinear_op_m or (lambda u: 0.0 * u)   # trainer.py:1007-1017
        self.nonlinear_op_p = nonlinear_op_p or (lambda u: 0.0 * u)
        # with a preconditioner the flat parameter vector is [network | preconditioner]
        self.precond = precond

    # evaluate_solution_fn (trainer.py:836-844) / solution_at_point_fn (:849-854) / MLP.py:98
    def solution(self, params: Tensor, R: Tensor, phi: Tensor = None) -> Tensor:
        if phi is None:
            phi = self.phi_fn(R)
        s = self.shape
        up = mlp_eval(params, s.Lp, s.Hp, 0, R.to(params.dtype))
        um = mlp_eval(params, s.Lm, s.Hm, s.n_p, R.to(params.dtype))
        return torch.where(phi.to(params.dtype) >= 0, up, um)


# ----------------------------------------------------------------------------------------
# a9: regression coefficients   (solvers/poisson/discretization.py:164-296)
# ----------------------------------------------------------------------------------------
def get_Xijk(dx, dy, dz, dtype) -> Tensor:
    """:164-197 : 27 offsets, x fastest, index 13 = centre."""
    rows = []
    for c in (-1.0, 0.0, 1.0):
        for b in (-1.0, 0.0, 1.0):
            for a in (-1.0, 0.0, 1.0):
                rows.append([a, b, c])
    d = torch.stack([torch.as_tensor(v, dtype=dtype) for v in (dx, dy, dz)])
    return torch.tensor(rows, dtype=dtype) * d


def normal_point_fn(points: Tensor, dx, dy, dz, phi_fn) -> Tensor:
    """:199-218 central differences of phi at +-(dx,dy,dz), normalised."""
    dt = points.dtype
    comps = []
    for a, d in enumerate((dx, dy, dz)):
        e = torch.zeros(3, dtype=dt); e[a] = 1.0
        d = torch.as_tensor(d, dtype=dt)
        comps.append((phi_fn(points + e * d).to(dt) - phi_fn(points - e * d).to(dt)) / (2 * d))
    g = torch.stack(comps, dim=1)
    norm = torch.sqrt(g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1] + g[:, 2] * g[:, 2])
    return g / norm[:, None]


PINV_RCOND = 10.0 * 3 * 1.1920929e-07  # jnp.linalg.pinv default: 10*max(M,N)*eps(float32)


def _pinv_sym3(A: Tensor) -> Tensor:
    return torch.linalg.pinv(A.to(torch.float64), rtol=PINV_RCOND).to(A.dtype) if A.dtype == torch.float32 \
        else torch.linalg.pinv(A, rtol=PINV_RCOND)


def regression_coeffs(points: Tensor, dx, dy, dz, prob: OracleProblem, pinv_in_working_dtype: bool = True):
    """get_regression_coeffs_at_point :238-296, vectorised over sites."""
    dt = points.dtype
    n = points.shape[0]
    X = get_Xijk(dx, dy, dz, dt)                                      # (27,3)
    verts = points[:, None, :] + X[None]                              # (n,27,3)
    phi_v = prob.phi_fn(verts.reshape(-1, 3)).reshape(n, 27).to(dt)
    Wp, Wm = sign_p_fn(phi_v), sign_m_fn(phi_v)                       # (n,27)
    Ap = torch.einsum("qa,nq,qb->nab", X, Wp, X)
    Am = torch.einsum("qa,nq,qb->nab", X, Wm, X)
    if pinv_in_working_dtype:
        pin = lambda A: torch.linalg.pinv(A, rtol=PINV_RCOND)
    else:
        pin = _pinv_sym3
    Dp = nan_to_num(pin(Ap) @ (Wp[:, :, None] * X[None]).transpose(1, 2))     # (n,3,27)
    Dm = nan_to_num(pin(Am) @ (Wm[:, :, None] * X[None]).transpose(1, 2))
    normal = normal_point_fn(points, dx, dy, dz, prob.phi_fn)         # (n,3)
    phi_pt = prob.phi_fn(points).to(dt)
    Cm = torch.einsum("na,naq->nq", normal, Dm)
    Cp = torch.einsum("na,naq->nq", normal, Dp)
    mu_p = prob.mu_p_fn(points).to(dt) * torch.ones(n, dtype=dt)
    mu_m = prob.mu_m_fn(points).to(dt) * torch.ones(n, dtype=dt)
    zeta_p_pqm = (((mu_p - mu_m) / mu_m) * phi_pt)[:, None] * Cp
    zeta_m_pqm = (((mu_p - mu_m) / mu_p) * phi_pt)[:, None] * Cm
    zeta_p = (zeta_p_pqm.sum(dim=1) - zeta_p_pqm[:, 13]) * (-1.0)
    zeta_m = (zeta_m_pqm.sum(dim=1) - zeta_m_pqm[:, 13]) * (-1.0)
    gamma_p_pqm = zeta_p_pqm / (1.0 + zeta_p[:, None])
    gamma_m_pqm = zeta_m_pqm / (1.0 - zeta_m[:, None])
    gamma_p = (gamma_p_pqm.sum(dim=1) - gamma_p_pqm[:, 13]) * (-1.0)
    gamma_m = (gamma_m_pqm.sum(dim=1) - gamma_m_pqm[:, 13]) * (-1.0)
    return dict(normal=normal, gamma_m=gamma_m, gamma_m_pqm=gamma_m_pqm, gamma_p=gamma_p,
                gamma_p_pqm=gamma_p_pqm, zeta_m=zeta_m, zeta_m_pqm=zeta_m_pqm, zeta_p=zeta_p,
                zeta_p_pqm=zeta_p_pqm, X=X, Dp=Dp, Dm=Dm, phi_v=phi_v)


# ----------------------------------------------------------------------------------------
# a8: u^-/u^+ at a stencil point   (discretization.py:426-519)
# ----------------------------------------------------------------------------------------
def u_mp_at_sites(params: Tensor, sites: Tensor, dx, dy, dz, prob: OracleProblem) -> Tuple[Tensor, Tensor]:
    dt = params.dtype
    n = sites.shape[0]
    sites = sites.to(dt)
    delta = prob.phi_fn(sites).to(dt)
    u = prob.solution(params, sites, delta)
    with torch.no_grad():
        rc = regression_coeffs(sites, dx, dy, dz, prob)
        flag = is_cell_crossed(sites, dx, dy, dz, prob.phi_fn)
        r_proj = sites - delta[:, None] * rc["normal"]
        mu_m = prob.mu_m_fn(sites).to(dt) * torch.ones(n, dtype=dt)
        mu_p = prob.mu_p_fn(sites).to(dt) * torch.ones(n, dtype=dt)
        alpha = prob.alpha_fn(r_proj).to(dt) * torch.ones(n, dtype=dt)
        beta = prob.beta_fn(r_proj).to(dt) * torch.ones(n, dtype=dt)
        b_over_mu_p = beta / (prob.mu_p_fn(r_proj).to(dt) * torch.ones(n, dtype=dt))
        b_over_mu_m = beta / (prob.mu_m_fn(r_proj).to(dt) * torch.ones(n, dtype=dt))
    verts = sites[:, None, :] + rc["X"][None]
    u_cube = prob.solution(params, verts.reshape(-1, 3)).reshape(n, 27)

    def dot(w):
        return (w * u_cube).sum(dim=1)

    # mu_minus_bigger_fn :465-486
    um_a = (-1.0 * dot(rc["gamma_m_pqm"]) + (1.0 - rc["gamma_m"] + rc["gamma_m_pqm"][:, 13]) * u
            + (-1.0) * (1.0 - rc["gamma_m"]) * (alpha + delta * b_over_mu_p))
    up_a = (-1.0 * dot(rc["zeta_m_pqm"]) + (1.0 - rc["zeta_m"] + rc["zeta_m_pqm"][:, 13]) * u
            + alpha + delta * b_over_mu_p)
    um_A = torch.where(delta > 0, um_a, u)
    up_A = torch.where(delta > 0, u, up_a)
    # mu_plus_bigger_fn :488-509
    um_b = (-1.0 * dot(rc["zeta_p_pqm"]) + (1.0 - rc["zeta_p"] + rc["zeta_p_pqm"][:, 13]) * u
            + (-1.0) * (alpha + delta * b_over_mu_m))
    up_b = (-1.0 * dot(rc["gamma_p_pqm"]) + (1.0 - rc["gamma_p"] + rc["gamma_p_pqm"][:, 13]) * u
            + (1.0 - rc["gamma_p"]) * (alpha + delta * b_over_mu_m))
    um_B = torch.where(delta > 0, um_b, u)
    up_B = torch.where(delta > 0, u, up_b)
    um_i = torch.where(mu_m > mu_p, um_A, um_B)
    up_i = torch.where(mu_m > mu_p, up_A, up_B)
    # bulk_point :456-462
    um_bulk = torch.where(flag == -1, u, torch.zeros_like(u))
    up_bulk = torch.where(flag == 1, u, torch.zeros_like(u))
    return torch.where(flag == 0, um_i, um_bulk), torch.where(flag == 0, up_i, up_bulk)


# ----------------------------------------------------------------------------------------
# a7: one finite-volume row per point   (discretization.py:299-423)
# ----------------------------------------------------------------------------------------
def is_box_boundary(points: Tensor, dx, dy, dz, bounds) -> Tensor:
    """:319-333"""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    xmin, xmax, ymin, ymax, zmin, zmax = bounds
    on = ((x - xmin).abs() < 1e-6 * dx) | ((x - xmax).abs() < 1e-6 * dx)
    on = on | ((y - ymin).abs() < 1e-6 * dy) | ((y - ymax).abs() < 1e-6 * dy)
    on = on | ((z - zmin).abs() < 1e-6 * dz) | ((z - zmax).abs() < 1e-6 * dz)
    return on


def compute_Ax_and_b(params: Tensor, points: Tensor, dx, dy, dz, prob: OracleProblem,
                     return_parts: bool = False):
    """compute_Ax_and_b_preconditioned_fn (discretization.py:299-423).  The preconditioner is 1.0 when
    disabled (trainer.py:246-253), else P(coeffs_) with the parameters stored after the network's.
    Returns (lhs/diag * P, rhs/diag * P), each (n,)."""
    dt = params.dtype
    points = points.to(dt)
    n = points.shape[0]
    dx, dy, dz = (torch.as_tensor(v, dtype=dt) for v in (dx, dy, dz))
    with torch.no_grad():
        coeffs_ = cell_faces_areas_values(points, dx, dy, dz, prob.phi_fn, prob.mu_m_fn, prob.mu_p_fn)
        c = coeffs_[:, :12]
        V_m, V_p = coeffs_[:, 12], coeffs_[:, 13]
        vol_nom = dx * dy * dz
        k_m = prob.k_m_fn(points).to(dt) * torch.ones(n, dtype=dt)
        k_p = prob.k_p_fn(points).to(dt) * torch.ones(n, dtype=dt)
        on_bnd = is_box_boundary(points, dx, dy, dz, prob.bounds)
    offs = torch.zeros((7, 3), dtype=dt)
    offs[1, 0], offs[2, 0] = -dx, dx
    offs[3, 1], offs[4, 1] = -dy, dy
    offs[5, 2], offs[6, 2] = -dz, dz
    um, up = [], []
    for s in range(7):
        a, b = u_mp_at_sites(params, points + offs[s], dx, dy, dz, prob)
        um.append(a); up.append(b)
    sum_m = c[:, 0] + c[:, 2] + c[:, 4] + c[:, 6] + c[:, 8] + c[:, 10]
    sum_p = c[:, 1] + c[:, 3] + c[:, 5] + c[:, 7] + c[:, 9] + c[:, 11]
    lhs = k_m * V_m * um[0]
    lhs = lhs + k_p * V_p * up[0]
    lhs = lhs + (prob.nonlinear_op_m(um[0]) * V_m + prob.nonlinear_op_p(up[0]) * V_p)
    lhs = lhs + (sum_m * um[0] + sum_p * up[0])
    for s in range(6):
        lhs = lhs + (-1.0 * c[:, 2 * s] * um[s + 1] - c[:, 2 * s + 1] * up[s + 1])
    diag = k_p * V_p + k_m * V_m + sum_m + sum_p
    # box boundary row :389-393
    u_b = prob.solution(params, points)
    lhs = torch.where(on_bnd, u_b * vol_nom, lhs)
    diag = torch.where(on_bnd, vol_nom * torch.ones_like(diag), diag)
    with torch.no_grad():
        rhs = (prob.f_m_fn(points).to(dt) * V_m + prob.f_p_fn(points).to(dt) * V_p
               + integrate_over_interface(points, dx, dy, dz, prob.phi_fn, prob.beta_fn))
        rhs_b = prob.dir_bc_fn(points).to(dt) * vol_nom
        rhs = torch.where(on_bnd, rhs_b * torch.ones_like(rhs), rhs)
    lhs_n = nan_to_num(lhs / diag)
    rhs_n = nan_to_num(rhs / diag)
    if prob.precond is not None:   # discretization.py:339, 418-419
        Pc = precond_eval(params[prob.shape.n_params:], prob.precond, coeffs_)
        lhs_n, rhs_n = lhs_n * Pc, rhs_n * Pc
    if return_parts:
        return lhs_n, rhs_n, dict(coeffs=coeffs_, diag=diag, rhs=rhs, on_bnd=on_bnd)
    return lhs_n, rhs_n


# ----------------------------------------------------------------------------------------
# a6: loss, a16: value_and_grad, a17: optax chain, a18: loops
# ----------------------------------------------------------------------------------------
def loss_fn(params: Tensor, points: Tensor, dx, dy, dz, prob: OracleProblem) -> Tensor:
    """Trainer.loss trainer.py:892-912 : mean(0.5*(lhs-rhs)^2)  (optax.l2_loss)."""
    lhs, rhs = compute_Ax_and_b(params, points, dx, dy, dz, prob)
    return (0.5 * (lhs - rhs) ** 2).mean()


def loss_and_grad(params: Tensor, points: Tensor, dx, dy, dz, prob: OracleProblem, chunk: int = 8192):
    """value_and_grad(self.loss) (trainer.py:786, 826), chunked over points to bound memory."""
    n = points.shape[0]
    p = params.detach().clone().requires_grad_(True)
    total = torch.zeros((), dtype=params.dtype)
    grad = torch.zeros_like(params)
    for s in range(0, n, chunk):
        lhs, rhs = compute_Ax_and_b(p, points[s:s + chunk], dx, dy, dz, prob)
        part = (0.5 * (lhs - rhs) ** 2).sum() / n
        g, = torch.autograd.grad(part, p)
        grad += g
        total += part.detach()
    return total, grad


class OptaxCustom:
    """chained_adam (solvers/optimizers.py:33-54): clip_by_global_norm(1.0) -> scale_by_adam ->
    scale_by_schedule(exponential_decay(lr, 1000, decay)) -> scale(-1).  optax 0.1.5 semantics."""

    def __init__(self, n: int, learning_rate=1e-3, decay_rate=0.96, transition_steps=1000,
                 max_norm=1.0, dtype=torch.float32, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.decay, self.ts, self.max_norm = learning_rate, decay_rate, transition_steps, max_norm
        self.b1, self.b2, self.eps = b1, b2, eps
        self.m = torch.zeros(n, dtype=dtype)
        self.v = torch.zeros(n, dtype=dtype)
        self.count = 0  # shared by scale_by_adam and scale_by_schedule

    def update(self, g: Tensor) -> Tensor:
        dt = g.dtype
        gnorm = torch.sqrt((g * g).sum())
        # optax.clip_by_global_norm: where(g_norm < max_norm, g, g / g_norm * max_norm)
        g = torch.where(gnorm < self.max_norm, g, (g / gnorm.to(dt)) * self.max_norm)
        self.m = (1 - self.b1) * g + self.b1 * self.m
        self.v = (1 - self.b2) * (g * g) + self.b2 * self.v
        t = self.count + 1
        m_hat = self.m / (1 - self.b1 ** t)
        v_hat = self.v / (1 - self.b2 ** t)
        upd = m_hat / (torch.sqrt(v_hat) + self.eps)
        step_size = self.lr * self.decay ** (self.count / self.ts)   # schedule sees the pre-increment count
        self.count += 1
        return -1.0 * (step_size * upd)


def batch_points(points: Tensor, batch_size: int, num_gpus: int = 1):
    """DatasetDict (data/data_management.py:79-186): per device a list of batches, contiguous blocks of the point list
    (:121-130).  Divisible sizes are the reference exactly.  For ragged sizes the reference fills the short last
    batch with random points from jax PRNGKey(0) (:70-76, :135-150), which cannot be restated without jax: here -
    as in the product - the short batch keeps its real points only, and a device whose block ends early gets empty
    batches (so that every device takes the same number of steps)."""
    n = points.shape[0]
    per = math.ceil(n / num_gpus)
    b = min(batch_size, per)
    nb = per // b + (1 if per % b else 0)
    out = []
    for g in range(num_gpus):
        dev = []
        for k in range(nb):
            lo = min(g * per + k * b, n)
            hi = min(lo + (b if k < per // b else per % b), n)
            dev.append(points[lo:hi])
        out.append(dev)
    return out


def single_gpu_train(params: Tensor, points: Tensor, grid_d, prob: OracleProblem, num_epochs: int,
                     batch_size: int, optimizer_dict=None, multires: bool = True):
    """single_GPU_train trainer.py:501-591 with alternate_res_sequentially (data_management.py:320-326)."""
    od = optimizer_dict or {"learning_rate": 1e-3, "sched": {"decay_rate": 0.96}}
    opt = OptaxCustom(params.numel(), od["learning_rate"], od["sched"]["decay_rate"], dtype=params.dtype)
    batches = batch_points(points, batch_size)[0]
    losses = []
    for epoch in range(num_epochs):
        zoom = epoch // (num_epochs // 4) if multires else 0
        d = [g * 0.5 ** zoom for g in grid_d]
        acc = 0.0
        for pts in batches:
            l, g = loss_and_grad(params, pts, d[0], d[1], d[2], prob)
            params = params + opt.update(g)
            acc += float(l)
        losses.append(acc / len(batches))
    return params, losses


def multi_gpu_train(params: Tensor, points: Tensor, grid_d, prob: OracleProblem, num_epochs: int,
                    batch_size: int, n_devices: int, optimizer_dict=None):
    """multi_GPU_train trainer.py:715-779 : per-device mean loss/grad, psum (SUM) over devices
    (:829-830), identical update everywhere, no multi-resolution schedule."""
    od = optimizer_dict or {"learning_rate": 1e-3, "sched": {"decay_rate": 0.96}}
    opt = OptaxCustom(params.numel(), od["learning_rate"], od["sched"]["decay_rate"], dtype=params.dtype)
    data = batch_points(points, n_devices * batch_size, n_devices)   # [device][batch] -> (B, 3)
    nb = len(data[0])
    losses = []
    for epoch in range(num_epochs):
        acc = 0.0
        for b in range(nb):
            gsum = torch.zeros_like(params); lsum = 0.0
            for dev in range(n_devices):
                if data[dev][b].shape[0] == 0:
                    continue
                l, g = loss_and_grad(params, data[dev][b], grid_d[0], grid_d[1], grid_d[2], prob)
                gsum += g; lsum += float(l)
            params = params + opt.update(gsum)
            acc += lsum
        losses.append(acc / nb)
    return params, losses


# ----------------------------------------------------------------------------------------
# 3.4 post-training evaluation  (trainer.py:960-977)
# ----------------------------------------------------------------------------------------
def evaluate_solution_and_gradients(params: Tensor, R: Tensor, dx, dy, dz, prob: OracleProblem):
    R = R.to(params.dtype).clone().requires_grad_(True)
    phi = prob.phi_fn(R.detach())
    u = prob.solution(params, R, phi)
    grad_u, = torch.autograd.grad(u.sum(), R)
    normals = normal_point_fn(R.detach(), dx, dy, dz, prob.phi_fn)
    grad_n = (normals * grad_u).sum(dim=1)
    return u.detach(), grad_u, grad_n
