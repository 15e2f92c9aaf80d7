"""Per-level, parameter-free set-up of the NBM step (host orchestration of K1/K2 over the C ABI).

Everything the reference recomputes for every point at every optimizer step although it does not
depend on the network parameters (crossing flags, cut-cell fractions, regression weights, face
coefficients; discretization.py:337-408) is computed here ONCE per (grid, cell size) and stored in
HBM as the row tables the step kernels stream.

User coefficient callables cannot run inside a CUDA kernel: the kernels emit the positions they
need (face centres, Gamma-triangle vertices, projected points), the host evaluates the batched
callables there with torch on the device, and K2 consumes the sample arrays.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import _cabi as cabi

NL_NONE, NL_SINH = 0, 1
INTERP = {"trilinear": 0, "quadratic": 1}


class Nonlinear:
    """N(u) = coef*sinh(u) (examples/biomolecules/coefficients.py:126-131) or nothing.  The
    reference takes an arbitrary callable (discretization.py:369); a CUDA kernel cannot, so the
    operator is named.  `Nonlinear.sinh(c)` objects are also callable so that the same object can be
    handed to code that wants the reference's callable form."""

    def __init__(self, kind: int = NL_NONE, coef: float = 0.0):
        self.kind, self.coef = kind, float(coef)

    @staticmethod
    def sinh(coef: float) -> "Nonlinear":
        return Nonlinear(NL_SINH, coef)

    def __call__(self, u):
        if self.kind == NL_NONE:
            return 0.0 * u
        return self.coef * torch.sinh(u)

    @staticmethod
    def coerce(op) -> "Nonlinear":
        if op is None:
            return Nonlinear()
        if isinstance(op, Nonlinear):
            return op
        # a plain callable: accept it only if it is identically zero (the reference default,
        # trainer.py:1007-1017); anything else is outside the kernel contract.
        probe = torch.linspace(-1.0, 1.0, 5)
        val = op(probe)
        val = val if isinstance(val, torch.Tensor) else torch.as_tensor(val, dtype=torch.float32)
        if bool((val == 0).all()):
            return Nonlinear()
        raise NotImplementedError(
            "nonlinear_op must be None, identically zero, or jax_dips_b200.Nonlinear.sinh(coef): "
            "arbitrary callables cannot be evaluated inside the CUDA step kernels")


class NetShape:
    """model_dict["mlp"] (trainer.py:116-123): tanh heads u^+ (p) and u^- (m)."""

    def __init__(self, layers_p=2, hidden_p=10, layers_m=1, hidden_m=1):
        self.layers_p, self.hidden_p, self.layers_m, self.hidden_m = layers_p, hidden_p, layers_m, hidden_m

    @staticmethod
    def from_model_dict(model_dict: dict) -> "NetShape":
        mt = model_dict.get("model_type", "mlp")
        if mt != "mlp":
            raise NotImplementedError(f"model_type {mt!r}: only the compact tanh 'mlp' surrogate is on this path")
        m = model_dict["mlp"]
        for k in ("activation_m", "activation_p"):
            if m.get(k, "jnp.tanh") not in ("jnp.tanh", "nn.tanh", "tanh"):
                raise NotImplementedError(f"{k}={m[k]!r}: the kernels implement tanh")
        return NetShape(m["hidden_layers_p"], m["hidden_dim_p"], m["hidden_layers_m"], m["hidden_dim_m"])

    def struct(self) -> cabi.Net:
        return cabi.Net(self.layers_p, self.hidden_p, self.layers_m, self.hidden_m)

    @staticmethod
    def _count(L, H):
        return 3 * H + H + (L - 1) * (H * H + H) + H + 1

    @property
    def n_p(self): return self._count(self.layers_p, self.hidden_p)
    @property
    def n_m(self): return self._count(self.layers_m, self.hidden_m)
    @property
    def n_params(self): return self.n_p + self.n_m


class PrecondShape:
    """model_dict["preconditioner"] (examples/benchmark_LPBE/conf/lpbe.yaml:62-67; nn/preconditioner.py:10-35):
    tanh MLP 26 -> layer_widths -> 1 on the point's 26 cell coefficients, P = 0.5 + scaling_coeff * sigmoid(.).
    Flat parameter order (after the network's): Dense_0.kernel (in,out) row-major, Dense_0.bias, Dense_1..."""
    N_IN = 26

    def __init__(self, layer_widths=(8, 4), scaling_coeff: float = 1.0):
        self.widths, self.scale = tuple(int(d) for d in layer_widths), float(scaling_coeff)
        if len(self.widths) != 2:
            raise NotImplementedError("the preconditioner kernel is compiled for two hidden layers (lpbe.yaml: [8, 4])")

    @staticmethod
    def from_model_dict(model_dict: dict) -> Optional["PrecondShape"]:
        pc = model_dict.get("preconditioner") or {}
        if not pc.get("enable", False):
            return None
        return PrecondShape(pc.get("layer_widths", (8, 4)), pc.get("scaling_coeff", 1.0))

    @property
    def n_params(self) -> int:
        return int(cabi.lib().nbm_precond_num_params(self.widths[0], self.widths[1]))

    def layer_dims(self):
        dims, fan_in = [], self.N_IN
        for d in list(self.widths) + [1]:
            dims.append((fan_in, d))
            fan_in = d
        return dims


class LevelSet:
    """phi on `lvl_gstate` + the reference's interpolant, evaluated inside the kernels
    (interpolate.py:906-1021 trilinear, :388-569 non-oscillatory quadratic; level_set.py:34-48)."""
    analytic = False

    def __init__(self, lvl_gstate, phi_values: torch.Tensor, interp: str = "trilinear",
                 perturb_eps: float = 1e-10, device=None):
        if interp not in INTERP:
            raise ValueError(f"unknown level-set interpolant {interp!r}")
        self.device = torch.device(device if device is not None else "cuda")
        nx, ny, nz = lvl_gstate.shape()
        phi = phi_values.to(self.device, torch.float32).contiguous().reshape(-1)
        if phi.numel() != nx * ny * nz:
            raise ValueError("phi_values does not match lvl_gstate")
        x, y, z = (a.to(self.device) for a in (lvl_gstate.x, lvl_gstate.y, lvl_gstate.z))
        self.phi_g = torch.empty((nx + 2) * (ny + 2) * (nz + 2), dtype=torch.float32, device=self.device)
        self.xg = torch.empty(nx + 2, dtype=torch.float32, device=self.device)
        self.yg = torch.empty(ny + 2, dtype=torch.float32, device=self.device)
        self.zg = torch.empty(nz + 2, dtype=torch.float32, device=self.device)
        L = cabi.lib()
        with torch.cuda.device(self.device):
            cabi.check(L.nbm_ghost_layer_f32(cabi.ptr(phi), cabi.ptr(x), cabi.ptr(y), cabi.ptr(z), nx, ny, nz,
                                             cabi.ptr(self.phi_g), cabi.ptr(self.xg), cabi.ptr(self.yg),
                                             cabi.ptr(self.zg), cabi.stream_ptr()), "nbm_ghost_layer_f32")
        self.struct = cabi.Lvl(cabi.ptr(self.phi_g), cabi.ptr(self.xg), cabi.ptr(self.yg), cabi.ptr(self.zg),
                               nx + 2, ny + 2, nz + 2, INTERP[interp], float(perturb_eps))
        self.bounds = [float(v) for v in (lvl_gstate.xmin(), lvl_gstate.xmax(), lvl_gstate.ymin(),
                                          lvl_gstate.ymax(), lvl_gstate.zmin(), lvl_gstate.zmax())]

    def __call__(self, pts: torch.Tensor) -> torch.Tensor:
        """phi at (n,3) points, on the device, through the CUDA interpolation kernel."""
        pts = pts.to(self.device, torch.float32).contiguous()
        out = torch.empty(pts.shape[0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().nbm_phi_interp_f32(C.byref(self.struct), cabi.ptr(pts), pts.shape[0],
                                                     cabi.ptr(out), cabi.stream_ptr()), "nbm_phi_interp_f32")
        return out


class AnalyticLevelSet:
    """The level set given by the user's own (batched) callable, called wherever phi is needed - what the reference does
    with an analytic `lvl_set_fn` (discretization.py:90; tests/test_poisson.py drives it that way).  A callable cannot
    run inside a kernel, but every phi evaluation of the path happens at positions that are known per level (cell
    corners, the 27-cube of crossed sites, evaluation points +- d): they are formed here with the kernels' own fp32
    arithmetic, the callable is evaluated with torch on the device, and the kernels take the samples (nbm_lvl_t).
    The callable is used as given (wrap it in `perturb_level_set_fn` yourself, as the reference's configs do)."""
    analytic = True

    def __init__(self, lvl_gstate, phi_fn: Callable, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.phi_fn = phi_fn
        self.struct = cabi.Lvl()          # no grid: the per-call sample pointers are filled in where they are used
        self.bounds = [float(v) for v in (lvl_gstate.xmin(), lvl_gstate.xmax(), lvl_gstate.ymin(),
                                          lvl_gstate.ymax(), lvl_gstate.zmin(), lvl_gstate.zmax())]

    def __call__(self, pts: torch.Tensor) -> torch.Tensor:
        return _sample(self.phi_fn, pts.to(self.device, torch.float32))

    def with_samples(self, corner_phi=None, cube_phi=None, eval_phi=None) -> cabi.Lvl:
        s = cabi.Lvl()
        s.corner_phi, s.cube_phi, s.eval_phi = cabi.ptr(corner_phi), cabi.ptr(cube_phi), cabi.ptr(eval_phi)
        return s


_CORNER_SIGN = ((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1), (-1, 1, -1), (1, 1, -1), (-1, 1, 1), (1, 1, 1))


def _classify_analytic(lvl: "AnalyticLevelSet", coords, lo, hi, d, dev):
    """flag / side of every lattice site from the callable (geometric_integrations_per_point.py:203-263), plane by
    plane; positions as in site_position() and corner_phis() of the kernels: x = xs[i] + shift, corner = (sign*d)*0.5 + x."""
    xs, ys, zs, shifts = coords
    nx, ny, nz = xs.numel(), ys.numel(), zs.numel()
    n = nx * ny * nz
    flag = torch.empty(len(shifts) * n, dtype=torch.int8, device=dev)
    side = torch.zeros(len(shifts) * n + 32, dtype=torch.uint8, device=dev)[:len(shifts) * n]
    f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    half = [[(f32(sg[a]) * f32(d[a])) * f32(0.5) for a in range(3)] for sg in _CORNER_SIGN]
    chunk = max(1, (1 << 21) // (ny * nz))
    for k, sh in enumerate(shifts):
        X, Y, Z = xs + f32(sh[0]), ys + f32(sh[1]), zs + f32(sh[2])
        for i0 in range(0, nx, chunk):
            i1 = min(nx, i0 + chunk)
            gx, gy, gz = torch.meshgrid(X[i0:i1], Y, Z, indexing="ij")
            P = torch.stack((gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)), dim=1)
            ph0 = lvl(P)
            neg = torch.zeros(P.shape[0], dtype=torch.int32, device=dev)
            first = None
            for c in range(8):
                pc = lvl(torch.stack((half[c][0] + P[:, 0], half[c][1] + P[:, 1], half[c][2] + P[:, 2]), dim=1))
                neg += (pc < 0).to(torch.int32)
                if c == 0:
                    first = pc
            fl = torch.where((neg == 0) | (neg == 8), torch.sign(first).to(torch.int8), torch.zeros_like(neg, dtype=torch.int8))
            ii = torch.arange(i0, i1, device=dev)
            ok = ((ii >= lo[0]) & (ii < hi[0]))[:, None, None] & \
                 ((torch.arange(ny, device=dev) >= lo[1]) & (torch.arange(ny, device=dev) < hi[1]))[None, :, None] & \
                 ((torch.arange(nz, device=dev) >= lo[2]) & (torch.arange(nz, device=dev) < hi[2]))[None, None, :]
            fl = torch.where(ok.reshape(-1), fl, torch.full_like(fl, 2))
            a, b = k * n + i0 * ny * nz, k * n + i1 * ny * nz
            flag[a:b] = fl
            side[a:b] = ((ph0 >= 0).to(torch.uint8) | ((ph0 > 0).to(torch.uint8) << 1))
    return flag, side


def _lattice(xs, ys, zs, lo=None, hi=None, shifts=None) -> cabi.Lattice:
    lat = cabi.Lattice()
    lat.xs, lat.ys, lat.zs = cabi.ptr(xs), cabi.ptr(ys), cabi.ptr(zs)
    lat.nx, lat.ny, lat.nz = xs.numel(), ys.numel(), zs.numel()
    lo = lo or (0, 0, 0)
    hi = hi or (lat.nx, lat.ny, lat.nz)
    for a in range(3):
        lat.lo[a], lat.hi[a] = int(lo[a]), int(hi[a])
    shifts = shifts or [(0.0, 0.0, 0.0)]
    lat.n_shift = len(shifts)
    for k, sh in enumerate(shifts):
        for a in range(3):
            lat.shift[k][a] = float(sh[a])
    return lat


def _sample(fn: Callable, pts: torch.Tensor) -> torch.Tensor:
    out = fn(pts)
    if not isinstance(out, torch.Tensor):
        out = torch.as_tensor(out, dtype=torch.float32, device=pts.device)
    out = out.to(device=pts.device, dtype=torch.float32)
    if out.numel() == 1 and pts.shape[0] != 1:
        out = out.reshape(1).expand(pts.shape[0])
    return out.reshape(pts.shape[0]).contiguous()


class CrossedSites:
    """Crossed sites of a lattice: compaction, K1 cut-cell, K2a regression, K2b jump weights."""

    def __init__(self, lvl: LevelSet, lat: cabi.Lattice, n_sites: int, d, fns, dev, n_cut_sites: int = None,
                 coords=None):
        """`n_cut_sites`: cut-cell geometry is only needed for crossed sites with id < n_cut_sites (the
        training points themselves); default all.  `coords` = (xs, ys, zs, shifts) of the lattice as tensors: needed
        with an analytic level set, whose samples are taken here."""
        L = cabi.lib()
        st = cabi.stream_ptr()
        dx, dy, dz = d
        if lvl.analytic:
            lo = [lat.lo[a] for a in range(3)]
            hi = [lat.hi[a] for a in range(3)]
            self.flag, self.side = _classify_analytic(lvl, coords, lo, hi, d, dev)
        else:
            self.flag = torch.empty(n_sites, dtype=torch.int8, device=dev)
            # (16 bytes of slack: the fused gradient kernel stages `side` with 16-byte bulk copies)
            self.side = torch.zeros(n_sites + 32, dtype=torch.uint8, device=dev)[:n_sites]
            cabi.check(L.nbm_classify_f32(C.byref(lvl.struct), C.byref(lat), dx, dy, dz, cabi.ptr(self.flag),
                                          cabi.ptr(self.side), st), "nbm_classify_f32")
        # compaction
        idx = torch.empty(n_sites, dtype=torch.int64, device=dev)
        self.cidx = torch.empty(n_sites, dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        ws = C.c_size_t(0)
        cabi.check(L.nbm_compact_crossed(cabi.ptr(self.flag), n_sites, cabi.ptr(idx), n_sites, cabi.ptr(self.cidx),
                                         cabi.ptr(count), None, C.byref(ws), st), "nbm_compact_crossed(size)")
        work = torch.empty(max(int(ws.value), 1), dtype=torch.uint8, device=dev)
        cabi.check(L.nbm_compact_crossed(cabi.ptr(self.flag), n_sites, cabi.ptr(idx), n_sites, cabi.ptr(self.cidx),
                                         cabi.ptr(count), cabi.ptr(work), C.byref(ws), st), "nbm_compact_crossed")
        nc = int(count.item())
        self.n = nc
        self.idx = idx[:nc].clone()
        del idx, work
        n1 = max(nc, 1)
        self.frac = torch.zeros(n1 * 14, dtype=torch.float32, device=dev)
        self.tri = torch.zeros(n1 * 90, dtype=torch.float32, device=dev)
        self.tri_area = torch.zeros(n1 * 10, dtype=torch.float32, device=dev)
        self.pos = torch.zeros(n1 * 3, dtype=torch.float32, device=dev)
        self.proj = torch.zeros(n1 * 3, dtype=torch.float32, device=dev)
        self.delta = torch.zeros(n1, dtype=torch.float32, device=dev)
        self.Cm = torch.zeros(n1 * 27, dtype=torch.float32, device=dev)
        self.Cp = torch.zeros(n1 * 27, dtype=torch.float32, device=dev)
        self.cube_side = torch.zeros(n1, dtype=torch.int32, device=dev)
        self.B = torch.zeros(n1 * 28, dtype=torch.float32, device=dev)
        self.beta_gamma = torch.zeros(n1, dtype=torch.float32, device=dev)
        if nc == 0:
            return
        n_cut = nc if n_cut_sites is None else int((self.idx < n_cut_sites).sum().item())
        self.n_cut = n_cut
        lvl_struct = lvl.struct
        if lvl.analytic:
            # positions of the crossed sites (as site_position()), their 8 corners and their 27-cube, sampled here
            xs, ys, zs, shifts = coords
            nx, ny, nz = xs.numel(), ys.numel(), zs.numel()
            f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
            kk = self.idx // (nx * ny * nz)
            e = self.idx - kk * (nx * ny * nz)
            sh = torch.tensor(shifts, dtype=torch.float32, device=dev)[kk]
            sx = xs[e // (ny * nz)] + sh[:, 0]
            sy = ys[(e // nz) % ny] + sh[:, 1]
            sz = zs[e % nz] + sh[:, 2]
            corner = torch.empty((nc, 8), dtype=torch.float32, device=dev)
            for c, sg in enumerate(_CORNER_SIGN):
                hx, hy, hz = ((f32(sg[a]) * f32(d[a])) * f32(0.5) for a in range(3))
                corner[:, c] = lvl(torch.stack((hx + sx, hy + sy, hz + sz), dim=1))
            cube = torch.empty((nc, 27), dtype=torch.float32, device=dev)
            for q in range(27):
                X0, X1, X2 = f32(float(q % 3 - 1)) * f32(dx), f32(float((q // 3) % 3 - 1)) * f32(dy), f32(float(q // 9 - 1)) * f32(dz)
                cube[:, q] = lvl(torch.stack((sx + X0, sy + X1, sz + X2), dim=1))
            self._corner_phi, self._cube_phi = corner.contiguous(), cube.contiguous()
            lvl_struct = lvl.with_samples(corner_phi=self._corner_phi, cube_phi=self._cube_phi)
        cabi.check(L.nbm_cutcell_f32(C.byref(lvl_struct), C.byref(lat), dx, dy, dz, cabi.ptr(self.idx), n_cut,
                                     cabi.ptr(self.frac), cabi.ptr(self.tri), cabi.ptr(self.tri_area), st),
                   "nbm_cutcell_f32")
        cabi.check(L.nbm_regression_f32(C.byref(lvl_struct), C.byref(lat), dx, dy, dz, cabi.ptr(self.idx), nc,
                                        cabi.ptr(self.pos), cabi.ptr(self.proj), cabi.ptr(self.delta),
                                        cabi.ptr(self.Cm), cabi.ptr(self.Cp), cabi.ptr(self.cube_side), st),
                   "nbm_regression_f32")
        # ---- host samples the user callables at the positions the kernels emitted
        pos, proj = self.pos.view(nc, 3), self.proj.view(nc, 3)
        mu_m_s, mu_p_s = _sample(fns.mu_m_fn, pos), _sample(fns.mu_p_fn, pos)
        alpha_p, beta_p = _sample(fns.alpha_fn, proj), _sample(fns.beta_fn, proj)
        mu_m_p, mu_p_p = _sample(fns.mu_m_fn, proj), _sample(fns.mu_p_fn, proj)
        cabi.check(L.nbm_site_weights_f32(nc, cabi.ptr(self.delta), cabi.ptr(self.Cm), cabi.ptr(self.Cp),
                                          cabi.ptr(mu_m_s), cabi.ptr(mu_p_s), cabi.ptr(alpha_p), cabi.ptr(beta_p),
                                          cabi.ptr(mu_m_p), cabi.ptr(mu_p_p), cabi.ptr(self.B), st),
                   "nbm_site_weights_f32")
        # integral over Gamma of beta (geometric_integrations_per_point.py:385-412): sum_t area_t * mean_v beta
        area = self.tri_area.view(nc, 10)
        # Unused triangle slots are all-zero vertices with zero area: like the reference (:397-410) the
        # callable IS evaluated at the origin and multiplied by 0, so a beta that is singular at the
        # origin turns the integral into NaN, which nan_to_num(rhs/diag) (discretization.py:419) then
        # maps to 0.  Kept bit-for-bit: results must match the reference on the same inputs.
        beta_v = _sample(fns.beta_fn, self.tri.view(nc * 30, 3)).view(nc, 10, 3).mean(dim=2)
        self.beta_gamma = (area * beta_v).sum(dim=1).contiguous()
        torch.cuda.current_stream().synchronize()  # sample tensors die here


class SharedPlan:
    """Shared-evaluation plan for the x-planes [xa, xb) of the training grid at the native cell size
    (d == grid spacing): one lattice = slab + 2 halo planes in x, 1 halo layer in y and z."""

    HX, HY, HZ = 2, 1, 1

    def __init__(self, lvl: LevelSet, tr_gstate, xa: int, xb: int, fns, net: NetShape,
                 nonlinear_m: Nonlinear, nonlinear_p: Nonlinear, n_mean: Optional[int] = None, device=None,
                 faces: Optional[bool] = None, fused: Optional[bool] = None, precond: Optional[PrecondShape] = None,
                 deterministic: bool = False, stencil_tma: Optional[bool] = None, stash: Optional[bool] = None,
                 overlap_lists: Optional[bool] = None):
        """`faces`: store one coefficient per cell FACE + 1/diag (16 B/node) instead of the 7 row weights
        (28 B/node); irregular rows move into the list.  Default: on (lattice rows are padded to 16-byte multiples).
        `fused`: evaluate the dense adjoint stencil inside the gradient kernel from TMA-staged row tables
        instead of a separate pass (needs faces and no nonlinear operator).  Default OFF: measured on B200 at
        256^3 the fused kernel takes 516-585 us against 370 + 122 us for gradient + adjoint kernels (the
        shared ring couples the 12 warps of a CTA to the slowest one).
        `stencil_tma`: residual rows and adjoint stencil of the faces table as ONE kernel whose x planes arrive as 3-D
        TMA boxes (halo included) in a shared-memory ring (`csrc/nbm_stencil_tma.cuh`).  Default: on where it applies.
        `stash`: keep the last hidden layer of every plus-side node from the forward kernel (48 bytes per lattice node
        for hidden_p = 10) so that the gradient kernel does not recompute it.  Default OFF: measured on B200 at 256^3
        the gradient kernel gains 21 us (343 -> 323) but the forward kernel loses 38 us to the 830 MB of stores
        (172 -> 210); see DESIGN.md.
        `overlap_lists`: run the list kernels (crossed sites, irregular rows) on a side stream BESIDE the TMA stencil
        kernel instead of after it (they need only U); their adjoint lands in a side buffer that a small merge kernel
        adds to G.  Default: on with the faces table (TMA stencil or the two separate kernels) when `deterministic` is off.
        `deterministic`: gather the adjoint of the lists (irregular rows, extrapolation) through their transposed
        incidence instead of scattering it with fp32 atomics: the whole step becomes bitwise reproducible, for
        ~5 us more per step at 256^3 (the atomics are faster than the doubly indirect gathers)."""
        dev = torch.device(device if device is not None else lvl.device)
        self.device, self.net, self.lvl = dev, net, lvl
        L = cabi.lib()
        with torch.cuda.device(dev):
            st = cabi.stream_ptr()
            Nx, Ny, Nz = tr_gstate.shape()
            assert 0 <= xa < xb <= Nx
            xs, ys, zs = (a.to(dev) for a in (tr_gstate.x, tr_gstate.y, tr_gstate.z))
            dx, dy, dz = float(tr_gstate.dx), float(tr_gstate.dy), float(tr_gstate.dz)
            self.d = (dx, dy, dz)
            nxl = xb - xa
            self.n_points = nxl * Ny * Nz

            def extend(a, lo, hi, h, step, h_hi=None):
                # coordinates of global indices [lo-h, hi+h_hi): grid values inside, a[0]-k*step / a[-1]+k*step outside
                n = a.numel()
                gi = torch.arange(lo - h, hi + (h if h_hi is None else h_hi), device=dev)
                inside = a[gi.clamp(0, n - 1)]
                st32 = torch.tensor(step, dtype=torch.float32, device=dev)
                below = a[0] - (-gi).clamp(min=0).to(torch.float32) * st32
                above = a[n - 1] + (gi - (n - 1)).clamp(min=0).to(torch.float32) * st32
                return torch.where(gi < 0, below, torch.where(gi > n - 1, above, inside)).contiguous()

            self.xe = extend(xs, xa, xb, self.HX, dx)
            self.ye = extend(ys, 0, Ny, self.HY, dy)
            # the z halo is widened on the high side until a lattice row is a whole number of 16-byte groups: rows
            # start 16-byte aligned (float4 stencil kernels on any grid) and the lattice arrays are legal TMA tensors
            self.hz_hi = self.HZ + (-(Nz + 2 * self.HZ)) % 4
            self.ze = extend(zs, 0, Nz, self.HZ, dz, self.hz_hi)
            ex, ey, ez = self.xe.numel(), self.ye.numel(), self.ze.numel()
            self.dims = (ex, ey, ez)
            ne = ex * ey * ez
            self.ne = ne
            # sites: real grid nodes within one plane of the slab
            gx_lo, gx_hi = max(xa - 1, 0), min(xb + 1, Nx)
            lo = (gx_lo - (xa - self.HX), self.HY, self.HZ)
            hi = (gx_hi - (xa - self.HX), self.HY + Ny, self.HZ + Nz)
            lat = _lattice(self.xe, self.ye, self.ze, lo, hi)
            self.sites = CrossedSites(lvl, lat, ne, self.d, fns, dev, coords=(self.xe, self.ye, self.ze, [(0.0, 0.0, 0.0)]))
            cs = self.sites

            # ---- per-point coefficient samples
            pxs = xs[xa:xb].contiguous()
            np_ = self.n_points
            mu_m_faces, mu_p_faces, k_m, k_p, f_m, f_p, g_dir = _point_samples(fns, pxs, ys, zs, self.d, dev)

            # ---- K2c row assembly into lattice layout
            use_nl = (nonlinear_m.kind != NL_NONE) or (nonlinear_p.kind != NL_NONE)
            can_faces = (ez % 4 == 0)
            if faces and not can_faces:
                raise ValueError("faces=True needs 16-byte aligned lattice rows")
            self.faces = can_faces if faces is None else bool(faces)
            self.w = None if self.faces else torch.zeros(7 * ne, dtype=torch.float32, device=dev)
            self.cface = torch.zeros(3 * ne, dtype=torch.float32, device=dev) if self.faces else None
            self.dinv = torch.zeros(ne, dtype=torch.float32, device=dev) if self.faces else None
            use_kv = self.faces and bool(((k_m != 0) | (k_p != 0)).any().item())
            self.kv = torch.zeros(ne, dtype=torch.float32, device=dev) if use_kv else None
            self.rhs = torch.zeros(ne, dtype=torch.float32, device=dev)
            can_fuse = self.faces and not use_nl
            if fused and not can_fuse:
                raise ValueError("fused=True needs the faces table and no nonlinear operator")
            self.fused = False if fused is None else bool(fused)
            self.S = torch.zeros(ne, dtype=torch.float32, device=dev) if self.fused else None
            self.nl = torch.zeros(2 * ne, dtype=torch.float32, device=dev) if use_nl else None
            irr = torch.full((ne,), -1, dtype=torch.int32, device=dev)
            cap = min(np_, 7 * cs.n) + 1
            irr_count = torch.zeros(1, dtype=torch.int64, device=dev)
            irr_point = torch.zeros(cap, dtype=torch.int64, device=dev)
            irr_wE = torch.zeros(cap * 7, dtype=torch.float32, device=dev)
            irr_c = torch.full((cap * 7,), -1, dtype=torch.int32, device=dev)
            irr_nl = torch.zeros(cap, dtype=torch.uint8, device=dev)
            irr_nlw = torch.zeros(cap, dtype=torch.float32, device=dev)
            if self.faces:
                # a row is also irregular when a neighbour sits on the other side of the interface
                cap = min(np_, 7 * cs.n + np_ // 64 + 1024) + 1
                irr_point = torch.zeros(cap, dtype=torch.int64, device=dev)
                irr_wE = torch.zeros(cap * 7, dtype=torch.float32, device=dev)
                irr_c = torch.full((cap * 7,), -1, dtype=torch.int32, device=dev)
                irr_nl = torch.zeros(cap, dtype=torch.uint8, device=dev)
                irr_nlw = torch.zeros(cap, dtype=torch.float32, device=dev)
            irr_wU = torch.zeros(cap * 7, dtype=torch.float32, device=dev) if self.faces else None
            irr_rhs = torch.zeros(cap, dtype=torch.float32, device=dev) if self.faces else None
            self.precond = precond
            self.coef26 = torch.zeros(26 * ne, dtype=torch.float32, device=dev) if precond is not None else None
            a = cabi.Assemble()
            a.coef26 = cabi.ptr(self.coef26)
            a.pts = _lattice(pxs, ys, zs)
            a.dx, a.dy, a.dz = dx, dy, dz
            for i, b in enumerate(lvl.bounds):
                a.bounds[i] = b
            a.shared = 1
            a.site_dims[0], a.site_dims[1], a.site_dims[2] = ex, ey, ez
            a.pt_off[0], a.pt_off[1], a.pt_off[2] = self.HX, self.HY, self.HZ
            a.flag, a.side, a.cidx = cabi.ptr(cs.flag), cabi.ptr(cs.side), cabi.ptr(cs.cidx)
            a.frac, a.beta_gamma = cabi.ptr(cs.frac), cabi.ptr(cs.beta_gamma)
            a.mu_m_faces, a.mu_p_faces = cabi.ptr(mu_m_faces), cabi.ptr(mu_p_faces)
            a.k_m, a.k_p, a.f_m, a.f_p, a.g_dir = (cabi.ptr(t) for t in (k_m, k_p, f_m, f_p, g_dir))
            a.w, a.rhs, a.nl, a.irr = cabi.ptr(self.w), cabi.ptr(self.rhs), cabi.ptr(self.nl), cabi.ptr(irr)
            a.n_out = ne
            a.out_stride[0], a.out_stride[1], a.out_stride[2] = ey * ez, ez, 1
            a.out_off = (self.HX * ey + self.HY) * ez + self.HZ
            a.irr_capacity = cap
            a.irr_count, a.irr_point = cabi.ptr(irr_count), cabi.ptr(irr_point)
            a.irr_wE, a.irr_c, a.irr_nl, a.irr_nlw = (cabi.ptr(t) for t in (irr_wE, irr_c, irr_nl, irr_nlw))
            a.faces = 1 if self.faces else 0
            a.cface, a.dinv, a.kv = cabi.ptr(self.cface), cabi.ptr(self.dinv), cabi.ptr(self.kv)
            a.irr_wU, a.irr_rhs = cabi.ptr(irr_wU), cabi.ptr(irr_rhs)
            cabi.check(L.nbm_assemble_f32(C.byref(a), st), "nbm_assemble_f32")
            n_irr = int(irr_count.item())
            if n_irr > cap:
                raise cabi.NbmError(f"irregular-row capacity exceeded ({n_irr} > {cap})")
            self.n_irr = n_irr
            # the assembly kernel appends irregular rows in atomic order: sort them by lattice node so that the list
            # kernels touch U / R / G with locality (and the order, hence the fp32 atomics' operands, is reproducible)
            m = max(n_irr, 1)
            perm = torch.argsort(irr_point[:n_irr]) if n_irr > 1 else torch.arange(m, device=dev)
            take = lambda t, w=1: (t[:m * w].view(m, w)[perm].reshape(-1).contiguous() if n_irr > 1 else t[:m * w].clone())
            self.irr_wU = take(irr_wU, 7) if self.faces else None
            self.irr_rhs = take(irr_rhs) if self.faces else None
            self.irr_point = take(irr_point)
            self.irr_wE = take(irr_wE, 7)
            self.irr_c = take(irr_c, 7)
            self.irr_nl = take(irr_nl)
            self.irr_nlw = take(irr_nlw)
            if n_irr > 1:   # lattice node -> slot map follows the permutation
                inv = torch.empty_like(perm)
                inv[perm] = torch.arange(n_irr, device=dev)
                has = irr >= 0
                irr[has] = inv[irr[has].long()].to(torch.int32)
            self.irr = irr
            del mu_m_faces, mu_p_faces, k_m, k_p, f_m, f_p, g_dir
            if self.fused and precond is not None:
                raise ValueError("the fused adjoint path does not take a preconditioner")
            if self.fused:
                # bit 2 of `side`: the node can receive a contribution from the lists (the 27-cube of a crossed
                # site, the 7 stencil sites of an irregular row); only those read (and re-zero) G in the step
                sxy, sy = ey * ez, ez
                tg = []
                if cs.n > 0:
                    o27 = torch.tensor([a * sxy + b * sy + c for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)],
                                       dtype=torch.int64, device=dev)
                    tg.append((cs.idx[:, None] + o27[None, :]).reshape(-1))
                if n_irr > 0:
                    o7 = torch.tensor([0, -sxy, sxy, -sy, sy, -1, 1], dtype=torch.int64, device=dev)
                    tg.append((self.irr_point[:n_irr, None] + o7[None, :]).reshape(-1))
                if tg:
                    nodes = torch.unique(torch.cat(tg))
                    assert int(nodes.min()) >= 0 and int(nodes.max()) < ne
                    cs.side[nodes] = cs.side[nodes] | 4

            # ---- transposed incidence of the lists (CSR): the adjoint of the irregular rows and of the extrapolation is
            # gathered per crossed site / per receiving node - no atomics, fixed summation order (bitwise reproducible)
            self.ge_ptr = self.ge_ent = self.g_ptr = self.g_ent = self.list_nodes = None
            self.n_list = 0
            if deterministic and not self.fused and (cs.n > 0 or n_irr > 0):
                def csr(targets, payload, n_targets):
                    order = torch.argsort(targets, stable=True)
                    ptr = torch.zeros(n_targets + 1, dtype=torch.int64, device=dev)
                    ptr[1:] = torch.cumsum(torch.bincount(targets, minlength=n_targets), 0)
                    assert int(ptr[-1]) < 2 ** 31
                    return ptr.to(torch.int32).contiguous(), payload[order].to(torch.int32).contiguous()

                sxy, sy = ey * ez, ez
                k7 = torch.arange(7, device=dev)
                if cs.n > 0:
                    ic = self.irr_c[:n_irr * 7].view(n_irr, 7).long() if n_irr > 0 else torch.zeros((0, 7), dtype=torch.int64, device=dev)
                    qq = torch.arange(ic.shape[0], device=dev)[:, None].expand(-1, 7)
                    m = ic >= 0
                    self.ge_ptr, self.ge_ent = csr(ic[m], (qq * 8 + k7[None, :])[m], cs.n)
                tg, pay = [], []
                if self.faces and n_irr > 0:
                    o7 = torch.tensor([0, -sxy, sxy, -sy, sy, -1, 1], dtype=torch.int64, device=dev)
                    wu = self.irr_wU[:n_irr * 7].view(n_irr, 7)
                    m = wu != 0
                    qq = torch.arange(n_irr, device=dev)[:, None].expand(-1, 7)
                    tg.append((self.irr_point[:n_irr, None] + o7[None, :])[m])
                    pay.append((qq * 8 + k7[None, :])[m])
                if cs.n > 0:
                    # cube vertex v of extrap_kernel: offset (v%3-1) sx + ((v/3)%3-1) sy + (v/9-1)
                    v27 = torch.arange(27, device=dev)
                    o27 = (v27 % 3 - 1) * sxy + ((v27 // 3) % 3 - 1) * sy + (v27 // 9 - 1)
                    Bm = cs.B.view(-1, 28)[:cs.n, :27] != 0
                    cc = torch.arange(cs.n, device=dev)[:, None].expand(-1, 27)
                    tg.append((cs.idx[:cs.n, None] + o27[None, :])[Bm])
                    pay.append((-(cc * 32 + v27[None, :]) - 1)[Bm])
                if tg:
                    tgt = torch.cat(tg)
                    assert int(tgt.min()) >= 0 and int(tgt.max()) < ne
                    self.list_nodes, inv = torch.unique(tgt, return_inverse=True)
                    self.n_list = int(self.list_nodes.numel())
                    self.g_ptr, self.g_ent = csr(inv, torch.cat(pay), self.n_list)
                    self.list_nodes = self.list_nodes.contiguous()
            # ---- work buffers + the step descriptor
            P = net.n_params + (precond.n_params if precond is not None else 0)
            self.n_total = P
            self.U = torch.zeros(ne, dtype=torch.float32, device=dev)
            self.R = torch.zeros(ne, dtype=torch.float32, device=dev)  # halo rows stay 0 for ever
            self.G = torch.zeros(ne, dtype=torch.float32, device=dev)
            self.E = torch.zeros(max(cs.n, 1), dtype=torch.float32, device=dev)
            self.gE = torch.zeros(max(cs.n, 1), dtype=torch.float32, device=dev)
            # the shared path launches at most one gradient CTA per SM: that many partial rows
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            n_pc_rows = 3 * sms if precond is not None else 0
            rows = min(L.nbm_step_partial_rows(), sms) + n_pc_rows
            self.partials = torch.zeros(rows * (P + 1), dtype=torch.float32, device=dev)
            self.loss_grad = torch.zeros(P + 1, dtype=torch.float32, device=dev)
            s = cabi.SharedStep()
            s.net = net.struct()
            s.nonlinear_m, s.nonlinear_p = nonlinear_m.kind, nonlinear_p.kind
            s.nl_coef_m, s.nl_coef_p = nonlinear_m.coef, nonlinear_p.coef
            s.xe, s.ye, s.ze = cabi.ptr(self.xe), cabi.ptr(self.ye), cabi.ptr(self.ze)
            s.ex, s.ey, s.ez = ex, ey, ez
            s.side, s.w, s.rhs, s.nl = cabi.ptr(cs.side), cabi.ptr(self.w), cabi.ptr(self.rhs), cabi.ptr(self.nl)
            s.n_crossed, s.c_node, s.B = cs.n, cabi.ptr(cs.idx), cabi.ptr(cs.B)
            s.n_irr = n_irr
            s.irr_point, s.irr_wE, s.irr_c = cabi.ptr(self.irr_point), cabi.ptr(self.irr_wE), cabi.ptr(self.irr_c)
            s.irr_nl, s.irr_nlw = cabi.ptr(self.irr_nl), cabi.ptr(self.irr_nlw)
            s.inv_n_points = 1.0 / float(n_mean if n_mean is not None else self.n_points)
            s.U, s.R, s.G, s.E, s.gE = (cabi.ptr(t) for t in (self.U, self.R, self.G, self.E, self.gE))
            s.partials, s.n_partial_rows, s.loss_grad = cabi.ptr(self.partials), rows, cabi.ptr(self.loss_grad)
            s.faces = 1 if self.faces else 0
            s.cface, s.dinv, s.kv = cabi.ptr(self.cface), cabi.ptr(self.dinv), cabi.ptr(self.kv)
            s.irr_wU, s.irr_rhs = cabi.ptr(self.irr_wU), cabi.ptr(self.irr_rhs)
            s.S = cabi.ptr(self.S)
            if self.g_ptr is not None:
                s.ge_ptr, s.ge_ent = cabi.ptr(self.ge_ptr), cabi.ptr(self.ge_ent)
                s.list_nodes, s.n_list = cabi.ptr(self.list_nodes), self.n_list
                s.g_ptr, s.g_ent = cabi.ptr(self.g_ptr), cabi.ptr(self.g_ent)
            # activation stash of the forward kernel
            self.stash = (os.environ.get("NBM_STASH", "0") != "0") if stash is None else bool(stash)
            self.Hst = torch.zeros(4 * ((net.hidden_p // 2 + 1) // 2) * ne, dtype=torch.float32, device=dev) if self.stash else None
            s.Hst = cabi.ptr(self.Hst)
            # dense stencil stage: one TMA-fed kernel for residual rows + adjoint (default) or the two separate kernels
            self.stencil_tma = (os.environ.get("NBM_STENCIL_TMA", "1") != "0") if stencil_tma is None else bool(stencil_tma)
            s.stencil_tma = 0 if self.stencil_tma else -1
            # (mirrors the library's own test, nbm_step.cu launch_shared)
            self.stencil_tma_active = bool(self.stencil_tma and self.faces and precond is None and not self.fused
                                           and ez % 4 == 0 and not use_nl)
            # list chain beside the dense stencil: side buffer G2, compact residuals Rq, the nodes the lists can reach
            want = (os.environ.get("NBM_OVERLAP_LISTS", "1") != "0") if overlap_lists is None else bool(overlap_lists)
            self.overlap_lists = bool(want and self.faces and precond is None and not self.fused and self.g_ptr is None
                                      and (cs.n > 0 or n_irr > 0))
            self.G2 = self.Rq = None
            if self.overlap_lists:
                sxy, sy = ey * ez, ez
                tg = []
                if cs.n > 0:
                    v27 = torch.arange(27, device=dev)
                    o27 = (v27 % 3 - 1) * sxy + ((v27 // 3) % 3 - 1) * sy + (v27 // 9 - 1)
                    tg.append((cs.idx[:cs.n, None] + o27[None, :]).reshape(-1))
                if n_irr > 0:
                    o7 = torch.tensor([0, -sxy, sxy, -sy, sy, -1, 1], dtype=torch.int64, device=dev)
                    tg.append((self.irr_point[:n_irr, None] + o7[None, :]).reshape(-1))
                self.list_nodes = torch.unique(torch.cat(tg)).contiguous()
                assert int(self.list_nodes.min()) >= 0 and int(self.list_nodes.max()) < ne
                self.n_list = int(self.list_nodes.numel())
                self.G2 = torch.zeros(ne, dtype=torch.float32, device=dev)
                self.Rq = torch.zeros(max(n_irr, 1), dtype=torch.float32, device=dev)
                s.list_nodes, s.n_list = cabi.ptr(self.list_nodes), self.n_list
                s.G2, s.Rq = cabi.ptr(self.G2), cabi.ptr(self.Rq)
                # slot-major copies of the list tables for the chain kernels (thread per site / per row)
                self.B_soa = self.irr_c_soa = self.irr_wE_soa = self.irr_wU_soa = None
                if os.environ.get("NBM_LIST_SOA", "1") != "0":
                    if cs.n > 0:
                        self.B_soa = cs.B.view(-1, 28)[:cs.n].t().contiguous()
                        s.B_soa = cabi.ptr(self.B_soa)
                    if n_irr > 0:
                        tr7 = lambda t: t[:n_irr * 7].view(n_irr, 7).t().contiguous()
                        self.irr_c_soa, self.irr_wE_soa, self.irr_wU_soa = tr7(self.irr_c), tr7(self.irr_wE), tr7(self.irr_wU)
                        s.irr_c_soa, s.irr_wE_soa, s.irr_wU_soa = (cabi.ptr(t) for t in (self.irr_c_soa, self.irr_wE_soa,
                                                                                         self.irr_wU_soa))
            if precond is not None:
                s.coef26 = cabi.ptr(self.coef26)
                s.pc_d1, s.pc_d2, s.pc_scale = precond.widths[0], precond.widths[1], precond.scale
                s.n_pc_rows = n_pc_rows
                # uncrossed row nodes per side: they take the per-side preconditioner kernels (6 inputs instead of 26)
                assert ne < 2 ** 31
                fl = torch.zeros((ex, ey, ez), dtype=torch.int8, device=dev)
                pv = (slice(self.HX, ex - self.HX), slice(self.HY, ey - self.HY), slice(self.HZ, ez - self.hz_hi))
                fl[pv] = cs.flag.view(ex, ey, ez)[pv]
                fl = fl.reshape(-1)
                self.pc_nodes_m = torch.nonzero(fl < 0).reshape(-1).to(torch.int32).contiguous()
                self.pc_nodes_p = torch.nonzero(fl > 0).reshape(-1).to(torch.int32).contiguous()
                del fl
                s.pc_nodes_m, s.n_pc_m = cabi.ptr(self.pc_nodes_m), self.pc_nodes_m.numel()
                s.pc_nodes_p, s.n_pc_p = cabi.ptr(self.pc_nodes_p), self.pc_nodes_p.numel()
                s.pc_d[0], s.pc_d[1], s.pc_d[2] = dx, dy, dz
            self.step = s
            self.xa, self.xb = xa, xb
            # the regression/cut-cell scratch is not needed by the step
            for name in ("tri", "tri_area", "pos", "proj", "Cm", "Cp"):
                pass  # kept: tests read them back; they are O(crossed sites)

    def loss_grad_launch(self, out: Optional[torch.Tensor] = None, comm=None, finalize=None) -> torch.Tensor:
        """Enqueue loss and d loss/d params for this plan's rows (parameters must have been uploaded
        with `upload_params`).  Returns the device buffer [grad(P), loss].  With `comm` (a PeerComm) the
        final partial-row reduction is fused with the all-reduce over the ranks (SUM, psum semantics)."""
        if out is not None:
            self.step.loss_grad = cabi.ptr(out)
        target = out if out is not None else self.loss_grad
        if comm is None:
            cabi.check(cabi.lib().nbm_loss_grad_shared_f32(C.byref(self.step), cabi.stream_ptr()),
                       "nbm_loss_grad_shared_f32")
            return target
        self.step.stages = 0x1f            # everything but the reduction
        try:
            cabi.check(cabi.lib().nbm_loss_grad_shared_f32(C.byref(self.step), cabi.stream_ptr()),
                       "nbm_loss_grad_shared_f32")
        finally:
            self.step.stages = 0
        if finalize is not None:   # exchange + optax chain + staging of the next step's parameters in one kernel
            comm.reduce_allreduce_finalize(self.partials, self.step.n_partial_rows, self.n_total + 1, target, finalize)
        else:
            comm.reduce_allreduce(self.partials, self.step.n_partial_rows, self.n_total + 1, target)
        return target

    def bind_params(self, params: torch.Tensor) -> None:
        """With a preconditioner the step reads its parameters straight from the tail of the flat device vector
        [network | preconditioner] (the network part travels through `upload_params`)."""
        if self.precond is not None:
            if params.numel() != self.n_total:
                raise ValueError(f"parameter vector has {params.numel()} entries, expected {self.n_total}")
            self.step.pc_params = params.data_ptr() + 4 * self.net.n_params
            self._bound = params   # keep the storage alive

    # ---- read-backs for tests -----------------------------------------------------------------
    def rhs_rows(self) -> torch.Tensor:
        """rhs/diag of every row in lattice layout (in faces mode the irregular rows' entries live in the list)"""
        if not self.faces or self.n_irr == 0:
            return self.rhs
        out = self.rhs.clone()
        out[self.irr_point[:self.n_irr]] = self.irr_rhs[:self.n_irr]
        return out

    def point_view(self, t: torch.Tensor) -> torch.Tensor:
        """lattice-layout array -> (n_points,) in the reference's point order"""
        ex, ey, ez = self.dims
        return t.view(ex, ey, ez)[self.HX:ex - self.HX, self.HY:ey - self.HY, self.HZ:ez - self.hz_hi].reshape(-1)


def balanced_slabs(lvl, tr_gstate, world: int, device=None, list_weight: Optional[float] = None):
    """Cost-weighted x-slab boundaries for `world` devices: [(xa, xb)] contiguous, covering every x plane.

    The reference gives every device the same number of planes (data_management.py:121-130).  A plane that holds
    interface costs more than one that does not (crossed sites and irregular rows run through the list kernels), so with
    equal slabs the interface-holding ranks set the pace.  The per-plane cost here is Ny*Nz + list_weight * (number of
    nodes of the plane with a 6-neighbour on the other side of the level set), evaluated with the plan's own level set;
    boundaries sit at the quantiles of its prefix sum.  Any contiguous partition gives the same [grad, loss] when every
    device normalises by the nominal N / world (`SharedPlan(n_mean=...)`): psum of per-device means with a common
    denominator is a sum over all points (trainer.py:829-830)."""
    Nx, Ny, Nz = tr_gstate.shape()
    if world <= 1:
        return [(0, Nx)]
    if list_weight is None:
        list_weight = float(os.environ.get("NBM_BALANCE_W", "10"))
    dev = torch.device(device if device is not None else lvl.device)
    ys, zs = tr_gstate.y.to(dev), tr_gstate.z.to(dev)
    Y, Z = torch.meshgrid(ys, zs, indexing="ij")
    Y, Z = Y.reshape(-1), Z.reshape(-1)
    sign = torch.empty((Nx, Ny, Nz), dtype=torch.bool, device=dev)
    for i in range(Nx):
        pts = torch.stack((torch.full_like(Y, float(tr_gstate.x[i])), Y, Z), dim=1)
        sign[i] = (lvl(pts) >= 0).view(Ny, Nz)
    ch = torch.zeros((Nx, Ny, Nz), dtype=torch.bool, device=dev)
    for ax in range(3):
        a = sign.narrow(ax, 0, sign.shape[ax] - 1) != sign.narrow(ax, 1, sign.shape[ax] - 1)
        ch.narrow(ax, 0, sign.shape[ax] - 1).logical_or_(a)
        ch.narrow(ax, 1, sign.shape[ax] - 1).logical_or_(a)
    cost = (Ny * Nz + list_weight * ch.view(Nx, -1).sum(dim=1).double()).cpu()
    cum = torch.cumsum(cost, 0)
    total = float(cum[-1])
    cuts = [0]
    for r in range(1, world):
        # first plane index whose inclusion reaches r/world of the total cost; every slab keeps at least one plane
        k = int(torch.searchsorted(cum, torch.tensor(total * r / world, dtype=cum.dtype)).item()) + 1
        k = max(k, cuts[-1] + 1)
        k = min(k, Nx - (world - r))
        cuts.append(k)
    cuts.append(Nx)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def upload_params(net: NetShape, params: torch.Tensor) -> None:
    s = net.struct()
    cabi.check(cabi.lib().nbm_upload_params(C.byref(s), cabi.ptr(params), cabi.stream_ptr()), "nbm_upload_params")


def _point_samples(fns, xs, ys, zs, d, dev):
    """coefficient samples every row needs: mu^-/+ at the 6 face centres
    (geometric_integrations_per_point.py:533-548), k, f, g_D at the point (discretization.py:355-356,
    :403-411).  Returns ([6][n] mu_m, [6][n] mu_p, k_m, k_p, f_m, f_p, g_dir)."""
    dx, dy, dz = d
    X, Y, Z = torch.meshgrid(xs, ys, zs, indexing="ij")
    R = torch.stack((X.reshape(-1), Y.reshape(-1), Z.reshape(-1)), dim=1)
    del X, Y, Z
    n = R.shape[0]
    mu_m_faces = torch.empty(6 * n, dtype=torch.float32, device=dev)
    mu_p_faces = torch.empty(6 * n, dtype=torch.float32, device=dev)
    half = [(-dx, 0, 0), (dx, 0, 0), (0, -dy, 0), (0, dy, 0), (0, 0, -dz), (0, 0, dz)]
    for f, off in enumerate(half):
        o = torch.tensor(off, dtype=torch.float32, device=dev) * 0.5
        Rf = R + o
        mu_m_faces[f * n:(f + 1) * n] = _sample(fns.mu_m_fn, Rf)
        mu_p_faces[f * n:(f + 1) * n] = _sample(fns.mu_p_fn, Rf)
        del Rf
    out = (mu_m_faces, mu_p_faces, _sample(fns.k_m_fn, R), _sample(fns.k_p_fn, R), _sample(fns.f_m_fn, R),
           _sample(fns.f_p_fn, R), _sample(fns.dir_bc_fn, R))
    return out


class GeneralLevel:
    """Row tables of the WHOLE training grid for one cell size d (any zoom level): 7 site lattices
    (the grid displaced by 0, -+dx, -+dy, -+dz), nothing shared between points."""

    def __init__(self, lvl: LevelSet, tr_gstate, d, fns, net: NetShape, nonlinear_m: Nonlinear,
                 nonlinear_p: Nonlinear, device=None, precond: Optional[PrecondShape] = None):
        dev = torch.device(device if device is not None else lvl.device)
        self.device, self.net, self.lvl, self.d = dev, net, lvl, tuple(float(v) for v in d)
        self.nonlinear_m, self.nonlinear_p = nonlinear_m, nonlinear_p
        L = cabi.lib()
        with torch.cuda.device(dev):
            st = cabi.stream_ptr()
            Nx, Ny, Nz = tr_gstate.shape()
            self.shape = (Nx, Ny, Nz)
            N = Nx * Ny * Nz
            self.n_points = N
            self.xs, self.ys, self.zs = (a.to(dev).contiguous() for a in (tr_gstate.x, tr_gstate.y, tr_gstate.z))
            dx, dy, dz = self.d
            shifts = [(0, 0, 0), (-dx, 0, 0), (dx, 0, 0), (0, -dy, 0), (0, dy, 0), (0, 0, -dz), (0, 0, dz)]
            lat = _lattice(self.xs, self.ys, self.zs, shifts=shifts)
            self.sites = CrossedSites(lvl, lat, 7 * N, self.d, fns, dev, n_cut_sites=N,
                                      coords=(self.xs, self.ys, self.zs, shifts))
            cs = self.sites
            mu_m_faces, mu_p_faces, k_m, k_p, f_m, f_p, g_dir = _point_samples(fns, self.xs, self.ys, self.zs,
                                                                               self.d, dev)
            use_nl = (nonlinear_m.kind != NL_NONE) or (nonlinear_p.kind != NL_NONE)
            self.w = torch.zeros(7 * N, dtype=torch.float32, device=dev)
            self.rhs = torch.zeros(N, dtype=torch.float32, device=dev)
            self.nl = torch.zeros(2 * N, dtype=torch.float32, device=dev) if use_nl else None
            self.irr = torch.full((N,), -1, dtype=torch.int32, device=dev)
            cap = min(N, cs.n) + 1
            irr_count = torch.zeros(1, dtype=torch.int64, device=dev)
            irr_point = torch.zeros(cap, dtype=torch.int64, device=dev)
            irr_wE = torch.zeros(cap * 7, dtype=torch.float32, device=dev)
            irr_c = torch.full((cap * 7,), -1, dtype=torch.int32, device=dev)
            irr_nl = torch.zeros(cap, dtype=torch.uint8, device=dev)
            irr_nlw = torch.zeros(cap, dtype=torch.float32, device=dev)
            self.precond = precond
            self.coef26 = torch.zeros(26 * N, dtype=torch.float32, device=dev) if precond is not None else None
            self.Pc = torch.zeros(N, dtype=torch.float32, device=dev) if precond is not None else None
            a = cabi.Assemble()
            a.coef26 = cabi.ptr(self.coef26)
            a.pts = _lattice(self.xs, self.ys, self.zs)
            a.dx, a.dy, a.dz = dx, dy, dz
            for i, b in enumerate(lvl.bounds):
                a.bounds[i] = b
            a.shared = 0
            a.site_dims[0], a.site_dims[1], a.site_dims[2] = Nx, Ny, Nz
            a.flag, a.side, a.cidx = cabi.ptr(cs.flag), cabi.ptr(cs.side), cabi.ptr(cs.cidx)
            a.frac, a.beta_gamma = cabi.ptr(cs.frac), cabi.ptr(cs.beta_gamma)
            a.mu_m_faces, a.mu_p_faces = cabi.ptr(mu_m_faces), cabi.ptr(mu_p_faces)
            a.k_m, a.k_p, a.f_m, a.f_p, a.g_dir = (cabi.ptr(t) for t in (k_m, k_p, f_m, f_p, g_dir))
            a.w, a.rhs, a.nl, a.irr = cabi.ptr(self.w), cabi.ptr(self.rhs), cabi.ptr(self.nl), cabi.ptr(self.irr)
            a.n_out = N
            a.out_stride[0], a.out_stride[1], a.out_stride[2] = Ny * Nz, Nz, 1
            a.out_off = 0
            a.irr_capacity = cap
            a.irr_count, a.irr_point = cabi.ptr(irr_count), cabi.ptr(irr_point)
            a.irr_wE, a.irr_c, a.irr_nl, a.irr_nlw = (cabi.ptr(t) for t in (irr_wE, irr_c, irr_nl, irr_nlw))
            cabi.check(L.nbm_assemble_f32(C.byref(a), st), "nbm_assemble_f32")
            n_irr = int(irr_count.item())
            if n_irr > cap:
                raise cabi.NbmError(f"irregular-row capacity exceeded ({n_irr} > {cap})")
            self.n_irr = n_irr
            m = max(n_irr, 1)
            self.irr_wE, self.irr_c = irr_wE[:m * 7].clone(), irr_c[:m * 7].clone()
            self.irr_nl, self.irr_nlw = irr_nl[:m].clone(), irr_nlw[:m].clone()
            self.E = torch.zeros(max(cs.n, 1), dtype=torch.float32, device=dev)
            self.gE = torch.zeros(max(cs.n, 1), dtype=torch.float32, device=dev)
            sh = torch.tensor(shifts, dtype=torch.float32, device=dev)                  # (7,3)
            self.xs7 = (self.xs[None, :] + sh[:, 0:1]).contiguous()                      # fp32 add, like point[0] - dx
            self.ys7 = (self.ys[None, :] + sh[:, 1:2]).contiguous()
            self.zs7 = (self.zs[None, :] + sh[:, 2:3]).contiguous()
            self.U7 = torch.zeros(7 * N, dtype=torch.float32, device=dev)
            self.G7 = torch.zeros(7 * N, dtype=torch.float32, device=dev)
            # Zoom level 1 (cell size = half the grid spacing): p + d e_a of a point is p' - d e_a of its neighbour, so the
            # nodes + the x-, y-, z-half-offset lattices (4 network evaluations per point) replace the 7 displaced
            # lattices for whole-plane batches (data_management.py:320-326).  Padded dims (Nx+1, Ny+1, Nz+1);
            # lattice l carries "- d" in its own direction, its last entry is the "+ d" site of the last point.
            g = (float(tr_gstate.dx), float(tr_gstate.dy), float(tr_gstate.dz))
            half = all(abs(2.0 * self.d[a] - g[a]) <= 1e-6 * g[a] for a in range(3))
            self.shared4 = bool(half and os.environ.get("NBM_ZOOM1_SHARED", "1") != "0")
            if self.shared4:
                def pad(base, minus=None, plus_last=None, step=0.0):
                    a0 = base if minus is None else minus
                    last = (base[-1:] + step) if plus_last is None else plus_last
                    return torch.cat((a0, last))
                xs0, ys0, zs0 = pad(self.xs, step=g[0]), pad(self.ys, step=g[1]), pad(self.zs, step=g[2])
                self.xs4 = torch.stack((xs0, pad(self.xs, self.xs7[1], self.xs7[2][-1:]), xs0, xs0)).contiguous()
                self.ys4 = torch.stack((ys0, ys0, pad(self.ys, self.ys7[3], self.ys7[4][-1:]), ys0)).contiguous()
                self.zs4 = torch.stack((zs0, zs0, zs0, pad(self.zs, self.zs7[5], self.zs7[6][-1:]))).contiguous()
                s7 = cs.side[:7 * N].view(7, Nx, Ny, Nz)
                s4 = torch.zeros((4, Nx + 1, Ny + 1, Nz + 1), dtype=torch.uint8, device=dev)
                s4[0, :Nx, :Ny, :Nz] = s7[0]
                s4[1, :Nx, :Ny, :Nz] = s7[1]
                s4[1, Nx, :Ny, :Nz] = s7[2][Nx - 1]
                s4[2, :Nx, :Ny, :Nz] = s7[3]
                s4[2, :Nx, Ny, :Nz] = s7[4][:, Ny - 1]
                s4[3, :Nx, :Ny, :Nz] = s7[5]
                s4[3, :Nx, :Ny, Nz] = s7[6][:, :, Nz - 1]
                self.side4 = s4.reshape(-1).contiguous()
                ne4 = (Nx + 1) * (Ny + 1) * (Nz + 1)
                self.U4 = torch.zeros(4 * ne4, dtype=torch.float32, device=dev)
                self.G4 = torch.zeros(4 * ne4, dtype=torch.float32, device=dev)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            self.n_pc_rows = sms if precond is not None else 0
            rows = L.nbm_step_partial_rows() + self.n_pc_rows
            self.rows = rows
            self.n_total = net.n_params + (precond.n_params if precond is not None else 0)
            self.partials = torch.zeros(rows * (self.n_total + 1), dtype=torch.float32, device=dev)


class PointsPlan:
    """General path for the contiguous batch [p0, p1) of the flattened training points at the cell
    size of `level`."""

    def __init__(self, level: GeneralLevel, p0: int, p1: int, n_mean: Optional[int] = None, keep_rows: bool = False):
        self.level, self.p0, self.p1 = level, int(p0), int(p1)
        keep_rows = keep_rows or level.precond is not None   # the preconditioner kernels exchange the raw residuals
        self.rows = torch.zeros(level.n_points, dtype=torch.float32, device=level.device) if keep_rows else None
        self.device = level.device
        self.n_points = self.p1 - self.p0
        net, cs = level.net, level.sites
        self.loss_grad = torch.zeros(level.n_total + 1, dtype=torch.float32, device=level.device)
        self.n_total, self.precond, self.net = level.n_total, level.precond, net
        s = cabi.PointsStep()
        s.net = net.struct()
        s.nonlinear_m, s.nonlinear_p = level.nonlinear_m.kind, level.nonlinear_p.kind
        s.nl_coef_m, s.nl_coef_p = level.nonlinear_m.coef, level.nonlinear_p.coef
        s.xs, s.ys, s.zs = cabi.ptr(level.xs), cabi.ptr(level.ys), cabi.ptr(level.zs)
        s.nx, s.ny, s.nz = level.shape
        s.p0, s.p1 = self.p0, self.p1
        s.dx, s.dy, s.dz = level.d
        s.side, s.w, s.rhs, s.nl, s.irr = (cabi.ptr(t) for t in (cs.side, level.w, level.rhs, level.nl, level.irr))
        s.n_crossed, s.c_site, s.c_pos = cs.n, cabi.ptr(cs.idx), cabi.ptr(cs.pos)
        s.c_cube_side, s.B = cabi.ptr(cs.cube_side), cabi.ptr(cs.B)
        s.n_irr = level.n_irr
        s.irr_wE, s.irr_c, s.irr_nl, s.irr_nlw = (cabi.ptr(t) for t in (level.irr_wE, level.irr_c, level.irr_nl,
                                                                        level.irr_nlw))
        s.inv_n_points = 1.0 / float(n_mean if n_mean is not None else self.n_points)
        s.E, s.gE = cabi.ptr(level.E), cabi.ptr(level.gE)
        s.partials, s.n_partial_rows, s.loss_grad = cabi.ptr(level.partials), level.rows, cabi.ptr(self.loss_grad)
        s.rows = cabi.ptr(self.rows)
        s.xs7, s.ys7, s.zs7 = cabi.ptr(level.xs7), cabi.ptr(level.ys7), cabi.ptr(level.zs7)
        s.U7, s.G7 = cabi.ptr(level.U7), cabi.ptr(level.G7)
        plane = level.shape[1] * level.shape[2]
        self.shared4 = bool(getattr(level, "shared4", False) and self.p0 % plane == 0 and self.p1 % plane == 0)
        if self.shared4:
            s.xs4, s.ys4, s.zs4 = cabi.ptr(level.xs4), cabi.ptr(level.ys4), cabi.ptr(level.zs4)
            s.side4, s.U4, s.G4 = cabi.ptr(level.side4), cabi.ptr(level.U4), cabi.ptr(level.G4)
        # the crossed sites of THIS batch (the cube kernels walk them instead of scanning the level's whole list)
        self.c_live = None
        if cs.n > 0 and (self.p0 > 0 or self.p1 < level.n_points):
            if getattr(level, "_site_point_host", None) is None:       # one device read per level, shared by its batches
                level._site_point_host = (cs.idx[:cs.n] % level.n_points).cpu().numpy()
            pts = level._site_point_host
            live = np.nonzero((pts >= self.p0) & (pts < self.p1))[0].astype(np.int32)
            self.c_live = torch.from_numpy(live).to(level.device)
            s.c_live, s.n_live = cabi.ptr(self.c_live), int(live.size)
        if level.precond is not None:
            s.coef26, s.Pc = cabi.ptr(level.coef26), cabi.ptr(level.Pc)
            s.pc_d1, s.pc_d2, s.pc_scale = level.precond.widths[0], level.precond.widths[1], level.precond.scale
            s.n_pc_rows = level.n_pc_rows
        self.step = s

    def bind_params(self, params: torch.Tensor) -> None:
        """see SharedPlan.bind_params"""
        if self.precond is not None:
            if params.numel() != self.n_total:
                raise ValueError(f"parameter vector has {params.numel()} entries, expected {self.n_total}")
            self.step.pc_params = params.data_ptr() + 4 * self.net.n_params
            self._bound = params

    def loss_grad_launch(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is not None:
            self.step.loss_grad = cabi.ptr(out)
        cabi.check(cabi.lib().nbm_loss_grad_points_f32(C.byref(self.step), cabi.stream_ptr()),
                   "nbm_loss_grad_points_f32")
        return out if out is not None else self.loss_grad


class EmptyPlan:
    """A batch with no points on this device (ragged multi-device partitions): contributes zeros, but still takes part
    in the gradient exchange so that every device performs the same sequence of collectives."""

    def __init__(self, n_total: int, device):
        self.n_total, self.device = int(n_total), torch.device(device)
        self.n_points = 0
        self.loss_grad = torch.zeros(self.n_total + 1, dtype=torch.float32, device=self.device)
        self._partials = torch.zeros(self.n_total + 1, dtype=torch.float32, device=self.device)

    def bind_params(self, params: torch.Tensor) -> None:
        pass

    def loss_grad_launch(self, out: Optional[torch.Tensor] = None, comm=None, finalize=None) -> torch.Tensor:
        target = out if out is not None else self.loss_grad
        if comm is None:
            target.zero_()
        elif finalize is not None:
            comm.reduce_allreduce_finalize(self._partials, 1, self.n_total + 1, target, finalize)
        else:
            comm.reduce_allreduce(self._partials, 1, self.n_total + 1, target)
        return target
