"""Problem definitions for tests and benchmarks: the same per-point callables the reference's
drivers define, written against `jax_dips_b200.numpy` (torch) instead of `jax.numpy`.

* `sphere()`, `star()`, `no_jump()`  : tests/confs/experiment_configs.py:21-186, :196-412, :420-555
  (analytic exact solutions -> the only end-to-end accuracy pin the reference has)
* `stars()`                          : examples/stars/solve_stars.py:103-282 (4x4x4 stars, mu 1/80)
* `dragon_like()`                    : examples/dragon/coefficients.py with a synthetic SDF
* `poisson_boltzmann()`              : examples/benchmark_LPBE + examples/biomolecules (kappa^2 sinh u)

Each returns a `Problem` whose callables take one point `r` (`r[0], r[1], r[2]`) and are batched
by `jax_dips_b200.numpy.vmap` (called once with the (3, n) view of the batch).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional

import numpy as np
import torch

from . import numpy as jnp
from .plan import Nonlinear


@dataclasses.dataclass
class Problem:
    initial_value_fn: Callable
    dirichlet_bc_fn: Callable
    phi_fn: Callable
    mu_m_fn: Callable
    mu_p_fn: Callable
    k_m_fn: Callable
    k_p_fn: Callable
    f_m_fn: Callable
    f_p_fn: Callable
    alpha_fn: Callable
    beta_fn: Callable
    exact_sol_m_fn: Optional[Callable] = None
    exact_sol_p_fn: Optional[Callable] = None
    nonlinear_op_m: Optional[Nonlinear] = None
    nonlinear_op_p: Optional[Nonlinear] = None
    box: tuple = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
    name: str = ""

    def setup_args(self):
        """positional arguments of trainer.setup (trainer.py:980-994)"""
        return (self.initial_value_fn, self.dirichlet_bc_fn, self.phi_fn, self.mu_m_fn, self.mu_p_fn,
                self.k_m_fn, self.k_p_fn, self.f_m_fn, self.f_p_fn, self.alpha_fn, self.beta_fn)


def grad(fn: Callable) -> Callable:
    """Per-point gradient of a scalar callable (the role jax.grad plays in experiment_configs.py):
    works on the (3, n) batched view through torch autograd."""
    def g(r):
        with torch.enable_grad():
            rr = r.detach().clone().requires_grad_(True)
            out = fn(rr)
            if not isinstance(out, torch.Tensor) or not out.requires_grad:
                return torch.zeros_like(rr)
            gr, = torch.autograd.grad(out.sum(), rr)
        return gr
    return g


def perturb_level_set_fn(phi_fn: Callable) -> Callable:
    """geometry/level_set.py:34-48"""
    def perturbed(r):
        lvl = phi_fn(r)
        return lvl + jnp.sign(jnp.sign(lvl) - 0.5) * 1.0e-10
    return perturbed


def _jump_beta(mu_m_fn, mu_p_fn, u_m_fn, u_p_fn, phi_fn, sign=-1.0, nan_safe=False):
    """beta = sign * (mu_p grad u_p - mu_m grad u_m) . grad phi   (experiment_configs.py:85-96)"""
    gp, gm, gn = grad(u_p_fn), grad(u_m_fn), grad(phi_fn)

    def beta_fn(r):
        v = mu_p_fn(r) * gp(r) - mu_m_fn(r) * gm(r)
        out = (v * gn(r)).sum(dim=0) * sign
        return torch.nan_to_num(out) if nan_safe else out
    return beta_fn


def sphere(center=(0.0, 0.0, 0.0), radius: float = 0.5) -> Problem:
    """experiment_configs.py:21-186 (centre and radius are the reference's (0,0,0), 0.5 by default; other values give
    test variants, e.g. an interface that reaches the box boundary)"""
    cx, cy, cz = center
    def exact_sol_m_fn(r): return jnp.exp(r[2])
    def exact_sol_p_fn(r): return jnp.sin(r[1]) * jnp.cos(r[0])
    def unperturbed_phi_fn(r): return jnp.sqrt((r[0] - cx) ** 2 + (r[1] - cy) ** 2 + (r[2] - cz) ** 2) - radius
    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)
    def mu_m_fn(r): return r[1] * r[1] * jnp.log(r[0] + 2.0) + 4.0
    def mu_p_fn(r): return jnp.exp(-1.0 * r[2])
    def alpha_fn(r): return exact_sol_p_fn(r) - exact_sol_m_fn(r)
    beta_fn = _jump_beta(mu_m_fn, mu_p_fn, exact_sol_m_fn, exact_sol_p_fn, phi_fn, sign=-1.0)
    def f_m_fn(r): return -1.0 * jnp.exp(r[2]) * (r[1] * r[1] * jnp.log(r[0] + 2) + 4)
    def f_p_fn(r): return 2.0 * jnp.exp(-1.0 * r[2]) * jnp.cos(r[0]) * jnp.sin(r[1])
    zero = lambda r: 0.0
    return Problem(zero, exact_sol_p_fn, phi_fn, mu_m_fn, mu_p_fn, zero, zero, f_m_fn, f_p_fn, alpha_fn, beta_fn,
                   exact_sol_m_fn, exact_sol_p_fn, name="sphere")


def star() -> Problem:
    """experiment_configs.py:196-412 (Guittet 2015 sec. 4.6)"""
    def exact_sol_m_fn(r): return jnp.sin(2.0 * r[0]) * jnp.cos(2.0 * r[1]) * jnp.exp(r[2])

    def exact_sol_p_fn(r):
        yx3 = (r[1] - r[0]) / 3.0
        return (16.0 * yx3 ** 5 - 20.0 * yx3 ** 3 + 5.0 * yx3) * jnp.log(r[0] + r[1] + 3) * jnp.cos(r[2])

    def unperturbed_phi_fn(r):
        x, y, z = r[0], r[1], r[2]
        r0 = 0.483
        th = jnp.arctan2(y, x)
        core = 0.1 * jnp.cos(3.0 * (th - 0.5)) + (-0.1) * jnp.cos(4.0 * (th - 1.8)) + 0.15 * jnp.cos(7.0 * (th - 0.0))
        phi_ = jnp.sqrt(x ** 2 + y ** 2 + z ** 2)
        phi_ = phi_ + -1.0 * r0 * (1.0 + ((x ** 2 + y ** 2) / (x ** 2 + y ** 2 + z ** 2)) ** 2 * core)
        return torch.where(torch.isnan(phi_), -r0 * core, phi_)

    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)

    def mu_m_fn(r):
        return 10.0 * (1 + 0.2 * jnp.cos(2 * jnp.pi * (r[0] + r[1])) * jnp.sin(2 * jnp.pi * (r[0] - r[1])) * jnp.cos(r[2]))

    def mu_p_fn(r): return 1.0 + 0.0 * r[0]
    def alpha_fn(r): return exact_sol_p_fn(r) - exact_sol_m_fn(r)
    beta_fn = _jump_beta(mu_m_fn, mu_p_fn, exact_sol_m_fn, exact_sol_p_fn, phi_fn, sign=-1.0, nan_safe=True)

    def f_m_fn(r):
        x, y, z = r[0], r[1], r[2]
        return (-1.0 * mu_m_fn(r) * (-7.0 * jnp.sin(2.0 * x) * jnp.cos(2.0 * y) * jnp.exp(z))
                + -4 * jnp.pi * jnp.cos(z) * jnp.cos(4 * jnp.pi * x) * 2 * jnp.cos(2 * x) * jnp.cos(2 * y) * jnp.exp(z)
                + -4 * jnp.pi * jnp.cos(z) * jnp.cos(4 * jnp.pi * y) * (-2) * jnp.sin(2 * x) * jnp.sin(2 * y) * jnp.exp(z)
                + 2 * jnp.cos(2 * jnp.pi * (x + y)) * jnp.sin(2 * jnp.pi * (x - y)) * jnp.sin(z) * jnp.sin(2 * x)
                * jnp.cos(2 * y) * jnp.exp(z))

    def f_p_fn(r):
        x, y, z = r[0], r[1], r[2]
        q = (y - x) / 3
        return -1.0 * ((16 * q ** 5 - 20 * q ** 3 + 5 * q) * (-2) * jnp.cos(z) / (x + y + 3) ** 2
                       + 2 * (16 * 5 * 4 * (1.0 / 9.0) * q ** 3 - 20 * 3 * 2 * (1.0 / 9.0) * q) * jnp.log(x + y + 3) * jnp.cos(z)
                       + -1 * (16 * q ** 5 - 20 * q ** 3 + 5 * q) * jnp.log(x + y + 3) * jnp.cos(z))

    zero = lambda r: 0.0
    return Problem(zero, exact_sol_p_fn, phi_fn, mu_m_fn, mu_p_fn, zero, zero, f_m_fn, f_p_fn, alpha_fn, beta_fn,
                   exact_sol_m_fn, exact_sol_p_fn, name="star")


def no_jump() -> Problem:
    """experiment_configs.py:420-555 (phi > 0 everywhere: no interface)"""
    def exact(r): return jnp.sin(r[1]) * jnp.cos(r[0]) * jnp.cos(r[2])
    def unperturbed_phi_fn(r): return jnp.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2) + 0.5
    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)
    one = lambda r: 1.0 + 0.0 * r[0]
    zero = lambda r: 0.0
    def f_p_fn(r): return 3.0 * jnp.sin(r[1]) * jnp.cos(r[0]) * jnp.cos(r[2])
    return Problem(lambda r: r[1], exact, phi_fn, one, one, zero, zero, zero, f_p_fn, zero, zero, exact, exact,
                   name="no_jump")


# ---------------------------------------------------------------------------------------------
# benchmark geometries (synthetic, seeded with numpy default_rng because the reference's
# jax PRNGKey streams cannot be reproduced without jax)
# ---------------------------------------------------------------------------------------------
def stars(n_side: int = 4, scale: float = 0.35, variable_mu: bool = True, seed: int = 0) -> Problem:
    """examples/stars/solve_stars.py:103-282: n_side^3 star-shaped inclusions in [-1,1]^3, mu^+ = 80,
    mu^- = 1 (variable_mu=False, as in the example) or the variable form of experiment_configs.star
    scaled to 1 ("variable mu/k jumps" of BASELINE.json); k = 0, alpha = beta = 0, g_D = 0,
    f^-/f^+ as solve_stars.py:244-282.  Per-star phase angles ~ N(0,1)*pi from numpy default_rng(seed)
    (the example draws them from jax PRNGKey(0), which cannot be reproduced without jax)."""
    rng = np.random.default_rng(seed)
    r0 = 0.483 * scale
    re = 0.911 * scale
    cen = np.linspace(-1 + 1.15 * re, 1 - 1.15 * re, n_side).astype(np.float32)
    # jnp.meshgrid default indexing="xy"
    Xc, Yc, Zc = np.meshgrid(cen, cen, cen)
    centres = np.column_stack((Xc.reshape(-1), Yc.reshape(-1), Zc.reshape(-1))).astype(np.float32)
    angles = (rng.standard_normal((centres.shape[0], 3)) * math.pi).astype(np.float32)
    betas = (0.1 * scale, -0.1 * scale, 0.15 * scale)
    ns = (3.0, 4.0, 7.0)

    def unperturbed_phi_fn(r):
        x, y, z = r[0], r[1], r[2]
        phi_ = torch.full_like(x, 1e9)
        for s in range(centres.shape[0]):
            xc, yc, zc = (float(v) for v in centres[s])
            th = torch.atan2(y - yc, x - xc)
            core = sum(betas[m] * torch.cos(ns[m] * (th - float(angles[s, m]))) for m in range(3))
            rho2 = (x - xc) ** 2 + (y - yc) ** 2
            rr2 = rho2 + (z - zc) ** 2
            cand = torch.sqrt(rr2) - 1.0 * r0 * (1.0 + (rho2 / rr2) ** 2 * core)
            phi_ = torch.minimum(phi_, cand)
            phi_ = torch.where(torch.isnan(phi_), -r0 * core, phi_)
        return phi_

    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)
    if variable_mu:
        def mu_m_fn(r):
            return 1.0 * (1 + 0.2 * jnp.cos(2 * jnp.pi * (r[0] + r[1])) * jnp.sin(2 * jnp.pi * (r[0] - r[1])) * jnp.cos(r[2]))
    else:
        def mu_m_fn(r): return 1.0 + 0.0 * r[0]
    def mu_p_fn(r): return 80.0 + 0.0 * r[0]
    zero = lambda r: 0.0
    base = star()
    def f_m_fn(r):
        x, y, z = r[0], r[1], r[2]
        return (-1.0 * mu_m_fn(r) * (-7.0 * jnp.sin(2.0 * x) * jnp.cos(2.0 * y) * jnp.exp(z))
                + -4 * jnp.pi * jnp.cos(z) * jnp.cos(4 * jnp.pi * x) * 2 * jnp.cos(2 * x) * jnp.cos(2 * y) * jnp.exp(z)
                + -4 * jnp.pi * jnp.cos(z) * jnp.cos(4 * jnp.pi * y) * (-2) * jnp.sin(2 * x) * jnp.sin(2 * y) * jnp.exp(z)
                + 2 * jnp.cos(2 * jnp.pi * (x + y)) * jnp.sin(2 * jnp.pi * (x - y)) * jnp.sin(z) * jnp.sin(2 * x)
                * jnp.cos(2 * y) * jnp.exp(z))
    return Problem(zero, zero, phi_fn, mu_m_fn, mu_p_fn, zero, zero, f_m_fn, base.f_p_fn, zero, zero, name="stars")


def _smooth_union(vals, k):
    # smooth minimum of signed distances (polynomial smooth-min)
    out = vals[0]
    for v in vals[1:]:
        h = torch.clamp(0.5 + 0.5 * (v - out) / k, 0.0, 1.0)
        out = v * (1 - h) + out * h - k * h * (1 - h)
    return out


def dragon_like(n_blobs: int = 50, seed: int = 1) -> Problem:
    """examples/dragon (solve_dragon.py:177, coefficients.py): an irregular SDF interface given on a
    grid and read through the non-oscillatory quadratic interpolant; mu^- = 1, mu^+ = 2,
    alpha = 0.1, beta = 1, k = 0, f^- = sin 40 pi x cos 40 pi y sin 40 pi z, f^+ = 0, g_D = 0.2.
    The dragon mesh itself is not in the reference tree: the SDF here is a smooth union of random
    ellipsoids."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.55, 0.55, size=(n_blobs, 3)).astype(np.float32)
    rad = rng.uniform(0.06, 0.16, size=(n_blobs, 3)).astype(np.float32)

    def unperturbed_phi_fn(r):
        C = torch.as_tensor(c, device=r[0].device)
        Rr = torch.as_tensor(rad, device=r[0].device)
        vals = []
        for s in range(C.shape[0]):
            q = torch.sqrt(((r[0] - C[s, 0]) / Rr[s, 0]) ** 2 + ((r[1] - C[s, 1]) / Rr[s, 1]) ** 2
                           + ((r[2] - C[s, 2]) / Rr[s, 2]) ** 2)
            vals.append((q - 1.0) * float(Rr[s].min()))
        return _smooth_union(vals, 0.05)

    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)
    c1 = lambda v: (lambda r: v + 0.0 * r[0])
    def f_m_fn(r): return jnp.sin(40 * jnp.pi * r[0]) * jnp.cos(40 * jnp.pi * r[1]) * jnp.sin(40 * jnp.pi * r[2])
    zero = lambda r: 0.0
    return Problem(zero, c1(0.2), phi_fn, c1(1.0), c1(2.0), zero, zero, f_m_fn, zero, c1(0.1), c1(1.0),
                   name="dragon_like")


def poisson_boltzmann(n_atoms: int = 200, seed: int = 2, half_width: float = 2.5) -> Problem:
    """examples/benchmark_LPBE/coefficients.py + examples/biomolecules/coefficients.py:126-131:
    union-of-spheres molecular surface, mu^- = 2, mu^+ = 80, k^- = 0, k^+ = 80 kappa^2 (linear part
    off here: the sinh carries it), N^+(u) = kappa_bar^2 sinh(u), alpha = Coulomb sum g,
    beta = mu^- grad g . n, g_D = Debye-Hueckel-like far field, f = 0."""
    rng = np.random.default_rng(seed)
    cen = rng.uniform(-0.9, 0.9, size=(n_atoms, 3)).astype(np.float32)
    sig = rng.uniform(0.18, 0.36, size=(n_atoms,)).astype(np.float32)
    chg = rng.choice([-1.0, 1.0], size=(n_atoms,)).astype(np.float32) * rng.uniform(0.2, 1.0, size=(n_atoms,)).astype(np.float32)
    kappa_sq = 8.4869e-5 / 80.0 * 1.0e4  # reduced units, scaled so the sinh term is visible at this box size
    eps_m, eps_p = 2.0, 80.0

    def unperturbed_phi_fn(r):
        C = torch.as_tensor(cen, device=r[0].device)
        S = torch.as_tensor(sig, device=r[0].device)
        best = None
        for s in range(C.shape[0]):
            v = torch.sqrt((r[0] - C[s, 0]) ** 2 + (r[1] - C[s, 1]) ** 2 + (r[2] - C[s, 2]) ** 2) - S[s]
            best = v if best is None else torch.minimum(best, v)
        return best

    phi_fn = perturb_level_set_fn(unperturbed_phi_fn)

    def coulomb(r):
        C = torch.as_tensor(cen, device=r[0].device)
        Q = torch.as_tensor(chg, device=r[0].device)
        g = 0.0
        for s in range(C.shape[0]):
            dist = torch.sqrt((r[0] - C[s, 0]) ** 2 + (r[1] - C[s, 1]) ** 2 + (r[2] - C[s, 2]) ** 2 + 1e-4)
            g = g + Q[s] / (eps_m * dist)
        return g * 0.05

    gn, gg = grad(unperturbed_phi_fn), grad(coulomb)

    def beta_fn(r):
        n = gn(r)
        n = n / torch.sqrt((n * n).sum(dim=0) + 1e-30)
        return eps_m * (gg(r) * n).sum(dim=0)

    def g_dir(r):
        return coulomb(r) * (eps_m / eps_p) * jnp.exp(-math.sqrt(kappa_sq) * jnp.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2))

    c1 = lambda v: (lambda r: v + 0.0 * r[0])
    zero = lambda r: 0.0
    hw = half_width
    return Problem(zero, g_dir, phi_fn, c1(eps_m), c1(eps_p), zero, zero, zero, zero, coulomb, beta_fn,
                   nonlinear_op_m=None, nonlinear_op_p=Nonlinear.sinh(eps_p * kappa_sq),
                   box=((-hw, -hw, -hw), (hw, hw, hw)), name="poisson_boltzmann")


def sphere_reaction() -> Problem:
    """the sphere problem with variable reaction coefficients k^-, k^+ != 0 and a nonlinear operator on both sides
    (N = c sinh u): every term of the row (discretization.py:366-386) is present.  No analytic solution is attached."""
    P = sphere()
    P.k_m_fn = lambda r: 0.7 + 0.2 * r[0]
    P.k_p_fn = lambda r: 1.3 + 0.1 * r[1]
    P.nonlinear_op_m = Nonlinear.sinh(40.0)
    P.nonlinear_op_p = Nonlinear.sinh(150.0)
    P.exact_sol_m_fn = P.exact_sol_p_fn = None
    P.name = "sphere_reaction"
    return P


def sphere_at_boundary() -> Problem:
    """the sphere problem with the interface pushed against the x+ face of the box: crossed cells whose
    27-point regression cube leaves the box (the halo layer of the lattice)"""
    P = sphere(center=(0.62, 0.1, -0.05), radius=0.36)
    P.name = "sphere_at_boundary"
    return P


PROBLEMS = {"sphere": sphere, "sphere_reaction": sphere_reaction, "sphere_at_boundary": sphere_at_boundary, "star": star, "no_jump": no_jump, "stars": stars, "dragon_like": dragon_like,
            "poisson_boltzmann": poisson_boltzmann}
