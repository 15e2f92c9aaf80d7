"""Shared helpers for the parity tests: build the ORACLE's view of a problem (batched CPU callables,
level set = the reference's interpolant of the gridded phi) next to the product's view."""
import torch

from jax_dips_b200 import numpy as jnp
from jax_dips_b200 import mesh
from oracle import nbm_oracle as O


def batched(fn, dtype):
    v = jnp.vmap(fn)

    def f(R):
        return v(R.to(dtype)).to(dtype)
    return f


def make_case(problem, n_train, n_lvl, interp="trilinear", dtype=torch.float32, box=None, perturb_eps=1e-10,
              net=None):
    lo, hi = box or problem.box
    tr = mesh.linspace_grid(lo, hi, [n_train] * 3 if isinstance(n_train, int) else n_train)
    lv = mesh.linspace_grid(lo, hi, [n_lvl] * 3 if isinstance(n_lvl, int) else n_lvl)
    phi_grid = jnp.vmap(problem.phi_fn)(lv.R)                    # float32 values on the lvl grid
    og = O.OracleGrid(lv.x, lv.y, lv.z)
    mk = O.multilinear_interpolation if interp == "trilinear" else O.nonoscillatory_quadratic_interpolation
    interp32 = mk(phi_grid, og)

    def phi_fn(R):                                               # always evaluated in float32
        v = interp32(R.to(torch.float32))
        if perturb_eps:
            v = v + O.sign_pm_fn(v) * perturb_eps
        return v.to(R.dtype)

    bounds = tuple(t.to(dtype) for t in (lv.xmin(), lv.xmax(), lv.ymin(), lv.ymax(), lv.zmin(), lv.zmax()))
    oprob = O.OracleProblem(
        phi_fn, batched(problem.mu_m_fn, dtype), batched(problem.mu_p_fn, dtype), batched(problem.k_m_fn, dtype),
        batched(problem.k_p_fn, dtype), batched(problem.f_m_fn, dtype), batched(problem.f_p_fn, dtype),
        batched(problem.alpha_fn, dtype), batched(problem.beta_fn, dtype), batched(problem.dirichlet_bc_fn, dtype),
        bounds, shape=net or O.NetShape(),
        nonlinear_op_m=problem.nonlinear_op_m, nonlinear_op_p=problem.nonlinear_op_p)
    return tr, lv, phi_grid, oprob


def rel_inf(a, b):
    """normwise relative error  max|a-b| / max|b|"""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
