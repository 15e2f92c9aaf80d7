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


# golden fixtures (tests/golden/<name>.npz, oracle/make_golden.py): name -> (problem, level-set interpolant)
GOLDEN_CASES = {"sphere_tri_z0": ("sphere", "trilinear"), "sphere_tri_z1": ("sphere", "trilinear"),
                "star_tri_z0": ("star", "trilinear"), "sphere_quad_z0": ("sphere", "quadratic"),
                "sphere_reaction_tri_z0": ("sphere_reaction", "trilinear"),
                "sphere_precond_tri_z0": ("sphere", "trilinear"),
                "sphere_dense_tri_z0": ("sphere", "trilinear"), "star_dense_tri_z0": ("star", "trilinear"),
                "stars_tri_z0": ("stars", "trilinear"), "dragon_quad_z0": ("dragon_like", "quadratic"),
                "pb_tri_z0": ("poisson_boltzmann", "trilinear")}


def make_case(problem, n_train, n_lvl, interp="trilinear", dtype=torch.float32, box=None, perturb_eps=1e-10,
              net=None, phi_grid=None):
    """`phi_grid`: level-set samples to use instead of evaluating `problem.phi_fn` on the lvl grid (the golden
    fixtures carry theirs, so that parity does not hinge on this machine's libm)."""
    lo, hi = box or problem.box
    tr = mesh.linspace_grid(lo, hi, [n_train] * 3 if isinstance(n_train, int) else n_train)
    lv = mesh.linspace_grid(lo, hi, [n_lvl] * 3 if isinstance(n_lvl, int) else n_lvl)
    if phi_grid is None:
        phi_grid = jnp.vmap(problem.phi_fn)(lv.R)                # float32 values on the lvl grid
    else:
        phi_grid = torch.as_tensor(phi_grid, dtype=torch.float32).reshape(-1)
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


def assert_rows_match(mine, ref, crossed, tol=1e-5, tol_crossed=1e-4):
    """rows (lhs or rhs) against reference-generated values: all rows against the problem's O(1) scale, uncrossed rows
    against their own scale at `tol`, rows of crossed cells (which carry the reference's float32 fraction noise through
    f^-V^- + f^+V^+ and the face coefficients) against their own scale at `tol_crossed`."""
    mine, ref = mine.double().cpu(), ref.double().cpu()
    assert float((mine - ref).abs().max()) < tol * max(float(ref.abs().max()), 1.0)
    un = ~crossed
    if un.any():
        assert rel_inf(mine[un], ref[un]) < tol, rel_inf(mine[un], ref[un])
    if crossed.any():
        assert rel_inf(mine[crossed], ref[crossed]) < max(tol, tol_crossed), rel_inf(mine[crossed], ref[crossed])
