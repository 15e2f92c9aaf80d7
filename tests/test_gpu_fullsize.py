"""BASELINE.json's configurations at their FULL sizes, checked through size-independent properties (the CPU oracle
evaluates 197 networks per point: it cannot follow to 64^3..256^3 in test time):

* two independent CUDA formulations of the same step agree: the shared-evaluation path (one network evaluation
  per lattice node, stencil tables) and the general per-point path at zoom 0 (7 evaluations per point, row
  weights applied per point);
* the three row-table layouts of the shared path agree (7 row weights | face coefficients | face coefficients with
  the adjoint fused into the gradient kernel);
* additivity over x-slabs (the multi-GPU partition, data_management.py:121-130);
* the gradient is the derivative of the loss: directional central differences of the CUDA loss;
* geometry: sum of the cut-cell volumes V^- = measure of Omega^- (the reference's own known-answer test,
  tests/test_geometric_integrations.py:181-182, 220-221), V^- + V^+ = cell volume, face fractions within [0, A];
* the oracle itself is compared on a bounded random SAMPLE of the full-size problem's rows.
"""
import pytest
import torch

import util
from jax_dips_b200 import mesh, problems
from jax_dips_b200 import numpy as jnp
from jax_dips_b200 import plan as nplan
from jax_dips_b200.simulation_states import PoissonSimStateFn
from oracle import nbm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def fns_of(problem):
    v = jnp.vmap
    return PoissonSimStateFn(v(problem.initial_value_fn), v(problem.dirichlet_bc_fn), v(problem.phi_fn),
                             v(problem.mu_m_fn), v(problem.mu_p_fn), v(problem.k_m_fn), v(problem.k_p_fn),
                             v(problem.f_m_fn), v(problem.f_p_fn), v(problem.alpha_fn), v(problem.beta_fn),
                             problem.nonlinear_op_m, problem.nonlinear_op_p)


def setup_case(name, n_tr, n_lvl, interp, **kw):
    P = problems.PROBLEMS[name](**kw)
    lo, hi = P.box
    tr = mesh.linspace_grid(lo, hi, [n_tr] * 3)
    lv = mesh.linspace_grid(lo, hi, [n_lvl] * 3)
    with torch.cuda.device(DEV):
        phi_grid = jnp.vmap(P.phi_fn)(lv.R.to(DEV))
        lvl = nplan.LevelSet(lv, phi_grid, interp=interp, device=DEV)
    return P, tr, lv, lvl


def shared(P, tr, lvl, xa=0, xb=None, **kw):
    return nplan.SharedPlan(lvl, tr, xa, xb if xb is not None else tr.shape()[0], fns_of(P), nplan.NetShape(),
                            nplan.Nonlinear.coerce(P.nonlinear_op_m), nplan.Nonlinear.coerce(P.nonlinear_op_p),
                            device=DEV, **kw)


def launch(pl, params):
    with torch.cuda.device(DEV):
        nplan.upload_params(nplan.NetShape(), params)
        out = pl.loss_grad_launch().clone()
        torch.cuda.synchronize()
    return out


CASES = [
    # BASELINE.json configs[1..3] (+ the 32^3 test problem of configs[0] for symmetry)
    pytest.param("sphere", 32, 128, "trilinear", {}, id="sphere-32-lvl128"),
    pytest.param("stars", 64, 128, "trilinear", {}, id="stars-64-lvl128"),
    pytest.param("dragon_like", 128, 128, "quadratic", {}, id="dragon-128-quadratic"),
]


@pytest.mark.parametrize("name,n_tr,n_lvl,interp,kw", CASES)
def test_formulations_agree_and_gradient_is_consistent(name, n_tr, n_lvl, interp, kw):
    P, tr, lv, lvl = setup_case(name, n_tr, n_lvl, interp, **kw)
    net = nplan.NetShape()
    params = O.init_params(O.NetShape(), seed=13).to(DEV)
    faces = shared(P, tr, lvl, faces=True, fused=False)
    a = launch(faces, params)
    assert torch.isfinite(a).all() and float(a[-1]) > 0
    # --- row-table layouts
    b = launch(shared(P, tr, lvl, faces=False), params)
    c = launch(shared(P, tr, lvl, faces=True, fused=True), params)
    assert util.rel_inf(b, a) < 2e-5, util.rel_inf(b, a)
    assert util.rel_inf(c, a) < 2e-5, util.rel_inf(c, a)
    # --- the per-point formulation at the native cell size
    d32 = [float(v) for v in (tr.dx, tr.dy, tr.dz)]
    level = nplan.GeneralLevel(lvl, tr, d32, fns_of(P), net, nplan.Nonlinear.coerce(P.nonlinear_op_m),
                               nplan.Nonlinear.coerce(P.nonlinear_op_p), device=DEV)
    g = launch(nplan.PointsPlan(level, 0, tr.num_points()), params)
    assert abs(float(g[-1]) - float(a[-1])) / float(a[-1]) < 2e-5
    assert util.rel_inf(g[:-1], a[:-1]) < 5e-5, util.rel_inf(g[:-1], a[:-1])
    del level
    # --- x-slab additivity (what the multi-GPU partition relies on)
    Nx = tr.shape()[0]
    acc = torch.zeros_like(a)
    for (xa, xb) in ((0, Nx // 4), (Nx // 4, Nx // 2 + 2), (Nx // 2 + 2, Nx)):
        pl = shared(P, tr, lvl, xa, xb)
        acc += launch(pl, params) * (pl.n_points / faces.n_points)
        del pl
    assert util.rel_inf(acc, a) < 2e-5, util.rel_inf(acc, a)
    # --- directional derivative of the loss (central differences in float32: a loose but size-independent check)
    gen = torch.Generator().manual_seed(1)
    for _ in range(2):
        d = torch.randn(params.numel(), generator=gen).to(DEV)
        d = d / d.norm()
        eps = 2e-3
        lp = float(launch(faces, params + eps * d)[-1])
        lm = float(launch(faces, params - eps * d)[-1])
        fd = (lp - lm) / (2 * eps)
        an = float((a[:-1] * d).sum())
        assert abs(fd - an) < 2e-2 * max(abs(an), float(a[:-1].norm()) * 0.05), (fd, an)


@pytest.mark.parametrize("name,n_tr,n_lvl,interp,kw", CASES)
def test_geometry_invariants_and_oracle_sample(name, n_tr, n_lvl, interp, kw):
    P, tr, lv, lvl = setup_case(name, n_tr, n_lvl, interp, **kw)
    pl = shared(P, tr, lvl)
    cs = pl.sites
    dx, dy, dz = (float(v) for v in (tr.dx, tr.dy, tr.dz))
    vol = dx * dy * dz
    frac = cs.frac.view(-1, 14)[:cs.n]
    assert cs.n > 100
    assert float((frac[:, 12] + frac[:, 13] - vol).abs().max()) < 1e-5 * vol
    areas = torch.tensor([dy * dz, dy * dz, dx * dz, dx * dz, dx * dy, dx * dy], device=DEV)
    am, ap = frac[:, 0:12:2], frac[:, 1:12:2]
    assert float(am.min()) >= 0 and float(ap.min()) >= 0
    # (A^- is a sum of float32 triangle areas formed from ABSOLUTE vertex coordinates, like the reference's: its
    #  rounding error relative to the face grows like eps / h - 3e-5 at 16^3 (test_reference_noise_floor), ~1e-4 at
    #  128^3 - and A^- may exceed the face by that much; A^+ = max(0, A - A^-))
    tol_a = 4e-6 / min(dx, dy, dz) * float(areas.max())
    assert float((am - areas).max()) < tol_a and float((am + ap - areas).abs().max()) < tol_a
    # measure of Omega^-: cut cells + full cells of the minus side, against a 2x finer sign count of the same interpolant
    flag = pl.point_view(cs.flag)
    cidx = pl.point_view(cs.cidx).long()
    crossed = flag == 0
    v_minus = float(frac[cidx[crossed], 12].double().sum()) + float((flag < 0).sum()) * vol
    lo, hi = P.box
    fine = mesh.linspace_grid(lo, hi, [2 * n_tr - 1] * 3)
    with torch.cuda.device(DEV):
        phi_f = lvl(fine.R.to(DEV))
    fv = float(fine.dx * fine.dy * fine.dz)
    v_count = float((phi_f < 0).sum()) * fv
    assert abs(v_minus - v_count) < 0.05 * v_count + 4 * vol, (v_minus, v_count)
    # the oracle on a bounded random sample of the rows
    dt = torch.float64
    trc, lvc, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, interp, dt)
    params = O.init_params(oprob.shape, seed=13, dtype=dt)
    launch(pl, params.float().to(DEV))
    gen = torch.Generator().manual_seed(3)
    idx_c = crossed.reshape(-1).nonzero().reshape(-1).cpu()
    pick = torch.cat((idx_c[torch.randperm(idx_c.numel(), generator=gen)[:96]],
                      torch.randint(0, tr.num_points(), (96,), generator=gen)))
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    lhs_o, rhs_o = O.compute_Ax_and_b(params, tr.R.to(dt)[pick], *d, oprob)
    rhs_k = pl.point_view(pl.rhs_rows()).cpu()[pick]
    lhs_k = pl.point_view(pl.R).cpu()[pick] + rhs_k
    scale = float(torch.maximum(lhs_o.abs().max(), rhs_o.abs().max()))
    assert float((lhs_k.double() - lhs_o).abs().max()) < 1e-5 * scale
    assert float((rhs_k.double() - rhs_o).abs().max()) < 1e-5 * scale


def test_poisson_boltzmann_256_layouts_slabs_and_derivative():
    """BASELINE.json configs[3]: nonlinear Poisson-Boltzmann, synthetic multi-atom surface, 256^3 (the per-GPU slab
    of the 2/4/8-GPU runs is a subset of this grid)."""
    P, tr, lv, lvl = setup_case("poisson_boltzmann", 256, 128, "trilinear", n_atoms=200)
    params = O.init_params(O.NetShape(), seed=13).to(DEV) * 2.0
    whole = shared(P, tr, lvl)
    assert whole.faces and whole.nl is not None
    a = launch(whole, params)
    assert torch.isfinite(a).all()
    b = launch(shared(P, tr, lvl, faces=False), params)
    assert util.rel_inf(b, a) < 2e-5, util.rel_inf(b, a)
    acc = torch.zeros_like(a)
    for (xa, xb) in ((0, 128), (128, 256)):
        pl = shared(P, tr, lvl, xa, xb)
        acc += launch(pl, params) * (pl.n_points / whole.n_points)
        del pl
    assert util.rel_inf(acc, a) < 2e-5
    gen = torch.Generator().manual_seed(2)
    d = torch.randn(params.numel(), generator=gen).to(DEV)
    d = d / d.norm()
    eps = 2e-3
    fd = (float(launch(whole, params + eps * d)[-1]) - float(launch(whole, params - eps * d)[-1])) / (2 * eps)
    an = float((a[:-1] * d).sum())
    assert abs(fd - an) < 2e-2 * max(abs(an), float(a[:-1].norm()) * 0.05), (fd, an)
