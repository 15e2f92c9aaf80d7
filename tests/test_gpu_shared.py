"""GPU parity of the shared-evaluation path (cell size == grid spacing) against the CPU oracle.

Tolerances (BASELINE.json north_star): cut-cell fractions and residual rows 1e-5, loss and
gradients 1e-4.  "Relative" is normwise: max|a-b| / max|b| over the array (rows of one problem share
one scale: the Dirichlet rows are O(1)); per-element relative error is meaningless where the
finite-volume row cancels to ~h^2.  The oracle runs in float64 (the exact-arithmetic value of the
reference's algorithm) on float32 level-set samples, so that every sign decision is identical.
"""
import ctypes as C

import pytest
import torch

import util
from jax_dips_b200 import _cabi as cabi
from jax_dips_b200 import numpy as jnp
from jax_dips_b200 import plan as nplan
from jax_dips_b200 import problems
from jax_dips_b200.simulation_states import PoissonSimStateFn
from oracle import nbm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_FRAC, TOL_ROW, TOL_LOSS = 1e-5, 1e-5, 1e-4


def fns_of(problem):
    v = jnp.vmap
    return PoissonSimStateFn(v(problem.initial_value_fn), v(problem.dirichlet_bc_fn), v(problem.phi_fn),
                             v(problem.mu_m_fn), v(problem.mu_p_fn), v(problem.k_m_fn), v(problem.k_p_fn),
                             v(problem.f_m_fn), v(problem.f_p_fn), v(problem.alpha_fn), v(problem.beta_fn),
                             problem.nonlinear_op_m, problem.nonlinear_op_p)


def build(problem, n_train, n_lvl, interp="trilinear", net=None, xa=0, xb=None, faces=None, fused=None, stencil_tma=None,
          phi_grid=None, stash=None, overlap_lists=None):
    tr, lv, phi_grid, oprob = util.make_case(problem, n_train, n_lvl, interp, torch.float64, net=net, phi_grid=phi_grid)
    lvl = nplan.LevelSet(lv, phi_grid, interp=interp, perturb_eps=1e-10, device=DEV)
    shape = nplan.NetShape(oprob.shape.Lp, oprob.shape.Hp, oprob.shape.Lm, oprob.shape.Hm)
    pl = nplan.SharedPlan(lvl, tr, xa, xb if xb is not None else tr.shape()[0], fns_of(problem), shape,
                          nplan.Nonlinear.coerce(problem.nonlinear_op_m),
                          nplan.Nonlinear.coerce(problem.nonlinear_op_p), device=DEV, faces=faces, fused=fused,
                          stencil_tma=stencil_tma, stash=stash, overlap_lists=overlap_lists)
    return tr, lv, lvl, oprob, pl, shape


@pytest.mark.parametrize("interp", ["trilinear", "quadratic"])
def test_phi_interp_matches_oracle(interp):
    P = problems.sphere()
    tr, lv, phi_grid, oprob = util.make_case(P, 8, 20, interp, torch.float32, perturb_eps=0.0)
    lvl = nplan.LevelSet(lv, phi_grid, interp=interp, perturb_eps=0.0, device=DEV)
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(20000, 3, generator=g) * 2.6 - 1.3).float()   # includes points outside the box
    got = lvl(pts.to(DEV)).cpu()
    want = oprob.phi_fn(pts)
    assert torch.equal(got, want) or util.rel_inf(got, want) < 1e-6


@pytest.mark.parametrize("name", ["sphere", "star"])
def test_classification_and_cut_cells(name):
    P = problems.PROBLEMS[name]()
    tr, lv, lvl, oprob, pl, _ = build(P, 16, 32)
    dt = torch.float64
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    R = tr.R.to(dt)
    flag_o = O.is_cell_crossed(R, *d, oprob.phi_fn)
    flag_k = pl.point_view(pl.sites.flag).cpu().to(dt)
    assert torch.equal(flag_k, flag_o), f"{(flag_k != flag_o).sum()} crossing flags differ"
    # fractions of crossed cells against the exact-arithmetic oracle (mu = 1 -> columns 14.. are areas)
    one = lambda X: torch.ones(X.shape[0], dtype=dt)
    crossed = (flag_o == 0).nonzero().reshape(-1)
    assert crossed.numel() > 50
    co = O.cell_faces_areas_values(R[crossed], *d, oprob.phi_fn, one, one)
    cidx = pl.point_view(pl.sites.cidx).cpu()[crossed].long()
    frac = pl.sites.frac.view(-1, 14).cpu().to(dt)[cidx]
    vol = float(d[0] * d[1] * d[2])
    area = float(d[1] * d[2])
    err_v = (frac[:, 12:14] - co[:, 12:14]).abs().max() / vol
    err_a = (frac[:, 0:12] - co[:, 14:26]).abs().max() / area
    assert err_v < TOL_FRAC and err_a < TOL_FRAC, (float(err_v), float(err_a))
    # Gamma integral of beta
    bg = pl.sites.beta_gamma.cpu().to(dt)[cidx]
    bo = O.integrate_over_interface(R[crossed], *d, oprob.phi_fn, oprob.beta_fn)
    # (sphere: beta is NaN at the origin, where the reference evaluates it for the unused triangle
    # slots -> the whole integral is NaN there, on both sides)
    assert torch.equal(torch.isnan(bg), torch.isnan(bo))
    fin = ~torch.isnan(bo)
    if fin.any():
        assert util.rel_inf(bg[fin], bo[fin]) < TOL_FRAC
    # and with a beta that is regular at the origin the integral itself is checked
    tri = pl.sites.tri.view(-1, 10, 3, 3)[cidx.to(pl.sites.tri.device)]
    area = pl.sites.tri_area.view(-1, 10)[cidx.to(pl.sites.tri.device)]
    smooth = lambda X: torch.cos(X[:, 0]) * torch.exp(X[:, 1]) + X[:, 2]
    got = (area.double().cpu() * smooth(tri.reshape(-1, 3).double().cpu()).view(-1, 10, 3).mean(-1)).sum(-1)
    want = O.integrate_over_interface(R[crossed], *d, oprob.phi_fn, smooth)
    assert util.rel_inf(got, want) < TOL_FRAC


# n = 15 gives an odd (y,z) plane: the scalar stencil kernels; the others take the 16-byte vector path.
# faces: one coefficient per cell face + 1/diag (the default on even grids) against the 7-weight row table.
# fused: the dense adjoint stencil evaluated inside the gradient kernel from TMA-staged tables (opt-in)
# against the separate adjoint pass (default).
@pytest.mark.parametrize("name,n,nl,faces,fused", [
    ("sphere", 16, 32, True, True), ("star", 16, 32, True, True), ("no_jump", 12, 16, True, False),
    ("sphere", 24, 24, True, False), ("star", 32, 32, True, True), ("star", 15, 32, False, False),
    ("sphere", 16, 32, True, False), ("star", 16, 32, True, False),
    ("sphere", 16, 32, False, False), ("star", 16, 32, False, False),
    # variable reaction coefficients k^-, k^+ (the kv table of the faces layout) + sinh on both sides
    ("sphere_reaction", 16, 32, True, False), ("sphere_reaction", 16, 32, False, False)])
def test_rows_loss_and_gradient(name, n, nl, faces, fused):
    P = problems.PROBLEMS[name]()
    tr, lv, lvl, oprob, pl, shape = build(P, n, nl, faces=faces, fused=fused)
    assert pl.faces == faces and pl.fused == fused
    dt = torch.float64
    params = O.init_params(oprob.shape, seed=7, dtype=dt)
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    lhs_o, rhs_o = O.compute_Ax_and_b(params, tr.R.to(dt), *d, oprob)
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *d, oprob)
    p_dev = params.float().to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, p_dev)
        lg = pl.loss_grad_launch()
        torch.cuda.synchronize()
    rhs_k = pl.point_view(pl.rhs_rows()).cpu()
    lhs_k = pl.point_view(pl.R).cpu() + rhs_k
    e_rhs, e_lhs = util.rel_inf(rhs_k, rhs_o), util.rel_inf(lhs_k, lhs_o)
    assert e_rhs < TOL_ROW and e_lhs < TOL_ROW, (e_lhs, e_rhs)
    loss_k, grad_k = float(lg[-1]), lg[:-1].cpu()
    assert abs(loss_k - float(loss_o)) / float(loss_o) < TOL_LOSS, (loss_k, float(loss_o))
    assert util.rel_inf(grad_k, grad_o) < TOL_LOSS, util.rel_inf(grad_k, grad_o)
    # per-coordinate check on the large entries too
    big = grad_o.abs() > 1e-2 * grad_o.abs().max()
    assert ((grad_k.double()[big] - grad_o[big]).abs() / grad_o[big].abs()).max() < 10 * TOL_LOSS
    # a second launch gives the same answer (the fused kernel re-zeroes the list contributions it consumed; the adjoint
    # of the lists uses fp32 atomics: order-dependent in the last bits only)
    with torch.cuda.device(DEV):
        lg2 = pl.loss_grad_launch().clone()
        torch.cuda.synchronize()
    assert util.rel_inf(lg2[:-1].cpu(), grad_k) < 1e-5


@pytest.mark.parametrize("name,faces", [("star", True), ("sphere", False)])
def test_deterministic_list_adjoint_is_bitwise_reproducible(name, faces):
    """deterministic=True: the adjoint of the irregular rows and of the extrapolation is gathered through the transposed
    incidence (no atomics): same gradient as the scattered form within fp32 rounding, and bitwise identical from launch to
    launch."""
    P = problems.PROBLEMS[name]()
    tr, lv, phi_grid, oprob = util.make_case(P, 16, 32, "trilinear", torch.float64)
    lvl = nplan.LevelSet(lv, phi_grid, device=DEV)
    shape = nplan.NetShape()
    mk = lambda det: nplan.SharedPlan(lvl, tr, 0, 16, fns_of(P), shape, nplan.Nonlinear(), nplan.Nonlinear(), device=DEV,
                                      faces=faces, deterministic=det)
    a, b = mk(True), mk(False)
    assert a.g_ptr is not None and b.g_ptr is None and a.n_list > 0
    params = O.init_params(oprob.shape, seed=7).to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        runs = [a.loss_grad_launch().clone() for _ in range(4)]
        ref = b.loss_grad_launch().clone()
        torch.cuda.synchronize()
    assert all(torch.equal(r, runs[0]) for r in runs[1:])
    assert util.rel_inf(runs[0], ref) < 1e-6


def test_anisotropic_grid_and_interface_at_the_box_boundary():
    """ragged sizes (Nx != Ny != Nz, dx != dy != dz) and crossed cells next to the Dirichlet boundary
    (regression cubes that reach the halo layer outside the box)"""
    P = problems.sphere_at_boundary()
    dt = torch.float64
    tr, lv, phi_grid, oprob = util.make_case(P, [14, 10, 12], [30, 22, 26], "trilinear", dt)
    lvl = nplan.LevelSet(lv, phi_grid, device=DEV)
    shape = nplan.NetShape()
    pl = nplan.SharedPlan(lvl, tr, 0, 14, fns_of(P), shape, nplan.Nonlinear(), nplan.Nonlinear(), device=DEV)
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    assert len({round(float(v), 6) for v in d}) == 3
    flag_o = O.is_cell_crossed(tr.R.to(dt), *d, oprob.phi_fn)
    R3 = tr.R.reshape(14, 10, 12, 3)
    near_face = (flag_o.reshape(14, 10, 12)[-2, :, :] == 0).sum()
    assert int(near_face) > 0, "the test problem must put crossed cells next to the x+ boundary"
    assert torch.equal(pl.point_view(pl.sites.flag).cpu().to(dt), flag_o)
    params = O.init_params(oprob.shape, seed=21, dtype=dt)
    lhs_o, rhs_o = O.compute_Ax_and_b(params, tr.R.to(dt), *d, oprob)
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *d, oprob)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params.float().to(DEV))
        lg = pl.loss_grad_launch().cpu()
    rhs_k = pl.point_view(pl.rhs_rows()).cpu()
    lhs_k = pl.point_view(pl.R).cpu() + rhs_k
    assert util.rel_inf(lhs_k, lhs_o) < TOL_ROW and util.rel_inf(rhs_k, rhs_o) < TOL_ROW
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
    assert util.rel_inf(lg[:-1], grad_o) < TOL_LOSS
    # the general path on the same problem, zoom level 1
    d1 = [float(torch.tensor(float(v), dtype=torch.float32) * 0.5) for v in (tr.dx, tr.dy, tr.dz)]
    level = nplan.GeneralLevel(lvl, tr, d1, fns_of(P), shape, nplan.Nonlinear(), nplan.Nonlinear(), device=DEV)
    pp = nplan.PointsPlan(level, 0, tr.num_points())
    dd = [torch.tensor(v, dtype=dt) for v in d1]
    loss_1, grad_1 = O.loss_and_grad(params, tr.R.to(dt), *dd, oprob)
    with torch.cuda.device(DEV):
        lg1 = pp.loss_grad_launch().cpu()
    assert abs(float(lg1[-1]) - float(loss_1)) / float(loss_1) < TOL_LOSS
    assert util.rel_inf(lg1[:-1], grad_1) < TOL_LOSS


def test_smoke_entry_point():
    import __graft_entry__ as g
    g.smoke()


def test_slab_plans_sum_to_the_whole_grid():
    """x-slabs (the multi-GPU partition, data_management.py:121-130) : sum of per-slab
    n_slab*mean-gradients equals the whole-grid n*mean-gradient."""
    P = problems.sphere()
    tr, lv, lvl, oprob, whole, shape = build(P, 16, 32)
    params = O.init_params(oprob.shape, seed=3).to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        ref = whole.loss_grad_launch().clone()
        acc = torch.zeros_like(ref)
        for (a, b) in ((0, 4), (4, 8), (8, 16)):
            tr2, lv2, lvl2, _, pl, _ = build(P, 16, 32, xa=a, xb=b)
            acc += pl.loss_grad_launch() * (pl.n_points / whole.n_points)
        torch.cuda.synchronize()
    assert util.rel_inf(acc, ref) < 1e-5


@pytest.mark.parametrize("name,faces", [("sphere", True), ("star", False), ("pb", True)])
def test_learned_preconditioner(name, faces):
    """P = 0.5 + s * sigmoid(MLP(coeffs_)) multiplies lhs/diag and rhs/diag of every row (nn/preconditioner.py:10-35,
    discretization.py:339, 418-419): loss, d loss/d network parameters and d loss/d preconditioner parameters."""
    if name == "pb":
        P = problems.poisson_boltzmann(n_atoms=6, seed=3, half_width=1.0)
        P.nonlinear_op_p = nplan.Nonlinear.sinh(3000.0)
    else:
        P = problems.PROBLEMS[name]()
    dt = torch.float64
    tr, lv, phi_grid, oprob = util.make_case(P, 16, 32, "trilinear", dt)
    oprob.precond = O.PrecondShape((8, 4), 1.0)
    lvl = nplan.LevelSet(lv, phi_grid, device=DEV)
    shape = nplan.NetShape()
    pc = nplan.PrecondShape((8, 4), 1.0)
    assert pc.n_params == oprob.precond.n_params == 257
    pl = nplan.SharedPlan(lvl, tr, 0, 16, fns_of(P), shape, nplan.Nonlinear.coerce(P.nonlinear_op_m),
                          nplan.Nonlinear.coerce(P.nonlinear_op_p), device=DEV, faces=faces, precond=pc)
    # inputs of the preconditioner are O(h): scale its first layer up so that P varies from point to point
    pp = O.init_precond_params(oprob.precond, seed=5, dtype=dt)
    pp[:26 * 8] *= 40.0
    params = torch.cat((O.init_params(oprob.shape, seed=7, dtype=dt), pp))
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *d, oprob)
    lhs_o, rhs_o, parts = O.compute_Ax_and_b(params, tr.R.to(dt), *d, oprob, return_parts=True)
    Pc = O.precond_eval(params[shape.n_params:], oprob.precond, parts["coeffs"])
    assert float(Pc.max() - Pc.min()) > 1e-2, "the test must exercise a varying preconditioner"
    # the 26 inputs the kernels hand to the preconditioner
    c26 = pl.coef26.view(26, *pl.dims)[:, pl.HX:-pl.HX, pl.HY:-pl.HY, pl.HZ:-pl.hz_hi].reshape(26, -1).T.cpu()
    assert util.rel_inf(c26, parts["coeffs"]) < TOL_FRAC
    p_dev = params.float().to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, p_dev)
        pl.bind_params(p_dev)
        lg = pl.loss_grad_launch().cpu()
    assert lg.numel() == shape.n_params + 257 + 1
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS, (float(lg[-1]), float(loss_o))
    n = shape.n_params
    assert util.rel_inf(lg[:n], grad_o[:n]) < TOL_LOSS, util.rel_inf(lg[:n], grad_o[:n])
    assert util.rel_inf(lg[n:-1], grad_o[n:]) < TOL_LOSS, util.rel_inf(lg[n:-1], grad_o[n:])
    # R now holds d loss/d r (times n) = P^2 r
    r_o = (lhs_o - rhs_o) / Pc
    got = pl.point_view(pl.R).cpu()
    assert util.rel_inf(got, Pc * Pc * r_o) < 10 * TOL_ROW


# The TMA-fed kernel computes residual rows and adjoint stencil in one pass (T stays in shared memory); its arithmetic
# order is that of the two separate kernels, so R and the dense part of G must be BITWISE equal; with the list kernels
# (fp32 atomics) the whole [grad, loss] agrees to the last bits.  Slabs (xa, xb) exercise the x halo planes, odd sizes the
# padded lattice rows, tiny/large grids the tile chooser, sphere_reaction the kv table and the nonlinear terms.
@pytest.mark.parametrize("name,n,nl,xa,xb", [
    ("sphere", 16, 32, 0, None), ("star", 32, 32, 0, None), ("star", 15, 32, 0, None), ("sphere", 24, 24, 5, 17),
    ("sphere_reaction", 16, 32, 0, None), ("sphere_reaction", 20, 32, 3, 12), ("sphere", 64, 64, 0, None),
    ("sphere", 6, 16, 0, None)])
def test_stencil_tma_is_bitwise_the_two_stencil_kernels(name, n, nl, xa, xb):
    P = problems.PROBLEMS[name]()
    tr, lv, lvl, oprob, pa, shape = build(P, n, nl, xa=xa, xb=xb, faces=True, stencil_tma=True)
    _, _, _, _, pb, _ = build(P, n, nl, xa=xa, xb=xb, faces=True, stencil_tma=False)
    # with a nonlinear operator the plan keeps the two separate kernels (sinh / cosh per cell sit badly on the TMA
    # kernel's one-barrier-per-plane critical path: 510 us against 127 + 131 us at Poisson-Boltzmann 256^3)
    nonlinear = P.nonlinear_op_m is not None or P.nonlinear_op_p is not None
    assert pa.stencil_tma_active == (not nonlinear) and not pb.stencil_tma_active
    params = O.init_params(oprob.shape, seed=11, dtype=torch.float64).float().to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        for pl in (pa, pb):
            pl.G.fill_(float("nan"))     # every node of G must be written by the dense stage
            pl.step.stages = 1 | 2 | 4 | 8 | 64   # forward, extrapolation, dense residual + adjoint, no list kernels
            pl.loss_grad_launch()
            pl.step.stages = 0
        torch.cuda.synchronize()
        assert torch.equal(pa.U, pb.U)
        assert torch.equal(pa.R, pb.R)
        assert not torch.isnan(pa.G).any()
        assert torch.equal(pa.G, pb.G)
        la = pa.loss_grad_launch().clone()
        lb = pb.loss_grad_launch().clone()
        la2 = pa.loss_grad_launch().clone()      # repeated launches (ring/barrier state starts clean every time)
        torch.cuda.synchronize()
    assert util.rel_inf(la.cpu(), lb.cpu()) < 2e-6, util.rel_inf(la.cpu(), lb.cpu())
    assert util.rel_inf(la2.cpu(), la.cpu()) < 2e-6
    assert abs(float(la[-1]) - float(lb[-1])) <= 1e-6 * abs(float(lb[-1]))


# The list kernels (crossed sites, irregular rows) run on a side stream beside the TMA stencil and leave their adjoint in
# a side buffer that a merge kernel adds to G: same rows bitwise, same G / [grad, loss] up to the order of the fp32
# atomics, the side buffer is left clean, and a CUDA graph of the step replays the fork/join.
@pytest.mark.parametrize("name,n,nl,xa,xb", [
    ("sphere", 16, 32, 0, None), ("star", 32, 32, 0, None), ("sphere", 24, 24, 5, 17),
    ("sphere_reaction", 16, 32, 0, None), ("sphere", 64, 64, 0, None), ("sphere", 24, 24, 0, 4)])
def test_list_chain_beside_the_stencil_equals_the_serial_chain(name, n, nl, xa, xb):
    P = problems.PROBLEMS[name]()
    tr, lv, lvl, oprob, pa, shape = build(P, n, nl, xa=xa, xb=xb, faces=True, stencil_tma=True, overlap_lists=True)
    _, _, _, _, pb, _ = build(P, n, nl, xa=xa, xb=xb, faces=True, stencil_tma=True, overlap_lists=False)
    if pa.sites.n == 0 and pa.n_irr == 0:
        assert not pa.overlap_lists      # a slab without interface has no list chain
        return
    assert pa.overlap_lists and not pb.overlap_lists and pa.G2 is not None and pb.G2 is None
    params = O.init_params(oprob.shape, seed=17, dtype=torch.float64).float().to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        la = pa.loss_grad_launch().clone()
        lb = pb.loss_grad_launch().clone()
        torch.cuda.synchronize()
        assert torch.equal(pa.U, pb.U)
        assert torch.equal(pa.R, pb.R)
        assert torch.equal(pa.E, pb.E)
        assert util.rel_inf(pa.gE.cpu(), pb.gE.cpu()) < 2e-6
        assert util.rel_inf(pa.G.cpu(), pb.G.cpu()) < 2e-6
        assert int((pa.G2 != 0).sum()) == 0
        assert util.rel_inf(la.cpu(), lb.cpu()) < 2e-6, util.rel_inf(la.cpu(), lb.cpu())
        # replayed as a CUDA graph (what the trainer does): the fork / join onto the library's side stream is captured
        side = torch.cuda.Stream(device=DEV)
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            pa.loss_grad_launch()
            with torch.cuda.graph(graph, stream=side):
                pa.loss_grad_launch()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            pa.loss_grad.zero_()
            graph.replay()
        torch.cuda.synchronize()
        assert util.rel_inf(pa.loss_grad.cpu(), lb.cpu()) < 2e-6
        assert int((pa.G2 != 0).sum()) == 0


# The forward kernel keeps the last hidden layer of every plus-side node (activation stash) and the gradient kernel reads
# it back instead of recomputing it: same [grad, loss] as the recomputing kernel up to the rounding of two tanh variants
# (shared-reciprocal in the forward kernel, per-element reciprocal in the recompute), for every compiled head shape.
@pytest.mark.parametrize("name,n,shape_args", [("sphere", 16, None), ("star", 20, None), ("sphere_reaction", 16, None),
                                               ("sphere", 12, (2, 10, 1, 3)), ("sphere", 12, (1, 10, 1, 1)),
                                               ("sphere", 12, (3, 10, 1, 1))])
def test_activation_stash_equals_recompute(name, n, shape_args):
    P = problems.PROBLEMS[name]()
    net = O.NetShape(*shape_args) if shape_args else None
    tr, lv, lvl, oprob, pa, shape = build(P, n, 32, net=net, stash=True)
    _, _, _, _, pb, _ = build(P, n, 32, net=net, stash=False)
    assert pa.Hst is not None and pb.Hst is None
    params = O.init_params(oprob.shape, seed=13, dtype=torch.float64).float().to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        la = pa.loss_grad_launch().clone()
        lb = pb.loss_grad_launch().clone()
        torch.cuda.synchronize()
        assert float(pa.Hst.abs().max()) > 0 and float(pa.Hst.abs().max()) <= 1.0    # tanh outputs were stored
    assert util.rel_inf(la.cpu(), lb.cpu()) < 5e-6, util.rel_inf(la.cpu(), lb.cpu())
