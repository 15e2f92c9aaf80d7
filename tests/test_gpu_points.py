"""GPU parity of the general per-point path (any cell size / batch) and of the training loops."""
import pytest
import torch

import util
from jax_dips_b200 import numpy as jnp
from jax_dips_b200 import mesh, plan as nplan, problems, trainer as ntrainer
from oracle import nbm_oracle as O
from test_gpu_shared import DEV, TOL_LOSS, TOL_ROW, build, fns_of

pytestmark = pytest.mark.gpu


def general(problem, n_train, n_lvl, zoom, interp="trilinear", precond=None, phi_grid=None):
    tr, lv, phi_grid, oprob = util.make_case(problem, n_train, n_lvl, interp, torch.float64, phi_grid=phi_grid)
    lvl = nplan.LevelSet(lv, phi_grid, interp=interp, perturb_eps=1e-10, device=DEV)
    shape = nplan.NetShape()
    d = [float(v) * 0.5 ** zoom for v in (tr.dx, tr.dy, tr.dz)]
    # the reference forms the zoomed cell in float32: gstate.dx * 0.5**zoom
    d32 = [float(torch.tensor(float(v), dtype=torch.float32) * torch.tensor(0.5 ** zoom, dtype=torch.float32))
           for v in (tr.dx, tr.dy, tr.dz)]
    level = nplan.GeneralLevel(lvl, tr, d32, fns_of(problem), shape, nplan.Nonlinear.coerce(problem.nonlinear_op_m),
                               nplan.Nonlinear.coerce(problem.nonlinear_op_p), device=DEV, precond=precond)
    return tr, oprob, level, shape, d32


@pytest.mark.parametrize("name,zoom", [("sphere", 0), ("sphere", 1), ("star", 2), ("sphere", 3), ("sphere_reaction", 1)])
def test_general_path_loss_and_gradient(name, zoom):
    P = problems.PROBLEMS[name]()
    tr, oprob, level, shape, d = general(P, 12, 32, zoom)
    dt = torch.float64
    params = O.init_params(oprob.shape, seed=11, dtype=dt)
    dd = [torch.tensor(v, dtype=dt) for v in d]
    n = tr.num_points()
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *dd, oprob)
    pl = nplan.PointsPlan(level, 0, n)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params.float().to(DEV))
        lg = pl.loss_grad_launch().cpu()
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
    assert util.rel_inf(lg[:-1], grad_o) < TOL_LOSS
    # a sub-batch: rows [a, b) only
    a, b = n // 4, n // 4 + n // 2
    loss_b, grad_b = O.loss_and_grad(params, tr.R[a:b].to(dt), *dd, oprob)
    pb = nplan.PointsPlan(level, a, b)
    with torch.cuda.device(DEV):
        lgb = pb.loss_grad_launch().cpu()
    assert abs(float(lgb[-1]) - float(loss_b)) / float(loss_b) < TOL_LOSS
    assert util.rel_inf(lgb[:-1], grad_b) < TOL_LOSS


@pytest.mark.parametrize("name", ["sphere", "star", "sphere_reaction"])
def test_zoom_1_on_four_shared_lattices_equals_the_seven_displaced_ones(name, monkeypatch):
    """zoom level 1: p + d e_a of a point is p' - d e_a of its neighbour, so 4 lattices (nodes + three half-offset ones)
    carry every stencil site; same [grad, loss] as the 7-lattice formulation up to the 1-ulp difference between
    fl(x_i + d) and fl(x_{i+1} - d), for the whole grid and for a batch of x planes; other batches keep 7 lattices."""
    P = problems.PROBLEMS[name]()
    tr, oprob, level4, shape, d = general(P, 12, 32, 1)
    monkeypatch.setenv("NBM_ZOOM1_SHARED", "0")
    _, _, level7, _, _ = general(P, 12, 32, 1)
    assert level4.shared4 and not level7.shared4
    _, _, level_z2, _, _ = general(P, 12, 32, 2)
    assert not level_z2.shared4                      # no sharing below half the spacing
    n, plane = tr.num_points(), 12 * 12
    params = O.init_params(oprob.shape, seed=19).to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        for (a, b) in ((0, n), (3 * plane, 9 * plane), (11 * plane, 12 * plane)):
            p4, p7 = nplan.PointsPlan(level4, a, b, keep_rows=True), nplan.PointsPlan(level7, a, b, keep_rows=True)
            assert p4.shared4 and not p7.shared4
            l4, l7 = p4.loss_grad_launch().clone(), p7.loss_grad_launch().clone()
            l4b = p4.loss_grad_launch().clone()      # G4 is cleared by the forward kernel: repeated launches agree
            torch.cuda.synchronize()
            assert util.rel_inf(p4.rows.cpu()[a:b], p7.rows.cpu()[a:b]) < 1e-5
            assert util.rel_inf(l4.cpu(), l7.cpu()) < 1e-5, util.rel_inf(l4.cpu(), l7.cpu())
            assert torch.equal(l4, l4b)
        ragged = nplan.PointsPlan(level4, 5, n - 7)
        assert not ragged.shared4


def test_general_and_shared_paths_agree_at_native_spacing():
    P = problems.star()
    tr, lv, lvl, oprob, shared, shape = build(P, 16, 32)
    _, _, level, _, _ = general(P, 16, 32, 0)
    pl = nplan.PointsPlan(level, 0, tr.num_points())
    params = O.init_params(oprob.shape, seed=5).to(DEV)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params)
        a = shared.loss_grad_launch().clone()
        b = pl.loss_grad_launch().clone()
    assert util.rel_inf(a, b) < 2e-5


def _solve(problem, n_tr, n_lvl, n_eval, num_epochs, batch_size, init, multi_gpu=False, optimizer_dict=None,
           model_dict=None, phi_interp="trilinear"):
    lo, hi = problem.box
    init_mesh_fn, _ = mesh.construct(3)
    tr = mesh.linspace_grid(lo, hi, [n_tr] * 3)
    lv = mesh.linspace_grid(lo, hi, [n_lvl] * 3)
    ev = mesh.linspace_grid(lo, hi, [n_eval] * 3)
    init_fn = ntrainer.setup(*problem.setup_args())
    od = optimizer_dict or {"optimizer_name": "custom", "learning_rate": 1e-2,
                            "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    sim_state, solve_fn = init_fn(lvl_gstate=lv, tr_gstate=tr, eval_gstate=ev, num_epochs=num_epochs,
                                  batch_size=batch_size, multi_gpu=multi_gpu, checkpoint_dir=None,
                                  optimizer_dict=od, model_dict=model_dict, init_params=init, device=DEV, print_rate=0,
                                  phi_interp=phi_interp)   # the oracle cases below use the gridded level set
    out = solve_fn(sim_state)
    return out, solve_fn.trainer, (tr, lv, ev)


def test_single_gpu_training_follows_the_oracle_trajectory():
    """setup/init_fn/solve_fn on sphere 8^3 (2 batches, 8 epochs = 4 zoom levels x 2) against the
    oracle's single_GPU_train from the same initial parameters."""
    P = problems.sphere()
    n_tr, n_lvl = 8, 24
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    p_o, losses_o = O.single_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=8, batch_size=256,
                                       optimizer_dict=od)
    (state, epoch_store, loss_epochs), T, _ = _solve(P, n_tr, n_lvl, 16, 8, 256, p0.float(), optimizer_dict=od)
    assert list(epoch_store) == list(range(8))
    lk = torch.as_tensor(loss_epochs).double()
    lo = torch.tensor(losses_o, dtype=torch.float64)
    assert ((lk - lo).abs() / lo).max() < 1e-3, (lk, lo)
    assert util.rel_inf(T.params.cpu(), p_o) < 1e-3
    # evaluation kernel against the oracle on the eval grid
    u_o, gu_o, gn_o = O.evaluate_solution_and_gradients(T.params.cpu().double(), T.eval_gstate.R.double(),
                                                        T.eval_gstate.dx.double(), T.eval_gstate.dy.double(),
                                                        T.eval_gstate.dz.double(), oprob)
    assert util.rel_inf(state.solution.cpu(), u_o) < 1e-5
    assert util.rel_inf(state.grad_solution.cpu(), gu_o) < 1e-5
    fin = torch.isfinite(gn_o)
    assert util.rel_inf(state.grad_normal_solution.cpu()[fin], gn_o[fin]) < 1e-4


def test_multi_gpu_loop_with_learned_preconditioner_follows_the_oracle():
    """model_dict["preconditioner"]["enable"] (examples/benchmark_LPBE/conf/lpbe.yaml:62-67) through
    setup/init_fn/solve_fn on the multi_gpu=True loop (one device here): loss trajectory and final parameters
    (network + preconditioner, trained jointly) against the oracle's multi_GPU_train."""
    P = problems.sphere()
    n_tr, n_lvl = 8, 24
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    oprob.precond = O.PrecondShape((8, 4), 1.0)
    pp = O.init_precond_params(oprob.precond, seed=2, dtype=torch.float64)
    pp[:26 * 8] *= 10.0
    p0 = torch.cat((O.init_params(oprob.shape, seed=42, dtype=torch.float64), pp))
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    md = {"name": None, "model_type": "mlp",
          "mlp": {"hidden_layers_m": 1, "hidden_dim_m": 1, "activation_m": "jnp.tanh",
                  "hidden_layers_p": 2, "hidden_dim_p": 10, "activation_p": "jnp.tanh"},
          "preconditioner": {"enable": True, "layer_widths": [8, 4], "scaling_coeff": 1.0}}
    grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    p_o, losses_o = O.multi_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=6, batch_size=256,
                                      n_devices=1, optimizer_dict=od)
    (state, epoch_store, loss_epochs), T, _ = _solve(P, n_tr, n_lvl, 8, 6, 256, p0.float(), optimizer_dict=od,
                                                     model_dict=md, multi_gpu=True)
    lk = torch.stack([torch.as_tensor(l).reshape(-1)[0] for l in loss_epochs]).double()
    lo = torch.tensor(losses_o, dtype=torch.float64)
    assert ((lk - lo).abs() / lo).max() < 1e-3, (lk, lo)
    assert util.rel_inf(T.params.cpu(), p_o) < 1e-3
    tree = ntrainer.params_to_tree(T.net, T.params, T.precond)
    assert tree["preconditioner"]["params"]["Dense_0"]["kernel"].shape == (26, 8)
    assert torch.equal(ntrainer.tree_to_params(T.net, tree, T.precond), T.params.cpu())
    # the single-GPU loop (4 zoom levels x 2 epochs, 2 batches: shared path at zoom 0, general path above)
    p_s, losses_s = O.single_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=8, batch_size=256,
                                       optimizer_dict=od)
    (state, _, loss_epochs), T, _ = _solve(P, n_tr, n_lvl, 8, 8, 256, p0.float(), optimizer_dict=od, model_dict=md)
    lk = torch.as_tensor(loss_epochs).double()
    lo = torch.tensor(losses_s, dtype=torch.float64)
    assert ((lk - lo).abs() / lo).max() < 1e-3, (lk, lo)
    assert util.rel_inf(T.params.cpu(), p_s) < 1e-3


@pytest.mark.parametrize("name,zoom", [("sphere", 1), ("star", 2)])
def test_general_path_with_learned_preconditioner(name, zoom):
    """loss, network gradient and preconditioner gradient of the per-point path (cell size = spacing / 2^zoom)"""
    P = problems.PROBLEMS[name]()
    dt = torch.float64
    tr, oprob, level, shape, d32 = general(P, 12, 32, zoom, precond=nplan.PrecondShape((8, 4), 1.0))
    oprob.precond = O.PrecondShape((8, 4), 1.0)
    n = tr.num_points()
    pp = O.init_precond_params(oprob.precond, seed=5, dtype=dt)
    pp[:26 * 8] *= 60.0
    params = torch.cat((O.init_params(oprob.shape, seed=11, dtype=dt), pp))
    dd = [torch.tensor(v, dtype=dt) for v in d32]
    for (a, b) in ((0, n), (n // 3, n // 3 + 700)):
        pl = nplan.PointsPlan(level, a, b)
        loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt)[a:b], *dd, oprob)
        p_dev = params.float().to(DEV)
        with torch.cuda.device(DEV):
            nplan.upload_params(shape, p_dev)
            pl.bind_params(p_dev)
            lg = pl.loss_grad_launch().cpu()
        assert lg.numel() == params.numel() + 1
        assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
        k = shape.n_params
        assert util.rel_inf(lg[:k], grad_o[:k]) < TOL_LOSS
        assert util.rel_inf(lg[k:-1], grad_o[k:]) < TOL_LOSS


def test_config0_sphere_32_error_metrics_track_the_oracle(tmp_path):
    """BASELINE.json configs[0] (tests/test_poisson.py, sphere of experiment_configs.py:21, 32^3 training grid, default
    MLP, optimizer "custom"): after the same epochs from the same initial vector the accuracy figures the reference's
    test logs (L_inf, RMSD, relative L_2 against the exact solution, test_poisson.py:276-283) must agree with the
    oracle's within 2 % (north star), and the fields go to a VTK file like test_poisson.py:249-253."""
    from jax_dips_b200 import io as nio
    P = problems.sphere()
    n_tr, n_lvl, n_ev, epochs, bs = 32, 64, 32, 4, 16384
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    p_o, losses_o = O.single_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=epochs, batch_size=bs,
                                       optimizer_dict=od)
    (state, _, loss_epochs), T, (trk, lvk, ev) = _solve(P, n_tr, n_lvl, n_ev, epochs, bs, p0.float(), optimizer_dict=od)
    phi_ev = jnp.vmap(P.phi_fn)(ev.R)
    exact = torch.where(phi_ev >= 0, jnp.vmap(P.exact_sol_p_fn)(ev.R), jnp.vmap(P.exact_sol_m_fn)(ev.R)).double()
    u_o = oprob.solution(p_o, ev.R.double())
    u_k = state.solution.cpu().double()

    def metrics(u):
        e = u - exact
        return (float(e.abs().max()), float((e ** 2).mean().sqrt()), float(((e ** 2).sum() / (exact ** 2).sum()).sqrt()))

    mk, mo = metrics(u_k), metrics(u_o)
    for a, b in zip(mk, mo):
        assert abs(a - b) / b < 0.02, (mk, mo)
    assert ((torch.as_tensor(loss_epochs).double() - torch.tensor(losses_o)).abs() / torch.tensor(losses_o)).max() < 1e-3
    path = nio.write_vtk_manual(ev, {"phi": phi_ev, "U": state.solution, "U_exact": exact, "U-U_exact": u_k - exact},
                                filename=str(tmp_path / "sphere"))
    pts, fields = nio.read_vts(path)
    assert set(fields) == {"phi", "U", "U_exact", "U-U_exact"} and pts.shape[0] == n_ev ** 3


def _analytic_case(P, n_tr, n_lvl):
    """oracle problem whose level set is the analytic callable itself (evaluated in float32 like everywhere else)"""
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    v = jnp.vmap(P.phi_fn)
    oprob.phi_fn = lambda R: v(R.to(torch.float32)).to(R.dtype)
    return tr, lv, oprob


@pytest.mark.parametrize("name", ["sphere", "star"])
def test_analytic_level_set_rows_loss_and_gradient(name):
    """the level set as the user's own callable (what tests/test_poisson.py hands to the reference,
    discretization.py:90) instead of its interpolant on lvl_gstate: shared path and per-point path (zoom 1)"""
    P = problems.PROBLEMS[name]()
    dt = torch.float64
    tr, lv, oprob = _analytic_case(P, 16, 32)
    lvl = nplan.AnalyticLevelSet(lv, jnp.vmap(P.phi_fn), device=DEV)
    shape = nplan.NetShape()
    params = O.init_params(oprob.shape, seed=7, dtype=dt)
    with torch.cuda.device(DEV):
        pl = nplan.SharedPlan(lvl, tr, 0, 16, fns_of(P), shape, nplan.Nonlinear(), nplan.Nonlinear(), device=DEV)
        nplan.upload_params(shape, params.float().to(DEV))
        lg = pl.loss_grad_launch().cpu()
    d = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    flag_o = O.is_cell_crossed(tr.R.to(dt), *d, oprob.phi_fn)
    assert torch.equal(pl.point_view(pl.sites.flag).cpu().to(dt), flag_o)
    # the interpolated level set is a different problem: the analytic one must not silently fall back to it
    trg, lvg, phi_grid, og = util.make_case(P, 16, 32, "trilinear", dt)
    lhs_o, rhs_o = O.compute_Ax_and_b(params, tr.R.to(dt), *d, oprob)
    lhs_g, rhs_g = O.compute_Ax_and_b(params, tr.R.to(dt), *d, og)
    assert util.rel_inf(lhs_g, lhs_o) > 1e-4
    rhs_k = pl.point_view(pl.rhs_rows()).cpu()
    lhs_k = pl.point_view(pl.R).cpu() + rhs_k
    assert util.rel_inf(lhs_k, lhs_o) < TOL_ROW and util.rel_inf(rhs_k, rhs_o) < TOL_ROW
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *d, oprob)
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
    assert util.rel_inf(lg[:-1], grad_o) < TOL_LOSS
    # per-point path, zoom level 1
    d32 = [float(torch.tensor(float(v), dtype=torch.float32) * torch.tensor(0.5, dtype=torch.float32))
           for v in (tr.dx, tr.dy, tr.dz)]
    with torch.cuda.device(DEV):
        level = nplan.GeneralLevel(lvl, tr, d32, fns_of(P), shape, nplan.Nonlinear(), nplan.Nonlinear(), device=DEV)
        lg1 = nplan.PointsPlan(level, 0, tr.num_points()).loss_grad_launch().cpu()
    dd = [torch.tensor(v, dtype=dt) for v in d32]
    loss_1, grad_1 = O.loss_and_grad(params, tr.R.to(dt), *dd, oprob)
    assert abs(float(lg1[-1]) - float(loss_1)) / float(loss_1) < TOL_LOSS
    assert util.rel_inf(lg1[:-1], grad_1) < TOL_LOSS


def test_analytic_level_set_through_the_trainer():
    """init_fn(..., phi_interp="analytic"): training loop and post-training evaluation (u, grad u, d u/d n) with the
    callable as the level set, against the oracle driven the same way"""
    P = problems.sphere()
    n_tr, n_lvl = 8, 24
    tr, lv, oprob = _analytic_case(P, n_tr, n_lvl)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    p_o, losses_o = O.single_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=8, batch_size=256,
                                       optimizer_dict=od)
    lo, hi = P.box
    trk = mesh.linspace_grid(lo, hi, [n_tr] * 3); lvk = mesh.linspace_grid(lo, hi, [n_lvl] * 3)
    ev = mesh.linspace_grid(lo, hi, [12] * 3)
    init_fn = ntrainer.setup(*P.setup_args())
    sim_state, solve_fn = init_fn(lvl_gstate=lvk, tr_gstate=trk, eval_gstate=ev, num_epochs=8, batch_size=256,
                                  checkpoint_dir=None, optimizer_dict=od, init_params=p0.float(), device=DEV,
                                  print_rate=0, phi_interp="analytic")
    state, _, loss_epochs = solve_fn(sim_state)
    T = solve_fn.trainer
    lk = torch.as_tensor(loss_epochs).double()
    lo_ = torch.tensor(losses_o, dtype=torch.float64)
    assert ((lk - lo_).abs() / lo_).max() < 1e-3, (lk, lo_)
    assert util.rel_inf(T.params.cpu(), p_o) < 1e-3
    u_o, gu_o, gn_o = O.evaluate_solution_and_gradients(T.params.cpu().double(), ev.R.double(), ev.dx.double(),
                                                        ev.dy.double(), ev.dz.double(), oprob)
    assert util.rel_inf(state.solution.cpu(), u_o) < 1e-5
    assert util.rel_inf(state.grad_solution.cpu(), gu_o) < 1e-5
    fin = torch.isfinite(gn_o)
    assert util.rel_inf(state.grad_normal_solution.cpu()[fin], gn_o[fin]) < 1e-4


def test_region_scaled_update_is_the_references_private_update():
    """Trainer.update_region_scaled = the reference's `__update` (trainer.py:791-816): minus-network gradient times
    mgrad_over_pgrad_scalefactor, one optimizer update, loss = m_loss + p_loss."""
    P = problems.sphere()
    n_tr, n_lvl, factor = 8, 24, 10.0
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    lo, hi = P.box
    g = [mesh.linspace_grid(lo, hi, [n] * 3) for n in (n_lvl, n_tr, 8)]
    init_fn = ntrainer.setup(*P.setup_args())
    sim_state, solve_fn = init_fn(lvl_gstate=g[0], tr_gstate=g[1], eval_gstate=g[2], num_epochs=4, batch_size=512,
                                  mgrad_over_pgrad_scalefactor=factor, checkpoint_dir=None, optimizer_dict=od,
                                  init_params=p0.float(), device=DEV, print_rate=0, phi_interp="trilinear")
    solve_fn(sim_state)
    T = solve_fn.trainer
    plan = T.plan_for(0, 0, n_tr ** 3)
    T.params.copy_(p0.float())      # back to the initial state: the loops above do not use the factor
    T.opt_state.zero_()
    T.opt_count.zero_()
    d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    opt = O.OptaxCustom(p0.numel(), learning_rate=1e-2, decay_rate=0.975, dtype=torch.float64)
    p_o = p0.clone()
    n_p = T.net.n_p
    for _ in range(3):
        loss_o, grad_o = O.loss_and_grad(p_o, tr.R.double(), *d, oprob)
        grad_o = grad_o.clone()
        grad_o[n_p:] *= factor
        p_o = p_o + opt.update(grad_o)
        loss_k = T.update_region_scaled(plan)
        assert abs(float(loss_k) - 2.0 * float(loss_o)) < 1e-4 * abs(2.0 * float(loss_o))
    assert util.rel_inf(T.params.cpu(), p_o) < 1e-4


def test_ragged_batches_keep_their_real_points():
    """512 points in batches of 200 (200 + 200 + 112; the reference would pad the last one with 88 random points from
    jax PRNGKey(0), data_management.py:70-76): unaligned batches run on the per-point path at every zoom level."""
    P = problems.sphere()
    n_tr, n_lvl = 8, 24
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    p_o, losses_o = O.single_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, num_epochs=4, batch_size=200,
                                       optimizer_dict=od)
    (state, epoch_store, loss_epochs), T, _ = _solve(P, n_tr, n_lvl, 8, 4, 200, p0.float(), optimizer_dict=od)
    lk = torch.as_tensor(loss_epochs).double()
    lo = torch.tensor(losses_o, dtype=torch.float64)
    assert ((lk - lo).abs() / lo).max() < 1e-3, (lk, lo)
    assert util.rel_inf(T.params.cpu(), p_o) < 1e-3
    assert int(T.opt_count.item()) == 12


def test_lbfgs_driver_minimises_the_first_batch():
    """optimizer_name "lbfgs" (trainer.py:197-208, 354-427): scipy L-BFGS-B, maxiter = num_epochs, on the first batch
    at the native cell size, driven by the CUDA loss/gradient.  The same scipy call on the oracle's float64
    loss/gradient gives the reference trajectory."""
    from scipy.optimize import minimize
    P = problems.no_jump()
    n_tr, n_lvl = 8, 12
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "lbfgs", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    (state, epoch_store, loss_epochs), T, _ = _solve(P, n_tr, n_lvl, 8, 15, 512, p0.float(), optimizer_dict=od)
    assert T.scipy_result.nit <= 15 and len(T.loss_history) >= 2
    assert T.loss_history[-1] < 0.2 * T.loss_history[0]
    d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]

    def fun(x):
        l, g = O.loss_and_grad(torch.from_numpy(x), tr.R.double(), *d, oprob)
        return float(l), g.numpy()

    l0, _ = fun(p0.numpy())
    assert abs(T.loss_history[0] - l0) / l0 < 1e-4
    sol = minimize(fun, p0.numpy(), jac=True, method="L-BFGS-B", tol=1e-15, options={"maxiter": 15})
    # the fp32 and fp64 line searches take the same path for the first iterations: same order of magnitude at the end
    assert 0.2 < float(T.scipy_result.fun) / float(sol.fun) < 5.0
    assert state.solution.shape[0] == 8 ** 3


def test_training_reduces_the_error_against_the_exact_solution():
    """no_jump problem 16^3, 200 epochs: the L-inf error against the analytic solution must fall
    well below its initial value (the reference only logs this number, tests/test_poisson.py:276-283)."""
    P = problems.no_jump()
    (state, _, losses), T, (tr, lv, ev) = _solve(P, 16, 16, 16, 200, 4096, None)
    from jax_dips_b200 import numpy as jnp
    exact = jnp.vmap(P.exact_sol_p_fn)(ev.R)
    err = float((state.solution.cpu() - exact).abs().max())
    assert float(losses[-1]) < 0.05 * float(losses[0])
    assert err < 0.2, err


# ---------------------------------------------------------------------------------------------
# golden vectors produced by the reference's own sources (oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------
import glob
import os

import numpy as np

TOL_FRAC_REF = 1e-4   # = tests/test_oracle_pinning.py: the reference's own float32 measure noise on 220-cell sets is 6.2e-5
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("grad_"))
CASE_PROBLEM = util.GOLDEN_CASES


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_rows_match_reference_golden(path):
    """CUDA rows and cut-cell fractions at the golden points against what the REFERENCE'S OWN CODE
    produced for the same inputs (x64 run).  Rows 1e-5; fractions 1e-4 = the reference's own
    float32 measure noise (tests/test_oracle_pinning.py::test_reference_noise_floor, TOL_FRAC_REF)."""
    name = os.path.basename(path)[:-4]
    pname, interp = CASE_PROBLEM[name]
    z = np.load(path)
    if "f64_pc_params" in z.files:
        pytest.skip("rows scaled by the learned preconditioner are not kept by the kernels (R <- P^2 r); the pinned "
                    "oracle covers this fixture (test_oracle_pinning) and the kernels are checked against the oracle")
    P = problems.PROBLEMS[pname]()
    zoom, idx = int(z["zoom"]), torch.from_numpy(z["point_idx"]).long()
    params = torch.from_numpy(z["f32_params"]).float()
    gold = torch.from_numpy(z["f64_lhs_rhs"])
    coeffs = torch.from_numpy(z["f64_coeffs"])
    d = z["f64_d"]
    pg = z["f32_phi_grid"]      # the fixture's own level-set samples
    if zoom == 0:
        tr, lv, lvl, oprob, pl, shape = build(P, int(z["n_tr"]), int(z["n_lvl"]), interp, phi_grid=pg)
        with torch.cuda.device(DEV):
            nplan.upload_params(shape, params.to(DEV))
            pl.loss_grad_launch()
            torch.cuda.synchronize()
        rhs_k = pl.point_view(pl.rhs_rows()).cpu()[idx]
        lhs_k = pl.point_view(pl.R).cpu()[idx] + rhs_k
        flag_k = pl.point_view(pl.sites.flag).cpu()[idx]
        cidx = pl.point_view(pl.sites.cidx).cpu()[idx]
        frac = pl.sites.frac.view(-1, 14).cpu()
    else:
        tr, oprob, level, shape, dd = general(P, int(z["n_tr"]), int(z["n_lvl"]), zoom, interp, phi_grid=pg)
        pp = nplan.PointsPlan(level, 0, tr.num_points(), keep_rows=True)
        with torch.cuda.device(DEV):
            nplan.upload_params(shape, params.to(DEV))
            pp.loss_grad_launch()
            torch.cuda.synchronize()
        rhs_k = level.rhs.cpu()[idx]
        lhs_k = pp.rows.cpu()[idx] + rhs_k
        n = tr.num_points()
        flag_k = level.sites.flag[:n].cpu()[idx]
        cidx = level.sites.cidx[:n].cpu()[idx]
        frac = level.sites.frac.view(-1, 14).cpu()
    assert torch.equal(flag_k.double(), torch.from_numpy(z["f64_flag"]))
    cr = flag_k == 0
    util.assert_rows_match(lhs_k, gold[:, 0], cr, TOL_ROW)
    util.assert_rows_match(rhs_k, gold[:, 1], cr, TOL_ROW)
    if cr.any():
        f = frac[cidx[cr].long()].double()
        vol, area = d.prod(), d[1] * d[2]
        ev = float((f[:, 12:14] - coeffs[cr][:, 12:14]).abs().max()) / vol
        ea = float((f[:, 0:12] - coeffs[cr][:, 14:26]).abs().max()) / area
        assert ev < TOL_FRAC_REF and ea < TOL_FRAC_REF, (ev, ea)


# ---------------------------------------------------------------------------------------------
# nonlinear operator (Poisson-Boltzmann), other network shapes, optimizers, checkpoints
# ---------------------------------------------------------------------------------------------
def _pb_small():
    P = problems.poisson_boltzmann(n_atoms=6, seed=3, half_width=1.0)
    # strong operators on BOTH sides so that the nonlinear terms are a visible part of every row
    P.nonlinear_op_p = nplan.Nonlinear.sinh(3000.0)
    P.nonlinear_op_m = nplan.Nonlinear.sinh(400.0)
    return P


@pytest.mark.parametrize("path_kind", ["shared", "general"])
def test_nonlinear_sinh_operator(path_kind):
    """N^+(u) = kappa^2 sinh(u) (examples/biomolecules/coefficients.py:126-131) in rows, loss and gradient."""
    P = _pb_small()
    dt = torch.float64
    if path_kind == "shared":
        tr, lv, lvl, oprob, pl, shape = build(P, 16, 32)
        dd = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    else:
        tr, oprob, level, shape, d32 = general(P, 12, 32, 1)
        pl = nplan.PointsPlan(level, 0, tr.num_points())
        dd = [torch.tensor(v, dtype=dt) for v in d32]
    params = O.init_params(oprob.shape, seed=9, dtype=dt) * 3.0      # larger u so that sinh is not ~linear
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *dd, oprob)
    # the nonlinear term must matter in this test
    lin = O.OracleProblem(oprob.phi_fn, oprob.mu_m_fn, oprob.mu_p_fn, oprob.k_m_fn, oprob.k_p_fn, oprob.f_m_fn,
                          oprob.f_p_fn, oprob.alpha_fn, oprob.beta_fn, oprob.dir_bc_fn, oprob.bounds, oprob.shape)
    loss_lin, _ = O.loss_and_grad(params, tr.R.to(dt), *dd, lin)
    assert abs(float(loss_lin) - float(loss_o)) / float(loss_o) > 1e-3
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params.float().to(DEV))
        lg = pl.loss_grad_launch().cpu()
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
    assert util.rel_inf(lg[:-1], grad_o) < TOL_LOSS


@pytest.mark.parametrize("shape_args", [(2, 10, 1, 3), (1, 10, 1, 1), (3, 10, 1, 1)])
def test_other_network_shapes(shape_args):
    P = problems.sphere()
    net = O.NetShape(*shape_args)
    tr, lv, lvl, oprob, pl, shape = build(P, 12, 24, net=net)
    dt = torch.float64
    params = O.init_params(net, seed=4, dtype=dt)
    dd = [tr.dx.to(dt), tr.dy.to(dt), tr.dz.to(dt)]
    loss_o, grad_o = O.loss_and_grad(params, tr.R.to(dt), *dd, oprob)
    with torch.cuda.device(DEV):
        nplan.upload_params(shape, params.float().to(DEV))
        lg = pl.loss_grad_launch().cpu()
    assert lg.numel() == net.n_params + 1
    assert abs(float(lg[-1]) - float(loss_o)) / float(loss_o) < TOL_LOSS
    assert util.rel_inf(lg[:-1], grad_o) < TOL_LOSS


def test_unsupported_network_shape_fails_loudly():
    from jax_dips_b200 import _cabi
    P = problems.sphere()
    with pytest.raises(_cabi.NbmError):
        tr, lv, lvl, oprob, pl, shape = build(P, 8, 16, net=O.NetShape(2, 7, 1, 1))
        with torch.cuda.device(DEV):
            nplan.upload_params(shape, O.init_params(O.NetShape(2, 7, 1, 1)).to(DEV))
            pl.loss_grad_launch()


def test_finalize_kernel_with_the_polynomial_schedule():
    """nbm_finalize_step_f32 (row reduction + optax chain + staging) with optimizer "custom" and the "polynomial"
    scheduler (optimizers.py:25-29: optax.polynomial_schedule(lr, 0, power 1, transition_steps)) against a float64
    restatement; the staged copies are what nbm_upload_staged_params hands to the next step."""
    import ctypes as C
    from jax_dips_b200 import _cabi as cabi
    net = nplan.NetShape()
    n, rows, T = net.n_params, 5, 4.0
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g, dtype=torch.float64) * 0.1
    lr = 1e-2
    p = p0.clone(); m = torch.zeros(n, dtype=torch.float64); v = torch.zeros(n, dtype=torch.float64)
    o = cabi.Optimizer(n, lr, 0.9, T, 1.0, 0.9, 0.999, 1e-8, 0, 1)
    s = net.struct()
    with torch.cuda.device(DEV):
        pk = p0.float().to(DEV); st = torch.zeros(2 * n, device=DEV); cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        lg = torch.zeros(n + 1, device=DEV)
        for t in range(6):                     # the schedule reaches 0 after transition_steps = 4
            part = torch.randn(rows, n + 1, generator=g, dtype=torch.float64) * (3.0 if t == 0 else 0.05)
            gr = part[:, :n].sum(0)
            gn = gr.norm()
            gr = gr if gn < 1.0 else gr / gn
            m = 0.9 * m + 0.1 * gr; v = 0.999 * v + 0.001 * gr * gr
            upd = (m / (1 - 0.9 ** (t + 1))) / (torch.sqrt(v / (1 - 0.999 ** (t + 1))) + 1e-8)
            p = p - lr * (1.0 - min(t, T) / T) * upd
            pd = part.float().to(DEV).contiguous()
            cabi.check(cabi.lib().nbm_finalize_step_f32(C.byref(o), C.byref(s), cabi.ptr(pd), rows, n + 1, cabi.ptr(lg),
                                                        cabi.ptr(pk), cabi.ptr(st), cabi.ptr(cnt), None, cabi.stream_ptr()))
            torch.cuda.synchronize()
            assert util.rel_inf(lg.cpu(), part.sum(0)) < 1e-6
        assert int(cnt.item()) == 6
        assert util.rel_inf(pk.cpu(), p) < 1e-5
        # the staged parameters are the updated ones: same network values as after a full upload
        cabi.check(cabi.lib().nbm_upload_staged_params(cabi.stream_ptr()))
        P = problems.sphere()
        tr, lv, phi_grid, oprob = util.make_case(P, 8, 16, "trilinear", torch.float64)
        lvl = nplan.LevelSet(lv, phi_grid, device=DEV)
        pts = tr.R[:64].to(DEV).contiguous()
        u1 = torch.empty(64, device=DEV); u2 = torch.empty(64, device=DEV)
        ev = lambda out: cabi.check(cabi.lib().nbm_evaluate_f32(C.byref(s), C.byref(lvl.struct), cabi.ptr(pts), 64, 1.0, 1.0,
                                                                1.0, cabi.ptr(out), None, None, cabi.stream_ptr()))
        ev(u1)
        nplan.upload_params(net, pk)
        ev(u2)
        torch.cuda.synchronize()
    assert torch.equal(u1, u2)


@pytest.mark.parametrize("name", ["custom", "adam", "rmsprop"])
def test_update_kernel_matches_optax_semantics(name):
    """nbm_apply_update_f32 against a float64 restatement of the optax chains (optimizers.py:33-54, 76-88)."""
    import ctypes as C
    from jax_dips_b200 import _cabi as cabi
    n, steps = 167, 5
    g = torch.Generator().manual_seed(0)
    params0 = torch.randn(n, generator=g, dtype=torch.float64) * 0.1
    grads = [torch.randn(n, generator=g, dtype=torch.float64) * s for s in (5.0, 0.01, 1.0, 0.2, 3.0)]
    lr, decay = 1e-2, 0.9
    # float64 restatement
    p = params0.clone(); m = torch.zeros(n, dtype=torch.float64); v = torch.zeros(n, dtype=torch.float64)
    for t, gr in enumerate(grads):
        if name == "custom":
            gn = gr.norm()
            gr = gr if gn < 1.0 else gr / gn
        if name == "rmsprop":
            v = 0.9 * v + 0.1 * gr * gr
            upd = gr / torch.sqrt(v + 1e-8)
            step = lr
        else:
            m = 0.9 * m + 0.1 * gr; v = 0.999 * v + 0.001 * gr * gr
            upd = (m / (1 - 0.9 ** (t + 1))) / (torch.sqrt(v / (1 - 0.999 ** (t + 1))) + 1e-8)
            step = lr * decay ** (t / 1000) if name == "custom" else lr
        p = p - step * upd
    kind = {"custom": 0, "adam": 1, "rmsprop": 2}[name]
    o = cabi.Optimizer(n, lr, decay, 1000.0, 1.0, 0.9, 0.9 if name == "rmsprop" else 0.999, 1e-8, kind, 0)
    with torch.cuda.device(DEV):
        pk = params0.float().to(DEV); st = torch.zeros(2 * n, device=DEV); cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        hist = torch.zeros(8, device=DEV)
        for t, gr in enumerate(grads):
            lg = torch.cat((gr.float(), torch.tensor([float(t)]))).to(DEV)
            cabi.check(cabi.lib().nbm_apply_update_f32(C.byref(o), cabi.ptr(lg), cabi.ptr(pk), cabi.ptr(st), cabi.ptr(cnt),
                                                       cabi.ptr(hist), cabi.stream_ptr()))
        torch.cuda.synchronize()
    assert int(cnt.item()) == steps
    assert torch.equal(hist[:steps].cpu(), torch.arange(steps).float())
    assert util.rel_inf(pk.cpu(), p) < 1e-5


def test_checkpoint_roundtrip_and_restart(tmp_path):
    """same dict keys as the reference's pickle (trainer.py:340-351); restart restores weights and optimizer state."""
    import pickle
    P = problems.no_jump()
    lo, hi = P.box
    tr = mesh.linspace_grid(lo, hi, [8] * 3); lv = mesh.linspace_grid(lo, hi, [12] * 3); ev = mesh.linspace_grid(lo, hi, [8] * 3)
    init_fn = ntrainer.setup(*P.setup_args())
    ck = str(tmp_path / "ck")
    kw = dict(lvl_gstate=lv, tr_gstate=tr, eval_gstate=ev, num_epochs=8, batch_size=512, device=DEV, print_rate=0)
    s0, solve = init_fn(checkpoint_dir=ck, **kw)
    solve(s0)
    T = solve.trainer
    files = sorted(os.listdir(ck))
    assert files == ["checkpoint_8"]
    state = pickle.load(open(os.path.join(ck, files[0]), "rb"))
    assert set(state) == {"opt_state", "params", "epoch", "batch_size", "resolution"}
    assert "double_mlp/~mlp_p_fn/linear_1" in state["params"]
    s1, solve2 = init_fn(checkpoint_dir=None, restart=True, restart_checkpoint_dir=ck, **kw)
    T2 = ntrainer.Trainer(lv, tr, ev, s1, T.sim_state_fn, num_epochs=8, batch_size=512, restart=True,
                          restart_checkpoint_dir=ck, device=DEV, checkpoint_dir=None)
    assert torch.equal(T2.params.cpu(), T.params.cpu())
    assert torch.equal(T2.opt_state.cpu(), T.opt_state.cpu())
    assert int(T2.opt_count.item()) == int(T.opt_count.item()) == 8
