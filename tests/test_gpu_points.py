"""GPU parity of the general per-point path (any cell size / batch) and of the training loops."""
import pytest
import torch

import util
from jax_dips_b200 import mesh, plan as nplan, problems, trainer as ntrainer
from oracle import nbm_oracle as O
from test_gpu_shared import DEV, TOL_LOSS, TOL_ROW, build, fns_of

pytestmark = pytest.mark.gpu


def general(problem, n_train, n_lvl, zoom, interp="trilinear"):
    tr, lv, phi_grid, oprob = util.make_case(problem, n_train, n_lvl, interp, torch.float64)
    lvl = nplan.LevelSet(lv, phi_grid, interp=interp, perturb_eps=1e-10, device=DEV)
    shape = nplan.NetShape()
    d = [float(v) * 0.5 ** zoom for v in (tr.dx, tr.dy, tr.dz)]
    # the reference forms the zoomed cell in float32: gstate.dx * 0.5**zoom
    d32 = [float(torch.tensor(float(v), dtype=torch.float32) * torch.tensor(0.5 ** zoom, dtype=torch.float32))
           for v in (tr.dx, tr.dy, tr.dz)]
    level = nplan.GeneralLevel(lvl, tr, d32, fns_of(problem), shape, nplan.Nonlinear.coerce(problem.nonlinear_op_m),
                               nplan.Nonlinear.coerce(problem.nonlinear_op_p), device=DEV)
    return tr, oprob, level, shape, d32


@pytest.mark.parametrize("name,zoom", [("sphere", 0), ("sphere", 1), ("star", 2), ("sphere", 3)])
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


def _solve(problem, n_tr, n_lvl, n_eval, num_epochs, batch_size, init, multi_gpu=False, optimizer_dict=None):
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
                                  optimizer_dict=od, init_params=init, device=DEV, print_rate=0)
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

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
CASE_PROBLEM = {"sphere_tri_z0": ("sphere", "trilinear"), "sphere_tri_z1": ("sphere", "trilinear"),
                "star_tri_z0": ("star", "trilinear"), "sphere_quad_z0": ("sphere", "quadratic")}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_rows_match_reference_golden(path):
    """CUDA rows and cut-cell fractions at the golden points against what the REFERENCE'S OWN CODE
    produced for the same inputs (x64 run).  Rows 1e-5; fractions 5e-5 = the reference's own
    float32-vs-x64 spread (tests/test_oracle_pinning.py::test_reference_noise_floor)."""
    name = os.path.basename(path)[:-4]
    pname, interp = CASE_PROBLEM[name]
    z = np.load(path)
    P = problems.PROBLEMS[pname]()
    zoom, idx = int(z["zoom"]), torch.from_numpy(z["point_idx"]).long()
    params = torch.from_numpy(z["f32_params"]).float()
    gold = torch.from_numpy(z["f64_lhs_rhs"])
    coeffs = torch.from_numpy(z["f64_coeffs"])
    d = z["f64_d"]
    if zoom == 0:
        tr, lv, lvl, oprob, pl, shape = build(P, int(z["n_tr"]), int(z["n_lvl"]), interp)
        with torch.cuda.device(DEV):
            nplan.upload_params(shape, params.to(DEV))
            pl.loss_grad_launch()
            torch.cuda.synchronize()
        rhs_k = pl.point_view(pl.rhs).cpu()[idx]
        lhs_k = pl.point_view(pl.R).cpu()[idx] + rhs_k
        flag_k = pl.point_view(pl.sites.flag).cpu()[idx]
        cidx = pl.point_view(pl.sites.cidx).cpu()[idx]
        frac = pl.sites.frac.view(-1, 14).cpu()
    else:
        tr, oprob, level, shape, dd = general(P, int(z["n_tr"]), int(z["n_lvl"]), zoom, interp)
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
    assert util.rel_inf(lhs_k, gold[:, 0]) < TOL_ROW
    assert util.rel_inf(rhs_k, gold[:, 1]) < TOL_ROW
    cr = flag_k == 0
    if cr.any():
        f = frac[cidx[cr].long()].double()
        vol, area = d.prod(), d[1] * d[2]
        assert float((f[:, 12:14] - coeffs[cr][:, 12:14]).abs().max()) / vol < 5e-5
        assert float((f[:, 0:12] - coeffs[cr][:, 14:26]).abs().max()) / area < 5e-5
