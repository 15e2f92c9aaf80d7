"""Golden LOSS and DIRECTIONAL DERIVATIVES of the loss, from the REFERENCE'S OWN SOURCE FILES.

    python oracle/make_golden_grad.py       # writes tests/golden/grad_*.npz   (build container only)

The reference differentiates `Trainer.loss` with `jax.value_and_grad` (trainer.py:786, 893); the numpy stand-in for
jax (oracle/jax_shim) has no autodiff, so the gradient cannot be taken the reference's way here.  What CAN be taken
from the reference's own code is the loss itself, `mean(0.5 (lhs - rhs)^2)` over a set of points (trainer.py:892-912 on
top of discretization.py:299-423, read in place from /root/reference), and therefore its central differences along
parameter directions in x64 mode.  These pin the oracle's autograd gradient (and, through the CUDA-vs-oracle tests,
the kernels') to the reference's own arithmetic at the level of the stated tolerance (1e-4).

Stored per case: the point set (the golden rows' points), the parameter vector, k directions, loss(theta) and
(loss(theta + eps d) - loss(theta - eps d)) / (2 eps).
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)   # sets up the stand-in and imports the reference modules

import numpy as np  # noqa: E402
import torch  # noqa: E402

jax, jnp = mg.jax, mg.jnp
EPS = 1e-5


def reference_loss_fn(problem, n_tr, n_lvl, interp, point_idx, zoom, precond=False):
    """loss(flat params) through the reference's Discretization (x64 mode), cf. make_golden.run_case"""
    dtype = np.float64
    jax.config.update("jax_enable_x64", True)
    lo, hi = problem.box
    init_mesh_fn, _ = mg.mesh.construct(3)
    ax = lambda n, a: jnp.linspace(lo[a], hi[a], n, dtype=jnp.float32)
    tr_d = [ax(n_tr, a)[1] - ax(n_tr, a)[0] for a in range(3)]
    lv = init_mesh_fn(ax(n_lvl, 0), ax(n_lvl, 1), ax(n_lvl, 2))
    phi_grid = mg.tnp.vmap(problem.phi_fn)(torch.from_numpy(np.asarray(lv.R, dtype=np.float32))).numpy()
    if interp == "trilinear":
        base = mg.interpolate.multilinear_interpolation(jnp.array(phi_grid), lv)
    else:
        base = jax.vmap(mg.interpolate.nonoscillatory_quadratic_interpolation_per_point(jnp.array(phi_grid), lv))
    phi_fn = mg.level_set.perturb_level_set_fn(base)
    b = lambda fn: mg.torch_batched(fn, dtype)
    fns = mg.PoissonSimStateFn(b(problem.initial_value_fn), b(problem.dirichlet_bc_fn), phi_fn, b(problem.mu_m_fn),
                               b(problem.mu_p_fn), b(problem.k_m_fn), b(problem.k_p_fn), b(problem.f_m_fn),
                               b(problem.f_p_fn), b(problem.alpha_fn), b(problem.beta_fn),
                               mg.nonlinear_callable(problem.nonlinear_op_m), mg.nonlinear_callable(problem.nonlinear_op_p))
    D = (mg.HookedPrecond if precond else mg.Hooked)(lv, None, fns, precondition=1, algorithm=0)
    if precond:
        D.set_precond(*mg.PRECOND)
    shape = mg.O.NetShape()
    d = [dtype(np.float32(v) * np.float32(0.5 ** zoom)) for v in tr_d]
    pts = mg.grid_points(lo, hi, n_tr, point_idx).astype(dtype)

    def loss(flat):
        flat = np.asarray(flat, dtype=dtype)
        D.set_net(jnp.Arr(flat[:shape.n_params]), shape)
        tree = {"preconditioner": mg.precond_tree(flat[shape.n_params:], mg.PRECOND[0])} if precond else None
        acc = 0.0
        for p in pts:
            lr = np.asarray(D.compute_Ax_and_b_fn(tree, jnp.Arr(p), *d)).reshape(2)
            acc += 0.5 * (lr[0] - lr[1]) ** 2          # optax.l2_loss (trainer.py:899-901)
        return acc / len(pts)                           # jnp.mean over the batch

    return loss, shape


CASES = [("sphere_tri_z0", "sphere", "trilinear"), ("star_tri_z0", "star", "trilinear"),
         ("sphere_tri_z1", "sphere", "trilinear"),      # zoom level 1: cell size = spacing / 2
         ("sphere_quad_z0", "sphere", "quadratic"),     # non-oscillatory quadratic level-set interpolant
         ("sphere_reaction_tri_z0", "sphere_reaction", "trilinear"),   # k != 0, N(u) = c sinh(u)
         ("sphere_precond_tri_z0", "sphere", "trilinear"),             # learned preconditioner: theta = [network | preconditioner]
         ("stars_tri_z0", "stars", "trilinear"), ("dragon_quad_z0", "dragon_like", "quadratic"),
         ("pb_tri_z0", "poisson_boltzmann", "trilinear")]


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, pname, interp in CASES:
        if only and name not in only:
            continue
        z = np.load(os.path.join(outdir, f"{name}.npz"))
        P = mg.problems.PROBLEMS[pname]()
        idx, n_tr, n_lvl, zoom = z["point_idx"], int(z["n_tr"]), int(z["n_lvl"]), int(z["zoom"])
        precond = "f64_pc_params" in z.files
        loss, shape = reference_loss_fn(P, n_tr, n_lvl, interp, idx, zoom, precond=precond)
        theta = mg.O.init_params(shape, seed=7, dtype=torch.float64).numpy()
        if precond:
            theta = np.concatenate((theta, mg.precond_flat(np.float64)))
        rng = np.random.default_rng(11)
        n = theta.size
        dirs = []
        for _ in range(5):                                  # dense random directions
            v = rng.standard_normal(n)
            dirs.append(v / np.linalg.norm(v))
        coords = list(rng.choice(shape.n_p, 4, replace=False)) + list(shape.n_p + rng.choice(shape.n_params - shape.n_p, 2, replace=False))
        if precond:     # + single coordinates of the preconditioner (first-layer kernel, deeper layers)
            coords += list(shape.n_params + rng.choice(26 * 8, 3, replace=False)) + list(shape.n_params + 26 * 8 + rng.choice(n - shape.n_params - 26 * 8, 3, replace=False))
        for i in coords:
            e = np.zeros(n)                                 # single coordinates of both heads
            e[i] = 1.0
            dirs.append(e)
        dirs = np.asarray(dirs)
        l0 = loss(theta)
        dl = np.asarray([(loss(theta + EPS * v) - loss(theta - EPS * v)) / (2 * EPS) for v in dirs])
        print(name, "loss", l0, "directional derivatives", dl, flush=True)
        np.savez_compressed(os.path.join(outdir, f"grad_{name}.npz"), point_idx=idx, n_tr=n_tr, n_lvl=n_lvl, zoom=zoom,
                            params=theta, dirs=dirs, loss=l0, dloss=dl, eps=EPS)


if __name__ == "__main__":
    main()
