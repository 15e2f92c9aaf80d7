#!/usr/bin/env python
"""The reference's tests/test_poisson.py flow (sphere of tests/confs/experiment_configs.py:21) on the B200 path:
same `setup -> init_fn -> solve_fn` calls, same dictionaries, the callables written against `jax_dips_b200.numpy`.

    python examples/solve_sphere.py [--n-train 32] [--epochs 400] [--analytic]

Prints the accuracy figures test_poisson.py logs (:276-283) and writes `results/sphere.vts` like :249-253.
"""
import argparse
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from jax_dips_b200 import io, mesh, problems, trainer
from jax_dips_b200 import numpy as jnp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-train", type=int, default=32)
    ap.add_argument("--n-lvl", type=int, default=128)
    ap.add_argument("--n-eval", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=400)
    ap.add_argument("--batch-size", type=int, default=131072)
    ap.add_argument("--analytic", action="store_true", help="use the callable itself as the level set (phi_interp='analytic')")
    ap.add_argument("--results", default="results")
    args = ap.parse_args()
    logging.basicConfig(level=logging.WARNING)

    P = problems.sphere()
    lo, hi = P.box
    init_mesh_fn, _ = mesh.construct(3)
    ax = lambda n, a: jnp.linspace(lo[a], hi[a], n, dtype=torch.float32)
    tr_gstate = init_mesh_fn(ax(args.n_train, 0), ax(args.n_train, 1), ax(args.n_train, 2))
    lvl_gstate = init_mesh_fn(ax(args.n_lvl, 0), ax(args.n_lvl, 1), ax(args.n_lvl, 2))
    eval_gstate = init_mesh_fn(ax(args.n_eval, 0), ax(args.n_eval, 1), ax(args.n_eval, 2))

    init_fn = trainer.setup(P.initial_value_fn, P.dirichlet_bc_fn, P.phi_fn, P.mu_m_fn, P.mu_p_fn, P.k_m_fn, P.k_p_fn,
                            P.f_m_fn, P.f_p_fn, P.alpha_fn, P.beta_fn)
    optimizer_dict = {"optimizer_name": "custom", "learning_rate": 1e-2,
                      "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    model_dict = {"name": None, "model_type": "mlp",
                  "mlp": {"hidden_layers_m": 1, "hidden_dim_m": 1, "activation_m": "jnp.tanh",
                          "hidden_layers_p": 2, "hidden_dim_p": 10, "activation_p": "jnp.tanh"},
                  "preconditioner": {"enable": False}}
    sim_state, solve_fn = init_fn(lvl_gstate=lvl_gstate, tr_gstate=tr_gstate, eval_gstate=eval_gstate,
                                  num_epochs=args.epochs, batch_size=args.batch_size, multi_gpu=False,
                                  checkpoint_dir=os.path.join(args.results, "checkpoints"), results_dir=args.results,
                                  optimizer_dict=optimizer_dict, model_dict=model_dict, print_rate=0,
                                  phi_interp="analytic" if args.analytic else "trilinear")
    t1 = time.time()
    sim_state, epoch_store, loss_epochs = solve_fn(sim_state=sim_state)
    t2 = time.time()
    print(f"solve took {t2 - t1:.2f} s for {args.epochs} epochs on {args.n_train}^3 points; "
          f"loss {float(loss_epochs[0]):.3e} -> {float(loss_epochs[-1]):.3e}")

    R = eval_gstate.R
    phi = jnp.vmap(P.phi_fn)(R)
    exact = torch.where(phi >= 0, jnp.vmap(P.exact_sol_p_fn)(R), jnp.vmap(P.exact_sol_m_fn)(R))
    sol = sim_state.solution.cpu()
    err = sol - exact
    print(f"Accuracy:\n L_inf : {float(err.abs().max()):.4e}\n L_2 : {float((err ** 2).sum().sqrt()):.4e}\n"
          f" Rel. L_2 : {float(((err ** 2).sum() / (exact ** 2).sum()).sqrt()):.4e}\n"
          f" RMSD error : {float((err ** 2).mean().sqrt()):.4e}")
    path = io.write_vtk_manual(eval_gstate, {"phi": phi, "U": sol, "U_exact": exact, "U-U_exact": err},
                               filename=os.path.join(args.results, "sphere"))
    print("wrote", path)


if __name__ == "__main__":
    main()
