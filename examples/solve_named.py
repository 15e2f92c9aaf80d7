#!/usr/bin/env python
"""BASELINE.json's other named configurations through the reference's API (`setup -> init_fn -> solve_fn`), with synthetic
geometry (no data files): the flows of examples/stars/solve_stars.py, examples/dragon/solve_dragon.py and
examples/benchmark_LPBE/main.py of the reference.

    python examples/solve_named.py stars              # 64^3 train / 128^3 level-set grid, variable mu jump, 1 GPU
    python examples/solve_named.py dragon_like        # 128^3, level set read through the quadratic interpolant
    python examples/solve_named.py poisson_boltzmann  # 256^3, kappa^2 sinh(u) outside the molecule
    torchrun --nproc-per-node 8 examples/solve_named.py poisson_boltzmann --multi-gpu     # one process per GPU
    python examples/solve_named.py dragon_like --multi-gpu      # one process driving every visible GPU (pmap model)
"""
import argparse
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from jax_dips_b200 import mesh, problems, trainer

CONFIGS = {   # problem -> (train points per axis, level-set points per axis, phi_interp)
    "stars": (64, 128, "trilinear"),
    "dragon_like": (128, 128, "quadratic"),
    "poisson_boltzmann": (256, 128, "analytic"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem", choices=sorted(CONFIGS))
    ap.add_argument("--epochs", type=int, default=40)
    ap.add_argument("--n-train", type=int, default=0, help="override the training grid (points per axis)")
    ap.add_argument("--n-eval", type=int, default=64)
    ap.add_argument("--multi-gpu", action="store_true")
    ap.add_argument("--preconditioner", action="store_true", help="train the learned preconditioner too (lpbe.yaml:62-67)")
    args = ap.parse_args()
    logging.basicConfig(level=logging.WARNING)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))

    n_tr, n_lvl, interp = CONFIGS[args.problem]
    n_tr = args.n_train or n_tr
    P = problems.PROBLEMS[args.problem]()
    lo, hi = P.box
    tr, lv, ev = (mesh.linspace_grid(lo, hi, [n] * 3) for n in (n_tr, n_lvl, args.n_eval))
    init_fn = trainer.setup(*P.setup_args())
    optimizer_dict = {"optimizer_name": "custom", "learning_rate": 1e-3,
                      "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    model_dict = {"name": None, "model_type": "mlp",
                  "mlp": {"hidden_layers_m": 1, "hidden_dim_m": 1, "activation_m": "jnp.tanh",
                          "hidden_layers_p": 2, "hidden_dim_p": 10, "activation_p": "jnp.tanh"},
                  "preconditioner": {"enable": bool(args.preconditioner), "layer_widths": [8, 4], "scaling_coeff": 1.0}}
    t0 = time.time()
    sim_state, solve_fn = init_fn(lvl_gstate=lv, tr_gstate=tr, eval_gstate=ev, num_epochs=args.epochs,
                                  batch_size=tr.num_points(), multi_gpu=args.multi_gpu, checkpoint_dir=None,
                                  optimizer_dict=optimizer_dict, model_dict=model_dict, print_rate=0, phi_interp=interp)
    sim_state, epoch_store, loss_epochs = solve_fn(sim_state=sim_state)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        first, last = loss_epochs[0], loss_epochs[-1]
        f = lambda v: float(v[0]) if getattr(v, "ndim", 0) or isinstance(v, (list, tuple)) else float(v)
        print(f"{args.problem}: {n_tr}^3 training points, {args.epochs} epochs on {max(world, 1)} process(es) x "
              f"{torch.cuda.device_count() if (args.multi_gpu and world == 1) else 1} device(s): {dt:.2f} s wall including "
              f"set-up and evaluation; loss {f(first):.3e} -> {f(last):.3e}; "
              f"solution on the {args.n_eval}^3 evaluation grid in [{float(sim_state.solution.min()):.3e}, "
              f"{float(sim_state.solution.max()):.3e}]")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
