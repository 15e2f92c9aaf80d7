#!/usr/bin/env python
"""Wall time of the public API (setup / init_fn / solve_fn) on one GPU: the reference's single_GPU_train schedule (four
cell sizes, a quarter of the epochs each), plan building included.   python tools/time_trainer.py [grid] [epochs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jax_dips_b200 import mesh, problems, trainer as ntrainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 40
batch = int(sys.argv[3]) if len(sys.argv) > 3 else n ** 3    # 131072 = the reference's default batch size
P = problems.sphere()
lo, hi = P.box
tr, lv, ev = (mesh.linspace_grid(lo, hi, [k] * 3) for k in (n, 128, 64))
od = {"optimizer_name": "custom", "learning_rate": 1e-3, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
init_fn = ntrainer.setup(*P.setup_args())
for rep, ep in enumerate((epochs, epochs, 10 * epochs, epochs)):     # first run: CUDA context, kernels and allocator cold
    torch.cuda.synchronize()
    t0 = time.time()
    sim_state, solve_fn = init_fn(lvl_gstate=lv, tr_gstate=tr, eval_gstate=ev, num_epochs=ep, batch_size=batch,
                                  checkpoint_dir=None, optimizer_dict=od, print_rate=0, phi_interp="trilinear",
                                  use_cuda_graph=os.environ.get("NBM_TRAINER_GRAPH", "1") != "0")
    torch.cuda.synchronize()
    t1 = time.time()
    state, epoch_store, loss_epochs = solve_fn(sim_state)
    torch.cuda.synchronize()
    t2 = time.time()
    T = solve_fn.trainer
    print(f"run {rep}: sphere {n}^3, batches of {batch}, {ep} epochs (4 cell sizes): init_fn {t1 - t0:.3f} s, solve_fn {t2 - t1:.3f} s "
          f"(set-up of the 4 levels, training, evaluation on 64^3), {n ** 3 * ep / (t2 - t0):.3e} point-evaluations/s overall; "
          f"loss {float(loss_epochs[0]):.3e} -> {float(loss_epochs[-1]):.3e}")
