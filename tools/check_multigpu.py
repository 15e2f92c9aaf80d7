"""torchrun --nproc-per-node N tools/check_multigpu.py : the multi-GPU training loop (one process per GPU,
NCCL all-reduce of [grad, loss]) against the oracle's multi-device loop (psum = SUM of per-device means,
trainer.py:829-830) from the same initial parameters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import util
from jax_dips_b200 import mesh, problems, trainer as ntrainer
from oracle import nbm_oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = problems.sphere()
    n_tr, n_lvl, epochs = 16, 32, 4
    tr, lv, phi_grid, oprob = util.make_case(P, n_tr, n_lvl, "trilinear", torch.float64)
    p0 = O.init_params(oprob.shape, seed=42, dtype=torch.float64)
    od = {"optimizer_name": "custom", "learning_rate": 1e-2, "sched": {"scheduler_name": "exponential", "decay_rate": 0.975}}
    ev = mesh.linspace_grid(*P.box, [16] * 3)
    init_fn = ntrainer.setup(*P.setup_args())
    sim_state, solve_fn = init_fn(lvl_gstate=lv, tr_gstate=tr, eval_gstate=ev, num_epochs=epochs, batch_size=131072,
                                  multi_gpu=True, checkpoint_dir=None, optimizer_dict=od, init_params=p0.float(),
                                  print_rate=0, phi_interp="trilinear")
    (state, epoch_store, loss_epochs) = solve_fn(sim_state)
    T = solve_fn.trainer
    if rank == 0:
        grid_d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
        p_o, losses_o = O.multi_gpu_train(p0.clone(), tr.R.double(), grid_d, oprob, epochs, 131072, world, od)
        lk = torch.stack([l[0] for l in loss_epochs]).double()
        lo = torch.tensor(losses_o, dtype=torch.float64)
        e_l = float(((lk - lo).abs() / lo).max())
        e_p = util.rel_inf(T.params.cpu(), p_o)
        print(f"world {world}: loss trajectory rel err {e_l:.3e}, final params rel-inf {e_p:.3e}, "
              f"losses {lk.tolist()} oracle {losses_o}")
        assert e_l < 1e-3 and e_p < 1e-3
        assert len(loss_epochs[0]) == world
    # every rank must hold identical parameters
    allp = [torch.zeros_like(T.params) for _ in range(world)]
    dist.all_gather(allp, T.params)
    assert all(torch.equal(a, allp[0]) for a in allp)
    # ragged partition: 15 x 16 x 16 points do not split into whole x planes per rank -> per-point path + NCCL
    tr2 = mesh.linspace_grid(*P.box, [15, 16, 16])
    tr2o, lv2, phi2, oprob2 = util.make_case(P, [15, 16, 16], n_lvl, "trilinear", torch.float64)
    sim2, solve2 = init_fn(lvl_gstate=lv, tr_gstate=tr2, eval_gstate=ev, num_epochs=2, batch_size=700,
                           multi_gpu=True, checkpoint_dir=None, optimizer_dict=od, init_params=p0.float(), print_rate=0,
                           phi_interp="trilinear")
    (_, _, le2) = solve2(sim2)
    T2 = solve2.trainer
    assert T2.allreduce_kind == "nccl"
    if rank == 0:
        gd = [tr2.dx.double(), tr2.dy.double(), tr2.dz.double()]
        p_o2, l_o2 = O.multi_gpu_train(p0.clone(), tr2.R.double(), gd, oprob2, 2, 700, world, od)
        lk2 = torch.stack([l[0] for l in le2]).double()
        e2 = float(((lk2 - torch.tensor(l_o2, dtype=torch.float64)).abs() / torch.tensor(l_o2, dtype=torch.float64)).max())
        print(f"ragged: loss trajectory rel err {e2:.3e}, params rel-inf {util.rel_inf(T2.params.cpu(), p_o2):.3e}")
        assert e2 < 1e-3 and util.rel_inf(T2.params.cpu(), p_o2) < 1e-3
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU OK")


if __name__ == "__main__":
    main()
