#!/usr/bin/env python
"""Time of the TMA stencil kernel alone (stage mask residual + adjoint, dense only) for the tile shape chosen by
choose_geom or forced through NBM_ST_* (tools/sweep: one process per shape).   python tools/time_stencil.py [grid] [nx]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jax_dips_b200 import plan as nplan
from jax_dips_b200.trainer import haiku_init

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nxp = int(sys.argv[2]) if len(sys.argv) > 2 else grid     # planes of the slab (strong-scaling shapes)
sys.argv = ["x", "--grid", str(grid)]
args = bench.parse()
dev = torch.device("cuda", 0)
problem = bench.make_problem("sphere")
tr, lv = bench.grids(problem, args, 1)
fns = bench.sim_fns(problem)
net = nplan.NetShape()
lvl = nplan.LevelSet(lv, fns.phi_fn(lv.R.to(dev)), interp="trilinear", perturb_eps=1e-10, device=dev)
nl = nplan.Nonlinear.coerce(None)
pl = nplan.SharedPlan(lvl, tr, 0, nxp, fns, net, nl, nl, device=dev)
nplan.upload_params(net, haiku_init(net, 42).to(dev))
pl.loss_grad_launch()
out = {}
for name, mask in (("stencil", 4 | 8 | 64), ("fwd", 1), ("node_grad", 16)):
    pl.step.stages = mask
    for _ in range(5):
        pl.loss_grad_launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        pl.loss_grad_launch()
    b.record()
    torch.cuda.synchronize()
    out[name] = a.elapsed_time(b) / 40 * 1e3
print(" ".join(f"{k} {v:7.1f} us" for k, v in out.items()), {k: v for k, v in os.environ.items() if k.startswith("NBM_ST")})
