#!/usr/bin/env python
"""How much of the list chain hides beside the TMA stencil: event timings of stage subsets (one GPU).
    python tools/time_lists.py [grid]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jax_dips_b200 import plan as nplan, _cabi as cabi
from jax_dips_b200.trainer import haiku_init

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sys.argv = ["x", "--grid", str(grid)]
args = bench.parse()
dev = torch.device("cuda", 0)
problem = bench.make_problem("sphere")
tr, lv = bench.grids(problem, args, 1)
fns = bench.sim_fns(problem)
net = nplan.NetShape()
lvl = nplan.LevelSet(lv, fns.phi_fn(lv.R.to(dev)), interp="trilinear", perturb_eps=1e-10, device=dev)
nl = nplan.Nonlinear.coerce(None)
res = {}
for ov in (True, False):
    pl = nplan.SharedPlan(lvl, tr, 0, tr.shape()[0], fns, net, nl, nl, device=dev, overlap_lists=ov)
    nplan.upload_params(net, haiku_init(net, 42).to(dev))
    for name, mask in (("fwd", 1), ("fwd+stencil, no lists", 1 | 2 | 4 | 8 | 64), ("fwd+stencil+lists", 1 | 2 | 4 | 8),
                       ("whole", 0)):
        pl.step.stages = mask
        for _ in range(5):
            pl.loss_grad_launch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        a.record()
        for _ in range(n):
            pl.loss_grad_launch()
        b.record()
        torch.cuda.synchronize()
        res[(ov, name)] = a.elapsed_time(b) / n * 1e3
    pl.step.stages = 0
for k, v in res.items():
    print(f"overlap={k[0]!s:5} {k[1]:28s} {v:8.1f} us")
