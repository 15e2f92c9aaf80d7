#!/usr/bin/env python
"""Step time of ONE batch of the reference's default size (131072 points = 8 x-planes of a 128^3 grid) on every level of the
single-GPU schedule: eager kernel time of loss + gradient (no optimizer).   python tools/time_batches.py [grid] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jax_dips_b200 import plan as nplan
from jax_dips_b200.trainer import haiku_init

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
sys.argv = ["x", "--grid", str(grid)]
args = bench.parse()
dev = torch.device("cuda", 0)
problem = bench.make_problem("sphere")
tr, lv = bench.grids(problem, args, 1)
fns = bench.sim_fns(problem)
net = nplan.NetShape()
lvl = nplan.LevelSet(lv, fns.phi_fn(lv.R.to(dev)), interp="trilinear", perturb_eps=1e-10, device=dev)
nl = nplan.Nonlinear.coerce(None)
nplan.upload_params(net, haiku_init(net, 42).to(dev))
plane = grid * grid
n = grid ** 3
p0 = (n // 2 // batch) * batch          # a batch through the middle of the box (it holds interface)
for zoom in (0, 1, 2, 3):
    if zoom == 0:
        pl = nplan.SharedPlan(lvl, tr, p0 // plane, (p0 + batch) // plane, fns, net, nl, nl, device=dev)
    else:
        f = 0.5 ** zoom
        level = nplan.GeneralLevel(lvl, tr, (float(tr.dx) * f, float(tr.dy) * f, float(tr.dz) * f), fns, net, nl, nl, device=dev)
        pl = nplan.PointsPlan(level, p0, p0 + batch)
    for _ in range(5):
        pl.loss_grad_launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        pl.loss_grad_launch()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 50 * 1e3
    print(f"zoom {zoom}: batch of {batch} points of {grid}^3: {us:7.1f} us per loss+gradient, {batch / us * 1e6:.3e} points/s")
