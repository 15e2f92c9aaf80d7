"""Time K5 (`nbm_evaluate_f32`: u, grad u, d u/d n on the evaluation grid, trainer.py:960-977) on one GPU.

    python tools/bench_eval.py [n]        # n^3 evaluation points (default 256), level set on a 128^3 grid
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from jax_dips_b200 import _cabi as cabi, mesh, plan as nplan, problems
from jax_dips_b200 import numpy as jnp
from jax_dips_b200.trainer import haiku_init


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    P = problems.sphere()
    lo, hi = P.box
    ev = mesh.linspace_grid(lo, hi, [n] * 3)
    lv = mesh.linspace_grid(lo, hi, [128] * 3)
    with torch.cuda.device(dev):
        lvl = nplan.LevelSet(lv, jnp.vmap(P.phi_fn)(lv.R.to(dev)), device=dev)
        net = nplan.NetShape()
        nplan.upload_params(net, haiku_init(net, 42).to(dev))
        pts = ev.R.to(dev).contiguous()
        N = pts.shape[0]
        u = torch.empty(N, device=dev); gu = torch.empty(3 * N, device=dev); gn = torch.empty(N, device=dev)
        s = net.struct()
        call = lambda: cabi.check(cabi.lib().nbm_evaluate_f32(C.byref(s), C.byref(lvl.struct), cabi.ptr(pts), N, float(ev.dx),
                                                              float(ev.dy), float(ev.dz), cabi.ptr(u), cabi.ptr(gu), cabi.ptr(gn),
                                                              cabi.stream_ptr()))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
    # bytes: 12 B point in, 4 + 12 + 4 B out (the level-set gathers hit L2: 8.8 MB grid)
    print(json.dumps({"kernel": "evaluate_kernel", "points": N, "ms": ms, "points_per_s": N / (ms * 1e-3),
                      "GBps_algorithmic": 32.0 * N / (ms * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
