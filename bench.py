#!/usr/bin/env python
"""Benchmark of the NBM training step (one optimizer step = residual + d loss/d params at every
training point + gradient all-reduce + optax chain) on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference, on the host cores

Prints ONE JSON line (rank 0).  metric = point-evaluations per second (BASELINE.json).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "point_evals_per_sec"
UNIT = "points/s"

# Algorithmic work per point-evaluation (DESIGN.md "Roofline"):
# default net p(3-10-10-1) | m(3-1-1), shared evaluation at native spacing
F_STEP = 1.0e3        # FLOP per point for the whole step (fwd 0.38k + bwd 0.55k + stencil 0.06k)
F_GRAD = 0.93e3       # FLOP per lattice node for the dominant kernel (forward recompute 0.38k + backward 0.55k)
# bytes per lattice node streamed by the two stencil passes: with the face table the residual
# pass moves cface 12 + dinv 4 + rhs 4 + U 4 + R 4 = 28 B and the adjoint pass cface 12 + dinv 4 + R 4 + G 4 = 24 B;
# with the 7-weight row table 40 B + 36 B (SURVEY section 8(d) quotes 64 B/point for a streamed K2 table)
B_STEP_FACES = 52.0
B_STEP_ROWS = 76.0
B_STEP_TMA = 36.0        # fused residual + adjoint: U, 3 face coefficients, 1/diag, rhs read once; R and G written
# DRAM bytes (read + write) of one node_grad launch at 256^3, `ncu --set full` (profiles/r2f_ncu_summary.md)
NODE_GRAD_TRAFFIC_256 = 160720384
NODE_GRAD_TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one node_grad launch at sphere 256^3, "
                            "`ncu --set full` capture profiles/r2f_ncu_full_raw.csv (not re-measured in this run)")
FP32_NOMINAL = 74.5                  # TFLOP/s: 148 SM x 128 lanes x 2 x 1.965 GHz


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sphere", choices=["sphere", "star", "stars", "dragon_like", "poisson_boltzmann"])
    ap.add_argument("--grid", type=int, default=256, help="training points per axis per GPU-slab (x grows with N: weak scaling)")
    ap.add_argument("--lvl", type=int, default=128)
    ap.add_argument("--interp", default="trilinear", choices=["trilinear", "quadratic", "analytic"],
                    help="level set: interpolant of the samples on the lvl grid, or (analytic) the callable itself")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="N > 1: strong = the named grid (256^3) sharded over the N GPUs (default, BASELINE.json's shape); "
                         "weak = --grid x-planes per GPU (the box is stretched in x)")
    ap.add_argument("--balance", type=int, default=1,
                    help="N > 1, strong scaling: 1 = cost-weighted x-slab boundaries (plan.balanced_slabs), 0 = equal slabs "
                         "(the reference's partition); the summed [grad, loss] is the same either way")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps when the working set fits in it")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement at N > 1")
    ap.add_argument("--zoom", type=int, default=0, help=">0: time the general per-point path at cell size = spacing/2^zoom")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="peer: partial-row reduction fused with the all-reduce over NVLink peer memory; nccl: ncclAllReduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--faces", type=int, default=-1, help="1/0: face-coefficient row table on/off (default: auto)")
    ap.add_argument("--stencil-tma", type=int, default=-1, help="1/0: residual + adjoint stencils as one TMA-fed kernel (default: auto)")
    ap.add_argument("--fused", type=int, default=-1, help="1/0: adjoint stencil fused into the gradient kernel (default: auto)")
    ap.add_argument("--precond", action="store_true",
                    help="train the learned preconditioner too (model_dict['preconditioner'], lpbe.yaml:62-67): 257 more "
                         "parameters, one more kernel per step")
    ap.add_argument("--emulate", default="", help="R/W: on ONE GPU, time the slab rank R would own in a W-GPU weak-scaling "
                                                  "run (load-balance diagnosis; not a bench line)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=12288)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def workload_string(args, world):
    """config.workload: the same string in both arms (ours and --impl reference) for the same flags"""
    nx = args.grid * world if args.scaling == "weak" else args.grid
    return (f"{args.workload}: train grid {nx}x{args.grid}x{args.grid} ({nx * args.grid * args.grid} points, x-slabs over "
            f"{world} GPU(s), {args.scaling} scaling), level set on {args.lvl}^3 lvl grid ({args.interp}), "
            "MLP p 3-10-10-1 | m 3-1-1 tanh, optimizer custom(adam), one batch per GPU")


def make_problem(name):
    from jax_dips_b200 import problems
    return problems.PROBLEMS[name]()


def grids(problem, args, world, scaling=None):
    from jax_dips_b200 import mesh
    lo, hi = problem.box
    nx = args.grid * world if (scaling or args.scaling) == "weak" else args.grid
    tr = mesh.linspace_grid(lo, hi, [nx, args.grid, args.grid])
    lv = mesh.linspace_grid(lo, hi, [args.lvl] * 3)
    return tr, lv


def sim_fns(problem):
    from jax_dips_b200 import numpy as jnp
    from jax_dips_b200.simulation_states import PoissonSimStateFn
    v = jnp.vmap
    return PoissonSimStateFn(v(problem.initial_value_fn), v(problem.dirichlet_bc_fn), v(problem.phi_fn),
                             v(problem.mu_m_fn), v(problem.mu_p_fn), v(problem.k_m_fn), v(problem.k_p_fn),
                             v(problem.f_m_fn), v(problem.f_p_fn), v(problem.alpha_fn), v(problem.beta_fn),
                             problem.nonlinear_op_m, problem.nonlinear_op_p)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (NVML, 10 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline(problem, args, n_sample, steps=1, warmup=0):
    """The oracle (CPU restatement of the reference: 197 network evaluations per point, JAX unavailable)
    on the host cores, on a bounded sample of the SAME workload: every (N/n_sample)-th training point."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import nbm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    tr, lv, phi_grid, oprob = util.make_case(problem, [args.grid] * 3, args.lvl,
                                             "trilinear" if args.interp == "analytic" else args.interp, torch.float32)
    if args.interp == "analytic":
        from jax_dips_b200 import numpy as jnp
        v = jnp.vmap(problem.phi_fn)
        oprob.phi_fn = lambda R: v(R.to(torch.float32)).to(R.dtype)
    n = tr.num_points()
    stride = max(1, n // n_sample)
    # build the sample without materialising the whole (n,3) point list
    idx = torch.arange(0, n, stride)[:n_sample]
    ny, nz = tr.shape()[1], tr.shape()[2]
    pts = torch.stack((tr.x[idx // (ny * nz)], tr.y[(idx // nz) % ny], tr.z[idx % nz]), dim=1)
    params = O.init_params(oprob.shape, seed=42)
    d = [tr.dx, tr.dy, tr.dz]
    for _ in range(warmup):
        O.loss_and_grad(params, pts[:256], *d, oprob)
    t0 = time.time()
    for _ in range(steps):
        O.loss_and_grad(params, pts, *d, oprob, chunk=4096)
    dt = (time.time() - t0) / steps
    return {"value": pts.shape[0] / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{pts.shape[0]} of the {n} training points (every {stride}-th), one loss+grad pass each; "
                      "torch-CPU restatement of the reference's literal formulation (JAX is not installed)",
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem = make_problem(args.workload)
    n_step = args.cpu_sample      # the same bounded sample as the cpu_baseline leg of our own arm
    res = cpu_baseline(problem, args, n_step, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args, int(os.environ.get("WORLD_SIZE", "1"))),
                       "sample": f"bounded sample of the workload: {n_step} of its training points per step (every k-th), "
                                 "on the host cores", "sample_points_per_step": n_step},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.scaling == "auto":
        args.scaling = "strong"
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from jax_dips_b200 import _cabi as cabi
    from jax_dips_b200 import plan as nplan
    from jax_dips_b200.optimizers import get_optimizer
    from jax_dips_b200.trainer import haiku_init

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = cabi.lib()

    problem = make_problem(args.workload)
    emu = tuple(int(v) for v in args.emulate.split("/")) if args.emulate else None
    tr, lv = grids(problem, args, emu[1] if emu else world)
    fns = sim_fns(problem)
    Nx, Ny, Nz = tr.shape()
    per = Nx // (emu[1] if emu else world)
    xa, xb = (emu[0] if emu else rank) * per, ((emu[0] if emu else rank) + 1) * per
    net = nplan.NetShape()
    precond = nplan.PrecondShape((8, 4), 1.0) if args.precond else None
    P = net.n_params + (precond.n_params if precond is not None else 0)
    if args.interp == "analytic":
        lvl = nplan.AnalyticLevelSet(lv, fns.phi_fn, device=dev)
    else:
        phi_lvl = fns.phi_fn(lv.R.to(dev))
        lvl = nplan.LevelSet(lv, phi_lvl, interp=args.interp, perturb_eps=1e-10, device=dev)
    t_setup = time.time()
    if args.zoom > 0:
        if world != 1:
            raise SystemExit("--zoom is a single-GPU measurement")
        f = 0.5 ** args.zoom
        level = nplan.GeneralLevel(lvl, tr, (float(tr.dx) * f, float(tr.dy) * f, float(tr.dz) * f), fns, net,
                                   nplan.Nonlinear.coerce(problem.nonlinear_op_m),
                                   nplan.Nonlinear.coerce(problem.nonlinear_op_p), device=dev)
        gp = nplan.PointsPlan(level, 0, tr.num_points())
        torch.cuda.synchronize()
        nplan.upload_params(net, haiku_init(net, 42).to(dev))
        for _ in range(3):
            gp.loss_grad_launch()
        torch.cuda.synchronize()
        z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        z0.record()
        for _ in range(args.steps):
            gp.loss_grad_launch()
        z1.record()
        torch.cuda.synchronize()
        msz = z0.elapsed_time(z1) / args.steps
        print(json.dumps({"metric": METRIC, "value": tr.num_points() / (msz * 1e-3), "unit": UNIT, "n_gpus": 1,
                          "ms_per_step": msz, "config": {"workload": f"{args.workload} {args.grid}^3 general path zoom {args.zoom}",
                                                         "crossed_sites": int(level.sites.n), "irregular_rows": int(level.n_irr)}}))
        return
    nl_m, nl_p = nplan.Nonlinear.coerce(problem.nonlinear_op_m), nplan.Nonlinear.coerce(problem.nonlinear_op_p)

    def make_plan(tr_, xa_, xb_, n_mean=None):
        return nplan.SharedPlan(lvl, tr_, xa_, xb_, fns, net, nl_m, nl_p, device=dev, n_mean=n_mean,
                                faces=None if args.faces < 0 else bool(args.faces),
                                fused=None if args.fused < 0 else bool(args.fused), precond=precond,
                                stencil_tma=None if args.stencil_tma < 0 else bool(args.stencil_tma))

    n_nominal = per * Ny * Nz           # every rank's mean runs over the nominal per-device batch (psum of means)
    slabs = None
    if world > 1 and args.scaling == "strong" and args.balance and not emu:
        slabs = nplan.balanced_slabs(lvl, tr, world, device=dev)
        xa, xb = slabs[rank]
    pl = make_plan(tr, xa, xb, n_mean=n_nominal)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    params0 = haiku_init(net, 42)
    if precond is not None:
        from jax_dips_b200.trainer import precond_init
        params0 = torch.cat((params0, precond_init(precond, 42)))
    params = params0.to(dev)
    opt_state = torch.zeros(2 * P, device=dev)
    opt_count = torch.zeros(1, dtype=torch.int32, device=dev)
    ostruct = cabi.Optimizer(P, 1e-3, 0.975, 1000.0, 1.0, 0.9, 0.999, 1e-8, 0, 0)
    comm = None
    if world > 1 and args.allreduce == "peer":
        from jax_dips_b200.comm import PeerComm
        try:
            comm = PeerComm(dev)
            ok = 1
        except Exception as exc:  # noqa: BLE001  (no CUDA IPC in this container, ...)
            print(f"# rank {rank}: peer all-reduce unavailable ({exc!r})", file=sys.stderr)
            comm, ok = None, 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks take the same path
        if int(flag.item()) == 0:
            comm = None

    net_struct = net.struct()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reset_state():
        params.copy_(params0.to(dev))
        opt_state.zero_()
        opt_count.zero_()
        nplan.upload_params(net, params)   # stages the initial parameters

    L2_MB = 126.0

    def plan_mb(p):
        """row tables + work arrays a step streams, per GPU"""
        return p.ne * ((16 if p.faces else 28) + 4 + 1 + 12) / 1e6

    class Stepper:
        """one optimizer step of plan `p` (what Trainer._step does): parameters into the constant bank, loss + gradient
        partial rows, [all-reduce], and ONE kernel for row reduction + optax chain + staging of the next step's
        parameter copies; replayed as a CUDA graph when possible"""

        def __init__(self, p):
            self.p, self.lg = p, p.loss_grad
            self.fits_l2 = plan_mb(p) <= 1.25 * L2_MB
            p.bind_params(params)
            self.graph_dev = self.graph_host = None
            if not args.no_graph and (world == 1 or comm is not None):
                try:
                    for _ in range(2):
                        self._step(True)
                    barrier()
                    self.graph_dev, self.graph_host = self._capture(True), self._capture(False)
                except Exception as exc:  # noqa: BLE001
                    if rank == 0:
                        print(f"# CUDA graph capture failed ({exc!r}); launching eagerly", file=sys.stderr)
                    self.graph_dev = self.graph_host = None
                    torch.cuda.synchronize()

        def _capture(self, staged):
            from jax_dips_b200.trainer import capture_graph      # the trainer's own capture (bare capture_begin / _end)
            return capture_graph(lambda: self._step(staged), dev)

        def _step(self, staged):
            if staged:
                return self._step_device(True)
            # the e2e leg: the host owns the parameters.  Pinned host -> device, the step, [grad, loss] and the updated
            # parameters device -> pinned host; captured, the three copies are memcpy nodes of the step's graph
            params.copy_(h_params, non_blocking=True)
            self._step_device(False)
            h_out.copy_(self.lg, non_blocking=True)
            h_params.copy_(params, non_blocking=True)   # the host keeps the parameters: next step's input

        def _step_device(self, staged):
            p = self.p
            if staged:   # device-resident training: the previous step staged the parameter copies
                cabi.check(L.nbm_upload_staged_params(cabi.stream_ptr()), "nbm_upload_staged_params")
            else:        # parameters handed in by the host every step (the e2e leg): prep kernel + copy
                nplan.upload_params(net, params)
            partials, rows = None, 0
            if comm is not None:   # exchange + optax chain + staging: one kernel
                p.loss_grad_launch(comm=comm, finalize=(ostruct, net_struct, params, opt_state, opt_count, None))
                return
            elif world > 1:
                p.loss_grad_launch()
                dist.all_reduce(self.lg, op=dist.ReduceOp.SUM)
            else:
                p.step.stages = 0x1f
                try:
                    p.loss_grad_launch()
                finally:
                    p.step.stages = 0
                partials, rows = p.partials, p.step.n_partial_rows
            cabi.check(L.nbm_finalize_step_f32(C.byref(ostruct), C.byref(net_struct), cabi.ptr(partials), rows, P + 1,
                                               cabi.ptr(self.lg), cabi.ptr(params), cabi.ptr(opt_state),
                                               cabi.ptr(opt_count), None, cabi.stream_ptr()), "nbm_finalize_step_f32")

        def step(self):
            self.graph_dev.replay() if self.graph_dev is not None else self._step(True)

        def step_host_params(self):
            self.graph_host.replay() if self.graph_host is not None else self._step(False)

        def time(self, fn, steps, warmup, sampler=None):
            """W untimed steps, barrier + synchronize, K timed steps bracketed by events, max over ranks (ms total).
            When the per-GPU working set fits in L2 (`flush_buf` set) the L2 is flushed before every step by writing a
            buffer twice its size; the flush is not timed (per-step event pairs, summed)."""
            for _ in range(warmup):
                fn()
            barrier()
            if sampler is not None:
                sampler.start()
            barrier()                      # nothing but the record sits between the last rendezvous and the first step
            if flush_buf is None or not self.fits_l2:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record()
                barrier()
                total = e0.elapsed_time(e1)
            else:
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                for a, b in evs:
                    flush_buf.fill_(1.0)
                    a.record()
                    fn()
                    b.record()
                barrier()
                total = sum(a.elapsed_time(b) for a, b in evs)
            t = torch.tensor([total], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

    # the timing rule: inputs larger than L2, or flush it between timed iterations
    table_mb = plan_mb(pl)
    flush_buf = None
    if table_mb <= 1.25 * L2_MB and not args.no_flush:
        flush_buf = torch.empty(int(2 * L2_MB * 1e6) // 4, dtype=torch.float32, device=dev)

    # pinned host buffers of the e2e leg (allocated before any capture: the graph's memcpy nodes hold their addresses)
    h_params = params.detach().cpu().pin_memory()
    h_out = torch.empty(P + 1).pin_memory()

    reset_state()
    st = Stepper(pl)
    used_graph = st.graph_dev is not None
    lg = st.lg

    # ---------------- parity of the exchange (N > 1), before anything is timed -----------------
    # (a) fused peer all-reduce vs NCCL all_reduce(SUM) of the un-reduced per-rank [grad, loss]; (b) every rank holds the
    # same bits; (c) slab additivity: the summed vector equals the whole-grid vector computed on ONE GPU (rank 0) with
    # the same 1/n (psum of per-device means, trainer.py:829-830).
    parity = None
    if world > 1 and not emu:
        reset_state()
        cabi.check(L.nbm_upload_staged_params(cabi.stream_ptr()), "nbm_upload_staged_params")
        own = pl.loss_grad_launch(out=torch.zeros(P + 1, device=dev)).clone()     # this rank's un-reduced vector
        pl.step.loss_grad = cabi.ptr(pl.loss_grad)
        ref = own.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        if comm is not None:
            got = pl.loss_grad_launch(comm=comm).clone()
        else:
            got = ref.clone()
        torch.cuda.synchronize()
        gathered = [torch.zeros_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        bitwise = all(torch.equal(g, gathered[0]) for g in gathered)
        scale = float(ref.abs().max())
        parity = {"peer_vs_nccl_rel": float((got - ref).abs().max()) / scale, "bitwise_equal_across_ranks": bool(bitwise),
                  "what": "fused peer all-reduce output vs dist.all_reduce(SUM) of the per-rank [grad, loss]; "
                          "rel = max|a-b| / max|b|"}
        if rank == 0 and args.scaling == "strong":
            whole = make_plan(tr, 0, Nx, n_mean=n_nominal)
            cabi.check(L.nbm_upload_staged_params(cabi.stream_ptr()), "nbm_upload_staged_params")
            w = whole.loss_grad_launch().clone()
            torch.cuda.synchronize()
            parity["slab_additivity_rel"] = float((got - w).abs().max()) / float(w.abs().max())
            parity["slab_additivity_loss_rel"] = abs(float(got[-1]) - float(w[-1])) / abs(float(w[-1]))
            del whole
            torch.cuda.empty_cache()
        flagp = torch.tensor([1 if (parity["peer_vs_nccl_rel"] <= 1e-6 and bitwise) else 0], device=dev)
        dist.all_reduce(flagp, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(int(flagp.item())) and parity.get("slab_additivity_rel", 0.0) <= 1e-4
        reset_state()

    # launches per step: constant-bank upload (memcpy node), fwd_nodes, residual, adjoint, node_grad, finalize
    # (+ 2 list kernels each for crossed sites / irregular rows); on several GPUs the peer all-reduce kernel in addition
    dense_launches = 3 if pl.stencil_tma_active else 4      # fwd_nodes, (residual + adjoint | stencil_tma), node_grad
    if pl.overlap_lists:     # extrap, extrap adjoint | irregular rows fwd + bwd | merge
        list_launches = (2 if pl.sites.n > 0 else 0) + (1 if pl.n_irr > 0 else 0) + 1
    else:
        list_launches = (1 if pl.sites.n > 0 else 0) * 2 + (1 if pl.n_irr > 0 else 0) * 2
    launches_per_step = (2 if (world > 1 and comm is None) else 1) + dense_launches + list_launches

    # ---------------- value: device-resident inputs -------------------------------------------
    sampler = ClockSampler(local)          # NVML initialised here, outside the timed window
    ms = st.time(st.step, args.steps, max(args.warmup, 3), sampler)
    clocks = sampler.stop()
    n_points_total = Nx * Ny * Nz
    value = n_points_total * args.steps / (ms * 1e-3)
    loss_now = float(lg[-1].item())

    # ---------------- e2e: the operator seam with HOST buffers --------------------------------
    # per step: the parameter vector from pinned host memory -> device, the step through the C ABI (prep kernel +
    # constant-bank upload + all kernels), then [grad, loss] and the updated parameters back to pinned host memory and a
    # stream synchronize (what a host framework that owns the parameters sees).  The grids and row tables are
    # per-level state, resident like the reference's jit constants.
    h_params.copy_(params)

    def e2e_step():
        st.step_host_params()          # H2D + step + 2 x D2H (one graph launch when the step is captured)
        torch.cuda.current_stream().synchronize()

    ms_e2e = st.time(e2e_step, args.steps, 2)
    e2e = {"value": n_points_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(h_params.numel() * 4),
           "d2h_bytes_per_step": int((P + 1) * 4 + P * 4),
           "what": "parameter vector H2D from pinned memory, step through the C ABI, [grad, loss] and updated "
                   "parameters D2H, stream sync, every step (the copies are memcpy nodes of the step's CUDA graph)"}

    # ---------------- secondary: weak scaling (the same per-GPU slab at every N) ---------------
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak and not emu:
        trw = grids(problem, args, world, "weak")[0]
        plw = make_plan(trw, rank * args.grid, (rank + 1) * args.grid)
        reset_state()
        stw = Stepper(plw)
        ms_w = stw.time(stw.step, args.steps, 3)
        nw = trw.num_points()
        weak = {"value": nw * args.steps / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w / args.steps,
                "workload": f"{trw.shape()[0]}x{trw.shape()[1]}x{trw.shape()[2]} ({args.grid} x-planes per GPU; box stretched "
                            "in x: not a BASELINE shape, kept for comparison with round 1)"}
        del stw, plw
        torch.cuda.empty_cache()
        reset_state()

    # ---------------- roofline of the dominant kernel (node_grad) -----------------------------
    roof = None
    cpu = None
    per_rank = None
    if world > 1:
        # per-rank device time of one step's kernels without the exchange (load balance), gathered to rank 0
        nplan.upload_params(net, params)
        for _ in range(2):
            pl.loss_grad_launch()
        torch.cuda.synchronize()
        samples = []
        for _ in range(7):
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            pl.loss_grad_launch()
            h1.record()
            torch.cuda.synchronize()
            samples.append(h0.elapsed_time(h1))
        tt = sorted(samples)[len(samples) // 2]     # median: eager launches, one slow sample must not read as imbalance
        allt = [torch.zeros(3, device=dev) for _ in range(world)]
        dist.all_gather(allt, torch.tensor([tt, float(pl.sites.n), float(pl.n_irr)], device=dev))
        per_rank = [{"rank": r, "ms_eager_no_exchange": float(a[0]), "crossed_sites": int(a[1]), "irregular_rows": int(a[2]),
                     "planes": (slabs[r][1] - slabs[r][0]) if slabs else per} for r, a in enumerate(allt)]
    if rank == 0:
        # measured FP32 FMA peak (MEASURED_PEAKS.json holds no FP32 figure)
        scratch = torch.zeros(4, device=dev)
        flops = C.c_double(0.0)
        L.nbm_ffma_probe_f32(64, cabi.ptr(scratch), C.byref(flops), cabi.stream_ptr())
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            L.nbm_ffma_probe_f32(4096, cabi.ptr(scratch), C.byref(flops), cabi.stream_ptr())
            g1.record()
            torch.cuda.synchronize()
            best = max(best, flops.value / (g0.elapsed_time(g1) * 1e-3) / 1e12)
        ffma_peak = best
        # per-stage device time, live, on the launching stream
        stage_ms = {}
        names = {1: "fwd_nodes", 2: "extrap", 4: "residual", 8: "adjoint", 16: "node_grad"}
        if pl.stencil_tma_active:
            # residual rows + adjoint stencil are one kernel; the list kernels (irregular rows, extrapolation adjoint) follow it
            names = {1: "fwd_nodes", 2: "extrap", 4 | 8 | 64: "stencil_tma", 4 | 8 | 128: "lists", 16: "node_grad",
                     4 | 64: "residual_alone", 8 | 64: "adjoint_alone"}
        nplan.upload_params(net, params)
        pl.step.stages = 0
        pl.loss_grad_launch()
        for bit, nm in names.items():
            pl.step.stages = bit
            tot = 0.0
            for _ in range(min(args.steps, 20)):
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                pl.loss_grad_launch()
                h1.record()
                torch.cuda.synchronize()
                tot += h0.elapsed_time(h1)
            stage_ms[nm] = tot / min(args.steps, 20)
        pl.step.stages = 0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ne = pl.ne
        ach = F_GRAD * ne / (stage_ms["node_grad"] * 1e-3) / 1e12
        step_ach = F_STEP * n_points_total / ((ms / args.steps) * 1e-3) / 1e12 / world
        b_step = B_STEP_FACES if pl.faces else B_STEP_ROWS
        t_stencil = stage_ms["residual"] + stage_ms["adjoint"] if "residual" in stage_ms else stage_ms["stencil_tma"]
        if pl.stencil_tma_active:
            b_step = B_STEP_TMA + (4.0 if pl.kv is not None else 0.0) + (8.0 if pl.nl is not None else 0.0)
        is_256 = (args.workload == "sphere" and args.grid == 256 and world == 1)
        roof = {"bound": "fp32", "kernel": "node_grad_kernel", "achieved": ach, "peak": ffma_peak, "unit": "TFLOP/s",
                "frac": ach / ffma_peak if ffma_peak else None,
                "peak_nominal": FP32_NOMINAL, "frac_nominal": ach / FP32_NOMINAL,
                "traffic": NODE_GRAD_TRAFFIC_256 if is_256 else None,
                "traffic_source": NODE_GRAD_TRAFFIC_SOURCE if is_256 else None,
                "peak_source": "FFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP32 figure; "
                               "nominal 148 SM x 128 x 2 x 1.965 GHz = 74.5)",
                "algorithmic_flop_per_node": F_GRAD, "nodes_per_launch": ne,
                "stage_ms": stage_ms,
                "step": {"achieved": step_ach, "per": "GPU", "frac": step_ach / ffma_peak if ffma_peak else None,
                         "frac_nominal": step_ach / FP32_NOMINAL, "algorithmic_flop_per_point": F_STEP},
                "hbm": {"achieved": b_step * ne / (t_stencil * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s",
                        "kernels": "stencil_tma (residual + adjoint, one kernel)" if pl.stencil_tma_active else "residual + adjoint",
                        "algorithmic_bytes_per_node": b_step,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"}}
        roof["hbm"]["frac"] = roof["hbm"]["achieved"] / hbm_peak
        if not args.no_cpu_baseline and world == 1:
            r = cpu_baseline(problem, args, args.cpu_sample, steps=1, warmup=1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args, world),
                           "slabs": ([b - a for a, b in slabs] if slabs else f"{per} x-planes per GPU (equal blocks)"),
                           "l2": (f"row tables + work arrays ({table_mb:.0f} MB per GPU) exceed the 126 MB L2; no flush needed"
                                  if flush_buf is None else
                                  f"row tables + work arrays are {table_mb:.0f} MB per GPU (fit in the 126 MB L2): L2 flushed "
                                  "before every timed step by writing a 252 MB buffer (untimed; per-step CUDA events, summed)"),
                           "row_table": "faces (3 face coefficients + 1/diag per node)" if pl.faces else "7 row weights per node",
                           "adjoint": ("fused into the gradient kernel (TMA ring)" if pl.fused else
                                       ("residual rows + adjoint stencil in one kernel fed by 3-D TMA boxes (T never leaves the SM)"
                                        if pl.stencil_tma_active else "separate stencil pass")),
                           "lists": ("side stream beside the dense stencil kernel(s), merged into G by one kernel" if pl.overlap_lists
                                     else "in line after the dense stencil"),
                           "cuda_graph": used_graph,
                           "allreduce": ("peer-memory kernel fused with the partial reduction" if comm is not None
                                         else ("nccl" if world > 1 else "none")),
                           "crossed_sites": int(pl.sites.n), "irregular_rows": int(pl.n_irr),
                           "setup_seconds": t_setup, "loss": loss_now,
                           "tolerances": "CUDA vs oracle: flags exact, fractions 1e-5 of the cell measure (1e-4 vs the "
                                         "reference-generated goldens: the reference's own f32/x64 spread is 3e-5 on the "
                                         "12-23-cell sets and 6.2e-5 on the 220-cell sets), rows 1e-5, loss and gradient 1e-4 "
                                         "(tests/, normwise)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "roofline": roof, "cpu_baseline": cpu}
        if parity is not None:
            line["parity_check"] = parity
        if weak is not None:
            line["weak_scaling"] = weak
        if per_rank is not None:
            line["per_rank"] = per_rank
        print(json.dumps(line), flush=True)
    if comm is not None:
        if comm.error():
            raise SystemExit("peer all-reduce reported a timeout")
        comm.close()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"parity check of the gradient exchange failed: {parity}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
