"""Drop-in for `jax_dips.solvers.poisson.trainer`: `setup -> init_fn -> solve_fn` with the same
names, keyword arguments, defaults, dict schemas and return shapes (trainer.py:980-1137), running
the NBM step on hand-written sm_100a CUDA kernels through the C ABI in include/nbm_b200.h.

Differences a user of the reference must know (all documented in DESIGN.md):

* coefficient callables are the reference's per-point callables written against
  `jax_dips_b200.numpy` (torch) instead of `jax.numpy`; they are batched by calling them once on
  the (3, n) view of the point list (the role `vmap` plays at trainer.py:995-1005);
* the level set: by default (`phi_interp="analytic"`) the user's callable itself is the level set
  everywhere, as in the reference (discretization.py:90; sampled on the host at the positions the
  kernels need).  `phi_interp="trilinear"` (interpolate.py:906) / `"quadratic"` (:388) sample it once on
  `lvl_gstate` and let the kernels read it through the reference's own grid interpolant - what a
  reference user gets by passing such an interpolant as `lvl_set_fn` (examples/dragon, solve_dragon.py:177);
* `nonlinear_op_m/p` must be None / zero / `Nonlinear.sinh(coef)`;
* the initial parameters follow haiku's initialisers (TruncatedNormal(0.1) hidden, 1/sqrt(fan_in)
  output, zero bias; MLP.py:65,70) but are drawn from a torch generator seeded with 42, because
  jax's PRNGKey(42) stream cannot be reproduced without jax; pass `init_params=` to pin them;
* multi-GPU = one process per GPU (`torchrun`), gradient `psum` = NCCL all-reduce(SUM) of the
  168-float [grad, loss] buffer (trainer.py:829-830);
* only `algorithm=0`, `model_type="mlp"`, optimizers "custom" / "adam" / "rmsprop" exist on this
  path; everything else raises the reference's exception types.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import os
import pickle
import signal
import time
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import _cabi as cabi
from . import data_management
from . import numpy as jnp
from .optimizers import OptimizerSpec, get_optimizer
from .plan import (AnalyticLevelSet, EmptyPlan, GeneralLevel, LevelSet, NetShape, Nonlinear, PointsPlan, PrecondShape,
                   SharedPlan, upload_params)
from .simulation_states import PoissonSimState, PoissonSimStateFn, replace

logger = logging.getLogger(__name__)

stop_training = False


def signalHandler(signal_num, frame):
    """trainer.py:68-78: SIGINT asks the multi-GPU loop to stop after the current epoch."""
    global stop_training
    stop_training = True
    logger.warning("Signal: %s. Training will stop after the completion of current epoch", signal_num)


def install_sigint_handler():
    signal.signal(signal.SIGINT, signalHandler)


_DEFAULT_OPT = {"optimizer_name": "custom", "learning_rate": 1e-3,
                "sched": {"scheduler_name": "exponential", "decay_rate": 0.9}}
_DEFAULT_MODEL = {
    "name": None, "model_type": "mlp",
    "mlp": {"hidden_layers_m": 1, "hidden_dim_m": 1, "activation_m": "jnp.tanh",
            "hidden_layers_p": 2, "hidden_dim_p": 10, "activation_p": "jnp.tanh"},
    "resnet": {"res_blocks_m": 3, "res_dim_m": 40, "activation_m": "nn.tanh",
               "res_blocks_p": 3, "res_dim_p": 80, "activation_p": "nn.tanh"},
}


def capture_graph(fn, device) -> "torch.cuda.CUDAGraph":
    """Capture `fn()` (C-ABI launches on the current stream, no torch allocations) into a CUDA graph on a side stream.
    `torch.cuda.graph(...)` would also synchronise the device, run the garbage collector and empty the caching AND the
    pinned-host allocators on every capture (26 ms each measured): a training loop with 16 batches x 4 levels captures
    64 graphs, so the bare capture calls are used."""
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    torch.cuda.current_stream(device).wait_stream(side)
    return g


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise cabi.NbmError("no CUDA device: the NBM path runs only on the GPU (no CPU fallback)")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())


def haiku_init(net: NetShape, seed: int = 42) -> torch.Tensor:
    """Flat parameter vector, C-ABI order (p head then m head; W (in,out) row-major then b)."""
    g = torch.Generator().manual_seed(seed)

    def trunc(n, std):
        out = torch.empty(n, dtype=torch.float64)
        torch.nn.init.trunc_normal_(out, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return out * std

    parts = []
    for (L, H) in ((net.layers_p, net.hidden_p), (net.layers_m, net.hidden_m)):
        fan_in = 3
        for _ in range(L):
            parts += [trunc(fan_in * H, 0.1), torch.zeros(H, dtype=torch.float64)]
            fan_in = H
        parts += [trunc(fan_in, 1.0 / math.sqrt(fan_in)), torch.zeros(1, dtype=torch.float64)]
    return torch.cat(parts).to(torch.float32)


def precond_init(pc: PrecondShape, seed: int = 42) -> torch.Tensor:
    """glorot_uniform kernels, zero biases (nn/preconditioner.py:20; flax nn.Dense defaults), torch generator."""
    g = torch.Generator().manual_seed(seed + 1000)
    parts = []
    for (fan_in, d) in pc.layer_dims():
        lim = math.sqrt(6.0 / (fan_in + d))
        parts += [(torch.rand(fan_in * d, generator=g, dtype=torch.float64) * 2 - 1) * lim,
                  torch.zeros(d, dtype=torch.float64)]
    return torch.cat(parts).to(torch.float32)


def params_to_tree(net: NetShape, flat: torch.Tensor, precond: Optional[PrecondShape] = None) -> dict:
    """haiku-style parameter tree (keys as the reference's checkpoints hold them, cf. trainer.py:323)."""
    flat = flat.detach().cpu().numpy()
    tree, off = {}, 0
    for head, (L, H) in (("p", (net.layers_p, net.hidden_p)), ("m", (net.layers_m, net.hidden_m))):
        fan_in = 3
        for l in range(L + 1):
            out = H if l < L else 1
            name = f"double_mlp/~mlp_{head}_fn/linear" + ("" if l == 0 else f"_{l}")
            w = flat[off: off + fan_in * out].reshape(fan_in, out).copy(); off += fan_in * out
            b = flat[off: off + out].copy(); off += out
            tree[name] = {"w": w, "b": b}
            fan_in = out
    tree["preconditioner"] = {}
    if precond is not None:
        # flax parameter tree of nn/preconditioner.py (trainer.py:236-243)
        dense = {}
        for l, (fan_in, d) in enumerate(precond.layer_dims()):
            k = flat[off: off + fan_in * d].reshape(fan_in, d).copy(); off += fan_in * d
            b = flat[off: off + d].copy(); off += d
            dense[f"Dense_{l}"] = {"kernel": k, "bias": b}
        tree["preconditioner"] = {"params": dense}
    return tree


def tree_to_params(net: NetShape, tree: dict, precond: Optional[PrecondShape] = None) -> torch.Tensor:
    parts = []
    for head, (L, H) in (("p", (net.layers_p, net.hidden_p)), ("m", (net.layers_m, net.hidden_m))):
        for l in range(L + 1):
            name = f"double_mlp/~mlp_{head}_fn/linear" + ("" if l == 0 else f"_{l}")
            parts += [np.asarray(tree[name]["w"], dtype=np.float32).reshape(-1),
                      np.asarray(tree[name]["b"], dtype=np.float32).reshape(-1)]
    if precond is not None:
        dense = tree["preconditioner"]["params"]
        for l in range(len(precond.layer_dims())):
            parts += [np.asarray(dense[f"Dense_{l}"]["kernel"], dtype=np.float32).reshape(-1),
                      np.asarray(dense[f"Dense_{l}"]["bias"], dtype=np.float32).reshape(-1)]
    return torch.from_numpy(np.concatenate(parts))


class Trainer:
    """Trainer for the NBM Poisson solver (trainer.py:81-977): owns the level set, the per-level
    plans (row tables in HBM), the parameters, the optimizer state and the training loops."""

    def __init__(self, lvl_gstate, tr_gstate, eval_gstate, sim_state: PoissonSimState,
                 sim_state_fn: PoissonSimStateFn, algorithm: int = 0, mgrad_over_pgrad_scalefactor: int = 1,
                 lvl_set_fn: Callable = None, num_epochs: int = 1000, multi_gpu: bool = False,
                 batch_size: int = 131072, checkpoint_dir: str = "./checkpoints", checkpoint_interval: int = 2,
                 results_dir: str = "./", loss_plot_name: str = "solver_loss", optimizer_dict: dict = None,
                 restart: bool = False, restart_checkpoint_dir: str = "./checkpoints", print_rate: int = 1,
                 model_dict: dict = None, phi_interp: str = "analytic", perturb_eps: float = 1e-10,
                 device=None, init_params: Optional[torch.Tensor] = None, use_cuda_graph: bool = True,
                 allreduce: str = "peer", n_devices: Optional[int] = None):
        global stop_training
        stop_training = False   # a SIGINT that stopped an earlier Trainer of this process must not stop this one
        if algorithm != 0:
            # discretization.py:146-148 references undefined attributes for algorithm=1
            raise NotImplementedError("only algorithm=0 (regression extrapolation) exists on this path")
        optimizer_dict = optimizer_dict or _DEFAULT_OPT
        model_dict = model_dict or _DEFAULT_MODEL
        self._optimizer_dict, self._model_dict = optimizer_dict, model_dict
        self._phi_interp, self._perturb_eps = phi_interp, perturb_eps
        self.device = torch.device(device) if device is not None else default_device()
        self.lvl_gstate, self.tr_gstate, self.eval_gstate = lvl_gstate, tr_gstate, eval_gstate
        self.sim_state, self.sim_state_fn = sim_state, sim_state_fn
        self.batch_size, self.num_epochs, self.multi_gpu = batch_size, num_epochs, multi_gpu
        self.checkpoint_dir, self.checkpoint_interval = checkpoint_dir, checkpoint_interval
        self.results_dir, self.loss_plot_name, self.print_rate = results_dir, loss_plot_name, print_rate
        # the reference stores this factor and reads it only in the private `__update` (trainer.py:791-816), which none
        # of its training loops call: `update` / `update_multi_gpu` ignore it.  Same here: the loops ignore it (and say
        # so), `update_region_scaled` is that private step for callers who want it.
        self.mgrad_over_pgrad_scalefactor = mgrad_over_pgrad_scalefactor
        if mgrad_over_pgrad_scalefactor != 1:
            logger.warning("mgrad_over_pgrad_scalefactor=%s does not enter the training loops (as in the reference, whose "
                           "loops never call `__update`, trainer.py:791-816); use Trainer.update_region_scaled for that step",
                           mgrad_over_pgrad_scalefactor)
        self.restart_checkpoint_dir = restart_checkpoint_dir
        self.use_cuda_graph = use_cuda_graph
        self.n_devices = n_devices       # single-process multi_gpu: how many local devices (default: all visible)
        if allreduce not in ("peer", "nccl"):
            raise ValueError("allreduce must be 'peer' (fused NVLink peer-memory kernel) or 'nccl'")
        self.allreduce_kind = allreduce
        self._comm = None
        cabi.lib()  # fail here, loudly, if the CUDA library is not built

        self.TD = data_management.TrainData(tr_gstate, lvl_set_fn, refine=False, refine_lod=False,
                                            refine_normals=False, v_cycle_period=2, rest_at_level=100)
        self.train_dx, self.train_dy, self.train_dz = (float(tr_gstate.dx), float(tr_gstate.dy), float(tr_gstate.dz))

        self.net = NetShape.from_model_dict(model_dict)
        # learned preconditioner (nn/preconditioner.py; trainer.py:229-243): parameters ride at the tail of the
        # flat vector and are trained jointly
        self.precond = PrecondShape.from_model_dict(model_dict)
        self.nonlinear_m = Nonlinear.coerce(sim_state_fn.nonlinear_op_m)
        self.nonlinear_p = Nonlinear.coerce(sim_state_fn.nonlinear_op_p)

        name = optimizer_dict["optimizer_name"]
        self.optimizer_name = name
        if name == "lbfgs":
            # trainer.py:197-208: no optax chain; solve = solve_jaxopt (scipy L-BFGS-B on the first batch)
            self.optimizer = None
            self.solve = self.solve_jaxopt
        else:
            self.optimizer: OptimizerSpec = get_optimizer(
                optimizer_name=name, scheduler_name=optimizer_dict["sched"]["scheduler_name"],
                learning_rate=optimizer_dict["learning_rate"], decay_rate=optimizer_dict["sched"]["decay_rate"])
            self.solve = self.solve_optax

        with torch.cuda.device(self.device):
            # level set on the lvl grid, read by the kernels through the reference's interpolant
            if phi_interp == "analytic":
                # the user's callable is the level set everywhere, as in the reference (discretization.py:90)
                self.lvl = AnalyticLevelSet(lvl_gstate, sim_state_fn.phi_fn, device=self.device)
            else:
                phi_lvl = sim_state_fn.phi_fn(lvl_gstate.R.to(self.device))
                self.lvl = LevelSet(lvl_gstate, phi_lvl, interp=phi_interp, perturb_eps=perturb_eps, device=self.device)
            P = self.n_params = self.net.n_params + (self.precond.n_params if self.precond is not None else 0)
            self.opt_state = torch.zeros(2 * P, dtype=torch.float32, device=self.device)
            self.opt_count = torch.zeros(1, dtype=torch.int32, device=self.device)
            if restart:
                state = self.fetch_checkpoint(self.restart_checkpoint_dir)
                if state is None:
                    raise FileNotFoundError(f"no checkpoint under {self.restart_checkpoint_dir}")
                self.params = tree_to_params(self.net, state["params"], self.precond).to(self.device)
                ost = state.get("opt_state") or {}
                if "moments" in ost:     # checkpoints of the lbfgs path carry scipy's result instead (warm start)
                    self.opt_state.copy_(torch.as_tensor(ost["moments"]).to(self.device))
                    self.opt_count.fill_(int(ost.get("count", 0)))
                self.batch_size = state["batch_size"]
                logger.info(f"Resuming training from epoch {state['epoch']} with batch_size {self.batch_size}, "
                            f"resolution {state['resolution']}.")
            else:
                p0 = init_params if init_params is not None else haiku_init(self.net, seed=42)
                if self.precond is not None and p0.numel() == self.net.n_params:
                    p0 = torch.cat((p0.detach().cpu().float(), precond_init(self.precond, seed=42)))
                if p0.numel() != P:
                    raise ValueError(f"init_params has {p0.numel()} entries, the network has {P}")
                self.params = p0.detach().to(self.device, torch.float32).contiguous().clone()
        self.epoch_start = 0  # trainer.py:261 (resume = warm start only)
        self.loss_epochs = torch.zeros(self.num_epochs - self.epoch_start)
        self.epoch_store = np.arange(self.epoch_start, self.num_epochs)
        self._plans = {}
        self._levels = {}
        self._graphs = {}
        self._opt_struct = None

    # ------------------------------------------------------------------------------------------
    # checkpoints (trainer.py:325-351): same dict keys
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def fetch_checkpoint(checkpoint_dir):
        if checkpoint_dir is None or not os.path.exists(checkpoint_dir):
            return None
        checkpoints = [p for p in os.listdir(checkpoint_dir) if "checkpoint_" in p]
        if not checkpoints:
            return None
        # the reference takes the lexicographic max (trainer.py:331-334); numeric max is what is meant
        checkpoint = os.path.join(checkpoint_dir, max(checkpoints, key=lambda p: int(p.split("_")[-1])))
        logger.info(f"Loading checkpoint {checkpoint}")
        with open(checkpoint, "rb") as f:
            return pickle.load(f)

    @staticmethod
    def save_checkpoint(checkpoint_dir, state):
        if checkpoint_dir is None:
            logger.info("No checkpoint dir. specified. Skipping checkpoint.")
            return None
        os.makedirs(checkpoint_dir, exist_ok=True)
        checkpoint = os.path.join(checkpoint_dir, "checkpoint_" + str(state["epoch"]))
        logger.info(f"Saving checkpoint {checkpoint}")
        with open(checkpoint, "wb") as f:
            pickle.dump(state, f)
        return checkpoint

    def _checkpoint_state(self, epoch: int) -> dict:
        return {"opt_state": {"moments": self.opt_state.cpu().numpy(), "count": int(self.opt_count.item())},
                "params": params_to_tree(self.net, self.params, self.precond), "epoch": epoch, "batch_size": self.batch_size,
                "resolution": f"{self.train_dx}, {self.train_dy}, {self.train_dz}"}

    # ------------------------------------------------------------------------------------------
    # plans: one per (zoom level, batch range); built lazily, kept in HBM
    # ------------------------------------------------------------------------------------------
    def plan_for(self, zoom: int, p0: int, p1: int, n_mean: Optional[int] = None):
        """`n_mean`: number of points the mean of the loss runs over when it is not the batch's own size (cost-balanced
        slabs of the multi-GPU loops: every device divides by the nominal per-device batch size)"""
        key = (zoom, p0, p1, n_mean)
        if key in self._plans:
            return self._plans[key]
        Nx, Ny, Nz = self.tr_gstate.shape()
        plane = Ny * Nz
        with torch.cuda.device(self.device):
            if p1 <= p0:
                pl = EmptyPlan(self.n_params, self.device)
            elif zoom == 0 and p0 % plane == 0 and p1 % plane == 0:
                pl = SharedPlan(self.lvl, self.tr_gstate, p0 // plane, p1 // plane, self.sim_state_fn, self.net,
                                self.nonlinear_m, self.nonlinear_p, device=self.device, precond=self.precond,
                                n_mean=n_mean)
                pl.bind_params(self.params)
            else:
                if n_mean is not None:
                    raise ValueError("n_mean is a property of whole-plane batches on the shared path")
                pl = PointsPlan(self.general_level(zoom), p0, p1)
                pl.bind_params(self.params)
        self._plans[key] = pl
        return pl

    def _device_ranges(self, DD, world: int):
        """per device the list of (p0, p1, n_mean) batches.  The reference's partition (contiguous equal blocks,
        data_management.py:121-130) unless every device holds ONE batch of whole x planes and no preconditioner is
        trained: then the slab boundaries are cost-weighted (`plan.balanced_slabs`; NBM_BALANCE_SLABS=0 keeps equal
        slabs) and every device divides by the nominal batch size, which leaves the summed [grad, loss] unchanged."""
        plane = self.tr_gstate.shape()[1] * self.tr_gstate.shape()[2]
        ref = [[(p0, p1, None) for (p0, p1) in DD.ranges(r)] for r in range(world)]
        one_batch = all(len(rr) == 1 and rr[0][1] > rr[0][0] and rr[0][0] % plane == 0 and rr[0][1] % plane == 0 for rr in ref)
        whole = sum(rr[0][1] - rr[0][0] for rr in ref) == self.tr_gstate.num_points() if one_batch else False
        if (world > 1 and one_batch and whole and self.precond is None
                and os.environ.get("NBM_BALANCE_SLABS", "1") != "0"):
            from .plan import balanced_slabs
            n_nom = ref[0][0][1] - ref[0][0][0]
            if all(rr[0][1] - rr[0][0] == n_nom for rr in ref):
                with torch.cuda.device(self.device):
                    slabs = balanced_slabs(self.lvl, self.tr_gstate, world, device=self.device)
                return [[(xa * plane, xb * plane, n_nom)] for (xa, xb) in slabs]
        return ref

    def general_level(self, zoom: int) -> GeneralLevel:
        if zoom not in self._levels:
            with torch.cuda.device(self.device):
                self._levels[zoom] = GeneralLevel(self.lvl, self.tr_gstate, self.TD.zoom_cell(zoom), self.sim_state_fn,
                                                  self.net, self.nonlinear_m, self.nonlinear_p, device=self.device,
                                                  precond=self.precond)
        return self._levels[zoom]

    def _warn_if_padded(self, DD) -> None:
        if DD.padded and not getattr(self, "_warned_padding", False):
            self._warned_padding = True
            logger.warning("%d points do not fold into %d device(s) x batches of %d: the reference pads the short batch "
                           "with random points from jax PRNGKey(0) (data_management.py:70-76); here it is trained on its "
                           "real points only", DD._len, DD.num_gpus, DD.batch_size)

    def _optimizer_struct(self) -> cabi.Optimizer:
        if self._opt_struct is None:
            o, s = self.optimizer, self.optimizer.scheduler
            self._opt_struct = cabi.Optimizer(self.n_params, float(o.learning_rate), float(s.decay_rate),
                                              float(s.transition_steps), float(o.max_norm), o.b1,
                                              0.9 if o.optimizer_name == "rmsprop" else o.b2, o.eps, o.kind,
                                              0 if s.scheduler_name == "exponential" else 1)
        return self._opt_struct

    # ------------------------------------------------------------------------------------------
    # the operator seam: loss and d loss/d params (Trainer.loss + value_and_grad, trainer.py:786, 893)
    # ------------------------------------------------------------------------------------------
    def loss_and_grad(self, params: torch.Tensor, plan) -> torch.Tensor:
        """[grad(P), loss] on the device for the rows of `plan` (enqueued on the current stream)."""
        upload_params(self.net, params)
        return plan.loss_grad_launch()

    def _step(self, plan, loss_hist: Optional[torch.Tensor], allreduce: bool, local_comm=None):
        """update (trainer.py:783-789) / update_multi_gpu (:824-834) on the current stream.  Inside the training loops
        the parameters reach the constant bank from the copies the previous step's finalize kernel staged
        (`_begin_training` stages the initial ones); the partial-row reduction, the optax chain and that staging are one
        kernel.  `local_comm`: this device's handle of an in-process peer exchange (single-process multi-device)."""
        L = cabi.lib()
        cabi.check(L.nbm_upload_staged_params(cabi.stream_ptr()), "nbm_upload_staged_params")
        net = self.net.struct()
        partials, rows = None, 0
        d = _dist() if allreduce else None
        # (peer exchange: partial rows -> psum over NVLink peer memory -> optax chain -> staging is ONE kernel)
        fin = (self._optimizer_struct(), net, self.params, self.opt_state, self.opt_count, loss_hist)
        if local_comm is not None:
            plan.loss_grad_launch(comm=local_comm, finalize=fin)
            return
        elif d is not None and d.get_world_size() > 1:
            comm = self._peer_comm() if (self.allreduce_kind == "peer" and isinstance(plan, (SharedPlan, EmptyPlan))) else None
            if comm is not None:
                plan.loss_grad_launch(comm=comm, finalize=fin)
                return
            else:
                lg = plan.loss_grad_launch()
                d.all_reduce(lg, op=d.ReduceOp.SUM)      # psum of grads and loss (:829-830)
        elif isinstance(plan, SharedPlan):
            plan.step.stages = 0x1f                      # everything but the reduction: the finalize kernel sums the rows
            try:
                lg = plan.loss_grad_launch()
            finally:
                plan.step.stages = 0
            partials, rows = plan.partials, plan.step.n_partial_rows
        else:
            lg = plan.loss_grad_launch()
        cabi.check(L.nbm_finalize_step_f32(C.byref(self._optimizer_struct()), C.byref(net), cabi.ptr(partials), rows,
                                           self.n_params + 1, cabi.ptr(lg), cabi.ptr(self.params),
                                           cabi.ptr(self.opt_state), cabi.ptr(self.opt_count), cabi.ptr(loss_hist),
                                           cabi.stream_ptr()), "nbm_finalize_step_f32")

    def update_region_scaled(self, plan, loss_hist: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference's private `__update` (trainer.py:791-816): the gradient of the minus-domain network
        (`mlp_m_fn` parameters) is multiplied by `mgrad_over_pgrad_scalefactor`, the plus-domain gradient is taken as
        is, one optimizer update is applied to both, and the returned loss is m_loss + p_loss (the same loss evaluated
        twice, i.e. 2 x loss).  One value_and_grad here: both of the reference's passes differentiate the same
        function.  Returns the device scalar; no training loop calls this (none does in the reference either)."""
        L = cabi.lib()
        with torch.cuda.device(self.device):
            upload_params(self.net, self.params)
            lg = plan.loss_grad_launch()
            n_p = self.net.n_p      # flat layout: [p-head | m-head | preconditioner]
            lg[n_p:self.net.n_params] *= float(self.mgrad_over_pgrad_scalefactor)
            lg[-1] *= 2.0
            cabi.check(L.nbm_apply_update_f32(C.byref(self._optimizer_struct()), cabi.ptr(lg), cabi.ptr(self.params),
                                              cabi.ptr(self.opt_state), cabi.ptr(self.opt_count), cabi.ptr(loss_hist),
                                              cabi.stream_ptr()), "nbm_apply_update_f32")
            return lg[-1]

    def _begin_training(self) -> None:
        """stage the current parameters (plain / pre-scaled / transposed copies) for the first step of a loop"""
        with torch.cuda.device(self.device):
            upload_params(self.net, self.params)

    def _peer_comm(self):
        if self._comm is None:
            from .comm import PeerComm
            d = _dist()
            try:
                comm, ok = PeerComm(self.device), 1
            except Exception as exc:  # noqa: BLE001
                logger.warning("peer all-reduce unavailable (%s); using NCCL", exc)
                comm, ok = None, 0
            flag = torch.tensor([ok], device=self.device)
            d.all_reduce(flag, op=d.ReduceOp.MIN)        # every rank must take the same path
            if int(flag.item()) == 0:
                self.allreduce_kind = "nccl"
                return None
            self._comm = comm
        return self._comm

    def _graph_step(self, plan, loss_hist, allreduce: bool):
        """replay the step as a CUDA graph (launch-bound at small grids).  The multi-GPU step is captured too when its
        exchange is the peer-memory kernel (an ordinary kernel launch); with NCCL the collective stays eager."""
        peer = False
        if allreduce:
            d = _dist()
            if d is not None and d.get_world_size() > 1:
                peer = (self.allreduce_kind == "peer" and isinstance(plan, (SharedPlan, EmptyPlan))
                        and self._peer_comm() is not None)
                if not peer:
                    return self._step(plan, loss_hist, allreduce)
            else:
                allreduce = False
        if not self.use_cuda_graph:
            return self._step(plan, loss_hist, allreduce)
        key = (id(plan), loss_hist.data_ptr() if loss_hist is not None else 0, bool(allreduce))
        g = self._graphs.get(key)
        if g is None:
            # (a warm-up outside capture would change the state: capture directly; the capture does not execute the step)
            g = capture_graph(lambda: self._step(plan, loss_hist, allreduce), self.device)
            self._graphs[key] = g
        g.replay()

    def _check_comm(self, where: str) -> None:
        """a peer wait that timed out poisons the result with NaN on that rank only: turn it into an exception on the
        host instead of letting the replicas diverge silently"""
        if self._comm is not None and self._comm.error():
            raise cabi.NbmError(f"peer all-reduce timed out ({where}): a rank did not reach the exchange within "
                                f"{self._comm.timeout_s:g} s (NBM_PEER_TIMEOUT_S); parameters on this rank are invalid")

    # ------------------------------------------------------------------------------------------
    # training loops
    # ------------------------------------------------------------------------------------------
    def solve_optax(self):
        start_time = time.time()
        if self.multi_gpu:
            self.epoch_store, self.loss_epochs = self.multi_GPU_train()
        else:
            self.epoch_store, self.loss_epochs = self.single_GPU_train()
        torch.cuda.synchronize(self.device)
        logger.info(f"solve took {time.time() - start_time} (sec)")
        d = _dist()
        last_epoch = int(self.epoch_store[-1]) + 1 if len(self.epoch_store) else self.epoch_start
        if d is None or d.get_rank() == 0:
            self.save_checkpoint(self.checkpoint_dir, self._checkpoint_state(last_epoch))
        final_solution, grad_u, grad_u_normal = self.evaluate_solution_and_gradients(self.params, self.eval_gstate)
        return final_solution, grad_u, grad_u_normal, self.epoch_store, self.loss_epochs

    def solve_jaxopt(self):
        """trainer.py:354-427: `jaxopt.ScipyMinimize(method="l-bfgs-b", fun=self.loss, tol=1e-15,
        maxiter=num_epochs)` on the FIRST batch at the native cell size.  jaxopt hands scipy a float64 copy of the
        flattened parameters and a value_and_grad callback; here the callback is the CUDA loss/gradient step
        (one launch sequence + a 168-float read-back per evaluation), scipy's L-BFGS-B does the rest on the host."""
        from scipy.optimize import minimize
        DD = data_management.DatasetDict(num_points=self.tr_gstate.num_points(), batch_size=self.batch_size)
        p0, p1 = DD.ranges(0)[0]
        plan = self.plan_for(0, p0, p1)
        n = self.n_params
        history = []

        def fun(x):
            with torch.cuda.device(self.device):
                self.params.copy_(torch.from_numpy(np.asarray(x, dtype=np.float32)))
                lg = self.loss_and_grad(self.params, plan).cpu().double().numpy()
            history.append(float(lg[n]))
            return float(lg[n]), lg[:n]

        start_time = time.time()
        with torch.cuda.device(self.device):
            x0 = self.params.detach().cpu().double().numpy()
        sol = minimize(fun, x0, jac=True, method="L-BFGS-B", tol=1e-15, options={"maxiter": int(self.num_epochs)})
        logger.info(f"solve took {time.time() - start_time} (sec)")
        with torch.cuda.device(self.device):
            self.params.copy_(torch.from_numpy(sol.x.astype(np.float32)))
        self.scipy_result = sol
        self.loss_history = history
        d = _dist()
        if d is None or d.get_rank() == 0:
            self.save_checkpoint(self.checkpoint_dir, {
                "opt_state": {"fun_val": float(sol.fun), "iter_num": int(sol.nit), "success": bool(sol.success),
                              "status": int(sol.status)},
                "params": params_to_tree(self.net, self.params, self.precond), "epoch": int(self.epoch_store[-1]) + 1,
                "batch_size": self.batch_size, "resolution": f"{self.train_dx}, {self.train_dy}, {self.train_dz}"})
        final_solution, grad_u, grad_u_normal = self.evaluate_solution_and_gradients(self.params, self.eval_gstate)
        # like the reference, epoch_store / loss_epochs are returned as initialised (trainer.py:260-263, 421-427)
        return final_solution, grad_u, grad_u_normal, self.epoch_store, self.loss_epochs

    def single_GPU_train(self):
        """trainer.py:501-591: epochs x batches, cell size halves every num_epochs//4 epochs
        (data_management.py:320-326), loss_epochs[e] = mean over batches of the batch losses."""
        DD = data_management.DatasetDict(num_points=self.tr_gstate.num_points(), batch_size=self.batch_size)
        self._warn_if_padded(DD)
        ranges = DD.ranges(0)
        nb = len(ranges)
        self._begin_training()
        with torch.cuda.device(self.device):
            loss_hist = torch.zeros(self.num_epochs * nb + 1, dtype=torch.float32, device=self.device)
            base = int(self.opt_count.item())
            hist = loss_hist if base == 0 else None
            per_step = [] if hist is None else None
            for epoch in range(self.num_epochs):
                zoom = self.TD.zoom_level(self.num_epochs, epoch)
                for (p0, p1) in ranges:
                    plan = self.plan_for(zoom, p0, p1)
                    self._graph_step(plan, hist, allreduce=False)
                    if per_step is not None:
                        per_step.append(plan.loss_grad[-1:].clone())
                if self.print_rate and epoch % self.print_rate == 0 and logger.isEnabledFor(logging.INFO):
                    logger.info(f"epoch # {epoch}")
            torch.cuda.synchronize(self.device)
            if hist is not None:
                losses = hist[: self.num_epochs * nb].view(self.num_epochs, nb).mean(dim=1).cpu()
            else:
                losses = torch.cat(per_step).view(self.num_epochs, nb).mean(dim=1).cpu()
        return np.arange(self.epoch_start, self.num_epochs), losses

    def multi_GPU_train(self):
        """trainer.py:715-779: the points are split in contiguous blocks (x-slabs, data_management.py:121-130), no
        multi-resolution schedule, grads and loss are SUMMED over devices (:829-830) and every device applies the
        identical update.  Two launch models: one process per GPU under `torchrun` (torch.distributed initialised), or -
        like the reference's single-process `pmap` over `jax.local_device_count()` (:727-743) - this process driving every
        local device itself (`_multi_device_train`)."""
        global stop_training
        d = _dist()
        if d is None:
            n_local = min(torch.cuda.device_count(), self.n_devices or torch.cuda.device_count())
            if n_local > 1:
                return self._multi_device_train(n_local)
            logger.warning("multi_gpu=True with one visible device and no torch.distributed group: training on one GPU "
                           "(jax.local_device_count() == 1 gives the reference the same)")
        world = d.get_world_size() if d is not None else 1
        rank = d.get_rank() if d is not None else 0
        DD = data_management.DatasetDict(num_points=self.tr_gstate.num_points(), batch_size=world * self.batch_size,
                                         num_gpus=world)
        self._warn_if_padded(DD)
        all_ranges = self._device_ranges(DD, world)
        ranges = all_ranges[rank]
        nb = len(ranges)
        # every rank must pick the same exchange for the same batch: the fused peer kernel serves whole-plane batches
        # (shared path) only
        plane = self.tr_gstate.shape()[1] * self.tr_gstate.shape()[2]
        if not all(p0 % plane == 0 and p1 % plane == 0 for rr in all_ranges for (p0, p1, _) in rr if p1 > p0):
            self.allreduce_kind = "nccl"
        loss_epochs, epoch_store = [], []
        t0 = time.time()
        self._begin_training()
        with torch.cuda.device(self.device):
            # every plan of an epoch is built BEFORE the first exchange (set-up cost differs a lot between ranks: the
            # interface may lie in a few slabs only), then the ranks meet at a barrier
            plans = [self.plan_for(0, p0, p1, n_mean) for (p0, p1, n_mean) in ranges]
            if d is not None and world > 1:
                if self.allreduce_kind == "peer":
                    self._peer_comm()
                torch.cuda.synchronize(self.device)
                d.barrier()
            n_steps = (self.num_epochs - self.epoch_start) * nb
            loss_hist = torch.zeros(n_steps + 1, dtype=torch.float32, device=self.device)
            base = int(self.opt_count.item())
            # the finalize kernel files the loss under the optimizer's step count: a restarted run (count != 0) gets
            # a view that starts `base` entries earlier so that its steps land in [0, n_steps)
            hist = loss_hist if base == 0 else None
            per_step = [] if hist is None else None
            stop_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            for epoch in range(self.epoch_start, self.num_epochs):
                # SIGINT reaches the ranks at different times (or only one of them): agree on the epoch to stop at
                stop = bool(stop_training)
                if d is not None and world > 1:
                    stop_flag.fill_(1 if stop else 0)
                    d.all_reduce(stop_flag, op=d.ReduceOp.MAX)
                    stop = bool(int(stop_flag.item()))
                if stop:
                    break
                for plan in plans:
                    self._graph_step(plan, hist, allreduce=True)
                    if per_step is not None:
                        per_step.append(plan.loss_grad[-1:].clone())
                epoch_store.append(epoch)
                if self.print_rate and epoch % self.print_rate == 0 and logger.isEnabledFor(logging.INFO):
                    dt_avg = (time.time() - t0) / self.print_rate
                    t0 = time.time()
                    logger.info(f"Epoch # {epoch} \t avg epoch time is {dt_avg} (sec)")
                if (epoch + 1) % self.checkpoint_interval == 0:
                    self._check_comm(f"epoch {epoch}")       # (host read: synchronises this rank's stream)
                    if rank == 0:
                        self.save_checkpoint(self.checkpoint_dir, self._checkpoint_state(epoch + 1))
            torch.cuda.synchronize(self.device)
            self._check_comm("end of training")
            n_done = len(epoch_store)
            if hist is not None:
                per_epoch = loss_hist[: n_done * nb].view(n_done, nb).mean(dim=1).cpu()
            else:
                per_epoch = (torch.cat(per_step).view(n_done, nb).mean(dim=1).cpu() if n_done
                             else torch.zeros(0))
        # the reference returns a list of per-device arrays (all entries equal after the psum)
        loss_epochs = [per_epoch[e].repeat(world) for e in range(n_done)]
        return epoch_store, loss_epochs

    def _multi_device_train(self, n_dev: int):
        """`multi_gpu=True` without a torch.distributed group: ONE process drives all local devices, as the reference's
        `pmap(update_multi_gpu, axis_name="devices")` does (trainer.py:727-756).  Every device holds a replica
        (parameters, optimizer state, level set, its x-slab's row tables); per step every device runs its slab's
        kernels on its own stream and the [grad, loss] SUM (psum, :829-830) is the peer-memory exchange kernel between
        the devices' blocks (peer access inside the process instead of CUDA IPC); each replica's step is one CUDA graph."""
        global stop_training
        from .comm import LocalPeerComm
        devices = [torch.device("cuda", r) for r in range(n_dev)]
        DD = data_management.DatasetDict(num_points=self.tr_gstate.num_points(), batch_size=n_dev * self.batch_size,
                                         num_gpus=n_dev)
        self._warn_if_padded(DD)
        plane = self.tr_gstate.shape()[1] * self.tr_gstate.shape()[2]
        if not all(p0 % plane == 0 and p1 % plane == 0 for r in range(n_dev) for (p0, p1) in DD.ranges(r) if p1 > p0):
            raise NotImplementedError(
                "single-process multi-device training needs batches of whole x planes on every device (the fused peer "
                "exchange serves the shared-evaluation path); launch one process per GPU with torchrun for ragged "
                "partitions (NCCL exchange)")
        reps = [self]
        for dev in devices:
            if dev == self.device:
                continue
            reps.append(Trainer(self.lvl_gstate, self.tr_gstate, self.eval_gstate, self.sim_state, self.sim_state_fn,
                                0, lvl_set_fn=None, num_epochs=self.num_epochs, multi_gpu=False,
                                batch_size=self.batch_size, checkpoint_dir=None, optimizer_dict=self._optimizer_dict,
                                model_dict=self._model_dict, phi_interp=self._phi_interp, perturb_eps=self._perturb_eps,
                                device=dev, init_params=self.params.detach().cpu(), use_cuda_graph=self.use_cuda_graph,
                                print_rate=0))
        reps.sort(key=lambda t: t.device.index)
        for t in reps:
            if t is not self:          # replicate the optimizer state too (a restarted run)
                t.opt_state.copy_(self.opt_state.to(t.device))
                t.opt_count.copy_(self.opt_count.to(t.device))
        comm = LocalPeerComm([t.device for t in reps])
        dev_ranges = self._device_ranges(DD, n_dev)
        nb = len(dev_ranges[0])
        n_steps = (self.num_epochs - self.epoch_start) * nb
        base = int(self.opt_count.item())
        hists, plans, graphs = [], [], []
        for r, t in enumerate(reps):
            with torch.cuda.device(t.device):
                t._begin_training()
                hists.append(torch.zeros(n_steps + 1, dtype=torch.float32, device=t.device))
                plans.append([t.plan_for(0, p0, p1, n_mean) for (p0, p1, n_mean) in dev_ranges[r]])
                torch.cuda.synchronize(t.device)
        per_step = [] if base != 0 else None

        def launch(r, b):
            t = reps[r]
            with torch.cuda.device(t.device):
                hist = hists[r] if base == 0 else None
                if not self.use_cuda_graph:
                    return t._step(plans[r][b], hist, True, local_comm=comm.handle(r))
                key = (r, b)
                g = t._graphs.get(key)
                if g is None:
                    g = capture_graph(lambda: t._step(plans[r][b], hist, True, local_comm=comm.handle(r)), t.device)
                    t._graphs[key] = g
                g.replay()

        epoch_store = []
        t0 = time.time()
        try:
            for epoch in range(self.epoch_start, self.num_epochs):
                if stop_training:
                    break
                for b in range(nb):
                    for r in range(n_dev):
                        launch(r, b)
                    if per_step is not None:
                        with torch.cuda.device(self.device):
                            per_step.append(plans[reps.index(self)][b].loss_grad[-1:].clone())
                epoch_store.append(epoch)
                if self.print_rate and epoch % self.print_rate == 0 and logger.isEnabledFor(logging.INFO):
                    dt_avg = (time.time() - t0) / self.print_rate
                    t0 = time.time()
                    logger.info(f"Epoch # {epoch} \t avg epoch time is {dt_avg} (sec)")
                if (epoch + 1) % self.checkpoint_interval == 0:
                    comm.raise_on_error(f"epoch {epoch}")
                    self.save_checkpoint(self.checkpoint_dir, self._checkpoint_state(epoch + 1))
            for t in reps:
                torch.cuda.synchronize(t.device)
            comm.raise_on_error("end of training")
        finally:
            for t in reps:
                t._graphs.clear()
            comm.close()
        n_done = len(epoch_store)
        with torch.cuda.device(self.device):
            if per_step is None:
                per_epoch = hists[reps.index(self)][: n_done * nb].view(n_done, nb).mean(dim=1).cpu()
            else:
                per_epoch = torch.cat(per_step).view(n_done, nb).mean(dim=1).cpu() if n_done else torch.zeros(0)
        self._replicas = reps       # (tests compare the replicas' parameters)
        return epoch_store, [per_epoch[e].repeat(n_dev) for e in range(n_done)]

    # ------------------------------------------------------------------------------------------
    # post-training evaluation (trainer.py:960-977)
    # ------------------------------------------------------------------------------------------
    def evaluate_solution_and_gradients(self, params: torch.Tensor, eval_gstate, chunk: int = 1 << 24):
        L = cabi.lib()
        n = eval_gstate.num_points()
        net = self.net.struct()
        with torch.cuda.device(self.device):
            upload_params(self.net, params.to(self.device))
            u = torch.empty(n, dtype=torch.float32, device=self.device)
            gu = torch.empty(n * 3, dtype=torch.float32, device=self.device)
            gn = torch.empty(n, dtype=torch.float32, device=self.device)
            R = eval_gstate.R
            dx, dy, dz = float(eval_gstate.dx), float(eval_gstate.dy), float(eval_gstate.dz)
            for s in range(0, n, chunk):
                e = min(n, s + chunk)
                pts = R[s:e].to(self.device).contiguous()
                lvl_struct = self._eval_lvl(pts, (dx, dy, dz), normals=True)
                cabi.check(L.nbm_evaluate_f32(C.byref(net), C.byref(lvl_struct), cabi.ptr(pts), e - s, dx, dy, dz,
                                              u[s:].data_ptr(), gu[3 * s:].data_ptr(), gn[s:].data_ptr(),
                                              cabi.stream_ptr()), "nbm_evaluate_f32")
                torch.cuda.current_stream().synchronize()
        return u, gu.view(n, 3), gn

    def _eval_lvl(self, pts: torch.Tensor, d, normals: bool):
        """the level-set descriptor for nbm_evaluate_f32: the grid form, or - analytic level set - its samples at the
        points and at +- d along the axes (discretization.py:199-218), positions formed in fp32 like the kernel's"""
        if not self.lvl.analytic:
            return self.lvl.struct
        n = pts.shape[0]
        ep = torch.zeros((n, 7), dtype=torch.float32, device=self.device)
        ep[:, 0] = self.lvl(pts)
        if normals:
            for a in range(3):
                h = torch.tensor(d[a], dtype=torch.float32, device=self.device)
                for k, sg in ((1 + 2 * a, -1.0), (2 + 2 * a, 1.0)):
                    q = pts.clone()
                    q[:, a] = pts[:, a] + sg * h
                    ep[:, k] = self.lvl(q)
        self._eval_phi = ep.contiguous()
        return self.lvl.with_samples(eval_phi=self._eval_phi)

    def evaluate_solution_fn(self, params: torch.Tensor, R_flat: torch.Tensor) -> torch.Tensor:
        """trainer.py:836-844"""
        L = cabi.lib()
        net = self.net.struct()
        with torch.cuda.device(self.device):
            upload_params(self.net, params.to(self.device))
            pts = R_flat.to(self.device, torch.float32).contiguous()
            u = torch.empty(pts.shape[0], dtype=torch.float32, device=self.device)
            lvl_struct = self._eval_lvl(pts, (1.0, 1.0, 1.0), normals=False)
            cabi.check(L.nbm_evaluate_f32(C.byref(net), C.byref(lvl_struct), cabi.ptr(pts), pts.shape[0], 1.0, 1.0,
                                          1.0, cabi.ptr(u), None, None, cabi.stream_ptr()), "nbm_evaluate_f32")
        return u


def setup(initial_value_fn, dirichlet_bc_fn, lvl_set_fn, mu_m_fn_, mu_p_fn_, k_m_fn_, k_p_fn_, f_m_fn_, f_p_fn_,
          alpha_fn_, beta_fn_, nonlinear_op_m=None, nonlinear_op_p=None):
    """trainer.py:980-1137.  Returns `init_fn`."""
    v = jnp.vmap
    u_0_fn, dir_bc_fn, phi_fn = v(initial_value_fn), v(dirichlet_bc_fn), v(lvl_set_fn)
    mu_m_fn, mu_p_fn, k_m_fn, k_p_fn = v(mu_m_fn_), v(mu_p_fn_), v(k_m_fn_), v(k_p_fn_)
    f_m_fn, f_p_fn, alpha_fn, beta_fn = v(f_m_fn_), v(f_p_fn_), v(alpha_fn_), v(beta_fn_)
    if nonlinear_op_m is None:
        logger.warning("nonlinear_op_m(u) is not defined. Setting it to zero.")
    if nonlinear_op_p is None:
        logger.warning("nonlinear_op_m(u) is not defined. Setting it to zero.")
    nonlinear_op_m = Nonlinear.coerce(nonlinear_op_m)
    nonlinear_op_p = Nonlinear.coerce(nonlinear_op_p)
    sim_state_fn = PoissonSimStateFn(u_0_fn, dir_bc_fn, phi_fn, mu_m_fn, mu_p_fn, k_m_fn, k_p_fn, f_m_fn, f_p_fn,
                                     alpha_fn, beta_fn, nonlinear_op_m, nonlinear_op_p)

    def init_fn(lvl_gstate=None, tr_gstate=None, eval_gstate=None, num_epochs: int = 1000, batch_size: int = 131072,
                algorithm: int = 0, mgrad_over_pgrad_scalefactor: int = 1, multi_gpu: bool = False,
                checkpoint_interval: int = 1000, checkpoint_dir: str = "./checkpoints", results_dir: str = "./",
                loss_plot_name: str = "solver_loss", optimizer_dict: dict = None, model_dict: dict = None,
                restart: bool = False, restart_checkpoint_dir: str = "./checkpoints", print_rate: int = 1,
                **extensions) -> Tuple[PoissonSimState, Callable]:
        """`extensions` (not in the reference): phi_interp, perturb_eps, device, init_params, use_cuda_graph, allreduce,
        n_devices."""
        device = torch.device(extensions["device"]) if extensions.get("device") is not None else default_device()
        optimizer_dict = optimizer_dict or {"optimizer_name": "custom", "learning_rate": 1e-3,
                                            "sched": {"scheduler_name": "exponential", "decay_rate": 0.9}}
        model_dict = model_dict or _DEFAULT_MODEL
        with torch.cuda.device(device):
            R = eval_gstate.R.to(device)
            PHI, DIRBC, U = phi_fn(R), dir_bc_fn(R), u_0_fn(R)
            MU_M, MU_P, K_M, K_P = mu_m_fn(R), mu_p_fn(R), k_m_fn(R), k_p_fn(R)
            F_M, F_P, ALPHA, BETA = f_m_fn(R), f_p_fn(R), alpha_fn(R), beta_fn(R)
            del R

        def solve_fn(sim_state: PoissonSimState):
            trainer = Trainer(lvl_gstate, tr_gstate, eval_gstate, sim_state, sim_state_fn, algorithm,
                              mgrad_over_pgrad_scalefactor=mgrad_over_pgrad_scalefactor, lvl_set_fn=lvl_set_fn,
                              num_epochs=num_epochs, multi_gpu=multi_gpu, batch_size=batch_size,
                              checkpoint_dir=checkpoint_dir, checkpoint_interval=checkpoint_interval,
                              results_dir=results_dir, loss_plot_name=loss_plot_name, optimizer_dict=optimizer_dict,
                              restart=restart, restart_checkpoint_dir=restart_checkpoint_dir, print_rate=print_rate,
                              model_dict=model_dict, **extensions)
            solve_fn.trainer = trainer
            final_solution, grad_u, grad_u_normal_to_interface, epoch_store, loss_epochs = trainer.solve()
            return (replace(sim_state, solution=final_solution, grad_solution=grad_u,
                            grad_normal_solution=grad_u_normal_to_interface), epoch_store, loss_epochs)

        return (PoissonSimState(PHI, U, DIRBC, MU_M, MU_P, K_M, K_P, F_M, F_P, ALPHA, BETA, None, None), solve_fn)

    return init_fn
