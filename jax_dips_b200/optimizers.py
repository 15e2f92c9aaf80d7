"""Optimizer factory mirroring jax_dips/solvers/optimizers.py (same names, arguments, errors).

The reference returns optax GradientTransformations; here `get_optimizer` returns an
`OptimizerSpec` that the CUDA update kernel (`nbm_apply_update_f32`) executes on the device:

* "custom"  : clip_by_global_norm(max_norm) -> scale_by_adam -> scale_by_schedule -> scale(-1)
              (optimizers.py:33-54)
* "adam"    : optax.adam(learning_rate)       (optimizers.py:76-81)
* "rmsprop" : optax.rmsprop(learning_rate)    (optimizers.py:83-88)
"""
from __future__ import annotations

import dataclasses
import logging

logger = logging.getLogger(__name__)


@dataclasses.dataclass
class Scheduler:
    scheduler_name: str = "exponential"
    learning_rate: float = 1e-2
    decay_rate: float = 0.96
    transition_steps: int = 1000

    def __call__(self, count: int) -> float:
        if self.scheduler_name == "exponential":
            return self.learning_rate * self.decay_rate ** (count / self.transition_steps)
        # optax.polynomial_schedule(init, end=0, power=1, transition_steps)
        frac = 1.0 - min(count, self.transition_steps) / self.transition_steps
        return self.learning_rate * frac


def get_scheduler(scheduler_name: str = "exponential", learning_rate: float = 1e-2, decay_rate: float = 0.96,
                  transition_steps: int = 1000, **kwargs) -> Scheduler:
    """optimizers.py:13-30"""
    if scheduler_name == "exponential":
        logger.info("Using Exponential Scheduler")
    elif scheduler_name == "polynomial":
        logger.info("Using Polynomial Scheduler")
    else:
        raise ValueError("Unknown scheduler: {}".format(scheduler_name))
    return Scheduler(scheduler_name, learning_rate, decay_rate, transition_steps)


@dataclasses.dataclass
class OptimizerSpec:
    optimizer_name: str = "custom"
    scheduler: Scheduler = dataclasses.field(default_factory=Scheduler)
    learning_rate: float = 1e-2
    max_norm: float = 1.0
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8

    @property
    def kind(self) -> int:
        return {"custom": 0, "adam": 1, "rmsprop": 2}[self.optimizer_name]


def chained_adam(scheduler_name: str = "exponential", learning_rate: float = 1e-2, decay_rate: float = 0.96,
                 transition_steps: int = 1000, max_norm: float = 1.0, **kwargs) -> OptimizerSpec:
    """optimizers.py:33-54"""
    sched = get_scheduler(scheduler_name, learning_rate, decay_rate, transition_steps)
    return OptimizerSpec("custom", sched, learning_rate, max_norm)


def get_optimizer(optimizer_name: str = "custom", scheduler_name: str = "exponential", learning_rate: float = 1e-2,
                  decay_rate: float = 0.96, max_norm: float = 1.0, loss_fn: object = None, **kwargs) -> OptimizerSpec:
    """optimizers.py:57-97"""
    if optimizer_name == "custom":
        logger.info("Using chained Adam optimizer")
        return chained_adam(scheduler_name=scheduler_name, learning_rate=learning_rate, decay_rate=decay_rate,
                            max_norm=max_norm, **kwargs)
    elif optimizer_name == "adam":
        logger.info("Using Adam optimizer")
        return OptimizerSpec("adam", Scheduler("exponential", learning_rate, 1.0, 1000), learning_rate, max_norm)
    elif optimizer_name == "rmsprop":
        logger.info("Using RMSprop optimizer")
        return OptimizerSpec("rmsprop", Scheduler("exponential", learning_rate, 1.0, 1000), learning_rate, max_norm)
    else:
        logger.error("Unknown optimizer: {}".format(optimizer_name))
        raise ValueError("Unknown optimizer: {}".format(optimizer_name))
