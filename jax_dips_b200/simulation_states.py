"""Containers mirroring jax_dips/solvers/simulation_states.py:11-25, 58-80."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional


@dataclasses.dataclass
class PoissonSimStateFn:
    """Batched coefficient callables (n,3) -> (n,) (simulation_states.py:11-25)."""
    u_0_fn: Callable
    dir_bc_fn: Callable
    phi_fn: Callable
    mu_m_fn: Callable
    mu_p_fn: Callable
    k_m_fn: Callable
    k_p_fn: Callable
    f_m_fn: Callable
    f_p_fn: Callable
    alpha_fn: Callable
    beta_fn: Callable
    nonlinear_op_m: Any
    nonlinear_op_p: Any


@dataclasses.dataclass
class PoissonSimState:
    """Eval-grid fields + the trained solution (simulation_states.py:58-80)."""
    phi: Any
    solution: Any
    dirichlet_bc: Any
    mu_m: Any
    mu_p: Any
    k_m: Any
    k_p: Any
    f_m: Any
    f_p: Any
    alpha: Any
    beta: Any
    grad_solution: Optional[Any]
    grad_normal_solution: Optional[Any]


def replace(state, **kw):
    return dataclasses.replace(state, **kw)
