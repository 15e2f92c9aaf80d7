"""A minimal numpy stand-in for the parts of the jax API that the reference's NBM path touches
(oracle infrastructure; runs ONLY in the build container, to execute the reference's own source files
under /root/reference and pin the oracle against them; see oracle/make_golden.py).

Semantics reproduced: float32 default dtype unless `config.update("jax_enable_x64", True)`, weak
python scalars, `.at[idx].set()`, stable `argsort`, `jnp.linalg.pinv`'s default cutoff
(10*max(M,N)*eps), `nan_to_num`, `isclose`, `lax.cond`, `vmap` (a Python loop), `jit` (identity).
out-of-range integer gathers clamp.  Not reproduced: tracing, autodiff (`grad` raises), XLA's
fusion/FMA choices.
"""
from . import numpy  # noqa: F401
from . import lax  # noqa: F401
from . import tree_util  # noqa: F401
from . import lib  # noqa: F401
from . import nn  # noqa: F401
from ._src.api import vmap, jit, grad, value_and_grad, pmap  # noqa: F401
from .numpy import config  # noqa: F401

tree_map = tree_util.tree_map
Array = numpy.ndarray     # annotation only (nn/preconditioner.py:23)


def local_device_count():
    return 1
