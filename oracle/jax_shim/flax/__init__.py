"""Stand-in for the sliver of flax that the reference's learned preconditioner touches
(jax_dips/nn/preconditioner.py:10-35 and trainer.py:229-243, 846-847): `flax.linen.Module` with a
`@compact` `__call__`, `nn.Dense`, `nn.sigmoid`, `nn.initializers.glorot_uniform`, `flax.core.freeze/unfreeze`.
Oracle infrastructure only (see oracle/jax_shim/jax/__init__.py); parameters are always INPUTS here
(`Module.init` raises: flax's PRNG-driven initialisation cannot be reproduced without jax)."""
from . import linen  # noqa: F401
from . import core  # noqa: F401
