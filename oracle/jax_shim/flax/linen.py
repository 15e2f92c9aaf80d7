"""flax.linen stand-in.  Semantics restated from flax 0.7 (the reference pins no version; `nn.Dense` has been
`y = lax.dot_general(x, kernel) + bias` with kernel of shape (in, out) in every release):
  * submodules created inside a `@compact` `__call__` are auto-named `<Class>_<k>` in call order;
  * `module.apply({"params": tree}, x)` binds that tree and calls `__call__`;
  * `nn.sigmoid` = jax.nn.sigmoid = 1 / (1 + exp(-x))."""
import numpy as _np
from jax import numpy as jnp

_scopes = []     # stack of {"params": subtree, "count": {class name: next index}}


class _Initializers:
    @staticmethod
    def glorot_uniform():
        def init(key, shape, dtype=None):
            raise NotImplementedError("the stand-in takes parameters as inputs")
        return init


initializers = _Initializers()


def compact(fn):
    return fn


def sigmoid(x):
    return 1.0 / (1.0 + jnp.exp(-x))


class Module:
    """dataclass-style module: annotated class attributes are constructor keywords"""

    def __init__(self, **kwargs):
        fields = {}
        for klass in reversed(type(self).__mro__):
            fields.update(getattr(klass, "__annotations__", {}))
        for name in fields:
            if name in kwargs:
                setattr(self, name, kwargs.pop(name))
            elif not hasattr(type(self), name):
                raise TypeError(f"{type(self).__name__}: missing field {name!r}")
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unexpected fields {sorted(kwargs)}")

    def _child_params(self):
        scope = _scopes[-1]
        k = scope["count"].get(type(self).__name__, 0)
        scope["count"][type(self).__name__] = k + 1
        return scope["params"][f"{type(self).__name__}_{k}"]

    def apply(self, variables, *args, **kwargs):
        _scopes.append({"params": variables["params"], "count": {}})
        try:
            return self(*args, **kwargs)
        finally:
            _scopes.pop()

    def init(self, *args, **kwargs):
        raise NotImplementedError("the stand-in takes parameters as inputs")


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = None

    def __init__(self, features=None, **kwargs):
        super().__init__(features=features, **kwargs)

    def __call__(self, x):
        p = self._child_params()
        kernel = jnp.asarray(p["kernel"])
        assert kernel.shape[-1] == self.features
        y = jnp.dot(x, kernel)
        if self.use_bias:
            y = y + jnp.asarray(p["bias"])
        return y
