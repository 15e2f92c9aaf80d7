import functools

import numpy as _np

from .. import numpy as jnp


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _stack(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([o[i] for o in outs]) for i in range(len(first)))
    return jnp.Arr(_np.stack([_np.asarray(o) for o in outs], axis=0))


def vmap(fn, in_axes=0, out_axes=0, axis_name=None):
    """A Python loop over the mapped axis (axis 0 or None per positional argument)."""
    @functools.wraps(fn)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = len(a)
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else a[i] for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return _stack(outs)
    return mapped


def grad(fn, *a, **k):
    def g(*args, **kwargs):
        raise NotImplementedError("the numpy stand-in has no autodiff")
    return g


value_and_grad = grad


def pmap(fn, **kwargs):
    return fn
