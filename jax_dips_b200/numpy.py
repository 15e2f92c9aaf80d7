"""A small `jax.numpy` look-alike on torch, so that coefficient callables written for the
reference (`from jax import numpy as jnp`; `r[0]`, `jnp.exp(z)`, ...) run unchanged here:
replace the import with `from jax_dips_b200 import numpy as jnp`.

The reference `vmap`s per-point callables `f(r: (3,)) -> ()` (trainer.py:995-1005).  Here the
same callable is invoked once with `r` = the (3, n) coordinate-major view of the batch, so
`r[0]`, `r[1]`, `r[2]` are length-n vectors and every element-wise expression broadcasts.
"""
from __future__ import annotations

import math

import torch

pi = math.pi
e = math.e
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
newaxis = None


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float32)


def _wrap(fn):
    def f(x, *a, **k):
        return fn(_t(x), *a, **k)
    f.__name__ = fn.__name__
    return f


sin, cos, tan = _wrap(torch.sin), _wrap(torch.cos), _wrap(torch.tan)
sinh, cosh, tanh = _wrap(torch.sinh), _wrap(torch.cosh), _wrap(torch.tanh)
arcsin, arccos, arctan = _wrap(torch.asin), _wrap(torch.acos), _wrap(torch.atan)
exp, log, sqrt, abs, sign = _wrap(torch.exp), _wrap(torch.log), _wrap(torch.sqrt), _wrap(torch.abs), _wrap(torch.sign)
floor, ceil, square = _wrap(torch.floor), _wrap(torch.ceil), _wrap(torch.square)
log10, log2, exp2 = _wrap(torch.log10), _wrap(torch.log2), _wrap(torch.exp2)
absolute = abs


def _pair(fn):
    def f(a, b):
        a, b = _t(a), _t(b)
        if a.device != b.device:
            b = b.to(a.device) if a.dim() >= b.dim() else b
            a = a.to(b.device)
        return fn(a, b)
    return f


arctan2 = _pair(torch.atan2)
minimum = _pair(torch.minimum)
maximum = _pair(torch.maximum)
power = _pair(torch.pow)


def where(c, a, b):
    c = _t(c)
    a = _t(a).to(c.device)
    b = _t(b).to(c.device)
    return torch.where(c.bool(), a, b)


def array(x, dtype=None):
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        ref = next(v for v in x if isinstance(v, torch.Tensor))
        return torch.stack([_t(v).to(ref.device).expand_as(ref) if _t(v).dim() == 0 else _t(v) for v in x])
    return torch.as_tensor(x, dtype=dtype or torch.float32)


asarray = array


def zeros_like(x): return torch.zeros_like(_t(x))
def ones_like(x): return torch.ones_like(_t(x))
def nan_to_num(x): return torch.nan_to_num(_t(x))
def clip(x, lo, hi): return torch.clamp(_t(x), lo, hi)
def dot(a, b): return (a * b).sum(dim=0)
def sum(x, axis=None): return _t(x).sum() if axis is None else _t(x).sum(dim=axis)
def min(x, axis=None): return _t(x).min() if axis is None else _t(x).min(dim=axis).values
def max(x, axis=None): return _t(x).max() if axis is None else _t(x).max(dim=axis).values
def linspace(a, b, n, dtype=None): return torch.linspace(a, b, n, dtype=torch.float64).to(dtype or torch.float32)


def vmap(fn):
    """Batched form of a per-point callable: (n,3) -> (n,).  See the module docstring."""
    def batched(R: torch.Tensor) -> torch.Tensor:
        n = R.shape[0]
        out = fn(R.t())
        if not isinstance(out, torch.Tensor):
            out = torch.as_tensor(out, dtype=R.dtype, device=R.device)
        out = out.to(device=R.device, dtype=R.dtype)
        if out.dim() == 0 or out.shape[-1] != n:
            out = out.reshape(-1)[:1].expand(n) if out.numel() == 1 else out
        return out.reshape(n).contiguous()
    batched.__wrapped__ = fn
    return batched
