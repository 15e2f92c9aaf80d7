import builtins

import numpy as _np

newaxis = None
pi = _np.pi
float32, float64, int32, int64, int16 = _np.float32, _np.float64, _np.int32, _np.int64, _np.int16
ndarray = _np.ndarray
index_exp = _np.index_exp


class _Config:
    x64 = False

    def update(self, key, value):
        if key == "jax_enable_x64":
            _Config.x64 = bool(value)


config = _Config()


def _default_float():
    return _np.float64 if _Config.x64 else _np.float32


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, value):
                out = _np.array(arr, copy=True)
                out[idx] = value
                return Arr(out)

            def add(self, value):
                out = _np.array(arr, copy=True)
                out[idx] += value
                return Arr(out)
        return _Setter()


class Arr(_np.ndarray):
    """ndarray with `.at[...]`.  Results of arithmetic keep the subclass."""

    def __new__(cls, a):
        return _np.asarray(a).view(cls)

    @property
    def at(self):
        return _At(self)

    def __array_finalize__(self, obj):
        pass

    def __getitem__(self, idx):
        # jax clamps out-of-range integer indices of a gather instead of raising
        if isinstance(idx, tuple) and len(idx) == self.ndim and builtins.all(
                isinstance(i, (int, _np.integer)) for i in idx):
            idx = tuple(builtins.min(int(i), n - 1) if int(i) >= 0 else int(i) for i, n in zip(idx, self.shape))
        return _np.ndarray.__getitem__(self, idx)

    def reshape(self, *shape, **kw):
        if not shape:  # jax: x.reshape() -> scalar-shaped array
            return _np.ndarray.reshape(self, ())
        return _np.ndarray.reshape(self, *shape, **kw)


def _fix(x):
    """x64 disabled: no float64 / int64 can come out of a jnp function."""
    if isinstance(x, tuple):
        return tuple(_fix(v) for v in x)
    if isinstance(x, (_np.ndarray, _np.generic)):
        if not _Config.x64:
            if x.dtype == _np.float64:
                x = x.astype(_np.float32)
        if isinstance(x, _np.ndarray):
            return Arr(x)
        return x
    if isinstance(x, float):
        return _default_float()(x)
    return x


def _wrap(fn):
    def f(*a, **k):
        return _fix(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "fn")
    return f


def array(x, dtype=None, copy=True):
    a = _np.array(x, dtype=dtype)
    if dtype is None and a.dtype == _np.float64 and not _Config.x64:
        # python floats / lists of python floats -> float32; float64 arrays cannot exist when x64 is off
        a = a.astype(_np.float32)
    return Arr(a)


asarray = array


def zeros(shape, dtype=None): return Arr(_np.zeros(shape, dtype=dtype or _default_float()))
def ones(shape, dtype=None): return Arr(_np.ones(shape, dtype=dtype or _default_float()))
def zeros_like(x, dtype=None): return Arr(_np.zeros_like(x, dtype=dtype))
def ones_like(x, dtype=None): return Arr(_np.ones_like(x, dtype=dtype))
def eye(n, dtype=None): return Arr(_np.eye(n, dtype=dtype or _default_float()))
def arange(*a, dtype=None): return Arr(_np.arange(*a, dtype=dtype))
def linspace(a, b, n, dtype=None): return Arr(_np.linspace(a, b, n).astype(dtype or _default_float()))
def diag(x): return Arr(_np.diag(x))
def argsort(x, axis=-1): return Arr(_np.argsort(x, axis=axis, kind="stable"))
def sort(x, axis=-1): return Arr(_np.sort(x, axis=axis, kind="stable"))


for _name in ("sign floor ceil sqrt abs absolute exp log sin cos tan tanh sinh cosh arctan2 where add subtract multiply "
              "divide minimum maximum sum min max mean dot matmul column_stack concatenate stack squeeze reshape "
              "count_nonzero isclose rint square power transpose swapaxes meshgrid heaviside floor_divide mod "
              "logical_and logical_or logical_not any all isnan isfinite clip cumsum prod outer cross "
              "expand_dims ravel take amin amax nonzero round").split():
    globals()[_name] = _wrap(getattr(_np, _name))


def split(a, indices_or_sections, axis=0):
    return [Arr(v) for v in _np.split(_np.asarray(a), indices_or_sections, axis=axis)]


def nan_to_num(x, nan=0.0, posinf=None, neginf=None):
    return _fix(_np.nan_to_num(x, nan=nan, posinf=posinf, neginf=neginf))


class linalg:
    @staticmethod
    def det(a):
        return _fix(_np.linalg.det(a))

    @staticmethod
    def pinv(a, rcond=None):
        a = _np.asarray(a)
        if rcond is None:
            rcond = 10.0 * builtins.max(a.shape[-2:]) * _np.finfo(a.dtype).eps
        # jnp.linalg.pinv: SVD in the array's dtype, singular values <= rcond*max dropped
        u, s, vh = _np.linalg.svd(a, full_matrices=False)
        cutoff = rcond * (s.max() if s.size else 0.0)
        sinv = _np.where(s > cutoff, 1.0 / _np.where(s > cutoff, s, 1.0), 0.0).astype(a.dtype)
        return _fix((vh.T * sinv) @ u.T)

    @staticmethod
    def norm(x, axis=None):
        return _fix(_np.linalg.norm(x, axis=axis))

    @staticmethod
    def inv(a):
        return _fix(_np.linalg.inv(a))
