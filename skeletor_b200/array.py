"""DeviceArray: a NumPy-flavoured proxy over a torch CUDA tensor view.

The reference's Field / Particles ARE NumPy arrays and tests poke them directly
(`ions['vx'] = f(x)`, `E['x'].active = ...`, `sources.rho.trim().sum()`,
`B['x'] += ...`).  This proxy keeps that surface on device memory: reads copy to
the host as NumPy (`__array__`), writes and in-place operators go through to the
device tensor, out-of-place arithmetic is evaluated by NumPy on the host copy
(identical IEEE semantics).  It is glue for set-up and diagnostics — the hot path
never goes through it.
"""
import numpy as np
import torch


def _to_tensor(val, like):
    """anything -> tensor broadcastable against `like` (device, float64/int)"""
    if isinstance(val, DeviceArray):
        return val.t
    if isinstance(val, torch.Tensor):
        return val.to(like.device)
    a = np.asarray(val)
    if a.dtype.names is not None:
        raise TypeError("structured value for a plain device array")
    return torch.as_tensor(np.ascontiguousarray(a), device=like.device).to(like.dtype)


def _index(idx, device):
    """translate NumPy-style indices (incl. DeviceArray/ndarray masks) for torch"""
    def one(i):
        if isinstance(i, DeviceArray):
            return i.t
        if isinstance(i, np.ndarray):
            return torch.as_tensor(i, device=device)
        if isinstance(i, (np.integer,)):
            return int(i)
        return i
    if isinstance(idx, tuple):
        return tuple(one(i) for i in idx)
    return one(idx)


class DeviceArray:
    __array_priority__ = 1000

    def __init__(self, t, on_write=None):
        self.t = t
        self._on_write = on_write

    # -- conversion ---------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def numpy(self):
        return self.__array__()

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    def __len__(self):
        return self.t.shape[0]

    def _wrap(self, t):
        return DeviceArray(t, self._on_write)

    def _wrote(self):
        if self._on_write is not None:
            self._on_write()

    # -- indexing -------------------------------------------------------------
    def __getitem__(self, idx):
        r = self.t[_index(idx, self.t.device)]
        if r.dim() == 0:
            return r.item()
        return self._wrap(r)

    def __setitem__(self, idx, val):
        self.t[_index(idx, self.t.device)] = _to_tensor(val, self.t)
        self._wrote()

    def fill(self, val):
        self.t.fill_(float(val))
        self._wrote()

    def copy(self):
        return self.__array__().copy()

    def squeeze(self):
        return self.__array__().squeeze()

    # -- in-place arithmetic: on the device -------------------------------------
    def __iadd__(self, o):
        self.t.add_(_to_tensor(o, self.t)); self._wrote(); return self

    def __isub__(self, o):
        self.t.sub_(_to_tensor(o, self.t)); self._wrote(); return self

    def __imul__(self, o):
        self.t.mul_(_to_tensor(o, self.t)); self._wrote(); return self

    def __itruediv__(self, o):
        self.t.div_(_to_tensor(o, self.t)); self._wrote(); return self

    # -- reductions -----------------------------------------------------------
    def sum(self, *a, **k):
        return self.__array__().sum(*a, **k)

    def mean(self, *a, **k):
        return self.__array__().mean(*a, **k)

    def min(self, *a, **k):
        return self.__array__().min(*a, **k)

    def max(self, *a, **k):
        return self.__array__().max(*a, **k)

    def std(self, *a, **k):
        return self.__array__().std(*a, **k)

    def __repr__(self):
        return "DeviceArray(%r)" % (self.__array__(),)


def _binary(name):
    def f(self, o):
        return getattr(self.__array__(), name)(np.asarray(o))
    f.__name__ = name
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__",
           "__truediv__", "__rtruediv__", "__pow__", "__rpow__", "__lt__", "__le__",
           "__gt__", "__ge__", "__eq__", "__ne__", "__mod__", "__floordiv__"):
    setattr(DeviceArray, _n, _binary(_n))
DeviceArray.__neg__ = lambda self: -self.__array__()
DeviceArray.__abs__ = lambda self: abs(self.__array__())
DeviceArray.__hash__ = None
