"""Field: a scalar or vector field on the extended slab grid, in device memory.

Mirrors skeletor.Field (reference skeletor/field.py:4-209): shape (myp, mx), dtype
float64 or a structured Float3/Float4; `active`, `trim()`, `copy_guards()`,
`translate()`, `boundaries_set`, `time`, `shear`.  Storage is one contiguous torch
CUDA tensor [myp][mx][nc] — byte-identical to the reference's NumPy structured
array — handed to libskeletor_b200 as a device pointer.
"""
import numpy as np
import torch

from . import _lib
from .array import DeviceArray, _index, _to_tensor


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _names_of(dtype):
    if dtype is None:
        return None
    dt = np.dtype(dtype)
    if dt.names is None:
        assert dt == np.float64, "only float64 fields are supported"
        return None
    assert all(dt[n] == np.float64 for n in dt.names)
    return tuple(dt.names)


class Field(DeviceArray):

    def __init__(self, grid, time=0.0, dtype=np.float64, _tensor=None, **kwds):
        self.grid = grid
        self.names = _names_of(dtype)
        self.np_dtype = np.dtype(dtype)
        if _tensor is None:
            _lib.require_cuda()
            shape = (grid.myp, grid.mx) if self.names is None else \
                (grid.myp, grid.mx, len(self.names))
            _tensor = torch.zeros(shape, dtype=torch.float64, device="cuda")
        DeviceArray.__init__(self, _tensor)
        # MPI-style neighbours (field.py:15-16)
        self.above = (grid.comm.rank + 1) % grid.comm.size
        self.below = (grid.comm.rank - 1) % grid.comm.size
        # Boolean indicating whether boundaries are set
        self.boundaries_set = False
        # Time of the field (reference tests sometimes pass `comm` here, Q8)
        self.time = time
        # Is there shear?
        self.shear = hasattr(grid, 'S')
        if self.shear:
            # Wave numbers for real-to-complex transforms (field.py:30-33)
            self.kx = 2*np.pi*np.fft.rfftfreq(grid.nx)/grid.dx
            self.y_kx = np.outer(grid.y, self.kx)

    # -- NumPy-like surface ---------------------------------------------------
    @property
    def dtype(self):
        return self.np_dtype

    @property
    def shape(self):
        return tuple(self.t.shape[:-1]) if self.names is not None else \
            tuple(self.t.shape)

    @property
    def nc(self):
        return 1 if self.names is None else len(self.names)

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().contiguous().cpu().numpy()
        if self.names is not None:
            a = a.view(self.np_dtype).reshape(a.shape[:-1])
        return a

    def _component(self, name):
        i = self.names.index(name)
        f = Field(self.grid, time=self.time, dtype=np.float64,
                  _tensor=self.t[..., i])
        f.boundaries_set = self.boundaries_set
        return f

    def _struct_to_tensor(self, val):
        """structured ndarray / Field / tuple / scalar -> tensor [...][nc]"""
        if isinstance(val, Field):
            return val.t
        if isinstance(val, (tuple, list)) and len(val) == self.nc and \
                all(np.isscalar(v) for v in val):
            return torch.tensor([float(v) for v in val], dtype=torch.float64,
                                device=self.t.device)
        a = np.asarray(val)
        if a.dtype.names is not None:
            a = np.ascontiguousarray(a).view(np.float64).reshape(a.shape + (self.nc,))
            return torch.as_tensor(a, device=self.t.device)
        # plain scalar / array broadcast over all components
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64),
                            device=self.t.device)
        return t.unsqueeze(-1) if t.dim() > 0 else t

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self._component(idx)
        if self.names is None:
            return DeviceArray.__getitem__(self, idx)
        # structured field: basic indexing over (y, x) returns a host copy
        return self.__array__()[idx]

    def __setitem__(self, idx, val):
        if isinstance(idx, str):
            self._component(idx)[...] = val
            return
        if self.names is None:
            DeviceArray.__setitem__(self, idx, val)
            return
        self.t[_index(idx, self.t.device)] = self._struct_to_tensor(val)

    def fill(self, val):
        if self.names is None:
            self.t.fill_(float(val))
        else:
            self.t[...] = self._struct_to_tensor(val)

    def __iadd__(self, other):
        self.t.add_(other.t if isinstance(other, Field) else
                    self._struct_to_tensor(other) if self.names else
                    _to_tensor(other, self.t))
        return self

    @property
    def ptr(self):
        assert self.t.is_contiguous(), "kernel called on a strided component view"
        return self.t.data_ptr()

    # -- active region ----------------------------------------------------------
    def _active_t(self):
        g = self.grid
        return self.t[g.lby:g.uby, g.lbx:g.ubx]

    @property
    def active(self):
        """Host copy of the active cells (the reference returns a NumPy view; use the
        setter — `f.active = rhs` — to write)."""
        a = self._active_t().contiguous().cpu().numpy()
        if self.names is not None:
            a = a.view(self.np_dtype).reshape(a.shape[:-1])
        return a

    @active.setter
    def active(self, rhs):
        if self.names is None:
            self._active_t()[...] = _to_tensor(rhs, self.t)
        else:
            self._active_t()[...] = self._struct_to_tensor(rhs)

    def trim(self):
        return self.active.squeeze()

    # -- guard cells --------------------------------------------------------------
    def _halo_exchange(self, up_rows, down_rows):
        """send `up_rows` to the rank above and `down_rows` to the rank below;
        returns (from_below, from_above).  Replaces send_up/send_dn (field.py:52-58)"""
        return self.grid.comm.halo_exchange(up_rows, down_rows)

    def _pack_rows(self, iy0, nrows):
        g = self.grid
        out = torch.empty((nrows, g.nx, self.nc) if self.names else (nrows, g.nx),
                          dtype=torch.float64, device=self.t.device)
        _lib.call("skb_pack_rows", self.ptr, self.nc, g.c, iy0, nrows,
                  out.data_ptr(), _stream())
        return out

    # -- the reference's piecewise guard API (compatibility; the step path uses the
    #    fused copy_guards / add_guards below) -----------------------------------------
    def send_up(self, sendbuf):
        """object-level sendrecv to the rank above (field.py:52-54)"""
        return self.grid.comm.sendrecv(sendbuf, dest=self.above, source=self.below)

    def send_dn(self, sendbuf):
        """object-level sendrecv to the rank below (field.py:56-58)"""
        return self.grid.comm.sendrecv(sendbuf, dest=self.below, source=self.above)

    def copy_guards_x(self):
        """Periodic boundary condition in x, all rows (field.py:73-83)"""
        _lib.call("skb_copy_guards_x_rows", self.ptr, self.nc, self.grid.c, 0,
                  self.grid.myp, _stream())

    def copy_guards_y(self):
        """Periodic boundary condition in y, active x only (field.py:85-98)"""
        g = self.grid
        t = self.t
        ax = slice(g.lbx, g.ubx)
        up = t[g.uby - g.lby:g.uby, ax].contiguous()     # my last active rows
        dn = t[g.lby:2*g.lby, ax].contiguous()           # my first active rows
        if g.comm.size > 1:
            from_below, from_above = self._halo_exchange(up, dn)
        else:
            from_below, from_above = up, dn
        t[:g.lby, ax] = from_below
        t[g.uby:, ax] = from_above

    def copy_guards_old(self):
        """reference test helper (field.py:128-151): same result as copy_guards when
        there is no shear"""
        shear, self.shear = self.shear, False
        try:
            self.copy_guards()
        finally:
            self.shear = shear

    def copy_guards(self):
        "Copy data to guard cells from corresponding active cells (field.py:100-126)."
        assert not self.boundaries_set, 'Boundaries are already set!'
        g = self.grid
        gc = g.c
        if g.comm.size == 1:
            _lib.call("skb_copy_guards", self.ptr, self.nc, gc, None, None, _stream())
        else:
            # my last active rows go up (they are the upper neighbour's lower
            # guards), my first active rows go down (field.py:91-98)
            up = self._pack_rows(g.uby - g.lby, g.lby)
            dn = self._pack_rows(g.lby, g.lby)
            from_below, from_above = self._halo_exchange(up, dn)
            _lib.call("skb_copy_guards", self.ptr, self.nc, gc,
                      from_below.data_ptr(), from_above.data_ptr(), _stream())
        if self.shear:
            # Translate the y-ghostzones (field.py:113-124)
            if g.comm.rank == g.comm.size - 1:
                self._translate_boundary(-g.Ly*g.S*self.time, g.uby, g.lby)
            if g.comm.rank == 0:
                self._translate_boundary(+g.Ly*g.S*self.time, 0, g.lby)
        self.boundaries_set = True

    def _translate_boundary(self, trans, iy0, nrows):
        """Spectral shift of guard rows [iy0, iy0+nrows) by `trans` along x
        (field.py:153-171): irfft(exp(-i kx trans) rfft(row)), all components at
        once through cuFFT, then refresh the rows' x guards."""
        g = self.grid
        fac = torch.as_tensor(np.exp(-1j*self.kx*trans), device=self.t.device)
        rows = self.t[iy0:iy0 + nrows, g.lbx:g.ubx]
        if self.names is not None:
            fac = fac.unsqueeze(-1)
        hat = torch.fft.rfft(rows, dim=1)
        rows[...] = torch.fft.irfft(fac*hat, n=g.nx, dim=1)
        _lib.call("skb_copy_guards_x_rows", self.ptr, self.nc, g.c, iy0, nrows,
                  _stream())

    def translate(self, time):
        """Translation along x by -S*t*y of the whole field (field.py:173-189)."""
        if not self.shear:
            return
        g = self.grid
        act = self._active_t()
        fac = torch.as_tensor(np.exp(1j*g.S*time*self.y_kx), device=self.t.device)
        if self.names is not None:
            fac = fac.unsqueeze(-1)
        act[...] = torch.fft.irfft(torch.fft.rfft(act, dim=1)*fac, n=g.nx, dim=1)
        self.boundaries_set = False

    def translate_vector(self, time):
        self.translate(time)
