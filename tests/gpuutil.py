"""Helpers for the -m gpu parity tests: move oracle-layout data to the device and
call the C ABI (libskeletor_b200.so) directly."""
import ctypes as C

import numpy as np
import torch

from oracle import oracle as orc
from skeletor_b200 import _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def cgrid(g):
    """skb_grid_t from an oracle Grid"""
    c = _lib.GridT(g.nx, g.ny, g.nyp, g.noff, g.lbx, g.lby, g.ubx, g.uby, g.dx, g.dy,
                   g.Lx, g.Ly, g.x0, g.y0)
    c.edges[0], c.edges[1] = g.edges
    return c


def soa(part, cap=None):
    """oracle AoS particles -> [5][cap] device tensor"""
    a = np.ascontiguousarray(part).view(np.float64).reshape(-1, 5)
    cap = cap or a.shape[0]
    t = torch.zeros((5, cap), dtype=torch.float64, device="cuda")
    t[:, :a.shape[0]] = torch.as_tensor(a.T.copy(), device="cuda")
    return t


def cparts(t):
    n = t.shape[1]*8
    p = t.data_ptr()
    return _lib.ParticlesT(p, p + n, p + 2*n, p + 3*n, p + 4*n)


def aos(t, n=None):
    """[5][cap] device tensor -> oracle AoS structured array"""
    a = t[:, :n].t().contiguous().cpu().numpy() if n is not None else \
        t.t().contiguous().cpu().numpy()
    return np.ascontiguousarray(a).view(orc.Particle).reshape(-1)


def dev(field):
    """structured / plain ndarray field -> contiguous device tensor [myp][mx][nc]"""
    a = np.ascontiguousarray(field)
    if a.dtype.names is not None:
        a = a.view(np.float64).reshape(a.shape + (len(a.dtype.names),))
    return torch.as_tensor(a.copy(), device="cuda")


def host(t, dtype=None):
    a = t.contiguous().cpu().numpy()
    if dtype is not None:
        a = a.view(dtype).reshape(a.shape[:-1])
    return a


class Tiling:
    """runs skb_tile_sort and keeps the buffers alive"""

    def __init__(self, g, order, tlx=4, tly=4, chunk=2048):
        self.g, self.order, self.tlx, self.tly, self.chunk = g, order, tlx, tly, chunk
        ntx, nty = C.c_int(), C.c_int()
        _lib.load().skb_tile_geometry(cgrid(g), tlx, tly, C.byref(ntx), C.byref(nty))
        self.ntx, self.nty = ntx.value, nty.value
        nt = self.ntx*self.nty
        i32 = dict(dtype=torch.int32, device="cuda")
        self.cell_counts = torch.zeros((nt << (tlx + tly)) + 1, **i32)
        self.block_sums = torch.zeros(4100, **i32)
        self.tile_offsets = torch.zeros(nt + 1, **i32)
        self.chunk_first = None
        self.n = 0
        self.use_cells = True

    def sort(self, t, n):
        out = torch.zeros_like(t)
        self.chunk_first = torch.zeros(n//self.chunk + 2, dtype=torch.int32, device="cuda")
        _lib.call("skb_tile_sort", cparts(t), cparts(out), n, cgrid(self.g), self.order,
                  self.tlx, self.tly, self.chunk, self.cell_counts.data_ptr(),
                  self.block_sums.data_ptr(), self.tile_offsets.data_ptr(),
                  self.chunk_first.data_ptr(), 0, None, stream())
        self.n = n
        return out

    def c(self):
        return C.pointer(_lib.TilingT(self.tile_offsets.data_ptr(),
                                      self.chunk_first.data_ptr(),
                                      self.cell_counts.data_ptr() if self.use_cells else None,
                                      self.ntx, self.nty,
                                      self.tlx, self.tly, self.chunk, self.n))


def sorted_rows(p):
    a = np.ascontiguousarray(p).view(np.float64).reshape(-1, 5)
    return a[np.lexsort((a[:, 4], a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
