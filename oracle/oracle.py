"""CPU oracle for skeletor's particle hot path (numpy + oracle/skeletor_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline legs may import this module.  The product
(skeletor_b200/) never imports it and has no CPU fallback.

Parity status: PINNED against the unmodified reference compiled into
oracle/_ref (tests/test_oracle_vs_reference.py) and against the committed
fixtures tests/golden/*.npz (tests/test_oracle_golden.py).

The per-particle / per-cell arithmetic is in skeletor_oracle.c; the pieces the
reference itself does with NumPy slicing (guard cells, spectral shear remap,
normalisation, Ohm, Faraday) are restated here with NumPy, citing the reference
file:line each follows.  The multi-rank communication steps (halo exchange,
cppmove2 migration) are restated as an "N-slab emulator" operating on lists of
per-slab arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "skeletor_oracle.c")
_LIB = os.path.join(HERE, "_build", "libskeletor_oracle.so")

Float3 = np.dtype([("x", "f8"), ("y", "f8"), ("z", "f8")])
Float4 = np.dtype([("t", "f8"), ("x", "f8"), ("y", "f8"), ("z", "f8")])
Particle = np.dtype([("x", "f8"), ("y", "f8"), ("vx", "f8"), ("vy", "f8"),
                     ("vz", "f8")], align=True)


def build(force=False):
    """gcc -O2, no FMA contraction (matches the reference's x86-64 SSE2 code)."""
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           "-o", _LIB, _SRC, "-lm"])
    return _LIB


class OGrid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nyp", C.c_int),
                ("noff", C.c_int), ("lbx", C.c_int), ("lby", C.c_int),
                ("ubx", C.c_int), ("uby", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("Lx", C.c_double),
                ("Ly", C.c_double), ("x0", C.c_double), ("y0", C.c_double),
                ("edges", C.c_double * 2)]


class Grid:
    """Slab geometry, reference skeletor/grid.py:7-67 (rank/size explicit)."""

    def __init__(self, nx, ny, rank=0, size=1, lbx=1, lby=1, Lx=1.0, Ly=1.0,
                 x0=0.0, y0=0.0, S=None, Omega=None):
        self.nx, self.ny, self.rank, self.size = nx, ny, rank, size
        self.Lx, self.Ly, self.x0, self.y0 = Lx, Ly, x0, y0
        self.dx, self.dy = Lx/nx, Ly/ny
        self.nyp = ny//size
        self.noff = self.nyp*rank
        self.edges = [float(self.noff), float(self.noff + self.nyp)]
        self.lbx, self.ubx = lbx, lbx + nx
        self.lby, self.uby = lby, lby + self.nyp
        self.mx, self.myp = nx + 2*lbx, self.nyp + 2*lby
        if S is not None:
            self.S = S
            self.Omega = Omega if Omega is not None else 0.0

    @property
    def shear(self):
        return hasattr(self, "S")

    @property
    def x(self):
        return self.x0 + (np.arange(self.nx) + 0.5)*self.dx

    @property
    def y(self):
        return self.y0 + (np.arange(self.noff, self.noff + self.nyp) + 0.5)*self.dy

    @property
    def yg(self):
        r = np.arange(self.noff - self.lby, self.noff + self.nyp + self.lby)
        return self.y0 + (r + 0.5)*self.dy

    def c(self):
        g = OGrid(self.nx, self.ny, self.nyp, self.noff, self.lbx, self.lby,
                  self.ubx, self.uby, self.dx, self.dy, self.Lx, self.Ly,
                  self.x0, self.y0)
        g.edges[0], g.edges[1] = self.edges
        return g

    def field(self, dtype=np.float64):
        return np.zeros((self.myp, self.mx), dtype)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for name in ("orc_push", "orc_drift", "orc_periodic_x",
                     "orc_shear_periodic_y", "orc_calculate_ihole",
                     "orc_deposit", "orc_push_and_deposit", "orc_gradient",
                     "orc_curl", "orc_divergence", "orc_interp"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _chk(a, dtype=None):
    assert a.flags["C_CONTIGUOUS"]
    if dtype is not None:
        assert a.dtype == dtype, (a.dtype, dtype)
    return a


# --- particle kernels ------------------------------------------------------

def push(part, E, B, grid, order, qtmh, dt, modified=False, Omega=0.0, S=0.0):
    """boris_push_* / modified_boris_push_* (particle_push.pyx:4-156)."""
    _chk(part, Particle), _chk(E, Float3), _chk(B, Float3)
    g = grid.c()
    lib().orc_push(_p(part), C.c_long(part.shape[0]), _p(E), _p(B),
                   C.byref(g), C.c_int(order), C.c_double(qtmh),
                   C.c_double(dt), C.c_int(int(modified)), C.c_double(Omega),
                   C.c_double(S))


def drift(part, grid, dt):
    g = grid.c()
    lib().orc_drift(_p(_chk(part, Particle)), C.c_long(part.shape[0]),
                    C.byref(g), C.c_double(dt))


def periodic_x(part, grid):
    g = grid.c()
    lib().orc_periodic_x(_p(_chk(part, Particle)), C.c_long(part.shape[0]),
                         C.byref(g))


def shear_periodic_y(part, grid, S, t):
    g = grid.c()
    lib().orc_shear_periodic_y(_p(_chk(part, Particle)),
                               C.c_long(part.shape[0]), C.byref(g),
                               C.c_double(S), C.c_double(t))


def calculate_ihole(part, ihole, grid):
    g = grid.c()
    _chk(ihole, np.int32)
    lib().orc_calculate_ihole(_p(_chk(part, Particle)),
                              C.c_long(part.shape[0]), _p(ihole),
                              C.c_int(ihole.shape[0] - 1), C.byref(g))


def deposit(part, cur, grid, order, S=0.0):
    """deposit_cic / deposit_tsc (deposit.pyx:6-34); accumulates into cur."""
    g = grid.c()
    lib().orc_deposit(_p(_chk(part, Particle)), C.c_long(part.shape[0]),
                      _p(_chk(cur, Float4)), C.byref(g), C.c_int(order),
                      C.c_double(S))


def push_and_deposit(part, E, B, grid, order, qtmh, dt, ihole, cur, S, update):
    g = grid.c()
    _chk(ihole, np.int32)
    lib().orc_push_and_deposit(
        _p(_chk(part, Particle)), C.c_long(part.shape[0]), _p(_chk(E, Float3)),
        _p(_chk(B, Float3)), C.byref(g), C.c_int(order), C.c_double(qtmh),
        C.c_double(dt), _p(ihole), C.c_int(ihole.shape[0] - 1),
        _p(_chk(cur, Float4)), C.c_double(S), C.c_int(int(update)))


# --- finite differences ------------------------------------------------------

def _plane(f, name):
    """component `name` of an interleaved field -> (pointer, element stride)."""
    es = f.dtype.itemsize//8
    off = f.dtype.fields[name][1]
    return C.c_void_p(f.ctypes.data + off), C.c_int(es)


def gradient(f, grad, grid):
    """f: plain float64 plane (finite_difference.pyx:5-13)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    g = grid.c()
    lib().orc_gradient(_p(f), C.c_int(1), _p(_chk(grad, Float3)), C.byref(g))


def curl(f, out, grid, down=True):
    g = grid.c()
    _chk(f), _chk(out, Float3)
    (px, es), (py, _), (pz, _) = _plane(f, "x"), _plane(f, "y"), _plane(f, "z")
    lib().orc_curl(px, py, pz, es, _p(out), C.byref(g), C.c_int(int(down)))


def divergence(f, div, grid):
    g = grid.c()
    _chk(f), _chk(div, np.float64)
    (px, es), (py, _) = _plane(f, "x"), _plane(f, "y")
    lib().orc_divergence(px, py, es, _p(div), C.byref(g))


def unstagger(f, out, grid):
    g = grid.c()
    (px, es), (py, _), (pz, _) = _plane(f, "x"), _plane(f, "y"), _plane(f, "z")
    lib().orc_interp(px, py, pz, es, _p(_chk(out, Float3)), C.byref(g), C.c_int(0))


def stagger(f, out, grid):
    g = grid.c()
    (px, es), (py, _), (pz, _) = _plane(f, "x"), _plane(f, "y"), _plane(f, "z")
    lib().orc_interp(px, py, pz, es, _p(_chk(out, Float3)), C.byref(g), C.c_int(1))


# --- guard cells (NumPy in the reference too) -------------------------------
# All functions take `fields`: a list with one array per slab (rank order), and
# `grids`: the matching list of Grid objects.  size == 1 is the single-rank case.

def _names(f):
    return f.dtype.names


def translate_boundary(f, grid, trans, iy):
    """Field._translate_boundary, field.py:153-171 (spectral shift of one row)."""
    g = grid
    kx = 2*np.pi*np.fft.rfftfreq(g.nx)/g.dx          # field.py:30
    fac = np.exp(-1j*kx*trans)
    if _names(f) is None:
        f[iy, g.lbx:g.ubx] = np.fft.irfft(fac*np.fft.rfft(f[iy, g.lbx:g.ubx]))
    else:
        for dim in _names(f):
            f[iy, g.lbx:g.ubx][dim] = np.fft.irfft(
                fac*np.fft.rfft(f[iy, g.lbx:g.ubx][dim]))
    f[iy, g.ubx:] = f[iy, g.lbx:g.lbx + g.lbx]
    f[iy, :g.lbx] = f[iy, g.ubx - g.lbx:g.ubx]


def copy_guards(fields, grids, time=0.0):
    """Field.copy_guards, field.py:73-126, for all slabs at once."""
    n = len(fields)
    g = grids[0]
    # copy_guards_y: local periodic copy, then neighbour overwrite (field.py:85-98)
    for f, g in zip(fields, grids):
        for iy in range(g.lby - 1, -1, -1):
            f[iy, g.lbx:g.ubx] = f[iy + g.nyp, g.lbx:g.ubx]
        for iy in range(g.uby, g.uby + g.lby):
            f[iy, g.lbx:g.ubx] = f[iy - g.nyp, g.lbx:g.ubx]
    # send_dn(self[uby:]) : my upper guards go down, I receive from above
    up_guard = [f[g.uby:, g.lbx:g.ubx].copy() for f, g in zip(fields, grids)]
    for r in range(n):
        g = grids[r]
        fields[r][g.uby:, g.lbx:g.ubx] = up_guard[(r + 1) % n]
    lo_guard = [f[:g.lby, g.lbx:g.ubx].copy() for f, g in zip(fields, grids)]
    for r in range(n):
        g = grids[r]
        fields[r][:g.lby, g.lbx:g.ubx] = lo_guard[(r - 1) % n]
    # copy_guards_x (field.py:73-83)
    for f, g in zip(fields, grids):
        for ix in range(g.lbx - 1, -1, -1):
            f[:, ix] = f[:, ix + g.nx]
        for ix in range(g.ubx, g.ubx + g.lbx):
            f[:, ix] = f[:, ix - g.nx]
    # shear remap of the edge ranks' y-guards (field.py:113-124)
    if grids[0].shear:
        g = grids[-1]
        for iy in range(g.uby, g.uby + g.lby):
            translate_boundary(fields[-1], g, -g.Ly*g.S*time, iy)
        g = grids[0]
        for iy in range(0, g.lby):
            translate_boundary(fields[0], g, +g.Ly*g.S*time, iy)


def add_guards(fields, grids, time=0.0):
    """Sources.add_guards, sources.py:91-150, for all slabs at once."""
    n = len(fields)
    # add_guards_x over all rows (sources.py:91-101)
    for f, g in zip(fields, grids):
        for dim in _names(f):
            a = f[dim]
            for ix in range(g.lbx):
                a[:, ix + g.nx] += a[:, ix]
            for ix in range(g.ubx + g.lbx - 1, g.ubx - 1, -1):
                a[:, ix - g.nx] += a[:, ix]
    # shear remap (sources.py:128-139)
    if grids[0].shear:
        g = grids[-1]
        for iy in range(g.uby, g.uby + g.lby):
            translate_boundary(fields[-1], g, g.Ly*g.S*time, iy)
        g = grids[0]
        for iy in range(0, g.lby):
            translate_boundary(fields[0], g, -g.Ly*g.S*time, iy)
    # add_guards_y (sources.py:103-115): guards travel to the neighbour, then fold
    up = [f[g.uby:, g.lbx:g.ubx].copy() for f, g in zip(fields, grids)]
    lo = [f[:g.lby, g.lbx:g.ubx].copy() for f, g in zip(fields, grids)]
    for r in range(n):
        g = grids[r]
        fields[r][g.uby:, g.lbx:g.ubx] = up[(r - 1) % n]   # send_up: recv from below
        fields[r][:g.lby, g.lbx:g.ubx] = lo[(r + 1) % n]   # send_dn: recv from above
    for f, g in zip(fields, grids):
        for dim in _names(f):
            a = f[dim]
            for iy in range(g.lby):
                a[iy + g.nyp, g.lbx:g.ubx] += a[iy, g.lbx:g.ubx]
            for iy in range(g.uby + g.lby - 1, g.uby - 1, -1):
                a[iy - g.nyp, g.lbx:g.ubx] += a[iy, g.lbx:g.ubx]
        # zero guards (sources.py:147-150)
        f[:g.lby, :] = 0.0
        f[g.uby:, :] = 0.0
        f[:, g.ubx:] = 0.0
        f[:, :g.lbx] = 0.0


def normalize(fields, grids, Ns, charge, n0):
    """Sources.normalize, sources.py:52-63 (global N, guards scaled too)."""
    g = grids[0]
    N = int(sum(Ns))
    fac = charge*n0*g.nx*g.ny/N
    for f in fields:
        for dim in _names(f):
            f[dim] *= fac


# --- particle migration: cppmove2 as a set operation -------------------------

def move(parts, Ns, grids):
    """Restates WHICH particles cppmove2 moves and what it does to y.

    picksc/ppic2/pplib2.c:607-981: a particle with y < edges[0] goes to the rank
    below (y += ny on rank 0, :676-677), y >= edges[1] to the rank above
    (y -= ny on the last rank, :692-693); particles keep moving until they are
    inside their slab (multi-hop, :756-866).  The resulting ORDER inside each
    slab is not contractual (reference tests sort, tests/test_skeletor.py:144);
    this function returns per-slab arrays in the order [stayers..., arrivals...].

    parts: list of Particle arrays (live prefix Ns[r]).  Returns new lists.
    """
    n = len(parts)
    ny = float(grids[0].ny)
    if n == 1:
        # Single rank: reproduce cppmove2's resulting ORDER too, so multi-step
        # single-rank runs stay bit-identical to the reference.  nvp == 1 branch
        # (pplib2.c:715-730): rbufl = sbufr, rbufr = sbufl; every leaver comes
        # back after the +-ny wrap, so holes (ascending, calculate_ihole order)
        # are refilled first from rbufl (the up-goers, :883-891) then from rbufr
        # (the down-goers, :899-907); nothing is appended or compacted.
        p = parts[0].copy()
        live = p[:Ns[0]]
        e0, e1 = grids[0].edges
        dn = live["y"] < e0
        up = ~dn & (live["y"] >= e1)
        holes = np.flatnonzero(dn | up)
        pd, pu = live[dn].copy(), live[up].copy()
        pd["y"] += ny
        pu["y"] -= ny
        back = np.concatenate([pu, pd])
        assert ((back["y"] >= e0) & (back["y"] < e1)).all(), "multi-hop on 1 rank"
        live[holes] = back
        return [p], [Ns[0]]
    stay = []
    flying = [[] for _ in range(n)]          # arrivals per destination rank
    for r in range(n):
        p = parts[r][:Ns[r]]
        e0, e1 = grids[r].edges
        dn = p["y"] < e0
        up = ~dn & (p["y"] >= e1)
        stay.append(p[~dn & ~up].copy())
        pd, pu = p[dn].copy(), p[up].copy()
        if r == 0:
            pd["y"] += ny
        if r == n - 1:
            pu["y"] -= ny
        flying[(r - 1) % n].append(pd)
        flying[(r + 1) % n].append(pu)
    arrived = [[] for _ in range(n)]
    for it in range(2000):
        nxt = [[] for _ in range(n)]
        busy = False
        for r in range(n):
            e0, e1 = grids[r].edges
            for q in flying[r]:
                if q.size == 0:
                    continue
                dn = q["y"] < e0
                up = ~dn & (q["y"] >= e1)
                arrived[r].append(q[~dn & ~up])
                qd, qu = q[dn].copy(), q[up].copy()
                if qd.size or qu.size:
                    busy = True
                if r == 0:
                    qd["y"] += ny
                if r == n - 1:
                    qu["y"] -= ny
                nxt[(r - 1) % n].append(qd)
                nxt[(r + 1) % n].append(qu)
        flying = nxt
        if not busy:
            break
    out, outN = [], []
    for r in range(n):
        new = np.concatenate([stay[r]] + arrived[r]) if arrived[r] else stay[r]
        buf = np.zeros(parts[r].shape[0], Particle)
        assert new.shape[0] <= buf.shape[0], "particle array overflow"
        buf[:new.shape[0]] = new
        out.append(buf)
        outN.append(new.shape[0])
    return out, outN


# --- field solvers (NumPy in the reference) ----------------------------------

def ohm(sources, B, E, grid, charge=1.0, temperature=0.0, eta=0.0, time=0.0):
    """Ohm.__call__, ohm.py:35-75, single slab (guards of sources and B set).

    Returns E with active cells updated (guards untouched apart from what the
    whole-array NumPy ops of the reference do to them)."""
    alpha = temperature/charge
    gradient(np.log(sources["t"]), E, grid)
    E["x"] *= -alpha
    E["y"] *= -alpha
    Je = grid.field(Float3)
    curl(B, Je, grid, down=True)
    for d in "xyz":
        E[d] += eta*Je[d]
    copy_guards([Je], [grid], time)
    for d in "xyz":
        Je[d] -= sources[d]
        Je[d] /= sources["t"]
    Bc = grid.field(Float3)
    unstagger(B, Bc, grid)
    E["x"] += Je["y"]*Bc["z"] - Je["z"]*Bc["y"]
    E["y"] += Je["z"]*Bc["x"] - Je["x"]*Bc["z"]
    E["z"] += Je["x"]*Bc["y"] - Je["y"]*Bc["x"]
    return Je, Bc


def faraday(E, B, grid, dt):
    """Faraday.__call__, faraday.py:16-30."""
    dB = grid.field(Float3)
    curl(E, dB, grid, down=False)
    for d in "xyz":
        B[d] -= dB[d]*dt
    return dB


# --- the sort oracle -----------------------------------------------------------

def cell_keys(part, grid, order, tile_log2=(4, 4)):
    """Integer sort key of each particle: tile-major index of the deposit/E-gather
    stencil base cell.  (New component, no reference counterpart; the GPU sort is
    checked bit-exactly against np.argsort(kind='stable') of these keys.)"""
    ox = grid.lbx - 0.5
    oy = grid.lby - 0.5 - grid.noff
    x = part["x"] + ox
    y = part["y"] + oy
    if order == 2:
        x = x + 0.5
        y = y + 0.5
    ix = x.astype(np.int32)
    iy = y.astype(np.int32)
    ix = np.clip(ix, 0, grid.mx - 1)
    iy = np.clip(iy, 0, grid.myp - 1)
    lx, ly = tile_log2
    ntx = (grid.mx + (1 << lx) - 1) >> lx
    key = (((iy >> ly)*ntx + (ix >> lx)) << (lx + ly)) \
        | ((iy & ((1 << ly) - 1)) << lx) | (ix & ((1 << lx) - 1))
    return key.astype(np.int32)
