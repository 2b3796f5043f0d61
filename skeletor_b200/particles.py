"""Particles: the ions of one y-slab, structure-of-arrays in device memory.

Mirrors skeletor.Particles (reference skeletor/particles.py:5-265): same
constructor, attributes (N, time, charge, mass, n0, order, ihole, sbufl/r, rbufl/r,
info, sources) and methods (initialize, push, push_modified, push_and_deposit,
drift, periodic_x, periodic_y, shear_periodic_y, move, deposit).  The reference is
a NumPy structured array of AoS records; here the five coordinates are five
contiguous device arrays (one [5][Nmax] tensor, plus a second one the tile sort
ping-pongs with), and `ions['x']`, `ions[:N]` return device-backed proxies.

Every method that changes positions ends with the tile sort (skb_tile_sort), so
push gathers from shared-memory field tiles and deposit accumulates runs of
same-cell particles in registers.  The ordering is a performance property only:
kernels are correct for any order (user code may overwrite coordinates at will).

With `gapped = True` the particles live in per-cell slot ranges between push calls
(skb_push_gapped keeps them ordered without a move pass); every other method and every
NumPy-style access first converts back to the dense arrays described above.
"""
import ctypes as C
import os
from warnings import warn

import numpy as np
import torch

from . import _lib
from .array import DeviceArray
from .field import _stream
from .sources import Sources
from .types import Particle

# tile = 2^TLX x 2^TLY stencil-base cells; chunk = particles per work-item unit
TLX, TLY, CHUNK = 4, 4, 2048

_ROW = {"x": 0, "y": 1, "vx": 2, "vy": 3, "vz": 4}


class ParticleSlice:
    """ions[a:b] — converts to a structured NumPy array; fields are device proxies"""

    def __init__(self, parent, sl):
        self.parent, self.sl = parent, sl

    def __array__(self, dtype=None, copy=None):
        self.parent._dense()
        a = self.parent._data[:, self.sl].t().contiguous().cpu().numpy()
        out = np.zeros(a.shape[0], Particle)
        for k, r in _ROW.items():
            out[k] = a[:, r]
        return out

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.parent[key][self.sl]
        return np.asarray(self)[key]

    def __setitem__(self, key, val):
        self.parent[key][self.sl] = val

    def __len__(self):
        return len(range(*self.sl.indices(self.parent.size)))

    @property
    def shape(self):
        return (len(self),)

    @property
    def size(self):
        return len(self)


class ParticleRow(DeviceArray):
    """ions['x'] — a LIVE view of one coordinate row, like the reference's NumPy field
    view: the tensor is looked up on every access, so a proxy kept across push / sort /
    layout changes (all of which swap the underlying buffers) never goes stale."""

    def __init__(self, parent, row):
        self._parent, self._row = parent, row
        self._on_write = parent._touched

    @property
    def t(self):
        self._parent._dense()
        return self._parent._data[self._row]


class Particles:
    """Container class for particles in a given subdomain"""

    def __init__(self, manifold, Nmax,
                 time=0.0, charge=1.0, mass=1.0, n0=1.0, order=1, nbmax=None):
        """Same signature as the reference (particles.py:10-11) plus `nbmax`: the
        reference always sizes the exchange buffers as 0.1*Nmax (particles.py:22),
        which is tens of GB at 1e9 particles; nbmax overrides that."""
        _lib.require_cuda()
        msg = 'Interpolation order {} needs more guard layers'.format(order)
        # The number of guard layers on each side needs to be equal to
        # int(ceil(order*0.5 + 0.5)) (particles.py:15-19)
        assert manifold.lbx >= order//2 + 1, msg

        Nmax = int(Nmax)
        # particle indices are 32-bit on the device (like the reference's C ints,
        # pplib2.c:673), one slab holds at most ~2.1e9 particles
        assert Nmax < 2**31 - 2**20, "Nmax too large for one slab: use more ranks"
        # Size of buffer for passing particles between processors
        nbmax = int(max(0.1*Nmax, 1)) if nbmax is None else int(max(nbmax, 1))
        # Size of ihole buffer for particles leaving processor
        ntmax = 2*nbmax
        self.nbmax, self.ntmax = nbmax, ntmax

        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        # phase space coordinates, SoA: rows x, y, vx, vy, vz
        self._data = torch.zeros((5, Nmax), **f64)
        # second [5][Nmax] tensor: the out-of-place tile sort, gap_build and densify ping-pong
        # with it.  Allocated on first use and released again once the particles sit in
        # the gapped layout (nothing needs it there): steady-state footprint of a big run
        # = ONE particle tensor (config 5 on one GPU: 58 GB instead of 117 GB).
        self._alt_t = None
        self.release_sort_buffer = 40*Nmax > (1 << 30)
        self.size = Nmax

        self.charge = charge
        self.mass = mass
        self.n0 = n0
        self.order = order
        self.manifold = manifold

        # Location of hole left in particle arrays (particles.py:40)
        self.ihole = torch.zeros(ntmax, **i32)
        # Buffer arrays for the neighbour exchange (AoS rows of 5 doubles); one
        # header row in front of each send buffer carries the row count when the
        # exchange goes through NVLink peer memory (skeletor_b200/peer.py)
        self._sbl = torch.zeros((nbmax + 1, 5), **f64)
        self._sbr = torch.zeros((nbmax + 1, 5), **f64)
        self.sbufl = self._sbl[1:]
        self.sbufr = self._sbr[1:]
        from . import peer
        self._mig = peer.PeerArena(manifold.comm, (nbmax + 1)*5) \
            if peer.available(manifold.comm) else None
        if self._mig is None and getattr(manifold.comm, "size", 1) == 1 and \
                os.environ.get("SKELETOR_B200_PEER", "1") != "0":
            # one rank: the same device-counted migration through a local arena (only
            # the gapped push uses it, see _gap_finish_single_sync)
            self._mig = peer.LocalArena((nbmax + 1)*5, self.device)
        self.rbufl = torch.zeros((nbmax, 5), **f64)
        self.rbufr = torch.zeros((nbmax, 5), **f64)
        self._keep = torch.zeros((2*nbmax, 5), **f64)
        self._counts = torch.zeros(8, **i32)
        self._counts_b = torch.zeros(8, **i32)
        self._move_scratch = torch.zeros(2*ntmax + 8, **i32)
        self._ihole_scratch = torch.zeros(
            int(_lib.load().skb_ihole_scratch_ints(Nmax)) + 1, **i32)

        # Create source array
        self.sources = Sources(manifold)
        # Info array used for checking errors in particle move
        self.info = np.zeros(7, np.int32)
        # Set initial time
        self.time = time
        self._N_global = None
        self.N = 0

        # tile-sort state
        ntx, nty = C.c_int(), C.c_int()
        _lib.load().skb_tile_geometry(manifold.c, TLX, TLY, C.byref(ntx), C.byref(nty))
        self._ntx, self._nty = ntx.value, nty.value
        ntiles = self._ntx*self._nty
        self._cell_counts = torch.zeros((ntiles << (TLX + TLY)) + 1, **i32)
        self._cell_counts_alt = torch.zeros_like(self._cell_counts)
        self._block_sums = torch.zeros(4100, **i32)
        self._tile_offsets = torch.zeros(ntiles + 1, **i32)
        self._chunk_first = torch.zeros(Nmax//CHUNK + 2, **i32)
        self._n_sorted = 0
        self._sorted = False
        self.sort_enabled = True
        self.deterministic = False    # canonical intra-cell order after every sort
        self.fused_push = False   # two-pass recompute path: correct, but the push is
        # instruction-bound, so it is not faster than push + precounted sort
        # gapped layout (see _push_gapped): particles stay ordered without a full move
        # pass.  Needs Nmax >= ~1.3 N; push()/push_modified() fall back to the dense
        # path when the slot ranges do not fit.
        # On by default (SKELETOR_B200_GAPPED=0 turns it off): instances whose slot ranges
        # do not fit (Nmax < ~1.3 N) or that are converted back after every push (code
        # that reads the particles each step) stay on / return to the dense path.
        self.gapped = os.environ.get("SKELETOR_B200_GAPPED", "1") != "0"
        self.gapped_push_and_deposit = True   # (only with gapped) fused sweep on that layout
        # push + the full-step deposit that follows it in ONE sweep (skb_push_deposit_gapped):
        # the push also accumulates the stencil sums of the new positions into a private
        # grid and Sources.deposit(ions) picks that grid up when the particles have not
        # been touched in between.  "auto": switched on by the first deposit that follows a
        # push, switched off again when a fused result goes unused.  OFF by default: on
        # config 5 the fused sweep takes 50 ms against 23.7 + 7.5 ms for the two kernels
        # (register pressure of the two roles in one kernel).
        self.fuse_deposit = os.environ.get("SKELETOR_B200_FUSE", "0")
        if self.fuse_deposit in ("0", "1"):
            self.fuse_deposit = self.fuse_deposit == "1"
        self._fuse_next = False
        self._fused_grid = None
        self._fused_valid = False
        self._fused_used = False
        self._fused_S = 0.0
        self._last_op = None
        # overlap_migration (SKELETOR_B200_OVERLAP=1): push() returns once its kernel is
        # queued; the migration that completes it (count read-back, neighbour exchange,
        # classification: three host round trips) runs on a second stream and is only
        # waited for when something needs the particles - Sources.deposit(ions) first
        # deposits the cells the kernel left behind and lets the exchange overlap that
        # sweep.  Parity-checked on 8 GPUs (profiles/mgpu_check_r02_n8_*.log), but OFF by
        # default: on 8 B200s it is 5 % slower than the in-line path (4.84 vs 4.61 ms per
        # step; the rows inserted afterwards have to be deposited one by one).
        self.overlap_migration = os.environ.get("SKELETOR_B200_OVERLAP", "0") == "1"
        # single_sync (SKELETOR_B200_SINGLE_SYNC=0 turns it off): with the NVLink peer
        # arena the whole migration of a gapped push - send, classify, insert - is queued
        # behind the push kernel with device-side counts and the host reads all counters
        # in ONE synchronisation at the end (three before)
        self.single_sync = os.environ.get("SKELETOR_B200_SINGLE_SYNC", "1") != "0"
        self._pending = None
        self._side = None
        self._cnt_host = None
        self._gap_fail = None         # N at which the slot ranges did not fit
        self._gap_thrash = 0          # gapped -> dense conversions after a single push
        self._gap_pushes = 0
        self.mover_fraction = 0.1     # size of the global mover list relative to Nmax
        self._rep = "dense"
        self._gap_start = None
        self._gap_nleft = 0
        self._gap_dirty = False

    @property
    def _alt(self):
        if self._alt_t is None:
            shape = (5, self.size)
            try:
                self._alt_t = torch.empty(shape, dtype=torch.float64, device=self.device)
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()      # hand cached blocks back and try once more
                self._alt_t = torch.empty(shape, dtype=torch.float64, device=self.device)
        return self._alt_t

    @_alt.setter
    def _alt(self, t):
        self._alt_t = t

    # -- particle counts ------------------------------------------------------------
    @property
    def N(self):
        """number of particles in this slab"""
        self._finish_pending()
        return self._N

    @N.setter
    def N(self, n):
        if getattr(self, "_rep", "dense") == "gapped":
            self._dense()              # "the first N entries" only means something there
        self._fused_valid = False
        self._N = int(n)
        self._N_global = None          # unknown until the next reduction

    def _set_N_after_migration(self, n):
        # migration only moves particles between slabs: the global count is conserved
        self._N = int(n)

    def N_global(self):
        """sum of N over all ranks (the reference recomputes it with an allreduce in
        every normalize, sources.py:55-59; it only changes when N is assigned)"""
        if self._N_global is None:
            from .comm import SUM
            self._finish_pending()
            self._N_global = self.manifold.comm.allreduce(int(self._N), op=SUM)
        return self._N_global

    # -- NumPy-like surface -------------------------------------------------------
    @property
    def shape(self):
        return (self.size,)

    @property
    def dtype(self):
        return Particle

    def __len__(self):
        return self.size

    def _touched(self):
        self._sorted = False
        self._fused_valid = False
        self._last_op = None

    def __getitem__(self, key):
        self._dense()
        if isinstance(key, str):
            return ParticleRow(self, _ROW[key])
        if isinstance(key, slice):
            return ParticleSlice(self, key)
        return np.asarray(ParticleSlice(self, slice(None)))[key]

    def __setitem__(self, key, val):
        if isinstance(key, str):
            self[key][...] = val
        elif isinstance(key, slice):
            a = np.asarray(val)
            for k in _ROW:
                self[k][key] = a[k]
        else:
            raise TypeError("unsupported particle assignment")

    def __array__(self, dtype=None, copy=None):
        self._dense()
        return np.asarray(ParticleSlice(self, slice(None)))

    # -- C ABI views ---------------------------------------------------------------
    @staticmethod
    def _soa(t):
        n = t.shape[1]*8
        p = t.data_ptr()
        return _lib.ParticlesT(p, p + n, p + 2*n, p + 3*n, p + 4*n)

    @property
    def _c(self):
        return self._soa(self._data)

    def _tiling_c(self):
        if self._rep == "gapped":
            # cell k = slots [gap_start[k], gap_start[k] + gap_count[k])
            return C.pointer(_lib.TilingT(
                None, None, None, self._ntx, self._nty, TLX, TLY, CHUNK, 0,
                self._gap_start.data_ptr(), self._gap_count.data_ptr()))
        if not (self._sorted and self._n_sorted > 0):
            return None
        return C.pointer(_lib.TilingT(
            self._tile_offsets.data_ptr(), self._chunk_first.data_ptr(),
            self._cell_counts.data_ptr(),
            self._ntx, self._nty, TLX, TLY, CHUNK, self._n_sorted))

    def _epilogue(self, flags, S=0.0):
        return C.pointer(_lib.EpilogueT(flags, float(S), float(self.time),
                                        self.ihole.data_ptr(), self.ntmax - 1,
                                        self._cell_counts.data_ptr(), self.order,
                                        TLX, TLY))

    # -- tile sort -------------------------------------------------------------------
    def sort(self, precounted=False):
        """Counting sort by tile-major stencil-base cell (skb_tile_sort).
        precounted: the key histogram is already in place (fused into the push
        epilogue + arrivals), so the 16 B/particle key pass is skipped."""
        self._dense()
        if self.N == 0 or not self.sort_enabled:
            self._sorted = False
            return
        if precounted:
            _lib.call("skb_tile_sort_precounted", self._c, self._soa(self._alt), self.N,
                      self.manifold.c, self.order, TLX, TLY, CHUNK,
                      self._cell_counts.data_ptr(), self._block_sums.data_ptr(),
                      self._tile_offsets.data_ptr(), self._chunk_first.data_ptr(),
                      _stream())
        else:
            _lib.call("skb_tile_sort", self._c, self._soa(self._alt), self.N,
                      self.manifold.c, self.order, TLX, TLY, CHUNK,
                      self._cell_counts.data_ptr(), self._block_sums.data_ptr(),
                      self._tile_offsets.data_ptr(), self._chunk_first.data_ptr(), 0,
                      None, _stream())
        self._data, self._alt = self._alt, self._data
        if self.deterministic:
            # claim order inside a cell depends on warp scheduling: rewrite every cell
            # in lexicographic order so the stored array (and every deposit sum that
            # follows) is bitwise reproducible
            _lib.call("skb_canonical_cells", self._c, self._soa(self._alt),
                      self._cell_counts.data_ptr(), self.manifold.c, TLX, TLY, _stream())
            self._data, self._alt = self._alt, self._data
        self._n_sorted = self.N
        self._sorted = True

    def _cellsums(self):
        """scratch of the deterministic deposit: [ncells][(order+1)^2 * 4] doubles"""
        if getattr(self, "_cellsums_buf", None) is None:
            nv = (self.order + 1)**2*4
            self._cellsums_buf = torch.zeros((self._cell_counts.numel() - 1)*nv,
                                             dtype=torch.float64, device=self.device)
        return self._cellsums_buf

    def _ensure_sorted(self):
        if self._rep == "gapped":
            return                  # ordered by construction
        if not self._sorted:
            self.sort()

    # -- gapped layout ---------------------------------------------------------------
    def _gap_alloc(self):
        if self._gap_start is None:
            i32 = dict(dtype=torch.int32, device=self.device)
            f64 = dict(dtype=torch.float64, device=self.device)
            nc = self._cell_counts.numel()                    # ncells + 1
            self._gap_start = torch.zeros(nc, **i32)
            self._gap_count = torch.zeros(nc, **i32)
            # every warp of skb_push_gapped reserves mover rows 64 at a time
            ntiles = self._ntx*self._nty
            warps = 8*max(ntiles, min(2368, 32*ntiles))
            mcap = max(int(self.mover_fraction*self.size), 1024) + 64*warps
            self._movers = torch.zeros((mcap, 5), **f64)
            # particles whose cell ran out of slots: small unordered SoA list beside
            # the cells, pushed / deposited by the generic kernels
            self._leftover = torch.zeros((5, max(mcap//8, 1024)), **f64)
            self._gcnt = torch.zeros(8, **i32)
            # scratch blocks for the movers a thread block re-inserts itself: sized for
            # ~3x the expected in-block movers of a uniform plasma
            ncta = max(ntiles, min(2368, 32*ntiles))
            self._scr_rows = int(min(max(1024, 0.5*self.size/ncta), 1 << 16)) & ~1
            self._npool = 1024
            self._scratch = torch.zeros((self._npool, self._scr_rows, 5), **f64)
            self._pool_owner = torch.zeros(self._npool, **i32)

    def _to_gapped(self):
        """dense ordered arrays -> per-cell slot ranges with slack (skb_gap_build).
        Returns False (and stays dense) when the ranges do not fit in Nmax slots."""
        if self.N == 0 or not self.sort_enabled or self.deterministic:
            return False
        if self.size % 2:
            return False        # rows of the [5][Nmax] tensor must be 16-byte aligned
        if self._gap_fail is not None and \
                abs(self.N - self._gap_fail) <= 0.02*self._gap_fail:
            return False        # did not fit last time and N has hardly changed
        if self._gap_thrash >= 16:
            # the caller reads the particles after every push: converting back and forth
            # costs more than the tile sort; try again every 64 pushes
            self._gap_thrash += 1
            if self._gap_thrash < 80:
                return False
            self._gap_thrash = 0
        if not (self._sorted and self._n_sorted == self.N):
            self.sort()
        self._gap_alloc()
        _lib.call("skb_gap_build", self._c, self._soa(self._alt),
                  self._cell_counts.data_ptr(), self.manifold.c, TLX, TLY,
                  self._gap_start.data_ptr(), self._gap_count.data_ptr(),
                  self._block_sums.data_ptr(), self.size, _stream())
        total = int(self._gap_start[-1].item())
        if total > self.size or total < 0:
            if self._gap_fail is None and self.size >= 1.25*self.N:
                warn("gapped particle layout needs Nmax >= ~1.36 N at this occupation "
                     "(N={}, Nmax={}): staying on the dense layout".format(self.N, self.size))
            self._gap_fail = self.N
            return False
        self._gap_fail = None
        self._data, self._alt = self._alt, self._data
        if self.release_sort_buffer:
            self._alt_t = None          # (the dense copy: back to the caching allocator)
        self._rep = "gapped"
        self._gap_nleft = 0
        self._gap_dirty = False
        self._gap_pushes = 0
        return True

    def _dense(self):
        """back to the dense representation every other method works on"""
        self._finish_pending()
        if self._rep != "gapped":
            return
        self._gap_thrash = self._gap_thrash + 1 if self._gap_pushes <= 1 else 0
        nleft = self._gap_nleft
        st = _stream()
        _lib.call("skb_gap_densify", self._c, self._soa(self._alt),
                  self._gap_start.data_ptr(), self._gap_count.data_ptr(),
                  self.manifold.c, TLX, TLY, self._cell_counts_alt.data_ptr(),
                  self._cell_counts.data_ptr(), self._tile_offsets.data_ptr(),
                  self._block_sums.data_ptr(), self._leftover.data_ptr(),
                  self._leftover.shape[1], nleft, self.N - nleft, st)
        _lib.call("skb_chunk_table", self._tile_offsets.data_ptr(), self.manifold.c,
                  TLX, TLY, CHUNK, self._chunk_first.data_ptr(), st)
        self._data, self._alt = self._alt, self._data
        self._rep = "dense"
        self._n_sorted = self.N - nleft
        # particles parked in a wrong cell (mover list overflow) break the ordering
        self._sorted = not self._gap_dirty and self._n_sorted > 0
        self._gap_nleft = 0
        self._gap_dirty = False

    def _push_gapped(self, E, B, dt, modified):
        """push + boundaries + migration on the gapped layout: every cell owns a slot
        range with slack, so only the particles that change cell are relocated; the
        rest are rewritten in place.  Movers that stay inside the cell range of their
        thread block are re-inserted by the block itself (L2-resident scratch), the
        others and the arrivals go through an AoS list and skb_gap_insert.  HBM traffic
        per particle and step drops from 176 B (push + move pass of the tile sort) to
        ~100 B.  Cells that run out of slots put their surplus on a small leftover list
        that is pushed and deposited by the generic kernels and re-inserted every step;
        when that list grows past half its size the slot ranges are rebuilt (densify ->
        tile sort -> skb_gap_build)."""
        if self._rep != "gapped" and not self._to_gapped():
            return False
        fuse = self.fuse_deposit is True or \
            (self.fuse_deposit == "auto" and self._fuse_next)
        cnt = self._gap_kernel(E, B, dt, modified, fuse)
        if self.overlap_migration and not fuse:
            self._start_pending(cnt)
        else:
            self._gap_finish(cnt, fused=fuse)
        self._gap_pushes += 1
        return True

    # -- migration overlapped with the deposit (multi-rank) ------------------------------
    def _start_pending(self, cnt):
        """queue the read-back of the push kernel's counters on the side stream"""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._cnt_host = torch.zeros(8, dtype=torch.int32).pin_memory()
            self._cnt_ev = torch.cuda.Event()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._cnt_host[:5].copy_(cnt[:5], non_blocking=True)
            self._cnt_host[5:7].copy_(self._gcnt[:2], non_blocking=True)
            self._cnt_ev.record(self._side)
        self._pending = cnt

    def _finish_pending(self, deposit_into=None):
        """complete the migration of the last push (no-op when nothing is pending).
        deposit_into: a Sources whose grid already holds the deposit of the cells as the
        push kernel left them; the rows that are inserted now are added to it."""
        if getattr(self, "_pending", None) is None:
            return
        self._pending = None
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._side):
            self._cnt_ev.synchronize()
            nm, nl, nr, fl, nlocal, nleft_b, lost_b = self._cnt_host[:7].tolist()
            self._gap_check_flags(fl)
            nkeep = self._exchange(nl, nr)       # on the side stream: overlaps the deposit
            done = torch.cuda.Event()
            done.record(self._side)
        main.wait_event(done)
        self._gap_finish_tail(nm, nl, nr, fl, nlocal, nkeep, False, deposit_into, nleft_b)

    def _fused_sources(self):
        """private grid of the fused push + deposit sweep (raw sums, [myp][mx][4])"""
        if self._fused_grid is None:
            m = self.manifold
            self._fused_grid = torch.zeros((m.myp, m.mx, 4), dtype=torch.float64,
                                           device=self.device)
        return self._fused_grid

    def _fused_for(self, sources, S):
        """the grid Sources.deposit(self) may copy instead of running the deposit kernel:
        the last push accumulated it and nothing has touched the particles since"""
        if self._fused_valid and sources.grid is self.manifold and \
                float(S) == self._fused_S and not self.deterministic:
            self._fused_used = True
            return self._fused_grid
        return None

    def _gap_kernel(self, E, B, dt, modified, fuse=False):
        m = self.manifold
        comm = m.comm
        self.time += dt
        args, flags = self._push_args(dt, modified)
        cnt = self._counts
        extra = ()
        if fuse:
            fs = self._fused_sources()
            fs.zero_()
            self._fused_S = float(getattr(m, 'S', 0.0))
            extra = (fs.data_ptr(), self._fused_S)
        _lib.call("skb_push_deposit_gapped" if fuse else "skb_push_gapped",
                  self._c, E.ptr, B.ptr, *args, flags,
                  float(getattr(m, 'S', 0.0)), float(self.time), *extra, TLX, TLY,
                  self._gap_start.data_ptr(), self._gap_count.data_ptr(),
                  self._movers.data_ptr(), self._movers.shape[0],
                  self.sbufl.data_ptr(), self.sbufr.data_ptr(), self.nbmax,
                  cnt.data_ptr(), comm.rank, comm.size, self._leftover.data_ptr(),
                  self._leftover.shape[1], self._gap_nleft, self._gcnt.data_ptr(),
                  self._scratch.data_ptr(), self._scr_rows, self._npool,
                  self._pool_owner.data_ptr(), _stream())
        return cnt

    def _push_and_deposit_gapped(self, E, B, dt, update):
        """push_and_deposit on the gapped layout (skb_push_and_deposit_gapped): the half-step
        deposit is fused into the gapped push; update=False only deposits."""
        m = self.manifold
        comm = m.comm
        self.time += dt
        qtmh = self.charge/self.mass*dt/2
        src = self.sources
        src.t.zero_()
        cnt = self._counts
        had_leftovers = self._gap_nleft > 0
        _lib.call("skb_push_and_deposit_gapped", self._c, E.ptr, B.ptr, m.c, self.order,
                  float(qtmh), float(dt), src.ptr, 0.0, int(bool(update)),
                  self.ihole.data_ptr(), self.ntmax - 1, TLX, TLY,
                  self._gap_start.data_ptr(), self._gap_count.data_ptr(),
                  self._movers.data_ptr(), self._movers.shape[0],
                  self.sbufl.data_ptr(), self.sbufr.data_ptr(), self.nbmax,
                  cnt.data_ptr(), comm.rank, comm.size, self._leftover.data_ptr(),
                  self._leftover.shape[1], self._gap_nleft, self._gcnt.data_ptr(),
                  self._scratch.data_ptr(), self._scr_rows, self._npool,
                  self._pool_owner.data_ptr(), _stream())
        src.boundaries_set = False
        src.normalize(self)
        src.set_boundaries()
        self._gap_pushes += 1
        cfl = had_leftovers and int(self.ihole[0].item()) < 0
        if update:
            self._gap_finish(cnt, cfl)
        elif cfl or (int(cnt[3].item()) & 4):
            # moved more than half a cell in half a step (push_and_deposit.pyx:66-68): the
            # reference's ihole[0] = -1, reported by particles.py:113-117
            msg = "ihole overflow error: ntmax={}, ierr={}"
            raise RuntimeError(msg.format(self.ihole.numel() - 1, 1))

    def _gap_check_flags(self, fl, cfl=False):
        if cfl or (fl & 4):
            msg = "ihole overflow error: ntmax={}, ierr={}"
            raise RuntimeError(msg.format(self.ihole.numel() - 1, 1))
        if fl & 2:
            raise RuntimeError("particle buffer overflow: nbmax={}".format(self.nbmax))
        if fl & 8:
            raise RuntimeError("gapped layout: mover lists overflowed, particles were "
                               "lost (raise Particles.mover_fraction; {} rows)".format(
                                   self._movers.shape[0]))

    def _gap_finish(self, cnt, cfl=False, fused=False):
        if self._mig is not None and self.single_sync and not fused:
            return self._gap_finish_single_sync(cnt, cfl)
        nm, nl, nr, fl, nlocal = cnt[:5].tolist()
        self._gap_check_flags(fl, cfl)
        nkeep = self._exchange(nl, nr)
        self._gap_finish_tail(nm, nl, nr, fl, nlocal, nkeep, fused)

    def _gap_finish_single_sync(self, cnt, cfl=False):
        """_gap_finish through NVLink peer memory with ONE host synchronisation: the
        leavers are sent with device-side counts (skb_peer_send), classified from the
        message headers, movers and arrivals are inserted with device-side counts
        (skb_gap_insert_counted), and only then the host reads every counter of the step.
        Everything is queued while the push kernel runs, so the migration costs the GPU
        its kernels and one barrier, not the host's launch latencies."""
        m = self.manifold
        comm = m.comm
        st = _stream()
        arena = self._mig
        c2 = self._counts_b
        p0 = cnt.data_ptr()
        # cnt = (rows on the mover list, leavers down, leavers up, flags, ...)
        from_below, from_above = arena.exchange_counted(self.sbufr, p0 + 8, self.sbufl, p0 + 4,
                                                        self.nbmax)
        c2.zero_()
        for buf in (from_below, from_above):
            _lib.call("skb_move_classify", buf.data_ptr(), -(self.nbmax + 1),
                      self._keep.data_ptr(), self.sbufl.data_ptr(),
                      self.sbufr.data_ptr(), self.nbmax, c2.data_ptr(), m.c,
                      comm.rank, comm.size, st)
        flags = arena.all_flags((c2[1] + c2[2]) > 0)
        g = self._gcnt          # (reset by skb_push_gapped)
        for rows, nptr in ((self._movers, p0), (self._keep, c2.data_ptr())):
            _lib.call("skb_gap_insert_counted", rows.data_ptr(), nptr, rows.shape[0],
                      self._c, self._gap_start.data_ptr(), self._gap_count.data_ptr(), m.c,
                      self.order, TLX, TLY, self._leftover.data_ptr(),
                      self._leftover.shape[1], g.data_ptr(), st)
        vals = torch.cat([cnt[:5], c2[:4], g[:2], flags.to(torch.int32)]).tolist()
        nm, nl, nr, fl, nlocal, nkeep, fnl, fnr, ovf, nleft, lost = vals[:11]
        self._gap_check_flags(fl, cfl)
        if ovf or nkeep > self._keep.shape[0]:
            raise RuntimeError("particle buffer overflow while forwarding")
        self.info[4] = max(self.info[4], 1)
        if any(vals[11:]):
            # somebody still forwards particles (more than one slab crossed in a step):
            # the remaining rounds of cppmove2 (pplib2.c:708-866), then their arrivals
            nkeep2 = self._exchange_peer(fnl, fnr, nkeep=nkeep, first=1)
            _lib.call("skb_gap_insert", self._keep[nkeep:].data_ptr(), nkeep2 - nkeep,
                      self._c, self._gap_start.data_ptr(), self._gap_count.data_ptr(), m.c,
                      self.order, TLX, TLY, self._leftover.data_ptr(),
                      self._leftover.shape[1], g.data_ptr(), st)
            nleft, lost = g[:2].tolist()
            nkeep = nkeep2
        self._gap_finish_tail(nm, nl, nr, fl, nlocal, nkeep, False, inserted=(nleft, lost))

    def _gap_finish_tail(self, nm, nl, nr, fl, nlocal, nkeep, fused, deposit_into=None,
                         nleft_b=0, inserted=None):
        """inserted: (leftover rows, overflow flag) when movers and arrivals have been
        inserted already (_gap_finish_single_sync)"""
        m = self.manifold
        st = _stream()
        if deposit_into is not None:
            # the cells were deposited as the push kernel left them: add what is inserted
            # now (rows on the global mover list, arrivals) and the rows that kernel put
            # on the leftover list itself
            S = float(getattr(deposit_into.grid, 'S', 0.0))
            for rows, n in ((self._movers, min(nm, self._movers.shape[0])),
                            (self._keep, nkeep)):
                _lib.call("skb_deposit_rows", rows.data_ptr(), n, deposit_into.ptr, m.c,
                          self.order, S, st)
            if nleft_b > 0:
                _lib.call("skb_deposit", self._soa(self._leftover),
                          min(nleft_b, self._leftover.shape[1]), deposit_into.ptr, m.c,
                          self.order, S, None, st)
        if fused:
            # the arrivals were not in this slab when the push accumulated its grid
            _lib.call("skb_deposit_rows", self._keep.data_ptr(), nkeep,
                      self._fused_grid.data_ptr(), m.c, self.order, self._fused_S, st)
        self._fused_valid = fused and not (fl & 16)
        self._fused_used = False
        new_n = self._N - nl - nr + nkeep
        if new_n > self.size:
            self.info[0] = new_n - self.size
            raise RuntimeError("particle overflow error, ierr = {}".format(
                new_n - self.size))
        g = self._gcnt          # (reset by skb_push_gapped)
        if inserted is None:
            for rows, n in ((self._movers, min(nm, self._movers.shape[0])),
                            (self._keep, nkeep)):
                _lib.call("skb_gap_insert", rows.data_ptr(), n, self._c,
                          self._gap_start.data_ptr(), self._gap_count.data_ptr(), m.c,
                          self.order, TLX, TLY, self._leftover.data_ptr(),
                          self._leftover.shape[1], g.data_ptr(), st)
            nleft, lost = g[:2].tolist()
        else:
            nleft, lost = inserted
        if lost:
            raise RuntimeError("gapped layout: leftover list overflow "
                               "({} > {})".format(nleft, self._leftover.shape[1]))
        self._set_N_after_migration(new_n)
        self.info[1] = self.info[2] = new_n
        self._gap_nleft = nleft
        self._gap_dirty = bool(fl & 1)
        # rows on the global list (incl. padding), leavers, flags, arrivals, leftovers,
        # movers re-inserted in place
        self._gap_stats = (nm, nl, nr, fl, nkeep, nleft, nlocal)
        if self._gap_dirty or 2*nleft > min(self._leftover.shape[1],
                                            self._movers.shape[0]):
            # slack exhausted in many cells (or parked particles): fall back to dense;
            # the next step re-sorts and rebuilds the slot ranges around the current
            # occupation
            self._dense()

    # -- reference API ---------------------------------------------------------------
    def initialize(self, x, y, vx, vy, vz):
        """particles.py:77-102 (positions in physical units, host arrays)"""
        m = self.manifold
        self._rep = "dense"
        x, y, vx, vy, vz = (np.asarray(a, dtype=np.float64) for a in (x, y, vx, vy, vz))
        ind = np.logical_and(y >= m.y0 + m.edges[0]*m.dy,
                             y < m.y0 + m.edges[1]*m.dy)
        self.N = int(np.sum(ind))
        assert self.size >= self.N
        if self.size < int(5/4*self.N):
            msg = "Particle array is probably not large enough"
            warn(msg + " (N={}, Nmax={})".format(self.N, self.size))
        host = np.stack([(x[ind] - m.x0)/m.dx, (y[ind] - m.y0)/m.dy,
                         vx[ind], vy[ind], vz[ind]])
        self._data[:, :self.N] = torch.as_tensor(host, device=self.device)
        self._sorted = False
        self._gap_fail = None

    def deposit(self, **kwds):
        self.sources.deposit(self, **kwds)

    def _ihole_count(self):
        n = int(self.ihole[0].item())
        # Check for ihole overflow error (particles.py:113-117)
        if n < 0:
            msg = "ihole overflow error: ntmax={}, ierr={}"
            raise RuntimeError(msg.format(self.ihole.numel() - 1, -n))
        return n

    def move(self):
        """Move particles that left the slab to the neighbouring ranks: ppic2's
        cppmove2 (pplib2.c:607-981) as pack -> NCCL ring exchange -> unpack."""
        self._dense()
        self._fused_valid = False
        g = self.manifold
        comm = g.comm
        gc = g.c
        st = _stream()
        cnt = self._counts
        # the hole count stays on the device for the pack; ONE D2H read afterwards
        # returns it together with the two buffer counts and the overflow flag
        _lib.call("skb_move_pack", self._c, self.ihole.data_ptr(), -self.ntmax,
                  self.sbufl.data_ptr(), self.sbufr.data_ptr(), self.nbmax,
                  cnt.data_ptr(), gc, comm.rank, comm.size, st)
        nl, nr, ovf, nh = cnt[:4].tolist()
        # Check for ihole overflow error (particles.py:113-117)
        if nh < 0:
            msg = "ihole overflow error: ntmax={}, ierr={}"
            raise RuntimeError(msg.format(self.ihole.numel() - 1, -nh))
        if ovf:
            raise RuntimeError("particle buffer overflow: nbmax={}".format(self.nbmax))
        nkeep = self._exchange(nl, nr)
        new_n = self.N + nkeep - nh
        if new_n > self.size:
            self.info[0] = new_n - self.size
            raise RuntimeError("particle overflow error, ierr = {}".format(
                new_n - self.size))
        _lib.call("skb_move_unpack", self._c, self.N, self.ihole.data_ptr(), nh,
                  self._keep.data_ptr(), nkeep, self._move_scratch.data_ptr(), st)
        self._set_N_after_migration(new_n)
        self.info[1] = self.info[2] = new_n
        self._sorted = False
        return nkeep

    def _exchange(self, nl, nr):
        """Neighbour exchange of the packed leavers (sbufl: nl rows going down, sbufr:
        nr rows going up) incl. multi-hop forwarding (pplib2.c:708-866).  Returns the
        number of arrivals, left as AoS rows in self._keep."""
        g = self.manifold
        comm = g.comm
        gc = g.c
        st = _stream()
        cnt = self._counts
        nkeep = 0
        if self._mig is not None and not self._mig.local:
            return self._exchange_peer(nl, nr)
        for it in range(2000):
            self.info[4] = max(self.info[4], it + 1)     # passes, pplib2.c:955
            if comm.size == 1:
                # nvp == 1: rbufl = sbufr, rbufr = sbufl (pplib2.c:715-730)
                from_below, from_above = self.sbufr[:nr], self.sbufl[:nl]
                nb_, na_ = nr, nl
            else:
                nb_, na_ = comm.exchange_counts(nr, nl)
                # (at least one row per message: no zero-byte NCCL transfers)
                from_below, from_above = comm.ring_exchange(
                    self.sbufr[:max(nr, 1)], self.sbufl[:max(nl, 1)],
                    self.rbufl[:max(nb_, 1)], self.rbufr[:max(na_, 1)])
            # keep what belongs here, pass the rest on (pplib2.c:756-866)
            cnt.zero_()
            cnt[0] = nkeep
            if comm.size == 1:
                # sbufl/sbufr are both source and destination: classify from copies
                from_below, from_above = from_below.clone(), from_above.clone()
            for buf, n in ((from_below, nb_), (from_above, na_)):
                _lib.call("skb_move_classify", buf.data_ptr(), n,
                          self._keep.data_ptr(), self.sbufl.data_ptr(),
                          self.sbufr.data_ptr(), self.nbmax, cnt.data_ptr(), gc,
                          comm.rank, comm.size, st)
            nkeep, nl, nr, ovf = cnt[:4].tolist()
            if ovf or nkeep > self._keep.shape[0]:
                raise RuntimeError("particle buffer overflow while forwarding")
            more = nl + nr
            if comm.size > 1:
                from .comm import MAX
                more = comm.allreduce(more, op=MAX)
            if more == 0:
                break
        return nkeep

    def _exchange_peer(self, nl, nr, nkeep=0, first=0):
        """_exchange through NVLink peer memory: header + rows are copied straight into
        the neighbours' slots, one device-side barrier, classification with device-side
        counts, and the "does anybody still forward?" agreement (cppimax, pplib2.c:873)
        through per-rank flag words — no NCCL call and ONE host sync per round."""
        g = self.manifold
        comm = g.comm
        gc = g.c
        st = _stream()
        cnt = self._counts
        arena = self._mig
        for it in range(first, 2000):
            self.info[4] = max(self.info[4], it + 1)     # passes, pplib2.c:955
            self._sbl[0, :1].fill_(float(nl))
            self._sbr[0, :1].fill_(float(nr))
            from_below, from_above = arena.exchange(self._sbr[:nr + 1], self._sbl[:nl + 1])
            cnt.zero_()
            cnt[:1].fill_(nkeep)
            for buf in (from_below, from_above):
                _lib.call("skb_move_classify", buf.data_ptr(), -(self.nbmax + 1),
                          self._keep.data_ptr(), self.sbufl.data_ptr(),
                          self.sbufr.data_ptr(), self.nbmax, cnt.data_ptr(), gc,
                          comm.rank, comm.size, st)
            flags = arena.all_flags((cnt[1] + cnt[2]) > 0)
            vals = torch.cat([cnt[:4].to(torch.float64), flags]).tolist()
            nkeep, nl, nr, ovf = (int(v) for v in vals[:4])
            if ovf or nkeep > self._keep.shape[0]:
                raise RuntimeError("particle buffer overflow while forwarding")
            if not any(vals[4:]):
                break
        return nkeep

    def periodic_x(self):
        """Applies periodic boundaries on particles along x"""
        self._dense()
        self._fused_valid = False
        self._last_op = None
        _lib.call("skb_periodic_x", self._c, self.N, self.manifold.c, _stream())

    def calculate_ihole(self):
        self._dense()
        _lib.call("skb_calculate_ihole", self._c, self.N, self.ihole.data_ptr(),
                  self.ntmax - 1, self.manifold.c, self._ihole_scratch.data_ptr(),
                  _stream())

    def periodic_y(self):
        """Applies periodic boundaries on particles along y: calculates ihole and
        then moves particles between processors (particles.py:133-143)."""
        self.calculate_ihole()
        self.move()

    def shear_periodic_y(self):
        """Shearing periodic boundaries along y (particles.py:145-157)."""
        self._dense()
        self._fused_valid = False
        self._last_op = None
        _lib.call("skb_shear_periodic_y", self._c, self.N, self.manifold.c,
                  float(self.manifold.S), float(self.time), _stream())
        self.periodic_y()

    def _push(self, E, B, dt, modified):
        self._finish_pending()
        if self._fused_valid and not self._fused_used:
            self._fuse_next = False        # the last fused deposit was not picked up
        self._fused_valid = False
        self._last_op = "push"
        if self.gapped and self.order in (1, 2) and \
                self._push_gapped(E, B, dt, modified):
            return
        self._dense()
        if self.sort_enabled and self.fused_push and self.N > 0:
            self._push_fused(E, B, dt, modified)
        else:
            count = self.sort_enabled and self.N > 0
            self._push_kernel(E, B, dt, modified, count=count)
            nkeep = self.move()
            if count:
                # arrivals complete the histogram the push epilogue started
                _lib.call("skb_sort_count_rows", self._keep.data_ptr(), nkeep,
                          self.manifold.c, self.order, TLX, TLY,
                          self._cell_counts.data_ptr(), _stream())
            self.sort(precounted=count)

    def _push_args(self, dt, modified):
        m = self.manifold
        qtmh = self.charge/self.mass*dt/2
        shear = hasattr(m, 'S')
        flags = _lib.EPI_PERIODIC_X | (_lib.EPI_SHEAR if shear else 0)
        return (m.c, self.order, float(qtmh), float(dt), int(modified),
                float(getattr(m, 'Omega', 0.0)) if modified else 0.0,
                float(getattr(m, 'S', 0.0)) if modified else 0.0), flags

    def _push_fused(self, E, B, dt, modified):
        """push + boundary epilogue + migration + tile sort as two passes that both
        recompute the push (skb_push_count / skb_push_scatter): 120 B of HBM traffic
        per particle instead of 176 B, no hole list, no hole filling."""
        if self.order not in (1, 2):
            msg = 'Interpolation order {} not implemented.'
            raise RuntimeError(msg.format(self.order))
        m = self.manifold
        comm = m.comm
        self.time += dt
        args, flags = self._push_args(dt, modified)
        self._ensure_sorted()
        st = _stream()
        til = self._tiling_c()
        epi = self._epilogue(flags, getattr(m, 'S', 0.0))
        cells = self._cell_counts.data_ptr()
        cnt = self._counts
        # pass 1: histogram of new cells; leavers go straight into sbufl / sbufr
        _lib.call("skb_push_count", self._c, self.N, E.ptr, B.ptr, *args, til, epi,
                  TLX, TLY, cells, self.sbufl.data_ptr(), self.sbufr.data_ptr(),
                  self.nbmax, cnt.data_ptr(), comm.rank, comm.size, st)
        nl, nr, ovf = cnt[:3].tolist()
        if ovf:
            raise RuntimeError("particle buffer overflow: nbmax={}".format(self.nbmax))
        nkeep = self._exchange(nl, nr)
        new_n = self.N - nl - nr + nkeep
        if new_n > self.size:
            self.info[0] = new_n - self.size
            raise RuntimeError("particle overflow error, ierr = {}".format(
                new_n - self.size))
        gc = args[0]
        _lib.call("skb_sort_count_rows", self._keep.data_ptr(), nkeep, gc, self.order,
                  TLX, TLY, cells, st)
        _lib.call("skb_sort_scan", cells, gc, TLX, TLY, CHUNK,
                  self._block_sums.data_ptr(), self._tile_offsets.data_ptr(),
                  self._chunk_first.data_ptr(), st)
        # pass 2: push again, write every staying particle to its sorted slot
        out = self._soa(self._alt)
        _lib.call("skb_push_scatter", self._c, out, self.N, E.ptr, B.ptr, *args, til, epi,
                  TLX, TLY, cells, st)
        _lib.call("skb_sort_scatter_rows", self._keep.data_ptr(), nkeep, out, gc,
                  self.order, TLX, TLY, cells, st)
        self._data, self._alt = self._alt, self._data
        self._set_N_after_migration(new_n)
        self.info[1] = self.info[2] = new_n
        self._n_sorted = new_n
        self._sorted = True

    def _push_kernel(self, E, B, dt, modified, count=False):
        if self.order not in (1, 2):
            msg = 'Interpolation order {} not implemented.'
            raise RuntimeError(msg.format(self.order))
        m = self.manifold
        self._dense()
        # Update time
        self.time += dt
        qtmh = self.charge/self.mass*dt/2
        shear = hasattr(m, 'S')
        # kernel + fused boundary epilogue: shear boost, hole list, x wrap
        # (particles.py:179-188 in one pass over the particles)
        flags = _lib.EPI_HOLES | _lib.EPI_PERIODIC_X | (_lib.EPI_SHEAR if shear else 0)
        self._ensure_sorted()
        if count:
            # the tiling read by this launch lives in tile_offsets / chunk_first; the
            # per-cell histogram array is free to be rebuilt for the NEXT ordering
            flags |= _lib.EPI_COUNT
            _lib.call("skb_sort_clear", self._cell_counts.data_ptr(), m.c, TLX, TLY,
                      _stream())
        _lib.call("skb_boris_push", self._c, self.N, E.ptr, B.ptr, m.c, self.order,
                  float(qtmh), float(dt), int(modified),
                  float(getattr(m, 'Omega', 0.0)) if modified else 0.0,
                  float(getattr(m, 'S', 0.0)) if modified else 0.0,
                  self._tiling_c(), self._epilogue(flags, getattr(m, 'S', 0.0)),
                  _stream())

    def kick(self, E, B, dt):
        """Velocity update alone: gather E, B at the particles and apply the Boris kick
        of `push` (kick_particle, particle_push.pxd:69-86) over dt; positions, time and
        slab membership do not change.  (The reference has no public kick - only the
        inlined cdef reachable through push - this is push without drift and boundary
        conditions: the same kernel with a zero drift factor.)"""
        if self.order not in (1, 2):
            msg = 'Interpolation order {} not implemented.'
            raise RuntimeError(msg.format(self.order))
        self._dense()
        self._fused_valid = False
        self._last_op = None
        qtmh = self.charge/self.mass*dt/2
        self._ensure_sorted()
        _lib.call("skb_boris_push", self._c, self.N, E.ptr, B.ptr, self.manifold.c,
                  self.order, float(qtmh), 0.0, 0, 0.0, 0.0, self._tiling_c(),
                  self._epilogue(0), _stream())

    def push(self, E, B, dt):
        """A standard Boris push which updates positions and velocities
        (particles.py:159-188).  If shear is turned on, E needs to be E_star and B
        needs to be B_star."""
        self._push(E, B, dt, False)

    def push_modified(self, E, B, dt):
        """particles.py:233-257"""
        self._push(E, B, dt, True)

    def push_and_deposit(self, E, B, dt, update=True):
        """Updates positions and velocities and deposits charge and currents at the
        half step; update=False only computes the new sources (predictor step).
        particles.py:190-231.  Does not work with shear (S = 0, as the reference)."""
        if self.order not in (1, 2):
            msg = 'Interpolation order {} not implemented.'
            raise RuntimeError(msg.format(self.order))
        self._finish_pending()
        self._fused_valid = False
        self._last_op = None
        if self.gapped and self.gapped_push_and_deposit and \
                (self._rep == "gapped" or self._to_gapped()) and \
                self._gap_nleft < self.ntmax - 1:
            return self._push_and_deposit_gapped(E, B, dt, update)
        self._dense()
        self.time += dt
        qtmh = self.charge/self.mass*dt/2
        S = 0.0
        src = self.sources
        src.t.zero_()
        self._ensure_sorted()
        # histogram of the new cell keys for the tile sort that follows an update, built
        # by the same kernel into the spare cell array (the live one is being read as
        # cell_end)
        count = bool(update) and self.sort_enabled and self._sorted and \
            self._n_sorted == self.N and self.N > 0
        nxt = None
        if count:
            nxt = self._cell_counts_alt.data_ptr()
            _lib.call("skb_sort_clear", nxt, self.manifold.c, TLX, TLY, _stream())
        _lib.call("skb_push_and_deposit", self._c, self.N, E.ptr, B.ptr,
                  self.manifold.c, self.order, float(qtmh), float(dt),
                  self.ihole.data_ptr(), self.ntmax - 1, src.ptr, S, int(bool(update)),
                  self._tiling_c(), nxt, TLX, TLY, _stream())
        src.boundaries_set = False
        src.normalize(self)
        src.set_boundaries()
        if update:
            nkeep = self.move()
            if count:
                self._cell_counts, self._cell_counts_alt = \
                    self._cell_counts_alt, self._cell_counts
                _lib.call("skb_sort_count_rows", self._keep.data_ptr(), nkeep,
                          self.manifold.c, self.order, TLX, TLY,
                          self._cell_counts.data_ptr(), _stream())
            self.sort(precounted=count)
        elif int(self.ihole[0].item()) < 0:
            self._ihole_count()

    def drift(self, dt):
        """particles.py:259-265: drift, then periodic_x and periodic_y"""
        self._finish_pending()
        self._fused_valid = False
        self._last_op = None
        if self.gapped and self.order in (1, 2) and \
                (self._rep == "gapped" or self._to_gapped()):
            # on the gapped layout: the cell-stream sweep without gather / kick
            m = self.manifold
            comm = m.comm
            cnt = self._counts
            _lib.call("skb_drift_gapped", self._c, m.c, self.order, float(dt), TLX, TLY,
                      self._gap_start.data_ptr(), self._gap_count.data_ptr(),
                      self._movers.data_ptr(), self._movers.shape[0],
                      self.sbufl.data_ptr(), self.sbufr.data_ptr(), self.nbmax,
                      cnt.data_ptr(), comm.rank, comm.size, self._leftover.data_ptr(),
                      self._leftover.shape[1], self._gap_nleft, self._gcnt.data_ptr(),
                      self._scratch.data_ptr(), self._scr_rows, self._npool,
                      self._pool_owner.data_ptr(), _stream())
            self._gap_finish(cnt)
            self._gap_pushes += 1
            return
        self._dense()
        flags = _lib.EPI_HOLES | _lib.EPI_PERIODIC_X
        _lib.call("skb_drift", self._c, self.N, float(dt), self.manifold.c,
                  self._epilogue(flags), _stream())
        self.move()
        self.sort()
