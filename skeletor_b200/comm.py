"""Communicator shim with the mpi4py surface skeletor uses, on torch.distributed.

The reference talks to MPI through two doors: mpi4py's pickled lowercase API from
Python (field.py:52-58, sources.py:59, time steppers) and raw MPI inside ppic2
(pplib2.c:741-753, 873).  Here both go through ONE torch.distributed process group
(NCCL on GPUs, gloo in the CPU tests): `ring_exchange` is the nearest-neighbour
sendrecv on device tensors (halo rows, migrating particles), `allreduce` the few
scalar reductions.  `COMM_WORLD`, `COMM_SELF`, `SUM`, `MAX` mirror mpi4py.MPI.
"""
import os

import numpy as np

SUM = "sum"
MAX = "max"
MIN = "min"


class SelfComm:
    """Single-rank communicator (mpi4py.MPI.COMM_SELF)."""
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allreduce(self, x, op=SUM):
        return x

    def allgather(self, x):
        return [x]

    def allreduce_tensor_(self, t, op=SUM):
        """in-place reduction of a device tensor over the ranks, ordered on the current
        stream (no host round trip)"""
        return t

    def gather(self, x, root=0):
        return [x]

    def bcast(self, x, root=0):
        return x

    def barrier(self):
        pass

    Barrier = barrier

    def sendrecv(self, sendobj, dest=0, source=0, **kw):
        return sendobj

    def ring_exchange(self, up, down, recv_below=None, recv_above=None):
        """send `up` to rank+1 and `down` to rank-1; returns (from_below, from_above)"""
        return up, down

    def exchange_counts(self, n_up, n_down):
        return n_up, n_down

    def halo_exchange(self, up, down):
        return up, down


class TorchComm:
    """torch.distributed-backed communicator (one process per GPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self._halo_arena = None

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def _device(self):
        import torch
        if self.backend == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def allreduce(self, x, op=SUM):
        import torch
        dist = self._dist
        rop = {SUM: dist.ReduceOp.SUM, MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN}[op]
        scalar = np.isscalar(x) or (isinstance(x, np.ndarray) and x.ndim == 0)
        a = np.asarray(x)
        dt = torch.int64 if a.dtype.kind in "iub" else torch.float64
        t = torch.as_tensor(a.astype(np.int64 if dt == torch.int64 else np.float64)
                            ).reshape(-1).to(self._device())
        dist.all_reduce(t, op=rop, group=self.group)
        out = t.cpu().numpy().reshape(a.shape)
        if scalar:
            return int(out) if dt == torch.int64 else float(out)
        return out

    def allgather_tensor(self, t):
        """concatenation (dim 0) of every rank's equally shaped device tensor, on the
        device: NCCL all-gather; host-staged over gloo"""
        import torch
        dist = self._dist
        t = t.contiguous()
        out = torch.empty((self.size*t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        if self.backend == "nccl":
            dist.all_gather_into_tensor(out, t, group=self.group)
        elif not t.is_cuda:
            parts = [torch.empty_like(t) for _ in range(self.size)]
            dist.all_gather(parts, t, group=self.group)
            out.copy_(torch.cat(parts))
        else:
            parts = [torch.empty(t.shape, dtype=t.dtype) for _ in range(self.size)]
            dist.all_gather(parts, t.cpu(), group=self.group)
            out.copy_(torch.cat(parts))
        return out

    def allreduce_tensor_(self, t, op=SUM):
        """in-place reduction of a device tensor over the ranks: NCCL on the current
        stream (no host round trip); host-staged over gloo"""
        dist = self._dist
        rop = {SUM: dist.ReduceOp.SUM, MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN}[op]
        if self.backend == "nccl" or not t.is_cuda:
            dist.all_reduce(t, op=rop, group=self.group)
        else:
            h = t.cpu()
            dist.all_reduce(h, op=rop, group=self.group)
            t.copy_(h)
        return t

    def allgather(self, x):
        out = [None]*self.size
        self._dist.all_gather_object(out, x, group=self.group)
        return out

    def gather(self, x, root=0):
        out = self.allgather(x)
        return out if self.rank == root else None

    def bcast(self, x, root=0):
        box = [x]
        self._dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def barrier(self):
        self._dist.barrier(group=self.group)

    Barrier = barrier

    def sendrecv(self, sendobj, dest=0, source=0, **kw):
        """object sendrecv (mpi4py lowercase API): used only off the hot path"""
        import torch
        if dest == self.rank and source == self.rank:
            return sendobj
        a = np.ascontiguousarray(sendobj)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(self._device())
        r = torch.empty_like(t)
        dist = self._dist
        ops = [dist.P2POp(dist.isend, t, dest, self.group),
               dist.P2POp(dist.irecv, r, source, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return r.cpu().numpy().view(a.dtype).reshape(a.shape)

    def ring_exchange(self, up, down, recv_below=None, recv_above=None):
        """Nearest-neighbour exchange of two tensors.

        `up` goes to rank+1, `down` to rank-1 (periodic ring, as field.py:15-16);
        returns (from_below, from_above).  One grouped NCCL send/recv.
        recv_below / recv_above: preallocated receive tensors when the incoming
        sizes differ from the outgoing ones (particle migration)."""
        import torch
        dist = self._dist
        above = (self.rank + 1) % self.size
        below = (self.rank - 1) % self.size
        from_below = torch.empty_like(up) if recv_below is None else recv_below
        from_above = torch.empty_like(down) if recv_above is None else recv_above
        if self.backend != "nccl" and up.is_cuda:
            # gloo has no device-side send/recv: the MESSAGES are staged through host
            # memory (several ranks sharing one GPU in the 1-GPU parity tests; every
            # kernel still runs on the device)
            fb, fa = self.ring_exchange(up.cpu(), down.cpu(),
                                        torch.empty(from_below.shape, dtype=up.dtype),
                                        torch.empty(from_above.shape, dtype=down.dtype))
            from_below.copy_(fb)
            from_above.copy_(fa)
            return from_below, from_above
        if self.size == 2:
            # both neighbours are the same peer: order the two messages explicitly
            ops = [dist.P2POp(dist.isend, up, above, self.group, tag=0),
                   dist.P2POp(dist.irecv, from_below, below, self.group, tag=0),
                   dist.P2POp(dist.isend, down, below, self.group, tag=1),
                   dist.P2POp(dist.irecv, from_above, above, self.group, tag=1)]
        else:
            ops = [dist.P2POp(dist.isend, up, above, self.group),
                   dist.P2POp(dist.isend, down, below, self.group),
                   dist.P2POp(dist.irecv, from_below, below, self.group),
                   dist.P2POp(dist.irecv, from_above, above, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return from_below, from_above

    def halo_exchange(self, up, down):
        """ring exchange of two equally-shaped halo messages; through NVLink peer
        memory (skeletor_b200/peer.py) when available, else NCCL send/recv"""
        from . import peer
        if not (up.is_cuda and peer.available(self)):
            return self.ring_exchange(up, down)
        n = max(up.numel(), down.numel())
        if self._halo_arena is None or self._halo_arena.cap < n:
            # (collective: every rank reaches the same exchange with the same sizes)
            self._halo_arena = peer.PeerArena(self, max(n, 1 << 16))
        fb, fa = self._halo_arena.exchange(up, down)
        return fb[:up.numel()].view(up.shape), fa[:down.numel()].view(down.shape)

    def exchange_counts(self, n_up, n_down):
        """tell the neighbours how many particles are coming; returns
        (n_from_below, n_from_above)"""
        import torch
        dev = self._device()
        up = torch.tensor([n_up], dtype=torch.int64, device=dev)
        dn = torch.tensor([n_down], dtype=torch.int64, device=dev)
        fb, fa = self.ring_exchange(up, dn)
        return int(fb.item()), int(fa.item())


COMM_SELF = SelfComm()
_world = None


def init_world():
    """COMM_WORLD: the torch.distributed world if launched under torchrun (one
    process per GPU; RANK/WORLD_SIZE/MASTER_* from the environment), else a
    single-rank communicator."""
    global _world
    if _world is not None:
        return _world
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        _world = TorchComm()
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        # SKELETOR_B200_BACKEND=gloo: several ranks on ONE GPU (NCCL refuses that);
        # used by the 1-GPU run of tests/test_gpu_multirank.py
        backend = os.environ.get("SKELETOR_B200_BACKEND", backend)
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) %
                                  torch.cuda.device_count())
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
        _world = TorchComm()
    else:
        _world = SelfComm()
    return _world


class _WorldProxy:
    """lazy COMM_WORLD so that importing the package never initialises NCCL"""

    def __getattr__(self, name):
        return getattr(init_world(), name)


COMM_WORLD = _WorldProxy()
