# cython: language_level=3
"""Single-rank stand-in for mpi4py.MPI: rank 0 of 1, collectives are identity.

`Comm(rank, size)` may also be constructed with a *fake* rank/size so that the
N-slab emulator (oracle/slab_emulator.py) can build per-slab reference grids;
the C side (cppinit2) still sees one rank.
"""
from mpi4py.libmpi cimport MPI_COMM_WORLD, MPI_COMM_SELF

SUM = 'sum'
MAX = 'max'


def Is_initialized():
    return True


def Is_finalized():
    return False


cdef class Comm:
    def __cinit__(self, int rank=0, int size=1):
        self.ob_mpi = MPI_COMM_WORLD
        self.rank = rank
        self.size = size

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def allreduce(self, x, op=SUM):
        return x

    def allgather(self, x):
        return [x]

    def bcast(self, x, root=0):
        return x

    def gather(self, x, root=0):
        return [x]

    def barrier(self):
        pass

    def sendrecv(self, sendobj, dest=0, source=0, **kw):
        return sendobj


COMM_WORLD = Comm()
COMM_SELF = Comm()
