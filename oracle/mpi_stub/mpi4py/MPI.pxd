from mpi4py.libmpi cimport MPI_Comm

cdef class Comm:
    cdef MPI_Comm ob_mpi
    cdef public int rank
    cdef public int size
