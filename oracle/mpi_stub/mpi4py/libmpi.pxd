cdef extern from "mpi.h":
    ctypedef int MPI_Comm
    MPI_Comm MPI_COMM_WORLD
    MPI_Comm MPI_COMM_SELF
