"""mpi4py import name for reference-style scripts (torch.distributed shim)."""
from . import MPI  # noqa: F401
