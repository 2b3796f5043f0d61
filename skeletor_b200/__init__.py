"""skeletor_b200 — B200-native drop-in for the particle hot path of
nbia-astro/skeletor (push, deposit, guard cells, migration) behind skeletor's own
Python API.  Compute lives in libskeletor_b200.so (hand-written CUDA for sm_100a,
include/skeletor_b200.h); PyTorch owns device memory, streams and NCCL.
There is no CPU fallback: objects that hold device data raise if CUDA or the
library is missing."""
# flake8: noqa
from .types import Complex, Complex2, Float, Float2, Float3, Float4, Int, Particle
from .grid import Grid
from .field import Field
from .sources import Sources
from .particles import Particles
from .ohm import Ohm
from .faraday import Faraday
from .poisson import Poisson
from .state import State
from .initial_condition import InitialCondition, DensityPertubation
from .io import IO
from .manifolds.second_order import Manifold, ShearingManifold
from . import comm
from .comm import COMM_WORLD, COMM_SELF


def cppinit(comm):
    """reference skeletor/cython/ppic2_wrapper.pyx:30-41 -> (idproc, nvp)"""
    return comm.rank, comm.size
