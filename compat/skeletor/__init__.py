"""`skeletor` import name for reference-style scripts: re-exports skeletor_b200."""
# flake8: noqa
from skeletor_b200 import *
from skeletor_b200 import (Complex, Complex2, Float, Float2, Float3, Float4, Int, Particle,
                           Grid, Field, Sources, Particles, Ohm, Faraday, Poisson, State,
                           InitialCondition, DensityPertubation, IO, cppinit)
