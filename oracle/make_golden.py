"""Generate tests/golden/*.npz by running tests/scenarios.py on the UNMODIFIED
reference package (single rank, serial MPI stand-in; oracle/ref.py: package()).

Build-container only (needs /root/reference).  The fixtures are committed; this
script is the record of how they were made:

    python oracle/build_ref.py && python oracle/make_golden.py

Test infrastructure, not product code.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def reference_namespace():
    from oracle import ref
    sk = ref.package()
    from mpi4py.MPI import COMM_WORLD
    from skeletor.manifolds.second_order import Manifold, ShearingManifold
    from skeletor.time_steppers.horowitz import TimeStepper as Horowitz
    from skeletor.time_steppers.predictor_corrector import TimeStepper as PC
    ns = types.SimpleNamespace(
        Manifold=Manifold, ShearingManifold=ShearingManifold,
        Particles=sk.Particles, Sources=sk.Sources, Field=sk.Field, Ohm=sk.Ohm,
        Faraday=sk.Faraday, State=sk.State, Float3=sk.Float3, comm=COMM_WORLD,
        Poisson=sk.Poisson,
        HorowitzStepper=Horowitz, PredictorCorrectorStepper=PC)
    return ns


def main():
    import scenarios
    ns = reference_namespace()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in scenarios.SCENARIOS.items():
        if only and name not in only:
            continue
        with contextlib.redirect_stdout(io.StringIO()):
            res = fn(ns)
        path = os.path.join(out, name + ".npz")
        np.savez_compressed(path, **res)
        print("%-28s %7.1f KB  %s" % (name, os.path.getsize(path)/1024,
                                      {k: v.shape for k, v in res.items()}))


if __name__ == "__main__":
    main()
