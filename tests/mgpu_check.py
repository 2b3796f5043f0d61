"""Multi-rank parity check, launched under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Runs the golden scenarios of tests/scenarios.py on N y-slabs and checks, on rank 0,
the reference's own invariant "N ranks == 1 rank" (reference
tests/test_skeletor.py:142-150) against the single-rank golden fixtures: particle
count exact, sorted particle coordinates and active-cell fields <= 1e-12 relative.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import scenarios as sc  # noqa: E402


def rows(p):
    a = np.ascontiguousarray(p).view(np.float64).reshape(-1, 5)
    return a[np.lexsort((a[:, 4], a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]


def main():
    import skeletor_b200 as sk
    from skeletor_b200.time_steppers.horowitz import TimeStepper as Horowitz
    from skeletor_b200.time_steppers.predictor_corrector import TimeStepper as PC
    comm = sk.comm.init_world()
    gapped = os.environ.get("MGPU_GAPPED", "0") == "1"
    pushes = [0]

    class GappedParticles(sk.Particles):
        """the same scenarios on the gapped particle layout"""

        def __init__(self, manifold, Nmax, **kw):
            super().__init__(manifold, 4*int(Nmax) + 8192, **kw)   # (empty cells own 4 slots each)
            self.gapped = True

        def _gap_kernel(self, *args, **kw):
            pushes[0] += 1
            return super()._gap_kernel(*args, **kw)

        def _push_and_deposit_gapped(self, *args, **kw):
            pushes[0] += 1
            return super()._push_and_deposit_gapped(*args, **kw)

    ns = types.SimpleNamespace(
        Manifold=sk.Manifold, ShearingManifold=sk.ShearingManifold,
        Particles=GappedParticles if gapped else sk.Particles, Sources=sk.Sources, Field=sk.Field, Ohm=sk.Ohm,
        Faraday=sk.Faraday, State=sk.State, Float3=sk.Float3, comm=comm,
        Poisson=sk.Poisson,
        HorowitzStepper=Horowitz, PredictorCorrectorStepper=PC)
    names = ["ionacoustic_cic", "ionacoustic_tsc", "gyro_cic", "gyro_tsc", "sheared_cic",
             "sheared_tsc", "predictor_corrector_tsc", "horowitz_cic", "poisson"]
    lb = {"ionacoustic_cic": 1, "poisson": 1}
    if gapped:
        names = names[:8]
    failed = 0
    for name in names:
        gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                res = sc.SCENARIOS[name](ns)
        except AssertionError as err:
            if "slab too small" not in str(err):
                raise
            # the fixture's grid has fewer rows per rank than guard layers at this rank
            # count (every rank hits the same assertion in Grid.__init__)
            if comm.rank == 0:
                print("%-26s ranks=%d SKIP  (%s)" % (name, comm.size, err), flush=True)
            continue
        # gather the slabs on every rank (object allgather: diagnostics path)
        g = lb.get(name, 2)
        rtol = 2e-12 if ("horowitz" in name or "predictor" in name) else 1e-12
        if name == "poisson":
            rtol = 2e-6       # float32 truncation in the reference (Q3)
        errs = {}
        ok, ntot = True, 0
        if "particles" in gold.files:
            parts = np.concatenate(comm.allgather(res["particles"]))
            ntot = comm.allreduce(int(res["N"]))
            ok = ntot == int(gold["N"])
            e = np.abs(rows(parts) - rows(gold["particles"])).max() / \
                np.abs(rows(gold["particles"])).max()
            errs["particles"] = e
            ok &= e <= rtol
        for key in gold.files:
            if key in ("N", "particles", "time", "t"):
                continue
            if key == "we":          # scalar diagnostic (field energy of the Poisson solve)
                e = abs(float(res[key]) - float(gold[key]))/abs(float(gold[key]))
                errs[key] = e
                ok &= e <= rtol
                continue
            act = np.concatenate(comm.allgather(
                np.ascontiguousarray(res[key][g:-g, g:-g])))
            ref = np.ascontiguousarray(gold[key][g:-g, g:-g])
            a = act.view(np.float64).ravel()
            b = ref.view(np.float64).ravel()
            e = np.abs(a - b).max()/np.abs(b).max()
            errs[key] = e
            ok &= e <= rtol
        if gapped:
            ok &= pushes[0] > 0          # the gapped push really ran
            errs["gapped_pushes"] = pushes[0]
            pushes[0] = 0
        if comm.rank == 0:
            print("%-26s ranks=%d N=%d %s  %s" % (
                name, comm.size, ntot, "OK  " if ok else "FAIL",
                " ".join("%s=%.1e" % kv for kv in errs.items())), flush=True)
        failed += (not ok)
    comm.barrier()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
