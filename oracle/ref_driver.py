"""CPU baseline driver: the particle-step of the metric on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE (used only by bench.py's cpu_baseline leg and
`bench.py --impl reference`).  One particle-step = `ions.push(E, B, dt)` +
`sources.deposit(ions)` + `add_guards` + `copy_guards` (the loop body of reference
tests/test_ionacoustic.py:160-178 without Ohm).

kind = "reference": the unmodified reference kernels compiled into oracle/_ref
  (boris_push_cic, calculate_ihole, cppmove2, periodic_x, deposit_cic — the calls
  Particles.push / Sources.deposit make, particles.py:159-188, sources.py:27-50)
  driven directly, i.e. without the Python-level `assert all(...)` residency checks
  (what `python -O` strips); the guard-cell / normalize steps, which
  are NumPy slicing in the reference too, run through the restatement in
  oracle/oracle.py.
kind = "port": the C restatement (oracle/skeletor_oracle.c), if oracle/_ref is
  missing.

There is no MPI on the box: P independent single-rank processes each own a
ny/P-row periodic slab of the workload — a communication-free UPPER bound on what
`mpirun -np P` of the reference could do.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _slab_setup(nx, rows, ppc, seed, vt_cells=0.1):
    """uniform Maxwellian slab: x~U[0,nx), y~U[0,rows), vt*dt/dx = vt_cells"""
    from oracle import oracle as orc
    g = orc.Grid(nx, rows, lbx=1, lby=1, Lx=1.0, Ly=rows/nx)
    n = nx*rows*ppc
    rng = np.random.default_rng(seed)
    p = np.zeros(int(1.25*n) + 16, orc.Particle)
    p["x"][:n] = rng.uniform(0, nx, n)
    p["y"][:n] = rng.uniform(0, rows, n)
    v = rng.normal(0.0, 1.0, (3, n))
    p["vx"][:n], p["vy"][:n], p["vz"][:n] = v
    dt = vt_cells*g.dx          # vt = 1
    E = g.field(orc.Float3)
    B = g.field(orc.Float3)
    xg, yg = np.meshgrid(g.x, g.y)
    a = (slice(g.lby, g.uby), slice(g.lbx, g.ubx))
    E["x"][a] = 0.01*np.sin(2*np.pi*xg/g.Lx)
    E["y"][a] = 0.01*np.cos(2*np.pi*yg/g.Ly)
    B["z"][a] = 1.0
    orc.copy_guards([E], [g])
    orc.copy_guards([B], [g])
    return g, p, n, E, B, dt


def run_slab(args):
    """worker: returns (particle_steps, seconds) for `steps` timed steps"""
    nx, rows, ppc, steps, warmup, seed, kind = args
    from oracle import oracle as orc
    g, p, n, E, B, dt = _slab_setup(nx, rows, ppc, seed)
    qtmh = 1.0/1.0*dt/2
    src = g.field(orc.Float4)
    if kind == "reference":
        from oracle import ref
        k = ref.kernels()
        rg = k.types.grid_t()
        rg.nx, rg.ny, rg.comm = g.nx, g.ny, k.MPI.COMM_WORLD
        rg.edges = list(g.edges)
        rg.nyp, rg.noff = g.nyp, g.noff
        rg.lbx, rg.lby, rg.ubx, rg.uby = g.lbx, g.lby, g.ubx, g.uby
        rg.dx, rg.dy, rg.Lx, rg.Ly, rg.x0, rg.y0 = g.dx, g.dy, g.Lx, g.Ly, g.x0, g.y0
        cwd = os.getcwd()
        os.chdir(os.environ.get("TMPDIR", "/tmp"))   # cppinit2 drops a file "C.2"
        k.ppic2_wrapper.cppinit(k.MPI.COMM_WORLD)
        os.chdir(cwd)
        nb = int(max(0.1*p.shape[0], 1))
        ihole = np.zeros(2*nb, np.int32)
        bufs = [np.zeros(nb, orc.Particle) for _ in range(4)]
        info = np.zeros(7, np.int32)

        def step(n):
            k.particle_push.boris_push_cic(p[:n], E, B, qtmh, dt, rg)
            k.particle_boundary.calculate_ihole(p[:n], ihole, rg)
            n = k.ppic2_wrapper.cppmove2(p, n, bufs[0], bufs[1], bufs[2], bufs[3],
                                         ihole, info, rg)
            k.particle_boundary.periodic_x(p[:n], rg)
            src.fill(0.0)
            k.deposit.deposit_cic(p[:n], src, rg, 0.0)
            orc.normalize([src], [g], [n], 1.0, 1.0)
            orc.add_guards([src], [g])
            orc.copy_guards([src], [g])
            return n
    else:
        def step(n):
            orc.push(p[:n], E, B, g, 1, qtmh, dt)
            (q,), (n2,) = orc.move([p], [n], [g])
            p[:] = q
            orc.periodic_x(p[:n2], g)
            src.fill(0.0)
            orc.deposit(p[:n2], src, g, 1)
            orc.normalize([src], [g], [n2], 1.0, 1.0)
            orc.add_guards([src], [g])
            orc.copy_guards([src], [g])
            return n2
    for _ in range(warmup):
        n = step(n)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        n = step(n)
        done += n
    return done, time.perf_counter() - t0


def measure(nx=2048, rows=8, ppc=256, steps=6, warmup=1, procs=None):
    """Run `procs` independent slabs concurrently; aggregate throughput =
    sum(particle-steps) / max(wall time)."""
    import multiprocessing as mp
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    procs = procs or host_cores()
    args = [(nx, rows, ppc, steps, warmup, 1234 + r, kind) for r in range(procs)]
    if procs == 1:
        res = [run_slab(args[0])]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(procs) as pool:
            res = pool.map(run_slab, args)
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return dict(value=total/wall, unit="particle-steps/s", cores=procs, kind=kind,
                sample="%d procs x (%dx%d cells x %d ppc = %d particles) x %d steps, "
                       "independent periodic slabs (no MPI on the box), CIC, float64"
                       % (procs, nx, rows, ppc, nx*rows*ppc, steps),
                seconds=wall, ms_per_step=1e3*wall/steps)


if __name__ == "__main__":
    print(measure(rows=int(sys.argv[1]) if len(sys.argv) > 1 else 2,
                  steps=3, procs=int(sys.argv[2]) if len(sys.argv) > 2 else 2))
