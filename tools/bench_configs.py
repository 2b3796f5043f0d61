"""Secondary measurements on BASELINE.json configs 2-4 (bench.py measures config 5, the
one the metric is quoted on).  Prints one JSON object; results are recorded in
profiles/ and DESIGN.md.  Synthetic uniform Maxwellian plasmas, float64, one GPU.

  python tools/bench_configs.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def uniform_plasma(sk, torch, m, ppc, order, seed=1234, gapped=False):
    n = m.nx*m.nyp*ppc
    nmax = int((1.36 if gapped else 1.05)*n) + 4096
    ions = sk.Particles(m, nmax + (nmax & 1), charge=1.0, mass=1.0, order=order,
                        nbmax=max(n//100, 1 << 16))
    ions.gapped = gapped
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    d = ions._data
    d[0, :n] = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*m.nx
    d[1, :n] = m.noff + torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)*m.nyp
    d[2:5, :n] = torch.randn((3, n), generator=gen, device="cuda", dtype=torch.float64)
    ions.N = n
    return ions, n


def timed(torch, fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale grid edge (testing)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--layout", default="gapped", choices=["gapped", "dense"],
                    help="particle layout for the push + deposit loops (configs 2, 4); "
                         "and the push_and_deposit sweeps of the time steppers (config 3)")
    a = ap.parse_args()
    gapped = a.layout == "gapped"
    import torch
    import skeletor_b200 as sk
    from skeletor_b200.time_steppers.horowitz import TimeStepper as Horowitz
    from skeletor_b200.time_steppers.predictor_corrector import TimeStepper as PC
    comm = sk.COMM_SELF
    out = {}
    sc = lambda v: max(32, int(v*a.scale))

    # config 2: Landau/ion-acoustic loop, 1024^2 x 64 ppc, CIC, Ohm included
    nx = sc(1024)
    m = sk.Manifold(nx, nx, comm, Lx=1.0, Ly=1.0)
    ions, n = uniform_plasma(sk, torch, m, 64, 1, gapped=gapped)
    dt = 0.1*m.dx
    E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3); B.copy_guards()
    src = sk.Sources(m)
    ohm = sk.Ohm(m, temperature=1.0, charge=1.0)

    def step2():
        ions.push(E, B, dt)
        src.deposit(ions)
        src.add_guards()
        src.copy_guards()
        ohm(src, B, E)
        E.copy_guards()
    src.deposit(ions, set_boundaries=True)
    ms = timed(torch, step2, a.steps)
    out["config2_landau_loop"] = {"grid": [nx, nx], "ppc": 64, "particles": n,
                                  "ms_per_step": ms, "particle_steps_per_s": n/ms*1e3,
                                  "layout": ions._rep}
    del ions, E, B, src
    torch.cuda.empty_cache()

    # config 3: hybrid steppers with B != 0, 2048^2 x 128 ppc, lbx = lby = 2, CIC
    nx = sc(2048)
    for name, cls, sweeps in (("config3_horowitz_iterate", Horowitz, 1),
                              ("config3_predictor_corrector_iterate", PC, 2)):
        # dx = 0.5 ion skin depths, dt = 1e-2 (example/ion_cyclotron_instability.py:
        # the Hall term makes the scheme stiff, dt ~ dx^2)
        m = sk.Manifold(nx, nx, comm, lbx=2, lby=2, Lx=0.5*nx, Ly=0.5*nx)
        # quiet start (11 x 11 sub-lattice = 121 ppc ~ the config's 128), cold-ish ions:
        # a noisy 128-ppc start drives O(1) electric fields through grad ln(rho)
        ions, n = uniform_plasma(sk, torch, m, 121, 1, gapped=gapped)
        sq = 11
        ax = (torch.arange(nx*sq, device="cuda", dtype=torch.float64) + 0.5)/sq
        ions._data[0, :n] = ax.repeat(nx*sq)
        ions._data[1, :n] = ax.repeat_interleave(nx*sq)
        ions._data[2:5, :n] *= 0.1
        ions._sorted = False
        B = sk.Field(m, dtype=sk.Float3)
        B.fill((1.0, 0.0, 0.0))
        B.copy_guards()
        e = cls(sk.State(ions, B), sk.Ohm(m, temperature=0.01, charge=1.0), m)
        dt = 1e-2
        # (no prepare(): the iteration to t=0 consistency is set-up, not the step)
        ions.deposit(set_boundaries=True)
        e.sources.t.copy_(ions.sources.t)
        e.sources.boundaries_set = True
        e.ohm(e.sources, e.B, e.E, set_boundaries=True)
        k0 = time.time()
        try:
            ms = timed(torch, lambda: e.iterate(dt), a.steps, warmup=1)
            t_pd = timed(torch, lambda: ions.push_and_deposit(e.E, e.B, dt, True), 3, 1)
            out[name] = {"grid": [nx, nx], "ppc": 121, "particles": n, "ms_per_iterate": ms,
                         "particle_sweeps_per_iterate": sweeps,
                         "particle_steps_per_s": sweeps*n/ms*1e3,
                         "push_and_deposit_update_ms": t_pd, "layout": ions._rep}
        except RuntimeError as err:
            out[name] = {"error": str(err)[:200]}
        del ions, B, e
        torch.cuda.empty_cache()

    # config 4: shearing sheet, push_modified + sheared guards, 2048^2 x 64 ppc
    m = sk.ShearingManifold(nx, nx, comm, lbx=2, lby=2, S=-1.5, Omega=1.0, Lx=1.0, Ly=1.0)
    ions, n = uniform_plasma(sk, torch, m, 64, 1, gapped=gapped)
    ions._data[2:5, :n] *= 0.05
    E = sk.Field(m, dtype=sk.Float3); E.copy_guards()
    B = sk.Field(m, dtype=sk.Float3); B.copy_guards()
    src = sk.Sources(m)
    dt = 0.1*m.dx
    tt = [0.0]

    def step4():
        ions.push_modified(E, B, dt)
        tt[0] += dt
        src.deposit(ions)
        src.time = tt[0]
        src.add_guards()
        src.copy_guards()
    ms = timed(torch, step4, a.steps)
    out["config4_shearing_sheet"] = {"grid": [nx, nx], "ppc": 64, "particles": n,
                                     "ms_per_step": ms, "particle_steps_per_s": n/ms*1e3,
                                     "layout": ions._rep}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
