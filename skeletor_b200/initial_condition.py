"""Particle loading (reference skeletor/initial_condition.py:4-143).  Default: on the
host with NumPy (the reference's tests seed np.random), then uploaded.  on_device=True:
quiet / noisy start and the density perturbation are generated in device memory - what
a 1e9-particle start needs."""
import numpy as np


class InitialCondition:

    def __init__(self, npc, quiet=False, vt=0.0, global_init=False, on_device=False,
                 seed=None):
        # Quiet start / particles per cell / ion thermal velocity /
        # initialize particles "globally" on each processor?
        self.quiet = quiet
        self.npc = npc
        self.vt = vt
        self.global_init = global_init
        # extension: draw the (noisy start) coordinates with the GPU's generator straight
        # into device memory — needed to even start a 1e9-particle run (the host path
        # is kept as the default because the reference's tests seed np.random)
        self.on_device = on_device
        self.seed = seed

    def positions(self, nx, ny):
        N = nx*ny*self.npc
        if self.quiet:
            # regular sub-lattice in every cell (quiet start)
            sq = int(np.sqrt(self.npc))
            assert sq**2 == self.npc
            x1 = (np.arange(nx*sq) + 0.5)/sq
            y1 = (np.arange(ny*sq) + 0.5)/sq
            x, y = [xy.flatten() for xy in np.meshgrid(x1, y1)]
        else:
            # noisy start
            x = nx*np.random.uniform(size=N)
            y = ny*np.random.uniform(size=N)
        return x, y

    def _call_on_device(self, manifold, ions):
        """Coordinates generated straight into device memory (per-slab start): the quiet
        start's sub-lattice is the same float64 arithmetic as `positions` (bit-identical
        coordinates, initial_condition.py:24-30); noisy positions and the Maxwellian
        velocities come from the GPU's generator (seed + rank)."""
        import torch
        assert not self.global_init, "on_device initialises every slab on its own rank"
        nx, nyp = manifold.nx, manifold.nyp
        N = nx*nyp*self.npc
        dev = ions.device
        gen = torch.Generator(device=dev)
        gen.manual_seed((self.seed if self.seed is not None else 0) + manifold.comm.rank)
        kw = dict(generator=gen, device=dev, dtype=torch.float64)
        ions._rep = "dense"           # (whatever was stored before is discarded)
        d = ions._data
        if self.quiet:
            sq = int(np.sqrt(self.npc))
            assert sq**2 == self.npc
            x1 = (torch.arange(nx*sq, device=dev, dtype=torch.float64) + 0.5)/sq
            y1 = (torch.arange(nyp*sq, device=dev, dtype=torch.float64) + 0.5)/sq
            # np.meshgrid(x1, y1) flattened: x runs fastest
            d[0, :N] = x1.repeat(nyp*sq)
            d[1, :N] = y1.repeat_interleave(nx*sq) + manifold.edges[0]
        else:
            d[0, :N] = torch.rand(N, **kw)*nx
            d[1, :N] = torch.rand(N, **kw)*nyp + manifold.edges[0]
        d[2:5, :N] = torch.randn((3, N), **kw)*self.vt
        ions.N = N
        ions._sorted = False

    def __call__(self, manifold, ions):
        if self.on_device:
            return self._call_on_device(manifold, ions)
        nx = manifold.nx
        ny = manifold.ny if self.global_init else manifold.nyp
        N = nx*ny*self.npc
        x, y = self.positions(nx, ny)
        vx = self.vt*np.random.normal(size=N)
        vy = self.vt*np.random.normal(size=N)
        vz = self.vt*np.random.normal(size=N)
        if self.global_init:
            ions.initialize(manifold.x0 + x*manifold.dx,
                            manifold.y0 + y*manifold.dy, vx, vy, vz)
        else:
            ions['x'][:N] = x
            ions['y'][:N] = y + manifold.edges[0]
            ions['vx'][:N] = vx
            ions['vy'][:N] = vy
            ions['vz'][:N] = vz
            ions.N = N


class DensityPertubation(InitialCondition):
    """Sinusoidal density perturbation n = 1 + ampl cos(kx x + ky y) by displacing
    uniformly placed particles along x (reference initial_condition.py:65-143).

    The reference inverts the CDF particle by particle with scipy's secant `newton`
    in a Python loop (minutes at 1e6 particles); here the closed-form CDF
        cdf(X, y) = [(X - x0) + ampl/kx (sin(kx X + ky y) - sin(kx x0 + ky y))] / Lx
    is inverted for all particles at once with Newton's method (NumPy, host)."""

    def __init__(self, npc, ikx, iky, ampl, **kwds):
        super().__init__(npc, **kwds)
        self.ikx = ikx
        self.iky = iky
        self.ampl = ampl
        if self.ikx == 0:
            msg = """This class unfortunately cannot currently handle density
            perturbations that do not have an x-dependence. The reason is
            that particle positions are assumed to be uniformly placed along x.
            The density perturbations are created by varying the interparticle
            distance in the y-direction only."""
            raise RuntimeError(msg)

    def __call__(self, manifold, ions):
        super().__call__(manifold, ions)
        Lx, x0, y0 = manifold.Lx, manifold.x0, manifold.y0
        kx = self.ikx*2*np.pi/manifold.Lx
        ky = self.iky*2*np.pi/manifold.Ly
        N = ions.N
        A = self.ampl
        self.f = lambda x, y: 1 + A*np.cos(kx*x + ky*y)
        if self.on_device:
            # the same Newton iteration on the device tensors (no host copy of the
            # particles: this is what makes a 1e9-particle perturbed start possible)
            import torch
            ions._dense()
            d = ions._data
            u = d[0, :N]/manifold.nx
            y = y0 + d[1, :N]*manifold.dy
            s0 = torch.sin(kx*x0 + ky*y)
            X = x0 + u*Lx
            for it in range(50):
                ph = kx*X + ky*y
                f = ((X - x0) + A/kx*(torch.sin(ph) - s0))/Lx - u
                dX = f/((1 + A*torch.cos(ph))/Lx)
                X = X - dX
                if float(dX.abs().max()) < 1e-15*Lx:
                    break
            d[0, :N] = (X - x0)/manifold.dx
            ions._touched()
            return
        # x-coordinate in units of the box size, y in "physical" units
        u = np.asarray(ions['x'])[:N]/manifold.nx
        y = y0 + np.asarray(ions['y'])[:N]*manifold.dy
        s0 = np.sin(kx*x0 + ky*y)
        X = x0 + u*Lx                       # exact for ampl = 0
        for it in range(50):
            f = ((X - x0) + A/kx*(np.sin(kx*X + ky*y) - s0))/Lx - u
            dX = f/((1 + A*np.cos(kx*X + ky*y))/Lx)
            X = X - dX
            if np.abs(dX).max() < 1e-15*Lx:
                break
        ions['x'][:N] = (X - x0)/manifold.dx
        ions['y'][:N] = (y - y0)/manifold.dy
