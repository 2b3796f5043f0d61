"""Host-side particle loading (reference skeletor/initial_condition.py:4-62): runs
once, on the host with NumPy, then uploads — out of the hot path by design."""
import numpy as np


class InitialCondition:

    def __init__(self, npc, quiet=False, vt=0.0, global_init=False):
        # Quiet start / particles per cell / ion thermal velocity /
        # initialize particles "globally" on each processor?
        self.quiet = quiet
        self.npc = npc
        self.vt = vt
        self.global_init = global_init

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

    def __call__(self, manifold, ions):
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
