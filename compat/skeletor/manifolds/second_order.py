from skeletor_b200.manifolds.second_order import Manifold, ShearingManifold  # noqa: F401
from skeletor_b200.poisson import PoissonSolver  # noqa: F401
