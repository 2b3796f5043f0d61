"""State handed to the time steppers (reference skeletor/state.py)."""


class State:
    """Particle species + magnetic field + time.  `species` may be one Particles object
    or a list of them; it is always stored as a list."""

    def __init__(self, species, B, time=0.0):
        self.species = species if isinstance(species, list) else [species]
        self.B = B
        self.t = time
