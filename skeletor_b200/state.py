class State:

    def __init__(self, species, B, time=0.0):
        """State class (reference skeletor/state.py:1-17).
        species: list of Particles objects or a single Particles object (stored
        as a one-element list)."""
        if isinstance(species, list):
            self.species = species
        else:
            self.species = [species]
        self.B = B
        self.t = time
