"""Single-rank stand-in for mpi4py (test infrastructure for oracle/_ref only)."""
