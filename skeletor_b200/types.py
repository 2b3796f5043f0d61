"""NumPy dtypes of the reference ABI (skeletor/cython/types.pyx:4-15)."""
from numpy import dtype

Int = dtype("i4")
Float = dtype("f8")
Complex = dtype("c16")

Float2 = [('x', Float), ('y', Float)]
Float3 = [('x', Float), ('y', Float), ('z', Float)]
Float4 = [('t', Float), ('x', Float), ('y', Float), ('z', Float)]
Complex2 = [('x', Complex), ('y', Complex)]

Particle = dtype([('x', Float), ('y', Float), ('vx', Float), ('vy', Float),
                  ('vz', Float)], align=True)
