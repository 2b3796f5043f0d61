"""Element types of the ABI, as NumPy dtypes.

Same names and memory layout as the reference's `skeletor.cython.types`
(types.pyx:4-15): float64 reals, C int indices, (x, y[, z]) / (t, x, y, z) records for
vector fields and sources, and the 40-byte particle record.  On the device the fields
keep exactly this interleaved layout; the particles are stored as five separate arrays.
"""
import numpy as _np


def _record(*names, base):
    return [(n, base) for n in names]


Int = _np.dtype(_np.int32)
Float = _np.dtype(_np.float64)
Complex = _np.dtype(_np.complex128)

Float2 = _record('x', 'y', base=Float)
Float3 = _record('x', 'y', 'z', base=Float)
Float4 = _record('t', 'x', 'y', 'z', base=Float)        # t = charge density, xyz = current
Complex2 = _record('x', 'y', base=Complex)

Particle = _np.dtype(_record('x', 'y', 'vx', 'vy', 'vz', base=Float), align=True)
