"""ctypes binding of libskeletor_b200.so (the C ABI declared in include/skeletor_b200.h).

This is the ONLY compute backend: there is no CPU or PyTorch fallback.  Loading
fails loudly if the shared library has not been built, and every call fails
loudly if it returns a CUDA error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKELETOR_B200_LIB") or \
    os.path.join(_HERE, "lib", "libskeletor_b200.so")

c_int, c_ll, c_dbl, c_vp = C.c_int, C.c_longlong, C.c_double, C.c_void_p


class GridT(C.Structure):
    """skb_grid_t == reference grid_t (skeletor/cython/types.pxd:25-37)"""
    _fields_ = [("nx", c_int), ("ny", c_int), ("nyp", c_int), ("noff", c_int),
                ("lbx", c_int), ("lby", c_int), ("ubx", c_int), ("uby", c_int),
                ("dx", c_dbl), ("dy", c_dbl), ("Lx", c_dbl), ("Ly", c_dbl),
                ("x0", c_dbl), ("y0", c_dbl), ("edges", c_dbl*2)]


class ParticlesT(C.Structure):
    _fields_ = [("x", c_vp), ("y", c_vp), ("vx", c_vp), ("vy", c_vp), ("vz", c_vp)]


class TilingT(C.Structure):
    _fields_ = [("tile_offsets", c_vp), ("chunk_first_tile", c_vp),
                ("cell_end", c_vp), ("ntx", c_int), ("nty", c_int), ("tlx", c_int), ("tly", c_int),
                ("chunk", c_int), ("n_sorted", c_ll), ("gap_start", c_vp),
                ("gap_count", c_vp)]


class EpilogueT(C.Structure):
    _fields_ = [("flags", c_int), ("S", c_dbl), ("t", c_dbl), ("ihole", c_vp),
                ("ntmax", c_int), ("cell_counts", c_vp), ("key_order", c_int),
                ("key_tlx", c_int), ("key_tly", c_int)]


EPI_SHEAR, EPI_PERIODIC_X, EPI_HOLES, EPI_COUNT = 1, 2, 4, 8

_P = ParticlesT
_G = C.POINTER(GridT)
_T = C.POINTER(TilingT)
_E = C.POINTER(EpilogueT)

# name -> argtypes; every function returns int (cudaError_t) unless noted
SIGNATURES = {
    "skb_boris_push": [_P, c_ll, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_int, c_dbl,
                       c_dbl, _T, _E, c_vp],
    "skb_drift": [_P, c_ll, c_dbl, _G, _E, c_vp],
    "skb_periodic_x": [_P, c_ll, _G, c_vp],
    "skb_shear_periodic_y": [_P, c_ll, _G, c_dbl, c_dbl, c_vp],
    "skb_calculate_ihole": [_P, c_ll, c_vp, c_int, _G, c_vp, c_vp],
    "skb_move_pack": [_P, c_vp, c_int, c_vp, c_vp, c_int, c_vp, _G, c_int, c_int, c_vp],
    "skb_move_classify": [c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, _G, c_int,
                          c_int, c_vp],
    "skb_move_unpack": [_P, c_ll, c_vp, c_int, c_vp, c_int, c_vp, c_vp],
    "skb_deposit": [_P, c_ll, c_vp, _G, c_int, c_dbl, _T, c_vp],
    "skb_deposit_deterministic": [_P, c_ll, c_vp, _G, c_int, c_dbl, _T, c_vp, c_vp],
    "skb_push_and_deposit": [_P, c_ll, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_vp,
                             c_int, c_vp, c_dbl, c_int, _T, c_vp, c_int, c_int, c_vp],
    "skb_tile_geometry": [_G, c_int, c_int, C.POINTER(c_int), C.POINTER(c_int)],
    "skb_cell_keys": [_P, c_ll, _G, c_int, c_int, c_int, c_vp, c_vp],
    "skb_tile_sort": [_P, _P, c_ll, _G, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                      c_vp, c_int, c_vp, c_vp],
    "skb_push_count": [_P, c_ll, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_int, c_dbl,
                       c_dbl, _T, _E, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp,
                       c_int, c_int, c_vp],
    "skb_push_scatter": [_P, _P, c_ll, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_int,
                         c_dbl, c_dbl, _T, _E, c_int, c_int, c_vp, c_vp],
    "skb_sort_clear": [c_vp, _G, c_int, c_int, c_vp],
    "skb_sort_count_rows": [c_vp, c_int, _G, c_int, c_int, c_int, c_vp, c_vp],
    "skb_sort_scan": [c_vp, _G, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "skb_sort_scatter_rows": [c_vp, c_int, _P, _G, c_int, c_int, c_int, c_vp, c_vp],
    "skb_tile_sort_precounted": [_P, _P, c_ll, _G, c_int, c_int, c_int, c_int, c_vp,
                                 c_vp, c_vp, c_vp, c_vp],
    "skb_canonical_cells": [_P, _P, c_vp, _G, c_int, c_int, c_vp],
    "skb_gap_build": [_P, _P, c_vp, _G, c_int, c_int, c_vp, c_vp, c_vp, c_ll, c_vp],
    "skb_push_gapped": [_P, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_int, c_dbl, c_dbl,
                        c_int, c_dbl, c_dbl, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp,
                        c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp,
                        c_vp, c_int, c_int, c_vp, c_vp],
    "skb_push_and_deposit_gapped": [_P, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_vp, c_dbl,
                                    c_int, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                    c_int, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp,
                                    c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "skb_push_deposit_gapped": [_P, c_vp, c_vp, _G, c_int, c_dbl, c_dbl, c_int, c_dbl, c_dbl,
                                c_int, c_dbl, c_dbl, c_vp, c_dbl, c_int, c_int, c_vp, c_vp,
                                c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp,
                                c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "skb_drift_gapped": [_P, _G, c_int, c_dbl, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp,
                         c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_int,
                         c_int, c_vp, c_vp],
    "skb_deposit_rows": [c_vp, c_int, c_vp, _G, c_int, c_dbl, c_vp],
    "skb_gap_insert": [c_vp, c_int, _P, c_vp, c_vp, _G, c_int, c_int, c_int, c_vp, c_int,
                       c_vp, c_vp],
    "skb_gap_insert_counted": [c_vp, c_vp, c_int, _P, c_vp, c_vp, _G, c_int, c_int, c_int,
                               c_vp, c_int, c_vp, c_vp],
    "skb_peer_send": [c_vp, c_vp, c_int, c_vp, c_vp],
    "skb_gap_densify": [_P, _P, c_vp, c_vp, _G, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                        c_vp, c_int, c_int, c_ll, c_vp],
    "skb_exclusive_scan": [c_vp, c_int, c_vp, c_vp],
    "skb_chunk_table": [c_vp, _G, c_int, c_int, c_int, c_vp, c_vp],
    "skb_copy_guards": [c_vp, c_int, _G, c_vp, c_vp, c_vp],
    "skb_add_guards": [c_vp, c_int, _G, c_int, c_vp, c_vp, c_vp],
    "skb_pack_rows": [c_vp, c_int, _G, c_int, c_int, c_vp, c_vp],
    "skb_copy_guards_x_rows": [c_vp, c_int, _G, c_int, c_int, c_vp],
    "skb_scale": [c_vp, c_ll, c_dbl, c_vp],
    "skb_gradient": [c_vp, c_int, c_vp, _G, c_vp],
    "skb_curl": [c_vp, c_vp, c_vp, c_int, c_vp, _G, c_int, c_vp],
    "skb_divergence": [c_vp, c_vp, c_int, c_vp, _G, c_vp],
    "skb_interp": [c_vp, c_vp, c_vp, c_int, c_vp, _G, c_int, c_vp],
    "skb_ohm": [c_vp, c_vp, c_vp, c_vp, c_vp, _G, c_dbl, c_dbl, c_vp],
    "skb_faraday": [c_vp, c_vp, c_vp, _G, c_dbl, c_vp],
    "skb_field_combine": [c_vp, c_vp, c_vp, c_ll, c_dbl, c_int, c_vp, c_vp],
    "skb_faraday_to": [c_vp, c_vp, c_vp, c_vp, _G, c_dbl, c_vp, c_vp],
    "skb_ohm_if": [c_vp, c_vp, c_vp, c_vp, c_vp, _G, c_dbl, c_dbl, c_vp, c_vp],
    "skb_horowitz_update": [c_vp, c_vp, c_vp, _G, c_vp, c_vp, c_vp],
    "skb_converged": [c_vp, c_dbl, c_dbl, c_int, c_vp, c_vp],
    "skb_poisson_kspace": [c_vp, c_vp, c_vp, c_int, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                           c_int, c_vp, c_vp],
}
OTHER = {
    "skb_version": ([], c_int),
    "skb_error_string": ([c_int], C.c_char_p),
    "skb_ihole_scratch_ints": ([c_ll], c_ll),
}

_lib = None
launches = 0          # number of C-ABI calls made
kernel_launches = 0   # CUDA kernels those calls launched (bench.py's gpu_launches)

# kernels launched per C-ABI call (memsets are not counted)
KERNELS_PER_CALL = {"skb_tile_sort": 6, "skb_calculate_ihole": 3, "skb_move_unpack": 3,
                    "skb_sort_scan": 4, "skb_sort_clear": 0, "skb_tile_sort_precounted": 5, "skb_deposit_deterministic": 2,
                    "skb_gap_build": 5, "skb_gap_densify": 5, "skb_exclusive_scan": 3}


class SkeletorCudaError(RuntimeError):
    pass


def load():
    """Load the library (no GPU needed for loading; calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SkeletorCudaError(
            "libskeletor_b200.so is not built (%s). Run `python -c 'import "
            "__graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = c_int
    for name, (args, res) in OTHER.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res
    _lib = lib
    return lib


# measurement hook (bench.py): {entry point name: [(start event, end event), ...]};
# when set, the named calls are bracketed by CUDA events on the current stream
trace = None


def call(name, *args):
    """Invoke a C-ABI entry point, raising on a CUDA error."""
    global launches, kernel_launches
    lib = load()
    if trace is not None and name in trace:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        trace[name].append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    launches += 1
    kernel_launches += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        msg = lib.skb_error_string(rc)
        raise SkeletorCudaError("%s failed: CUDA error %d (%s)" % (
            name, rc, msg.decode() if msg else "?"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SkeletorCudaError(
            "skeletor_b200 needs a CUDA device (B200, sm_100a); there is no CPU "
            "fallback.")
    load()
