"""CPU: the C-ABI shared library loads and exports every symbol that
include/skeletor_b200.h declares, and the ctypes binding covers all of them."""
import ctypes
import os
import re

from skeletor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    txt = open(os.path.join(ROOT, "include", "skeletor_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_binding_covers_header():
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER)
    assert set(declared()) == bound


def test_no_gpu_call_needed_to_load():
    lib = _lib.load()
    assert lib.skb_version() >= 100
    assert lib.skb_ihole_scratch_ints(1000) >= 4


def test_struct_layout_matches_header():
    # skb_grid_t: 8 ints + 6 doubles + 2 doubles
    assert ctypes.sizeof(_lib.GridT) == 8*4 + 8*8
    assert ctypes.sizeof(_lib.ParticlesT) == 5*8
    assert _lib.TilingT.n_sorted.offset % 8 == 0


def test_product_does_not_import_the_oracle():
    """the product path must not route through oracle/ (the smoke check lives in
    __graft_entry__.smoke(), outside the package)"""
    for pkg in ("skeletor_b200", "compat"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in src.replace("# oracle", ""), f
