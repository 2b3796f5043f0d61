"""Loader for the compiled, unmodified reference (oracle/_ref/).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this.  Product code (skeletor_b200/) never does.

Two levels:

* ``kernels()`` — the reference's compiled Cython/C kernels only
  (skeletor/cython/*.so built by oracle/build_ref.py).  Self-contained: works on
  the GPU box, where /root/reference does not exist.
* ``package()`` — the whole reference Python package (`skeletor.Particles`,
  `Sources`, `Field`, `Manifold`, time steppers ...) imported from
  /root/reference with its `skeletor.cython` sub-package redirected to the
  compiled modules in oracle/_ref.  Only available in the build container; it is
  what oracle/make_golden.py uses to generate tests/golden/*.npz.
"""
import importlib
import importlib.util
import os
import sys
import types as _types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
REFERENCE_TREE = os.environ.get("SKELETOR_REFERENCE", "/root/reference")

_KERNEL_MODULES = ["types", "particle_push", "deposit", "particle_boundary",
                   "push_and_deposit", "finite_difference", "operators",
                   "ppic2_wrapper"]


def available():
    cy = os.path.join(REF, "skeletor", "cython")
    return os.path.isdir(cy) and any(f.startswith("particle_push") and
                                     f.endswith(".so") for f in os.listdir(cy))


def package_available():
    return available() and os.path.isfile(
        os.path.join(REFERENCE_TREE, "skeletor", "particles.py"))


def _ensure_stub_mpi():
    if "mpi4py" in sys.modules and not getattr(
            sys.modules["mpi4py"], "__file__", "").startswith(REF):
        raise RuntimeError("a real mpi4py is already imported; the serial "
                           "oracle needs the stand-in from oracle/_ref")
    if REF not in sys.path:
        sys.path.insert(0, REF)


class _Kernels:
    pass


_kernels = None


def kernels():
    """Namespace with the compiled reference modules as attributes."""
    global _kernels
    if _kernels is not None:
        return _kernels
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py)")
    _ensure_stub_mpi()
    if "skeletor" not in sys.modules:
        if package_available():
            package()
        else:
            pkg = _types.ModuleType("skeletor")
            pkg.__path__ = [os.path.join(REF, "skeletor")]
            sys.modules["skeletor"] = pkg
            cy = _types.ModuleType("skeletor.cython")
            cy.__path__ = [os.path.join(REF, "skeletor", "cython")]
            sys.modules["skeletor.cython"] = cy
    k = _Kernels()
    cwd = os.getcwd()
    for m in _KERNEL_MODULES:
        setattr(k, m, importlib.import_module("skeletor.cython." + m))
    os.chdir(cwd)
    import mpi4py.MPI as MPI
    k.MPI = MPI
    _kernels = k
    return k


_package = None


def package():
    """Import the reference `skeletor` package (build container only)."""
    global _package
    if _package is not None:
        return _package
    if not package_available():
        raise RuntimeError("reference tree or oracle/_ref missing")
    _ensure_stub_mpi()
    root = os.path.join(REFERENCE_TREE, "skeletor")
    # cppinit2 truncates a file "C.2" in the CWD on every Grid construction
    # (reference picksc/ppic2/pplib2.c:92): keep that out of the repo.
    scratch = os.path.join(REF, "scratch")
    os.makedirs(scratch, exist_ok=True)
    os.chdir(scratch)
    cy = _types.ModuleType("skeletor.cython")
    cy.__path__ = [os.path.join(REF, "skeletor", "cython")]
    sys.modules["skeletor.cython"] = cy
    spec = importlib.util.spec_from_file_location(
        "skeletor", os.path.join(root, "__init__.py"),
        submodule_search_locations=[root])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["skeletor"] = pkg
    pkg.cython = cy
    spec.loader.exec_module(pkg)
    _package = pkg
    return pkg
