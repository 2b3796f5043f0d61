"""Build the UNMODIFIED reference kernels into oracle/_ref/ (test oracle + CPU baseline).

Test infrastructure, not product code.  Nothing under skeletor_b200/ may import
what this produces; only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs do.

What it does
------------
* compiles the serial mpi4py stand-in (oracle/mpi_stub/mpi4py/MPI.pyx) to
  oracle/_ref/mpi4py/MPI*.so;
* cythonizes the reference's own skeletor/cython/*.pyx **where they lie** under
  /root/reference (generated C goes to oracle/_ref/build, never into the
  reference tree, never into git) and links them with the reference's
  picksc/ppic2/{pplib2,ppush2}.c against oracle/mpi_stub/mpi.h
  (rank 0 of 1; cppmove2 / cpptpose take their nvp==1 branches,
  reference picksc/ppic2/pplib2.c:715-730, 446-453);
* installs the resulting extension modules as oracle/_ref/skeletor/cython/*.so.

Compiler directives follow the reference's setup.py:8-18 (boundscheck=False,
cdivision=True, wraparound=False) plus language_level=2 (the reference uses
implicit relative cimports, e.g. `from types cimport ...`) and
legacy_implicit_noexcept=True (Cython 0.2x semantics the reference was written
for).  gcc -O2 on x86-64 without -march emits SSE2 and no FMA, which is what
the bit-parity claims in DESIGN.md rest on.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
No-op (exit 0) when the reference tree is absent (e.g. on the GPU box, where the
prebuilt oracle/_ref/ that travelled with the snapshot is used as is).
"""
import argparse
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
STUB = os.path.join(HERE, "mpi_stub")

MODULES = ["types", "particle_push", "deposit", "particle_boundary",
           "push_and_deposit", "finite_difference", "operators",
           "ppic2_wrapper"]


def built():
    so = glob.glob(os.path.join(OUT, "skeletor", "cython", "*.so"))
    names = {os.path.basename(s).split(".")[0] for s in so}
    return all(m in names for m in MODULES) and \
        bool(glob.glob(os.path.join(OUT, "mpi4py", "MPI*.so")))


def build(reference="/root/reference", force=False, quiet=True):
    if not os.path.isdir(os.path.join(reference, "skeletor", "cython")):
        return built()
    if built() and not force:
        return True

    from setuptools import Extension
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext
    from Cython.Build import cythonize
    from numpy import get_include

    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["LDSHARED"] = "/usr/bin/gcc -shared"
    os.environ.setdefault("CFLAGS", "-O2")
    build_dir = os.path.join(OUT, "build")
    os.makedirs(build_dir, exist_ok=True)
    cflags = ["-O2", "-Wno-unused-function", "-w", "-fopenmp"]

    def run(exts, cwd):
        old = os.getcwd()
        os.chdir(cwd)
        try:
            dist = Distribution({"ext_modules": exts})
            cmd = build_ext(dist)
            cmd.build_lib = OUT
            cmd.build_temp = os.path.join(build_dir, "temp")
            cmd.inplace = 0
            cmd.ensure_finalized()
            cmd.run()
        finally:
            os.chdir(old)

    # 1. the mpi4py stand-in (our own code)
    stub_ext = cythonize(
        [Extension("mpi4py.MPI", [os.path.join(STUB, "mpi4py", "MPI.pyx")],
                   include_dirs=[STUB], extra_compile_args=cflags)],
        include_path=[STUB], build_dir=build_dir, quiet=quiet,
        compiler_directives={"language_level": 3})
    run(stub_ext, STUB)
    shutil.copy(os.path.join(STUB, "mpi4py", "__init__.py"),
                os.path.join(OUT, "mpi4py", "__init__.py"))

    # 2. the reference's own extensions, sources left where they are
    cy = os.path.join(reference, "skeletor", "cython")
    exts = [Extension("skeletor.cython." + m,
                      [os.path.join("skeletor", "cython", m + ".pyx")],
                      include_dirs=[get_include(), STUB, cy,
                                    os.path.join(reference, "picksc", "ppic2")],
                      extra_compile_args=cflags,
                      extra_link_args=["-fopenmp"])
            for m in MODULES]
    old = os.getcwd()
    os.chdir(reference)   # `# distutils: sources = picksc/...` is cwd-relative
    try:
        exts = cythonize(
            exts, include_path=[STUB, cy], build_dir=build_dir, quiet=quiet,
            compiler_directives={"boundscheck": False, "cdivision": True,
                                 "wraparound": False, "language_level": 2,
                                 "legacy_implicit_noexcept": True})
    finally:
        os.chdir(old)
    run(exts, reference)
    for d in ("skeletor", os.path.join("skeletor", "cython")):
        init = os.path.join(OUT, d, "__init__.py")
        if not os.path.exists(init):
            # empty package markers so the compiled kernels import standalone
            open(init, "w").close()
    return built()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force, quiet=not a.verbose)
    print("oracle/_ref:", "ready" if ok else "NOT available")
    sys.exit(0)
