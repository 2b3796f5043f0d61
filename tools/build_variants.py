"""Tuning aid: build variants of libskeletor_b200.so with different cell-stream kernel
configurations (CS_* macros of csrc/cellstream.cu) -> skeletor_b200/lib/variants/lib_<name>.so;
tools/run_variants.sh benches each of them through SKELETOR_B200_LIB."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

VARIANTS = {
    "I8": ["GAP_INS_ITEMS=8"],
    "I4M4": ["GAP_INS_MINB=4"],
    "I2M6": ["GAP_INS_ITEMS=2", "GAP_INS_MINB=6"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    d = os.path.join(ROOT, "skeletor_b200", "lib", "variants")
    os.makedirs(d, exist_ok=True)
    for n in names:
        out = os.path.join(d, "lib_%s.so" % n)
        g.build_cuda(defines=VARIANTS[n], out=out)
        print(out)
