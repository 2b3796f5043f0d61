"""Tuning aid: build variants of libskeletor_b200.so with different cell-stream kernel
configurations (CS_* macros of csrc/cellstream.cu) -> skeletor_b200/lib/variants/lib_<name>.so;
tools/run_variants.sh benches each of them through SKELETOR_B200_LIB."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

VARIANTS = {
    "G_w4_b4_nst2": ["CS_WARPS=4", "CS_MINB=4", "CS_NST=2"],
    "G2_w4_b5_nst2": ["CS_WARPS=4", "CS_MINB=5", "CS_NST=2"],
    "G3_w4_b4_nst2_s32": ["CS_WARPS=4", "CS_MINB=4", "CS_NST=2", "CS_STAGE=32"],
    "G4_w4_b4_nst2_hints": ["CS_WARPS=4", "CS_MINB=4", "CS_NST=2", "CS_L2HINTS=1"],
    "G5_w4_b4_nst2_np2": ["CS_WARPS=4", "CS_MINB=4", "CS_NST=2", "CS_NP2=1"],
    "G6_w4_b4_nst4_s32": ["CS_WARPS=4", "CS_MINB=4", "CS_NST=4", "CS_STAGE=32"],
    "B2_w8_b2_nst2": ["CS_NST=2"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    d = os.path.join(ROOT, "skeletor_b200", "lib", "variants")
    os.makedirs(d, exist_ok=True)
    for n in names:
        out = os.path.join(d, "lib_%s.so" % n)
        g.build_cuda(defines=VARIANTS[n], out=out)
        print(out)
