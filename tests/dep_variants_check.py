"""Helper of test_deposit_kernel_variants: deposit 256 and 20 particles per cell, both
orders, with the kernel selection of the environment (SKB_DEP_RING / SKB_DEP_PAIR, read
once per process by csrc/deposit.cu) and compare with the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from refutil import random_particles  # noqa: E402
import gpuutil as gu  # noqa: E402
from skeletor_b200 import _lib  # noqa: E402

worst = 0.0
for ppc in (256, 20):
    for order in (1, 2):
        g = orc.Grid(nx=32, ny=16, lbx=2, lby=2)
        rng = np.random.default_rng(50 + ppc + order)
        n = 32*16*ppc
        p = random_particles(g, n, rng)
        exp = g.field(orc.Float4)
        orc.deposit(p, exp, g, order, 0.0)
        tl = gu.Tiling(g, order)
        t = tl.sort(gu.soa(p), n)
        cur = torch.zeros((g.myp, g.mx, 4), dtype=torch.float64, device="cuda")
        _lib.call("skb_deposit", gu.cparts(t), n, cur.data_ptr(), gu.cgrid(g), order, 0.0,
                  tl.c(), gu.stream())
        got = gu.host(cur, orc.Float4).view(np.float64)
        e = exp.view(np.float64)
        err = float(np.abs(got - e).max()/np.abs(e).max())
        worst = max(worst, err)
        assert err < 1e-12, (ppc, order, err)
print("OK", worst)
