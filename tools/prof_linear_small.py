#!/usr/bin/env python
"""Short launches of the small-system tensor kernel for ncu (k_linear_tc, one warp per path tile): d = 32 and d = 8, euler and heun.
    ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -c 4 -o gpurun_out/prof python tools/prof_linear_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdetools_b200 as sde  # noqa: E402

eng = sde.default_engine(0)
dev = torch.device("cuda", 0)
for d, scheme, P in ((32, "euler", 1 << 18), (32, "heun", 1 << 18), (8, "euler", 1 << 20), (16, "euler", 1 << 19)):
    R = np.random.default_rng(d).standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    params = np.concatenate([A.ravel(order="F"), np.eye(d).ravel(order="F")])
    x0 = np.zeros(d)
    x0[0] = 1.0
    n = 100
    XT = torch.empty(P * d, dtype=torch.float64, device=dev)
    ms = eng.run("linear", scheme, params, np.arange(n + 1) / n, x0, nx=d, nB=d, npaths=P, seed=4, XT=XT, want_time=True)
    print(f"d={d} {scheme}: {ms:.3f} ms, {P * n / ms / 1e-3:.3g} path-steps/s")
