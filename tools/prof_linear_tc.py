#!/usr/bin/env python
"""Short launches of the general tensor kernel for ncu (k_linear_tc): config C4 with heun at d = 256 and euler at d = 512.
    ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -c 2 -o gpurun_out/prof python tools/prof_linear_tc.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdetools_b200 as sde  # noqa: E402

eng = sde.default_engine(0)
dev = torch.device("cuda", 0)
for d, scheme, P in ((256, "heun", 148 * 32), (512, "euler", 148 * 32)):
    R = np.random.default_rng(d).standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    params = np.concatenate([A.ravel(order="F"), np.eye(d).ravel(order="F")])
    x0 = np.zeros(d)
    x0[0] = 1.0
    n = 40
    XT = torch.empty(P * d, dtype=torch.float64, device=dev)
    ms = eng.run("linear", scheme, params, np.arange(n + 1) / n, x0, nx=d, nB=d, npaths=P, seed=4, XT=XT, want_time=True)
    fl = (2.0 * d * d + 4 * d) * (2 if scheme == "heun" else 1)
    print(f"d={d} {scheme}: {ms:.3f} ms, {fl * P * n / ms / 1e9:.2f} TFLOP/s")
