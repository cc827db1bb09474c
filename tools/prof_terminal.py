#!/usr/bin/env python
"""One short launch pattern of the bench's dominant kernel for ncu (k_terminal<cir,heun,abs>, uniform grid):
    ncu --set full --clock-control none --import-source on -k regex:k_terminal -s 1 -c 2 -o gpurun_out/prof python tools/prof_terminal.py
The workload is config C2 shortened to 2,000 steps x 148*768*8 paths so that ~40 replays stay short."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdetools_b200 as sde  # noqa: E402

eng = sde.default_engine(0)
dev = torch.device("cuda", 0)
model, scheme, proj = (sys.argv[1:4] + ["cir", "heun", "abs"][len(sys.argv) - 1:])[:3]
nsteps = int(os.environ.get("PROF_STEPS", "2000"))
P = int(os.environ.get("PROF_PATHS", str(148 * 768 * 8)))
t = np.arange(nsteps + 1) * (1.0 / nsteps)
ps = torch.zeros(5, dtype=torch.float64, device=dev)
hc = torch.zeros(259, dtype=torch.int64, device=dev)
params = {"cir": [1.0, 1.0, 1.0], "ou": [1.0, 0.0, 1.0]}[model]
for _ in range(4):
    ms = eng.run(model, scheme, params, t, [1.0], nx=1, npaths=P, projection=proj, seed=7, power_sums=ps, center=[1.0],
                 hist_counts=hc, hist=(0.0, 8.0, 256), want_time=True)
torch.cuda.synchronize()
print(f"{model} {scheme} {proj}: {P} paths x {nsteps} steps, {ms:.3f} ms, {P * nsteps / ms * 1e3:.4g} path-steps/s")
