#!/usr/bin/env python
"""Times the trajectory kernel on config C3 (van der Pol, full trajectory, fused generator) for one build:
    TRAJ_WARPS=6 python tools/tune_traj.py [vdp|ou|cir]
Path count = whole waves of (148 SMs x TRAJ_WARPS warps x 32 paths), ~24 GB of trajectory for vdp."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdetools_b200 as sde  # noqa: E402

eng = sde.default_engine(0)
dev = torch.device("cuda", 0)
warps = int(os.environ.get("TRAJ_WARPS", "6"))
which = sys.argv[1] if len(sys.argv) > 1 else "vdp"
nt = 10001
cfg = {"vdp": ("vdp", "euler", [1.0, 0.5], [0.0, 0.0], 2, "identity", 0.01),
       "ou": ("ou", "euler", [1.3, 0.7, 0.9], [1.0], 1, "identity", 1e-3),
       "cir": ("cir", "heun", [1.0, 1.0, 1.0], [1.0], 1, "abs", 1e-3)}[which]
model, scheme, params, x0, nx, proj, h = cfg
wave = eng.info["sm_count"] * warps * 32
P = wave * max(1, round(151552 * 2 / nx / wave))
t = np.arange(nt) * h
X = torch.empty(P * nx * nt, dtype=torch.float64, device=dev)
run = lambda: eng.run(model, scheme, params, t, x0, nx=nx, npaths=P, projection=proj, seed=3, X=X, want_time=True)
run()
ms = min(run() for _ in range(3))
steps = P * (nt - 1)
print(f"{which} warps={warps} P={P}: {ms:.3f} ms  {steps / ms * 1e3:.4g} path-steps/s  {8.0 * nx * steps / ms / 1e6:.0f} GB/s  "
      f"frac {8.0 * nx * steps / ms / 1e6 / 6549.1:.3f}", flush=True)
