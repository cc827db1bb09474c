import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import sdetools_b200 as sde
eng = sde.default_engine(0)
dev = torch.device("cuda", 0)
P, n = 1 << 22, 1000
t = np.arange(n + 1) * 1e-3
ps = torch.zeros(5, dtype=torch.float64, device=dev)
for model, scheme, params, proj in (("ou", "euler", [1.3, 0.7, 0.9], "identity"), ("cir", "heun", [1.0, 1.0, 1.0], "abs")):
    for fc in (None, np.sin(2 * t)[:, None]):
        ms = min(eng.run(model, scheme, params, t, [1.0], nx=1, npaths=P, projection=proj, seed=3, power_sums=ps, center=[1.0],
                         drift_forcing=fc, want_time=True) for _ in range(3))
        print(model, scheme, "forcing" if fc is not None else "autonomous", f"{P * n / ms * 1e3:.3e} path-steps/s")
