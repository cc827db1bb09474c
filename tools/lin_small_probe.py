import sys, numpy as np, torch
sys.path.insert(0, '.')
import sdetools_b200 as sde
eng = sde.default_engine(0)
dev = torch.device('cuda', 0)
for d, nB, P, n in ((2, 2, 1 << 20, 1000), (4, 2, 1 << 19, 1000), (8, 8, 1 << 18, 1000), (16, 16, 1 << 17, 500)):
    rng = np.random.default_rng(d)
    A = -0.5 * np.eye(d) + 0.5 * (rng.standard_normal((d, d)) - rng.standard_normal((d, d)).T) / np.sqrt(d)
    G = rng.standard_normal((d, nB)) * 0.5
    params = np.concatenate([A.ravel(order='F'), G.ravel(order='F')])
    XT = torch.empty(P * d, dtype=torch.float64, device=dev)
    t = np.arange(n + 1) / n
    for scheme in ('euler', 'heun'):
        eng.run('linear', scheme, params, t, np.ones(d), nx=d, nB=nB, npaths=P, seed=1, XT=XT, want_time=True)
        ms = min(eng.run('linear', scheme, params, t, np.ones(d), nx=d, nB=nB, npaths=P, seed=1, XT=XT, want_time=True) for _ in range(2))
        print(d, nB, scheme, P, n, round(ms, 2), 'ms', '%.3g path-steps/s' % (P * n / ms * 1e3), flush=True)
