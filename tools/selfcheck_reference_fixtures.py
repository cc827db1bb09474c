#!/usr/bin/env python
"""Plumbing check of tests/test_reference_fixtures.py where R does not exist.

Writes, with the ORACLE standing in for R, every file tests/golden/make_reference_fixtures.R would write (same names,
same C99-hex format) into a temporary directory and runs the fixture tests against it.  This proves that the tests read
the right files, shapes and layouts and that the CUDA path reproduces them (with -m gpu on a B200) — NOT parity with R:
oracle-vs-oracle comparisons pass by construction.  Parity needs the real outputs (run the R script on a box with R).

    python tools/selfcheck_reference_fixtures.py [-m gpu]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import read_hex, write_hex  # noqa: E402
from oracle import laws  # noqa: E402
from oracle import sde_oracle as O  # noqa: E402
import test_reference_fixtures as T  # noqa: E402


def main():
    out = tempfile.mkdtemp(prefix="reference_R_selfcheck_")
    w = lambda name, a: write_hex(os.path.join(out, name + ".txt"), a)
    t, B, B2 = T.inp("times"), T.inp("B"), T.inp("lin_B")
    nt, P = t.size, B.shape[2]
    for k, (model, params, x0, proj) in enumerate(T.CASES):
        f, g, _, _ = O.catalogue(model, params)
        for scheme in ("euler", "heun"):
            X = np.stack([getattr(O, scheme)(f, g, t, x0, B=B[:, :, p], p=O.projection(proj))["X"].reshape(nt, -1) for p in range(P)], axis=2)
            w(f"X_{scheme}_{k}", X)
    A, G, x0 = T.inp("lin_A"), T.inp("lin_G"), T.inp("lin_x0")
    fl, gl, _, _ = O.catalogue("linear", (A, G))
    cfun, sfun, Fin = T._td()
    for scheme in ("euler", "heun"):
        fn = getattr(O, scheme)
        w(f"lin_X_{scheme}", np.stack([fn(fl, gl, t, x0, B=B2[:, :, p])["X"] for p in range(P)], axis=2))
        w(f"td_X_{scheme}", np.stack([fn(lambda tt, x: 1.0 * (1.0 - x) + cfun(tt), lambda tt, x: sfun(tt) * (1.0 * np.sqrt(np.abs(x))), t, [1.0],
                                         B=B[:, :, p], p=abs)["X"].reshape(nt, 1) for p in range(P)], axis=2))
        w(f"td_lin_X_{scheme}", np.stack([fn(lambda tt, x: A @ x + Fin * np.sin(2 * tt), lambda x: G, t, x0, B=B2[:, :, p])["X"] for p in range(P)], axis=2))
        for what in ("kill", "stop", "both"):
            X, S, tau = T._oracle_killed(scheme, what)
            w(f"kill_X_{scheme}_{what}", X), w(f"kill_S_{scheme}_{what}", S), w(f"kill_tau_{scheme}_{what}", tau)
    S0 = np.diag(np.arange(1, x0.size + 1) / 10)
    w("law_lyap", laws.lyap(A, G @ G.T))
    l0 = laws.dLinSDE(A, G, 0.7)
    w("law_eAt", l0["eAt"]), w("law_St", l0["St"])
    w("law_St_S0", laws.dLinSDE(A, G, 0.7, S0=S0)["St"])
    w("law_EX", np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0)["EX"]))
    w("law_EX_zoh", np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0, u=Fin)["EX"]))
    w("law_EX_foh", np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0, u=np.stack([Fin, 2 * Fin], axis=1))["EX"]))
    f, g = T.inp("int_f"), T.inp("int_g")
    nc = f.shape[0]
    for rule in "lrc":
        w(f"int_stochint_{rule}", np.stack([O.stochint(f[c], g[c], rule) for c in range(nc)]))
    w("int_ito", np.stack([O.itointegral(f[c], g[c]) for c in range(nc)]))
    w("int_qv", np.stack([O.QuadraticVariation(g[c]) for c in range(nc)]))
    w("int_cv", np.stack([O.CrossVariation(f[c], g[c]) for c in range(nc)]))
    z, tb = T.inp("rbm_z"), T.inp("rbm_times")
    w("rbm_B", np.stack([O.rBM(tb, sigma=0.7, B0=0.25, u=-0.1, z=z[:, p]) for p in range(z.shape[1])], axis=1))
    env = dict(os.environ, SDEGPU_REFERENCE_R_DIR=out)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_reference_fixtures.py"), "-q", *sys.argv[1:]], env=env)
    print(f"(oracle stand-in outputs in {out}: a check of the tests' plumbing, not of parity with R)")
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
