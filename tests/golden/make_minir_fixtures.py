#!/usr/bin/env python
"""Writes tests/golden/reference_minir/: outputs of the REFERENCE'S OWN SOURCE TEXT, evaluated by oracle/minir.

    python tests/golden/make_minir_fixtures.py [--ref /root/reference] [--c1] [--jobs 8]

GNU R is not installed in the build image.  `oracle/minir` is a small evaluator for the R subset the reference is written in; it
parses the unmodified `/root/reference/R/sdetools.R` and runs `euler`, `heun`, `stochint`, `itointegral`, `QuadraticVariation`,
`CrossVariation`, `rBM`, `lyap`, `dLinSDE` from their own function bodies.  This script

1. runs the SAME recipe a box with GNU R would run (`tests/golden/make_reference_fixtures.R`, written for Rscript, unchanged) under
   that evaluator, on the committed inputs `tests/golden/reference_R_inputs/`, writing `reference_minir/*.txt` (C99 hex doubles);
2. with `--c1`, runs BASELINE config C1 at full size — scalar OU `euler()`, 1,000 paths x 1,000 steps, supplied `B`, both parameter
   sets of SURVEY §8(d) — one `euler()` call per path as the reference's users do (`vignettes/Killing.Rmd:183-189`), and stores the
   SHA-256 of the 8 MB result (reference layout, path-major, time fastest) together with every path's terminal value in
   `reference_minir/c1_ou_euler.json` (about 15 CPU-minutes per parameter set, spread over `--jobs` processes).

`/root/reference` does not exist on the GPU box; these outputs are committed so the `-m gpu` tests can compare the CUDA path with
them there.  tests/test_minir.py re-runs step 1 when the reference is present and asserts the committed files are current.
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

C1_SETS = [(1.0, 0.0, 1.0, 0.0), (1.3, 0.7, 0.9, 2.0)]          # lambda, xi, sigma, x0


def run_recipe(ref, outname="reference_minir", here=HERE):
    from oracle.minir import Interp
    R = Interp(args=[ref, here, outname])
    R.script_file = os.path.join(HERE, "make_reference_fixtures.R")
    R.out = open(os.devnull, "w")
    R.source(R.script_file)
    return os.path.join(here, outname)


def c1_inputs():
    """C1's grid and Brownian paths exactly as tests/test_gpu_parity.py builds them: times = seq(0, 10, 0.01), B = rBM's
    cumsum(sqrt(dt) * Z) with the 80-bit accumulator, Z from PCG64(20261019) drawn path-major, B[1] = 0."""
    from oracle import sde_oracle as O
    rng = np.random.default_rng(20261019)
    t = O.r_seq(0, 10, 0.01)
    nt, P = t.size, 1000
    z = rng.standard_normal((P, nt)).T
    z[0] = 0.0
    B = O.r_cumsum_cols(np.sqrt(np.concatenate(([0.0], O.r_diff(t))))[:, None] * z)
    return t, B, z


def _c1_chunk(job):
    ref, lam, xi, sig, x0, lo, hi = job
    from oracle.minir import Interp, from_numpy, to_numpy
    R = Interp()
    R.source(os.path.join(ref, "R", "sdetools.R"))
    # the catalogue's OU closures (r/R/sdegpu.R): f(x) = lambda*(xi - x), g(x) = sigma
    R.run(f"f <- function(x) {lam!r} * ({xi!r} - x); g <- function(x) {sig!r}")
    f, g = R.globalenv.lookup("f"), R.globalenv.lookup("g")
    t, B, _ = c1_inputs()
    out = np.empty((hi - lo, t.size))
    for p in range(lo, hi):
        sol = R.call("euler", f, g, from_numpy(t), from_numpy([x0]), B=from_numpy(B[:, p]))
        out[p - lo] = to_numpy(sol)["X"].reshape(-1)
    return lo, out


def run_c1(ref, jobs, outdir):
    import multiprocessing as mp
    from oracle.minir import Interp, from_numpy, to_numpy
    t, B, z = c1_inputs()
    # the reference's own rBM (R/sdetools.R:16-22) with the recorded normals reproduces B: times[1] = 0 makes the first increment 0
    R = Interp()
    R.source(os.path.join(ref, "R", "sdetools.R"))
    R.run(".z <- NULL; rnorm <- function(n, mean = 0, sd = 1) mean + sd * .z[seq_len(n)]")
    for p in (0, 1, 500, 999):
        R.globalenv.set(".z", from_numpy(z[:, p]))
        assert np.array_equal(to_numpy(R.call("rBM", from_numpy(t))), B[:, p]), "rBM under minir != test construction of B"
    res = {"source": "R/sdetools.R:375-470 (euler) evaluated by oracle/minir, one call per path", "nt": int(t.size), "npaths": 1000,
           "layout": "sha256 over float64 little-endian bytes of X[p, 0, i] (path-major, time fastest)", "sets": []}
    with mp.Pool(jobs) as pool:
        for lam, xi, sig, x0 in C1_SETS:
            step = 1000 // (jobs * 4) or 1
            parts = pool.map(_c1_chunk, [(ref, lam, xi, sig, x0, lo, min(lo + step, 1000)) for lo in range(0, 1000, step)])
            X = np.concatenate([o for _, o in sorted(parts, key=lambda q: q[0])], axis=0)           # (P, nt)
            res["sets"].append({"lambda": lam, "xi": xi, "sigma": sig, "x0": x0,
                                "sha256_X": hashlib.sha256(np.ascontiguousarray(X).astype("<f8").tobytes()).hexdigest(),
                                "XT": [float(v).hex() for v in X[:, -1]],
                                "X_path7_every50": [float(v).hex() for v in X[7, ::50]]})
            print("C1", (lam, xi, sig, x0), res["sets"][-1]["sha256_X"], flush=True)
    with open(os.path.join(outdir, "c1_ou_euler.json"), "w") as fh:
        json.dump(res, fh, indent=0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("SDETOOLS_REF", "/root/reference"))
    ap.add_argument("--c1", action="store_true")
    ap.add_argument("--jobs", type=int, default=8)
    a = ap.parse_args()
    out = run_recipe(a.ref)
    print("wrote", len(os.listdir(out)), "files to", out)
    if a.c1:
        run_c1(a.ref, a.jobs, out)
