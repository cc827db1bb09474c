#!/usr/bin/env python
"""Writes the fixtures in this directory.

    python tests/golden/make_golden.py

Two files, two different kinds of evidence:

* reference_known_answers.json — every number the reference's OWN tests and documentation pin on or
  next to the sample-path hot path (tests/testthat/test_solvers.R:3-36, R/sdetools.R:86-102,:587-590,
  vignettes/CoxIngersollRoss.Rmd:56-71) plus the Random123 Philox4x32-10 known-answer vectors.
  These are expectations copied as data, each with its source line.
* oracle_vectors.npz — seeded inputs and the outputs of the numpy restatement of the reference loop
  (oracle/sde_oracle.py), one path per call exactly as R would run it. R is not installed in the image,
  so these are NOT outputs of R: they freeze the oracle (a later edit that changes a single bit fails
  tests/test_golden.py) and give the CUDA path a fixed target that does not depend on the oracle being
  importable on the GPU box. Parity against R itself stays "unpinned" (DESIGN.md §2).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import philox_ref                      # noqa: E402
from oracle import sde_oracle as O                 # noqa: E402

CASES = [("ou", [1.3, 0.7, 0.9], [2.0], "identity"), ("cir", [1.0, 1.0, 1.0], [1.0], "abs"),
         ("cir", [0.5, 0.2, 1.5], [0.05], "relu"), ("vdp", [1.0, 0.5], [0.1, -0.2], "identity"),
         ("logistic", [0.5, 1.0], [0.3], "identity"), ("gbm", [0.4, 1.0], [1.0], "abs"),
         ("doublewell", [0.5], [0.2], "identity")]


def known_answers():
    return {
        "dims": {"source": "tests/testthat/test_solvers.R:3-18", "times": list(range(11)), "nx": 3, "nB": [1, 2],
                 "dim_X": [11, 3]},
        "dLinSDE": {"source": "tests/testthat/test_solvers.R:20-36", "cases": [
            {"A": [[-1.0]], "G": [[1.0]], "t": 1.0, "S0": [[0.5]], "eAt": [[float(np.exp(-1.0))]], "St": [[0.5]]},
            {"A": [[0.0, 1.0], [-1.0, 0.0]], "G": [[1.0, 0.0], [0.0, 1.0]], "t": float(2 * np.pi), "x0": [1.0, 0.0],
             "EX": [1.0, 0.0], "St": [[float(2 * np.pi), 0.0], [0.0, float(2 * np.pi)]]},
            {"A": [[-1.0]], "G": [[1.0]], "t": 2.0, "x0": [1.0], "u": [[1.0]], "EX": [1.0]},
            {"A": [[0.0]], "G": [[1.0]], "t": 2.0, "x0": [1.0], "u": [[1.0, 3.0]], "EX": [5.0]}]},
        "lyap": {"source": "R/sdetools.R:587-590", "cases": [
            {"A": [[-1.0]], "Q": [[1.0]], "P": [[0.5]]}]},
        "philox4x32_10": {"source": "Random123 kat_vectors (SURVEY.md Appendix F)", "cases": [
            {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
            {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
            {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
             "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]}]},
    }


def oracle_vectors():
    out = {}
    rng = np.random.default_rng(20261020)
    # a non-uniform grid (dt_i is never assumed constant, R/sdetools.R:446) with an odd number of points
    t = np.concatenate(([0.0], np.cumsum(rng.uniform(0.004, 0.02, 40))))
    P = 5
    B = np.stack([O.rvBM(t, 1, rng=rng) for _ in range(P)], axis=2)            # (nt, 1, P)
    out["times"], out["B"] = t, B
    for k, (model, params, x0, proj) in enumerate(CASES):
        f, g, _, _ = O.catalogue(model, params)
        for scheme, fn in (("euler", O.euler), ("heun", O.heun)):
            X = np.stack([fn(f, g, t, x0, B=B[:, :, p], p=O.projection(proj))["X"] for p in range(P)], axis=2)
            out[f"X_{scheme}_{k}"] = X
    # linear system, d=3, two noise channels
    d, nB = 3, 2
    A = -np.eye(d) + 0.3 * rng.standard_normal((d, d))
    G = rng.standard_normal((d, nB))
    x0 = rng.standard_normal(d)
    B2 = np.stack([O.rvBM(t, nB, rng=rng) for _ in range(P)], axis=2)
    out["lin_A"], out["lin_G"], out["lin_x0"], out["lin_B"] = A, G, x0, B2
    f, g, _, _ = O.catalogue("linear", (A, G))
    for scheme, fn in (("euler", O.euler), ("heun", O.heun)):
        out[f"lin_X_{scheme}"] = np.stack([fn(f, g, t, x0, B=B2[:, :, p])["X"] for p in range(P)], axis=2)
    # integrals on 4 columns of 57 points (sums chosen to cancel heavily so the 80-bit accumulator matters)
    n, nc = 57, 4
    fcol = rng.standard_normal((nc, n)) * 10.0 ** rng.integers(-6, 7, (nc, n))
    gcol = (rng.standard_normal((nc, n)) * 10.0 ** rng.integers(-3, 4, (nc, n))).cumsum(axis=1)
    out["int_f"], out["int_g"] = fcol, gcol
    for rule in "lrc":
        out[f"int_stochint_{rule}"] = np.stack([O.stochint(fcol[c], gcol[c], rule) for c in range(nc)])
    out["int_ito"] = np.stack([O.itointegral(fcol[c], gcol[c]) for c in range(nc)])
    out["int_qv"] = np.stack([O.QuadraticVariation(gcol[c]) for c in range(nc)])
    out["int_cv"] = np.stack([O.CrossVariation(fcol[c], gcol[c]) for c in range(nc)])
    # the fused generator: normals of paths 0..3 and 2^40+7, 13 steps, channel 0
    paths = np.array([0, 1, 2, 3, (1 << 40) + 7], dtype=np.uint64)
    out["gen_paths"], out["gen_seed"] = paths, np.array([0x5DE70015], dtype=np.uint64)
    out["gen_z"] = philox_ref.normals(paths, 13, 0, 0x5DE70015)
    return out


def write_hex(path, a):
    """C99 hexadecimal doubles, one per line, column-major, after a line with the dimensions: what
    make_reference_fixtures.R reads with as.numeric() and writes back with sprintf("%a")."""
    a = np.asarray(a, dtype=np.float64)
    with open(path, "w") as fh:
        fh.write(" ".join(str(d) for d in (a.shape if a.ndim else (1,))) + "\n")
        fh.write("\n".join(float(v).hex() for v in a.ravel(order="F")) + "\n")


def read_hex(path):
    with open(path) as fh:
        dims = [int(d) for d in fh.readline().split()]
        vals = [float("nan") if tok in ("NA", "NaN") else (float("inf") if tok == "Inf" else (float("-inf") if tok == "-Inf" else float.fromhex(tok)))
                for tok in fh.read().split()]
    return np.array(vals).reshape(dims, order="F")


def reference_inputs(vec):
    """Inputs of tests/golden/make_reference_fixtures.R (run where R exists; see that file)."""
    d = os.path.join(HERE, "reference_R_inputs")
    os.makedirs(d, exist_ok=True)
    for k in ("times", "B", "lin_A", "lin_G", "lin_x0", "lin_B", "int_f", "int_g"):
        write_hex(os.path.join(d, k + ".txt"), vec[k])
    rng = np.random.default_rng(20261021)                  # its own stream: oracle_vectors.npz does not move
    P = vec["B"].shape[2]
    write_hex(os.path.join(d, "Stau.txt"), rng.uniform(0.5, 0.95, P))
    tb = np.concatenate(([0.25], 0.25 + np.cumsum(rng.uniform(0.01, 0.05, 30))))   # times[1] > 0: first increment spans [0, t1]
    write_hex(os.path.join(d, "rbm_times.txt"), tb)
    write_hex(os.path.join(d, "rbm_z.txt"), rng.standard_normal((tb.size, 3)))


def main():
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as fh:
        json.dump(known_answers(), fh, indent=1)
    vec = oracle_vectors()
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vec)
    reference_inputs(vec)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
