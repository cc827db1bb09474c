#!/usr/bin/env python
"""End-to-end (host buffers in, host buffers out) rates of the I/O-heavy configurations through the public API:
C1 (OU euler, 1000 x 1000, supplied B) and C3 (van der Pol, full trajectory), next to the kernel-only time the
library reports.  One JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdetools_b200 as sde  # noqa: E402

eng = sde.default_engine(0)


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def c1():
    t = np.arange(1001) * 0.01
    rng = np.random.default_rng(1)
    B = np.asfortranarray(np.cumsum(np.concatenate([np.zeros((1, 1, 1000)), 0.1 * rng.standard_normal((1000, 1, 1000))]), axis=0))
    X = np.empty((1001, 1, 1000), order="F")
    ms = [0.0]

    def run():
        ms[0] = eng.run("ou", "euler", [1.0, 0.0, 1.0], t, [0.0], nx=1, npaths=1000, B=B, X=X, want_time=True)
    s = timed(run, 20)
    return {"config": "C1 OU euler 1000 paths x 1000 steps, supplied B", "e2e_s": s, "kernel_ms": ms[0], "path_steps_per_s": 1e6 / s,
            "h2d_bytes": B.nbytes, "d2h_bytes": X.nbytes, "gb_per_s": (B.nbytes + X.nbytes) / s / 1e9}


def c3(P, supplied_b=False):
    t = np.arange(10001) * 0.01
    X = np.empty((10001, 2, P), order="F")          # untouched pages: the first pass pays the page faults, like a fresh R result
    B = None
    if supplied_b:
        B = np.empty((10001, 1, P), order="F")
        eng.rvbm(t, P, B.reshape(10001, P, order="F"), seed=5)
    ms = [0.0]

    def run():
        ms[0] = eng.run("vdp", "euler", [1.0, 0.5], t, [0.0, 0.0], nx=2, npaths=P, seed=3, X=X, B=B, want_time=True)
    t0 = time.perf_counter()
    run()
    first = time.perf_counter() - t0
    s = timed(run, 2)
    return {"config": f"C3 van der Pol euler {P} paths x 10000 steps, full trajectory" + (", supplied B" if supplied_b else ""),
            "e2e_s": s, "e2e_first_touch_s": first, "kernel_ms": ms[0], "path_steps_per_s": P * 1e4 / s, "h2d_bytes": 0 if B is None else B.nbytes,
            "d2h_bytes": X.nbytes, "gb_per_s": (X.nbytes + (0 if B is None else B.nbytes)) / s / 1e9}


def pcie_peak(nbytes=1 << 30, reps=5):
    """The denominators: one pinned buffer of 1 GiB moved by the driver, each direction alone and both at once."""
    import torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}

    def bw(fn, moved):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return moved * reps / (time.perf_counter() - t0) / 1e9

    out["d2h_pinned_gb_per_s"] = bw(lambda: host.copy_(dev, non_blocking=True), nbytes)
    out["h2d_pinned_gb_per_s"] = bw(lambda: dev.copy_(host, non_blocking=True), nbytes)

    def both():
        with torch.cuda.stream(s1):
            host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s2):
            dev2.copy_(host2, non_blocking=True)
    out["bidirectional_pinned_gb_per_s"] = bw(both, 2 * nbytes)
    out["host_threads"] = os.cpu_count()
    return {"config": "PCIe peak, pinned 1 GiB copies by the driver", **out}


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
    if os.environ.get("E2E_PCIE_PEAK", "1") != "0":
        print(json.dumps(pcie_peak()))
    print(json.dumps(c1()))
    print(json.dumps(c3(P)))
    print(json.dumps(c3(P, True)))
