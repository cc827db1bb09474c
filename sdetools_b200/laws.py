"""Transition laws of linear SDEs at the dimension the device integrates.

Host-side mirror of `dLinSDE` (R/sdetools.R:508-572) and `lyap` (:593-600) with the same arguments and
return values, over `sdegpu_lyap` / `sdegpu_dlinsde` of libsdegpu.so (csrc/linlaws.cu): O(d^3) and GEMM-shaped
(Pade scaling-and-squaring, bilinear transform + doubling), where the reference solves the d^2 x d^2 Kronecker
system (O(d^6), unusable at the d = 256 the DMMA kernel runs).  Where the Lyapunov equation has no solution the
library takes Van Loan's block matrix exponential — the reference's own fallback (:525-534).

`linear_exact` turns the law into a catalogue model whose *Euler* step is the exact transition
X_{k+1} = e^{A dt} X_k + chol(S_dt) Z (vignettes/NoisyOscillator.Rmd:70-80), so euler() on a uniform grid
samples the linear SDE without discretisation error on the kernels that already exist.
"""
from __future__ import annotations

import numpy as np

from . import catalogue as _cat
from ._lib import SdeGpuError
from .engine import default_engine


def _as_matrix(A):
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    if A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    return A


def lyap(A, Q, *, device=None):
    """Solve A X + X A' + Q = 0 (R/sdetools.R:593-600).  Raises LinAlgError where the reference's
    `solve` fails: when the equation is singular."""
    A, Q = _as_matrix(A), _as_matrix(Q)
    if A.shape != Q.shape:
        raise ValueError("A and Q must have the same dimensions")
    try:
        return default_engine(device).lyap(A, Q)
    except SdeGpuError as e:
        if e.code == -1:
            raise np.linalg.LinAlgError(str(e)) from None
        raise


def dLinSDE(A, G, t, x0=None, u=None, S0=None, *, device=None):
    """Transition law of dX = (A X + u) dt + G dB over time t (R/sdetools.R:508-572).
    Returns {"eAt", "St"} or, with x0, {"EX", "St"}; u of length nx is held constant (zero-order hold,
    :545-553), u with two columns is interpolated linearly from u[:,0] to u[:,1] (first-order hold, :559-571)."""
    A = _as_matrix(A)
    d = A.shape[0]
    if x0 is not None and np.size(x0) != d:
        raise ValueError("x0 must have one entry per row of A")
    if u is not None and np.size(u) not in (d, 2 * d):
        raise ValueError("u must have length nx or be nx x 2")
    first, St = default_engine(device).dlinsde(A, G, t, x0, u, None if S0 is None else _as_matrix(S0))
    St = 0.5 * (St + St.T)
    return {"eAt": first, "St": St} if x0 is None else {"EX": first, "St": St}


def linear_exact(A, G, dt):
    """Catalogue pair (f, g) whose Euler step of size dt is the exact transition of dX = A X dt + G dB:
    f(x) = ((e^{A dt} - I)/dt) x,  g = L/sqrt(dt) with L L' = S_dt.  Use with euler() on a grid of constant dt."""
    A = _as_matrix(A)
    d = A.shape[0]
    law = dLinSDE(A, G, dt)
    St = law["St"]
    try:
        L = np.linalg.cholesky(St)
    except np.linalg.LinAlgError:                       # rank-deficient noise: symmetric square root
        w, V = np.linalg.eigh(St)
        L = V * np.sqrt(np.clip(w, 0.0, None))
    off = np.abs(L) < 1e-14 * np.max(np.abs(L))
    L = np.where(off & ~np.eye(d, dtype=bool), 0.0, L)  # rounding-level off-diagonals would only hide a diagonal g
    return _cat.linear((law["eAt"] - np.eye(d)) / dt, L / np.sqrt(dt))
