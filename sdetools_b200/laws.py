"""Transition laws of linear SDEs at the dimension the device integrates.

Host-side mirror of `dLinSDE` (R/sdetools.R:508-572) and `lyap` (:593-600) with the same arguments and
return values, but O(d^3): the reference solves the Lyapunov equation through the d^2 x d^2 Kronecker
system (O(d^6), unusable at the d = 256 the DMMA kernel runs), here it is Bartels-Stewart (LAPACK via
scipy) and, where the Lyapunov equation is singular, Van Loan's block matrix exponential — the same
fallback the reference takes (:525-534), reached by the same condition (lambda_i + lambda_j = 0).

`linear_exact` turns the law into a catalogue model whose *Euler* step is the exact transition
X_{k+1} = e^{A dt} X_k + chol(S_dt) Z (vignettes/NoisyOscillator.Rmd:70-80), so euler() on a uniform grid
samples the linear SDE without discretisation error on the kernels that already exist.

This is host logic of the product (numpy/scipy), not part of the test oracle.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg

from . import catalogue as _cat


def _as_matrix(A):
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    if A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    return A


def _lyapunov_is_singular(A) -> bool:
    """kron(I,A)+kron(A,I) is singular iff lambda_i + lambda_j = 0 for some pair of eigenvalues."""
    lam = linalg.eigvals(A)
    s = np.abs(lam[:, None] + lam[None, :])
    scale = max(1.0, float(np.max(np.abs(lam))))
    return bool(np.min(s) <= 1e-12 * scale)


def lyap(A, Q):
    """Solve A X + X A' + Q = 0 (R/sdetools.R:593-600).  Raises LinAlgError where the reference's
    `solve` fails: when the equation is singular."""
    A, Q = _as_matrix(A), _as_matrix(Q)
    if _lyapunov_is_singular(A):
        raise np.linalg.LinAlgError("Lyapunov equation is singular (lambda_i + lambda_j = 0)")
    return linalg.solve_continuous_lyapunov(A, -Q)


def _van_loan(A, GG, S0, t):
    """S_t = e^{At} S0 e^{A't} + int_0^t e^{As} GG' e^{A's} ds from one 2d x 2d matrix exponential."""
    d = A.shape[0]
    M = np.zeros((2 * d, 2 * d))
    M[:d, :d], M[:d, d:], M[d:, d:] = -A, GG, A.T
    E = linalg.expm(M * t)
    eAt = E[d:, d:].T
    return eAt @ E[:d, d:] + eAt @ S0 @ eAt.T


def dLinSDE(A, G, t, x0=None, u=None, S0=None):
    """Transition law of dX = (A X + u) dt + G dB over time t (R/sdetools.R:508-572).
    Returns {"eAt", "St"} or, with x0, {"EX", "St"}; u of length nx is held constant (zero-order hold,
    :545-553), u with two columns is interpolated linearly from u[:,0] to u[:,1] (first-order hold, :559-571)."""
    A = _as_matrix(A)
    d = A.shape[0]
    G = np.asarray(G, dtype=np.float64).reshape(d, -1)
    GG = G @ G.T
    S0 = np.zeros((d, d)) if S0 is None else _as_matrix(S0)
    eAt = linalg.expm(A * t)
    if _lyapunov_is_singular(A):
        St = _van_loan(A, GG, S0, t)
    else:
        Sinf = linalg.solve_continuous_lyapunov(A, -GG)
        St = Sinf - eAt @ (Sinf - S0) @ eAt.T
    St = 0.5 * (St + St.T)
    if x0 is None:
        return {"eAt": eAt, "St": St}
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
    if u is None:
        return {"EX": eAt @ x0, "St": St}
    u = np.asarray(u, dtype=np.float64)
    O, I = np.zeros((d, d)), np.eye(d)
    if u.size == d:
        AA = np.block([[A, I], [O, O]])
        return {"EX": (linalg.expm(AA * t) @ np.concatenate([x0, u.ravel()]))[:d], "St": St}
    u = u.reshape(d, 2)
    AAA = np.block([[O, O, O], [I, O, O], [O, I, A]])
    return {"EX": (linalg.expm(AAA * t) @ np.concatenate([(u[:, 1] - u[:, 0]) / t, u[:, 0], x0]))[-d:], "St": St}


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
