"""Analytic transition laws of the reference (validation side) + the exact laws
of the *discretised* schemes.  TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/R/sdetools.R`: `ou.parameters` :706-712, `dOU/pOU/qOU`
:745-778, `cir.parameters` :606-617, `dCIR/pCIR/qCIR` :652-689, `dLinSDE`
:508-572, `lyap` :593-600.  R's nmath `dnorm/pnorm/dchisq/pchisq` are replaced
by scipy.stats (`norm`, `ncx2`); `Matrix::expm` by `scipy.linalg.expm`.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg, stats


# ---- Ornstein-Uhlenbeck ---------------------------------------------------
def ou_parameters(x0, lam, xi, sigma, t):
    """R/sdetools.R:706-712."""
    mean = xi + (x0 - xi) * np.exp(-lam * t)
    sd = sigma * np.sqrt(t) if lam == 0 else sigma * np.sqrt((1 - np.exp(-2 * lam * t)) / 2 / lam)
    return mean, sd


def dOU(x, x0, lam, xi, sigma, t, log=False):
    """R/sdetools.R:745-752."""
    m, s = ou_parameters(x0, lam, xi, sigma, t)
    ld = stats.norm.logpdf(x, m, s)
    return ld if log else np.exp(ld)


def pOU(x, x0, lam, xi, sigma, t, lower_tail=True):
    """R/sdetools.R:758-765."""
    m, s = ou_parameters(x0, lam, xi, sigma, t)
    return stats.norm.cdf(x, m, s) if lower_tail else stats.norm.sf(x, m, s)


def qOU(ps, x0, lam, xi, sigma, t):
    """R/sdetools.R:771-778."""
    m, s = ou_parameters(x0, lam, xi, sigma, t)
    return stats.norm.ppf(ps, m, s)


def rOU(n, x0, lam, xi, sigma, t, rng=None, z=None):
    """R/sdetools.R:784-791: rnorm(n, mean, sd) = mean + sd*Z (z: supplied standard normals)."""
    m, s = ou_parameters(np.asarray(x0, dtype=float), lam, xi, sigma, t)
    z = (rng or np.random.default_rng()).standard_normal(n) if z is None else np.asarray(z, dtype=float)
    return m + s * z


# ---- Cox-Ingersoll-Ross ---------------------------------------------------
def cir_parameters(x0, lam, xi, gamma, t, Stratonovich=False):
    """R/sdetools.R:606-617."""
    if Stratonovich:
        xi = xi + gamma ** 2 / 4 / lam
    c = 2 * lam / gamma ** 2 / (1 - np.exp(-lam * t))
    nu = 2 * c * x0 * np.exp(-lam * t)
    n = 4 * lam * xi / gamma ** 2
    return c, nu, n


def dCIR(x, x0, lam, xi, gamma, t, Stratonovich=False, log=False):
    """R/sdetools.R:652-663: log(2c) + dchisq(2c x, n, nu, log=TRUE)."""
    c, nu, n = cir_parameters(x0, lam, xi, gamma, t, Stratonovich)
    if np.isinf(t):       # dchisq with ncp=0 is the central law; scipy's ncx2 needs nc>0
        ld = np.log(2 * c) + stats.chi2.logpdf(2 * c * np.asarray(x, dtype=float), n)
    else:
        ld = np.log(2 * c) + stats.ncx2.logpdf(2 * c * np.asarray(x, dtype=float), n, nu)
    return ld if log else np.exp(ld)


def pCIR(x, x0, lam, xi, gamma, t, Stratonovich=False, lower_tail=True):
    """R/sdetools.R:669-676."""
    c, nu, n = cir_parameters(x0, lam, xi, gamma, t, Stratonovich)
    d = stats.chi2(n) if np.isinf(t) else stats.ncx2(n, nu)
    q = 2 * c * np.asarray(x, dtype=float)
    return d.cdf(q) if lower_tail else d.sf(q)


def qCIR(ps, x0, lam, xi, gamma, t, Stratonovich=False):
    """R/sdetools.R:682-689."""
    c, nu, n = cir_parameters(x0, lam, xi, gamma, t, Stratonovich)
    d = stats.chi2(n) if np.isinf(t) else stats.ncx2(n, nu)
    return d.ppf(ps) / 2 / c


def rCIR(n, x0, lam, xi, gamma, t, Stratonovich=False, rng=None):
    """R/sdetools.R:695-702: rchisq(n, df, ncp)/2/c; R's rchisq with ncp is the Poisson mixture
    rpois(ncp/2) -> rchisq(2 r) + rgamma(df/2, scale 2) (nmath/rnchisq.c), restated with numpy draws."""
    rng = rng or np.random.default_rng()
    c, nu, df = cir_parameters(np.broadcast_to(np.asarray(x0, dtype=float), (n,)), lam, xi, gamma, t, Stratonovich)
    r = rng.poisson(nu / 2).astype(float)
    chi = np.where(r > 0, 2.0 * rng.standard_gamma(np.maximum(r, 1e-300)), 0.0)
    if df > 0:
        chi = chi + 2.0 * rng.standard_gamma(df / 2, n)
    return chi / 2 / c


# ---- linear SDE -----------------------------------------------------------
def lyap(A, Q):
    """R/sdetools.R:593-600: P = kron(I,A)+kron(A,I); X = -solve(P, vec(Q)).
    Dense n^2 x n^2 solve; raises LinAlgError when P is singular (the reference
    relies on that error in dLinSDE's tryCatch)."""
    A = np.atleast_2d(np.asarray(A, dtype=float))
    Q = np.atleast_2d(np.asarray(Q, dtype=float))
    n = A.shape[0]
    I = np.eye(n)
    P = np.kron(I, A) + np.kron(A, I)
    if np.linalg.cond(P) > 1 / np.finfo(float).eps:      # R's solve() errors on computational singularity
        raise np.linalg.LinAlgError("system is computationally singular")
    X = -np.linalg.solve(P, Q.reshape(-1, order="F"))
    return X.reshape(n, n, order="F")


def dLinSDE(A, G, t, x0=None, u=None, S0=None):
    """R/sdetools.R:508-572 including the Hamiltonian fallback (:525-534) and the
    zero-/first-order-hold expectations (:544-571)."""
    A = np.atleast_2d(np.asarray(A, dtype=float))
    G = np.asarray(G, dtype=float)
    G = G.reshape(A.shape[0], -1)
    GG = G @ G.T
    nx = A.shape[0]
    S0 = np.zeros_like(A) if S0 is None else np.atleast_2d(np.asarray(S0, dtype=float))
    eAt = linalg.expm(A * t)
    try:
        Sinf = lyap(A, GG)
        St = Sinf - eAt @ (Sinf - S0) @ eAt.T
    except np.linalg.LinAlgError:
        O, I = np.zeros((nx, nx)), np.eye(nx)
        H = np.block([[-A.T, O], [GG, A]])
        XY = linalg.expm(H * t) @ np.vstack([I, S0])
        St = XY[nx:, :] @ np.linalg.inv(XY[:nx, :])
    if x0 is None:
        return {"eAt": eAt, "St": St}
    x0 = np.atleast_1d(np.asarray(x0, dtype=float))
    if u is None:
        return {"EX": eAt @ x0, "St": St}
    u = np.asarray(u, dtype=float)
    if u.size == nx:                                        # zero-order hold :545-553
        AA = np.vstack([np.hstack([A, np.eye(nx)]), np.zeros((nx, 2 * nx))])
        Exu = linalg.expm(AA * t) @ np.concatenate([x0, u.ravel()])
        return {"EX": Exu[:nx], "St": St}
    u = u.reshape(nx, 2)                                    # first-order hold :559-571
    O, I = np.zeros((nx, nx)), np.eye(nx)
    AAA = np.vstack([np.hstack([O, O, O]), np.hstack([I, O, O]), np.hstack([O, I, A])])
    Exuu = linalg.expm(AAA * t) @ np.concatenate([(u[:, 1] - u[:, 0]) / t, u[:, 0], x0])
    return {"EX": Exuu[-nx:], "St": St}


# ---- exact laws of the discretised schemes -------------
def ou_discrete_law(scheme, x0, lam, xi, sigma, h, n):
    """Terminal law of n uniform steps of size h for OU: exactly Gaussian.
    euler (:453): a = 1-lam h, var = sigma^2 h (1-a^2n)/(1-a^2).
    heun (:318,:323, additive noise): b = 1-lam h+(lam h)^2/2, c = sigma (1-lam h/2)."""
    if scheme == "euler":
        a, c = 1 - lam * h, sigma
    else:
        a, c = 1 - lam * h + 0.5 * (lam * h) ** 2, sigma * (1 - lam * h / 2)
    mean = xi + (x0 - xi) * a ** n
    var = c * c * h * (1 - a ** (2 * n)) / (1 - a * a) if a != 1 else c * c * h * n
    return mean, np.sqrt(var)


def linear_discrete_law(A, G, x0, h, n):
    """Euler on dX = AX dt + G dB: m <- (I+Ah) m ; S <- (I+Ah) S (I+Ah)^T + G G^T h."""
    A = np.atleast_2d(np.asarray(A, dtype=float))
    G = np.asarray(G, dtype=float).reshape(A.shape[0], -1)
    M = np.eye(A.shape[0]) + A * h
    m = np.asarray(x0, dtype=float).copy()
    S = np.zeros_like(A)
    Q = G @ G.T * h
    for _ in range(n):
        m = M @ m
        S = M @ S @ M.T + Q
    return m, S
