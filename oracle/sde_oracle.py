"""numpy restatement of the reference sample-path layer.  TEST INFRASTRUCTURE ONLY.

Every function cites the lines of `/root/reference/R/sdetools.R` it follows.
R semantics reproduced on purpose:
  * every binary operator is one correctly rounded IEEE-double operation, no
    fused multiply-add, left-associative; `%*%` binds tighter than `*`;
  * `cumsum` accumulates in x87 80-bit `long double` and rounds each partial
    sum to double (R `src/main/cum.c`); numpy `longdouble` is that type here;
  * matrices are column-major; `X` is `nt x nx` (time fastest).

PARITY (see oracle/__init__.py): no numeric golden vector exists in the reference
for these functions; this restatement is pinned bit for bit to the reference's own
source executed by oracle/minir (tests/golden/reference_minir/, tests/test_minir.py,
tests/test_reference_fixtures.py) and by the identities in tests/test_oracle.py.
"""
from __future__ import annotations

import inspect

import numpy as np

assert np.finfo(np.longdouble).nmant == 63, "oracle needs x87 80-bit long double"


# --------------------------------------------------------------------------
# base-R primitives the path uses
# --------------------------------------------------------------------------
def r_seq(frm: float, to: float, by: float) -> np.ndarray:
    """R `seq(from,to,by)` = from + (0:n)*by with n = floor((to-from)/by + 1e-10)
    (R `seq.default`); grids built this way differ by ulps from step to step,
    which is why `dt` is never treated as a constant."""
    n = int(np.floor((to - frm) / by + 1e-10))
    return frm + np.arange(n + 1, dtype=np.float64) * by


def r_cumsum(x) -> np.ndarray:
    """R `cumsum` on doubles: `LDOUBLE sum=0; sum += x[i]; s[i]=(double)sum`."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    acc = np.longdouble(0.0)
    for i in range(x.shape[0]):
        acc = acc + np.longdouble(x[i])
        out[i] = np.float64(acc)
    return out


def r_cumsum_cols(x) -> np.ndarray:
    """`cumsum` applied to every column of an (n, ncols) array at once (same
    80-bit accumulator per column; vectorised over columns only)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    acc = np.zeros(x.shape[1:], dtype=np.longdouble)
    for i in range(x.shape[0]):
        acc = acc + x[i].astype(np.longdouble)
        out[i] = acc.astype(np.float64)
    return out


def r_diff(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return x[1:] - x[:-1]


# --------------------------------------------------------------------------
# Brownian motion generators  (R/sdetools.R:16-73)
# --------------------------------------------------------------------------
def rBM(times, sigma=1.0, B0=0.0, u=0.0, rng=None, z=None) -> np.ndarray:
    """R/sdetools.R:16-22.  `dt <- c(times[1],diff(times))`;
    `dB <- rnorm(n, mean=u*dt, sd=sigma*sqrt(dt))`; `B <- B0+cumsum(dB)`.
    R's Mersenne-Twister/inversion stream cannot be reproduced without R, so the
    standard normals come from `z` (if given) or a numpy Generator."""
    times = np.asarray(times, dtype=np.float64)
    dt = np.concatenate(([times[0]], r_diff(times)))
    if z is None:
        rng = np.random.default_rng() if rng is None else rng
        z = rng.standard_normal(times.shape[0])
    dB = u * dt + (sigma * np.sqrt(dt)) * np.asarray(z, dtype=np.float64)
    return B0 + r_cumsum(dB)


def rvBM(times, n=1, sigma=None, B0=None, u=None, rng=None, z=None) -> np.ndarray:
    """R/sdetools.R:40-46: `sapply(1:n, rBM)` -> `nt x n`, one column per channel."""
    def rep(v, d):
        v = np.full(n, d, dtype=np.float64) if v is None else np.atleast_1d(np.asarray(v, dtype=np.float64))
        return np.repeat(v, n) if v.shape[0] == 1 and n > 1 else v
    sigma, B0, u = rep(sigma, 1.0), rep(B0, 0.0), rep(u, 0.0)
    rng = np.random.default_rng() if rng is None else rng
    cols = [rBM(times, sigma[i], B0[i], u[i], rng=rng, z=None if z is None else z[:, i]) for i in range(n)]
    return np.stack(cols, axis=1)


def rBrownianBridge(times, B0T=(0.0, 0.0), sigma=1.0, rng=None, z=None) -> np.ndarray:
    """R/sdetools.R:62-73."""
    times = np.asarray(times, dtype=np.float64)
    B = rBM(times, sigma=sigma, rng=rng, z=z)
    n = times.shape[0]
    slope = (B0T[1] - B0T[0] - B[n - 1] + B[0]) / (times[n - 1] - times[0])
    return B + B0T[0] - B[0] + slope * (times - times[0])


# --------------------------------------------------------------------------
# stochastic integrals on a stored path  (R/sdetools.R:105-187)
# --------------------------------------------------------------------------
def stochint(f, g, rule="l"):
    """R/sdetools.R:105-113.  Returns a vector for one rule, a dict of columns
    (in `rule` order, like the data.frame) for several."""
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    dg = r_diff(g)
    l = np.concatenate(([0.0], r_cumsum(f[:-1] * dg)))
    r = np.concatenate(([0.0], r_cumsum(f[1:] * dg)))
    c = 0.5 * (l + r)
    cols = {"l": l, "r": r, "c": c}
    if isinstance(rule, str):
        return cols[rule]
    return {k: cols[k] for k in rule}


def QuadraticVariation(X):
    """R/sdetools.R:134-137: c(0,cumsum(diff(X)^2)); R special-cases ^2 to x*x."""
    d = r_diff(X)
    return np.concatenate(([0.0], r_cumsum(d * d)))


def itointegral(G, B):
    """R/sdetools.R:155-158: c(0,cumsum(G[-length(G)]*diff(B)))."""
    G = np.asarray(G, dtype=np.float64)
    return np.concatenate(([0.0], r_cumsum(G[:-1] * r_diff(B))))


def CrossVariation(X, Y):
    """R/sdetools.R:184-187: c(0,cumsum(diff(X)*diff(Y)))."""
    return np.concatenate(([0.0], r_cumsum(r_diff(X) * r_diff(Y))))


# --------------------------------------------------------------------------
# euler / heun, one path per call, arbitrary closures  (R/sdetools.R:240-470)
# --------------------------------------------------------------------------
def _nformals(fn) -> int:
    return len(inspect.signature(fn).parameters)


def _prologue(f, g, times, x0, B, h, r, rng):
    """R/sdetools.R:377-446 (euler) == :242-311 (heun), textually duplicated there."""
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
    times = np.asarray(times, dtype=np.float64)
    nx, nt = x0.shape[0], times.shape[0]
    ff = (lambda t, x: f(x)) if _nformals(f) == 1 else f                     # :381-386
    ggg = (lambda t, x: g(x)) if _nformals(g) == 1 else g                    # :389-394
    if h is None:                                                            # :397
        h = lambda t, x: 1.0
    no_mortality = r is None                                                 # :406-409
    if r is None:
        r = lambda t, x: 0.0
    rr = (lambda t, x: r(x)) if _nformals(r) == 1 else r                     # :411-416
    g0 = np.asarray(ggg(times[0], x0), dtype=np.float64)                     # :420
    if g0.ndim < 2:                                                          # :422-424 vector g => nB=1
        nB = 1
        gg = lambda t, x: np.reshape(np.broadcast_to(np.asarray(ggg(t, x), dtype=np.float64), (nx,)), (nx, 1))
    else:                                                                    # :425-428
        nB = g0.shape[1]
        gg = lambda t, x: np.asarray(ggg(t, x), dtype=np.float64)
    if B is None:                                                            # :431-433
        B = rvBM(times, nB, rng=rng)
    B = np.asarray(B, dtype=np.float64)
    if B.ndim < 2:                                                           # :435 recycled into nB columns
        B = np.stack([B] * nB, axis=1)
    dB = B[1:, :] - B[:-1, :]                                                # :437 apply(B,2,diff)
    dt = r_diff(times)                                                       # :446
    return x0, times, nx, nt, ff, gg, h, rr, no_mortality, dB, dt


def _matvec(gX, dBi):
    """`gX %*% dB[i,]` (:453): column-axpy order of reference dgemv, no FMA.
    For nB=1 it is exactly one multiply per state component."""
    acc = gX[:, 0] * dBi[0]
    for k in range(1, gX.shape[1]):
        acc = acc + gX[:, k] * dBi[k]
    return acc


def _epilogue(times, X, S, i, no_mortality):
    if no_mortality:                                                         # :462-465
        return {"times": times, "X": X}
    tau = times[i + 1]                                                       # :467-468
    Xo = X[: i + 2, :]
    return {"times": times[: i + 2], "X": Xo[:, 0] if Xo.shape[1] == 1 else Xo, "S": S[: i + 2], "tau": tau}


def euler(f, g, times, x0, B=None, p=lambda x: x, h=None, r=None, S0=1.0, Stau=None, rng=None):
    """R/sdetools.R:375-470.  Returns dict(times=, X=) with X (nt, nx); rows after
    an h/S stop stay NaN (R: NA) when r is None, truncated otherwise."""
    x0, times, nx, nt, ff, gg, h, rr, no_mort, dB, dt = _prologue(f, g, times, x0, B, h, r, rng)
    if Stau is None:                                                         # lazy default Stau=runif(1)
        Stau = (np.random.default_rng() if rng is None else rng).random()
    X = np.full((nt, nx), np.nan)
    X[0, :] = x0
    S = np.zeros(nt)
    S[0] = S0
    i = 0
    for i in range(nt - 1):                                                  # :448
        fX = np.asarray(ff(times[i], X[i, :].copy()), dtype=np.float64).reshape(nx)   # :450
        gX = gg(times[i], X[i, :].copy())                                    # :451
        X[i + 1, :] = p((X[i, :] + fX * dt[i]) + _matvec(gX, dB[i, :]))      # :453
        if h(times[i + 1], X[i + 1, :]) <= 0:                                # :455
            break
        S[i + 1] = np.ravel(S[i] * np.exp(-rr(times[i], X[i, :]) * dt[i]))[0]             # :456 (App. B quirk 2: `t` there is base::t)
        if S[i + 1] < Stau:                                                  # :457
            break
    return _epilogue(times, X, S, i, no_mort)


def heun(f, g, times, x0, B=None, p=lambda x: x, h=None, r=None, S0=1.0, Stau=None, rng=None):
    """R/sdetools.R:240-338."""
    x0, times, nx, nt, ff, gg, h, rr, no_mort, dB, dt = _prologue(f, g, times, x0, B, h, r, rng)
    if Stau is None:
        Stau = (np.random.default_rng() if rng is None else rng).random()
    X = np.full((nt, nx), np.nan)
    X[0, :] = x0
    S = np.zeros(nt)
    S[0] = S0
    i = 0
    for i in range(nt - 1):                                                  # :313
        fX = np.asarray(ff(times[i], X[i, :].copy()), dtype=np.float64).reshape(nx)   # :316
        gX = gg(times[i], X[i, :].copy())                                    # :317
        Y = np.asarray(p((X[i, :] + fX * dt[i]) + _matvec(gX, dB[i, :])), dtype=np.float64)  # :318
        fY = np.asarray(ff(times[i + 1], Y.copy()), dtype=np.float64).reshape(nx)     # :321
        gY = gg(times[i + 1], Y.copy())                                      # :322
        # :323  `0.5*(gX+gY) %*% dB` parses as 0.5*((gX+gY) %*% dB)
        X[i + 1, :] = p((X[i, :] + (0.5 * (fX + fY)) * dt[i]) + 0.5 * _matvec(gX + gY, dB[i, :]))
        if h(times[i + 1], X[i + 1, :]) <= 0:                                # :325
            break
        S[i + 1] = np.ravel(S[i] * np.exp(-rr(times[i], X[i, :]) * dt[i]))[0]             # :326
        if S[i + 1] < Stau:                                                  # :327
            break
    return _epilogue(times, X, S, i, no_mort)


# --------------------------------------------------------------------------
# model catalogue: the closures the R-side constructors would return, with the
# operation order fixed as in SURVEY.md App. A / App. C
# --------------------------------------------------------------------------
MODEL_IDS = {"ou": 0, "cir": 1, "vdp": 2, "logistic": 3, "gbm": 4, "doublewell": 5, "linear": 6, "dwsqrt": 7, "loglogistic": 8}
PROJ_IDS = {"identity": 0, "abs": 1, "relu": 2}


def catalogue(model: str, params):
    """Return (f, g, nx, nB) for a catalogue model.  f,g act on arrays shaped
    (nx,) or (nx, npaths) elementwise, so the same closures drive the one-path
    `euler` above and the vectorised `ensemble` below."""
    P = [float(v) for v in np.ravel(params)] if model != "linear" else None
    if model == "ou":          # R/sdetools.R:718  lambda*(xi-x) / sigma ; Killing.Rmd:157-161
        lam, xi, sig = P
        return (lambda x: lam * (xi - x)), (lambda x: sig + 0.0 * x), 1, 1
    if model == "cir":         # HMM-filtering-of-scalar-SDEs.Rmd:33-51  gamma*sqrt(abs(x))
        lam, xi, gam = P
        return (lambda x: lam * (xi - x)), (lambda x: gam * np.sqrt(np.abs(x))), 1, 1
    if model == "vdp":         # vanderPol.Rmd:27-53  c(v, v*(mu-x^2)-x) / c(0,sigma)
        mu, sig = P
        def f(x):
            return np.stack([x[1], x[1] * (mu - x[0] * x[0]) - x[0]])
        def g(x):
            return np.stack([0.0 * x[0], sig + 0.0 * x[0]])
        return f, g, 2, 1
    if model == "logistic":    # FisheriesManagement.Rmd:30-51  x*(1-x) - c*x^2 / sigma*x
        sig, c = P
        return (lambda x: x * (1.0 - x) - c * (x * x)), (lambda x: sig * x), 1, 1
    if model == "gbm":         # R/sdetools.R:222-228  mu*x / sigma*x
        mu, sig = P
        return (lambda x: mu * x), (lambda x: sig * x), 1, 1
    if model == "doublewell":  # SDEtools.Rmd:50-55  x - x^3 written x - x*x*x / sigma
        sig, = P
        return (lambda x: x - x * x * x), (lambda x: sig + 0.0 * x), 1, 1
    if model == "dwsqrt":      # ItoIntegrals.Rmd:66-69  x - x^3 / sqrt(1 + x^2)
        return (lambda x: x - x * x * x), (lambda x: np.sqrt(1.0 + x * x)), 1, 1
    if model == "loglogistic":  # ItosFormula.Rmd:113-116  1 - exp(y) - 0.5*sigma^2 / sigma
        sig, = P
        return (lambda x: (1.0 - np.exp(x)) - 0.5 * (sig * sig)), (lambda x: sig + 0.0 * x), 1, 1
    if model == "linear":      # R/sdetools.R:360-368  A %*% x / G
        A, G = params
        A = np.asarray(A, dtype=np.float64)
        G = np.asarray(G, dtype=np.float64).reshape(A.shape[0], -1)
        d, nB = A.shape[0], G.shape[1]
        def f(x):
            acc = A[:, 0].reshape((d,) + (1,) * (x.ndim - 1)) * x[0]
            for k in range(1, d):
                acc = acc + A[:, k].reshape((d,) + (1,) * (x.ndim - 1)) * x[k]
            return acc
        def g(x):
            return G if x.ndim == 1 else G.reshape(d, nB, 1) + 0.0 * x[0]
        return f, g, d, nB
    raise ValueError(f"unknown catalogue model {model!r}")


def projection(name: str):
    return {"identity": (lambda x: x), "abs": np.abs, "relu": (lambda x: np.where(x < 0.0, 0.0, x))}[name]


def ensemble(scheme, model, params, times, x0, B=None, dB=None, proj="identity", full=True):
    """Vectorised-over-paths restatement of the euler (:448-458) / heun (:313-328)
    loop for catalogue models; per element it performs exactly the operations of
    `euler`/`heun` above, so results are bit-identical to calling them per path.

    B : (nt, nB, npaths) Brownian paths (reference layout, time first), or
    dB: (nt-1, nB, npaths) increments used directly (Philox-stream checks).
    x0: (nx,) shared or (nx, npaths).  Returns X (nt, nx, npaths) if `full`
    else the terminal state (nx, npaths)."""
    f, g, nx, nB = catalogue(model, params)
    p = projection(proj)
    times = np.asarray(times, dtype=np.float64)
    dt = r_diff(times)
    if dB is None:
        B = np.asarray(B, dtype=np.float64)
        dB = B[1:] - B[:-1]
    npaths = dB.shape[2]
    x = np.empty((nx, npaths))
    x[...] = np.asarray(x0, dtype=np.float64).reshape(nx, -1)
    X = np.empty((times.shape[0], nx, npaths)) if full else None
    if full:
        X[0] = x

    def mv(gx, dBi):                       # gx (nx,npaths) [nB=1 vector g] or (nx,nB,npaths)
        if gx.ndim == 2:
            return gx * dBi[0]
        acc = gx[:, 0] * dBi[0]
        for k in range(1, gx.shape[1]):
            acc = acc + gx[:, k] * dBi[k]
        return acc

    for i in range(times.shape[0] - 1):
        fX, gX = f(x), g(x)
        if scheme == "euler":
            x = p((x + fX * dt[i]) + mv(gX, dB[i]))
        else:
            Y = p((x + fX * dt[i]) + mv(gX, dB[i]))
            fY, gY = f(Y), g(Y)
            x = p((x + (0.5 * (fX + fY)) * dt[i]) + 0.5 * mv(gX + gY, dB[i]))
        if full:
            X[i + 1] = x
    return X if full else x


# --------------------------------------------------------------------------
# terminal statistics as the device kernels define them (include/sdegpu.h)
# --------------------------------------------------------------------------
def hist_counts(x, lo, hi, nbins):
    """Bin convention of R `hist(right=TRUE, include.lowest=TRUE)` on equal-width
    breaks: bin k covers (lo+k*w, lo+(k+1)*w], the first bin also contains lo.
    Returns int64[nbins+3] = [under, bins..., over, nan].  Index arithmetic is
    fixed as  k = ceil((x-lo)*(nbins/(hi-lo))) - 1  with individually rounded
    operations, clamped to [0, nbins-1] for lo <= x <= hi."""
    x = np.asarray(x, dtype=np.float64).ravel()
    inv_w = np.float64(nbins) / (np.float64(hi) - np.float64(lo))
    out = np.zeros(nbins + 3, dtype=np.int64)
    nan = np.isnan(x)
    under = x < lo
    over = x > hi
    ok = ~(nan | under | over)
    k = np.ceil((x[ok] - lo) * inv_w).astype(np.int64) - 1
    k = np.clip(k, 0, nbins - 1)
    out[1:nbins + 1] = np.bincount(k, minlength=nbins)
    out[0], out[nbins + 1], out[nbins + 2] = under.sum(), over.sum(), nan.sum()
    return out


def power_sums(x, center):
    """[n, S1, S2, S3, S4] with S_k = sum (x-center)^k over finite x."""
    x = np.asarray(x, dtype=np.float64).ravel()
    d = x[np.isfinite(x)] - center
    d2 = d * d
    return np.array([d.size, d.sum(), d2.sum(), (d2 * d).sum(), (d2 * d2).sum()])
