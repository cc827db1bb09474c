"""Host-side mirror of the reference's sample-path API (same names, argument
order and return layout as R/sdetools.R), running on libsdegpu.so.

    euler(f, g, times, x0, B=None, p=..., h=None, r=None, S0=1, Stau=None)   R/sdetools.R:375
    heun (f, g, times, x0, B=None, p=..., h=None, r=None, S0=1, Stau=None)   R/sdetools.R:240
    stochint(f, g, rule="l")  :105    itointegral(G, B)  :155
    QuadraticVariation(X)     :134    CrossVariation(X, Y) :184
    rBM(times, sigma, B0, u)  :16     rvBM(times, n, sigma, B0, u) :40     rBrownianBridge :62

New arguments are keyword-only and appended after the reference's, so every
positional call of the reference keeps its meaning.  Arrays use the reference's
column-major layout: numpy results are Fortran-ordered, `X[i, j]` (one path) or
`X[i, j, p]` (ensemble) is time i, component j, path p.
"""
from __future__ import annotations

import numpy as np

from . import catalogue as _cat
from .engine import default_engine


def _catalogue_model(f, g):
    fm, gm = getattr(f, "sdegpu_model", None), getattr(g, "sdegpu_model", None)
    if fm is None or gm is None:
        raise TypeError("f and g must come from sdetools_b200.catalogue: arbitrary closures cannot run on the "
                        "device and this package has no interpreted fallback")
    if fm != gm or not np.array_equal(f.sdegpu_params, g.sdegpu_params) or f.role != "drift" or g.role != "diffusion":
        raise TypeError("f and g must be the (drift, diffusion) pair returned by one catalogue constructor")
    return fm, f.sdegpu_params, f.nx, f.nB


def _tabulate(fn, times, width):
    """fn evaluated at every grid time, as the reference's loop would (`ff(times[i], .)`): nt x width."""
    return np.array([np.broadcast_to(np.asarray(fn(float(t)), dtype=np.float64).reshape(-1), (width,)) for t in times])


def _simulate(scheme, f, g, times, x0, B, p, h, r, S0, Stau, npaths, seed, output, moments, center, hist, arith,
              device, path_offset, integrals=False, model_integrals=False):
    if h is not None and not hasattr(h, "sdegpu_stop"):
        raise TypeError("h must be a catalogue stop rule (catalogue.stop_below/stop_above/stop_outside)")
    if r is not None and not hasattr(r, "sdegpu_kill"):
        raise TypeError("r must be a catalogue killing rate (catalogue.kill_const/kill_exp/kill_quad)")
    model, params, nx, nB = _catalogue_model(f, g)
    proj = _cat.projection_id(p)
    times = np.ascontiguousarray(times, dtype=np.float64)
    nt = times.size
    x0 = np.asarray(x0, dtype=np.float64)
    names = None
    if x0.ndim == 2:                       # (nx, npaths): one initial condition per path
        if x0.shape[0] != nx:
            raise ValueError(f"x0 must have {nx} rows")
        x0c = np.ascontiguousarray(x0.T)   # x0[p*nx+j]
        npaths_x0 = x0.shape[1]
    else:
        if x0.size != nx:
            raise ValueError(f"length(x0) must be {nx} for this model")
        x0c, npaths_x0 = np.ascontiguousarray(x0.ravel()), None
    ensemble = True
    if B is not None:
        Bh = np.asarray(B, dtype=np.float64)
        if Bh.ndim == 1:                   # :435  matrix(B, nrow=nt, ncol=nB): a vector is recycled into nB columns
            Bh = np.repeat(Bh[:, None], nB, axis=1)
        if Bh.ndim == 2:
            if Bh.shape != (nt, nB):
                raise ValueError(f"B must be length(times) x nB = {nt} x {nB}")
            Bh, ensemble = Bh[:, :, None], npaths is not None and npaths != 1
        if Bh.shape[:2] != (nt, nB):
            raise ValueError(f"B must be {nt} x {nB} x npaths")
        if npaths is not None and npaths != Bh.shape[2]:
            raise ValueError("npaths disagrees with the third dimension of B")
        npaths = Bh.shape[2]
        Bdev = np.asfortranarray(Bh)       # memory: B[(p*nB+k)*nt+i]
    else:
        Bdev = None
        if npaths is None:
            npaths, ensemble = (npaths_x0 or 1), npaths_x0 is not None
    if npaths_x0 is not None and npaths_x0 != npaths:
        raise ValueError("x0 has one column per path; its column count must equal npaths")
    if seed is None:
        seed = int(np.random.default_rng().integers(0, 2 ** 63))
    eng = default_engine(device)
    if (integrals or model_integrals) and output == "path":
        output = "terminal"                # the integrals are accumulated in the step loop instead of storing the path
    forcing = _tabulate(f.sdegpu_forcing, times, nx) if getattr(f, "sdegpu_forcing", None) is not None else None
    gscale = _tabulate(g.sdegpu_scale, times, 1)[:, 0] if getattr(g, "sdegpu_scale", None) is not None else None
    want_path = output == "path"
    if output not in ("path", "terminal", "none"):
        raise ValueError("output must be 'path', 'terminal' or 'none'")
    PI = np.empty((4, nx, npaths), order="F") if integrals else None
    MI = np.empty((2, nx, npaths), order="F") if model_integrals else None
    X = np.empty((nt, nx, npaths), order="F") if want_path else None
    XT = np.empty((nx, npaths), order="F") if output in ("path", "terminal") else None
    ps = np.zeros((5, nx), order="F") if moments else None
    hc = np.zeros(hist[2] + 3, dtype=np.uint64) if hist is not None else None
    killed = h is not None or r is not None
    tau = np.empty(npaths) if killed else None
    S_tau = np.empty(npaths) if r is not None else None
    S_path = np.empty((nt, npaths), order="F") if (r is not None and want_path) else None
    Stau_arr = None
    if killed and Stau is not None:              # one threshold per path (the reference: one runif(1) per call)
        Stau_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(Stau, dtype=np.float64), (npaths,)))
    eng.run(model, scheme, params, times, x0c, nx=nx, nB=nB, npaths=npaths, B=Bdev, projection=proj, arith=arith,
            seed=seed, path_offset=path_offset, X=X, XT=XT, power_sums=ps, center=center, hist_counts=hc, hist=hist,
            kill=r, stop=h, S0=S0, Stau=Stau_arr, tau=tau, S_tau=S_tau, path_integrals=PI, S_path=S_path,
            model_integrals=MI, drift_forcing=forcing, diffusion_scale=gscale)
    res = {"times": times}
    if MI is not None:                     # itointegral(f(X), times), itointegral(g(X), B): their sum is X_T - X_0 for euler()
        res["model_integrals"] = {k: (MI[i] if ensemble else MI[i, :, 0]) for i, k in enumerate(("drift", "noise"))}
    if PI is not None:                     # terminal entries of the reference's integral functions on each path
        res["integrals"] = {k: (PI[i] if ensemble else PI[i, :, 0]) for i, k in enumerate(("QV", "ito", "strat", "time"))}
    if killed:
        # with mortality the reference returns tau = times[i+1] and truncates (:466-468); an ensemble keeps
        # the full grid with NaN rows after each path's tau (the r = NULL convention, :439,:464) plus tau/S
        res["tau"] = tau if ensemble else float(tau[0])
        if S_tau is not None:
            res["S"] = S_tau if ensemble else float(S_tau[0])
        if S_path is not None:
            res["S_path"] = S_path         # nt x npaths: every path's S vector, NaN after its tau
    if want_path and killed and r is not None and not ensemble:
        # one path with mortality: exactly the reference's list(times[1:(i+1)], X[1:(i+1),], S[1:(i+1)], tau)  (:466-468)
        k = int(np.searchsorted(times, tau[0], side="left")) + 1
        res["times"] = times[:k]
        res["X"] = X[:k, 0, 0] if nx == 1 else X[:k, :, 0]
        res["S"] = S_path[:k, 0]
    elif want_path:
        res["X"] = X if ensemble else X[:, :, 0]
    if XT is not None:
        res["XT"] = XT if ensemble else XT[:, 0]
    if ps is not None:
        res["power_sums"] = ps
        res["moments"] = moments_from_power_sums(ps, np.zeros(nx) if center is None else np.ravel(center))
    if hc is not None:
        lo, hi, nb = hist[:3]
        res["hist"] = {"breaks": lo + (hi - lo) * np.arange(nb + 1) / nb, "counts": hc[1:nb + 1].astype(np.int64),
                       "under": int(hc[0]), "over": int(hc[nb + 1]), "nan": int(hc[nb + 2])}
    res["seed"] = seed
    return res


def moments_from_power_sums(ps, center):
    """{n, mean, var, skewness, kurtosis} per component from shifted power sums (5, nx)."""
    ps = np.asarray(ps, dtype=np.float64).reshape(5, -1)
    n = ps[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        m1, m2, m3, m4 = ps[1] / n, ps[2] / n, ps[3] / n, ps[4] / n
        var = m2 - m1 * m1
        mu3 = m3 - 3 * m1 * m2 + 2 * m1 ** 3
        mu4 = m4 - 4 * m1 * m3 + 6 * m1 * m1 * m2 - 3 * m1 ** 4
        return {"n": n, "mean": np.asarray(center) + m1, "var": var * n / (n - 1),
                "skewness": mu3 / var ** 1.5, "kurtosis": mu4 / (var * var)}


def euler(f, g, times, x0, B=None, p=None, h=None, r=None, S0=1.0, Stau=None, *, npaths=None, seed=None,
          output="path", moments=False, center=None, hist=None, arith=None, device=None, path_offset=0, integrals=False,
          model_integrals=False):
    """Euler-Maruyama for the Ito SDE dX = f dt + g dB (R/sdetools.R:375-470).

    With the reference's arguments only, returns dict(times=, X=) with X the `nt x nx` matrix of one
    path, exactly like the reference's list(times=,X=).  `npaths` / a 3-d `B` (nt x nB x npaths) /
    a 2-d `x0` (nx x npaths) turn the call into the ensemble the reference builds with sapply
    (vignettes/Killing.Rmd:183-189): X is then `nt x nx x npaths`.  If B is None the Brownian
    increments are drawn on the device (Philox, `seed`).  `output='terminal'` skips the trajectory;
    `moments=True` / `hist=(lo, hi, nbins[, component])` reduce the terminal states on the device.
    `integrals=True` accumulates, in the step loop and without storing the path, the terminal entries of
    QuadraticVariation(X_j), itointegral(X_j,B), stochint(X_j,B,"c") and itointegral(X_j,times) for every path
    and component: res["integrals"] = {"QV","ito","strat","time"}, each nx x npaths (nB = 1 models).
    `model_integrals=True` does the same for the model's own drift and diffusion: res["model_integrals"] = {"drift":
    itointegral(f(X),times)[nt], "noise": itointegral(g(X),B)[nt]} (vignettes/ItoIntegrals.Rmd:82-93).
    Time-dependent closures: f = catalogue.forced(f0, c), g = catalogue.scaled(g0, s) (R/sdetools.R:381-394).
    `device` may be a list of CUDA ordinals: the ensemble is then split over those GPUs (sdegpu_create_multi)."""
    return _simulate("euler", f, g, times, x0, B, p, h, r, S0, Stau, npaths, seed, output, moments, center, hist,
                     arith, device, path_offset, integrals, model_integrals)


def heun(f, g, times, x0, B=None, p=None, h=None, r=None, S0=1.0, Stau=None, *, npaths=None, seed=None,
         output="path", moments=False, center=None, hist=None, arith=None, device=None, path_offset=0, integrals=False,
         model_integrals=False):
    """Heun's method for the Stratonovich SDE dX = f dt + g o dB (R/sdetools.R:240-338); see `euler`."""
    return _simulate("heun", f, g, times, x0, B, p, h, r, S0, Stau, npaths, seed, output, moments, center, hist,
                     arith, device, path_offset, integrals, model_integrals)


# ---- integrals ------------------------------------------------------------------
def _cols(*arrs):
    """Reference vectors (n,) or ensembles (n, ncols) -> Fortran buffers (column c at c*n)."""
    out = []
    shape = None
    for a in arrs:
        a = np.asarray(a, dtype=np.float64)
        a2 = a.reshape(a.shape[0], -1)
        if shape is None:
            shape = a.shape
        elif a.shape != shape:
            raise ValueError("arguments must have the same shape")
        out.append(np.asfortranarray(a2))
    return out, shape


def stochint(f, g, rule="l", *, device=None):
    """R/sdetools.R:105-113.  One rule -> array like f; several -> dict of columns in `rule` order."""
    (fb, gb), shape = _cols(f, g)
    n, nc = fb.shape
    rules = [rule] if isinstance(rule, str) else list(rule)
    if any(k not in ("l", "r", "c") for k in rules):
        raise ValueError("rule must be drawn from 'l', 'r', 'c'")
    bufs = {k: np.empty((n, nc), order="F") for k in set(rules)}
    default_engine(device).stochint(fb, gb, n, nc, bufs.get("l"), bufs.get("r"), bufs.get("c"))
    if isinstance(rule, str):
        return bufs[rule].reshape(shape, order="F")
    return {k: bufs[k].reshape(shape, order="F") for k in rules}


def itointegral(G, B, *, device=None):
    """R/sdetools.R:155-158."""
    (gb, bb), shape = _cols(G, B)
    out = np.empty(gb.shape, order="F")
    default_engine(device).itointegral(gb, bb, gb.shape[0], gb.shape[1], out)
    return out.reshape(shape, order="F")


def QuadraticVariation(X, *, device=None):
    """R/sdetools.R:134-137."""
    (xb,), shape = _cols(X)
    out = np.empty(xb.shape, order="F")
    default_engine(device).qv(xb, xb.shape[0], xb.shape[1], out)
    return out.reshape(shape, order="F")


def CrossVariation(X, Y, *, device=None):
    """R/sdetools.R:184-187."""
    (xb, yb), shape = _cols(X, Y)
    out = np.empty(xb.shape, order="F")
    default_engine(device).crossvar(xb, yb, xb.shape[0], xb.shape[1], out)
    return out.reshape(shape, order="F")


# ---- Brownian motion -----------------------------------------------------------------
def rvBM(times, n=1, sigma=1.0, B0=0.0, u=0.0, *, seed=None, device=None, path_offset=0):
    """R/sdetools.R:40-46: `nt x n`, one column per sample path.  (Vector-valued sigma/B0/u as in the
    reference are applied column-wise on the host side by calling once per distinct value.)"""
    times = np.ascontiguousarray(times, dtype=np.float64)
    if seed is None:
        seed = int(np.random.default_rng().integers(0, 2 ** 63))
    sig, b0, uu = (np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)) for v in (sigma, B0, u))
    out = np.empty((times.size, n), order="F")
    eng = default_engine(device)
    if np.all(sig == sig[0]) and np.all(b0 == b0[0]) and np.all(uu == uu[0]):
        eng.rvbm(times, n, out, sig[0], b0[0], uu[0], seed, path_offset)
    else:
        col = np.empty((times.size, 1), order="F")
        for i in range(n):
            eng.rvbm(times, 1, col, sig[i], b0[i], uu[i], seed, path_offset + i)
            out[:, i] = col[:, 0]
    return out


def _rlaw(law, n, x0, params, t, stratonovich, steps, path, seed, device, path_offset):
    if seed is None:
        seed = int(np.random.default_rng().integers(0, 2 ** 63))
    n = int(n)
    out = np.empty(n)
    traj = np.empty((steps + 1, n), order="F") if path else None
    default_engine(device).rlaw(law, params, t, x0, n, out, nsteps=steps, path=traj, stratonovich=stratonovich, seed=seed,
                                path_offset=path_offset)
    return traj if path else out


def rOU(n, x0, lam, xi, sigma, t, *, steps=1, path=False, seed=None, device=None, path_offset=0):
    """R/sdetools.R:784-791: n draws of X_t | X_0 = x0 for dX = lambda (xi - X) dt + sigma dB.
    `steps` > 1 applies the transition repeatedly (X_{k+1} ~ law(. | X_k, t)); `path=True` returns the
    (steps+1) x n states.  The normals are the ones euler()/heun() give the same path and step under `seed`."""
    return _rlaw("ou", n, x0, (lam, xi, sigma), t, False, steps, path, seed, device, path_offset)


def rCIR(n, x0, lam, xi, gamma, t, Stratonovich=False, *, steps=1, path=False, seed=None, device=None, path_offset=0):
    """R/sdetools.R:695-702: n draws of X_t | X_0 = x0 for dX = lambda (xi - X) dt + gamma sqrt(X) dB, a scaled
    non-central chi-square.  `steps`/`path` as in rOU: the recursive sampling of vignettes/CoxIngersollRoss.Rmd:76-84."""
    return _rlaw("cir", n, x0, (lam, xi, gamma), t, Stratonovich, steps, path, seed, device, path_offset)


def _law(law, what, x, x0, params, t, stratonovich, device):
    xs = np.asarray(x, dtype=np.float64)
    out = np.empty(xs.size)
    default_engine(device).law_eval(law, what, params, t, x0, np.ascontiguousarray(xs.reshape(-1)), out, stratonovich)
    return out.reshape(xs.shape) if xs.ndim else float(out[0])


def dOU(x, x0, lam, xi, sigma, t, *, device=None):
    """R/sdetools.R:745-752, evaluated on the device (x0 scalar or one per point)."""
    return _law("ou", "d", x, x0, (lam, xi, sigma), t, False, device)


def pOU(x, x0, lam, xi, sigma, t, *, device=None):
    """R/sdetools.R:758-765 (lower tail)."""
    return _law("ou", "p", x, x0, (lam, xi, sigma), t, False, device)


def dCIR(x, x0, lam, xi, gamma, t, Stratonovich=False, *, device=None):
    """R/sdetools.R:652-663: 2c dchisq(2c x, n, nu)."""
    return _law("cir", "d", x, x0, (lam, xi, gamma), t, Stratonovich, device)


def pCIR(x, x0, lam, xi, gamma, t, Stratonovich=False, *, device=None):
    """R/sdetools.R:669-676 (lower tail)."""
    return _law("cir", "p", x, x0, (lam, xi, gamma), t, Stratonovich, device)


def rvBM_functionals(times, n=1, sigma=1.0, B0=0.0, u=0.0, level=None, *, seed=None, device=None, path_offset=0):
    """max / min / which.max / which.min / first passage of each column of rvBM(times, n, sigma, B0, u) without
    storing the n paths (vignettes/BrownianMotion.Rmd:101-144: `apply(rvBM(0:Nt,Np),2,max)`).  With the same
    `seed` the values equal those computed from rvBM()'s matrix bit for bit.  Returns a dict of length-n
    arrays: max, min, tmax, tmin and, when `level` is given, thit (NaN where the level is not reached)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    if seed is None:
        seed = int(np.random.default_rng().integers(0, 2 ** 63))
    out = {k: np.empty(n) for k in ("max", "min", "tmax", "tmin")}
    if level is not None:
        out["thit"] = np.empty(n)
    default_engine(device).bm_functionals(times, n, sigma, B0, u, 0.0 if level is None else level, seed, path_offset,
                                          vmax=out["max"], vmin=out["min"], tmax=out["tmax"], tmin=out["tmin"],
                                          thit=out.get("thit"))
    return out


def rBM(times, sigma=1.0, B0=0.0, u=0.0, *, seed=None, device=None):
    """R/sdetools.R:16-22."""
    return rvBM(times, 1, sigma, B0, u, seed=seed, device=device)[:, 0]


def rBrownianBridge(times, B0T=(0.0, 0.0), sigma=1.0, *, n=None, seed=None, device=None):
    """R/sdetools.R:62-73: linear detrending of an rBM path.  `n` gives the nt x n matrix that
    `replicate(n, rBrownianBridge(...))` builds (vignettes/BrownianMotion.Rmd:78-98), one device ensemble."""
    times = np.asarray(times, dtype=np.float64)
    B = rvBM(times, 1 if n is None else n, sigma=sigma, seed=seed, device=device)
    nt = times.size
    slope = (B0T[1] - B0T[0] - B[nt - 1] + B[0]) / (times[nt - 1] - times[0])
    BB = B + B0T[0] - B[0] + slope * (times - times[0])[:, None]
    return BB[:, 0] if n is None else BB
