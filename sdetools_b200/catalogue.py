"""Compiled model catalogue.

The reference takes arbitrary R closures `f`, `g` (R/sdetools.R:375, :240); a
device kernel cannot call back into an interpreter, so drift and diffusion come
from this catalogue.  Each constructor returns a pair of ordinary callables
`(f, g)` — usable exactly like the closures in the reference's examples — that
carry the attributes `sdegpu_model` / `sdegpu_params`, which is what `euler` and
`heun` dispatch on.  Operation order of every formula is the one the matching
R closure evaluates and is what the kernels reproduce in
STRICT arithmetic.

Calling the closures evaluates the formula with numpy; that is a convenience
for plotting drift/diffusion, never the simulation path.
"""
from __future__ import annotations

import numpy as np


class CatalogueFunction:
    """A drift or diffusion closure tagged with its catalogue identity."""

    def __init__(self, role, model, params, fn, nx, nB):
        self.role, self.sdegpu_model, self.fn = role, model, fn
        self.sdegpu_params = np.ascontiguousarray(params, dtype=np.float64)
        self.nx, self.nB = nx, nB

    def __call__(self, *args):          # f(x) or f(t,x), like the reference accepts (:381-394)
        v = self.fn(np.asarray(args[-1], dtype=np.float64))
        c, s = getattr(self, "sdegpu_forcing", None), getattr(self, "sdegpu_scale", None)
        if (c is not None or s is not None) and len(args) < 2:
            raise TypeError("a time-dependent catalogue function is called as f(t, x)")
        if c is not None:
            v = v + np.asarray(c(args[0]), dtype=np.float64).reshape(np.shape(v)) if np.ndim(v) else v + float(c(args[0]))
        if s is not None:
            v = float(s(args[0])) * v
        return v

    def __repr__(self):
        return f"<catalogue {self.role} {self.sdegpu_model} params={self.sdegpu_params.tolist()}>"


def _pair(model, params, f, g, nx=1, nB=1):
    return (CatalogueFunction("drift", model, params, f, nx, nB),
            CatalogueFunction("diffusion", model, params, g, nx, nB))


def ou(lambda_=1.0, xi=0.0, sigma=1.0):
    """dX = lambda*(xi-X) dt + sigma dB  (R/sdetools.R:718; Killing.Rmd:157-161)."""
    return _pair("ou", [lambda_, xi, sigma], lambda x: lambda_ * (xi - x), lambda x: sigma + 0.0 * x)


def cir(lambda_=1.0, xi=1.0, gamma=1.0):
    """dX = lambda*(xi-X) dt + gamma*sqrt(abs(X)) dB; use p=abs (HMM-filtering-of-scalar-SDEs.Rmd:33-51)."""
    return _pair("cir", [lambda_, xi, gamma], lambda x: lambda_ * (xi - x), lambda x: gamma * np.sqrt(np.abs(x)))


def vdp(mu=1.0, sigma=0.5):
    """van der Pol: f = c(v, v*(mu-x^2)-x), g = c(0,sigma)  (vanderPol.Rmd:27-53); nx = 2, one channel."""
    return _pair("vdp", [mu, sigma],
                 lambda x: np.stack([x[1], x[1] * (mu - x[0] * x[0]) - x[0]]),
                 lambda x: np.stack([0.0 * x[0], sigma + 0.0 * x[0]]), nx=2)


def logistic(sigma=0.5, harvest=0.0):
    """Logistic fisheries: f = x*(1-x) - harvest*x^2, g = sigma*x  (FisheriesManagement.Rmd:30-51)."""
    return _pair("logistic", [sigma, harvest], lambda x: x * (1.0 - x) - harvest * (x * x), lambda x: sigma * x)


def gbm(mu=0.0, sigma=1.0):
    """Geometric Brownian motion f = mu*x, g = sigma*x  (R/sdetools.R:222-228)."""
    return _pair("gbm", [mu, sigma], lambda x: mu * x, lambda x: sigma * x)


def doublewell(sigma=0.5):
    """f = x - x^3 (evaluated x - x*x*x), g = sigma  (SDEtools.Rmd:50-55)."""
    return _pair("doublewell", [sigma], lambda x: x - x * x * x, lambda x: sigma + 0.0 * x)


def dwsqrt():
    """f = x - x^3 (evaluated x - x*x*x), g = sqrt(1 + x^2): the example of vignettes/ItoIntegrals.Rmd:66-69."""
    return _pair("dwsqrt", [], lambda x: x - x * x * x, lambda x: np.sqrt(1.0 + x * x))


def loglogistic(sigma=0.25):
    """dY = (1 - exp(Y) - sigma^2/2) dt + sigma dB: Y = log X of stochastic logistic growth (vignettes/ItosFormula.Rmd:113-116)."""
    return _pair("loglogistic", [sigma], lambda y: (1.0 - np.exp(y)) - 0.5 * (sigma * sigma), lambda y: sigma + 0.0 * y)


def linear(A, G):
    """dX = A X dt + G dB with matrix-valued g: nB = ncol(G) channels (R/sdetools.R:360-368)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    d = A.shape[0]
    G = np.asarray(G, dtype=np.float64).reshape(d, -1)
    params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
    return _pair("linear", params, lambda x: A @ x, lambda x: G, nx=d, nB=G.shape[1])


def forced(f, c):
    """The time-dependent drift f(t,x) = f0(x) + c(t): `f` a catalogue drift, `c` a callable t -> scalar or nx-vector
    (the reference's f(t,x) form, R/sdetools.R:381-386; its own example f <- function(t,x) A %*% x + F * sin(2*t), :211,:363,
    is forced(linear(A,G)[0], lambda t: F * sin(2*t))).  euler/heun evaluate c only at the grid times, as the reference does."""
    if getattr(f, "role", None) != "drift":
        raise TypeError("forced() takes the drift of a catalogue pair")
    out = CatalogueFunction("drift", f.sdegpu_model, f.sdegpu_params, f.fn, f.nx, f.nB)
    out.sdegpu_forcing = c
    return out


def scaled(g, s):
    """The time-dependent diffusion g(t,x) = s(t) * g0(x): `g` a catalogue diffusion with one channel, `s` a callable
    t -> scalar (the reference's g(t,x) form, R/sdetools.R:388-394)."""
    if getattr(g, "role", None) != "diffusion":
        raise TypeError("scaled() takes the diffusion of a catalogue pair")
    if g.sdegpu_model == "linear":
        raise TypeError("scaled() is not available for the linear model")
    out = CatalogueFunction("diffusion", g.sdegpu_model, g.sdegpu_params, g.fn, g.nx, g.nB)
    out.sdegpu_scale = s
    return out


class KillRate:
    """Killing rate r(x) for the `r` argument of euler/heun (R/sdetools.R:405-416,456; Killing.Rmd:152-190)."""

    def __init__(self, kind, a, b=0.0, component=0, fn=None):
        self.sdegpu_kill, self.a, self.b, self.component, self.fn = kind, float(a), float(b), int(component), fn

    def __call__(self, *args):          # r(x) or r(t,x)
        x = np.atleast_1d(np.asarray(args[-1], dtype=np.float64))
        return self.fn(x[self.component])


class StopRule:
    """Stopping function h(t,x) for the `h` argument: the path stops when h <= 0 (R/sdetools.R:455)."""

    def __init__(self, kind, lo=0.0, hi=0.0, component=0, fn=None):
        self.sdegpu_stop, self.lo, self.hi, self.component, self.fn = kind, float(lo), float(hi), int(component), fn

    def __call__(self, t, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return self.fn(x[self.component])


def kill_const(a):
    return KillRate("const", a, fn=lambda v: a + 0.0 * v)


def kill_exp(a, b, component=0):
    """r = a*exp(b*x)   (Killing.Rmd: r <- function(x) 0.5*exp(2*x))."""
    return KillRate("exp", a, b, component, fn=lambda v: a * np.exp(b * v))


def kill_quad(a, b=0.0, component=0):
    """r = a*(x-b)*(x-b)."""
    return KillRate("quad", a, b, component, fn=lambda v: a * ((v - b) * (v - b)))


def stop_below(lo, component=0):
    """h = x - lo: stop when x <= lo."""
    return StopRule("below", lo=lo, component=component, fn=lambda v: v - lo)


def stop_above(hi, component=0):
    """h = hi - x: stop when x >= hi."""
    return StopRule("above", hi=hi, component=component, fn=lambda v: hi - v)


def stop_outside(lo, hi, component=0):
    """h = (x-lo)*(hi-x): stop on leaving (lo, hi)."""
    return StopRule("outside", lo=lo, hi=hi, component=component, fn=lambda v: (v - lo) * (hi - v))


def relu(x):
    """Projection pmax(x, 0)."""
    return np.where(np.asarray(x) < 0, 0.0, x)


def identity(x):
    return x


def projection_id(p) -> str:
    """Map the reference's `p` argument onto a catalogue projection."""
    if p is None or p is identity:
        return "identity"
    if isinstance(p, str) and p in ("identity", "abs", "relu"):
        return p
    if p is abs or p is np.abs or p is np.absolute or p is np.fabs:
        return "abs"
    if p is relu:
        return "relu"
    raise TypeError("projection p must be None, abs, catalogue.relu or one of 'identity'/'abs'/'relu': "
                    "arbitrary closures cannot run on the device")
