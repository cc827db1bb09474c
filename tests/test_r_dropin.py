"""The R side of the drop-in, EXECUTED end to end: R user code -> r/R/sdegpu.R -> .Call -> r/src/sdegpu_shim.c -> libsdegpu.so.

GNU R is not installed in the image.  `oracle/minir` (an evaluator for the R subset involved; tests/test_minir.py shows it runs the
reference's own source and tests) evaluates the UNMODIFIED wrapper file `r/R/sdegpu.R` — the file a maintainer adds to the package —
and its `.Call()`s are bridged to the compiled shim running on the functional mock of R's C runtime (tests/test_r_shim.py::MockR).
So every line between an R user's `euler(f, g, times, x0, B)` and the C ABI is executed, and what comes back is compared

* on the GPU with the reference's own `euler()`/`heun()`/integral outputs (tests/golden/reference_minir/: the reference source
  evaluated on the same inputs), bit for bit — the calls below are the reference's call forms, `sol$X` and the truncated
  `list(times, X, S, tau)` are the reference's return shapes (R/sdetools.R:464,468);
* on the CPU (reference checkout present) for the dispatch rule: closures that are not from the catalogue run the reference's own
  R loop (`SDEtools::euler`) with identical results, so existing user code keeps working unchanged.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden import read_hex  # noqa: E402
from oracle.minir import NULL, Interp, RList, RNull, RVec, from_numpy, to_numpy  # noqa: E402
from oracle.minir.interp import RError as MiniRError, chr_, int_  # noqa: E402

REF = os.environ.get("SDETOOLS_REF", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "R", "sdetools.R"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="the reference checkout is not present (it does not travel to the GPU box)")
IN = os.path.join(ROOT, "tests", "golden", "reference_R_inputs")
MINIR = os.path.join(ROOT, "tests", "golden", "reference_minir")
ENTRY = ["C_sdegpu_run", "C_sdegpu_stochint", "C_sdegpu_itointegral", "C_sdegpu_qv", "C_sdegpu_crossvar", "C_sdegpu_rvbm",
         "C_sdegpu_bm_functionals", "C_sdegpu_rlaw", "C_sdegpu_law_eval", "C_sdegpu_lyap", "C_sdegpu_dlinsde", "C_sdegpu_info"]


def inp(name):
    return read_hex(os.path.join(IN, name + ".txt"))


def ref(name):
    return read_hex(os.path.join(MINIR, name + ".txt"))


def to_mock(x):
    """An R value of the evaluator as the numpy/dict form MockR.sexp() turns into a SEXP (integer and logical -> INTSXP)."""
    if x is None or isinstance(x, RNull):
        return None
    if isinstance(x, RList):
        return {n: to_mock(v) for n, v in zip(x.names(), x.v)}
    a = np.array(x.v, dtype=np.int32 if x.kind in ("integer", "logical") else np.float64)
    d = x.dim()
    return a if d is None else a.reshape(d, order="F")


def from_mock(v):
    if v is None:
        return NULL
    if isinstance(v, dict):
        return RList([from_mock(e) for e in v.values()], list(v.keys()))
    if isinstance(v, list):
        return chr_(v) if (v and isinstance(v[0], str)) else RList([from_mock(e) for e in v])
    a = np.asarray(v)
    r = from_numpy(a.astype(np.float64))
    if a.dtype.kind in "iu":
        r.kind = "integer"
    return r


class RSession:
    """An R session with the package attached: the reference as namespace SDEtools (when present), r/R/*.R sourced on top."""

    def __init__(self):
        self.R = R = Interp()
        if HAVE_REF:
            R.load_namespace("SDEtools", os.path.join(REF, "R", "sdetools.R"))
        for name in ENTRY:                                   # what useDynLib(.registration = TRUE) binds in the namespace
            R.globalenv.set(name, chr_(name))
        self.mock = None
        self.calls = []
        R.dotcall = self.dotcall
        R.source(os.path.join(ROOT, "r", "R", "fallbacks.R"))
        R.source(os.path.join(ROOT, "r", "R", "sdegpu.R"))

    def dotcall(self, name, args):
        if self.mock is None:
            from test_r_shim import MockR
            self.mock = MockR()
        from test_r_shim import RError
        self.calls.append(name)
        try:
            return from_mock(self.mock.call(name, *[to_mock(a) for a in args]))
        except RError as e:
            raise MiniRError(str(e))

    def set(self, **kw):
        for k, v in kw.items():
            self.R.globalenv.set(k, from_numpy(v) if isinstance(v, (np.ndarray, list, float, int)) else v)

    def run(self, src):
        return to_numpy(self.R.run(src))


@pytest.fixture()
def S():
    return RSession()


# (catalogue constructor call in R, x0, projection argument in R, fixture index) — make_reference_fixtures.R's cases
CASES = [("sde.ou(1.3, 0.7, 0.9)", "2.0", "", 0), ("sde.cir(1.0, 1.0, 1.0)", "1.0", ", p = abs", 1),
         ("sde.cir(0.5, 0.2, 1.5)", "0.05", ", p = 'relu'", 2), ("sde.vdp(1.0, 0.5)", "c(0.1, -0.2)", "", 3),
         ("sde.logistic(0.5, 1.0)", "0.3", "", 4), ("sde.gbm(0.4, 1.0)", "1.0", ", p = abs", 5),
         ("sde.doublewell(0.5)", "0.2", "", 6), ("sde.dwsqrt()", "0.2", "", 7), ("sde.loglogistic(0.25)", "-0.5", "", 8)]


# ---- CPU: the wrapper file itself, and the dispatch rule ------------------------------------------------------------------------
def test_wrapper_file_evaluates_and_keeps_the_reference_signatures(S):
    """euler/heun keep (f, g, times, x0, B, p, h, r, S0, Stau) in place with the reference's defaults (R/sdetools.R:375,:240);
    everything new is appended."""
    for fn in ("euler", "heun"):
        names = list(S.run(f"names(formals({fn}))"))
        assert names[:10] == ["f", "g", "times", "x0", "B", "p", "h", "r", "S0", "Stau"], names
        assert names[10:] == ["npaths", "seed", "output", "moments", "center", "hist", "arith", "integrals", "model_integrals", "devices"]
    assert list(S.run("names(formals(stochint))")) == ["f", "g", "rule"]
    assert list(S.run("names(formals(itointegral))")) == ["G", "B"]
    assert list(S.run("names(formals(dLinSDE))")) == ["A", "G", "t", "x0", "u", "S0"]
    # catalogue constructors return ordinary closures that compute what they say (usable by the unmodified reference loop)
    assert S.run("m <- sde.cir(2, 3, 0.5); c(m$f(1.5), m$g(4), attr(m$f, 'sdegpu_params'))").tolist() == [3.0, 1.0, 2.0, 3.0, 0.5]
    assert S.run("m <- sde.vdp(1, 0.5); m$f(c(0.1, -0.2))").tolist() == [-0.2, -0.2 * (1.0 - 0.1 * 0.1) - 0.1]
    assert S.run("c(.projection_id(function(x) x), .projection_id(abs), .projection_id('relu'), .projection_id(NULL))").tolist() == [0, 1, 2, 0]
    assert np.isnan(S.run(".projection_id(function(x) pmax(x, 0))")[0])


@needs_ref
def test_arbitrary_closures_fall_through_to_the_reference_loop(S):
    """Closures that are not from the catalogue (or a projection the device does not know) take the reference's own R code:
    same numbers as calling SDEtools::euler directly, no native call made."""
    t, B = inp("times"), inp("B")
    S.set(times=t, B1=B[:, 0, 0])
    for scheme in ("euler", "heun"):
        a = S.run(f"{scheme}(function(x) sin(x) - x, function(x) 0.3 + 0.1 * x, times, 0.4, B = B1)$X")
        b = S.run(f"SDEtools::{scheme}(function(x) sin(x) - x, function(x) 0.3 + 0.1 * x, times, 0.4, B = B1)$X")
        assert a.shape == (t.size, 1) and np.array_equal(a, b)
        # a catalogue model with a projection the device does not have: still the reference loop
        c = S.run(f"m <- sde.cir(0.5, 0.2, 1.5); {scheme}(m$f, m$g, times, 0.05, B = B1, p = function(x) pmax(x, 0))$X")
        assert np.array_equal(c, ref(f"X_{scheme}_2")[:, :, 0])
        # f(t, x) / g(t, x) closures and a stopping rule written by the user
        d = S.run(f"{scheme}(function(t, x) -x + sin(t), function(t, x) 1, times, 0, B = B1, h = function(t, x) 0.5 - x)$X")
        e = S.run(f"SDEtools::{scheme}(function(t, x) -x + sin(t), function(t, x) 1, times, 0, B = B1, h = function(t, x) 0.5 - x)$X")
        assert np.array_equal(d, e, equal_nan=True)
    assert S.calls == []
    # small integrals stay in R (one vector is a dependent chain R's cumsum finishes before a device round trip would)
    f, g = inp("int_f")[0], inp("int_g")[0]
    S.set(fv=f, gv=g)
    assert np.array_equal(S.run("itointegral(fv, gv)"), ref("int_ito")[0])
    si = S.run("stochint(fv, gv, c('l', 'c'))")
    assert list(si) == ["l", "c"] and np.array_equal(si["c"], ref("int_stochint_c")[0])
    assert S.calls == []


# ---- GPU: R user code through wrapper, shim and C ABI, against the reference's outputs ---------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("ctor,x0,parg,k", CASES)
def test_r_user_call_returns_the_reference_result(S, ctor, x0, parg, k, scheme):
    """sol <- euler(m$f, m$g, times, x0, B); sol$X — one path per call like the reference, and all paths in one call."""
    t, B = inp("times"), inp("B")
    want = ref(f"X_{scheme}_{k}")
    tol = 1e-13 if "loglogistic" in ctor else None           # exp(): libm vs CUDA, last ulp
    S.set(times=t, B=B)
    S.run(f"m <- {ctor}")
    for p in range(B.shape[2]):
        sol = S.run(f"{scheme}(m$f, m$g, times, {x0}, B = B[, 1, {p + 1}]{parg})")
        assert list(sol)[:2] == ["times", "X"] and sol["X"].shape == want.shape[:2]            # list(times=, X=), nt x nx (:464)
        assert np.array_equal(sol["times"], t)
        assert np.array_equal(sol["X"], want[:, :, p]) if tol is None else np.allclose(sol["X"], want[:, :, p], rtol=tol, atol=tol)
    ens = S.run(f"{scheme}(m$f, m$g, times, {x0}, B = B{parg})$X")                                  # nt x nx x npaths in one launch
    assert np.array_equal(ens, want) if tol is None else np.allclose(ens, want, rtol=tol, atol=tol)
    assert S.calls and set(S.calls) == {"C_sdegpu_run"}


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_r_user_call_linear_system(S, scheme):
    S.set(times=inp("times"), A=inp("lin_A"), G=inp("lin_G"), x0=inp("lin_x0"), B2=inp("lin_B"))
    X = S.run(f"m <- sde.linear(A, G); {scheme}(m$f, m$g, times, x0, B = B2)$X")
    assert np.allclose(X, ref(f"lin_X_{scheme}"), rtol=1e-15, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["kill", "stop", "both"])
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_r_user_call_with_mortality_returns_the_truncated_list(S, scheme, what):
    """With `r` the reference returns list(times[1:(i+1)], X[1:(i+1),], S[1:(i+1)], tau) (R/sdetools.R:468,:336); X drops to a
    vector for nx = 1.  One call per path with its own Stau, as Killing.Rmd:183-189 does."""
    t, B, Stau = inp("times"), inp("B"), inp("Stau")
    wX, wS, wtau = ref(f"kill_X_{scheme}_{what}"), ref(f"kill_S_{scheme}_{what}"), ref(f"kill_tau_{scheme}_{what}")
    S.set(times=t, B=B, Stau=Stau)
    h = "NULL" if what == "kill" else "sde.stop.outside(-0.3, 1.2)"
    r = "sde.kill.const(0)" if what == "stop" else "sde.kill.exp(0.5, 2)"
    S.run("m <- sde.ou(1.0, 0.0, 1.0)")
    for p in range(B.shape[2]):
        sol = S.run(f"{scheme}(m$f, m$g, times, 0.5, B = B[, 1, {p + 1}], h = {h}, r = {r}, Stau = Stau[{p + 1}])")
        assert list(sol) == ["times", "X", "S", "tau"]
        n = sol["times"].size
        assert sol["tau"][0] == wtau[p] and np.array_equal(sol["times"], t[:n]) and t[n - 1] == wtau[p]
        assert sol["X"].shape == (n,) and np.array_equal(sol["X"], wX[:n, 0, p])
        assert np.all(np.isnan(wX[n:, 0, p]))
        assert np.allclose(sol["S"], wS[:n, p], rtol=1e-15, atol=0)


@pytest.mark.gpu
def test_r_integral_wrappers_on_the_device(S):
    """With the size threshold lowered, stochint / itointegral / QuadraticVariation / CrossVariation go through .Call and return
    the reference's values bit for bit, in the reference's shapes (vector for one rule, data.frame columns in `rule` order)."""
    f, g = inp("int_f"), inp("int_g")
    S.run(".sdegpu_integral_min <- 0")
    for c in range(f.shape[0]):
        S.set(fv=f[c], gv=g[c])
        assert np.array_equal(S.run("stochint(fv, gv)"), ref("int_stochint_l")[c])
        si = S.run("stochint(fv, gv, rule = c('c', 'l', 'r'))")
        assert list(si) == ["c", "l", "r"]
        for rule in "lrc":
            assert np.array_equal(si[rule], ref(f"int_stochint_{rule}")[c]), rule
        assert np.array_equal(S.run("itointegral(fv, gv)"), ref("int_ito")[c])
        assert np.array_equal(S.run("QuadraticVariation(gv)"), ref("int_qv")[c])
        assert np.array_equal(S.run("CrossVariation(fv, gv)"), ref("int_cv")[c])
    assert {"C_sdegpu_stochint", "C_sdegpu_itointegral", "C_sdegpu_qv", "C_sdegpu_crossvar"} <= set(S.calls)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_r_time_dependent_closures(S, scheme):
    """f(t,x) = f0(x) + c(t), g(t,x) = s(t) g0(x): the wrapper tabulates c and s at `times` with R's own arithmetic."""
    S.set(times=inp("times"), B=inp("B"), A=inp("lin_A"), G=inp("lin_G"), x0=inp("lin_x0"), B2=inp("lin_B"))
    X = S.run(f"m <- sde.cir(1.0, 1.0, 1.0); f <- sde.forced(m$f, function(t) 0.3 * sin(2 * t)); "
              f"g <- sde.scaled(m$g, function(t) 1 + 0.5 * cos(3 * t)); {scheme}(f, g, times, 1.0, B = B, p = abs)$X")
    assert np.allclose(X, ref(f"td_X_{scheme}"), rtol=1e-15, atol=1e-15)
    XL = S.run(f"l <- sde.linear(A, G); Fin <- c(0.5, -1, 0.25); fl <- sde.forced(l$f, function(t) Fin * sin(2 * t)); "
               f"{scheme}(fl, l$g, times, x0, B = B2)$X")
    assert np.allclose(XL, ref(f"td_lin_X_{scheme}"), rtol=1e-14, atol=1e-14)


@pytest.mark.gpu
def test_r_linear_laws_through_the_library(S):
    S.set(A=inp("lin_A"), G=inp("lin_G"), x0=inp("lin_x0"))
    S.run(".sdegpu_lyap_min <- 0L; Fin <- c(0.5, -1, 0.25)")
    assert np.allclose(S.run("lyap(A, G %*% t(G))"), ref("law_lyap"), rtol=1e-12, atol=1e-12)
    l0 = S.run("dLinSDE(A, G, 0.7)")
    assert np.allclose(l0["eAt"], ref("law_eAt"), rtol=1e-12, atol=1e-12) and np.allclose(l0["St"], ref("law_St"), rtol=1e-12, atol=1e-12)
    assert np.allclose(np.ravel(S.run("dLinSDE(A, G, 0.7, x0 = x0, u = Fin)$EX")), ref("law_EX_zoh"), rtol=1e-12, atol=1e-12)
    assert np.allclose(np.ravel(S.run("dLinSDE(A, G, 0.7, x0 = x0, u = cbind(Fin, 2 * Fin))$EX")), ref("law_EX_foh"), rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_r_ensemble_call_of_the_killing_vignette(S):
    """vignettes/Killing.Rmd:152-207: OU with a constant killing rate; lifetimes are Exp(rate).  The R-level call with npaths
    replaces the vignette's sapply over euler(); the default Stau draws one threshold per path."""
    res = S.run("m <- sde.ou(1, 0, 1); euler(m$f, m$g, seq(0, 20, 0.01), 0, r = sde.kill.const(0.5), npaths = 20000, "
                "seed = 11, output = 'none')")
    tau = res["tau"]
    assert tau.shape == (20000,) and abs(tau.mean() - 2.0) < 5 * 2.0 / np.sqrt(20000) + 0.01
