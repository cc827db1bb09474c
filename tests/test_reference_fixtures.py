"""Parity against outputs of the reference ITSELF, from two independent executions of its source.

`tests/golden/reference_minir/` (committed): the unmodified `/root/reference/R/sdetools.R`, parsed and evaluated by `oracle/minir`
— a minimal evaluator for the R subset the reference is written in (GNU R is not installed in the build image).  The functions that
ran are the reference's own bodies; only the base-R primitives underneath them are restated (oracle/minir/interp.py).  Made by
`python tests/golden/make_minir_fixtures.py`, which runs the recipe below under that evaluator.

`tests/golden/reference_R/` (absent in the repository): the same recipe run by GNU R on any box that has it:

    Rscript tests/golden/make_reference_fixtures.R /path/to/SDEtools     # writes tests/golden/reference_R/*.txt
    python -m pytest tests/test_reference_fixtures.py                    # CPU: oracle == R;  -m gpu: CUDA == R

Every test below runs once per source that is present ([minir] always, [R] skipped with that reason when absent).
The inputs (tests/golden/reference_R_inputs/, C99 hex doubles written by make_golden.py) are committed; outputs
use the same exact format, so every comparison below is bit for bit unless a tolerance is written next to it.
Reference lines: euler R/sdetools.R:375-470, heun :240-338, stochint :105-113, QuadraticVariation :134-137,
itointegral :155-158, CrossVariation :184-187, rBM :16-22.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden import read_hex  # noqa: E402

IN = os.path.join(ROOT, "tests", "golden", "reference_R_inputs")
# SDEGPU_REFERENCE_R_DIR: another directory of outputs in the same format (tools/selfcheck_reference_fixtures.py uses it to
# exercise this file's plumbing with the oracle's outputs standing in for R's — a check of the tests, not of parity)
SOURCES = {"R": os.environ.get("SDEGPU_REFERENCE_R_DIR") or os.path.join(ROOT, "tests", "golden", "reference_R"),
           "minir": os.path.join(ROOT, "tests", "golden", "reference_minir")}
OUT = SOURCES["minir"]


@pytest.fixture(autouse=True, params=["minir", "R"])
def reference_source(request):
    """Each test runs against every set of reference outputs that exists."""
    global OUT
    d = SOURCES[request.param]
    if not (os.path.isdir(d) and any(f.endswith(".txt") for f in os.listdir(d))):
        pytest.skip("no outputs of GNU R under tests/golden/reference_R/ (R is not installed here): run `Rscript "
                    "tests/golden/make_reference_fixtures.R <reference checkout>` on a box with R" if request.param == "R" else
                    "tests/golden/reference_minir/ is missing: run `python tests/golden/make_minir_fixtures.py`")
    OUT = d
    yield request.param


def needs_R(fn):                           # kept as a marker of which tests read reference outputs; skipping is the fixture's job
    return fn


# (model, params, x0, projection) — the closures make_reference_fixtures.R writes out, in its order
CASES = [("ou", [1.3, 0.7, 0.9], [2.0], "identity"), ("cir", [1.0, 1.0, 1.0], [1.0], "abs"),
         ("cir", [0.5, 0.2, 1.5], [0.05], "relu"), ("vdp", [1.0, 0.5], [0.1, -0.2], "identity"),
         ("logistic", [0.5, 1.0], [0.3], "identity"), ("gbm", [0.4, 1.0], [1.0], "abs"),
         ("doublewell", [0.5], [0.2], "identity"), ("dwsqrt", [], [0.2], "identity"), ("loglogistic", [0.25], [-0.5], "identity")]
# models whose closures call exp(): R uses the C library's exp, numpy and CUDA their own, all within 1 ulp of each
# other but not bit-compatible, so these are compared to a tolerance instead
TRANSCENDENTAL = {"loglogistic"}
KILL = dict(a=0.5, b=2.0)                 # r(x) = 0.5 exp(2x)   (vignettes/Killing.Rmd)
STOP = dict(lo=-0.3, hi=1.2)              # h(t,x) = (x - lo)(hi - x)


def inp(name):
    return read_hex(os.path.join(IN, name + ".txt"))


def ref(name):
    return read_hex(os.path.join(OUT, name + ".txt"))


def same(a, b, tol=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    if tol is None:
        return bool(np.array_equal(a, b, equal_nan=True))
    return bool(np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=tol, atol=tol))


def test_inputs_round_trip():
    """The committed hex inputs are the arrays of oracle_vectors.npz, bit for bit (no R needed)."""
    vec = np.load(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"))
    for k in ("times", "B", "lin_A", "lin_G", "lin_x0", "lin_B", "int_f", "int_g"):
        assert np.array_equal(inp(k), vec[k]), k


# ---- the oracle against R ------------------------------------------------------------------------------
@needs_R
@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_matches_R_paths(k, scheme):
    from oracle import sde_oracle as O
    model, params, x0, proj = CASES[k]
    t, B = inp("times"), inp("B")
    f, g, _, _ = O.catalogue(model, params)
    fn = getattr(O, scheme)
    X = np.stack([fn(f, g, t, x0, B=B[:, :, p], p=O.projection(proj))["X"].reshape(t.size, -1) for p in range(B.shape[2])], axis=2)
    assert same(X, ref(f"X_{scheme}_{k}"), 1e-13 if model in TRANSCENDENTAL else None)


@needs_R
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_oracle_matches_R_linear(scheme):
    from oracle import sde_oracle as O
    t, B = inp("times"), inp("lin_B")
    f, g, _, _ = O.catalogue("linear", (inp("lin_A"), inp("lin_G")))
    X = np.stack([getattr(O, scheme)(f, g, t, inp("lin_x0"), B=B[:, :, p])["X"] for p in range(B.shape[2])], axis=2)
    # %*% goes through R's BLAS: summation order inside dgemv is the vendor's, so 2 ulp instead of bit for bit
    assert same(X, ref(f"lin_X_{scheme}"), 1e-15)


def _oracle_killed(scheme, what):
    from oracle import sde_oracle as O
    t, B, Stau = inp("times"), inp("B"), inp("Stau")
    f, g, _, _ = O.catalogue("ou", [1.0, 0.0, 1.0])
    kill = (lambda x: KILL["a"] * np.exp(KILL["b"] * x))
    stop = (lambda tt, x: (x - STOP["lo"]) * (STOP["hi"] - x))
    nt, P = t.size, B.shape[2]
    X, S, tau = np.full((nt, 1, P), np.nan), np.full((nt, P), np.nan), np.zeros(P)
    for p in range(P):
        sol = getattr(O, scheme)(f, g, t, [0.5], B=B[:, :, p], h=None if what == "kill" else stop,
                                 r=(lambda x: 0.0 * x) if what == "stop" else kill, Stau=float(Stau[p]))
        n = sol["times"].size
        X[:n, 0, p], S[:n, p], tau[p] = np.ravel(sol["X"]), sol["S"], sol["tau"]
    return X, S, tau


@needs_R
@pytest.mark.parametrize("what", ["kill", "stop", "both"])
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_oracle_matches_R_mortality(scheme, what):
    X, S, tau = _oracle_killed(scheme, what)
    assert same(tau, ref(f"kill_tau_{scheme}_{what}"))
    assert same(X, ref(f"kill_X_{scheme}_{what}"))
    assert same(S, ref(f"kill_S_{scheme}_{what}"), 1e-15)          # exp(): C library vs numpy, last ulp


@needs_R
def test_oracle_matches_R_integrals_and_rbm():
    from oracle import sde_oracle as O
    f, g = inp("int_f"), inp("int_g")
    nc = f.shape[0]
    for rule in "lrc":
        assert same(np.stack([O.stochint(f[c], g[c], rule) for c in range(nc)]), ref(f"int_stochint_{rule}")), rule
    assert same(np.stack([O.itointegral(f[c], g[c]) for c in range(nc)]), ref("int_ito"))
    assert same(np.stack([O.QuadraticVariation(g[c]) for c in range(nc)]), ref("int_qv"))
    assert same(np.stack([O.CrossVariation(f[c], g[c]) for c in range(nc)]), ref("int_cv"))
    z, tb = inp("rbm_z"), inp("rbm_times")
    assert same(np.stack([O.rBM(tb, sigma=0.7, B0=0.25, u=-0.1, z=z[:, p]) for p in range(z.shape[1])], axis=1), ref("rbm_B"))
    if os.path.exists(os.path.join(OUT, "rvbm_B.txt")):          # fixtures made since rvBM / rBrownianBridge joined the recipe
        assert same(O.rvBM(tb, z.shape[1], sigma=[0.7, 1.3, 0.2], B0=[0.25, -1, 0], u=[-0.1, 0, 0.4], z=z), ref("rvbm_B"))
        assert same(O.rBrownianBridge(tb, B0T=(1.0, -2.0), sigma=0.8, z=z[:, 1]), ref("bridge_B"))


def _td():
    cfun = lambda t: 0.3 * np.sin(2 * t)
    sfun = lambda t: 1 + 0.5 * np.cos(3 * t)
    Fin = np.array([0.5, -1.0, 0.25])[: inp("lin_x0").size]
    return cfun, sfun, Fin


@needs_R
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_oracle_matches_R_time_dependent_closures(scheme):
    from oracle import sde_oracle as O
    if not os.path.exists(os.path.join(OUT, f"td_X_{scheme}.txt")):
        pytest.skip("fixtures predate the time-dependent cases: re-run make_reference_fixtures.R")
    cfun, sfun, Fin = _td()
    t, B, B2 = inp("times"), inp("B"), inp("lin_B")
    fn = getattr(O, scheme)
    X = np.stack([fn(lambda tt, x: 1.0 * (1.0 - x) + cfun(tt), lambda tt, x: sfun(tt) * (1.0 * np.sqrt(np.abs(x))), t, [1.0],
                     B=B[:, :, p], p=abs)["X"].reshape(t.size, 1) for p in range(B.shape[2])], axis=2)
    assert same(X, ref(f"td_X_{scheme}"), 1e-15)                  # sin/cos: C library vs numpy, last ulp
    A, G, x0 = inp("lin_A"), inp("lin_G"), inp("lin_x0")
    XL = np.stack([fn(lambda tt, x: A @ x + Fin * np.sin(2 * tt), lambda x: G, t, x0, B=B2[:, :, p])["X"] for p in range(B2.shape[2])], axis=2)
    assert same(XL, ref(f"td_lin_X_{scheme}"), 1e-14)


@needs_R
def test_oracle_matches_R_linear_laws():
    from oracle import laws
    if not os.path.exists(os.path.join(OUT, "law_lyap.txt")):
        pytest.skip("no law fixtures (the Matrix package was not available to make_reference_fixtures.R)")
    A, G, x0 = inp("lin_A"), inp("lin_G"), inp("lin_x0")
    _, _, Fin = _td()
    S0 = np.diag(np.arange(1, x0.size + 1) / 10)
    assert same(laws.lyap(A, G @ G.T), ref("law_lyap"), 1e-12)
    l0 = laws.dLinSDE(A, G, 0.7)
    assert same(l0["eAt"], ref("law_eAt"), 1e-12) and same(l0["St"], ref("law_St"), 1e-12)
    assert same(laws.dLinSDE(A, G, 0.7, S0=S0)["St"], ref("law_St_S0"), 1e-12)
    assert same(np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0)["EX"]), ref("law_EX"), 1e-12)
    assert same(np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0, u=Fin)["EX"]), ref("law_EX_zoh"), 1e-12)
    assert same(np.ravel(laws.dLinSDE(A, G, 0.7, x0=x0, u=np.stack([Fin, 2 * Fin], axis=1))["EX"]), ref("law_EX_foh"), 1e-12)


# ---- the CUDA path against R (through the C ABI) -----------------------------------------------------------
@needs_R
@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("k", range(len(CASES)))
def test_cuda_matches_R_paths(k, scheme):
    import sdetools_b200 as sde
    model, params, x0, proj = CASES[k]
    f, g = getattr(sde.catalogue, model)(*params)
    fn = getattr(sde, scheme)
    X = fn(f, g, inp("times"), x0, inp("B"), p={"identity": None, "abs": abs, "relu": "relu"}[proj])["X"]   # STRICT with a supplied B
    assert same(X, ref(f"X_{scheme}_{k}"), 1e-13 if model in TRANSCENDENTAL else None)


@needs_R
@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_cuda_matches_R_linear(scheme):
    import sdetools_b200 as sde
    f, g = sde.catalogue.linear(inp("lin_A"), inp("lin_G"))
    X = getattr(sde, scheme)(f, g, inp("times"), inp("lin_x0"), inp("lin_B"))["X"]
    assert same(X, ref(f"lin_X_{scheme}"), 1e-15)


@needs_R
@pytest.mark.gpu
@pytest.mark.parametrize("what", ["kill", "stop", "both"])
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_cuda_matches_R_mortality(scheme, what):
    import sdetools_b200 as sde
    cat = sde.catalogue
    f, g = cat.ou(1.0, 0.0, 1.0)
    res = getattr(sde, scheme)(f, g, inp("times"), 0.5, inp("B"),
                               h=None if what == "kill" else cat.stop_outside(STOP["lo"], STOP["hi"]),
                               r=cat.kill_const(0.0) if what == "stop" else cat.kill_exp(KILL["a"], KILL["b"]), Stau=inp("Stau"))
    assert same(res["tau"], ref(f"kill_tau_{scheme}_{what}"))
    assert same(res["X"], ref(f"kill_X_{scheme}_{what}"))
    assert same(res["S_path"], ref(f"kill_S_{scheme}_{what}"), 1e-15)


@needs_R
@pytest.mark.gpu
def test_cuda_matches_R_integrals():
    import sdetools_b200 as sde
    f, g = np.ascontiguousarray(inp("int_f").T), np.ascontiguousarray(inp("int_g").T)      # one column per vector
    for rule in "lrc":
        assert same(sde.stochint(f, g, rule).T, ref(f"int_stochint_{rule}")), rule
    assert same(sde.itointegral(f, g).T, ref("int_ito"))
    assert same(sde.QuadraticVariation(g).T, ref("int_qv"))
    assert same(sde.CrossVariation(f, g).T, ref("int_cv"))


@needs_R
@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_cuda_matches_R_time_dependent_closures(scheme):
    import sdetools_b200 as sde
    if not os.path.exists(os.path.join(OUT, f"td_X_{scheme}.txt")):
        pytest.skip("fixtures predate the time-dependent cases: re-run make_reference_fixtures.R")
    cfun, sfun, Fin = _td()
    f0, g0 = sde.catalogue.cir(1.0, 1.0, 1.0)
    X = getattr(sde, scheme)(sde.catalogue.forced(f0, cfun), sde.catalogue.scaled(g0, sfun), inp("times"), 1.0, inp("B"), p=abs)["X"]
    assert same(X, ref(f"td_X_{scheme}"), 1e-15)
    fl, gl = sde.catalogue.linear(inp("lin_A"), inp("lin_G"))
    XL = getattr(sde, scheme)(sde.catalogue.forced(fl, lambda tt: Fin * np.sin(2 * tt)), gl, inp("times"), inp("lin_x0"), inp("lin_B"))["X"]
    assert same(XL, ref(f"td_lin_X_{scheme}"), 1e-14)


@needs_R
@pytest.mark.gpu
def test_cuda_library_matches_R_linear_laws():
    import sdetools_b200 as sde
    if not os.path.exists(os.path.join(OUT, "law_lyap.txt")):
        pytest.skip("no law fixtures (the Matrix package was not available to make_reference_fixtures.R)")
    A, G, x0 = inp("lin_A"), inp("lin_G"), inp("lin_x0")
    _, _, Fin = _td()
    assert same(sde.lyap(A, G @ G.T), ref("law_lyap"), 1e-12)
    l0 = sde.dLinSDE(A, G, 0.7)
    assert same(l0["eAt"], ref("law_eAt"), 1e-12) and same(l0["St"], ref("law_St"), 1e-12)
    assert same(sde.dLinSDE(A, G, 0.7, x0=x0, u=Fin)["EX"], ref("law_EX_zoh"), 1e-12)
    assert same(sde.dLinSDE(A, G, 0.7, x0=x0, u=np.stack([Fin, 2 * Fin], axis=1))["EX"], ref("law_EX_foh"), 1e-12)
