"""The reference's own source text as the checker: `oracle/minir` evaluates `/root/reference/R/sdetools.R`.

Three layers:
1. the evaluator against R-language facts (precedence, recycling, drop, lazy defaults, cumsum's 80-bit accumulator): answers any
   R session prints, each citing the R documentation section or the reference line whose behaviour depends on it;
2. when the reference checkout is present (this container; not the GPU box): the reference's own testthat file passes under the
   evaluator, and re-running the fixture recipe reproduces the committed `tests/golden/reference_minir/` byte for byte;
3. BASELINE config C1 at full size (scalar OU euler(), 1,000 paths x 1,000 steps, supplied B): the SHA-256 of the result of the
   reference's `euler()` (one call per path, under the evaluator; tests/golden/make_minir_fixtures.py --c1) equals the C oracle's on
   CPU and the CUDA path's through the C ABI on the GPU.
"""
import filecmp
import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle.minir import Interp, from_numpy, to_numpy  # noqa: E402

REF = os.environ.get("SDETOOLS_REF", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "R", "sdetools.R"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="the reference checkout is not present (it does not travel to the GPU box)")
MINIR = os.path.join(ROOT, "tests", "golden", "reference_minir")


def R_eval(src):
    return to_numpy(Interp().run(src))


# ---- 1. the evaluator against the R language ---------------------------------------------------------------------------------
@pytest.mark.parametrize("src,want", [
    ("-2^2", -4.0),                                   # ?Syntax: ^ binds tighter than unary minus
    ("-1:3", [-1, 0, 1, 2, 3]),                       # unary minus binds tighter than :
    ("2^-1", 0.5),
    ("2^3^2", 512.0),                                 # ^ is right-associative
    ("1:3*2", [2, 4, 6]),                             # : binds tighter than *
    ("10-3-2", 5.0),                                  # left-associative
    ("2*3 %*% 4", [[24.0]]),                              # %any% binds tighter than * (R/sdetools.R:323 depends on it)
    ("c(1,2) %*% c(3,4)", [[11.0]]),                  # two vectors of equal length: inner product, a 1 x 1 matrix
    ("!TRUE == FALSE", 1.0),                          # ! binds looser than comparison: !(TRUE == FALSE)
    ("1:6 + 1:2", [2, 4, 4, 6, 6, 8]),                # recycling
    ("matrix(1:6, nrow = 2)[2, ]", [2, 4, 6]),        # column-major fill, drop = TRUE
    ("dim(matrix(1:6, nrow = 2)[, 2:3])", [2, 2]),
    ("dim(matrix(1:6, nrow = 2)[2, , drop = FALSE])", [1, 3]),
    ("x <- c(5, 6, 7); x[-1]", [6, 7]),
    ("x <- c(5, 6, 7); x[-length(x)]", [5, 6]),       # the idiom of stochint/itointegral (R/sdetools.R:109,157)
    ("x <- c(a = 1, b = 2); x[['b']]", 2.0),
    ("diff(c(1, 4, 9, 16))", [3, 5, 7]),
    ("apply(matrix(c(1, 2, 4, 8, 16, 32), nrow = 3), 2, diff)", [[1.0, 8.0], [2.0, 16.0]]),
    ("is.null(dim(apply(matrix(1:4, nrow = 2), 2, diff)))", 1.0),   # nt = 2 collapses dB to a vector
    ("sapply(1:3, function(i) c(i, i^2))", [[1, 2, 3], [1, 4, 9]]),
    ("f <- function(x, y = x * 2) { x <- 10; y }; f(1)", 20.0),      # defaults are promises, forced on first use (R-lang 4.3.3)
    ("f <- function(a, b) { if (missing(b)) 1 else 2 }; f(3)", 1.0),
    ("f <- function(n, mean = 0) mean; f(1, me = 5)", 5.0),          # partial argument matching
    ("c <- 3; c(c, 1)", [3, 1]),                      # function lookup skips non-functions (stochint's `c <- 0.5*(l+r)`, :111)
    ("length(formals(function(t, x) 1))", 2.0),       # R/sdetools.R:381
    ("is.null(formals(abs))", 1.0),
    ("x <- numeric(3); x[2] <- 5; x", [0, 5, 0]),
    ("l <- list(a = 1); l$b$c <- 2; l$b$c", 2.0),
    ("r <- list(m = matrix(0, 2, 2)); r$m[2, ] <- c(1, 2); r$m", [[0, 0], [1, 2]]),          # nested replacement
    ("X <- array(NA, c(2, 2)); X[1, ] <- c(1, 2); is.na(X[2, 1])", 1.0),
    ("seq(0, 1, 0.25)", [0, 0.25, 0.5, 0.75, 1.0]),
    ("length(seq(0, 10, 0.01))", 1001.0),
    ("seq(0, 2 * pi, length = 5)[5] == 2 * pi", 1.0),
    ("s <- 0; for (i in 1:10) { if (i == 3) next; if (i > 5) break; s <- s + i }; s", 12.0),
    ("tryCatch(solve(matrix(0, 2, 2)), error = function(e) -1)", -1.0),   # dLinSDE's fallback (R/sdetools.R:520-534)
    ("data.frame(l = 1:2, r = 3:4)[, 'r']", [3, 4]),
    ("names(data.frame(l = 1:2, r = 3:4, c = 5:6)[, c('l', 'c')])", ["l", "c"]),
    ("x <- c(1, 2, 3); x^2", [1, 4, 9]),
    ("sprintf('%a', 3)", ["0x1.8p+1"]),
    ("sprintf('%a', 0.1)", ["0x1.999999999999ap-4"]),
    ("as.numeric('0x1.8p+1')", 3.0),
    ("paste(c(2, 3), collapse = ' ')", ["2 3"]),
    ("kronecker(diag(2), matrix(1:4, 2))[3:4, 3:4]", [[1, 3], [2, 4]]),
])
def test_r_language_facts(src, want):
    got = R_eval(src)
    if isinstance(want, list) and want and isinstance(want[0], str):
        assert list(got) == want
    else:
        assert np.array_equal(np.asarray(got, dtype=float), np.atleast_1d(np.asarray(want, dtype=float))), (src, got)


def test_one_rounding_per_operator_no_fma():
    """a*b + c is a rounded multiply followed by a rounded add (R evaluates every operator separately)."""
    a, b, c = 1.0 + 2.0**-30, 1.0 - 2.0**-30, -1.0
    got = R_eval(f"{a!r} * {b!r} + {c!r}")
    assert got[0] == (a * b) + c == 0.0                      # the fused result would be -2^-60


def test_cumsum_uses_the_x87_accumulator():
    """R's cumsum keeps an 80-bit accumulator and rounds per element (cum.c); a double accumulator loses the 1s here."""
    got = R_eval("cumsum(c(2^53, 1, 1, -2^53))")
    assert list(got) == [2.0**53, 2.0**53, 2.0**53 + 2, 2.0]
    assert list(np.cumsum([2.0**53, 1, 1, -2.0**53])) != list(got)


def test_lazy_default_Stau_is_forced_once_per_call():
    """`Stau = runif(1)` (R/sdetools.R:375) is a promise: evaluated at the first `S[i+1] < Stau`, once."""
    R = Interp()
    R.run("n <- 0; runif <- function(k) { n <<- n + 1; 0.5 }; f <- function(a, Stau = runif(1)) { for (i in 1:5) if (a < Stau) a <- a + 1; a }")
    assert to_numpy(R.run("f(0)"))[0] == 1.0 and to_numpy(R.run("n"))[0] == 1.0
    assert to_numpy(R.run("f(0, Stau = 3)"))[0] == 3.0 and to_numpy(R.run("n"))[0] == 1.0


# ---- 2. the reference's own files -----------------------------------------------------------------------------------------------
@needs_ref
def test_reference_testthat_file_passes_under_the_evaluator():
    """tests/testthat/test_solvers.R of the reference, unmodified: 2 test_that blocks, 8 expectations (dim(X) for nB = 1, 2 with
    the increments drawn by rvBM; the six dLinSDE answers, the second through tryCatch's Van Loan fallback)."""
    R = Interp()
    R.source(os.path.join(REF, "R", "sdetools.R"))
    R.source(os.path.join(REF, "tests", "testthat", "test_solvers.R"))
    assert [(t[1], t[3]) for t in R.tests] == [(True, 2), (True, 6)], R.tests


@needs_ref
def test_every_exported_function_of_the_reference_file_is_defined():
    R = Interp()
    R.source(os.path.join(REF, "R", "sdetools.R"))
    for name in ("euler", "heun", "stochint", "itointegral", "QuadraticVariation", "CrossVariation", "rBM", "rvBM", "rBrownianBridge",
                 "dLinSDE", "lyap", "dOU", "pOU", "dCIR", "pCIR", "rOU", "rCIR"):
        assert name in R.globalenv.vars, name


@needs_ref
def test_committed_minir_fixtures_are_current():
    """Re-running the recipe on the reference source reproduces tests/golden/reference_minir/*.txt byte for byte."""
    import make_minir_fixtures as M
    with tempfile.TemporaryDirectory() as tmp:
        out = M.run_recipe(REF, os.path.relpath(tmp, M.HERE))
        names = sorted(f for f in os.listdir(MINIR) if f.endswith(".txt"))
        assert names and sorted(f for f in os.listdir(out) if f.endswith(".txt")) == names
        match, mismatch, errors = filecmp.cmpfiles(MINIR, out, names, shallow=False)
        assert not mismatch and not errors, (mismatch, errors)


@needs_ref
def test_reference_laws_under_the_evaluator_agree_with_the_oracle_laws():
    """dOU/pOU/dCIR/pCIR (R/sdetools.R:652-676,745-765) from the reference's own bodies (base R's dnorm/pchisq/... replaced by
    scipy's) against oracle/laws.py, incl. SURVEY §8(c)'s spot values."""
    from oracle import laws
    R = Interp()
    R.source(os.path.join(REF, "R", "sdetools.R"))
    x = np.linspace(0.05, 4.0, 23)
    for strat in (False, True):
        d = to_numpy(R.run(f"dCIR(c({','.join(repr(float(v)) for v in x)}), 1, 1.2, 0.8, 0.9, 0.7, Stratonovich = {'TRUE' if strat else 'FALSE'})"))
        assert np.allclose(d, laws.dCIR(x, 1, 1.2, 0.8, 0.9, 0.7, Stratonovich=strat), rtol=1e-11, atol=0)
        p = to_numpy(R.run(f"pCIR(c({','.join(repr(float(v)) for v in x)}), 1, 1.2, 0.8, 0.9, 0.7, Stratonovich = {'TRUE' if strat else 'FALSE'})"))
        assert np.allclose(p, laws.pCIR(x, 1, 1.2, 0.8, 0.9, 0.7, Stratonovich=strat), rtol=1e-11, atol=0)
    assert abs(to_numpy(R.run("dCIR(1,1,1,1,1,1)"))[0] - 0.5799888552024156) < 1e-13
    assert abs(to_numpy(R.run("dCIR(1,1,1,1,1,1,Stratonovich=TRUE)"))[0] - 0.6184781231754529) < 1e-13
    d = to_numpy(R.run(f"dOU(c({','.join(repr(float(v)) for v in x)}), 2, 1.3, 0.7, 0.9, 0.4)"))
    assert np.allclose(d, laws.dOU(x, 2, 1.3, 0.7, 0.9, 0.4), rtol=1e-12, atol=0)
    p = to_numpy(R.run(f"pOU(c({','.join(repr(float(v)) for v in x)}), 2, 1.3, 0.7, 0.9, 0.4)"))
    assert np.allclose(p, laws.pOU(x, 2, 1.3, 0.7, 0.9, 0.4), rtol=1e-12, atol=0)
    par = to_numpy(R.run("ou.parameters(1,1,1,1,1)"))
    assert par["mean"][0] == 1.0 and abs(par["sd"][0] - 0.6575198539828996) < 1e-15


# ---- 3. config C1 at full size ----------------------------------------------------------------------------------------------------
def _c1():
    with open(os.path.join(MINIR, "c1_ou_euler.json")) as fh:
        return json.load(fh)


def _sha(X_path_major):
    return hashlib.sha256(np.ascontiguousarray(X_path_major).astype("<f8").tobytes()).hexdigest()


def test_c1_full_size_c_oracle_equals_the_reference_source():
    """1,000 paths x 1,000 steps: the plain-C oracle (also the bench's CPU baseline) gives the reference source's bytes."""
    import make_minir_fixtures as M
    from oracle import c_oracle
    from oracle import sde_oracle as O
    t, B, _ = M.c1_inputs()
    fix = _c1()
    assert fix["npaths"] == 1000 and fix["nt"] == t.size == 1001
    for s in fix["sets"]:
        X, _ = c_oracle.ensemble(0, O.MODEL_IDS["ou"], [s["lambda"], s["xi"], s["sigma"]], 1, 1, O.PROJ_IDS["identity"], t,
                                 np.array([s["x0"]]), B=np.ascontiguousarray(B.T.reshape(1000, 1, t.size)), nthreads=4)
        X = X.reshape(1000, t.size)
        assert [float(v).hex() for v in X[:, -1]] == s["XT"]
        assert _sha(X) == s["sha256_X"]


@needs_ref
def test_c1_sample_of_paths_reproduced_live():
    """Three of the 1,000 paths re-run now through the reference's euler() under the evaluator hit the stored terminal values."""
    import make_minir_fixtures as M
    fix = _c1()
    s = fix["sets"][1]
    lo, out = M._c1_chunk((REF, s["lambda"], s["xi"], s["sigma"], s["x0"], 7, 8))
    assert [float(v).hex() for v in out[0, ::50]] == s["X_path7_every50"] and float(out[0, -1]).hex() == s["XT"][7]


@pytest.mark.gpu
def test_c1_full_size_cuda_equals_the_reference_source():
    """The CUDA path through the C ABI (STRICT arithmetic, supplied B): same SHA-256 as the reference's euler() source."""
    import make_minir_fixtures as M
    import sdetools_b200 as sde
    t, B, _ = M.c1_inputs()
    for s in _c1()["sets"]:
        f, g = sde.catalogue.ou(s["lambda"], s["xi"], s["sigma"])
        X = sde.euler(f, g, t, s["x0"], B.reshape(t.size, 1, 1000))["X"]             # (nt, 1, P)
        Xp = np.ascontiguousarray(X.transpose(2, 1, 0)).reshape(1000, t.size)
        assert [float(v).hex() for v in Xp[:, -1]] == s["XT"]
        assert _sha(Xp) == s["sha256_X"]
