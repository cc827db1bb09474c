"""Exact transition-law samplers rOU / rCIR (R/sdetools.R:784-791, :695-702).

CPU half: the oracle's restatement of the two samplers against the laws the reference pairs them with.
GPU half: the device samplers (through the C ABI) — deterministic where the law is Gaussian (the normals
are a known stream), Kolmogorov-Smirnov against pOU/pCIR otherwise.  Seeds are fixed, so the p-value
thresholds (1e-3) are not flaky: they were chosen after looking at the statistics once."""
import numpy as np
import pytest
from scipy import stats

from oracle import laws


# ---- CPU: oracle ---------------------------------------------------------------------------------
def test_oracle_rOU_is_mean_plus_sd_z():
    z = np.random.default_rng(1).standard_normal(1000)
    x = laws.rOU(1000, 2.0, 1.3, 0.7, 0.9, 0.4, z=z)
    m, s = laws.ou_parameters(2.0, 1.3, 0.7, 0.9, 0.4)
    assert np.array_equal(x, m + s * z)
    assert laws.rOU(3, 1.0, 0.0, 0.0, 2.0, 4.0, z=np.ones(3)).tolist() == [5.0, 5.0, 5.0]      # lambda = 0: sd = sigma sqrt(t)


@pytest.mark.parametrize("lam,xi,gam,t,x0,strat", [(1.0, 1.0, 1.0, 0.5, 1.0, False), (0.5, 0.2, 1.5, 1.0, 0.05, False),
                                                   (1.0, 1.0, 1.0, 1e-3, 1.0, False), (2.0, 0.5, 0.8, 0.3, 0.1, True)])
def test_oracle_rCIR_follows_pCIR(lam, xi, gam, t, x0, strat):
    x = laws.rCIR(100000, x0, lam, xi, gam, t, strat, rng=np.random.default_rng(7))
    ks = stats.kstest(x, lambda q: laws.pCIR(q, x0, lam, xi, gam, t, strat))
    assert ks.pvalue > 1e-3, ks


# ---- GPU -----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sde(gpu):
    import sdetools_b200
    return sdetools_b200


@pytest.mark.gpu
def test_rOU_is_the_reference_formula_on_the_known_normal_stream(sde):
    eng = sde.default_engine(0)
    n, seed = 5000, 77
    z = eng.normals(seed, 10, n, 1)[0]
    x0 = np.random.default_rng(0).uniform(-1, 3, n)
    got = sde.rOU(n, x0, 1.3, 0.7, 0.9, 0.4, seed=seed, path_offset=10)
    assert np.allclose(got, laws.rOU(n, x0, 1.3, 0.7, 0.9, 0.4, z=z), rtol=1e-13, atol=1e-15)
    got0 = sde.rOU(n, 2.0, 0.0, 0.0, 2.0, 4.0, seed=seed, path_offset=10)                        # shared x0, lambda = 0
    assert np.allclose(got0, 2.0 + 4.0 * z, rtol=1e-13)


@pytest.mark.gpu
def test_rOU_recursion_is_coupled_to_euler_and_exact_in_law(sde):
    lam, xi, sig, x0, n, nsteps = 1.3, 0.7, 0.9, 2.0, 100000, 400
    t = np.arange(nsteps + 1) * (1.0 / nsteps)
    X = sde.rOU(n, x0, lam, xi, sig, t[1], steps=nsteps, path=True, seed=5)
    assert X.shape == (nsteps + 1, n) and np.all(X[0] == x0)
    ks = stats.kstest(X[-1], lambda q: laws.pOU(q, x0, lam, xi, sig, 1.0))
    assert ks.pvalue > 1e-3, ks
    f, g = sde.catalogue.ou(lam, xi, sig)
    E = sde.euler(f, g, t, x0, npaths=2000, seed=5)["X"][:, 0, :]
    # same seed => same Brownian increments: the Euler path shadows the exact one (strong error O(dt))
    assert np.max(np.abs(E - X[:, :2000])) < 0.02
    assert np.corrcoef(E[-1], X[-1, :2000])[0, 1] > 0.9999


CIR_CASES = [(1.0, 1.0, 1.0, 0.5, 1.0, False),       # df = 4
             (0.5, 0.2, 1.5, 1.0, 0.05, False),      # df = 0.18 < 1: gamma shape boost, mass near 0
             (1.0, 1.0, 1.0, 1e-3, 1.0, False),      # ncp/2 ~ 2000: PTRS Poisson, large gamma shape
             (1.0, 1.0, 0.3, 0.01, 2.0, False),      # ncp/2 ~ 4400
             (2.0, 0.5, 0.8, 0.3, 0.1, True),        # Stratonovich parameters (:610)
             (1.0, 1.0, 1.0, np.inf, 1.0, False)]    # stationary law (CoxIngersollRoss.Rmd:56-71)


@pytest.mark.gpu
@pytest.mark.parametrize("lam,xi,gam,t,x0,strat", CIR_CASES)
def test_rCIR_follows_pCIR(sde, lam, xi, gam, t, x0, strat):
    n = 200000
    x = sde.rCIR(n, x0, lam, xi, gam, t, strat, seed=11)
    assert x.shape == (n,) and np.all(x >= 0)
    ks = stats.kstest(x, lambda q: laws.pCIR(q, x0, lam, xi, gam, t, strat))
    assert ks.pvalue > 1e-3, ks
    xi_ito = xi + gam ** 2 / 4 / lam if strat else xi
    mean = xi_ito + (x0 - xi_ito) * np.exp(-lam * t)
    assert abs(x.mean() - mean) < 5 * x.std() / np.sqrt(n)


@pytest.mark.gpu
def test_rCIR_recursion_chapman_kolmogorov_and_sharding(sde):
    lam, xi, gam, x0, n = 1.0, 1.0, 1.0, 0.3, 200000
    x = sde.rCIR(n, x0, lam, xi, gam, 0.02, steps=50, seed=3)                 # 50 exact transitions of 0.02 = one of 1.0
    ks = stats.kstest(x, lambda q: laws.pCIR(q, x0, lam, xi, gam, 1.0))
    assert ks.pvalue > 1e-3, ks
    # a draw is a function of (seed, global id, transition): two shards == one call; per-draw x0
    x0v = np.random.default_rng(2).uniform(0.1, 2.0, 1000)
    whole = sde.rCIR(1000, x0v, lam, xi, gam, 0.1, steps=3, seed=9)
    a = sde.rCIR(400, x0v[:400], lam, xi, gam, 0.1, steps=3, seed=9)
    b = sde.rCIR(600, x0v[400:], lam, xi, gam, 0.1, steps=3, seed=9, path_offset=400)
    assert np.array_equal(whole, np.concatenate([a, b]))
    P = sde.rCIR(1000, x0v, lam, xi, gam, 0.1, steps=3, path=True, seed=9)
    assert np.array_equal(P[0], x0v) and np.array_equal(P[-1], whole)


@pytest.mark.gpu
def test_exact_ensemble_exposes_the_euler_bias(sde):
    """What next-2 is for: C2-style CIR ensemble, Heun (Stratonovich) against the exact law."""
    lam, xi, gam, x0, n = 1.0, 1.0, 1.0, 1.0, 400000
    exact = sde.rCIR(n, x0, lam, xi, gam, 1.0, True, seed=21)
    f, g = sde.catalogue.cir(lam, xi, gam)
    for nsteps, tol in ((20, None), (2000, 0.004)):
        t = np.arange(nsteps + 1) / nsteps
        XT = sde.heun(f, g, t, x0, p=abs, npaths=n, seed=22, output="terminal")["XT"][0]
        d = stats.ks_2samp(XT, exact).statistic
        if tol is None:
            coarse = d
        else:
            assert d < tol and d < coarse          # the fine grid is within sampling error, the coarse one is not


@pytest.mark.gpu
def test_rlaw_argument_errors(sde):
    from sdetools_b200._lib import SdeGpuError
    with pytest.raises(SdeGpuError, match="lambda > 0"):
        sde.rCIR(10, 1.0, 0.0, 1.0, 1.0, 1.0)
    with pytest.raises(SdeGpuError, match="t must be"):
        sde.rOU(10, 1.0, 1.0, 1.0, 1.0, -1.0)
    with pytest.raises(ValueError):
        sde.rOU(10, np.ones(7), 1.0, 1.0, 1.0, 1.0)
    assert sde.rOU(0, 1.0, 1.0, 1.0, 1.0, 1.0).shape == (0,)
