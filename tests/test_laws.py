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


# ---- dLinSDE / lyap of the product (sdegpu_dlinsde / sdegpu_lyap, O(d^3)) against the oracle's restatement of the reference;
# ---- the host logic of the same source runs without a GPU in tests/test_linlaws_host.py ----
@pytest.mark.gpu
def test_product_dLinSDE_reproduces_reference_known_answers_and_oracle(gpu):
    from sdetools_b200 import laws as P
    tol = 1.5e-8
    sol = P.dLinSDE(-1, 1, 1, S0=0.5)                        # tests/testthat/test_solvers.R:20-36
    assert sol["eAt"].item() == pytest.approx(np.exp(-1), rel=tol) and sol["St"].item() == pytest.approx(0.5, rel=tol)
    A = np.array([[0.0, 1.0], [-1.0, 0.0]])
    sol = P.dLinSDE(A, np.eye(2), 2 * np.pi, x0=[1, 0])     # singular Lyapunov equation: Van Loan route
    assert np.allclose(sol["EX"], [1, 0], atol=tol) and np.allclose(sol["St"], 2 * np.pi * np.eye(2), atol=10 * tol)
    assert P.dLinSDE(-1, 1, 2, x0=1, u=1)["EX"][0] == pytest.approx(1, rel=tol)
    assert P.dLinSDE(0, 1, 2, x0=1, u=np.array([[1.0, 3.0]]))["EX"][0] == pytest.approx(5, rel=tol)
    assert P.lyap(-1, 1).item() == pytest.approx(0.5)       # R/sdetools.R:587-590
    Ao = np.array([[0.0, 1.0], [-1.0, -0.1]])
    assert np.allclose(P.lyap(Ao, np.diag([0.0, 1.0])), 5 * np.eye(2))
    with pytest.raises(np.linalg.LinAlgError):
        P.lyap(A, np.eye(2))
    rng = np.random.default_rng(3)
    for d in (1, 3, 7):
        Ar = -np.eye(d) + 0.4 * rng.standard_normal((d, d))
        Gr = rng.standard_normal((d, max(1, d - 1)))
        S0 = np.diag(rng.uniform(0, 1, d))
        x0, u1, u2 = rng.standard_normal(d), rng.standard_normal(d), rng.standard_normal((d, 2))
        for kw in ({}, {"x0": x0}, {"x0": x0, "u": u1}, {"x0": x0, "u": u2}, {"S0": S0}):
            a, b = P.dLinSDE(Ar, Gr, 0.7, **kw), laws.dLinSDE(Ar, Gr, 0.7, **kw)
            assert a.keys() == b.keys()
            for k in a:
                assert np.allclose(a[k], b[k], rtol=1e-10, atol=1e-12), (d, kw.keys(), k)
        assert np.allclose(P.lyap(Ar, Gr @ Gr.T), laws.lyap(Ar, Gr @ Gr.T), rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [256, 600])
def test_product_dLinSDE_at_d256_matches_closed_form(gpu, d):
    """C4's matrix: A + A' = -I and A normal, so S_t = sigma^2 (1 - e^{-t}) I — out of reach of the
    reference's 65536 x 65536 Kronecker solve.  The matrix products run on the device (k_dgemm)."""
    from sdetools_b200 import laws as P
    R = np.random.default_rng(256).standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    sol = P.dLinSDE(A, np.eye(d), 1.0)
    assert np.allclose(sol["St"], (1 - np.exp(-1.0)) * np.eye(d), atol=1e-10)
    assert np.allclose(sol["eAt"] @ sol["eAt"].T, np.exp(-1.0) * np.eye(d), atol=1e-10)
    assert np.allclose(P.lyap(A, np.eye(d)), np.eye(d), atol=1e-10)
    x0, u = np.ones(d), np.linspace(-1, 1, d)
    from scipy import linalg
    AA = np.block([[A, np.eye(d)], [np.zeros((d, 2 * d))]])
    assert np.allclose(P.dLinSDE(A, np.eye(d), 0.5, x0=x0, u=u)["EX"], (linalg.expm(0.5 * AA) @ np.concatenate([x0, u]))[:d], rtol=1e-9, atol=1e-11)


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


@pytest.mark.gpu
def test_linear_exact_stepping_has_no_discretisation_error(sde):
    """vignettes/NoisyOscillator.Rmd:70-80 as an ensemble: X_{k+1} = e^{A dt} X_k + chol(S_dt) Z on a coarse grid
    reproduces the law at T exactly, where plain Euler on the same grid is visibly biased."""
    rng = np.random.default_rng(6)
    d, n, nsteps, T = 4, 200000, 10, 2.0
    A = np.array([[-0.2, 1.0, 0, 0], [-1.0, -0.2, 0, 0], [0.3, 0, -0.5, 0.2], [0, 0.1, 0, -1.0]])
    G = rng.standard_normal((d, 2))
    x0 = np.array([1.0, 0.0, -1.0, 0.5])
    t = np.arange(nsteps + 1) * (T / nsteps)
    want = sde.dLinSDE(A, G, T, x0=x0)
    fe, ge = sde.linear_exact(A, G, T / nsteps)
    XT = sde.euler(fe, ge, t, x0, npaths=n, seed=4, output="terminal")["XT"]          # (d, n)
    se = np.sqrt(np.diag(want["St"]) / n)
    assert np.all(np.abs(XT.mean(axis=1) - want["EX"]) < 5 * se)
    C = np.cov(XT)
    assert np.max(np.abs(C - want["St"])) < 5 * np.max(np.diag(want["St"])) * np.sqrt(2.0 / n)
    f, g = sde.catalogue.linear(A, G)
    XE = sde.euler(f, g, t, x0, npaths=n, seed=4, output="terminal")["XT"]
    assert np.max(np.abs(np.cov(XE) - want["St"])) > 20 * np.max(np.diag(want["St"])) * np.sqrt(2.0 / n)   # dt = 0.2: biased


# ---- next-3: Brownian ensembles and their path functionals -------------------------------------
@pytest.mark.gpu
def test_bm_functionals_equal_those_of_the_stored_ensemble(sde):
    t = np.concatenate(([0.0], np.cumsum(np.random.default_rng(1).uniform(0.01, 0.2, 333))))
    n, seed = 2000, 17
    B = sde.rvBM(t, n, sigma=1.3, B0=0.4, u=-0.1, seed=seed, path_offset=5)                   # (nt, n)
    F = sde.rvBM_functionals(t, n, sigma=1.3, B0=0.4, u=-0.1, level=1.5, seed=seed, path_offset=5)
    assert np.array_equal(F["max"], B.max(axis=0)) and np.array_equal(F["min"], B.min(axis=0))
    assert np.array_equal(F["tmax"], t[B.argmax(axis=0)]) and np.array_equal(F["tmin"], t[B.argmin(axis=0)])
    hit = B >= 1.5
    want = np.where(hit.any(axis=0), t[hit.argmax(axis=0)], np.nan)
    assert np.array_equal(F["thit"], want, equal_nan=True) and 0.05 < np.isnan(want).mean() < 0.95
    down = sde.rvBM_functionals(t, n, sigma=1.3, B0=0.4, u=-0.1, level=-1.0, seed=seed, path_offset=5)["thit"]
    hit = B <= -1.0
    assert np.array_equal(down, np.where(hit.any(axis=0), t[hit.argmax(axis=0)], np.nan), equal_nan=True)
    assert "thit" not in sde.rvBM_functionals(t, 3, seed=1)


@pytest.mark.gpu
def test_maximum_of_brownian_motion_reflection_principle(sde):
    """vignettes/BrownianMotion.Rmd:101-144: P(S_t >= x) = 2 P(B_t >= x); on a grid the maximum is short by
    about 0.5826 sqrt(dt) (the vignette's Nt = 10 vs 1000 discussion), which the comparison allows for."""
    Nt, Np = 4000, 400000
    t = np.arange(Nt + 1, dtype=float)
    m = sde.rvBM_functionals(t, Np, seed=23)["max"]
    assert np.all(m >= 0)                                                       # B_0 = 0 is on the grid
    shift = 0.5826
    for x in (0.5, 1.0, 2.0):
        lvl = x * np.sqrt(Nt)
        emp = np.mean(m + shift >= lvl)
        assert abs(emp - 2 * stats.norm.sf(x)) < 4 * np.sqrt(0.25 / Np) + 2e-3
    coarse = sde.rvBM_functionals(np.arange(11.0), Np, seed=23)["max"]
    assert np.mean(coarse >= np.sqrt(10.0)) < 2 * stats.norm.sf(1.0) - 0.05    # Nt = 10: visibly too few crossings


@pytest.mark.gpu
def test_brownian_bridge_ensemble(sde):
    """vignettes/BrownianMotion.Rmd:78-98: mean on the chord, variance sigma^2 (t-t0)(1-(t-t0)/(T-t0))."""
    tv, n, sigma, B0T = np.arange(10.0, 21.0), 100000, 2.0, (15.0, 5.0)
    BB = sde.rBrownianBridge(tv, B0T, sigma, n=n, seed=8)
    assert BB.shape == (tv.size, n)
    assert np.allclose(BB[0], 15.0) and np.allclose(BB[-1], 5.0)
    s = tv - tv[0]
    var = sigma ** 2 * s * (1 - s / s[-1])
    assert np.max(np.abs(BB.mean(axis=1) - (15.0 - s))) < 5 * np.sqrt(var.max() / n)
    assert np.max(np.abs(BB.var(axis=1) - var)) < 5 * var.max() * np.sqrt(2.0 / n)
    one = sde.rBrownianBridge(tv, B0T, sigma, seed=8)
    assert one.shape == (tv.size,) and np.array_equal(one, BB[:, 0])


# ---- next-2: the laws themselves on the device ------------------------------------------------------
@pytest.mark.gpu
def test_device_laws_match_scipy(sde):
    """dOU/pOU/dCIR/pCIR evaluated on the device against the oracle's scipy-based restatement of the reference
    (R/sdetools.R:652-676, :745-765).  Tolerances: 1e-11 absolute on distribution functions, 1e-9 relative on
    densities that are not denormal."""
    x = np.linspace(-3, 5, 401)
    assert np.max(np.abs(sde.pOU(x, 2.0, 1.3, 0.7, 0.9, 0.4) - laws.pOU(x, 2.0, 1.3, 0.7, 0.9, 0.4))) < 1e-14
    assert np.allclose(sde.dOU(x, 2.0, 1.3, 0.7, 0.9, 0.4), laws.dOU(x, 2.0, 1.3, 0.7, 0.9, 0.4), rtol=1e-12, atol=0)
    for lam, xi, gam, t, x0, strat in CIR_CASES + [(1.0, 1.0, 0.2, 1e-3, 3.0, False)]:       # last: ncp/2 ~ 75000
        c, nu, n = laws.cir_parameters(x0, lam, xi, gam, t, strat)
        mean, sd = (n + nu) / (2 * c), np.sqrt(2 * (n + 2 * nu)) / (2 * c)
        q = np.concatenate([np.linspace(max(mean - 8 * sd, 1e-12), mean + 10 * sd, 300), [0.0, 1e-9, mean]])
        P, D = sde.pCIR(q, x0, lam, xi, gam, t, strat), sde.dCIR(q, x0, lam, xi, gam, t, strat)
        Pw, Dw = laws.pCIR(q, x0, lam, xi, gam, t, strat), laws.dCIR(q, x0, lam, xi, gam, t, strat)
        assert np.max(np.abs(P - Pw)) < 1e-11, (lam, xi, gam, t)
        ok = Dw > 1e-250
        assert np.allclose(D[ok], Dw[ok], rtol=1e-9), (lam, xi, gam, t)
        assert np.all(np.diff(P[:300]) >= -1e-15) and P[300] == 0.0
    # per-point x0 and scalars
    x0v = np.random.default_rng(0).uniform(0.2, 2.0, 50)
    got = sde.pCIR(np.full(50, 1.0), x0v, 1.0, 1.0, 1.0, 0.5)
    assert np.allclose(got, [laws.pCIR(1.0, v, 1.0, 1.0, 1.0, 0.5) for v in x0v], atol=1e-12)
    assert isinstance(sde.pOU(0.3, 0.0, 1.0, 0.0, 1.0, 1.0), float)


@pytest.mark.gpu
def test_probability_integral_transform_stays_on_the_device(sde):
    """What the device laws are for: u = pCIR(X_T) of an ensemble's terminal states, computed where they live.
    Exact sampler -> uniform; Heun on a coarse grid -> visibly not."""
    import torch
    eng = sde.default_engine(0)
    n = 400000
    dev = torch.device("cuda", 0)
    XT = torch.empty(n, dtype=torch.float64, device=dev)
    U = torch.empty_like(XT)
    x0 = torch.full((1,), 1.0, dtype=torch.float64, device=dev)
    eng.rlaw("cir", (1.0, 1.0, 1.0), 1.0, x0, n, XT, nsteps=1, stratonovich=True, seed=2)
    eng.law_eval("cir", "p", (1.0, 1.0, 1.0), 1.0, x0, XT, U, stratonovich=True)
    counts = torch.histc(U, bins=50, min=0.0, max=1.0).cpu().numpy()
    chi2 = float(((counts - n / 50) ** 2 / (n / 50)).sum())
    assert stats.chi2.sf(chi2, 49) > 1e-3
    t = np.arange(11) / 10
    eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=n, projection="abs", seed=3, XT=XT)
    eng.law_eval("cir", "p", (1.0, 1.0, 1.0), 1.0, x0, XT, U, stratonovich=True)
    counts = torch.histc(U, bins=50, min=0.0, max=1.0).cpu().numpy()
    assert stats.chi2.sf(float(((counts - n / 50) ** 2 / (n / 50)).sum()), 49) < 1e-6


@pytest.mark.gpu
def test_edge_cases_of_the_law_and_brownian_entry_points(sde):
    from sdetools_b200._lib import SdeGpuError
    # zero transitions: the start comes back; one draw; a start beyond 2^32 draws
    x0 = np.array([0.3, 1.7, 2.2])
    assert np.array_equal(sde.rCIR(3, x0, 1.0, 1.0, 1.0, 0.5, steps=0), x0)
    assert sde.rOU(1, 1.0, 1.0, 0.0, 1.0, 0.1, seed=3, path_offset=(1 << 40)).shape == (1,)
    a = sde.rCIR(64, 1.0, 1.0, 1.0, 1.0, 0.1, seed=5, path_offset=(1 << 33))
    b = sde.rCIR(64, 1.0, 1.0, 1.0, 1.0, 0.1, seed=5, path_offset=(1 << 33) + 32)
    assert np.array_equal(a[32:], b[:32]) and not np.array_equal(a[:32], b[:32])
    # CIR started at 0 stays non-negative and leaves 0 (df = 4 > 0); with xi = 0 it is absorbed (ncp = 0 and df = 0)
    z = sde.rCIR(1000, 0.0, 1.0, 1.0, 1.0, 0.1, seed=1)
    assert np.all(z >= 0) and z.max() > 0
    assert np.all(sde.rCIR(100, 0.0, 1.0, 0.0, 1.0, 0.1, seed=1) == 0.0)
    # OU with sigma = 0 is the deterministic relaxation; its density/distribution are refused (sd = 0)
    assert np.allclose(sde.rOU(4, 2.0, 1.0, 0.5, 0.0, 1.0, seed=1), 0.5 + 1.5 * np.exp(-1.0))
    with pytest.raises(SdeGpuError, match="degenerate"):
        sde.pOU(0.0, 1.0, 1.0, 0.0, 0.0, 1.0)
    # distribution functions at and beyond the ends
    assert sde.pCIR(np.array([-1.0, 0.0]), 1.0, 1.0, 1.0, 1.0, 0.5).tolist() == [0.0, 0.0]
    assert sde.pCIR(1e6, 1.0, 1.0, 1.0, 1.0, 0.5) == 1.0 and sde.dCIR(-0.5, 1.0, 1.0, 1.0, 1.0, 0.5) == 0.0
    assert sde.pOU(np.array([-1e3, 1e3]), 0.0, 1.0, 0.0, 1.0, 1.0).tolist() == [0.0, 1.0]
    # Brownian motion on a single time point: B = B0 + N(u t0, sigma^2 t0); functionals of a one-point path
    one = sde.rvBM([0.25], 5, sigma=2.0, B0=1.0, seed=9)
    F = sde.rvBM_functionals([0.25], 5, sigma=2.0, B0=1.0, level=1.0, seed=9)
    assert one.shape == (1, 5) and np.array_equal(F["max"], one[0]) and np.array_equal(F["min"], one[0])
    assert np.all(F["tmax"] == 0.25) and np.array_equal(np.isnan(F["thit"]), one[0] < 1.0)
    # times[0] = 0: the first increment has zero length, so B starts at B0 exactly
    W = sde.rvBM([0.0, 1.0, 2.0], 3, B0=-2.0, seed=2)
    assert np.all(W[0] == -2.0)
    with pytest.raises(SdeGpuError, match="non-decreasing"):
        sde.rvBM_functionals([0.0, 2.0, 1.0], 3, seed=1)


@pytest.mark.gpu
def test_fuzz_shapes_rvbm_and_functionals(sde):
    """Random small shapes through the Brownian kernels (sector stores with per-column phases): nt from 1, column counts
    around the warp size, times[0] zero or not — against the oracle's rBM fed with the device's normals, bit for bit,
    and the functionals against those of the stored matrix."""
    from oracle import sde_oracle as O
    eng = sde.default_engine(0)
    rng = np.random.default_rng(5)
    for it in range(40):
        nt = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 33]))
        n = int(rng.choice([1, 2, 5, 31, 32, 33, 100]))
        t0 = float(rng.choice([0.0, 0.3]))
        t = t0 + np.concatenate(([0.0], np.cumsum(rng.uniform(0.01, 0.2, nt - 1))))
        sigma, B0, u, seed, off = float(rng.uniform(0.5, 2)), float(rng.normal()), float(rng.normal()), int(rng.integers(1 << 30)), int(rng.integers(1 << 36))
        B = sde.rvBM(t, n, sigma=sigma, B0=B0, u=u, seed=seed, path_offset=off)
        z = eng.normals(seed, off, n, nt)
        for c in range(n):
            assert np.array_equal(B[:, c], O.rBM(t, sigma, B0, u, z=z[:, c])), (it, nt, n, c)
        lvl = B0 + 0.3
        F = sde.rvBM_functionals(t, n, sigma=sigma, B0=B0, u=u, level=lvl, seed=seed, path_offset=off)
        assert np.array_equal(F["max"], B.max(axis=0)) and np.array_equal(F["tmin"], t[B.argmin(axis=0)]), (it, nt, n)
        hit = B >= lvl
        assert np.array_equal(F["thit"], np.where(hit.any(axis=0), t[hit.argmax(axis=0)], np.nan), equal_nan=True), (it, nt, n)
