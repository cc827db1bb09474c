"""Pins the oracle: the reference's own expectations (tests/testthat/test_solvers.R)
plus the machine-precision identities its vignettes assert."""
import numpy as np
import pytest

from oracle import c_oracle, laws, philox_ref
from oracle import sde_oracle as O


# ---- tests/testthat/test_solvers.R:3-18 ------------------------------------
def test_reference_dims_nB1_and_nB2():
    times = np.arange(0, 11, dtype=float)
    nx = 3
    rng = np.random.default_rng(1)
    sol = O.euler(lambda t, x: -x, lambda t, x: x, times, np.ones(nx), rng=rng)
    assert sol["X"].shape == (len(times), nx)
    sol = O.euler(lambda t, x: -x, lambda t, x: rng.random((nx, 2)), times, np.ones(nx), rng=rng)
    assert sol["X"].shape == (len(times), nx)
    assert np.isfinite(sol["X"]).all()


# ---- tests/testthat/test_solvers.R:20-36 (tolerance = testthat expect_equal) -
def test_reference_dLinSDE_known_answers():
    tol = 1.5e-8
    sol = laws.dLinSDE(-1, 1, 1, S0=0.5)
    assert sol["eAt"].item() == pytest.approx(np.exp(-1), rel=tol)
    assert sol["St"].item() == pytest.approx(0.5, rel=tol)
    A = np.array([[0.0, 1.0], [-1.0, 0.0]])          # matrix(c(0,-1,1,0),2,2) column-major
    sol = laws.dLinSDE(A, np.eye(2), 2 * np.pi, x0=[1, 0])
    assert np.allclose(sol["EX"], [1, 0], atol=tol)
    assert np.allclose(sol["St"], 2 * np.pi * np.eye(2), atol=10 * tol)   # only via the Hamiltonian fallback
    assert laws.dLinSDE(-1, 1, 2, x0=1, u=1)["EX"][0] == pytest.approx(1, rel=tol)
    assert laws.dLinSDE(0, 1, 2, x0=1, u=np.array([[1.0, 3.0]]))["EX"][0] == pytest.approx(5, rel=tol)


def test_lyap_doc_examples():       # R/sdetools.R:587-590
    assert laws.lyap(-1, 1).item() == pytest.approx(0.5)
    A = np.array([[0.0, 1.0], [-1.0, -0.1]])         # array(c(0,-1,1,-0.1),c(2,2))
    assert np.allclose(laws.lyap(A, np.diag([0.0, 1.0])), 5 * np.eye(2))
    with pytest.raises(np.linalg.LinAlgError):
        laws.lyap(np.array([[0.0, 1.0], [-1.0, 0.0]]), np.eye(2))


# ---- vignettes/ItoIntegrals.Rmd:66-115,143-153 -------------------------------
def _ito_setup():
    t = O.r_seq(0, 1, 0.01)
    B = O.rBM(t, rng=np.random.default_rng(7))
    f = lambda x: x - x ** 3
    g = lambda x: np.sqrt(1 + x * x)
    X = O.euler(f, g, t, 0.0, B=B)["X"][:, 0]
    return t, B, f, g, X


def test_euler_equals_sum_of_ito_integrals():
    t, B, f, g, X = _ito_setup()
    assert np.max(np.abs(X - O.itointegral(f(X), t) - O.itointegral(g(X), B))) < 1e-14


def test_driver_recovered_by_ito_integration():
    t, B, f, g, X = _ito_setup()
    W = O.itointegral(1 / g(X), X) - O.itointegral(f(X) / g(X), t)
    assert np.max(np.abs(B - W)) < 1e-14


def test_stratonovich_is_ito_plus_half_crossvariation():
    t, B, f, g, X = _ito_setup()
    G = X + B
    assert np.max(np.abs(O.stochint(G, X, "c") - O.itointegral(G, X) - 0.5 * O.CrossVariation(G, X))) < 1e-14
    both = O.stochint(G, X, ["l", "c"])
    assert list(both) == ["l", "c"] and np.array_equal(both["l"], O.itointegral(G, X))


def test_stochint_doc_examples():   # R/sdetools.R:86-102
    times = np.linspace(0, 2 * np.pi, 21)
    I = O.stochint(np.cos(times), np.sin(times))
    assert np.max(np.abs(I - (0.5 * times + 0.25 * np.sin(2 * times)))) < 0.35   # left rule, 20 coarse steps
    t = O.r_seq(0, 10, 0.01)
    BM = O.rBM(t, rng=np.random.default_rng(3))
    I = O.stochint(BM, BM, ["l", "c", "r"])
    assert np.allclose(I["c"], 0.5 * BM ** 2, atol=1e-12)                       # exact telescoping
    assert np.max(np.abs(I["l"] - (0.5 * BM ** 2 - 0.5 * t))) < 0.5
    assert np.allclose(I["r"] - I["l"], O.QuadraticVariation(BM), atol=1e-12)


def test_heun_vs_euler_ito_correction():   # R/sdetools.R:222-228
    t = O.r_seq(0, 10, 0.01)
    BM = O.rBM(t, rng=np.random.default_rng(11))
    Xs = O.heun(lambda x: -0.1 * x, lambda x: x, t, 1.0, BM)["X"][:, 0]
    Xi = O.euler(lambda x: 0.4 * x, lambda x: x, t, 1.0, BM)["X"][:, 0]
    exact = np.exp(-0.1 * t + BM)
    assert np.max(np.abs(Xs / exact - 1)) < 0.06
    assert np.max(np.abs(Xi / exact - 1)) < 0.6


def test_cumsum_is_80bit():
    x = np.array([1.0, 2.0 ** -60, 2.0 ** -60, -1.0])
    assert O.r_cumsum(x)[-1] == 2.0 ** -59          # plain double cumsum would give 0
    assert np.cumsum(x)[-1] == 0.0


# ---- analytic laws -----------------------------------------------------------
def test_law_spot_values():          # SURVEY.md §8(c) restated spot values
    assert laws.dCIR(1, 1, 1, 1, 1, 1) == pytest.approx(0.5799888552024156, rel=1e-10)
    assert laws.dCIR(1, 1, 1, 1, 1, 1, Stratonovich=True) == pytest.approx(0.6184781231754529, rel=1e-10)
    m, s = laws.ou_parameters(1, 1, 1, 1, 1)
    assert m == 1 and s == pytest.approx(0.6575198539828996, rel=1e-14)
    assert laws.ou_parameters(0, 0, 0, 2, 4)[1] == 4.0


def test_dCIR_stationary_is_gamma():   # vignettes/CoxIngersollRoss.Rmd:48-71
    from scipy import stats
    lam, xi, gam = 1.3, 0.8, 0.6
    x = np.linspace(0.05, 3, 50)
    ref = stats.gamma.pdf(x, a=2 * lam * xi / gam ** 2, scale=gam ** 2 / (2 * lam))
    assert np.max(np.abs(laws.dCIR(x, 1.0, lam, xi, gam, np.inf) - ref)) < 1e-12
    assert np.allclose(laws.pCIR(laws.qCIR([0.1, 0.5, 0.9], 1, 1, 1, 1, 1), 1, 1, 1, 1, 1), [0.1, 0.5, 0.9])


def test_ou_discrete_law_limits():
    m, s = laws.ou_discrete_law("euler", 2.0, 1.3, 0.7, 0.9, 1e-5, 100000)
    m0, s0 = laws.ou_parameters(2.0, 1.3, 0.7, 0.9, 1.0)
    assert m == pytest.approx(m0, rel=1e-4) and s == pytest.approx(s0, rel=1e-4)
    # SURVEY App. D scratch numbers
    assert laws.ou_discrete_law("euler", 2.0, 1.3, 0.7, 0.9, 0.01, 100)[1] ** 2 == pytest.approx(0.29068, abs=1e-5)
    assert laws.ou_discrete_law("heun", 2.0, 1.3, 0.7, 0.9, 0.01, 100)[1] ** 2 == pytest.approx(0.28839, abs=1e-5)


# ---- Philox4x32-10 known answers (Random123 kat_vectors) ----
KATS = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.mark.parametrize("ctr,key,out", KATS)
def test_philox_kat(ctr, key, out):
    got = philox_ref.philox4x32_10(*ctr, *key)
    assert tuple(int(v) for v in got) == out
    import ctypes as C
    arr = lambda v: (C.c_uint32 * len(v))(*v)
    o = (C.c_uint32 * 4)()
    c_oracle.lib().oracle_philox4x32_10(arr(ctr), arr(key), o)
    assert tuple(o) == out


def test_normals_c_vs_numpy_and_moments():
    seed = 0x5DE70015
    z = philox_ref.normals(np.arange(2000), 101, 0, seed)
    zc = np.stack([c_oracle.normals(seed, p, 0, 101) for p in (0, 7, 1999)], axis=1)
    assert np.max(np.abs(z[:, [0, 7, 1999]] - zc)) < 1e-13
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert abs((z ** 4).mean() - 3) < 0.1


# ---- C restatement == numpy restatement, bit for bit --------------------------
CASES = [("ou", [1.3, 0.7, 0.9], [2.0], "identity"), ("cir", [1.0, 1.0, 1.0], [1.0], "abs"),
         ("cir", [0.5, 0.2, 1.5], [0.05], "relu"), ("vdp", [1.0, 0.5], [0.1, -0.2], "identity"),
         ("logistic", [0.5, 1.0], [0.3], "identity"), ("gbm", [0.4, 1.0], [1.0], "identity"),
         ("doublewell", [0.5], [0.2], "identity")]


def _paths(nt, nB, npaths, seed):
    rng = np.random.default_rng(seed)
    t = O.r_seq(0, (nt - 1) * 0.01, 0.01)
    B = np.stack([O.rvBM(t, nB, rng=rng) for _ in range(npaths)], axis=2)      # (nt, nB, npaths)
    return t, B


@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("model,params,x0,proj", CASES)
def test_c_oracle_bitwise_equals_numpy(scheme, model, params, x0, proj):
    t, B = _paths(201, 1, 16, 5)
    Xn = O.ensemble(scheme, model, params, t, x0, B=B, proj=proj)                # (nt, nx, np)
    nx = len(x0)
    Xc, XTc = c_oracle.ensemble(O_SCHEME[scheme], O.MODEL_IDS[model], params, nx, 1, O.PROJ_IDS[proj], t,
                                np.array(x0), B=np.ascontiguousarray(B.transpose(2, 1, 0)))
    assert np.array_equal(Xc.transpose(2, 1, 0), Xn)
    assert np.array_equal(XTc.T, Xn[-1])
    # and the vectorised ensemble equals the one-path reference loop
    f, g, _, _ = O.catalogue(model, params)
    fn = O.euler if scheme == "euler" else O.heun
    for p in (0, 9):
        Xp = fn(f, g, t, x0, B=B[:, :, p], p=O.projection(proj))["X"]
        assert np.array_equal(Xp, Xn[:, :, p])


O_SCHEME = {"euler": 0, "heun": 1}


@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_linear_c_vs_numpy(scheme):
    rng = np.random.default_rng(2)
    d, nB = 5, 3
    A = -np.eye(d) + 0.3 * rng.standard_normal((d, d))
    G = rng.standard_normal((d, nB))
    t, B = _paths(101, nB, 8, 9)
    x0 = rng.standard_normal(d)
    Xn = O.ensemble(scheme, "linear", (A, G), t, x0, B=B)
    params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
    Xc, _ = c_oracle.ensemble(O_SCHEME[scheme], 6, params, d, nB, 0, t, x0, B=np.ascontiguousarray(B.transpose(2, 1, 0)))
    assert np.array_equal(Xc.transpose(2, 1, 0), Xn)
    f, g, _, _ = O.catalogue("linear", (A, G))
    fn = O.euler if scheme == "euler" else O.heun
    assert np.array_equal(fn(f, g, t, x0, B=B[:, :, 3])["X"], Xn[:, :, 3])


def test_c_integrals_equal_numpy():
    rng = np.random.default_rng(4)
    a, b = rng.standard_normal((6, 300)), rng.standard_normal((6, 300)).cumsum(axis=1)
    for c in range(6):
        assert np.array_equal(c_oracle.integral(0, a, b)[c], O.itointegral(a[c], b[c]))
        assert np.array_equal(c_oracle.integral(1, a, b)[c], O.stochint(a[c], b[c], "r"))
        assert np.array_equal(c_oracle.integral(2, a, b)[c], O.stochint(a[c], b[c], "c"))
        assert np.array_equal(c_oracle.integral(3, a)[c], O.QuadraticVariation(a[c]))
        assert np.array_equal(c_oracle.integral(4, a, b)[c], O.CrossVariation(a[c], b[c]))


def test_c_fused_generator_matches_supplied_increments():
    """Philox mode of the C oracle == supplied-increment mode fed with the same normals."""
    seed, npaths, nt = 12345, 6, 52
    t = O.r_seq(0, (nt - 1) * 0.02, 0.02)
    _, XT = c_oracle.ensemble(1, 1, [1.0, 1.0, 1.0], 1, 1, 1, t, [1.0], seed=seed, path_offset=100, npaths=npaths, full=False)
    z = np.stack([c_oracle.normals(seed, 100 + p, 0, nt - 1) for p in range(npaths)], axis=1)   # (nt-1, np)
    dB = (np.sqrt(O.r_diff(t))[:, None] * z)[:, None, :]
    Xn = O.ensemble("heun", "cir", [1.0, 1.0, 1.0], t, [1.0], dB=dB, proj="abs", full=False)
    assert np.array_equal(XT.T, Xn)


def test_hist_convention():
    c = O.hist_counts([0.0, 0.5, 1.0, 1.0000001, -1e-9, 2.0, np.nan, 0.25], 0.0, 2.0, 4)
    # bins (0,.5],(.5,1],(1,1.5],(1.5,2] with 0 in the first
    assert c.tolist() == [1, 3, 1, 1, 1, 0, 1]


def test_as241_qnorm_against_scipy():
    """The oracle's inversion (Wichura AS241, the algorithm of R's qnorm / default rnorm)."""
    import ctypes as C
    from scipy.special import ndtri
    lib = c_oracle.lib()
    lib.oracle_qnorm.restype, lib.oracle_qnorm.argtypes = C.c_double, [C.c_double]
    ps = np.concatenate([np.linspace(1e-6, 1 - 1e-6, 501), 10.0 ** -np.arange(1, 17, 0.37), 2.0 ** -np.arange(2, 54)])
    got = np.array([lib.oracle_qnorm(float(p)) for p in ps])
    assert np.max(np.abs(got - ndtri(ps)) / np.maximum(np.abs(ndtri(ps)), 1e-3)) < 1e-13
