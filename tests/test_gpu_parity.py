"""GPU parity: the CUDA path (through the C ABI / host mirror) against the CPU oracle.
Bar: bit-exact for supplied-B STRICT runs, histogram counts and the integrals;
1e-12 relative for FAST arithmetic; Monte-Carlo tolerances (stated per test) for
fused-generator ensembles against the exact laws."""
import numpy as np
import pytest

from oracle import c_oracle, laws, philox_ref
from oracle import sde_oracle as O

pytestmark = pytest.mark.gpu

SCH = {"euler": 0, "heun": 1}


@pytest.fixture(scope="module")
def sde(gpu):
    import sdetools_b200
    return sdetools_b200


@pytest.fixture(scope="module")
def eng(sde):
    return sde.default_engine(0)


def brownian(nt, nB, npaths, seed, dt=0.01):
    """B as rBM builds it (80-bit cumsum of sqrt(dt)*Z), reference layout (nt, nB, npaths)."""
    rng = np.random.default_rng(seed)
    t = O.r_seq(0, (nt - 1) * dt, dt)
    z = rng.standard_normal((nt, nB * npaths))
    z[0] = 0.0
    B = O.r_cumsum_cols(np.sqrt(np.concatenate(([0.0], O.r_diff(t))))[:, None] * z)
    return t, B.reshape(nt, nB, npaths)


def oracle_X(scheme, model, params, nx, nB, proj, t, x0, B):
    """C oracle on reference-layout B (nt,nB,np) -> X (nt,nx,np)."""
    X, _ = c_oracle.ensemble(SCH[scheme], O.MODEL_IDS[model], params, nx, nB, O.PROJ_IDS[proj], t, np.asarray(x0, float),
                             B=np.ascontiguousarray(B.transpose(2, 1, 0)), nthreads=8)
    return X.transpose(2, 1, 0)


# ---- config C1: scalar OU euler(), 1000 paths x 1000 steps, supplied B, bit-level -------------
@pytest.mark.parametrize("lam,xi,sig,x0", [(1.0, 0.0, 1.0, 0.0), (1.3, 0.7, 0.9, 2.0)])
def test_c1_ou_euler_supplied_B_bit_identical(sde, lam, xi, sig, x0):
    rng = np.random.default_rng(20261019)
    t = O.r_seq(0, 10, 0.01)
    nt, P = t.size, 1000
    z = rng.standard_normal((P, nt)).T                       # drawn path-major
    z[0] = 0.0
    B = O.r_cumsum_cols(np.sqrt(np.concatenate(([0.0], O.r_diff(t))))[:, None] * z).reshape(nt, 1, P)
    f, g = sde.catalogue.ou(lam, xi, sig)
    got = sde.euler(f, g, t, x0, B)
    assert got["X"].shape == (nt, 1, P) and got["times"] is not None
    want = oracle_X("euler", "ou", [lam, xi, sig], 1, 1, "identity", t, [x0], B)
    assert np.array_equal(got["X"], want)
    # the numpy restatement of R's loop, one path per call, agrees too
    fo, go, _, _ = O.catalogue("ou", [lam, xi, sig])
    for p in (0, 499, 999):
        assert np.array_equal(O.euler(fo, go, t, x0, B=B[:, :, p])["X"], got["X"][:, :, p])
    # single-path call returns the reference's nt x nx matrix
    one = sde.euler(f, g, t, x0, B[:, 0, 7])
    assert one["X"].shape == (nt, 1) and np.array_equal(one["X"], want[:, :, 7])


CASES = [("ou", [1.3, 0.7, 0.9], [2.0], "identity"), ("cir", [1.0, 1.0, 1.0], [1.0], "abs"),
         ("cir", [0.5, 0.2, 1.5], [0.05], "relu"), ("vdp", [1.0, 0.5], [0.1, -0.2], "identity"),
         ("logistic", [0.5, 1.0], [0.3], "identity"), ("gbm", [0.4, 1.0], [1.0], "abs"),
         ("doublewell", [0.5], [0.2], "identity"), ("dwsqrt", [], [0.0], "identity")]


@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("model,params,x0,proj", CASES)
def test_all_models_strict_bit_identical_and_fast_close(eng, scheme, model, params, x0, proj):
    nx = len(x0)
    t, B = brownian(204, 1, 333, 5)                         # ragged: 333 paths = 2.6 tiles, 203 steps = 6.3 tiles
    want = oracle_X(scheme, model, params, nx, 1, proj, t, x0, B)
    Bf = np.asfortranarray(B)
    X = np.empty((t.size, nx, 333), order="F")
    XT = np.empty((nx, 333), order="F")
    eng.run(model, scheme, params, t, x0, nx=nx, npaths=333, B=Bf, projection=proj, arith="strict", X=X, XT=XT)
    assert np.array_equal(X, want)
    assert np.array_equal(XT, want[-1])
    eng.run(model, scheme, params, t, x0, nx=nx, npaths=333, B=Bf, projection=proj, arith="fast", X=X)
    assert np.allclose(X, want, rtol=1e-12, atol=1e-13)     # north star: 1e-12 relative with the same B


def test_per_path_x0_and_terminal_only(eng):
    t, B = brownian(65, 1, 130, 8)
    x0 = np.random.default_rng(1).uniform(0.5, 2.0, (1, 130))
    Xo, XTo = c_oracle.ensemble(1, 1, [1.0, 1.0, 0.7], 1, 1, 1, t, np.ascontiguousarray(x0.T),
                                B=np.ascontiguousarray(B.transpose(2, 1, 0)))
    XT = np.empty((1, 130), order="F")
    eng.run("cir", "heun", [1.0, 1.0, 0.7], t, np.ascontiguousarray(x0.T), nx=1, npaths=130, B=np.asfortranarray(B),
            projection="abs", XT=XT)
    assert np.array_equal(XT, XTo.T)


# ---- the fused generator ----------------------------------------------------------------------
KATS = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.mark.parametrize("ctr,key,out", KATS)
def test_philox_known_answers_on_device(eng, ctr, key, out):
    assert eng.philox4x32_10(ctr, key) == out


def test_device_normals_match_reference_stream(eng):
    seed, off = 0x5DE70015, (1 << 33) + 5                    # path ids beyond 32 bits
    z = eng.normals(seed, off, 257, 101)
    want = philox_ref.normals(off + np.arange(257), 101, 0, seed)
    assert np.max(np.abs(z - want)) < 1e-9                   # table quantile vs exact qnorm (ICDF_MAX_ERR 7.4e-10)
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert np.array_equal(eng.normals(seed, off + 100, 3, 7), z[:7, 100:103])     # keyed by global id
    assert not np.array_equal(eng.normals(seed, off, 4, 6, channel=1), z[:6, :4])


@pytest.mark.parametrize("model,params,x0,nx", [("ou", [1.3, 0.7, 0.9], [2.0], 1), ("vdp", [1.0, 0.5], [0.1, -0.2], 2)])
def test_tail_draws_agree_across_the_fast_kernels(eng, model, params, x0, nx):
    """A draw beyond 2^-12 in the tail leaves the fast table: the register-resident kernel replaces the block out of line, the
    trajectory kernel evaluates it on a clamped segment first and replaces the word behind the block (philox.cuh, NormalGen).
    Both must give every path that holds such a draw the value the generic evaluation (sdegpu_normals) gives it — a normal is a
    function of its own word, whichever kernel draws it."""
    P, nt, seed, off = 4096, 1030, 20261020, (1 << 35) + 17                    # 4.2e6 draws: ~1000 of them in the tail
    t = np.arange(nt) * 1e-3
    X = np.empty((nt, nx, P), order="F")
    XT = np.empty((nx, P), order="F")
    eng.run(model, "euler", params, t, x0, nx=nx, npaths=P, seed=seed, path_offset=off, X=X)
    eng.run(model, "euler", params, t, x0, nx=nx, npaths=P, seed=seed, path_offset=off, XT=XT)
    z = eng.normals(seed, off, P, nt - 1)
    from scipy.stats import norm
    zt = -norm.ppf(2.0 ** -13)                                                # |z| beyond it: u < 2^-12
    tail_paths = np.unique(np.nonzero(np.abs(z) > zt)[1])
    assert 600 < tail_paths.size < 1400, tail_paths.size
    dB = (np.sqrt(O.r_diff(t))[:, None] * z)[:, None, :]
    want = O.ensemble("euler", model, params, t, x0, dB=dB)
    assert np.allclose(X[..., tail_paths], want[..., tail_paths], rtol=1e-12, atol=1e-13)
    assert np.allclose(X, want, rtol=1e-12, atol=1e-13)
    assert np.array_equal(XT, X[-1])                                          # both kernels take the folded FAST step


@pytest.mark.parametrize("scheme,model,params,x0,proj", [("heun", "cir", [1.0, 1.0, 1.0], [1.0], "abs"),
                                                         ("euler", "vdp", [1.0, 0.5], [0.0, 0.0], "identity"),
                                                         ("euler", "ou", [1.3, 0.7, 0.9], [2.0], "identity")])
def test_fused_generator_paths_follow_oracle(eng, scheme, model, params, x0, proj):
    """Same increments (device normals) through the oracle's loop reproduce the device paths."""
    nx, P, nt, seed, off = len(x0), 200, 78, 99, 1000
    t = O.r_seq(0, (nt - 1) * 0.01, 0.01)
    X = np.empty((nt, nx, P), order="F")
    XT = np.empty((nx, P), order="F")
    eng.run(model, scheme, params, t, x0, nx=nx, npaths=P, projection=proj, seed=seed, path_offset=off, X=X)
    eng.run(model, scheme, params, t, x0, nx=nx, npaths=P, projection=proj, seed=seed, path_offset=off, XT=XT)
    z = eng.normals(seed, off, P, nt - 1)
    dB = (np.sqrt(O.r_diff(t))[:, None] * z)[:, None, :]
    want = O.ensemble(scheme, model, params, t, x0, dB=dB, proj=proj)
    assert np.allclose(X, want, rtol=1e-12, atol=1e-13)      # FAST arithmetic vs reference order
    # register-resident kernel vs trajectory kernel: same normals, but the terminal kernel's step is regrouped (folded
    # coefficients, second-order sqrt, a grid that is uniform to rounding taken as uniform), so they agree to FAST
    # tolerance, not bit for bit
    assert np.allclose(XT, X[-1], rtol=1e-10, atol=1e-12)
    # and the CPU oracle's own fused generator (libm log/sin/cos) gives the same paths to rounding
    _, XTc = c_oracle.ensemble(SCH[scheme], O.MODEL_IDS[model], params, nx, 1, O.PROJ_IDS[proj], t, np.array(x0),
                               seed=seed, path_offset=off, npaths=P, full=False)
    assert np.allclose(XT, XTc.T, rtol=1e-8, atol=1e-9)


def test_sharding_by_path_offset_is_invariant(eng):
    t = O.r_seq(0, 0.5, 0.01)
    args = dict(nx=1, projection="abs", seed=7, hist=(0.0, 4.0, 64), center=[1.0])
    full_XT, full_h, full_ps = np.empty((1, 1000), order="F"), np.zeros(67, np.uint64), np.zeros((5, 1), order="F")
    eng.run("cir", "heun", [1, 1, 1], t, [1.0], npaths=1000, XT=full_XT, hist_counts=full_h, power_sums=full_ps, **args)
    h, ps = np.zeros(67, np.uint64), np.zeros((5, 1), order="F")
    parts = []
    for lo, hi in ((0, 300), (300, 1000)):
        xt = np.empty((1, hi - lo), order="F")
        eng.run("cir", "heun", [1, 1, 1], t, [1.0], npaths=hi - lo, path_offset=lo, XT=xt, hist_counts=h, power_sums=ps,
                accumulate=True, **args)
        parts.append(xt)
    assert np.array_equal(np.concatenate(parts, axis=1), full_XT)
    assert np.array_equal(h, full_h) and h.sum() == 1000
    assert np.allclose(ps, full_ps, rtol=1e-13)


def test_terminal_statistics_match_oracle_definitions(eng):
    t = O.r_seq(0, 1, 0.02)
    P = 5000
    XT, h, ps = np.empty((2, P), order="F"), np.zeros(35, np.uint64), np.zeros((5, 2), order="F")
    eng.run("vdp", "euler", [1.0, 0.5], t, [0.5, 0.0], nx=2, npaths=P, seed=3, XT=XT, hist_counts=h, hist=(-0.6, 0.6, 32, 1),
            power_sums=ps, center=[0.3, -0.1])
    assert np.array_equal(h.astype(np.int64), O.hist_counts(XT[1], -0.6, 0.6, 32))          # integer work: bit-exact
    for j, c in enumerate((0.3, -0.1)):
        assert np.allclose(ps[:, j], O.power_sums(XT[j], c), rtol=1e-12)


# ---- statistical acceptance ---------------------------------------------------
def test_ou_ensemble_matches_exact_discrete_law(sde):
    lam, xi, sig, x0, n, h = 1.3, 0.7, 0.9, 2.0, 100, 0.01
    N = 1 << 21
    t = np.arange(n + 1) * h
    for scheme in ("euler", "heun"):
        m, s = laws.ou_discrete_law(scheme, x0, lam, xi, sig, h, n)
        fn = sde.euler if scheme == "euler" else sde.heun
        res = fn(*sde.catalogue.ou(lam, xi, sig), t, x0, npaths=N, seed=0x5DE70005, output="none", moments=True, center=[m],
                 hist=(m - 6 * s, m + 6 * s, 128))
        mom = res["moments"]
        assert abs(mom["mean"][0] - m) < 5 * s / np.sqrt(N)                       # MC error only
        assert abs(mom["var"][0] - s * s) < 5 * s * s * np.sqrt(2 / N)
        assert abs(mom["kurtosis"][0] - 3) < 5 * np.sqrt(24 / N)
        from scipy import stats
        edges = res["hist"]["breaks"]
        prob = np.diff(stats.norm.cdf(edges, m, s))
        keep = prob * N > 50
        chi2 = (((res["hist"]["counts"] - prob * N) ** 2 / (prob * N))[keep]).sum()
        assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4
        # against dOU the Euler variance is biased by ~ lam*h/2 (App. D): visible at this N, as it should be
        m0, s0 = laws.ou_parameters(x0, lam, xi, sig, n * h)
        if scheme == "euler":
            assert abs(mom["var"][0] / s0 ** 2 - 1 - lam * h / 2) < 3e-3


def test_cir_heun_histogram_matches_stratonovich_law(sde):
    """Config C2 at test size: terminal histogram vs pCIR(Stratonovich=TRUE) (R/sdetools.R:610,669-676);
    discretisation bias estimated by step halving, as SURVEY.md §8(d) prescribes."""
    from scipy import stats
    N, T = 1 << 20, 1.0
    f, g = sde.catalogue.cir(1.0, 1.0, 1.0)
    hs = {}
    for n in (500, 1000):
        hs[n] = sde.heun(f, g, np.arange(n + 1) * (T / n), 1.0, p=abs, npaths=N, seed=0x5DE70015, output="none",
                         hist=(0.0, 8.0, 64))["hist"]
    edges = hs[1000]["breaks"]
    prob = np.diff(laws.pCIR(edges, 1.0, 1.0, 1.0, 1.0, T, Stratonovich=True))
    cnt = hs[1000]["counts"].astype(float)
    bias = hs[1000]["counts"].astype(float) - hs[500]["counts"].astype(float)     # ~ h-halving difference (same seed)
    keep = prob * N > 100
    z = (cnt - prob * N)[keep] / np.sqrt((prob * N * (1 - prob))[keep] + bias[keep] ** 2)
    assert np.all(np.abs(z) < 5), z
    assert hs[1000]["over"] + hs[1000]["under"] + hs[1000]["nan"] < 1e-4 * N


# ---- integrals: bit-exact against R's 80-bit cumsum ------------------------------------------------
def test_integrals_bit_identical_to_long_double_cumsum(sde):
    rng = np.random.default_rng(4)
    n, nc = 1001, 333
    f = rng.standard_normal((n, nc)) * np.exp(rng.uniform(-20, 20, (1, nc)))
    g = np.cumsum(rng.standard_normal((n, nc)), axis=0) * 0.1
    fa, ga = np.ascontiguousarray(f.T), np.ascontiguousarray(g.T)
    res = sde.stochint(f, g, ["l", "r", "c"])
    assert np.array_equal(res["l"], c_oracle.integral(0, fa, ga).T)
    assert np.array_equal(res["r"], c_oracle.integral(1, fa, ga).T)
    assert np.array_equal(res["c"], c_oracle.integral(2, fa, ga).T)
    assert np.array_equal(sde.itointegral(f, g), res["l"])
    assert np.array_equal(sde.QuadraticVariation(g), c_oracle.integral(3, ga).T)
    assert np.array_equal(sde.CrossVariation(f, g), c_oracle.integral(4, fa, ga).T)
    # the reference's one-vector calls, against the numpy restatement
    assert np.array_equal(sde.stochint(f[:, 0], g[:, 0], "c"), O.stochint(f[:, 0], g[:, 0], "c"))
    assert np.array_equal(sde.itointegral(f[:, 1], g[:, 1]), O.itointegral(f[:, 1], g[:, 1]))


def test_integrals_edge_cases(sde):
    assert sde.itointegral(np.array([3.0]), np.array([1.0])).tolist() == [0.0]          # n = 1 -> c(0)
    assert sde.itointegral(np.array([3.0, 9.0]), np.array([1.0, 1.5])).tolist() == [0.0, 1.5]
    x = np.array([1.0, 2.0 ** -60 + 1.0, np.inf, 1.0, np.nan, 2.0])
    got, want = sde.QuadraticVariation(x), O.QuadraticVariation(x)
    assert np.array_equal(got, want, equal_nan=True)
    # a sum a plain-double prefix sum gets wrong: 1 + 2^-60 + 2^-60 - 1
    G = np.array([1.0, 2.0 ** -60, 2.0 ** -60, -1.0, 0.0])
    assert sde.itointegral(G, np.arange(5.0))[-1] == 2.0 ** -59 and np.cumsum(G)[-2] == 0.0


def test_cumsum_accumulator_adversarial_columns(sde):
    """The accumulator's floating-point fast path and its integer fallback (float80.cuh) on columns built to
    hit exact 64-bit ties, sums just below powers of two, cancellation, subnormals, near-overflow and
    Inf/NaN — against numpy's long double cumsum (the x87 FPU), bit for bit.  G*dB with B = 0,1,2,.. is G."""
    rng = np.random.default_rng(80)
    n = 600
    cols = []
    for kind in range(12):
        v = rng.standard_normal(n)
        if kind == 1: v = np.ldexp(v, rng.integers(-60, 60, n))
        if kind == 2: v = np.ldexp(v, rng.integers(-1070, 1000, n))                 # leaves the fast path's range
        if kind == 3: v[1::2] = -v[0::2] * (1 + 2.0 ** -52 * rng.integers(0, 9, n // 2))
        if kind == 4: v = np.ldexp(rng.integers(1, 2 ** 53, n).astype(float), rng.integers(-113, 7, n)) * rng.choice([-1, 1], n)
        if kind == 5: v = np.where(rng.random(n) < 0.2, rng.choice([0.0, -0.0], n), np.ldexp(1.0, rng.integers(-60, 60, n)))
        if kind == 6: v = np.ldexp(1.0, -(np.arange(n) % 120)) * np.where(np.arange(n) % 3, 1.0, -1.0)
        if kind == 7: v = v * v * 1e-4
        if kind == 8:                                                                # ties of the running sum, broken or not
            cur = np.longdouble(0)
            for i in range(n):
                if i % 3 == 0 or cur == 0:
                    x = float(np.ldexp(float(rng.integers(1, 2 ** 53) | 1), int(rng.integers(-70, -30))))
                else:
                    e = int(np.frexp(np.float64(cur))[1])
                    x = float(np.ldexp(1.0, e - 65)) * float(2 * rng.integers(0, 9) - 7)
                    x += [0.0, 1.0, -1.0][int(rng.integers(0, 3))] * float(np.ldexp(1.0, e - 65 - 20 - int(rng.integers(0, 30))))
                v[i] = x
                cur = cur + np.longdouble(x)
        if kind == 9:                                                                # crossing powers of two
            cur = np.longdouble(0)
            for i in range(n):
                if i % 4 == 0 or cur == 0:
                    x = float(np.ldexp(1.0 + float(rng.integers(0, 4096)) * 2.0 ** -52, int(rng.integers(-20, 20))))
                else:
                    e = int(np.frexp(np.float64(cur))[1])
                    p2 = np.ldexp(np.longdouble(1), e - 1 + int(rng.integers(0, 2)))
                    tgt = (-p2 if cur < 0 else p2) + np.ldexp(np.longdouble(int(rng.integers(-3, 4))), e - 66 + int(rng.integers(0, 3)))
                    x = float(tgt - cur)
                v[i] = x
                cur = cur + np.longdouble(x)
        if kind == 10: v[[50, 300]] = [np.inf, -np.inf]
        if kind == 11: v = np.concatenate(([5e-324, 5e-324, -5e-324, 1e308, 1e308, 1e308, -1e308, 4.9e-324], v[8:]))
        cols.append(v)
    G = np.stack(cols, axis=1)                                   # (n, 12)
    G = np.concatenate([G, np.zeros((1, G.shape[1]))])           # itointegral uses G[:-1]
    Bt = np.repeat(np.arange(n + 1.0)[:, None], G.shape[1], axis=1)
    with np.errstate(all="ignore"):
        got = sde.itointegral(G, Bt)
        want = np.stack([O.itointegral(G[:, c], Bt[:, c]) for c in range(G.shape[1])], axis=1)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isinf(want[:, 2]).any() or np.abs(want[:, 2]).max() > 1e290          # the wide column did leave the fast range


def test_vignette_identities_on_device(sde):
    """vignettes/ItoIntegrals.Rmd:82-115,143-153 end to end on the GPU path."""
    t, B = brownian(101, 1, 64, 7)
    f, g = sde.catalogue.doublewell(0.5)
    X = sde.euler(f, g, t, 0.0, B)["X"][:, 0, :]
    Bm = B[:, 0, :]
    tt = np.repeat(t[:, None], 64, axis=1)
    resid = X - sde.itointegral(X - X * X * X, tt) - sde.itointegral(np.full_like(X, 0.5), Bm)
    assert np.max(np.abs(resid)) < 1e-14
    G = X + Bm
    assert np.max(np.abs(sde.stochint(G, X, "c") - sde.itointegral(G, X) - 0.5 * sde.CrossVariation(G, X))) < 1e-14


def test_rvbm_matches_reference_construction(sde, eng):
    t = np.concatenate(([0.5], 0.5 + np.cumsum(np.random.default_rng(2).uniform(0.01, 0.1, 200))))
    B = sde.rvBM(t, 300, sigma=1.7, B0=0.3, u=-0.2, seed=11, path_offset=40)
    z = eng.normals(11, 40, 300, t.size)
    for c in (0, 150, 299):
        assert np.array_equal(B[:, c], O.rBM(t, 1.7, 0.3, -0.2, z=z[:, c]))
    # BrownianMotion.Rmd:38-48: cov(B_s, B_t) ~ min(s,t)
    tt = np.arange(1, 11) * 0.1
    W = sde.rvBM(tt, 200000, seed=5)
    assert abs(np.var(W[-1]) - 1.0) < 0.02 and abs(np.mean(W[4] * W[9]) - 0.5) < 0.02
    bb = sde.rBrownianBridge(tt, (1.0, -1.0), seed=3)
    assert bb[0] == pytest.approx(1.0) and bb[-1] == pytest.approx(-1.0)


# ---- linear SDE -----------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_linear_strict_bit_identical(sde, scheme):
    rng = np.random.default_rng(2)
    d, nB, P = 5, 3, 70
    A = -np.eye(d) + 0.3 * rng.standard_normal((d, d))
    G = rng.standard_normal((d, nB))
    t, B = brownian(52, nB, P, 9)
    x0 = rng.standard_normal(d)
    want = O.ensemble(scheme, "linear", (A, G), t, x0, B=B)
    fn = sde.euler if scheme == "euler" else sde.heun
    got = fn(*sde.catalogue.linear(A, G), t, x0, B)["X"]
    assert np.array_equal(got, want)
    fast = fn(*sde.catalogue.linear(A, G), t, x0, B, arith="fast")["X"]
    assert np.allclose(fast, want, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("d,nB", [(1, 1), (2, 2), (2, 1), (3, 2), (4, 4), (4, 3),          # k_linear_small
                                  (5, 1), (7, 7), (8, 8), (12, 5), (16, 16), (20, 7), (32, 32), (32, 1),   # k_linear_warp
                                  (6, 9), (40, 3), (200, 3)])                             # k_linear_generic (200,3: odd smem split)
def test_small_and_medium_linear_systems(sde, eng, scheme, d, nB):
    """d, nB <= 4 (the noisy oscillator's size) run one thread per path with A, G, x in registers; d <= 32 one lane
    per (path, row) with shuffles; the rest in the shared-memory kernel.  Each: bit-identical to the oracle with a
    supplied B, the oracle's own fused generator to the quantile table's accuracy, moments = those of the returned
    terminal states."""
    rng = np.random.default_rng(10 * d + nB)
    A = -0.6 * np.eye(d) + 0.4 * rng.standard_normal((d, d)) / np.sqrt(d)
    G = rng.standard_normal((d, nB))
    x0 = rng.standard_normal(d)
    t, B = brownian(77, nB, 150, 3)
    fn = sde.euler if scheme == "euler" else sde.heun
    want = O.ensemble(scheme, "linear", (A, G), t, x0, B=B)
    got = fn(*sde.catalogue.linear(A, G), t, x0, B)
    assert np.array_equal(got["X"], want) and np.array_equal(got["XT"], want[-1])
    assert np.allclose(fn(*sde.catalogue.linear(A, G), t, x0, B, arith="fast")["X"], want, rtol=1e-12, atol=1e-13)
    # fused generator against the C oracle's (same Philox stream, libm quantile instead of the table)
    P, seed, off = 1003, 77, 1 << 34
    params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
    _, XTc = c_oracle.ensemble(SCH[scheme], 6, params, d, nB, 0, t, x0, seed=seed, path_offset=off, npaths=P, full=False, nthreads=8)
    XT = np.empty((d, P), order="F")
    ps = np.zeros((5, d), order="F")
    eng.run("linear", scheme, params, t, x0, nx=d, nB=nB, npaths=P, seed=seed, path_offset=off, XT=XT, power_sums=ps,
            center=np.zeros(d))
    assert np.allclose(XT, XTc.T, rtol=1e-7, atol=1e-7)
    assert np.array_equal(ps[0], np.full(d, P)) and np.allclose(ps[1], XT.sum(axis=1), rtol=1e-10, atol=1e-9)
    assert np.allclose(ps[2], (XT ** 2).sum(axis=1), rtol=1e-10)
    # and the trajectory variant ends where the terminal-only variant does
    X = fn(*sde.catalogue.linear(A, G), t, x0, npaths=64, seed=seed, path_offset=off)["X"]
    tol = 1e-10 if d > 32 else 1e-13          # d > 32, euler: the terminal-only run is a tensor-core GEMM (other summation order)
    assert np.allclose(X[-1], XT[:, :64], rtol=tol, atol=tol * 1e-2)


def test_reference_dims_test_on_device(sde):
    """tests/testthat/test_solvers.R:3-18 with catalogue closures: dim(X) = c(nt, nx), nB = 1 and nB = 2."""
    times = np.arange(0, 11, dtype=float)
    sol = sde.euler(*sde.catalogue.linear(-np.eye(3), np.ones((3, 1))), times, np.ones(3), seed=1)
    assert sol["X"].shape == (11, 3)
    sol = sde.euler(*sde.catalogue.linear(-np.eye(3), np.random.default_rng(0).random((3, 2))), times, np.ones(3), seed=1)
    assert sol["X"].shape == (11, 3) and np.isfinite(sol["X"]).all()


def test_linear_oscillator_covariance_vs_dLinSDE(sde):
    """vignettes/NoisyOscillator.Rmd:38-54,94-97: ensemble mean/cov vs dLinSDE and the exact Euler law."""
    lam, om, s = 0.2, 1.0, 0.5
    A, G = np.array([[-lam, -om], [om, -lam]]), s * np.eye(2)
    n, h, N = 400, 0.01, 1 << 18
    x0 = np.array([1.0, 0.0])
    res = sde.euler(*sde.catalogue.linear(A, G), np.arange(n + 1) * h, x0, npaths=N, seed=42, output="terminal")
    XT = res["XT"]
    m, S = laws.linear_discrete_law(A, G, x0, h, n)
    assert np.all(np.abs(XT.mean(axis=1) - m) < 5 * np.sqrt(np.diag(S) / N))
    assert np.all(np.abs(np.cov(XT) - S) < 5 * np.sqrt(2 / N) * np.sqrt(np.outer(np.diag(S), np.diag(S))) + 1e-12)
    ref = laws.dLinSDE(A, G, n * h, x0=x0)
    assert np.allclose(XT.mean(axis=1), ref["EX"], atol=0.02) and np.allclose(np.cov(XT), ref["St"], atol=0.02)   # O(h) Euler bias


# ---- error behaviour ----------------------------------------------------------------------------------
def test_argument_errors(sde, eng):
    from sdetools_b200._lib import SdeGpuError
    f, g = sde.catalogue.ou()
    with pytest.raises(SdeGpuError, match="at least 2 time points"):
        sde.euler(f, g, [0.0], 0.0)
    with pytest.raises(SdeGpuError, match="non-decreasing"):
        sde.euler(f, g, [0.0, 1.0, 0.5], 0.0)
    with pytest.raises(SdeGpuError, match="STRICT"):
        sde.euler(f, g, [0.0, 1.0, 2.0], 0.0, arith="strict")
    with pytest.raises(ValueError):
        sde.euler(f, g, [0.0, 1.0, 2.0], [0.0, 1.0])
    with pytest.raises(TypeError):
        sde.euler(f, g, [0.0, 1.0, 2.0], 0.0, h=lambda t, x: 1)
    assert sde.euler(f, g, [0.0, 1.0], 0.5, np.array([0.0, 2.0]))["X"].tolist() == [[0.5], [0.5 + -0.5 * 1.0 + 2.0]]   # nt = 2
    eng.run("ou", "euler", [1, 0, 1], [0.0, 1.0, 2.0], [0.0], nx=1, npaths=0)                  # empty ensemble is a no-op


# ---- high-dimensional linear SDE on the FP64 tensor path (DMMA) ------------------------------------------
@pytest.mark.parametrize("force_dmma", [False, True])
@pytest.mark.parametrize("d", [32, 40, 96, 100, 250, 256])
def test_linear_dmma_matches_generic_kernel_and_oracle(eng, d, force_dmma, monkeypatch):
    """(force_dmma: 32 < d <= 128 is routed to k_linear_tc with several path tiles per CTA by default; SDEGPU_LINEAR_DMMA=1 keeps
    the tuned C4 kernel k_linear_dmma on those sizes so that both stay covered.)
    Config C4 at test size: dX = AX dt + diag(g) dB, euler, fused generator.  The DMMA kernel (terminal
    output; any 32 < d <= 256, zero-padded to a multiple of 32 rows) must agree with the generic kernel (taken
    when the trajectory is requested) to summation order, and with the CPU oracle's own fused generator to the
    quantile table's accuracy.  (d = 32 runs the warp-shuffle kernel.)"""
    if force_dmma:
        monkeypatch.setenv("SDEGPU_LINEAR_DMMA", "1")
    rng = np.random.default_rng(d)
    R = rng.standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    g = rng.uniform(0.5, 1.5, d)
    params = np.concatenate([A.ravel(order="F"), np.diag(g).ravel(order="F")])
    P, nt, seed, off = 150, 41, 77, 12345                       # 150 paths: two full 64-path tiles + a ragged one
    t = O.r_seq(0, (nt - 1) * 0.01, 0.01)
    x0 = rng.standard_normal(d)
    XT = np.empty((d, P), order="F")
    eng.run("linear", "euler", params, t, x0, nx=d, nB=d, npaths=P, seed=seed, path_offset=off, XT=XT)
    X = np.empty((nt, d, P), order="F")
    eng.run("linear", "euler", params, t, x0, nx=d, nB=d, npaths=P, seed=seed, path_offset=off, X=X)
    assert np.allclose(XT, X[-1], rtol=1e-11, atol=1e-12)
    _, XTc = c_oracle.ensemble(0, 6, params, d, d, 0, t, x0, seed=seed, path_offset=off, npaths=P, full=False, nthreads=8)
    assert np.allclose(XT, XTc.T, rtol=1e-7, atol=1e-8)
    # per-path x0 and power sums through the same kernel
    x0p = rng.standard_normal((P, d))
    ps = np.zeros((5, d), order="F")
    eng.run("linear", "euler", params, t, np.ascontiguousarray(x0p), nx=d, nB=d, npaths=P, seed=seed, XT=XT, power_sums=ps,
            center=np.zeros(d))
    assert np.allclose(ps[1], XT.sum(axis=1), rtol=1e-10, atol=1e-10) and np.all(ps[0] == P)


def test_linear_dmma_covariance_matches_lyapunov_law(eng):
    """A + A' = -I, G = s I  =>  lyap(A, s^2 I) = s^2 I and St = s^2 (1 - exp(-t)) I;
    the ensemble is compared with the exact Euler-discrete covariance (MC error only)."""
    d, s, n, h, P = 64, 0.7, 200, 0.005, 1 << 15
    rng = np.random.default_rng(1)
    R = rng.standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    G = s * np.eye(d)
    x0 = np.zeros(d)
    x0[0] = 1.0
    XT = np.empty((d, P), order="F")
    eng.run("linear", "euler", np.concatenate([A.ravel(order="F"), G.ravel(order="F")]), np.arange(n + 1) * h, x0, nx=d, nB=d,
            npaths=P, seed=5, XT=XT)
    m, S = laws.linear_discrete_law(A, G, x0, h, n)
    C = np.cov(XT)
    sd = np.sqrt(np.diag(S))
    assert np.all(np.abs(XT.mean(axis=1) - m) < 5.5 * sd / np.sqrt(P))
    assert np.max(np.abs(C - S) / np.outer(sd, sd)) < 6 * np.sqrt(2 / P)
    assert np.allclose(np.diag(S), s * s * (1 - np.exp(-n * h)), rtol=0.02)         # dLinSDE/lyap closed form, O(h) bias
    assert np.allclose(laws.lyap(A[:4, :4] * 0 - 0.5 * np.eye(4), s * s * np.eye(4)), s * s * np.eye(4))


# ---- mortality and stopping (R/sdetools.R:397-416,455-469; Killing.Rmd:152-207) ---------
def test_killing_matches_reference_loop(sde):
    """euler(f, g, times, x0, B, r=r, Stau=Stau) path by path against the numpy restatement of R's loop:
    tau and the trajectory up to tau are bit-identical, S agrees to exp()'s last ulp."""
    t, B = brownian(601, 1, 96, 21)
    f, g = sde.catalogue.ou(1.0, 0.0, 1.0)                   # Killing.Rmd: f = -x, g = 1, r = 0.5*exp(2x)
    r = sde.catalogue.kill_exp(0.5, 2.0)
    Stau = np.random.default_rng(3).random(96)
    got = sde.euler(f, g, t, 0.0, B, r=r, Stau=Stau)
    fo, go, _, _ = O.catalogue("ou", [1.0, 0.0, 1.0])
    ro = lambda x: 0.5 * np.exp(2.0 * x[0])
    assert got["X"].shape == (601, 1, 96) and got["tau"].shape == (96,)
    n_killed = 0
    for p in range(96):
        ref = O.euler(fo, go, t, 0.0, B=B[:, :, p], r=ro, Stau=Stau[p])
        k = ref["times"].size                                 # rows kept by the reference (1:(i+1))
        assert got["tau"][p] == ref["tau"]
        assert np.array_equal(got["X"][:k, 0, p], np.atleast_1d(ref["X"]).ravel())
        assert np.all(np.isnan(got["X"][k:, 0, p]))
        assert got["S"][p] == pytest.approx(ref["S"][-1], rel=1e-13)
        assert got["XT"][0, p] == np.atleast_1d(ref["X"]).ravel()[-1]
        # the survival path: the reference's S[1:(i+1)] (:443-444,456,468), NaN after tau
        assert np.allclose(got["S_path"][:k, p], ref["S"], rtol=1e-13, atol=0) and got["S_path"][0, p] == 1.0
        assert np.all(np.isnan(got["S_path"][k:, p]))
        n_killed += ref["tau"] < t[-1]
    assert 20 < n_killed < 96                                 # both outcomes are exercised
    # one path: the reference's own return shape list(times[1:(i+1)], X[1:(i+1),], S[1:(i+1)], tau)  (:466-468)
    p = int(np.argmin(got["tau"]))
    one = sde.euler(f, g, t, 0.0, B[:, :, p], r=r, Stau=float(Stau[p]))
    ref = O.euler(fo, go, t, 0.0, B=B[:, :, p], r=ro, Stau=Stau[p])
    assert one["tau"] == ref["tau"] and np.array_equal(one["times"], ref["times"])
    assert np.array_equal(one["X"], np.atleast_1d(ref["X"]).ravel()) and np.allclose(one["S"], ref["S"], rtol=1e-13, atol=0)


def test_survival_path_with_stopping_reads_zero_at_the_stop(sde):
    """h and r together (R/sdetools.R:455-457): a path stopped by h keeps S = 0 at tau (the reference's S <- numeric(nt)
    is never written there), a killed one its last survival; both NaN afterwards."""
    t, B = brownian(401, 1, 80, 23)
    f, g = sde.catalogue.ou(1.0, 0.0, 1.0)
    r, h = sde.catalogue.kill_exp(0.5, 2.0), sde.catalogue.stop_outside(-0.4, 0.9)
    Stau = np.random.default_rng(4).random(80)
    for scheme in ("euler", "heun"):
        got = getattr(sde, scheme)(f, g, t, 0.1, B, r=r, h=h, Stau=Stau)
        fo, go, _, _ = O.catalogue("ou", [1.0, 0.0, 1.0])
        nstop = 0
        for p in range(80):
            ref = getattr(O, scheme)(fo, go, t, 0.1, B=B[:, :, p], r=lambda x: 0.5 * np.exp(2.0 * x[0]),
                                     h=lambda tt, x: (x[0] + 0.4) * (0.9 - x[0]), Stau=Stau[p])
            k = ref["times"].size
            assert got["tau"][p] == ref["tau"]
            assert np.allclose(got["S_path"][:k, p], ref["S"], rtol=1e-13, atol=0)
            assert np.all(np.isnan(got["S_path"][k:, p]))
            nstop += ref["S"][-1] == 0.0
        assert 5 < nstop < 80


def test_stopping_rule_matches_reference_loop(sde):
    t, B = brownian(301, 1, 64, 22)
    f, g = sde.catalogue.gbm(0.1, 0.6)
    h = sde.catalogue.stop_outside(0.7, 1.5)
    got = sde.heun(f, g, t, 1.0, B, h=h)
    fo, go, _, _ = O.catalogue("gbm", [0.1, 0.6])
    stops = 0
    for p in range(64):
        ref = O.heun(fo, go, t, 1.0, B=B[:, :, p], h=lambda tt, x: (x[0] - 0.7) * (1.5 - x[0]))
        X = ref["X"][:, 0]                                    # r = NULL: full length, NA after the stop (:439,:464)
        k = int(np.sum(~np.isnan(X)))
        assert np.array_equal(got["X"][:k, 0, p], X[:k]) and np.all(np.isnan(got["X"][k:, 0, p]))
        assert got["tau"][p] == t[k - 1]
        stops += k < t.size
    assert 10 < stops <= 64 and "S" not in got


def test_constant_killing_rate_gives_exponential_lifetimes(sde):
    a, T, n, N = 0.7, 2.0, 200, 1 << 18
    t = np.arange(n + 1) * (T / n)
    res = sde.euler(*sde.catalogue.ou(1.0, 0.0, 1.0), t, 0.0, r=sde.catalogue.kill_const(a), npaths=N, seed=9,
                    output="terminal")
    tau = res["tau"]
    for tk in (0.5, 1.0, 1.5):
        pk = np.exp(-a * tk)                                  # P(tau > t_k) = P(Stau <= exp(-a t_k))
        assert abs(np.mean(tau > tk + 1e-12) - pk) < 4.5 * np.sqrt(pk * (1 - pk) / N)
    alive = tau == t[-1]
    assert abs(alive.mean() - np.exp(-a * T)) < 4.5 * np.sqrt(np.exp(-a * T) / N)
    assert np.all(res["S"][~alive] < 1) and res["XT"].shape == (1, N)


def test_lane_refill_kernel_equals_one_path_per_thread(sde):
    """Without a trajectory the kill/stop kernel lets a lane whose path ended take the warp's next path
    (k_killed_refill).  Per path it must give exactly what the one-path-per-thread kernel (taken when X is
    requested) gives; the run is reproducible although lanes pick paths dynamically."""
    f, g = sde.catalogue.ou(1.0, 0.0, 1.0)
    t = np.arange(301) * 0.01
    kw = dict(r=sde.catalogue.kill_quad(2.0, 0.0), h=sde.catalogue.stop_above(1.2), seed=44)
    N = 1_000_000                                               # ~280 paths per warp: every lane refills many times
    big = sde.euler(f, g, t, 0.3, npaths=N, output="terminal", moments=True, center=[0.0], hist=(-3.0, 3.0, 64), **kw)
    lifetimes = big["tau"]
    assert 0.05 < np.mean(lifetimes < t[-1]) < 0.999 and np.unique(lifetimes).size > 100
    for off in (0, 123_457, N - 500):
        ref = sde.euler(f, g, t, 0.3, npaths=500, path_offset=off, output="path", **kw)       # k_killed (stores X)
        sl = slice(off, off + 500)
        assert np.array_equal(ref["tau"], big["tau"][sl]) and np.array_equal(ref["S"], big["S"][sl])
        assert np.array_equal(ref["XT"], big["XT"][:, sl])
    again = sde.euler(f, g, t, 0.3, npaths=N, output="terminal", moments=True, center=[0.0], hist=(-3.0, 3.0, 64), **kw)
    assert np.array_equal(again["power_sums"], big["power_sums"]) and np.array_equal(again["hist"]["counts"], big["hist"]["counts"])
    assert big["hist"]["counts"].sum() + big["hist"]["under"] + big["hist"]["over"] + big["hist"]["nan"] == N
    # supplied B (STRICT): enough paths that a warp's chunk exceeds its 32 lanes
    rng = np.random.default_rng(5)
    tb, P = np.arange(102) * 0.02, 400_000
    B = np.cumsum(np.concatenate([np.zeros((1, P)), rng.standard_normal((101, P)) * np.sqrt(0.02)]), axis=0)[:, None, :]
    Stau = rng.random(P)
    a = sde.euler(f, g, tb, 0.3, B, output="terminal", Stau=Stau, r=kw["r"], h=kw["h"])
    b = sde.euler(f, g, tb, 0.3, B, output="path", Stau=Stau, r=kw["r"], h=kw["h"])
    assert np.array_equal(a["tau"], b["tau"]) and np.array_equal(a["S"], b["S"]) and np.array_equal(a["XT"], b["XT"])
    assert np.unique(a["tau"]).size > 50


# ---- BASELINE.json full sizes, through size-independent properties ---------------------------------------
def test_c2_full_size_histogram_and_moments(sde):
    """Config C2 as benchmarked: CIR heun(), 10^7 paths x 10^4 steps.  Properties: every path is counted once,
    terminal mean/variance match the Stratonovich CIR law (noncentral chi-square) within MC error (the O(h)
    bias at h = 1e-4 is below it), the histogram passes a chi-square test against pCIR(Stratonovich=TRUE)."""
    from scipy import stats
    N, n = 10 ** 7, 10 ** 4
    t = np.arange(n + 1) * (1.0 / n)
    res = sde.heun(*sde.catalogue.cir(1.0, 1.0, 1.0), t, 1.0, p=abs, npaths=N, seed=0x5DE70015, output="none", moments=True,
                   center=[1.0], hist=(0.0, 8.0, 256))
    h = res["hist"]
    assert int(h["counts"].sum()) + h["under"] + h["over"] + h["nan"] == N and h["nan"] == 0 and h["under"] == 0
    c, nu, dof = laws.cir_parameters(1.0, 1.0, 1.0, 1.0, 1.0, Stratonovich=True)
    mean, var = (dof + nu) / (2 * c), 2 * (dof + 2 * nu) / (2 * c) ** 2
    assert abs(res["moments"]["mean"][0] - mean) < 5 * np.sqrt(var / N)
    assert abs(res["moments"]["var"][0] / var - 1) < 5 * np.sqrt((res["moments"]["kurtosis"][0] - 1) / N) + 2e-4
    prob = np.diff(laws.pCIR(h["breaks"], 1.0, 1.0, 1.0, 1.0, 1.0, Stratonovich=True))
    keep = prob * N > 200
    chi2 = (((h["counts"] - prob * N) ** 2 / (prob * N))[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4, chi2


def test_c5_share_ou_against_exact_discrete_law(sde):
    """One GPU's share of config C5 (OU euler, 10^3 steps): 5*10^7 paths against the scheme's exact Gaussian law
     — at this N the variance is resolved to 2e-4, finer than the O(lambda h) bias against
    dOU, which must therefore be visible with the right sign and size."""
    lam, xi, sig, x0, n, T, N = 1.3, 0.7, 0.9, 2.0, 1000, 1.0, 5 * 10 ** 7
    m, s = laws.ou_discrete_law("euler", x0, lam, xi, sig, T / n, n)
    res = sde.euler(*sde.catalogue.ou(lam, xi, sig), np.arange(n + 1) * (T / n), x0, npaths=N, seed=0x5DE70005, output="none",
                    moments=True, center=[m], hist=(m - 6 * s, m + 6 * s, 512))
    mom = res["moments"]
    assert mom["n"][0] == N
    assert abs(mom["mean"][0] - m) < 5 * s / np.sqrt(N)
    assert abs(mom["var"][0] / s ** 2 - 1) < 5 * np.sqrt(2 / N)
    assert abs(mom["skewness"][0]) < 5 * np.sqrt(6 / N) and abs(mom["kurtosis"][0] - 3) < 5 * np.sqrt(24 / N)
    s0 = laws.ou_parameters(x0, lam, xi, sig, T)[1]
    bias = s ** 2 / s0 ** 2 - 1                                  # O(lambda h) discretisation bias against dOU (9e-4 here)
    assert bias > 3 * np.sqrt(2 / N) and mom["var"][0] / s0 ** 2 - 1 == pytest.approx(bias, abs=5 * np.sqrt(2 / N))
    from scipy import stats
    prob = np.diff(stats.norm.cdf(res["hist"]["breaks"], m, s))
    keep = prob * N > 200
    chi2 = (((res["hist"]["counts"] - prob * N) ** 2 / (prob * N))[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4, chi2


def test_c3_trajectory_properties_at_scale(eng):
    """Config C3 at 2 GB: the trajectory kernel's last row equals the register-resident kernel's terminal
    state, the first row is x0, and an arbitrary sub-range run with path_offset reproduces its slice."""
    import torch
    P, nt = 13000, 10001                                       # odd nt, ragged last wave
    t = np.arange(nt) * 0.01
    dev = torch.device("cuda", 0)
    X = torch.empty(P * 2 * nt, dtype=torch.float64, device=dev)
    XT = torch.empty(P * 2, dtype=torch.float64, device=dev)
    eng.run("vdp", "euler", [1.0, 0.5], t, [0.3, -0.2], nx=2, npaths=P, seed=11, X=X)
    eng.run("vdp", "euler", [1.0, 0.5], t, [0.3, -0.2], nx=2, npaths=P, seed=11, XT=XT)
    eng.synchronize()
    Xv = X.view(P, 2, nt)
    assert torch.all(Xv[:, 0, 0] == 0.3) and torch.all(Xv[:, 1, 0] == -0.2) and bool(torch.isfinite(Xv).all())
    assert torch.allclose(Xv[:, :, -1], XT.view(P, 2), rtol=1e-9, atol=1e-11)   # terminal kernel: folded step, FAST tolerance
    sub = torch.empty(37 * 2 * nt, dtype=torch.float64, device=dev)
    eng.run("vdp", "euler", [1.0, 0.5], t, [0.3, -0.2], nx=2, npaths=37, seed=11, path_offset=5003, X=sub)
    eng.synchronize()
    assert torch.equal(sub.view(37, 2, nt), Xv[5003:5040])


# ---- integrals accumulated in the step loop (no trajectory stored) -------------------------------------
@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("model,params,x0,proj", [CASES[1], CASES[3], CASES[6]])
def test_fused_path_integrals_equal_integral_functions_on_the_stored_path(sde, scheme, model, params, x0, proj):
    """With a supplied B the fused functionals are the terminal entries of QuadraticVariation / itointegral /
    stochint(.., "c") / itointegral(X, times) evaluated on the stored trajectory — bit for bit (same terms, same
    80-bit accumulator), for both arithmetic modes."""
    fn = sde.euler if scheme == "euler" else sde.heun
    f, g = getattr(sde.catalogue, model)(*params)
    p = {"identity": None, "abs": abs, "relu": "relu"}[proj]
    t, B = brownian(157, 1, 70, 12)
    for arith in ("strict", "fast"):
        path = fn(f, g, t, x0, B, p=p, arith=arith)["X"]                      # (nt, nx, P)
        got = fn(f, g, t, x0, B, p=p, arith=arith, integrals=True)
        assert "X" not in got and np.array_equal(got["XT"], path[-1])
        nx, P = path.shape[1], path.shape[2]
        for j in range(nx):
            Xj, Bj = path[:, j, :], B[:, 0, :]
            T = np.repeat(t[:, None], P, axis=1)
            assert np.array_equal(got["integrals"]["QV"][j], sde.QuadraticVariation(Xj)[-1])
            assert np.array_equal(got["integrals"]["ito"][j], sde.itointegral(Xj, Bj)[-1])
            assert np.array_equal(got["integrals"]["strat"][j], sde.stochint(Xj, Bj, "c")[-1])
            assert np.array_equal(got["integrals"]["time"][j], sde.itointegral(Xj, T)[-1])


def test_fused_path_integrals_with_the_generator(sde):
    """Fused generator (no B to compare with): the ensemble means of the four functionals of an OU path against
    their closed forms — E[QV] = sigma^2 T, E[strat - ito] = sigma T / 2 (half the cross-variation [X,B]_T,
    vignettes/ItoIntegrals.Rmd:143-153) and E int X dt."""
    lam, xi, sig, x0, n, nsteps = 1.3, 0.7, 0.9, 2.0, 50000, 1000
    t = np.arange(nsteps + 1) / nsteps
    f, g = sde.catalogue.ou(lam, xi, sig)
    r = sde.euler(f, g, t, x0, npaths=n, seed=31, integrals=True)
    I, XT = r["integrals"], r["XT"][0]
    assert abs(I["QV"][0].mean() - sig ** 2) < 5 * I["QV"][0].std() / np.sqrt(n) + 2e-3      # E QV = sigma^2 T (+ O(dt))
    assert np.all(I["QV"][0] > 0)
    # Stratonovich - Ito = 0.5 [X,B]_T = 0.5 sigma T in expectation (additive noise)
    diff = I["strat"][0] - I["ito"][0]
    assert abs(diff.mean() - 0.5 * sig) < 5 * diff.std() / np.sqrt(n) + 1e-3
    # E int X dt for OU started at x0
    want = xi + (x0 - xi) * (1 - np.exp(-lam)) / lam
    assert abs(I["time"][0].mean() - want) < 5 * I["time"][0].std() / np.sqrt(n) + 2e-3
    assert np.isfinite(XT).all()


def test_path_integrals_argument_errors(sde):
    from sdetools_b200._lib import SdeGpuError
    f, g = sde.catalogue.ou()
    with pytest.raises(SdeGpuError, match="path_integrals"):
        sde.euler(f, g, [0.0, 0.5, 1.0], 0.0, npaths=4, seed=1, integrals=True, r=sde.catalogue.kill_const(1.0))


def test_single_channel_linear_model_shares_the_scalar_stream(sde):
    """A linear model with one noise channel draws from the same stream as the catalogue's scalar models:
    dX = -lam X dt + sig dB as `linear` and as `ou` give the same paths under one seed (to FMA-level rounding)."""
    t = np.arange(201) * 0.01
    a = sde.euler(*sde.catalogue.linear([[-1.3]], [[0.9]]), t, 2.0, npaths=500, seed=8)["X"]
    b = sde.euler(*sde.catalogue.ou(1.3, 0.0, 0.9), t, 2.0, npaths=500, seed=8)["X"]
    assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_device_buffers_at_8_byte_offsets_take_the_elementwise_paths(eng):
    """The sector (256-bit) loads/stores need 32-byte aligned buffers; device tensors that start 8 bytes into an
    allocation must fall back to element-wise access and give the same bits (k_traj, k_stream, k_integral, k_rvbm)."""
    import torch
    dev = torch.device("cuda", 0)
    eng.use_torch_stream()                                      # device-pointer calls are only enqueued: share torch's stream

    def off8(n):                                                # n doubles starting 8 bytes past an aligned address
        return torch.empty(n + 1, dtype=torch.float64, device=dev)[1:]

    P, nt = 333, 205
    t = np.arange(nt) * 0.01
    # fused generator trajectory (k_traj), two components
    Xa = torch.empty(P * 2 * nt, dtype=torch.float64, device=dev)
    Xo = off8(P * 2 * nt)
    assert Xo.data_ptr() % 32 == 8 and Xa.data_ptr() % 32 == 0
    for X in (Xa, Xo):
        eng.run("vdp", "heun", [1.0, 0.5], t, [0.1, -0.2], nx=2, npaths=P, seed=6, X=X)
    assert torch.equal(Xa, Xo)
    # supplied B (k_stream): misaligned input and output
    Ba = torch.empty(P * nt, dtype=torch.float64, device=dev)
    eng.rvbm(t, P, Ba, seed=4)
    Bo = off8(P * nt)
    Bo.copy_(Ba)
    X1a = torch.empty(P * nt, dtype=torch.float64, device=dev)
    X1o = off8(P * nt)
    eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=P, B=Ba, projection="abs", X=X1a)
    eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=P, B=Bo, projection="abs", X=X1o)
    assert torch.equal(X1a, X1o)
    # integrals and rvBM
    la, lo = torch.empty(P * nt, dtype=torch.float64, device=dev), off8(P * nt)
    eng.itointegral(X1a, Ba, nt, P, la)
    eng.itointegral(X1o, Bo, nt, P, lo)
    assert torch.equal(la, lo)
    eng.rvbm(t, P, Bo, seed=4)
    assert torch.equal(Ba, Bo)
    eng.use_own_stream()


def test_output_options_agree_across_kernels(eng):
    """One ensemble, every way of asking for its terminal statistics: trajectory kernel, register-resident kernel and
    fused-integral kernel must report the same terminal states, histogram (on the second component) and power sums;
    per-path x0 with the generator; the ACCUMULATE flag adds a second shard to the first."""
    P, nt, seed = 3001, 130, 19
    t = np.arange(nt) * 0.02
    rng = np.random.default_rng(3)
    x0 = np.ascontiguousarray(rng.uniform(-1, 1, (P, 2)))                         # x0[p*nx+j]
    kw = dict(nx=2, npaths=P, seed=seed, center=[0.1, -0.2], hist=(-4.0, 4.0, 37, 1))
    out = {}
    for name, extra in (("traj", dict(X=np.empty((nt, 2, P), order="F"))), ("terminal", {}),
                        ("functionals", dict(path_integrals=np.empty((4, 2, P), order="F")))):
        XT, ps, hc = np.empty((2, P), order="F"), np.zeros((5, 2), order="F"), np.zeros(40, np.uint64)
        eng.run("vdp", "heun", [1.0, 0.5], t, x0, XT=XT, power_sums=ps, hist_counts=hc, **kw, **extra)
        out[name] = (XT, ps, hc, extra)
    XT0, ps0, hc0, ex0 = out["traj"]
    assert np.array_equal(ex0["X"][0].T, x0) and np.array_equal(ex0["X"][-1], XT0)
    assert int(hc0.sum()) == P
    for name in ("terminal", "functionals"):
        XT, ps, hc, _ = out[name]
        assert np.allclose(XT, XT0, rtol=1e-12, atol=1e-14), name
        assert np.array_equal(hc, hc0), name
        assert np.allclose(ps, ps0, rtol=1e-10, atol=1e-9), name
    d = XT0 - np.array([[0.1], [-0.2]])
    assert np.array_equal(ps0[0], [P, P]) and np.allclose(ps0[1], d.sum(axis=1), rtol=1e-10, atol=1e-9)
    assert np.allclose(ps0[4], (d ** 4).sum(axis=1), rtol=1e-10)
    want_h = O.hist_counts(XT0[1], -4.0, 4.0, 37)
    assert np.array_equal(hc0, want_h)
    # ACCUMULATE: shard A then shard B on top == the whole ensemble
    ps, hc = np.zeros((5, 2), order="F"), np.zeros(40, np.uint64)
    kw2 = dict(nx=2, seed=seed, center=[0.1, -0.2], hist=(-4.0, 4.0, 37, 1), power_sums=ps, hist_counts=hc)
    eng.run("vdp", "heun", [1.0, 0.5], t, np.ascontiguousarray(x0[:1000]), npaths=1000, **kw2)
    eng.run("vdp", "heun", [1.0, 0.5], t, np.ascontiguousarray(x0[1000:]), npaths=P - 1000, path_offset=1000, accumulate=True, **kw2)
    assert np.array_equal(hc, hc0) and np.allclose(ps, ps0, rtol=1e-10, atol=1e-9)


def test_c_abi_from_plain_c(gpu, tmp_path):
    """tests/native/abi_smoke.c: the library driven from C with host pointers, as the R shim would."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "abi_smoke"
    lib = os.path.join(root, "sdetools_b200")
    subprocess.run(["/usr/bin/gcc", "-O1", "-std=c11", "-I", os.path.join(root, "include"), "-o", str(exe),
                    os.path.join(root, "tests", "native", "abi_smoke.c"), "-L", lib, "-lsdegpu", "-lm", f"-Wl,-rpath,{lib}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "abi smoke ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("force_dmma", [False, True])
@pytest.mark.parametrize("d,nB", [(64, 64), (100, 7), (128, 128), (160, 160), (256, 5), (40, 1)])
def test_linear_dmma_dense_noise_matrix(eng, d, nB, force_dmma, monkeypatch):
    """(force_dmma as above: d <= 96 runs on k_linear_tc by default, k_linear_dmma_dense when forced.)
    Correlated noise (dense G, any nB) on the tensor path: the increment term is a second GEMM on a tile of
    normals.  Must agree with the generic kernel (taken when the trajectory is requested) to summation order and
    with the CPU oracle's fused generator to the quantile table's accuracy; moments = those of the terminal states."""
    if force_dmma:
        monkeypatch.setenv("SDEGPU_LINEAR_DMMA", "1")
    rng = np.random.default_rng(1000 * d + nB)
    R = rng.standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    G = np.tril(rng.standard_normal((d, d)))[:, :nB] / np.sqrt(d) if nB > 1 else rng.standard_normal((d, 1))
    params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
    P, nt, seed, off = 150, 31, 78, 4321
    t = O.r_seq(0, (nt - 1) * 0.01, 0.01)
    x0 = rng.standard_normal(d)
    XT = np.empty((d, P), order="F")
    ps = np.zeros((5, d), order="F")
    eng.run("linear", "euler", params, t, x0, nx=d, nB=nB, npaths=P, seed=seed, path_offset=off, XT=XT, power_sums=ps,
            center=np.zeros(d))
    X = np.empty((nt, d, P), order="F")
    eng.run("linear", "euler", params, t, x0, nx=d, nB=nB, npaths=P, seed=seed, path_offset=off, X=X)
    assert np.allclose(XT, X[-1], rtol=1e-11, atol=1e-12)
    _, XTc = c_oracle.ensemble(0, 6, params, d, nB, 0, t, x0, seed=seed, path_offset=off, npaths=P, full=False, nthreads=8)
    assert np.allclose(XT, XTc.T, rtol=1e-7, atol=1e-8)
    assert np.all(ps[0] == P) and np.allclose(ps[1], XT.sum(axis=1), rtol=1e-10, atol=1e-10)


def test_exact_stepping_at_d128_runs_on_the_tensor_path(sde):
    """laws.linear_exact at d = 128 gives a dense lower-triangular g: X_{k+1} = e^{A dt} X_k + chol(S_dt) Z in 8 coarse
    steps reproduces dLinSDE's mean and covariance at T (MC error only), where Euler on that grid could not."""
    d, n, T, P = 128, 8, 2.0, 1 << 16
    rng = np.random.default_rng(5)
    R = rng.standard_normal((d, d))
    A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
    G = rng.standard_normal((d, d)) / np.sqrt(d)
    x0 = np.zeros(d)
    x0[:3] = [1.0, -1.0, 0.5]
    fe, ge = sde.linear_exact(A, G, T / n)
    XT = sde.euler(fe, ge, np.arange(n + 1) * (T / n), x0, npaths=P, seed=9, output="terminal")["XT"]
    law = sde.dLinSDE(A, G, T, x0=x0)
    se = np.sqrt(np.diag(law["St"]) / P)
    assert np.max(np.abs(XT.mean(axis=1) - law["EX"]) / se) < 5.0
    C = np.cov(XT)
    assert np.max(np.abs(C - law["St"])) < 5.5 * np.max(np.diag(law["St"])) * np.sqrt(2.0 / P)


@pytest.mark.parametrize("nbins", [1, 8192])
def test_histogram_bin_count_extremes(eng, nbins):
    """1 bin and the maximum of 8192 bins (32 KB of shared memory next to the 151 KB quantile table), through the
    register-resident, trajectory and supplied-B kernels, against the oracle's histogram of the terminal states."""
    from sdetools_b200._lib import SdeGpuError
    P, nt = 5000, 41
    t = np.arange(nt) * 0.025
    lo, hi = 0.0, 3.0
    for want_x in (False, True):
        XT = np.empty((1, P), order="F")
        hc = np.zeros(nbins + 3, np.uint64)
        X = np.empty((nt, 1, P), order="F") if want_x else None
        eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=P, projection="abs", seed=2, XT=XT, X=X,
                hist_counts=hc, hist=(lo, hi, nbins))
        assert np.array_equal(hc, O.hist_counts(XT[0], lo, hi, nbins)) and int(hc.sum()) == P
    tb, B = brownian(nt, 1, 300, 4, dt=0.025)
    XT = np.empty((1, 300), order="F")
    hc = np.zeros(nbins + 3, np.uint64)
    eng.run("cir", "heun", [1.0, 1.0, 1.0], tb, [1.0], nx=1, npaths=300, B=np.asfortranarray(B), projection="abs", XT=XT,
            hist_counts=hc, hist=(lo, hi, nbins))
    assert np.array_equal(hc, O.hist_counts(XT[0], lo, hi, nbins))
    with pytest.raises(SdeGpuError):
        eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=10, seed=2, hist_counts=np.zeros(8196, np.uint64), hist=(lo, hi, 8193))
    with pytest.raises(SdeGpuError):
        eng.run("cir", "heun", [1.0, 1.0, 1.0], t, [1.0], nx=1, npaths=10, seed=2, hist_counts=np.zeros(7, np.uint64), hist=(1.0, 1.0, 4))


def test_killing_and_stopping_on_the_second_component(sde):
    """nx = 2 (van der Pol): rate and stopping rule read the velocity component; path by path against the numpy
    restatement of R's loop, and the lane-refill kernel against the trajectory kernel."""
    t, B = brownian(401, 1, 64, 31)
    f, g = sde.catalogue.vdp(1.0, 0.5)
    r = sde.catalogue.kill_quad(0.8, 0.0, component=1)
    h = sde.catalogue.stop_outside(-1.5, 1.2, component=1)
    Stau = np.random.default_rng(6).random(64)
    got = sde.heun(f, g, t, [0.1, -0.2], B, r=r, h=h, Stau=Stau)
    fo, go, _, _ = O.catalogue("vdp", [1.0, 0.5])
    ends = set()
    for p in range(64):
        ref = O.heun(fo, go, t, [0.1, -0.2], B=B[:, :, p], r=lambda x: 0.8 * (x[1] * x[1]),
                     h=lambda tt, x: (x[1] + 1.5) * (1.2 - x[1]), Stau=Stau[p])
        k = ref["times"].size
        assert got["tau"][p] == ref["tau"]
        assert np.array_equal(got["X"][:k, :, p], ref["X"]) and np.all(np.isnan(got["X"][k:, :, p]))
        ends.add(k == t.size)
    assert ends == {True, False}
    term = sde.heun(f, g, t, [0.1, -0.2], B, r=r, h=h, Stau=Stau, output="terminal")
    assert np.array_equal(term["tau"], got["tau"]) and np.array_equal(term["XT"], got["XT"]) and np.array_equal(term["S"], got["S"])


def test_fuzz_shapes_supplied_B_bit_identical(eng):
    """Sixty random shapes — nt from 2, a single path, ragged counts, non-uniform grids, shared or per-path x0, every
    model, scheme and projection — through the supplied-B STRICT path (sector loads/stores with per-row phases)
    and the integral kernels, against the C oracle bit for bit."""
    rng = np.random.default_rng(20261020)
    models = [c for c in CASES]
    for it in range(60):
        model, params, x0, proj = models[int(rng.integers(len(models)))]
        scheme = ("euler", "heun")[int(rng.integers(2))]
        nt = int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 12, 17, 33, 70]))
        P = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 90]))
        nx = len(x0)
        t = np.concatenate(([0.0], np.cumsum(rng.uniform(0.002, 0.03, nt - 1))))
        B = np.cumsum(np.concatenate([np.zeros((1, 1, P)), rng.standard_normal((nt - 1, 1, P)) * np.sqrt(np.diff(t))[:, None, None]]), axis=0)
        per_path = bool(rng.integers(2))
        x0a = np.abs(np.array(x0)[None, :] + 0.1 * rng.standard_normal((P, nx))) if per_path else np.array(x0, float)
        Xo, XTo = c_oracle.ensemble(SCH[scheme], O.MODEL_IDS[model], params, nx, 1, O.PROJ_IDS[proj], t,
                                    np.ascontiguousarray(x0a), B=np.ascontiguousarray(B.transpose(2, 1, 0)))
        X = np.empty((nt, nx, P), order="F")
        XT = np.empty((nx, P), order="F")
        eng.run(model, scheme, params, t, np.ascontiguousarray(x0a), nx=nx, npaths=P, B=np.asfortranarray(B), projection=proj,
                arith="strict", X=X, XT=XT)
        assert np.array_equal(X, Xo.transpose(2, 1, 0)), (it, model, scheme, nt, P, per_path)
        assert np.array_equal(XT, XTo.T), (it, model, scheme, nt, P)
        # integrals of the first component against the long-double cumsum
        Xc = np.ascontiguousarray(X[:, 0, :].T)
        Bc = np.ascontiguousarray(B[:, 0, :].T)
        out = np.empty((nt, P), order="F")
        eng.itointegral(np.asfortranarray(X[:, 0, :]), np.asfortranarray(B[:, 0, :]), nt, P, out)
        assert np.array_equal(out, c_oracle.integral(0, Xc, Bc).T), (it, "ito", nt, P)
        eng.qv(np.asfortranarray(X[:, 0, :]), nt, P, out)
        assert np.array_equal(out, c_oracle.integral(3, Xc).T), (it, "qv", nt, P)


def test_fuzz_shapes_fused_generator(eng):
    """Random small shapes through the generator kernels: the trajectory kernel, the register-resident kernel and the
    oracle's loop fed with the device's normals must agree (1e-12), for nt from 2 and path counts around the warp size."""
    rng = np.random.default_rng(77)
    for it in range(80):
        model, params, x0, proj = CASES[int(rng.integers(len(CASES)))]
        scheme = ("euler", "heun")[int(rng.integers(2))]
        # nt around the trajectory kernel's 16-point chunks, its two-chunk stores and its four-slot ring
        nt = int(rng.choice([2, 3, 4, 5, 6, 8, 9, 13, 17, 18, 19, 20, 21, 33, 36, 40, 53, 70, 85, 99, 100, 101, 102, 131]))
        P = int(rng.choice([1, 5, 31, 33, 64, 65, 100, 129, 257]))
        nx, seed, off = len(x0), int(rng.integers(1 << 30)), int(rng.integers(1 << 40))
        t = np.concatenate(([0.0], np.cumsum(rng.uniform(0.002, 0.03, nt - 1))))
        X = np.empty((nt, nx, P), order="F")
        XT = np.empty((nx, P), order="F")
        eng.run(model, scheme, params, t, x0, nx=nx, npaths=P, projection=proj, seed=seed, path_offset=off, X=X)
        eng.run(model, scheme, params, t, x0, nx=nx, npaths=P, projection=proj, seed=seed, path_offset=off, XT=XT)
        z = eng.normals(seed, off, P, nt - 1)
        dB = (np.sqrt(O.r_diff(t))[:, None] * z)[:, None, :]
        want = O.ensemble(scheme, model, params, t, x0, dB=dB, proj=proj)
        # a path that a projection held at (or rounding put within 1e-6 of) zero has no 1e-12 bar under sqrt noise: the
        # next increment g(x) dB = gamma sqrt(x) dB turns a last-bit difference at x ~ 0 into sqrt(1e-16) ~ 1e-8
        calm = np.abs(want).min(axis=(0, 1)) > 1e-6 if model == "cir" else np.ones(P, bool)
        # CIR takes the folded step with the second-order square root (models.cuh FastStep, 2e-14 per step): FAST bar 1e-10
        rtol = 1e-10 if model == "cir" else 1e-12
        assert np.allclose(X[..., calm], want[..., calm], rtol=rtol, atol=1e-13), (it, model, scheme, nt, P)
        assert np.allclose(X, want, rtol=1e-6, atol=1e-6), (it, model, scheme, nt, P)
        assert np.array_equal(XT, X[-1]), (it, model, scheme, nt, P)   # both kernels take the folded FAST step


def test_fuzz_shapes_killed_functionals_linear(sde, eng):
    """Small and ragged shapes (nt from 2) through the remaining kernels: lane-refill vs one-path-per-thread killing,
    fused path integrals vs the integral functions on the stored path, linear kernels with and without trajectory."""
    rng = np.random.default_rng(99)
    f, g = sde.catalogue.ou(1.0, 0.0, 1.0)
    for it in range(25):
        nt = int(rng.choice([2, 3, 4, 5, 6, 9, 30]))
        P = int(rng.choice([1, 7, 33, 200]))
        t = np.concatenate(([0.0], np.cumsum(rng.uniform(0.01, 0.2, nt - 1))))
        seed = int(rng.integers(1 << 30))
        kw = dict(r=sde.catalogue.kill_exp(0.8, 1.0), h=sde.catalogue.stop_outside(-0.7, 0.9), seed=seed, npaths=P)
        a = sde.euler(f, g, t, 0.1, output="path", **kw)
        b = sde.euler(f, g, t, 0.1, output="terminal", **kw)
        assert np.array_equal(np.asarray(a["tau"]), np.asarray(b["tau"])) and np.array_equal(np.asarray(a["S"]), np.asarray(b["S"])), (it, nt, P)
        assert np.array_equal(np.asarray(a["XT"]), np.asarray(b["XT"])), (it, nt, P)
        # fused integrals, supplied B
        B = np.cumsum(np.concatenate([np.zeros((1, 1, P)), rng.standard_normal((nt - 1, 1, P)) * np.sqrt(np.diff(t))[:, None, None]]), axis=0)
        fv, gv = sde.catalogue.vdp(1.0, 0.5)
        path = sde.heun(fv, gv, t, [0.1, -0.2], B)["X"].reshape(nt, 2, P)
        I = sde.heun(fv, gv, t, [0.1, -0.2], B, integrals=True)["integrals"]
        for j in range(2):
            Xj, Bj = path[:, j, :], B[:, 0, :]
            assert np.array_equal(np.asarray(I["QV"]).reshape(2, P)[j], np.atleast_2d(sde.QuadraticVariation(Xj).reshape(nt, P))[-1]), (it, nt, P)
            assert np.array_equal(np.asarray(I["strat"]).reshape(2, P)[j], np.atleast_2d(sde.stochint(Xj, Bj, "c").reshape(nt, P))[-1]), (it, nt, P)
        # linear kernels: small, warp and generic shapes
        d, nB = [(2, 2), (3, 1), (7, 4), (20, 20), (40, 3)][int(rng.integers(5))]
        A = -0.5 * np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
        G = rng.standard_normal((d, nB))
        x0 = rng.standard_normal(d)
        fl, gl = sde.catalogue.linear(A, G)
        for fn in (sde.euler, sde.heun):
            X = fn(fl, gl, t, x0, npaths=P, seed=seed)["X"].reshape(nt, d, P)
            XT = np.asarray(fn(fl, gl, t, x0, npaths=P, seed=seed, output="terminal")["XT"]).reshape(d, P)
            assert np.array_equal(X[0], np.repeat(x0[:, None], P, axis=1)), (it, d, nB, nt, P)
            assert np.allclose(X[-1], XT, rtol=1e-10, atol=1e-12), (it, d, nB, nt, P)


def test_fuzz_shapes_tensor_path(eng):
    """The DMMA kernels on awkward shapes: d just above 32 and just below 256, one step, path counts around the
    64-path tile, diagonal and dense G — terminal output against the generic kernel's trajectory."""
    rng = np.random.default_rng(123)
    for it in range(14):
        d = int(rng.choice([33, 63, 65, 127, 129, 255]))
        dense = bool(rng.integers(2))
        nB = d if not dense else int(rng.choice([1, 2, 5, min(d, 60)]))
        nt = int(rng.choice([2, 3, 6]))
        P = int(rng.choice([1, 63, 64, 65, 130]))
        R = rng.standard_normal((d, d))
        A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
        G = np.diag(rng.uniform(0.5, 1.5, d)) if not dense else rng.standard_normal((d, nB)) / np.sqrt(d)
        params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
        t = np.concatenate(([0.0], np.cumsum(rng.uniform(0.005, 0.02, nt - 1))))
        x0 = rng.standard_normal(d)
        seed = int(rng.integers(1 << 30))
        XT = np.empty((d, P), order="F")
        X = np.empty((nt, d, P), order="F")
        eng.run("linear", "euler", params, t, x0, nx=d, nB=nB, npaths=P, seed=seed, XT=XT)
        eng.run("linear", "euler", params, t, x0, nx=d, nB=nB, npaths=P, seed=seed, X=X)
        assert np.allclose(XT, X[-1], rtol=1e-11, atol=1e-12), (it, d, nB, nt, P, dense)


def test_fuzz_shapes_integral_functions(sde):
    """stochint (every subset of rules), itointegral, QuadraticVariation, CrossVariation on vectors and matrices with
    n from 1 and column counts around the warp size, against the long-double oracle bit for bit."""
    rng = np.random.default_rng(314)
    for it in range(40):
        n = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 64, 129]))
        nc = int(rng.choice([1, 2, 31, 33, 70]))
        f = rng.standard_normal((n, nc)) * 10.0 ** rng.integers(-3, 4, (n, nc))
        g = np.cumsum(rng.standard_normal((n, nc)), axis=0)
        fa, ga = np.ascontiguousarray(f.T), np.ascontiguousarray(g.T)
        rules = [r for r in "lrc" if rng.integers(2)] or ["c"]
        got = sde.stochint(f, g, rules)
        for r in rules:
            want = c_oracle.integral({"l": 0, "r": 1, "c": 2}[r], fa, ga).T
            assert np.array_equal(np.asarray(got[r]).reshape(n, nc), want), (it, n, nc, r)
        assert np.array_equal(sde.itointegral(f, g).reshape(n, nc), c_oracle.integral(0, fa, ga).T), (it, n, nc)
        assert np.array_equal(sde.QuadraticVariation(g).reshape(n, nc), c_oracle.integral(3, ga).T), (it, n, nc)
        assert np.array_equal(sde.CrossVariation(f, g).reshape(n, nc), c_oracle.integral(4, fa, ga).T), (it, n, nc)
        one = sde.stochint(f[:, 0], g[:, 0], "r")
        assert one.shape == (n,) and np.array_equal(one, O.stochint(f[:, 0], g[:, 0], "r"))


def test_vignette_models_ito_integrals_and_itos_formula(sde):
    """The two vignette examples with their own models.  ItoIntegrals.Rmd:66-93: for dX = (x - x^3) dt + sqrt(1 + x^2) dB
    the Euler path equals int f dt + int g dB to machine precision.  ItosFormula.Rmd:36-116: Y = log X of stochastic
    logistic growth solves dY = (1 - e^Y - sigma^2/2) dt + sigma dB; both Euler paths driven by the same B stay close."""
    t, B = brownian(1001, 1, 32, 17, dt=0.01)
    f, g = sde.catalogue.dwsqrt()
    X = sde.euler(f, g, t, 0.0, B)["X"][:, 0, :]
    Bm = B[:, 0, :]
    tt = np.repeat(t[:, None], 32, axis=1)
    resid = X - sde.itointegral(X - X * X * X, tt) - sde.itointegral(np.sqrt(1.0 + X * X), Bm)
    assert np.max(np.abs(resid)) < 5e-14 * max(1.0, np.max(np.abs(X)))
    fo, go, _, _ = O.catalogue("dwsqrt", [])
    assert np.array_equal(O.euler(fo, go, t, 0.0, B=B[:, :, 3])["X"][:, 0], X[:, 3])            # bit-identical to the R loop restated
    sigma, x0 = 0.25, 0.01
    fl, gl = sde.catalogue.logistic(sigma, 0.0)
    fy, gy = sde.catalogue.loglogistic(sigma)
    Xl = sde.euler(fl, gl, t, x0, B)["X"][:, 0, :]
    Y = sde.euler(fy, gy, t, np.log(x0), B)["X"][:, 0, :]
    assert np.max(np.abs(Y - np.log(Xl))) < 0.05                                                    # O(dt) apart, as in the vignette's plot
    foy, goy, _, _ = O.catalogue("loglogistic", [sigma])
    want = O.euler(foy, goy, t, np.log(x0), B=B[:, :, 5])["X"][:, 0]
    assert np.allclose(Y[:, 5], want, rtol=1e-12, atol=1e-13)                                       # exp(): last-ulp differences only
    # fused generator and heun run as well; terminal-only kernel agrees with the trajectory kernel
    for fq, gq, x0q in ((f, g, 0.0), (fy, gy, np.log(x0))):
        a = sde.heun(fq, gq, t[:101], x0q, npaths=100, seed=3)["X"][-1, 0]
        b = sde.heun(fq, gq, t[:101], x0q, npaths=100, seed=3, output="terminal")["XT"][0]
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13)
