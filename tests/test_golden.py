"""Committed fixtures (tests/golden/, written by tests/golden/make_golden.py).

CPU half: the oracle still reproduces the frozen vectors bit for bit, and the reference's own
known answers (tests/testthat/test_solvers.R:3-36, R/sdetools.R:587-590) hold for the restatement.
GPU half: the CUDA path, through the C ABI, reproduces the same frozen vectors without the oracle
being involved at run time (bit-exact for STRICT arithmetic and the integrals, 1e-9 for the
generator whose table inversion is specified to 7.4e-10)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("ou", [1.3, 0.7, 0.9], [2.0], "identity"), ("cir", [1.0, 1.0, 1.0], [1.0], "abs"),
         ("cir", [0.5, 0.2, 1.5], [0.05], "relu"), ("vdp", [1.0, 0.5], [0.1, -0.2], "identity"),
         ("logistic", [0.5, 1.0], [0.3], "identity"), ("gbm", [0.4, 1.0], [1.0], "abs"),
         ("doublewell", [0.5], [0.2], "identity")]


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(HERE, "oracle_vectors.npz"))


@pytest.fixture(scope="module")
def known():
    with open(os.path.join(HERE, "reference_known_answers.json")) as fh:
        return json.load(fh)


# ---- CPU: the oracle against the fixtures ------------------------------------------------------
def test_generator_script_reproduces_fixture(vec):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.oracle_vectors()
    assert sorted(fresh) == sorted(vec.files)
    for k in vec.files:
        assert np.array_equal(fresh[k], vec[k]), k
    with open(os.path.join(HERE, "reference_known_answers.json")) as fh:
        assert json.load(fh) == json.loads(json.dumps(mod.known_answers()))


def test_c_oracle_reproduces_fixture(vec):
    from oracle import c_oracle
    from oracle import sde_oracle as O
    t, B = vec["times"], vec["B"]
    for k, (model, params, x0, proj) in enumerate(CASES):
        for s, scheme in enumerate(("euler", "heun")):
            X, _ = c_oracle.ensemble(s, O.MODEL_IDS[model], params, len(x0), 1, O.PROJ_IDS[proj], t, np.array(x0),
                                     B=np.ascontiguousarray(B.transpose(2, 1, 0)))
            assert np.array_equal(X.transpose(2, 1, 0), vec[f"X_{scheme}_{k}"]), (model, scheme)
    params = np.concatenate([vec["lin_A"].ravel(order="F"), vec["lin_G"].ravel(order="F")])
    for s, scheme in enumerate(("euler", "heun")):
        X, _ = c_oracle.ensemble(s, 6, params, 3, 2, 0, t, vec["lin_x0"],
                                 B=np.ascontiguousarray(vec["lin_B"].transpose(2, 1, 0)))
        assert np.array_equal(X.transpose(2, 1, 0), vec[f"lin_X_{scheme}"])
    f, g = vec["int_f"], vec["int_g"]
    assert np.array_equal(c_oracle.integral(0, f, g), vec["int_ito"])
    assert np.array_equal(c_oracle.integral(1, f, g), vec["int_stochint_r"])
    assert np.array_equal(c_oracle.integral(2, f, g), vec["int_stochint_c"])
    assert np.array_equal(c_oracle.integral(3, g), vec["int_qv"])
    assert np.array_equal(c_oracle.integral(4, f, g), vec["int_cv"])
    seed = int(vec["gen_seed"][0])
    for j, p in enumerate(vec["gen_paths"]):
        assert np.allclose(c_oracle.normals(seed, int(p), 0, 13), vec["gen_z"][:, j], rtol=0, atol=1e-13)


def test_reference_known_answers(known):
    from oracle import laws, philox_ref
    from oracle import sde_oracle as O
    tol = 1.5e-8                                            # testthat::expect_equal
    for c in known["dLinSDE"]["cases"]:
        kw = {k: np.array(c[k]) for k in ("x0", "u", "S0") if k in c}
        sol = laws.dLinSDE(np.array(c["A"]), np.array(c["G"]), c["t"], **kw)
        for key in ("eAt", "St", "EX"):
            if key in c:
                assert np.allclose(np.asarray(sol[key]).ravel(), np.array(c[key]).ravel(), rtol=tol, atol=10 * tol), (c, key)
    for c in known["lyap"]["cases"]:
        assert np.allclose(laws.lyap(np.array(c["A"]), np.array(c["Q"])), np.array(c["P"]))
    for c in known["philox4x32_10"]["cases"]:
        got = philox_ref.philox4x32_10(*[np.uint64(v) for v in c["ctr"]], c["key"][0], c["key"][1])
        assert [int(v) for v in got] == c["out"]
    d = known["dims"]
    t = np.array(d["times"], float)
    rng = np.random.default_rng(0)
    f = lambda t_, x: -x
    assert list(O.euler(f, lambda t_, x: x, t, np.ones(d["nx"]), rng=rng)["X"].shape) == d["dim_X"]
    g2 = lambda t_, x: rng.uniform(size=(d["nx"], 2))
    assert list(O.euler(f, g2, t, np.ones(d["nx"]), rng=rng)["X"].shape) == d["dim_X"]


# ---- GPU: the CUDA path against the fixtures, oracle not involved ------------------------------
@pytest.fixture(scope="module")
def eng(gpu):
    import sdetools_b200
    return sdetools_b200.default_engine(0)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
@pytest.mark.parametrize("k", range(len(CASES)))
def test_gpu_reproduces_golden_paths(eng, vec, scheme, k):
    model, params, x0, proj = CASES[k]
    t, B = vec["times"], vec["B"]
    want = vec[f"X_{scheme}_{k}"]
    X = np.empty(want.shape, order="F")
    eng.run(model, scheme, params, t, x0, nx=len(x0), npaths=B.shape[2], B=np.asfortranarray(B), projection=proj,
            arith="strict", X=X)
    assert np.array_equal(X, want)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_gpu_reproduces_golden_linear(eng, vec, scheme):
    params = np.concatenate([vec["lin_A"].ravel(order="F"), vec["lin_G"].ravel(order="F")])
    want = vec[f"lin_X_{scheme}"]
    X = np.empty(want.shape, order="F")
    eng.run("linear", scheme, params, vec["times"], vec["lin_x0"], nx=3, nB=2, npaths=want.shape[2],
            B=np.asfortranarray(vec["lin_B"]), arith="strict", X=X)
    assert np.array_equal(X, want)


@pytest.mark.gpu
def test_gpu_reproduces_golden_integrals(gpu, vec):
    import sdetools_b200 as sde
    f, g = vec["int_f"], vec["int_g"]
    for c in range(f.shape[0]):
        got = sde.stochint(f[c], g[c], rule=["l", "r", "c"])
        for rule in "lrc":
            assert np.array_equal(np.asarray(got[rule]), vec[f"int_stochint_{rule}"][c]), (c, rule)
        assert np.array_equal(sde.itointegral(f[c], g[c]), vec["int_ito"][c])
        assert np.array_equal(sde.QuadraticVariation(g[c]), vec["int_qv"][c])
        assert np.array_equal(sde.CrossVariation(f[c], g[c]), vec["int_cv"][c])


@pytest.mark.gpu
def test_gpu_reproduces_golden_generator(eng, vec, known):
    seed = int(vec["gen_seed"][0])
    for j, p in enumerate(vec["gen_paths"]):
        z = eng.normals(seed, int(p), 1, 13)[:, 0]
        assert np.max(np.abs(z - vec["gen_z"][:, j])) < 1e-9
    for c in known["philox4x32_10"]["cases"]:
        assert list(eng.philox4x32_10(tuple(c["ctr"]), tuple(c["key"]))) == c["out"]
