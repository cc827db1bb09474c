"""The R .Call shim (r/src/sdegpu_shim.c) EXECUTED, not only type-checked.  R is not installable in this image, so the
shim is compiled against tests/mock_r/Rinternals.h and linked with tests/mock_r/mock_runtime.c, a functional stand-in
for the slice of R's runtime the shim touches (typed vectors with names/dim, PROTECT counting, Rf_error and interrupts
as longjmps, unif_rand between Get/PutRNGstate, R_ToplevelExec).  ctypes plays the part of R's .Call.

CPU tests: every argument check fires as an R error before a pointer reaches the library, the PROTECT stack is balanced
on every path, seeds round-trip through doubles.  GPU tests: each entry point returns what the C ABI returns for the
same inputs (which tests/test_gpu_parity.py pins against the oracle), on one and on several devices; a pending
interrupt stops a long run between chunks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "native", "_build")
REALSXP, INTSXP, VECSXP, STRSXP, NILSXP = 14, 13, 19, 16, 0


class MockR:
    """.Call on the mock runtime: numpy in, numpy (or dicts of numpy) out."""

    def __init__(self):
        from sdetools_b200 import build
        lib_dir = os.path.dirname(build.build())
        os.makedirs(OUT, exist_ok=True)
        so = os.path.join(OUT, "libshim_mock.so")
        srcs = [os.path.join(ROOT, "r", "src", "sdegpu_shim.c"), os.path.join(ROOT, "tests", "mock_r", "mock_runtime.c")]
        deps = srcs + [os.path.join(ROOT, "tests", "mock_r", "Rinternals.h"), os.path.join(ROOT, "include", "sdegpu.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
            r = subprocess.run(["/usr/bin/gcc", "-std=gnu99", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_r"),
                                "-I", os.path.join(ROOT, "include"), *srcs, "-o", so, "-L", lib_dir, "-lsdegpu", f"-Wl,-rpath,{lib_dir}", "-lm"],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        L = self.L = C.CDLL(so)
        vp = C.c_void_p
        for name, res, args in [("mock_call", C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]), ("mock_last_error", C.c_char_p, []),
                                ("mock_reset", None, []), ("mock_set_seed", None, [C.c_uint64]), ("mock_interrupt_after", None, [C.c_int]),
                                ("mock_interrupt_checks", C.c_int, []), ("mock_nil", vp, []), ("mock_real", vp, [vp, C.c_ssize_t]),
                                ("mock_int", vp, [vp, C.c_ssize_t]), ("mock_list", vp, [C.c_ssize_t]),
                                ("mock_list_set", None, [vp, C.c_ssize_t, C.c_char_p, vp]), ("mock_set_dim", None, [vp, C.c_int, vp]),
                                ("mock_type", C.c_int, [vp]), ("mock_length", C.c_ssize_t, [vp]), ("mock_data", vp, [vp]),
                                ("mock_ndim", C.c_int, [vp]), ("mock_dim", C.c_int, [vp, C.c_int]), ("mock_elt", vp, [vp, C.c_ssize_t]),
                                ("mock_name", C.c_char_p, [vp, C.c_ssize_t])]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    # -- R values ------------------------------------------------------------------------------------------------
    def sexp(self, v):
        L = self.L
        if v is None:
            return L.mock_nil()
        if isinstance(v, dict):
            lst = L.mock_list(len(v))
            for i, (k, x) in enumerate(v.items()):
                L.mock_list_set(lst, i, k.encode(), self.sexp(x))
            return lst
        a = np.asarray(v)
        if a.dtype.kind in "iu" and not isinstance(v, np.ndarray):
            a = a.astype(np.int32)
        if a.dtype == np.int32:
            buf = np.ascontiguousarray(a.ravel(order="F"))
            s = L.mock_int(buf.ctypes.data, buf.size)
        else:
            buf = np.ascontiguousarray(a.astype(np.float64).ravel(order="F"))
            s = L.mock_real(buf.ctypes.data, buf.size)
        if a.ndim >= 2:
            d = np.array(a.shape, dtype=np.int32)
            L.mock_set_dim(s, a.ndim, d.ctypes.data)
        return s

    def value(self, s):
        L = self.L
        t = L.mock_type(s)
        if t == NILSXP:
            return None
        n = L.mock_length(s)
        if t == VECSXP:
            names = [L.mock_name(s, i).decode() for i in range(n)]
            vals = [self.value(L.mock_elt(s, i)) for i in range(n)]
            return dict(zip(names, vals)) if all(names) else vals        # an unnamed list stays a list
        if t == STRSXP:
            return [C.string_at(L.mock_data(L.mock_elt(s, i))).decode() for i in range(n)]
        ct = C.c_double if t == REALSXP else C.c_int
        a = np.ctypeslib.as_array((ct * n).from_address(L.mock_data(s))).copy() if n else np.empty(0)
        nd = L.mock_ndim(s)
        return a.reshape([L.mock_dim(s, k) for k in range(nd)], order="F") if nd else a

    def call(self, name, *args):
        """.Call(name, ...): the value, or raises RError / RInterrupt like R would signal a condition."""
        L = self.L
        fn = C.cast(getattr(L, name), C.c_void_p)
        argv = (C.c_void_p * max(1, len(args)))(*[self.sexp(a) for a in args])
        res = C.c_void_p()
        rc = L.mock_call(fn, len(args), argv, C.byref(res))
        try:
            if rc == 1:
                raise RError(L.mock_last_error().decode())
            if rc == 2:
                raise RInterrupt()
            assert rc == 0, L.mock_last_error().decode()          # 3: PROTECT stack imbalance
            return self.value(res)
        finally:
            L.mock_reset()


class RError(RuntimeError):
    pass


class RInterrupt(KeyboardInterrupt):
    pass


@pytest.fixture(scope="module")
def R():
    return MockR()


T = np.arange(11) * 0.1
OU = [1.0, 0.0, 1.0]


def run_args(**opts):
    base = {"npaths": 3.0, "nx": 1.0, "nB": 1.0, "projection": 0, "output": 2.0, "seed": 5.0}
    base.update(opts)
    return (0, 0, np.array(OU), T, np.array([0.5]), None, base)


# ---- CPU: argument checks (no device is touched before they pass) ------------------------------------------------------
@pytest.mark.parametrize("patch,msg", [
    ({"x0": np.array([0.5, 0.1])}, "x0 has length 2"),
    ({"B": np.zeros(7)}, "B (nt x nB x npaths) has length 7, expected 33"),
    ({"opts": {"Stau": np.array([0.5]), "kill_kind": 1, "kill_a": 1.0}}, "Stau (one threshold per path) has length 1, expected 3"),
    ({"opts": {"moments": 1.0, "center": np.array([0.0, 1.0])}}, "center (one shift per component) has length 2"),
    ({"opts": {"hist": np.array([0.0, 1.0])}}, "hist must be c(lo, hi, nbins"),
    ({"opts": {"hist": np.array([0.0, 1.0, -4.0])}}, "nbins must be an integer in [1, 8192]"),
    ({"opts": {"hist": np.array([0.0, 1.0, 8.0, 3.0])}}, "hist component out of range"),
    ({"opts": {"npaths": -1.0}}, "option 'npaths' must be an integer"),
    ({"opts": {"npaths": float("nan")}}, "option 'npaths' must be an integer"),
    ({"opts": {"npaths": 2.5}}, "option 'npaths' must be an integer"),
    ({"opts": {"seed": -3.0}}, "seed must be a whole number in [0, 2^53]"),
    ({"opts": {"seed": 2.0 ** 60}}, "seed must be a whole number in [0, 2^53]"),
    ({"opts": {"forcing": np.zeros(5)}}, "forcing (nt x nx) has length 5, expected 11"),
    ({"opts": {"gscale": np.zeros(12)}}, "gscale (nt) has length 12, expected 11"),
    ({"opts": {"devices": np.array([0.5])}}, "devices must be CUDA ordinals"),
    ({"opts": {"kill_comp": 4.0}}, "option 'kill_comp' must be an integer in [0, 0]"),
    ({"times": np.array([0.0])}, "times must be a double vector of length >= 2"),
    ({"x0": np.array([1], dtype=np.int32)}, "x0 has length"),
])
def test_run_argument_checks_are_r_errors(R, patch, msg):
    args = list(run_args(**patch.get("opts", {})))
    if "x0" in patch:
        args[4] = patch["x0"]
    if "B" in patch:
        args[5] = patch["B"]
    if "times" in patch:
        args[3] = patch["times"]
    with pytest.raises(RError) as e:
        R.call("C_sdegpu_run", *args)
    assert msg in str(e.value)


def test_integral_and_law_argument_checks(R):
    x = np.linspace(0, 1, 9)
    for name, args, msg in [("C_sdegpu_itointegral", (x, x[:5]), "same length"), ("C_sdegpu_crossvar", (x, x[:5]), "same length"),
                            ("C_sdegpu_stochint", (x, x, np.array([1, 0], dtype=np.int32)), "three integer flags"),
                            ("C_sdegpu_qv", (np.arange(4, dtype=np.int32),), "double vector"),
                            ("C_sdegpu_rvbm", (T, -2.0, 1.0, 0.0, 0.0, {}), "n must be a whole number"),
                            ("C_sdegpu_rlaw", (0, 5.0, np.array([1.0]), np.array([1.0, 2.0]), 0.5, {}), "bad arguments"),
                            ("C_sdegpu_lyap", (np.zeros(6), np.zeros(6)), "square matrix"),
                            ("C_sdegpu_lyap", (np.zeros((2, 2)), np.zeros((3, 3))), "same length"),
                            ("C_sdegpu_dlinsde", (np.zeros((2, 2)), np.zeros(3), 1.0, None, None, None), "as many rows"),
                            ("C_sdegpu_dlinsde", (np.zeros((2, 2)), np.eye(2), 1.0, np.zeros(3), None, None), "x0 has length 3, expected 2"),
                            ("C_sdegpu_dlinsde", (np.zeros((2, 2)), np.eye(2), 1.0, np.zeros(2), np.zeros(3), None), "u must have length")]:
        with pytest.raises(RError) as e:
            R.call(name, *args)
        assert msg in str(e.value), (name, str(e.value))


def test_without_a_device_the_shim_raises_an_r_error(R):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(RError) as e:
        R.call("C_sdegpu_run", *run_args())
    assert "no CUDA device" in str(e.value)                        # loud failure, no fallback


def test_registration_table_matches_the_entry_points():
    src = open(os.path.join(ROOT, "r", "src", "sdegpu_shim.c")).read()
    import re
    defined = {m.group(1): m.group(2).count("SEXP") for m in re.finditer(r"^SEXP (C_sdegpu_\w+)\(([^)]*)\)", src, re.M)}
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(C_sdegpu_\w+)", \(DL_FUNC\)&\1, (\d+)\}', src)}
    assert defined == table and len(table) == 12
    wrapper = open(os.path.join(ROOT, "r", "R", "sdegpu.R")).read()
    for name in table:
        assert f".Call({name}" in wrapper, name                    # every entry point has an R-side caller


# ---- GPU: the shim returns what the C ABI returns ----------------------------------------------------------------------
@pytest.fixture(scope="module")
def sde(gpu):
    import sdetools_b200
    return sdetools_b200


@pytest.mark.gpu
def test_run_through_the_shim_equals_the_c_abi(R, sde):
    rng = np.random.default_rng(1)
    t = np.arange(41) * 0.025
    B = np.cumsum(np.concatenate([np.zeros((1, 1, 6)), np.sqrt(0.025) * rng.standard_normal((40, 1, 6))]), axis=0)
    f, g = sde.catalogue.cir(1.0, 1.0, 0.8)
    want = sde.heun(f, g, t, 1.0, B, p=abs, moments=True, center=[1.0], hist=(0.0, 3.0, 12))
    got = R.call("C_sdegpu_run", 1, 1, np.array([1.0, 1.0, 0.8]), t, np.array([1.0]), B,
                 {"npaths": 6.0, "nx": 1.0, "nB": 1.0, "projection": 1, "output": 2.0, "moments": 1.0, "center": np.array([1.0]),
                  "hist": np.array([0.0, 3.0, 12.0]), "seed": 1.0})
    assert got["X"].shape == (41, 1, 6) and np.array_equal(got["X"], want["X"]) and np.array_equal(got["XT"], want["XT"])
    assert np.array_equal(got["power_sums"], want["power_sums"])
    assert np.array_equal(got["hist"][1:13], want["hist"]["counts"]) and got["seed"][0] == 1.0
    # fused generator, per-path x0, integrals and model integrals
    x0 = np.linspace(0.5, 1.5, 50)[None, :]
    want = sde.euler(f, g, t, x0, p=abs, seed=77, integrals=True, model_integrals=True)
    got = R.call("C_sdegpu_run", 0, 1, np.array([1.0, 1.0, 0.8]), t, x0, None,
                 {"npaths": 50.0, "nx": 1.0, "projection": 1, "output": 1.0, "integrals": 1.0, "model_integrals": 1.0, "seed": 77.0})
    assert got["X"] is None and np.array_equal(got["XT"], want["XT"])
    assert np.array_equal(got["integrals"][1], want["integrals"]["ito"]) and np.array_equal(got["model_integrals"][0], want["model_integrals"]["drift"])
    # time-dependent closures tabulated by the wrapper
    c, s = np.sin(2 * t), 1 + 0.1 * t
    fo = sde.catalogue.forced(f, lambda tt: np.sin(2 * tt))
    go = sde.catalogue.scaled(g, lambda tt: 1 + 0.1 * tt)
    want = sde.heun(fo, go, t, 1.0, B, p=abs)
    got = R.call("C_sdegpu_run", 1, 1, np.array([1.0, 1.0, 0.8]), t, np.array([1.0]), B,
                 {"npaths": 6.0, "projection": 1, "forcing": c, "gscale": s})
    assert np.array_equal(got["X"], want["X"])


@pytest.mark.gpu
def test_mortality_through_the_shim(R, sde):
    """S path (R/sdetools.R:443-444,456,468), tau, per-path thresholds: supplied ones are used as given, a missing Stau is
    one Philox uniform per path (the reference draws one runif per euler() call, i.e. per path of an sapply ensemble)."""
    t = np.arange(201) * 0.01
    f, g = sde.catalogue.ou(1.0, 0.0, 1.0)
    r = sde.catalogue.kill_exp(0.5, 2.0)
    stau = np.linspace(0.1, 0.9, 40)
    want = sde.euler(f, g, t, 0.0, r=r, npaths=40, seed=9, Stau=stau)
    opts = {"npaths": 40.0, "seed": 9.0, "kill_kind": 2, "kill_comp": 0, "kill_a": 0.5, "kill_b": 2.0, "S0": 1.0, "Spath": 1.0}
    got = R.call("C_sdegpu_run", 0, 0, np.array(OU), t, np.array([0.0]), None, {**opts, "Stau": stau})
    assert np.array_equal(got["tau"], want["tau"]) and np.array_equal(got["S"], want["S"])
    assert np.array_equal(got["Spath"], want["S_path"], equal_nan=True) and np.array_equal(got["X"], want["X"], equal_nan=True)
    none = R.call("C_sdegpu_run", 0, 0, np.array(OU), t, np.array([0.0]), None, opts)
    want = sde.euler(f, g, t, 0.0, r=r, npaths=40, seed=9)
    assert np.array_equal(none["tau"], want["tau"]) and len(set(none["tau"])) > 5      # independent thresholds, not one shared


@pytest.mark.gpu
def test_seed_from_r_stream_round_trips(R, sde):
    """seed = NULL draws 53 bits from unif_rand(): the returned double reproduces the run when passed back."""
    R.L.mock_set_seed(123)
    a = R.call("C_sdegpu_run", *run_args(seed=None, npaths=8.0))
    seed = a["seed"][0]
    assert seed == np.floor(seed) and 0 <= seed < 2.0 ** 53
    b = R.call("C_sdegpu_run", *run_args(seed=seed, npaths=8.0))
    assert np.array_equal(a["X"], b["X"])
    R.L.mock_set_seed(124)
    c = R.call("C_sdegpu_run", *run_args(seed=None, npaths=8.0))
    assert c["seed"][0] != seed and not np.array_equal(a["X"], c["X"])


@pytest.mark.gpu
def test_vector_utilities_through_the_shim(R, sde):
    rng = np.random.default_rng(4)
    f, g = rng.standard_normal((300, 4)), np.cumsum(rng.standard_normal((300, 4)), axis=0)
    got = R.call("C_sdegpu_stochint", f, g, np.array([1, 0, 1], dtype=np.int32))
    want = sde.stochint(f, g, rule=["l", "c"])
    assert np.array_equal(got[0], want["l"]) and got[1] is None and np.array_equal(got[2], want["c"])
    assert np.array_equal(R.call("C_sdegpu_itointegral", f[:, 0], g[:, 0]), sde.itointegral(f[:, 0], g[:, 0]))
    assert np.array_equal(R.call("C_sdegpu_qv", g), sde.QuadraticVariation(g))
    assert np.array_equal(R.call("C_sdegpu_crossvar", f, g), sde.CrossVariation(f, g))
    assert np.array_equal(R.call("C_sdegpu_rvbm", T, 7.0, 2.0, 0.5, 0.1, {"seed": 3.0}), sde.rvBM(T, 7, 2.0, 0.5, 0.1, seed=3))
    fun = R.call("C_sdegpu_bm_functionals", T, 7.0, 1.0, 0.0, 0.0, {"seed": 3.0, "level": 0.4})
    want = sde.rvBM_functionals(T, 7, level=0.4, seed=3)
    assert all(np.array_equal(fun[k], want[k], equal_nan=True) for k in ("max", "min", "tmax", "tmin", "thit"))
    assert np.array_equal(R.call("C_sdegpu_rlaw", 1, 9.0, np.array([1.0]), np.array([1.0, 1.0, 0.5]), 0.3, {"seed": 2.0, "steps": 3.0}),
                          sde.rCIR(9, 1.0, 1.0, 1.0, 0.5, 0.3, steps=3, seed=2))
    x = np.linspace(0.1, 2, 11)
    assert np.array_equal(R.call("C_sdegpu_law_eval", 1, 1, x, np.array([1.0]), np.array([1.0, 1.0, 0.5]), 0.3, 0), sde.pCIR(x, 1.0, 1.0, 1.0, 0.5, 0.3))
    A = np.array([[0.0, 1.0], [-1.0, -0.1]])
    assert np.allclose(R.call("C_sdegpu_lyap", A, np.diag([0.0, 1.0])), 5 * np.eye(2))
    sol = R.call("C_sdegpu_dlinsde", np.array([[-1.0]]), np.array([[1.0]]), 1.0, None, None, np.array([[0.5]]))
    assert list(sol) == ["eAt", "St"] and sol["eAt"].item() == pytest.approx(np.exp(-1)) and sol["St"].item() == pytest.approx(0.5)
    sol = R.call("C_sdegpu_dlinsde", np.array([[0.0]]), np.array([[1.0]]), 2.0, np.array([1.0]), np.array([[1.0, 3.0]]), None)
    assert list(sol) == ["EX", "St"] and sol["EX"][0] == pytest.approx(5)          # tests/testthat/test_solvers.R:34-36


@pytest.mark.gpu
def test_pending_interrupt_stops_a_long_run(R, sde):
    """The callback the shim registers polls R_CheckUserInterrupt under R_ToplevelExec between chunks; the run ends with an
    R condition, the PROTECT stack unwound, and the context stays usable."""
    t = np.arange(20001) * 1e-4
    big = (1, 1, np.array([1.0, 1.0, 1.0]), t, np.array([1.0]), None,
           {"npaths": 1.2e7, "projection": 1, "output": 0.0, "moments": 1.0, "seed": 1.0})
    R.L.mock_interrupt_after(2)                                    # the third poll finds Ctrl-C pending
    with pytest.raises((RInterrupt, RError)) as e:
        R.call("C_sdegpu_run", *big)
    assert e.type is RInterrupt or "interrupted" in str(e.value)
    small = R.call("C_sdegpu_run", *run_args(npaths=4.0))
    assert small["X"].shape == (11, 1, 4) and np.isfinite(small["X"]).all()


@pytest.mark.gpu
def test_devices_option_runs_on_several_gpus(R, sde):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    opts = {"npaths": 30001.0, "projection": 1, "output": 1.0, "moments": 1.0, "center": np.array([1.0]), "hist": np.array([0.0, 6.0, 24.0]), "seed": 4.0}
    t = np.arange(101) * 0.01
    one = R.call("C_sdegpu_run", 1, 1, np.array([1.0, 1.0, 1.0]), t, np.array([1.0]), None, {**opts, "devices": np.array([0.0])})
    two = R.call("C_sdegpu_run", 1, 1, np.array([1.0, 1.0, 1.0]), t, np.array([1.0]), None, {**opts, "devices": np.array([0.0, 1.0])})
    assert np.array_equal(one["XT"], two["XT"]) and np.array_equal(one["hist"], two["hist"])
    assert np.allclose(one["power_sums"], two["power_sums"], rtol=1e-12)
    info = R.call("C_sdegpu_info")
    assert list(info["devices"]) == [0, 1] and info["reduction"][0] in ("nccl", "host")
