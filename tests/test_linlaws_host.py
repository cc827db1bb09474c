"""CPU: the host logic of csrc/linlaws.cu (sdegpu_lyap / sdegpu_dlinsde) against the oracle's restatement of the
reference's lyap / dLinSDE (R/sdetools.R:593-600, :508-572) and against scipy.  A test-only shared object is
built from the same source with the device-GEMM threshold raised out of reach, so it runs here without a GPU; the GPU
tests cover the library entry points and the device GEMM."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import linalg

from oracle import laws

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "native", "_build")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "liblinlaws_test.so")
    srcs = [os.path.join(ROOT, "sdetools_b200", "csrc", "linlaws.cu"), os.path.join(ROOT, "tests", "native", "linlaws_host.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        r = subprocess.run(["nvcc", "-shared", "-Xcompiler", "-fPIC", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-DSDEGPU_LINLAWS_DEVICE_MIN=1000000", "-I", os.path.join(ROOT, "include"), "-x", "cu", *srcs, "-o", so, "-cudart", "static"],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    L.t_why.restype = C.c_char_p
    return L


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lyap(lib, A, Q):
    A, Q = np.asfortranarray(A, dtype=float), np.asfortranarray(Q, dtype=float)
    X = np.empty_like(A, order="F")
    return lib.t_lyap(P(A), P(Q), A.shape[0], P(X)), X


def dlinsde(lib, A, G, t, x0=None, u=None, S0=None):
    A = np.asfortranarray(np.atleast_2d(A), dtype=float)
    d = A.shape[0]
    G = np.asfortranarray(np.asarray(G, dtype=float).reshape(d, -1))
    x0k = None if x0 is None else np.ascontiguousarray(np.ravel(x0), dtype=float)
    uk = None if u is None else np.asfortranarray(np.asarray(u, dtype=float).reshape(d, -1))
    S0k = None if S0 is None else np.asfortranarray(S0, dtype=float)
    St, eAt, EX = np.empty((d, d), order="F"), np.empty((d, d), order="F"), np.empty(d)
    rc = lib.t_dlinsde(P(A), P(G), d, G.shape[1], C.c_double(t), P(x0k), P(uk), 0 if uk is None else uk.shape[1], P(S0k),
                       P(eAt), P(EX) if x0 is not None else None, P(St))
    assert rc == 0, lib.t_why()
    return ({"eAt": eAt, "St": St} if x0 is None else {"EX": EX, "St": St})


def test_reference_known_answers(lib):
    """tests/testthat/test_solvers.R:20-36 and the doc examples R/sdetools.R:587-590."""
    tol = 1.5e-8
    sol = dlinsde(lib, -1.0, 1.0, 1.0, S0=np.array([[0.5]]))
    assert sol["eAt"].item() == pytest.approx(np.exp(-1), rel=tol) and sol["St"].item() == pytest.approx(0.5, rel=tol)
    A = np.array([[0.0, 1.0], [-1.0, 0.0]])
    sol = dlinsde(lib, A, np.eye(2), 2 * np.pi, x0=[1, 0])          # singular Lyapunov equation: Van Loan route
    assert np.allclose(sol["EX"], [1, 0], atol=tol) and np.allclose(sol["St"], 2 * np.pi * np.eye(2), atol=10 * tol)
    assert dlinsde(lib, -1.0, 1.0, 2.0, x0=[1.0], u=[1.0])["EX"][0] == pytest.approx(1, rel=tol)
    assert dlinsde(lib, 0.0, 1.0, 2.0, x0=[1.0], u=np.array([[1.0, 3.0]]))["EX"][0] == pytest.approx(5, rel=tol)
    rc, X = lyap(lib, [[-1.0]], [[1.0]])
    assert rc == 0 and X.item() == pytest.approx(0.5)
    rc, X = lyap(lib, np.array([[0.0, 1.0], [-1.0, -0.1]]), np.diag([0.0, 1.0]))
    assert rc == 0 and np.allclose(X, 5 * np.eye(2))
    rc, _ = lyap(lib, A, np.eye(2))                                  # lambda_i + lambda_j = 0: the reference's solve() fails too
    assert rc == -1 and b"singular" in lib.t_why()


@pytest.mark.parametrize("d", [1, 3, 7, 24, 41, 64, 90])
def test_against_the_oracle_restatement_and_scipy(lib, d):
    """d <= 40 takes the reference's Kronecker system, d > 40 the bilinear transform + doubling."""
    rng = np.random.default_rng(d)
    A = -np.eye(d) + 0.4 * rng.standard_normal((d, d)) / np.sqrt(d)
    G = rng.standard_normal((d, max(1, d - 1))) / np.sqrt(d)
    S0 = np.diag(rng.uniform(0, 1, d))
    x0, u1, u2 = rng.standard_normal(d), rng.standard_normal(d), rng.standard_normal((d, 2))
    ref = laws if d <= 7 else None                              # the restated Kronecker solve is O(d^6)
    for kw in ({}, {"x0": x0}, {"x0": x0, "u": u1}, {"x0": x0, "u": u2}, {"S0": S0}):
        a = dlinsde(lib, A, G, 0.7, **kw)
        if ref is not None:
            b = ref.dLinSDE(A, G, 0.7, **kw)
            assert a.keys() == b.keys()
            for k in a:
                assert np.allclose(a[k], np.asarray(b[k]).reshape(a[k].shape), rtol=1e-10, atol=1e-12), (d, list(kw), k)
    eAt = linalg.expm(0.7 * A)
    Sinf = linalg.solve_continuous_lyapunov(A, -G @ G.T)
    sol = dlinsde(lib, A, G, 0.7)
    assert np.allclose(sol["eAt"], eAt, rtol=1e-11, atol=1e-13)
    assert np.allclose(sol["St"], Sinf - eAt @ Sinf @ eAt.T, rtol=1e-9, atol=1e-12)
    rc, X = lyap(lib, A, G @ G.T)
    assert rc == 0 and np.allclose(X, Sinf, rtol=1e-10, atol=1e-12)
    assert np.allclose(A @ X + X @ A.T + G @ G.T, 0, atol=1e-12)
    AA = np.block([[A, np.eye(d)], [np.zeros((d, 2 * d))]])
    assert np.allclose(dlinsde(lib, A, G, 0.7, x0=x0, u=u1)["EX"], (linalg.expm(0.7 * AA) @ np.concatenate([x0, u1]))[:d], rtol=1e-10, atol=1e-12)


def test_anti_stable_and_mixed_spectra_beyond_the_kronecker_range(lib):
    d = 48
    rng = np.random.default_rng(5)
    A = np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)      # all eigenvalues in the right half plane
    Q = np.eye(d)
    rc, X = lyap(lib, A, Q)
    assert rc == 0 and np.allclose(A @ X + X @ A.T + Q, 0, atol=1e-11)
    Am = np.diag(np.concatenate([-np.ones(d // 2), 2 * np.ones(d // 2)]))  # mixed: refused with an explanation ...
    rc, _ = lyap(lib, Am, Q)
    assert rc == -1 and b"imaginary axis" in lib.t_why()
    sol = dlinsde(lib, Am, np.eye(d), 0.5)                              # ... and dLinSDE falls back to Van Loan, like :525-533
    lam = np.diag(Am)
    assert np.allclose(np.diag(sol["St"]), (np.exp(2 * lam * 0.5) - 1) / (2 * lam), rtol=1e-10)


def test_stiff_system_and_large_norm(lib):
    """A spectrum spread over six decades (doubling needs ~25 squarings) and a norm that needs 12 Pade squarings."""
    d = 50
    lam = -np.logspace(-3, 3, d)
    V = np.linalg.qr(np.random.default_rng(1).standard_normal((d, d)))[0]
    A = V @ np.diag(lam) @ V.T
    rc, X = lyap(lib, A, np.eye(d))
    assert rc == 0 and np.allclose(X, V @ np.diag(-0.5 / lam) @ V.T, rtol=1e-8, atol=1e-12)
    sol = dlinsde(lib, A, np.eye(d), 10.0)
    assert np.allclose(sol["eAt"], V @ np.diag(np.exp(10.0 * lam)) @ V.T, atol=1e-11)
