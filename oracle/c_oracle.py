"""ctypes front-end to oracle/c/sde_oracle.c.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdeoracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "sde_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_ensemble.restype = C.c_int
        _lib.oracle_ensemble.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                         _dp, C.c_uint64, C.c_uint64, C.c_uint64, _dp, _dp, C.c_int]
        _lib.oracle_integral.argtypes = [C.c_int, _dp, _dp, C.c_size_t, C.c_size_t, _dp]
        _lib.oracle_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _dp]
        _lib.oracle_ld_add.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint16), C.c_double]
        _lib.oracle_rBM.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        _lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def ensemble(scheme, model, params, nx, nB, proj, times, x0, B=None, seed=0, path_offset=0, npaths=None,
             full=True, terminal=True, nthreads=1):
    """B: (npaths, nB, nt) C-contiguous == reference layout B[(p*nB+k)*nt+i].
    Returns (X (npaths, nx, nt) or None, XT (npaths, nx) or None)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    params = np.ascontiguousarray(np.ravel(params), dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    stride = 0 if x0.size == nx else nx
    if B is not None:
        B = np.ascontiguousarray(B, dtype=np.float64)
        npaths = B.shape[0]
    nt = times.shape[0]
    X = np.empty((npaths, nx, nt)) if full else None
    XT = np.empty((npaths, nx)) if terminal else None
    rc = lib().oracle_ensemble(scheme, model, _p(params), nx, nB, proj, nt, _p(times), _p(x0), stride, _p(B),
                               seed, path_offset, npaths, _p(X), _p(XT), nthreads)
    if rc != 0:
        raise RuntimeError("oracle_ensemble failed")
    return X, XT


def integral(which, a, b=None):
    """a,b: (ncols, n) C-contiguous (== n x ncols column-major)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    a2 = a.reshape(-1, a.shape[-1])
    out = np.empty_like(a2)
    lib().oracle_integral(which, _p(a2), _p(b), a2.shape[1], a2.shape[0], _p(out))
    return out.reshape(a.shape)


def normals(seed, path, channel, nsteps):
    z = np.empty(nsteps)
    lib().oracle_normals(seed, path, channel, nsteps, _p(z))
    return z


def ld_add(mant: int, sexp: int, d: float):
    m, s = C.c_uint64(mant), C.c_uint16(sexp)
    lib().oracle_ld_add(C.byref(m), C.byref(s), d)
    return m.value, s.value
