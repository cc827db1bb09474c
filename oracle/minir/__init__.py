"""TEST INFRASTRUCTURE ONLY — `oracle.minir`: a minimal evaluator for the R the reference is written in.

It exists so that parity can be pinned to the reference's own source text (`/root/reference/R/sdetools.R`) in an image that has
no GNU R.  See interp.py for the scope; tests/golden/make_minir_fixtures.py for how the committed fixtures were made.

    from oracle.minir import Interp, to_numpy, from_numpy
    R = Interp(); R.source("/root/reference/R/sdetools.R")
    sol = R.call("euler", f, g, from_numpy(times), from_numpy(x0), B=from_numpy(B))

    python -m oracle.minir script.R [args...]          # an Rscript stand-in
"""
import numpy as np

from .interp import NULL, Interp, RError, RList, RNull, RVec, dbl  # noqa: F401


def from_numpy(a):
    """A numpy array (any rank, C or F order) as an R double vector / matrix / array (column-major data, dim attribute)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim <= 1:
        return dbl(a.reshape(-1))
    from .interp import int_
    return RVec("double", a.reshape(-1, order="F"), {"dim": int_(list(a.shape))})


def to_numpy(x):
    """An R vector / matrix / array as a numpy array of the same shape (lists become dicts of arrays)."""
    if isinstance(x, RNull):
        return None
    if isinstance(x, RList):
        names = x.names()
        return {n or i: to_numpy(v) for i, (n, v) in enumerate(zip(names, x.v))}
    if not isinstance(x, RVec):
        return x                                   # functions, language objects: handed back as they are
    d = x.dim()
    v = np.array(x.v, dtype=np.float64) if x.kind != "character" else np.array(list(x.v), dtype=object)
    return v if d is None else v.reshape(d, order="F")
