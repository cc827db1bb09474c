"""TEST INFRASTRUCTURE ONLY — an evaluator for the subset of R that the reference's sample-path code is written in.

Purpose: GNU R is not installed in the build image, so the reference (`/root/reference/R/sdetools.R`, pure R) cannot
run here.  Instead of trusting a hand restatement alone, this evaluator runs the reference's SOURCE TEXT: `source()` parses the
unmodified file and `euler`, `heun`, `stochint`, `itointegral`, `QuadraticVariation`, `CrossVariation`, `rBM`, `rvBM`, `lyap`,
`dLinSDE`, ... are then the reference's own function bodies, evaluated statement by statement.  What is restated here is only the
slice of BASE R those bodies call (base R is not part of the reference either):

* vectors with attributes (names, dim, dimnames), recycling arithmetic — ONE correctly rounded IEEE-754 operation per operator
  and element (numpy ufuncs; no fused multiply-add), `x^2` as `x*x` (R's arithmetic.c R_POW), left-to-right evaluation;
* `cumsum`/`sum` accumulate in the 80-bit x87 `long double` and round to double per element, as R's cum.c/summary.c do
  (numpy.longdouble is that type on x86-64);
* `%*%`: dot products accumulated in ascending k in double (what the reference BLAS dgemm/dgemv that R falls back to does);
* lazy arguments (promises): a default such as `Stau = runif(1)` is evaluated on first use; function lookup skips non-functions;
* `[`, `[[`, `$` and their replacement forms incl. nested targets, drop = TRUE semantics, negative and logical subscripts;
* `apply`, `sapply`, `matrix`, `array`, `diff`, `seq`, `rep`, `c`, `list`, `data.frame` column selection, `tryCatch(error=)`.
Stand-ins that are NOT base R's algorithms (results agree to rounding, never claimed bit-exact): `Matrix::expm` -> scipy.linalg.expm,
`solve` -> LAPACK through numpy (R also calls dgesv), `exp/sin/cos/log` -> numpy's libm, `dnorm/pnorm/dchisq/pchisq/...` -> scipy.

Nothing in the product imports this package (tests/test_abi.py::test_product_never_imports_oracle).
"""
import math
import os
import re
import struct
import sys

import numpy as np

from .parser import parse

NA_REAL = struct.unpack("<d", struct.pack("<Q", 0x7FF00000000007A2))[0]     # R's NA_real_: a NaN whose low word is 1954
KINDS = ["logical", "integer", "double", "character"]


def is_na_payload(x):
    return x != x and (struct.unpack("<Q", struct.pack("<d", x))[0] & 0xFFFFFFFF) == 1954


class RError(Exception):
    pass


class RNull:
    attrs = {}

    def __repr__(self):
        return "NULL"


NULL = RNull()


class RVec:
    __slots__ = ("kind", "v", "attrs")

    def __init__(self, kind, v, attrs=None):
        self.kind = kind
        if kind == "character":
            a = np.empty(len(v), dtype=object)
            a[:] = list(v)
            self.v = a
        else:
            self.v = np.array(v, dtype=np.float64).reshape(-1)
        self.attrs = dict(attrs) if attrs else {}

    def __len__(self):
        return len(self.v)

    def dim(self):
        d = self.attrs.get("dim")
        return None if d is None else [int(x) for x in d.v]

    def __repr__(self):
        return f"RVec({self.kind}, {list(self.v)[:8]}, dim={self.dim()})"


class RList:
    __slots__ = ("v", "attrs")

    def __init__(self, v, names=None, attrs=None):
        self.v = list(v)
        self.attrs = dict(attrs) if attrs else {}
        if names is not None and any(n for n in names):
            self.attrs["names"] = chr_(["" if n is None else n for n in names])

    def __len__(self):
        return len(self.v)

    def names(self):
        n = self.attrs.get("names")
        return [""] * len(self.v) if n is None else list(n.v)


def dbl(v, **attrs):
    return RVec("double", np.atleast_1d(np.asarray(v, dtype=np.float64)), attrs)


def int_(v, **attrs):
    return RVec("integer", np.atleast_1d(np.asarray(v, dtype=np.float64)), attrs)


def lgl(v, **attrs):
    return RVec("logical", np.atleast_1d(np.asarray(v, dtype=np.float64)), attrs)


def chr_(v, **attrs):
    return RVec("character", [v] if isinstance(v, str) else list(v), attrs)


class Closure:
    def __init__(self, params, body, env, src=None):
        self.params, self.body, self.env, self.src = params, body, env, src
        self.attrs = {}


class Builtin:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn
        self.attrs = {}


class RLang:
    """An unevaluated expression (quote(x), body(f))."""
    attrs = {}

    def __init__(self, node):
        self.node = node


class Promise:
    __slots__ = ("expr", "env", "value", "forced", "is_default")

    def __init__(self, expr, env, is_default=False):
        self.expr, self.env, self.value, self.forced, self.is_default = expr, env, None, False, is_default


class Missing:
    pass


MISSING = Missing()


class Env:
    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e.vars[name]
            e = e.parent
        raise RError(f"object '{name}' not found")

    def has(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return True
            e = e.parent
        return False

    def set(self, name, value):
        self.vars[name] = value

    def set_super(self, name, value):
        e = self.parent
        while e is not None:
            if name in e.vars or e.parent is None or e.parent.parent is None:
                e.vars[name] = value
                return
            e = e.parent


class ReturnEx(Exception):
    def __init__(self, value):
        self.value = value


class BreakEx(Exception):
    pass


class NextEx(Exception):
    pass


def is_fn(x):
    return isinstance(x, (Closure, Builtin))


def as_bool(x, what="condition"):
    if not isinstance(x, RVec) or len(x) == 0:
        raise RError(f"argument is of length zero ({what})")
    v = x.v[0]
    if x.kind == "character":
        raise RError("argument is not interpretable as logical")
    if v != v:
        raise RError("missing value where TRUE/FALSE needed")
    return bool(v != 0)


def scalar(x):
    if isinstance(x, RVec) and len(x) >= 1:
        return x.v[0]
    raise RError("scalar expected")


def sint(x):
    return int(scalar(x))


def sstr(x):
    v = scalar(x)
    return v if isinstance(v, str) else fmt_num(v)


def fmt_num(v, kind="double"):
    if v != v:
        return "NA" if is_na_payload(v) else "NaN"
    if kind == "logical":
        return "TRUE" if v else "FALSE"
    if math.isinf(v):
        return "Inf" if v > 0 else "-Inf"
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return "%.15g" % v


def c_hexfloat(x):
    """C99 printf("%a") as glibc prints it (what R's sprintf("%a") gives): the shortest exact hexadecimal mantissa."""
    if x != x:
        return "NA" if is_na_payload(x) else "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    h = float(x).hex()                               # [-]0x1.8000000000000p+1
    m = re.match(r"(-?)0x([01])\.([0-9a-f]+)p([+-]\d+)", h)
    sign, lead, frac, ex = m.groups()
    frac = frac.rstrip("0")
    return f"{sign}0x{lead}{'.' + frac if frac else ''}p{ex}"


def higher_kind(a, b):
    return KINDS[max(KINDS.index(a), KINDS.index(b))]


def recycle(v, n):
    m = len(v)
    if m == n:
        return v
    if m == 0:
        return v
    if m == 1:
        return np.repeat(v, n)
    return np.resize(v, n)


def copy_vec(x):
    return RVec(x.kind, x.v.copy(), x.attrs)


def strip(x):
    return RVec(x.kind, x.v, None)


def mat(x):
    """An RVec as a 2-D numpy array (column-major data)."""
    d = x.dim()
    if d is None:
        return x.v.reshape(-1, 1)
    if len(d) != 2:
        raise RError("matrix expected")
    return x.v.reshape(d, order="F")


def from_mat(a, kind="double", dimnames=None):
    a = np.asarray(a, dtype=np.float64)
    r = RVec(kind, a.reshape(-1, order="F"), {"dim": int_(list(a.shape))})
    if dimnames is not None:
        r.attrs["dimnames"] = dimnames
    return r


def dimnames_of(x):
    dn = x.attrs.get("dimnames")
    if dn is None:
        return None
    return [None if isinstance(e, RNull) else list(e.v) for e in dn.v]


def make_dimnames(lst):
    if lst is None or all(e is None for e in lst):
        return None
    return RList([NULL if e is None else chr_(e) for e in lst])


# ------------------------------------------------------------------------------------------------------------------------------
# arithmetic: one IEEE operation per operator and element
# ------------------------------------------------------------------------------------------------------------------------------
def arith(op, a, b):
    if not isinstance(a, RVec) or not isinstance(b, RVec) or a.kind == "character" or b.kind == "character":
        raise RError(f"non-numeric argument to binary operator {op}")
    na, nb = len(a), len(b)
    n = 0 if (na == 0 or nb == 0) else max(na, nb)
    x, y = recycle(a.v, n), recycle(b.v, n)
    both_int = a.kind in ("integer", "logical") and b.kind in ("integer", "logical")
    with np.errstate(all="ignore"):
        if op == "+":
            r = x + y
        elif op == "-":
            r = x - y
        elif op == "*":
            r = x * y
        elif op == "/":
            r, both_int = x / y, False
        elif op == "^":
            # R's arithmetic.c (R_POW / R_pow): y == 2 is x*x, everything else C's pow
            if nb == 1 and y.size and y[0] == 2.0:
                r = x * x
            else:
                r = np.where(y == 2.0, x * x, np.power(x, y))
                r = np.where((x == 1.0) | (y == 0.0), 1.0, r)
            both_int = False
        elif op == "%%":
            r = x - np.floor(x / y) * y
        elif op == "%/%":
            r = np.floor(x / y)
        else:
            raise RError(f"unknown operator {op}")
    attrs = {}
    for s in (b, a):                                  # attributes of both operands, the first one's taking precedence
        if len(s) == n:
            attrs.update(s.attrs)
    return RVec("integer" if both_int else "double", r, attrs)


def compare(op, a, b):
    if not isinstance(a, RVec) or not isinstance(b, RVec):
        raise RError("comparison of these types is not implemented")
    na, nb = len(a), len(b)
    n = 0 if (na == 0 or nb == 0) else max(na, nb)
    x, y = recycle(a.v, n), recycle(b.v, n)
    if a.kind == "character" or b.kind == "character":
        xs = [v if isinstance(v, str) else fmt_num(v) for v in x]
        ys = [v if isinstance(v, str) else fmt_num(v) for v in y]
        f = {"==": lambda p, q: p == q, "!=": lambda p, q: p != q, "<": lambda p, q: p < q, ">": lambda p, q: p > q,
             "<=": lambda p, q: p <= q, ">=": lambda p, q: p >= q}[op]
        return lgl([f(p, q) for p, q in zip(xs, ys)])
    with np.errstate(all="ignore"):
        r = {"==": np.equal, "!=": np.not_equal, "<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal}[op](x, y)
    r = r.astype(np.float64)
    r[np.isnan(x) | np.isnan(y)] = NA_REAL
    attrs = {}
    for s in (b, a):
        if len(s) == n and "dim" in s.attrs:
            attrs["dim"] = s.attrs["dim"]
    return RVec("logical", r, attrs)


def matprod(a, b):
    """R's %*% (array.c do_matprod): vectors are promoted to row or column matrices so that the product conforms; each
    element is sum_k a[i,k]*b[k,j] accumulated in ascending k with one rounded multiply and one rounded add per term."""
    da, db = a.dim(), b.dim()
    if da is None and db is None:
        if len(a) == len(b):
            A, B = a.v.reshape(1, -1), b.v.reshape(-1, 1)
        elif len(a) == 1:
            A, B = a.v.reshape(1, 1), b.v.reshape(1, -1)
        else:
            A, B = a.v.reshape(-1, 1), b.v.reshape(1, -1)
    elif da is None:
        B = mat(b)
        A = a.v.reshape(1, -1) if len(a) == B.shape[0] else a.v.reshape(-1, 1)
    elif db is None:
        A = mat(a)
        B = b.v.reshape(-1, 1) if len(b) == A.shape[1] else b.v.reshape(1, -1)
    else:
        A, B = mat(a), mat(b)
    if A.shape[1] != B.shape[0]:
        raise RError("non-conformable arguments")
    K = A.shape[1]
    out = np.zeros((A.shape[0], B.shape[1]))
    with np.errstate(all="ignore"):
        for j in range(B.shape[1]):
            if K == 0:
                continue
            acc = A[:, 0] * B[0, j]
            for k in range(1, K):
                acc = acc + A[:, k] * B[k, j]
            out[:, j] = acc
    dn = None
    dna, dnb = (dimnames_of(a) if da else None), (dimnames_of(b) if db else None)
    if dna or dnb:
        dn = make_dimnames([dna[0] if dna else None, dnb[1] if dnb else None])
    return from_mat(out, dimnames=dn)


def cumsum_ld(v):
    """R's cumsum (cum.c): LDOUBLE accumulator, each element rounded to double on store."""
    out = np.empty(len(v))
    acc = np.longdouble(0.0)
    for i, x in enumerate(v):
        acc = acc + np.longdouble(x)
        out[i] = np.float64(acc)
    return out


def sum_ld(v):
    acc = np.longdouble(0.0)
    for x in v:
        acc = acc + np.longdouble(x)
    return float(np.float64(acc))


# ------------------------------------------------------------------------------------------------------------------------------
# subscripts
# ------------------------------------------------------------------------------------------------------------------------------
def resolve_sub(ix, n, names=None, allow_extend=False):
    """One subscript -> array of 0-based positions."""
    if ix is None:
        return np.arange(n)
    if isinstance(ix, RNull):
        return np.arange(0)
    if not isinstance(ix, RVec):
        raise RError("invalid subscript type")
    if ix.kind == "character":
        out = []
        for s in ix.v:
            if names is None or s not in names:
                if allow_extend:
                    out.append(-1)
                    continue
                raise RError(f"subscript out of bounds: '{s}'")
            out.append(names.index(s))
        return np.array(out, dtype=np.int64)
    if ix.kind == "logical":
        m = recycle(ix.v, max(n, len(ix.v)))
        return np.nonzero(m != 0)[0]
    v = ix.v
    if len(v) == 0:
        return np.arange(0)
    if np.any(np.isnan(v)):
        raise RError("NA subscripts are not supported")
    v = np.trunc(v).astype(np.int64)
    if np.all(v <= 0):
        keep = np.ones(n, dtype=bool)
        drop = -v[v < 0] - 1
        keep[drop[drop < n]] = False
        return np.nonzero(keep)[0]
    if np.any(v < 0):
        raise RError("can't mix positive and negative subscripts")
    v = v[v > 0] - 1
    if not allow_extend and np.any(v >= n):
        raise RError("subscript out of bounds")
    return v


def index_vec(x, args, named, double):
    drop = True
    if "drop" in named:
        drop = as_bool(named["drop"])
    if isinstance(x, RNull):
        return NULL
    if isinstance(x, RList):
        return index_list(x, args, double, drop)
    if not isinstance(x, RVec):
        raise RError("object of this type is not subsettable")
    d = x.dim()
    if len(args) == 1 or d is None:
        if len(args) != 1:
            raise RError("incorrect number of dimensions")
        nm = x.attrs.get("names")
        names = list(nm.v) if nm is not None else None
        if isinstance(args[0], RVec) and args[0].dim() is not None and d is not None and args[0].kind != "logical" and \
                len(args[0].dim()) == 2 and args[0].dim()[1] == len(d) and len(d) > 1:
            m = mat(args[0]).astype(np.int64) - 1
            pos = np.zeros(m.shape[0], dtype=np.int64)
            mul = 1
            for k, dk in enumerate(d):
                pos += m[:, k] * mul
                mul *= dk
        else:
            pos = resolve_sub(args[0], len(x), names)
        if double and len(pos) != 1:
            raise RError("subscript out of bounds")
        if np.any(pos >= len(x)):
            vals = np.array([x.v[p] if p < len(x) else (NA_REAL if x.kind != "character" else "NA") for p in pos],
                            dtype=x.v.dtype)
        else:
            vals = x.v[pos]
        r = RVec(x.kind, vals)
        if names is not None and not double:
            r.attrs["names"] = chr_([names[p] for p in pos])
        return r
    if len(args) != len(d):
        raise RError("incorrect number of dimensions")
    dn = dimnames_of(x) or [None] * len(d)
    subs = [resolve_sub(a, d[k], dn[k]) for k, a in enumerate(args)]
    arr = x.v.reshape(d, order="F")
    res = arr[np.ix_(*subs)]
    shape = list(res.shape)
    rdn = [None if dn[k] is None else [dn[k][p] for p in subs[k]] for k in range(len(d))]
    if drop:
        keep = [k for k in range(len(shape)) if shape[k] != 1]
        if len(keep) <= 1:
            r = RVec(x.kind, res.reshape(-1, order="F"))
            if len(keep) == 1 and rdn[keep[0]] is not None:
                r.attrs["names"] = chr_(rdn[keep[0]])
            return r
        shape = [shape[k] for k in keep]
        rdn = [rdn[k] for k in keep]
    r = RVec(x.kind, res.reshape(-1, order="F"), {"dim": int_(shape)})
    mdn = make_dimnames(rdn)
    if mdn is not None:
        r.attrs["dimnames"] = mdn
    return r


def index_list(x, args, double, drop=True):
    names = x.names()
    is_df = "class" in x.attrs and "data.frame" in list(x.attrs["class"].v)
    if is_df and len(args) == 2:
        rows, cols = args
        pos = resolve_sub(cols, len(x), names)
        colv = [x.v[p] for p in pos]
        if rows is not None:
            colv = [index_vec(c, [rows], {}, False) for c in colv]
        if len(pos) == 1 and drop:
            return colv[0]
        return RList(colv, [names[p] for p in pos], {"class": chr_("data.frame")})
    if len(args) != 1:
        raise RError("incorrect number of subscripts on a list")
    pos = resolve_sub(args[0], len(x), names, allow_extend=True)
    if double:
        if len(pos) != 1:
            raise RError("subscript out of bounds")
        if pos[0] < 0 or pos[0] >= len(x):
            if args[0].kind == "character":
                return NULL
            raise RError("subscript out of bounds")
        return x.v[pos[0]]
    keep_attrs = {k: v for k, v in x.attrs.items() if k == "class"}
    return RList([x.v[p] if 0 <= p < len(x) else NULL for p in pos], [names[p] if 0 <= p < len(x) else "NA" for p in pos], keep_attrs)


def assign_index(x, args, named, value, double):
    if isinstance(x, RNull):
        x = RList([]) if (double or isinstance(value, (RList, Closure, Builtin))) else RVec(value.kind if isinstance(value, RVec) else "logical", [])
    if isinstance(x, RList):
        names = x.names()
        if len(args) != 1:
            raise RError("incorrect number of subscripts on a list")
        pos = resolve_sub(args[0], len(x), names, allow_extend=True)
        v, nm = list(x.v), list(names)
        vals = [value] if double else (list(value.v) if isinstance(value, RList) else [value])
        for j, p in enumerate(pos):
            val = vals[j % len(vals)]
            if p < 0:                                           # a new name
                v.append(val)
                nm.append(str(args[0].v[j]))
            else:
                while p >= len(v):
                    v.append(NULL)
                    nm.append("")
                v[p] = val
        return RList(v, nm, {k: a for k, a in x.attrs.items() if k != "names"})
    if not isinstance(x, RVec):
        raise RError("object is not subsettable for assignment")
    if isinstance(value, RList):
        raise RError("assigning a list into an atomic vector is not supported")
    if not isinstance(value, RVec):
        raise RError("invalid replacement value")
    kind = higher_kind(x.kind, value.kind)
    d = x.dim()

    def conv(vec, from_kind):
        if kind == "character" and from_kind != "character":
            return np.array([fmt_num(e, from_kind) for e in vec], dtype=object)
        return vec

    data = conv(x.v, x.kind).copy()
    val = conv(value.v, value.kind)
    if len(args) == 1 or d is None:
        if len(args) != 1:
            raise RError("incorrect number of subscripts")
        nm = x.attrs.get("names")
        names = list(nm.v) if nm is not None else None
        pos = resolve_sub(args[0], len(x), names, allow_extend=True)
        new_names = None
        if len(pos) and (np.any(pos >= len(data)) or np.any(pos < 0)):
            if args[0].kind == "character":
                new_names = list(names) if names is not None else [""] * len(data)
                pos = pos.copy()
                for j, p in enumerate(pos):
                    if p < 0:
                        s = args[0].v[j]
                        if s in new_names:
                            pos[j] = new_names.index(s)
                        else:
                            new_names.append(s)
                            pos[j] = len(new_names) - 1
            n2 = int(pos.max()) + 1
            fill = "NA" if kind == "character" else NA_REAL
            ext = np.empty(n2, dtype=data.dtype)
            ext[:len(data)] = data
            ext[len(data):] = fill
            data = ext
            if new_names is None and names is not None:
                new_names = names + [""] * (n2 - len(names))
        if len(pos):
            if len(val) == 0:
                raise RError("replacement has length zero")
            data[pos] = recycle(val, len(pos))
        attrs = dict(x.attrs)
        if len(data) != len(x.v):
            attrs.pop("dim", None)
            attrs.pop("dimnames", None)
        if new_names is not None:
            attrs["names"] = chr_(new_names)
        return RVec(kind, data, attrs)
    if len(args) != len(d):
        raise RError("incorrect number of subscripts")
    dn = dimnames_of(x) or [None] * len(d)
    subs = [resolve_sub(a, d[k], dn[k]) for k, a in enumerate(args)]
    cnt = 1
    for s in subs:
        cnt *= len(s)
    if cnt:
        if len(val) == 0:
            raise RError("replacement has length zero")
        if cnt % len(val) != 0:
            raise RError("number of items to replace is not a multiple of replacement length")
        arr = data.reshape(d, order="F")
        arr[np.ix_(*subs)] = recycle(val, cnt).reshape([len(s) for s in subs], order="F")
        data = arr.reshape(-1, order="F")
    return RVec(kind, data, x.attrs)


# ------------------------------------------------------------------------------------------------------------------------------
# the evaluator
# ------------------------------------------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, args=(), seed=1):
        self.base = Env()
        self.globalenv = Env(self.base)
        self.rng = np.random.default_rng(seed)
        self.script_args = list(args)
        self.script_file = None
        self.namespaces = {}                      # package name -> Env (load_namespace); `pkg::name` looks there first
        self.dotcall = None                       # hook for .Call(symbol, ...): callable(symbol_name, [values]) -> value
        self.out = sys.stdout
        from . import builtins as B
        B.install(self)

    # -- entry points
    def run(self, src, env=None):
        env = env or self.globalenv
        val = NULL
        for node in parse(src):
            val = self.ev(node, env)
        return val

    def source(self, path, env=None):
        with open(path) as fh:
            return self.run(fh.read(), env)

    def load_namespace(self, name, *paths):
        """Source files into an environment of their own, reachable as `name::fn` (what library(name) provides)."""
        env = self.namespaces.setdefault(name, Env(self.base))
        for p in paths:
            self.source(p, env)
        return env

    def call(self, fn, *args, **kwargs):
        """Call an R function value from Python with already evaluated arguments."""
        if isinstance(fn, str):
            fn = self.get_function(fn, self.globalenv)
        supplied = [(None, self.const_promise(a)) for a in args] + [(k, self.const_promise(v)) for k, v in kwargs.items()]
        return self.apply_fn(fn, supplied)

    @staticmethod
    def const_promise(value):
        p = Promise(None, None)
        p.value, p.forced = value, True
        return p

    # -- helpers
    def force(self, v):
        if isinstance(v, Promise):
            if not v.forced:
                v.value = self.ev(v.expr, v.env)
                v.forced = True
                v.expr = v.env = None
            return v.value
        if isinstance(v, Missing):
            raise RError("argument is missing, with no default")
        return v

    def get_function(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = e.vars[name]
                if isinstance(v, Promise):
                    v = self.force(v)
                if is_fn(v):
                    return v
            e = e.parent
        raise RError(f"could not find function \"{name}\"")

    # -- evaluation
    def ev(self, node, env):
        k = node[0]
        if k == "num":
            return RVec("integer" if node[2] else "double", [node[1]])
        if k == "id":
            return self.force(env.lookup(node[1]))
        if k == "call":
            return self.ev_call(node, env)
        if k == "assign":
            val = self.ev(node[2], env)
            self.assign(node[1], val, env, node[3])
            return val
        if k == "index":
            obj = self.ev(node[1], env)
            args, named = [], {}
            for name, e in node[2]:
                if name in ("drop", "exact"):
                    named[name] = self.ev(e, env)
                else:
                    args.append(None if e is None else self.ev(e, env))
            return index_vec(obj, args, named, node[3])
        if k == "dollar":
            return self.dollar(self.ev(node[1], env), node[2])
        if k == "str":
            return chr_(node[1])
        if k == "block":
            val = NULL
            for e in node[1]:
                val = self.ev(e, env)
            return val
        if k == "if":
            if as_bool(self.ev(node[1], env), "if"):
                return self.ev(node[2], env)
            return self.ev(node[3], env) if node[3] is not None else NULL
        if k == "for":
            seq = self.ev(node[2], env)
            items = [] if isinstance(seq, RNull) else (list(seq.v) if isinstance(seq, RList) else [RVec(seq.kind, [e]) for e in seq.v])
            for it in items:
                env.set(node[1], it)
                try:
                    self.ev(node[3], env)
                except BreakEx:
                    break
                except NextEx:
                    continue
            return NULL
        if k == "while" or k == "repeat":
            while k == "repeat" or as_bool(self.ev(node[1], env), "while"):
                try:
                    self.ev(node[-1], env)
                except BreakEx:
                    break
                except NextEx:
                    continue
            return NULL
        if k == "function":
            return Closure(node[1], node[2], env, node[3])
        if k == "null":
            return NULL
        if k == "lgl":
            return lgl([1.0 if node[1] else 0.0])
        if k == "na":
            if node[1] == "NA_character_":
                return chr_(["NA"])
            return RVec({"NA": "logical", "NA_real_": "double", "NA_integer_": "integer"}[node[1]], [NA_REAL])
        if k == "break":
            raise BreakEx()
        if k == "next":
            raise NextEx()
        if k == "ns":
            name = f"{node[1]}::{node[2]}"
            if node[1] in self.namespaces and node[2] in self.namespaces[node[1]].vars:
                return self.force(self.namespaces[node[1]].vars[node[2]])
            if self.base.has(name):
                return self.base.lookup(name)
            return self.force(env.lookup(node[2]))
        raise RError(f"cannot evaluate node {k}")

    def dollar(self, obj, name):
        if isinstance(obj, RList):
            names = obj.names()
            if name in names:
                return obj.v[names.index(name)]
            cand = [i for i, n in enumerate(names) if n.startswith(name)]            # `$` matches partially
            return obj.v[cand[0]] if len(cand) == 1 else NULL
        if isinstance(obj, RNull):
            return NULL
        raise RError("$ operator is invalid for atomic vectors")

    SPECIAL = {"&&", "||", "missing", "tryCatch", "quote", "function", "on.exit", "suppressWarnings", "suppressMessages", "test_that",
               "switch", "substitute", "Recall", "library", "require"}

    def ev_call(self, node, env):
        fnode, argnodes = node[1], node[2]
        if fnode[0] == "id":
            name = fnode[1]
            if name in self.SPECIAL and not self._shadowed(name, env):
                return self.special(name, argnodes, env)
            fn = self.get_function(name, env)
        else:
            fn = self.ev(fnode, env)
            if not is_fn(fn):
                raise RError("attempt to apply non-function")
        supplied = []
        for name, e in argnodes:
            if e is not None and e[0] == "id" and e[1] == "...":
                dots = env.lookup("...")
                supplied.extend(dots)
            elif e is None:
                supplied.append((name, MISSING))
            else:
                supplied.append((name, Promise(e, env)))
        return self.apply_fn(fn, supplied)

    def _shadowed(self, name, env):
        e = env
        while e is not None and e is not self.base:
            if name in e.vars and is_fn(e.vars[name]):
                return True
            e = e.parent
        return False

    def apply_fn(self, fn, supplied):
        if isinstance(fn, Builtin):
            args, kwargs = [], {}
            for name, p in supplied:
                v = None if isinstance(p, Missing) else self.force(p)
                if name is None:
                    args.append(v)
                else:
                    kwargs[name] = v
            return fn.fn(self, args, kwargs)
        # closure: exact names, then unique partial names, then position (R-lang 4.3.2)
        fenv = Env(fn.env)
        pnames = [p[0] for p in fn.params]
        bound = {}
        rest = []
        has_dots = "..." in pnames
        for name, p in supplied:
            if name is not None and name in pnames and name != "..." and name not in bound:
                bound[name] = p
            else:
                rest.append((name, p))
        rest2 = []
        for name, p in rest:
            if name is not None:
                before_dots = pnames[:pnames.index("...")] if has_dots else pnames
                cand = [q for q in before_dots if q.startswith(name) and q not in bound]
                if len(cand) == 1:
                    bound[cand[0]] = p
                    continue
                if not has_dots:
                    raise RError(f"unused argument ({name})")
            rest2.append((name, p))
        dots = []
        free = [q for q in pnames if q not in bound and q != "..."]
        if has_dots:
            di = pnames.index("...")
            free_before = [q for q in pnames[:di] if q not in bound]
            for name, p in rest2:
                if name is None and free_before:
                    bound[free_before.pop(0)] = p
                else:
                    dots.append((name, p))
        else:
            for name, p in rest2:
                if not free:
                    raise RError("unused argument")
                bound[free.pop(0)] = p
        for pname, default in fn.params:
            if pname == "...":
                fenv.vars["..."] = dots
            elif pname in bound and not isinstance(bound[pname], Missing):
                fenv.vars[pname] = bound[pname]
            elif default is not None:
                fenv.vars[pname] = Promise(default, fenv, is_default=True)
            else:
                fenv.vars[pname] = MISSING
        try:
            return self.ev(fn.body, fenv)
        except ReturnEx as r:
            return r.value

    def special(self, name, argnodes, env):
        if name == "&&" or name == "||":
            a = as_bool(self.ev(argnodes[0][1], env), name)
            if name == "&&" and not a:
                return lgl([0.0])
            if name == "||" and a:
                return lgl([1.0])
            return lgl([1.0 if as_bool(self.ev(argnodes[1][1], env), name) else 0.0])
        if name == "missing":
            v = env.vars.get(argnodes[0][1][1])
            return lgl([1.0 if (isinstance(v, Missing) or (isinstance(v, Promise) and v.is_default)) else 0.0])
        if name in ("suppressWarnings", "suppressMessages"):
            return self.ev(argnodes[0][1], env)
        if name == "quote":
            return RLang(argnodes[0][1])
        if name in ("substitute", "on.exit", "library", "require"):
            return NULL
        if name == "tryCatch":
            handlers, body = {}, None
            for nm, e in argnodes:
                if nm is None or nm == "expr":
                    body = e
                else:
                    handlers[nm] = e
            try:
                try:
                    return self.ev(body, env) if body is not None else NULL
                except (RError, np.linalg.LinAlgError, ArithmeticError, ValueError) as err:
                    if "error" not in handlers:
                        raise
                    h = self.ev(handlers["error"], env)
                    cond = RList([chr_(str(err))], ["message"], {"class": chr_(["simpleError", "error", "condition"])})
                    return self.call(h, cond)
            finally:
                if "finally" in handlers:
                    self.ev(handlers["finally"], env)
        if name == "test_that":
            desc = sstr(self.ev(argnodes[0][1], env))
            self.tests = getattr(self, "tests", [])
            before = getattr(self, "expectations", 0)
            try:
                self.ev(argnodes[1][1], Env(env))
                self.tests.append((desc, True, "", getattr(self, "expectations", 0) - before))
            except RError as err:
                self.tests.append((desc, False, str(err), getattr(self, "expectations", 0) - before))
            return NULL
        if name == "switch":
            sel = self.ev(argnodes[0][1], env)
            if sel.kind == "character":
                key = sel.v[0]
                alts = argnodes[1:]
                for i, (nm, e) in enumerate(alts):
                    if nm == key:
                        while e is None and i + 1 < len(alts):
                            i += 1
                            e = alts[i][1]
                        return self.ev(e, env)
                for nm, e in alts:
                    if nm is None:
                        return self.ev(e, env)
                return NULL
            k = int(sel.v[0])
            return self.ev(argnodes[k][1], env) if 1 <= k < len(argnodes) else NULL
        raise RError(f"special form {name} is not implemented")

    # -- assignment, including nested replacement targets:  res$l[i, ] <- v ; colnames(X) <- v ; X[i + 1, ] <- v
    def assign(self, target, value, env, superassign=False):
        k = target[0]
        if k == "id" or k == "str":
            if superassign:
                env.set_super(target[1], value)
            else:
                env.set(target[1], value)
            return
        if k == "index":
            old = self.current(target[1], env)
            args, named = [], {}
            for name, e in target[2]:
                if name is not None and name in ("drop", "exact"):
                    named[name] = self.ev(e, env)
                else:
                    args.append(None if e is None else self.ev(e, env))
            self.assign(target[1], assign_index(old, args, named, value, target[3]), env, superassign)
            return
        if k == "dollar":
            old = self.current(target[1], env)
            self.assign(target[1], assign_index(old, [chr_(target[2])], {}, value, True), env, superassign)
            return
        if k == "call" and target[1][0] == "id":
            fname = target[1][1] + "<-"
            inner = target[2][0][1]
            old = self.current(inner, env)
            fn = self.get_function(fname, env)
            extra = [(nm, Promise(e, env)) for nm, e in target[2][1:]]
            new = self.apply_fn(fn, [(None, self.const_promise(old))] + extra + [("value", self.const_promise(value))])
            self.assign(inner, new, env, superassign)
            return
        raise RError("invalid assignment target")

    def current(self, node, env):
        if node[0] == "id":
            return self.force(env.lookup(node[1])) if env.has(node[1]) else NULL
        return self.ev(node, env)
