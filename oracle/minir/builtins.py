"""TEST INFRASTRUCTURE ONLY — the slice of base R (plus `Matrix::expm` and the `testthat` expectations) that the reference's
sample-path code, its tests and the fixture recipe call.  See interp.py for what is restated exactly and what is a stand-in."""
import math
import os
import re
import sys

import numpy as np

from .interp import (NA_REAL, NULL, Builtin, Closure, Env, RError, RLang, RList, RNull, RVec, ReturnEx, arith, as_bool, c_hexfloat, chr_,
                     compare, cumsum_ld, dbl, dimnames_of, fmt_num, from_mat, higher_kind, index_vec, int_, is_fn, lgl,
                     make_dimnames, mat, matprod, recycle, scalar, sint, sstr, strip, sum_ld, KINDS)


def install(I):
    base = I.base

    def reg(name, fn):
        base.set(name, Builtin(name, fn))

    def simple(name):
        def deco(fn):
            reg(name, lambda interp, a, k: fn(*a, **k))
            return fn
        return deco

    base.set("T", lgl([1.0]))
    base.set("F", lgl([0.0]))
    base.set("pi", dbl([math.pi]))
    base.set("LETTERS", chr_([chr(65 + i) for i in range(26)]))
    base.set("letters", chr_([chr(97 + i) for i in range(26)]))

    # ---- operators -------------------------------------------------------------------------------------------------------
    for op in ("+", "-", "*", "/", "^", "%%", "%/%"):
        def f(interp, a, k, op=op):
            if len(a) == 1:
                if op == "-":
                    x = a[0]
                    return RVec("integer" if x.kind in ("integer", "logical") else "double", -x.v, x.attrs)
                if op == "+":
                    return a[0]
            return arith(op, a[0], a[1])
        reg(op, f)
    for op in ("==", "!=", "<", ">", "<=", ">="):
        reg(op, lambda interp, a, k, op=op: compare(op, a[0], a[1]))

    def logic(op):
        def f(interp, a, k):
            x, y = a
            n = 0 if (len(x) == 0 or len(y) == 0) else max(len(x), len(y))
            p, q = recycle(x.v, n) != 0, recycle(y.v, n) != 0
            return lgl((p & q) if op == "&" else (p | q))
        return f
    reg("&", logic("&"))
    reg("|", logic("|"))
    reg("!", lambda interp, a, k: RVec("logical", np.where(np.isnan(a[0].v), NA_REAL, (a[0].v == 0).astype(float)),
                                       {q: v for q, v in a[0].attrs.items() if q in ("dim", "dimnames", "names")}))
    reg("(", lambda interp, a, k: a[0])
    reg("%*%", lambda interp, a, k: matprod(a[0], a[1]))
    reg("%o%", lambda interp, a, k: from_mat(np.multiply.outer(a[0].v, a[1].v)))
    reg("%in%", lambda interp, a, k: lgl([1.0 if any(e == f for f in a[1].v) else 0.0 for e in a[0].v]))

    def colon(interp, a, k):
        lo, hi = scalar(a[0]), scalar(a[1])
        n = int(math.floor(abs(hi - lo) + 1e-10)) + 1
        step = 1.0 if hi >= lo else -1.0
        v = lo + step * np.arange(n)
        return RVec("integer" if lo == int(lo) else "double", v)
    reg(":", colon)

    def do_return(interp, a, k):
        raise ReturnEx(a[0] if a else NULL)
    reg("return", do_return)
    reg("invisible", lambda interp, a, k: a[0] if a else NULL)
    reg("identity", lambda interp, a, k: a[0])
    reg("force", lambda interp, a, k: a[0])

    # ---- construction ------------------------------------------------------------------------------------------------------
    def combine(interp, a, k):
        items = [(None, x) for x in a] + list(k.items())
        if any(isinstance(x, RList) or is_fn(x) for _, x in items):
            out, names = [], []
            for nm, x in items:
                if isinstance(x, RList):
                    out.extend(x.v)
                    names.extend(x.names())
                elif isinstance(x, RVec):
                    for e in x.v:
                        out.append(RVec(x.kind, [e]))
                        names.append(nm or "")
                elif not isinstance(x, RNull):
                    out.append(x)
                    names.append(nm or "")
            return RList(out, names)
        vecs = [(nm, x) for nm, x in items if isinstance(x, RVec)]
        if not vecs:
            return NULL
        kind = "logical"
        for _, x in vecs:
            kind = higher_kind(kind, x.kind)
        data, names, any_names = [], [], False
        for nm, x in vecs:
            xn = x.attrs.get("names")
            for j, e in enumerate(x.v):
                data.append(fmt_num(e, x.kind) if (kind == "character" and x.kind != "character") else e)
                if xn is not None and xn.v[j]:
                    names.append((nm + "." if nm else "") + xn.v[j])
                    any_names = True
                elif nm:
                    names.append(nm if len(x) == 1 else f"{nm}{j + 1}")
                    any_names = True
                else:
                    names.append("")
        r = RVec(kind, data)
        if any_names:
            r.attrs["names"] = chr_(names)
        return r
    reg("c", combine)

    def mklist(interp, a, k):
        return RList(list(a) + list(k.values()), [""] * len(a) + list(k.keys()))
    reg("list", mklist)

    def data_frame(interp, a, k):
        k = {q: v for q, v in k.items() if q not in ("stringsAsFactors", "check.names")}
        cols, names = [], []
        for i, x in enumerate(a):
            if isinstance(x, RList):                             # data.frame(<list>): one column per element
                cols.extend(x.v)
                names.extend(n or f"V{j + 1}" for j, n in enumerate(x.names()))
            else:
                cols.append(x)
                names.append(f"V{i + 1}")
        r = RList(cols + list(k.values()), names + list(k.keys()))
        r.attrs["class"] = chr_("data.frame")
        return r
    reg("data.frame", data_frame)

    def numeric(interp, a, k):
        n = sint(a[0]) if a else sint(k.get("length", dbl([0])))
        return dbl(np.zeros(n))
    reg("numeric", numeric)
    reg("double", numeric)
    reg("integer", lambda interp, a, k: int_(np.zeros(sint(a[0]) if a else 0)))
    reg("logical", lambda interp, a, k: lgl(np.zeros(sint(a[0]) if a else 0)))
    reg("character", lambda interp, a, k: chr_([""] * (sint(a[0]) if a else 0)))

    def vector(interp, a, k):
        mode = sstr(a[0]) if a else sstr(k.get("mode", chr_("logical")))
        n = sint(a[1]) if len(a) > 1 else sint(k.get("length", dbl([0])))
        if mode == "list":
            return RList([NULL] * n)
        return RVec({"numeric": "double"}.get(mode, mode), np.zeros(n) if mode != "character" else [""] * n)
    reg("vector", vector)

    def rep(interp, a, k):
        x = a[0]
        if "each" in k:
            v = np.repeat(x.v, sint(k["each"]))
            x = RVec(x.kind, v)
        times = a[1] if len(a) > 1 else k.get("times")
        if times is not None:
            if len(times) == 1:
                v = np.tile(x.v, sint(times))
            else:
                v = np.repeat(x.v, times.v.astype(int))
            x = RVec(x.kind, v)
        lo = k.get("length.out", k.get("length", k.get("len")))
        if lo is not None:
            x = RVec(x.kind, np.resize(x.v, sint(lo)))
        return RVec(x.kind, x.v)
    reg("rep", rep)
    reg("rep_len", lambda interp, a, k: RVec(a[0].kind, np.resize(a[0].v, sint(a[1]))))
    reg("seq_len", lambda interp, a, k: int_(np.arange(1, sint(a[0]) + 1)))
    reg("seq_along", lambda interp, a, k: int_(np.arange(1, len(a[0]) + 1)))

    def seq(interp, a, k):
        # seq.default: from + (0:n)*by with n = floor((to - from)/by + 1e-10); length.out: from + (0:(n-1))*by, by = (to-from)/(n-1)
        names = ["from", "to", "by"]
        vals = dict(zip(names, a))
        for q, v in k.items():
            full = [n for n in ("from", "to", "by", "length.out", "along.with") if n.startswith(q)]
            vals[full[0]] = v
        if "along.with" in vals:
            return int_(np.arange(1, len(vals["along.with"]) + 1))
        if "from" in vals and len(vals) == 1:
            x = vals["from"]
            n = len(x) if len(x) != 1 else sint(x)
            return int_(np.arange(1, n + 1))
        frm = scalar(vals["from"]) if "from" in vals else 1.0
        if "length.out" in vals:
            n = sint(vals["length.out"])
            if "to" in vals and "by" not in vals:
                to = scalar(vals["to"])
                if n == 1:
                    return dbl([frm])
                by = (to - frm) / (n - 1)
                v = frm + np.arange(n) * by
                if n > 2:
                    v[-1] = to                                   # seq.default: c(from + (0:(n-2))*by, to)
                return dbl(v)
            by = scalar(vals["by"]) if "by" in vals else 1.0
            return dbl(frm + np.arange(n) * by)
        to = scalar(vals["to"])
        if "by" not in vals:
            return colon(interp, [dbl([frm]), dbl([to])], {})
        by = scalar(vals["by"])
        n = int(math.floor((to - frm) / by + 1e-10))
        v = frm + np.arange(n + 1) * by
        int_all = all(float(x) == int(x) for x in (frm, by)) and vals["from"].kind == "integer" and vals["by"].kind == "integer"
        return RVec("integer" if int_all else "double", v)
    reg("seq", seq)

    def matrix(interp, a, k):
        data = a[0] if a else k.get("data", RVec("logical", [NA_REAL]))
        pos = list(a[1:])
        nrow = k.get("nrow", pos[0] if len(pos) > 0 else None)
        ncol = k.get("ncol", pos[1] if len(pos) > 1 else None)
        byrow = as_bool(k["byrow"]) if "byrow" in k else (as_bool(pos[2]) if len(pos) > 2 else False)
        n = len(data)
        if nrow is None and ncol is None:
            nr, nc = n, 1
        elif ncol is None:
            nr = sint(nrow)
            nc = int(math.ceil(n / nr)) if nr else 0
        elif nrow is None:
            nc = sint(ncol)
            nr = int(math.ceil(n / nc)) if nc else 0
        else:
            nr, nc = sint(nrow), sint(ncol)
        v = np.resize(data.v, nr * nc) if n else np.full(nr * nc, NA_REAL)
        if byrow:
            v = v.reshape(nr, nc).reshape(-1, order="F")
        r = RVec(data.kind, v, {"dim": int_([nr, nc])})
        if "dimnames" in k and not isinstance(k["dimnames"], RNull):
            r.attrs["dimnames"] = k["dimnames"]
        return r
    reg("matrix", matrix)

    def array(interp, a, k):
        data = a[0] if a else k.get("data", RVec("logical", [NA_REAL]))
        d = a[1] if len(a) > 1 else k.get("dim", int_([len(data)]))
        dims = [int(x) for x in d.v]
        n = int(np.prod(dims)) if dims else 0
        v = np.resize(data.v, n) if len(data) else np.full(n, NA_REAL)
        return RVec(data.kind, v, {"dim": int_(dims)})
    reg("array", array)

    def diag(interp, a, k):
        x = a[0] if a else None
        if x is None:
            n = sint(k["n"] if "n" in k else k["nrow"])
            return from_mat(np.eye(n))
        if x.dim() is not None and len(x.dim()) == 2:
            return dbl(np.diag(mat(x)).copy())
        if len(x) == 1 and len(a) == 1 and "nrow" not in k:
            n = sint(x)
            return from_mat(np.eye(n))
        n = sint(a[1]) if len(a) > 1 else (sint(k["nrow"]) if "nrow" in k else len(x))
        return from_mat(np.diag(np.resize(x.v, n)))
    reg("diag", diag)

    def bind(axis):
        def f(interp, a, k):
            items = [(None, x) for x in a] + [(q, v) for q, v in k.items() if q != "deparse.level"]
            items = [(q, x) for q, x in items if not isinstance(x, RNull)]
            mats, other = [], None
            for _, x in items:
                if x.dim() is not None:
                    other = mat(x).shape[1 - axis] if other is None else other
            if other is None:
                other = max(len(x) for _, x in items)
            names = []
            for q, x in items:
                if x.dim() is not None:
                    m = mat(x)
                    dn = dimnames_of(x)
                    names.extend(dn[axis] if dn and dn[axis] else [""] * m.shape[axis])
                else:
                    v = np.resize(x.v, other)
                    m = v.reshape(1, -1) if axis == 0 else v.reshape(-1, 1)
                    names.append(q or "")
                mats.append(m)
            out = np.concatenate(mats, axis=axis)
            dn = None
            if any(names):
                dn = make_dimnames([names, None] if axis == 0 else [None, names])
            return from_mat(out, dimnames=dn)
        return f
    reg("rbind", bind(0))
    reg("cbind", bind(1))

    # ---- attributes --------------------------------------------------------------------------------------------------------
    reg("length", lambda interp, a, k: int_([len(a[0]) if not isinstance(a[0], RNull) and not is_fn(a[0]) else (1 if is_fn(a[0]) else 0)]))
    reg("dim", lambda interp, a, k: a[0].attrs.get("dim", NULL) if not is_fn(a[0]) else NULL)
    reg("nrow", lambda interp, a, k: int_([a[0].dim()[0]]) if isinstance(a[0], RVec) and a[0].dim() else NULL)
    reg("ncol", lambda interp, a, k: int_([a[0].dim()[1]]) if isinstance(a[0], RVec) and a[0].dim() else NULL)
    reg("NROW", lambda interp, a, k: int_([a[0].dim()[0] if a[0].dim() else len(a[0])]))
    reg("NCOL", lambda interp, a, k: int_([a[0].dim()[1] if a[0].dim() else 1]))

    def names(interp, a, k):
        x = a[0]
        if isinstance(x, RVec) and x.dim() is not None and len(x.dim()) == 1:
            dn = dimnames_of(x)
            return chr_(dn[0]) if dn and dn[0] else NULL
        return x.attrs.get("names", NULL) if not is_fn(x) else NULL
    reg("names", names)

    def set_attr(x, key, value):
        if isinstance(x, RVec):
            r = RVec(x.kind, x.v, x.attrs)
        elif isinstance(x, RList):
            r = RList(x.v, None, x.attrs)
        else:
            raise RError("cannot set attributes on this object")
        if isinstance(value, RNull):
            r.attrs.pop(key, None)
        else:
            r.attrs[key] = value
        return r

    def set_names(interp, a, k):
        x, value = a[0], k.get("value", a[1] if len(a) > 1 else NULL)
        if isinstance(value, RVec) and value.kind != "character":
            value = chr_([fmt_num(e, value.kind) for e in value.v])
        return set_attr(x, "names", value)
    reg("names<-", set_names)
    reg("setNames", lambda interp, a, k: set_names(interp, [a[0]], {"value": a[1] if len(a) > 1 else k.get("nm", NULL)}))

    def set_dim(interp, a, k):
        x, value = a[0], k.get("value", a[1] if len(a) > 1 else NULL)
        r = set_attr(x, "dim", value if isinstance(value, RNull) else int_(value.v))
        r.attrs.pop("names", None)
        r.attrs.pop("dimnames", None)
        return r
    reg("dim<-", set_dim)

    def dimnames_get(which):
        def f(interp, a, k):
            dn = dimnames_of(a[0]) if isinstance(a[0], RVec) else None
            if isinstance(a[0], RList) and which == 1:
                return a[0].attrs.get("names", NULL)
            if which is None:
                return a[0].attrs.get("dimnames", NULL)
            return chr_(dn[which]) if dn and len(dn) > which and dn[which] else NULL
        return f
    reg("dimnames", dimnames_get(None))
    reg("rownames", dimnames_get(0))
    reg("colnames", dimnames_get(1))

    def dimnames_set(which):
        def f(interp, a, k):
            x, value = a[0], k.get("value", a[1] if len(a) > 1 else NULL)
            if which is None:
                return set_attr(x, "dimnames", value)
            d = x.dim()
            if d is None or len(d) < 2:
                raise RError("attempt to set 'colnames' on an object with less than two dimensions")
            dn = dimnames_of(x) or [None] * len(d)
            if isinstance(value, RNull):
                dn[which] = None
            else:
                if len(value) != d[which]:
                    raise RError("length of 'dimnames' not equal to array extent")
                dn[which] = [e if isinstance(e, str) else fmt_num(e) for e in value.v]
            r = RVec(x.kind, x.v, x.attrs)
            m = make_dimnames(dn)
            if m is None:
                r.attrs.pop("dimnames", None)
            else:
                r.attrs["dimnames"] = m
            return r
        return f
    reg("dimnames<-", dimnames_set(None))
    reg("rownames<-", dimnames_set(0))
    reg("colnames<-", dimnames_set(1))

    def attr_get(interp, a, k):
        return a[0].attrs.get(sstr(a[1]), NULL)
    reg("attr", attr_get)
    reg("attr<-", lambda interp, a, k: set_attr_any(a[0], sstr(a[1]), k.get("value", a[2] if len(a) > 2 else NULL)))

    def copy_fn(x):
        y = Closure(x.params, x.body, x.env, x.src) if isinstance(x, Closure) else Builtin(x.name, x.fn)
        y.attrs = dict(x.attrs)
        return y

    def set_attr_any(x, key, value):
        if is_fn(x):                                             # attributes on a function: copy, like every R value
            y = copy_fn(x)
            if isinstance(value, RNull):
                y.attrs.pop(key, None)
            else:
                y.attrs[key] = value
            return y
        return set_attr(x, key, value)

    def attributes(interp, a, k):
        at = a[0].attrs
        return RList(list(at.values()), list(at.keys())) if at else NULL
    reg("attributes", attributes)

    def set_attributes(interp, a, k):
        x, value = a[0], k.get("value", a[1] if len(a) > 1 else NULL)
        new = {} if isinstance(value, RNull) else dict(zip(value.names(), value.v))
        if is_fn(x):
            y = copy_fn(x)
            y.attrs = new
            return y
        y = RVec(x.kind, x.v) if isinstance(x, RVec) else RList(x.v)
        y.attrs = new
        return y
    reg("attributes<-", set_attributes)

    def set_storage_mode(interp, a, k):
        x, value = a[0], sstr(k.get("value", a[1] if len(a) > 1 else NULL))
        return RVec({"numeric": "double"}.get(value, value), x.v, x.attrs)
    reg("storage.mode<-", set_storage_mode)
    reg("body", lambda interp, a, k: RLang(a[0].body) if isinstance(a[0], Closure) else NULL)

    def rmatch(interp, a, k):
        table = list(a[1].v)
        return int_([table.index(e) + 1 if e in table else NA_REAL for e in a[0].v])
    reg("match", rmatch)

    def dotcall(interp, a, k):
        if interp.dotcall is None:
            raise RError(".Call: no native library is attached to this evaluator")
        return interp.dotcall(sstr(a[0]), list(a[1:]))
    reg(".Call", dotcall)

    reg("class", lambda interp, a, k: a[0].attrs.get("class", chr_("function" if is_fn(a[0]) else ("list" if isinstance(a[0], RList)
                else ("NULL" if isinstance(a[0], RNull) else ("matrix" if a[0].dim() and len(a[0].dim()) == 2 else
                {"double": "numeric"}.get(a[0].kind, a[0].kind)))))))
    reg("class<-", lambda interp, a, k: set_attr(a[0], "class", k.get("value", a[1] if len(a) > 1 else NULL)))

    def formals(interp, a, k):
        f = a[0]
        if isinstance(f, Closure):
            return RList([NULL for _ in f.params], [p[0] for p in f.params])
        return NULL                                              # primitives have no formals
    reg("formals", formals)

    # ---- type tests and conversion -----------------------------------------------------------------------------------------
    reg("is.null", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RNull) else 0.0]))
    reg("is.function", lambda interp, a, k: lgl([1.0 if is_fn(a[0]) else 0.0]))
    reg("is.matrix", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].dim() is not None and len(a[0].dim()) == 2 else 0.0]))
    reg("is.array", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].dim() is not None else 0.0]))
    reg("is.numeric", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].kind in ("double", "integer") else 0.0]))
    reg("is.character", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].kind == "character" else 0.0]))
    reg("is.logical", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].kind == "logical" else 0.0]))
    reg("is.list", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RList) else 0.0]))
    reg("is.na", lambda interp, a, k: lgl([1.0 if (e != e if not isinstance(e, str) else e == "NA") else 0.0 for e in a[0].v],
                                          **({"dim": a[0].attrs["dim"]} if "dim" in a[0].attrs else {})))
    reg("is.nan", lambda interp, a, k: lgl([1.0 if (e != e and not (isinstance(e, float) and False)) else 0.0 for e in a[0].v]))
    reg("is.finite", lambda interp, a, k: lgl(np.isfinite(a[0].v).astype(float)))

    def parse_num(s):
        s = s.strip()
        if s in ("NA", ""):
            return NA_REAL
        try:
            if re.match(r"^[+-]?0[xX]", s):
                return float.fromhex(s)
            if s in ("Inf", "-Inf", "NaN", "inf", "-inf", "nan"):
                return float(s)
            return float(s)
        except ValueError:
            return NA_REAL

    def as_numeric(interp, a, k):
        x = a[0]
        if isinstance(x, RNull):
            return dbl([])
        if isinstance(x, RList):
            return dbl([scalar(e) for e in x.v])
        if x.kind == "character":
            return dbl([parse_num(s) for s in x.v])
        return RVec("double", x.v)
    reg("as.numeric", as_numeric)
    reg("as.double", as_numeric)
    reg("as.vector", lambda interp, a, k: a[0] if isinstance(a[0], RList) else RVec(a[0].kind, a[0].v))

    def as_integer(interp, a, k):
        x = as_numeric(interp, a, k)
        with np.errstate(all="ignore"):
            return int_(np.trunc(x.v))
    reg("as.integer", as_integer)
    reg("as.logical", lambda interp, a, k: lgl(np.where(np.isnan(a[0].v), NA_REAL, (a[0].v != 0).astype(float))))
    reg("as.character", lambda interp, a, k: a[0] if a[0].kind == "character" else chr_([fmt_num(e, a[0].kind) for e in a[0].v]))

    def as_matrix(interp, a, k):
        x = a[0]
        if isinstance(x, RVec):
            if x.dim() is not None and len(x.dim()) == 2:
                return x
            r = RVec(x.kind, x.v, {"dim": int_([len(x), 1])})
            if "names" in x.attrs:
                r.attrs["dimnames"] = RList([x.attrs["names"], NULL])
            return r
        raise RError("as.matrix: unsupported object")
    reg("as.matrix", as_matrix)
    reg("as.list", lambda interp, a, k: a[0] if isinstance(a[0], RList) else RList([RVec(a[0].kind, [e]) for e in a[0].v],
                                                                                    list(a[0].attrs["names"].v) if "names" in a[0].attrs else None))
    reg("unlist", lambda interp, a, k: combine(interp, list(a[0].v), {}) if isinstance(a[0], RList) else a[0])

    # ---- mathematics -------------------------------------------------------------------------------------------------------
    def math1(name, f):
        def g(interp, a, k):
            x = a[0]
            with np.errstate(all="ignore"):
                return RVec("double", f(x.v), x.attrs)
        reg(name, g)
    for nm, f in (("sqrt", np.sqrt), ("exp", np.exp), ("sin", np.sin), ("cos", np.cos), ("tan", np.tan), ("expm1", np.expm1),
                  ("log1p", np.log1p), ("floor", np.floor), ("ceiling", np.ceil), ("trunc", np.trunc), ("tanh", np.tanh),
                  ("sinh", np.sinh), ("cosh", np.cosh), ("atan", np.arctan), ("asin", np.arcsin), ("acos", np.arccos),
                  ("log2", np.log2), ("log10", np.log10), ("sign", np.sign),
                  ("gamma", np.vectorize(math.gamma, otypes=[float])), ("lgamma", np.vectorize(math.lgamma, otypes=[float]))):
        math1(nm, f)
    reg("abs", lambda interp, a, k: RVec(a[0].kind if a[0].kind != "logical" else "integer", np.abs(a[0].v), a[0].attrs))

    def rlog(interp, a, k):
        x = a[0]
        with np.errstate(all="ignore"):
            r = np.log(x.v)
            base_ = a[1] if len(a) > 1 else k.get("base")
            if base_ is not None:
                r = r / np.log(scalar(base_))
        return RVec("double", r, x.attrs)
    reg("log", rlog)
    reg("round", lambda interp, a, k: RVec(a[0].kind, np.round(a[0].v, sint(a[1]) if len(a) > 1 else (sint(k["digits"]) if "digits" in k else 0)), a[0].attrs))
    reg("signif", lambda interp, a, k: RVec("double", [float("%.*g" % (sint(a[1]) if len(a) > 1 else 6, e)) for e in a[0].v], a[0].attrs))

    def pextreme(f):
        def g(interp, a, k):
            n = max(len(x) for x in a)
            out = recycle(a[0].v, n).copy()
            for x in a[1:]:
                out = f(out, recycle(x.v, n))
            return RVec(a[0].kind if all(x.kind == a[0].kind for x in a) else "double", out, a[0].attrs if len(a[0]) == n else {})
        return g
    reg("pmax", pextreme(np.maximum))
    reg("pmin", pextreme(np.minimum))

    def flat(a):
        return np.concatenate([x.v for x in a if isinstance(x, RVec)]) if a else np.zeros(0)

    def na_rm(a, k):
        v = flat(a)
        if "na.rm" in k and as_bool(k["na.rm"]):
            v = v[~np.isnan(v)]
        return v
    reg("sum", lambda interp, a, k: RVec("integer" if all(x.kind in ("integer", "logical") for x in a) else "double", [sum_ld(na_rm(a, k))]))
    reg("prod", lambda interp, a, k: dbl([float(np.prod(na_rm(a, k).astype(np.longdouble)))]))
    reg("max", lambda interp, a, k: RVec(a[0].kind, [np.max(na_rm(a, k))]) if len(na_rm(a, k)) else dbl([-math.inf]))
    reg("min", lambda interp, a, k: RVec(a[0].kind, [np.min(na_rm(a, k))]) if len(na_rm(a, k)) else dbl([math.inf]))
    reg("range", lambda interp, a, k: dbl([np.min(na_rm(a, k)), np.max(na_rm(a, k))]))

    def mean(interp, a, k):
        v = na_rm(a[:1], k)
        n = len(v)
        if n == 0:
            return dbl([math.nan])
        s = np.longdouble(sum_ld(v)) / n                          # summary.c: a second pass refines the long double mean
        t = np.longdouble(0)
        for e in v:
            t += np.longdouble(e) - s
        return dbl([float(np.float64(s + t / n))])
    reg("mean", mean)

    def var(interp, a, k):
        v = a[0].v
        return dbl([float(np.var(v, ddof=1))])
    reg("var", var)
    reg("sd", lambda interp, a, k: dbl([float(np.std(a[0].v, ddof=1))]))
    reg("all", lambda interp, a, k: lgl([1.0 if np.all(flat(a) != 0) else 0.0]))
    reg("any", lambda interp, a, k: lgl([1.0 if np.any(flat(a) != 0) else 0.0]))
    reg("which", lambda interp, a, k: int_(np.nonzero(a[0].v == 1)[0] + 1))
    reg("which.max", lambda interp, a, k: int_([int(np.argmax(a[0].v)) + 1]))
    reg("which.min", lambda interp, a, k: int_([int(np.argmin(a[0].v)) + 1]))
    reg("rev", lambda interp, a, k: RVec(a[0].kind, a[0].v[::-1].copy()))
    reg("sort", lambda interp, a, k: RVec(a[0].kind, np.sort(a[0].v)))
    reg("order", lambda interp, a, k: int_(np.argsort(a[0].v, kind="stable") + 1))
    reg("cumsum", lambda interp, a, k: RVec("integer" if a[0].kind in ("integer", "logical") else "double", cumsum_ld(a[0].v),
                                            {"names": a[0].attrs["names"]} if "names" in a[0].attrs else None))
    reg("cumprod", lambda interp, a, k: dbl(np.array(np.cumprod(a[0].v.astype(np.longdouble)), dtype=np.float64)))

    def diff(interp, a, k):
        x = a[0]
        lag = sint(k["lag"]) if "lag" in k else (sint(a[1]) if len(a) > 1 else 1)
        d = x.dim()
        if d is not None and len(d) == 2:                         # diff.default on a matrix: r[i, ] = x[i + lag, ] - x[i, ]
            m = mat(x)
            return from_mat(m[lag:, :] - m[:-lag, :], kind=x.kind if x.kind == "integer" else "double")
        v = x.v
        if len(v) <= lag:
            return RVec(x.kind, [])
        return RVec("integer" if x.kind in ("integer", "logical") else "double", v[lag:] - v[:-lag])
    reg("diff", diff)

    def head_tail(which):
        def f(interp, a, k):
            x = a[0]
            n = sint(a[1]) if len(a) > 1 else (sint(k["n"]) if "n" in k else 6)
            d = x.dim() if isinstance(x, RVec) else None
            if d is not None and len(d) == 2:
                m = mat(x)
                nn = n if n >= 0 else max(d[0] + n, 0)
                sub = m[:nn, :] if which == "head" else m[d[0] - min(nn, d[0]):, :]
                return from_mat(sub, kind=x.kind)
            L = len(x)
            nn = min(n, L) if n >= 0 else max(L + n, 0)
            idx = np.arange(nn) if which == "head" else np.arange(L - nn, L)
            if isinstance(x, RList):
                nm = x.names()
                return RList([x.v[i] for i in idx], [nm[i] for i in idx])
            return RVec(x.kind, x.v[idx])
        return f
    reg("head", head_tail("head"))
    reg("tail", head_tail("tail"))

    # ---- linear algebra ----------------------------------------------------------------------------------------------------
    def transpose(interp, a, k):
        x = a[0]
        d = x.dim()
        if d is None:
            r = RVec(x.kind, x.v, {"dim": int_([1, len(x)])})
            return r
        dn = dimnames_of(x)
        return from_mat(mat(x).T, kind=x.kind, dimnames=make_dimnames([dn[1], dn[0]]) if dn else None)
    reg("t", transpose)

    def solve(interp, a, k):
        A = mat(a[0])
        if A.shape[0] != A.shape[1]:
            raise RError("'a' must be a square matrix")
        # R's solve.default: LAPACK dgesv, error "system is exactly singular" when a pivot is exactly zero, error when
        # rcond < tol = .Machine$double.eps ("system is computationally singular")
        if len(a) > 1:
            b = a[1]
            B = mat(b) if b.dim() is not None else b.v.reshape(-1, 1)
        else:
            B = np.eye(A.shape[0])
        if np.any(~np.isfinite(A)):
            raise RError("NA/NaN/Inf in foreign function call")
        try:
            X = np.linalg.solve(A, B)
        except np.linalg.LinAlgError:
            raise RError("Lapack routine dgesv: system is exactly singular")
        rc = 1.0 / np.linalg.cond(A, 1) if A.size else 1.0
        if rc < np.finfo(float).eps:
            raise RError(f"system is computationally singular: reciprocal condition number = {rc:g}")
        if len(a) > 1 and a[1].dim() is None:
            return dbl(X.reshape(-1))
        return from_mat(X)
    reg("solve", solve)
    reg("kronecker", lambda interp, a, k: from_mat(np.kron(mat(a[0]), mat(a[1]))))
    reg("crossprod", lambda interp, a, k: matprod(transpose(interp, [a[0]], {}), a[1] if len(a) > 1 else a[0]))
    reg("det", lambda interp, a, k: dbl([float(np.linalg.det(mat(a[0])))]))
    reg("outer", lambda interp, a, k: from_mat(np.multiply.outer(a[0].v, a[1].v)))

    def expm(interp, a, k):
        import scipy.linalg
        return from_mat(scipy.linalg.expm(mat(a[0])))
    reg("Matrix::expm", expm)                                   # stand-in: scipy's Pade scaling-and-squaring, not the Matrix package's code
    reg("expm", expm)
    reg("Matrix::Diagonal", lambda interp, a, k: from_mat(np.eye(sint(a[0]) if a else sint(k["n"])) * (scalar(k["x"]) if "x" in k else 1.0)))
    reg("Matrix::sparseMatrix", lambda interp, a, k: from_mat(np.zeros([int(e) for e in k["dims"].v])))

    def eigen(interp, a, k):
        w, v = np.linalg.eig(mat(a[0]))
        if np.all(np.isreal(w)):
            o = np.argsort(-w.real)
            return RList([dbl(w.real[o]), from_mat(v.real[:, o])], ["values", "vectors"])
        raise RError("complex eigenvalues are not supported by minir")
    reg("eigen", eigen)
    reg("chol", lambda interp, a, k: from_mat(np.linalg.cholesky(mat(a[0])).T))

    # ---- apply family ------------------------------------------------------------------------------------------------------
    def simplify(results, names=None):
        if all(isinstance(r, RVec) for r in results) and results:
            lens = {len(r) for r in results}
            if lens == {1}:
                kind = "logical"
                for r in results:
                    kind = higher_kind(kind, r.kind)
                out = RVec(kind, [r.v[0] for r in results])
                if names:
                    out.attrs["names"] = chr_(names)
                return out
            if len(lens) == 1 and lens != {0}:
                n = lens.pop()
                m = np.stack([r.v for r in results], axis=1)
                out = from_mat(m, kind=results[0].kind if results[0].kind != "logical" else "logical")
                out.v = m.reshape(-1, order="F")
                rn = results[0].attrs.get("names")
                if rn is not None or names:
                    out.attrs["dimnames"] = RList([rn if rn is not None else NULL, chr_(names) if names else NULL])
                return out
        return RList(results, names)

    def elements(x):
        if isinstance(x, RList):
            return list(x.v)
        return [RVec(x.kind, [e]) for e in x.v]

    def lapply(interp, a, k):
        x, f = a[0], a[1] if len(a) > 1 else k["FUN"]
        extra = list(a[2:])
        kw = {q: v for q, v in k.items() if q not in ("FUN", "simplify", "USE.NAMES")}
        return [interp.call(f, e, *extra, **kw) for e in elements(x)]

    def sapply(interp, a, k):
        res = lapply(interp, a, k)
        x = a[0]
        names = None
        if isinstance(x, RVec) and x.kind == "character":
            names = list(x.v)
        elif "names" in getattr(x, "attrs", {}):
            names = list(x.attrs["names"].v)
        if "simplify" in k and not as_bool(k["simplify"]):
            return RList(res, names)
        return simplify(res, names)
    reg("sapply", sapply)
    reg("lapply", lambda interp, a, k: RList(lapply(interp, a, k), list(a[0].attrs["names"].v) if "names" in getattr(a[0], "attrs", {}) else None))
    reg("vapply", lambda interp, a, k: sapply(interp, [a[0], a[1]], {}))

    def replicate(interp, a, k):
        raise RError("replicate needs an unevaluated expression; not supported by minir")
    reg("replicate", replicate)

    def apply(interp, a, k):
        x, margin, f = a[0], sint(a[1]), a[2] if len(a) > 2 else k["FUN"]
        extra = list(a[3:])
        d = x.dim()
        if d is None or len(d) != 2:
            raise RError("apply: dim(X) must have a positive length (2-D arrays only in minir)")
        m = mat(x)
        dn = dimnames_of(x) or [None, None]
        res = []
        for j in range(d[margin - 1]):
            sl = m[j, :] if margin == 1 else m[:, j]
            v = RVec(x.kind, sl.copy())
            other = dn[2 - margin]
            if other:
                v.attrs["names"] = chr_(other)
            res.append(interp.call(f, v, *extra))
        return simplify(res, dn[margin - 1])
    reg("apply", apply)
    reg("do.call", lambda interp, a, k: interp.call(a[0] if is_fn(a[0]) else sstr(a[0]), *[e for e, n in zip(a[1].v, a[1].names()) if not n],
                                                    **{n: e for e, n in zip(a[1].v, a[1].names()) if n}))
    reg("Reduce", lambda interp, a, k: __import__("functools").reduce(lambda p, q: interp.call(a[0], p, q), elements(a[1])))
    reg("mapply", lambda interp, a, k: simplify([interp.call(a[0], *els) for els in zip(*[elements(x) for x in a[1:]])]))

    # ---- random numbers (numpy's generator: no stream parity with R, none is claimed) -----------------------------------------
    def rnorm(interp, a, k):
        n = sint(a[0]) if a else sint(k["n"])
        mean = a[1] if len(a) > 1 else k.get("mean", dbl([0.0]))
        sd = a[2] if len(a) > 2 else k.get("sd", dbl([1.0]))
        z = interp.rng.standard_normal(n)
        return dbl(recycle(mean.v, n) + recycle(sd.v, n) * z)      # rnorm(n, mu, sigma) = mu + sigma * norm_rand()
    reg("rnorm", rnorm)

    def runif(interp, a, k):
        n = sint(a[0]) if a else sint(k["n"])
        lo = scalar(a[1] if len(a) > 1 else k.get("min", dbl([0.0])))
        hi = scalar(a[2] if len(a) > 2 else k.get("max", dbl([1.0])))
        return dbl(lo + (hi - lo) * interp.rng.random(n))
    reg("runif", runif)

    def set_seed(interp, a, k):
        interp.rng = np.random.default_rng(sint(a[0]))
        return NULL
    reg("set.seed", set_seed)

    # ---- distributions: scipy stand-ins for R's nmath ---------------------------------------------------------------------------
    def dist(name, kind):
        def f(interp, a, k):
            import scipy.stats as st
            names = {"norm": ["mean", "sd"], "chisq": ["df", "ncp"], "exp": ["rate"], "gamma": ["shape", "rate"]}[name]
            x = a[0]
            par = {}
            for i, nm in enumerate(names):
                if len(a) > 1 + i:
                    par[nm] = a[1 + i].v
                elif nm in k:
                    par[nm] = k[nm].v
            log_ = as_bool(k["log"]) if "log" in k else False
            lower = as_bool(k["lower.tail"]) if "lower.tail" in k else True
            logp = as_bool(k["log.p"]) if "log.p" in k else False
            if name == "norm":
                d = st.norm(loc=par.get("mean", 0.0), scale=par.get("sd", 1.0))
            elif name == "chisq":
                ncp = par.get("ncp")
                d = st.chi2(par["df"]) if ncp is None or np.all(ncp == 0) else st.ncx2(par["df"], ncp)
            elif name == "exp":
                d = st.expon(scale=1.0 / par.get("rate", 1.0))
            else:
                d = st.gamma(par["shape"], scale=1.0 / par.get("rate", 1.0))
            with np.errstate(all="ignore"):
                if kind == "d":
                    r = d.logpdf(x.v) if log_ else d.pdf(x.v)
                elif kind == "p":
                    r = (d.logcdf(x.v) if logp else d.cdf(x.v)) if lower else (d.logsf(x.v) if logp else d.sf(x.v))
                elif kind == "q":
                    p = np.exp(x.v) if logp else x.v
                    r = d.ppf(p) if lower else d.isf(p)
                else:
                    r = d.rvs(size=sint(x), random_state=interp.rng)
            return dbl(np.atleast_1d(r), **({"dim": x.attrs["dim"]} if "dim" in x.attrs and kind != "r" else {}))
        return f
    for nm in ("norm", "chisq", "exp", "gamma"):
        for kd in "dpqr":
            if not (nm == "norm" and kd == "r"):
                reg(kd + nm, dist(nm, kd))

    # ---- strings, files, the script environment -----------------------------------------------------------------------------
    def paste_impl(a, k, default_sep):
        sep = sstr(k["sep"]) if "sep" in k else default_sep
        vecs = [x for x in a if not isinstance(x, RNull) and len(x) > 0]
        if not vecs:
            return chr_([])
        n = max(len(x) for x in vecs)
        cols = [[e if isinstance(e, str) else fmt_num(e, x.kind) for e in recycle(x.v, n)] for x in vecs]
        rows = [sep.join(parts) for parts in zip(*cols)]
        if "collapse" in k and not isinstance(k["collapse"], RNull):
            return chr_(sstr(k["collapse"]).join(rows))
        return chr_(rows)
    reg("paste", lambda interp, a, k: paste_impl(a, k, " "))
    reg("paste0", lambda interp, a, k: paste_impl(a, k, ""))

    def sprintf(interp, a, k):
        fmt = a[0]
        vecs = a[1:]
        if any(len(v) == 0 for v in vecs):
            return chr_([])
        n = max([len(fmt)] + [len(v) for v in vecs])
        out = []
        for i in range(n):
            f = recycle(fmt.v, n)[i]
            vals = [(recycle(v.v, n)[i], v.kind) for v in vecs]
            pos = [0]

            def sub(m):
                spec = m.group(0)
                conv = spec[-1]
                if conv == "%":
                    return "%"
                val, kind = vals[pos[0]]
                pos[0] += 1
                if conv == "a":
                    return c_hexfloat(float(val))
                if conv in "di":
                    if val != val:
                        return "NA"
                    if val != int(val):
                        raise RError("invalid format '%d'; use format %f, %e, %g or %a for numeric objects")
                    return (spec[:-1] + "d") % int(val)
                if conv == "s":
                    return (spec) % (val if isinstance(val, str) else fmt_num(val, kind))
                if val != val:
                    return "NA" if conv != "s" else "NA"
                return spec % float(val)
            out.append(re.sub(r"%[-+ 0#]*\d*(?:\.\d+)?[disfgeEGax%]", sub, f))
        return chr_(out)
    reg("sprintf", sprintf)
    reg("format", lambda interp, a, k: chr_([e if isinstance(e, str) else fmt_num(e, a[0].kind) for e in a[0].v]))
    reg("nchar", lambda interp, a, k: int_([len(e) for e in a[0].v]))
    reg("toupper", lambda interp, a, k: chr_([e.upper() for e in a[0].v]))
    reg("tolower", lambda interp, a, k: chr_([e.lower() for e in a[0].v]))
    reg("substr", lambda interp, a, k: chr_([e[sint(a[1]) - 1:sint(a[2])] for e in a[0].v]))

    def rx(pattern, k):
        if "fixed" in k and as_bool(k["fixed"]):
            return re.compile(re.escape(pattern))
        return re.compile(pattern)

    def strsplit(interp, a, k):
        pat = sstr(a[1]) if len(a) > 1 else sstr(k["split"])
        out = []
        for s in a[0].v:
            if pat == "":
                out.append(chr_(list(s)))
            else:
                parts = rx(pat, k).split(s)
                if parts and parts[-1] == "":
                    parts = parts[:-1]
                out.append(chr_(parts))
        return RList(out)
    reg("strsplit", strsplit)
    reg("sub", lambda interp, a, k: chr_([rx(sstr(a[0]), k).sub(sstr(a[1]).replace("\\", "\\\\") if ("fixed" in k) else re.sub(r"\\(\d)", r"\\g<\1>", sstr(a[1])), s, count=1) for s in a[2].v]))
    reg("gsub", lambda interp, a, k: chr_([rx(sstr(a[0]), k).sub(re.sub(r"\\(\d)", r"\\g<\1>", sstr(a[1])), s) for s in a[2].v]))

    def grep(interp, a, k):
        pat, x = rx(sstr(a[0]), k), a[1]
        hits = [i for i, s in enumerate(x.v) if pat.search(s)]
        if "value" in k and as_bool(k["value"]):
            return chr_([x.v[i] for i in hits])
        return int_([i + 1 for i in hits])
    reg("grep", grep)
    reg("grepl", lambda interp, a, k: lgl([1.0 if rx(sstr(a[0]), k).search(s) else 0.0 for s in a[1].v]))
    reg("file.path", lambda interp, a, k: paste_impl(a, {"sep": chr_("/")}, "/"))
    reg("dirname", lambda interp, a, k: chr_([os.path.dirname(s) or "." for s in a[0].v]))
    reg("basename", lambda interp, a, k: chr_([os.path.basename(s) for s in a[0].v]))
    reg("normalizePath", lambda interp, a, k: chr_([os.path.abspath(s) for s in a[0].v]))
    reg("file.exists", lambda interp, a, k: lgl([1.0 if os.path.exists(s) else 0.0 for s in a[0].v]))

    def dir_create(interp, a, k):
        os.makedirs(sstr(a[0]), exist_ok=True)
        return lgl([1.0])
    reg("dir.create", dir_create)

    def list_files(interp, a, k):
        path = sstr(a[0]) if a else sstr(k.get("path", chr_(".")))
        names = sorted(os.listdir(path))
        if "pattern" in k:
            names = [n for n in names if re.search(sstr(k["pattern"]), n)]
        if "full.names" in k and as_bool(k["full.names"]):
            names = [os.path.join(path, n) for n in names]
        return chr_(names)
    reg("list.files", list_files)

    def read_lines(interp, a, k):
        with open(sstr(a[0] if a else k["con"])) as fh:
            return chr_(fh.read().splitlines())
    reg("readLines", read_lines)

    def write_lines(interp, a, k):
        text = a[0] if a else k["text"]
        con = a[1] if len(a) > 1 else k.get("con")
        body = "".join(s + "\n" for s in text.v)
        if con is None:
            interp.out.write(body)
        else:
            with open(sstr(con), "w") as fh:
                fh.write(body)
        return NULL
    reg("writeLines", write_lines)

    def cat(interp, a, k):
        sep = sstr(k["sep"]) if "sep" in k else " "
        parts = []
        for x in a:
            if isinstance(x, RVec):
                parts.extend(e if isinstance(e, str) else fmt_num(e, x.kind) for e in x.v)
        s = sep.join(parts)
        interp.out.write(s.replace(" \n", "\n") if sep == " " else s)
        return NULL
    reg("cat", cat)

    def rprint(interp, a, k):
        x = a[0]
        if isinstance(x, RVec):
            interp.out.write(" ".join(e if isinstance(e, str) else fmt_num(e, x.kind) for e in x.v) + "\n")
        else:
            interp.out.write(repr(x) + "\n")
        return x
    reg("print", rprint)
    reg("message", lambda interp, a, k: (sys.stderr.write("".join(sstr(x) for x in a) + "\n"), NULL)[1])

    def commandArgs(interp, a, k):
        trailing = as_bool(a[0]) if a else (as_bool(k["trailingOnly"]) if "trailingOnly" in k else False)
        if trailing:
            return chr_(interp.script_args)
        return chr_(["minir", "--file=" + (interp.script_file or "")] + (["--args"] + interp.script_args if interp.script_args else []))
    reg("commandArgs", commandArgs)
    reg("Sys.getenv", lambda interp, a, k: chr_(os.environ.get(sstr(a[0]), sstr(a[1]) if len(a) > 1 else sstr(k.get("unset", chr_(""))))))
    reg("Sys.time", lambda interp, a, k: dbl([__import__("time").time()]))
    reg("nargs", lambda interp, a, k: int_([0]))

    def rsource(interp, a, k):
        interp.source(sstr(a[0]))
        return NULL
    reg("source", rsource)

    def rget(interp, a, k):
        return interp.force(interp.globalenv.lookup(sstr(a[0])))
    reg("get", rget)
    reg("match.fun", lambda interp, a, k: a[0] if is_fn(a[0]) else interp.get_function(sstr(a[0]), interp.globalenv))
    reg("exists", lambda interp, a, k: lgl([1.0 if interp.globalenv.has(sstr(a[0])) else 0.0]))
    reg("requireNamespace", lambda interp, a, k: lgl([1.0 if sstr(a[0]) in ("Matrix",) else 0.0]))
    reg("library", lambda interp, a, k: NULL)
    reg("require", lambda interp, a, k: lgl([1.0]))

    def stop(interp, a, k):
        raise RError("".join(sstr(x) for x in a if isinstance(x, RVec)))
    reg("stop", stop)
    reg("warning", lambda interp, a, k: (sys.stderr.write("Warning: " + "".join(sstr(x) for x in a) + "\n"), NULL)[1])

    def stopifnot(interp, a, k):
        for i, x in enumerate(a):
            if not (isinstance(x, RVec) and len(x) > 0 and np.all(x.v == 1)):
                raise RError(f"stopifnot: condition {i + 1} is not all TRUE")
        return NULL
    reg("stopifnot", stopifnot)

    # ---- testthat's expectations (tests/testthat/test_solvers.R) ----------------------------------------------------------------
    def all_equal_num(x, y, tol=1.5e-8):
        # all.equal.numeric: mean relative difference sum|x-y| / sum|x| (absolute when the target's mean is below tol)
        if len(x) != len(y):
            return f"Lengths ({len(x)}, {len(y)}) differ"
        a, b = x.v.astype(float), y.v.astype(float)
        if np.any(np.isnan(a) != np.isnan(b)):
            return "'is.NA' value mismatch"
        ok = ~np.isnan(a)
        a, b = a[ok], b[ok]
        diff_ = np.abs(a - b)
        if not len(a) or np.all(diff_ == 0):
            return None
        xy = np.sum(diff_) / len(a)
        xn = np.sum(np.abs(a)) / len(a)
        what = "relative"
        if np.isfinite(xn) and xn > tol:
            xy = xy / xn
        else:
            what = "absolute"
        return None if xy < tol else f"Mean {what} difference: {xy:g}"

    def expect_equal(interp, a, k):
        interp.expectations = getattr(interp, "expectations", 0) + 1
        x, y = a[0], a[1]
        tol = scalar(k["tolerance"]) if "tolerance" in k else 1.5e-8
        dx, dy = x.attrs.get("dim"), y.attrs.get("dim")
        if (dx is None) != (dy is None) or (dx is not None and list(dx.v) != list(dy.v)):
            raise RError("expect_equal: attributes (dim) differ")
        msg = all_equal_num(x, y, tol)
        if msg:
            raise RError(f"expect_equal: {msg}")
        return NULL
    reg("expect_equal", expect_equal)

    def expect_true(interp, a, k):
        interp.expectations = getattr(interp, "expectations", 0) + 1
        if not as_bool(a[0]):
            raise RError("expect_true: FALSE")
        return NULL
    reg("expect_true", expect_true)
    def same_obj(x, y):
        if isinstance(x, RNull) or isinstance(y, RNull):
            return isinstance(x, RNull) and isinstance(y, RNull)
        if is_fn(x) or is_fn(y):
            if isinstance(x, Builtin) and isinstance(y, Builtin):
                return x.fn is y.fn and same_attrs(x, y)
            return x is y
        if isinstance(x, RLang) or isinstance(y, RLang):
            return isinstance(x, RLang) and isinstance(y, RLang) and x.node == y.node
        if isinstance(x, RList) or isinstance(y, RList):
            return (isinstance(x, RList) and isinstance(y, RList) and len(x) == len(y) and all(same_obj(p, q) for p, q in zip(x.v, y.v))
                    and same_attrs(x, y))
        return (x.kind == y.kind and len(x) == len(y) and all((p == q) or (p != p and q != q) for p, q in zip(x.v, y.v))
                and same_attrs(x, y))

    def same_attrs(x, y):
        return set(x.attrs) == set(y.attrs) and all(same_obj(x.attrs[q], y.attrs[q]) for q in x.attrs)

    def all_equal(interp, a, k):
        x, y = a[0], a[1]
        if isinstance(x, RVec) and isinstance(y, RVec) and x.kind != "character" and y.kind != "character":
            msg = all_equal_num(x, y, scalar(k["tolerance"]) if "tolerance" in k else 1.5e-8)
            return lgl([1.0]) if msg is None else chr_(msg)
        return lgl([1.0]) if same_obj(x, y) else chr_("objects differ")
    reg("all.equal", all_equal)
    reg("identical", lambda interp, a, k: lgl([1.0 if same_obj(a[0], a[1]) else 0.0]))
    reg("isTRUE", lambda interp, a, k: lgl([1.0 if isinstance(a[0], RVec) and a[0].kind == "logical" and len(a[0]) == 1 and a[0].v[0] == 1 else 0.0]))
