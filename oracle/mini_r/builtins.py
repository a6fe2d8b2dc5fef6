"""Base-R functions used by the reference's sparse-GP path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Each builtin receives (interp, positional values, named values). Names follow base R; argument names are the
documented ones (help pages of base / stats), matched exactly or by unique prefix like R does for closures.
"""
from __future__ import annotations

import math
import sys

import numpy as np

from .interp import (MISSING, Builtin, Closure, Lang, RError, RList, Vec, _format_num, _Return, arith, as_float, chrv,
                     dbl, from_matrix, index_get, index_set, is_chr, lgl, matrix_of, scalar, truthy)


def _args(pos, kw, names, defaults=None):
    """Match named (exact, then unique prefix) and positional values against the formal `names`."""
    out = {}
    kw = dict(kw)
    for k in list(kw):
        if k in names:
            out[k] = kw.pop(k)
    for k in list(kw):
        hits = [n for n in names if n.startswith(k) and n not in out]
        if len(hits) == 1:
            out[hits[0]] = kw.pop(k)
    free = [n for n in names if n not in out]
    for n, v in zip(free, pos):
        out[n] = v
    for n, dv in (defaults or {}).items():
        out.setdefault(n, dv)
    for n in names:
        if n not in out:
            raise RError('argument "%s" is missing, with no default' % n)
    if kw:
        out["..."] = kw
    return out


def _num1(f, keep_dim=True):
    def g(I, pos, kw):
        if len(pos) + len(kw) != 1:          # never drop an argument silently (plogis(log.p = TRUE) once was)
            raise RError("unsupported extra arguments %r to a one-argument math builtin" % (list(kw),))
        x = pos[0] if pos else list(kw.values())[0]
        with np.errstate(all="ignore"):
            r = f(as_float(x))
        return Vec(r, dim=x.dim if keep_dim else None, names=x.names)
    return g


def _flatten_c(vals, names_in):
    """c(...)"""
    if any(isinstance(v, (RList, Closure, Builtin)) for v in vals):
        items, names = [], []
        for nm, v in zip(names_in, vals):
            if isinstance(v, RList):
                items += v.items
                names += (v.names or [""] * len(v.items))
            elif isinstance(v, Vec):
                for k in range(len(v.v)):
                    items.append(Vec(v.v[k:k + 1]))
                    names.append((v.names[k] if v.names else "") or (nm or ""))
            elif v is not None:
                items.append(v)
                names.append(nm or "")
        return RList(items, names if any(names) else None)
    vs = [v for v in vals if v is not None]
    if not vs:
        return None
    kinds = [v.v.dtype for v in vs]
    if any(k == object for k in kinds):
        parts = []
        for v in vs:
            parts += [e if v.v.dtype == object else _format_num(e) for e in v.v]
        out = chrv(parts)
    elif any(k == np.float64 for k in kinds):
        out = Vec(np.concatenate([v.v.astype(np.float64) for v in vs]))
    elif any(k == np.int64 for k in kinds):
        out = Vec(np.concatenate([v.v.astype(np.int64) for v in vs]))
    else:
        out = Vec(np.concatenate([v.v for v in vs]))
    names, has = [], False
    for nm, v in zip(names_in, vals):
        if v is None:
            continue
        for k in range(len(v.v)):
            inner = v.names[k] if v.names else ""
            if nm and len(v.v) == 1:
                names.append(nm)
                has = True
            elif nm and inner:
                names.append(nm + "." + inner)
                has = True
            elif nm:
                names.append(nm + str(k + 1))
                has = True
            else:
                names.append(inner)
                has = has or bool(inner)
    out.names = names if has else None
    return out


def install(I):
    G = I.globalenv.vars

    def reg(name, fn, special=False):
        G[name] = Builtin(name, fn, special)

    G["pi"] = dbl(math.pi)
    G["letters"] = chrv(list("abcdefghijklmnopqrstuvwxyz"))

    # ---------------------------------------------------------------- control / language
    def f_return(I, args, env):
        raise _Return(I.eval(args[0][1], env) if args and args[0][1] is not None else None)
    reg("return", f_return, special=True)

    def f_paren(I, pos, kw):
        return pos[0]
    reg("(", f_paren)

    def f_quote(I, args, env):
        return Lang(args[0][1])
    reg("quote", f_quote, special=True)

    def subst(ast, mapping):
        if not isinstance(ast, tuple):
            return ast
        if ast[0] == "sym" and ast[1] in mapping:
            return ("value", mapping[ast[1]])
        if ast[0] == "dollar":
            nm = ast[2]
            if nm in mapping and is_chr(mapping[nm]):
                nm = mapping[nm].v[0]
            return ("dollar", subst(ast[1], mapping), nm)
        if ast[0] == "call":
            return ("call", subst(ast[1], mapping), [(nm, subst(a, mapping) if a is not None else None) for nm, a in ast[2]])
        if ast[0] == "index":
            return ("index", subst(ast[1], mapping), [(nm, subst(a, mapping) if a is not None else None) for nm, a in ast[2]], ast[3])
        if ast[0] in ("binop",):
            return ("binop", ast[1], subst(ast[2], mapping), subst(ast[3], mapping))
        if ast[0] == "unop":
            return ("unop", ast[1], subst(ast[2], mapping))
        if ast[0] == "block":
            return ("block", [subst(a, mapping) for a in ast[1]])
        return ast

    def f_substitute(I, args, env):
        named = {nm: a for nm, a in args if nm}
        pos = [a for nm, a in args if not nm]
        expr = named.get("expr", pos[0] if pos else None)
        envarg = named.get("env", pos[1] if len(pos) > 1 else None)
        mapping = {}
        if envarg is not None:
            lst = I.eval(envarg, env)
            if isinstance(lst, RList):
                mapping = {nm: v for nm, v in zip(lst.names or [], lst.items) if nm}
        else:                      # inside a function: substitute the promise expressions -- not needed here
            mapping = {}
        return Lang(subst(expr, mapping))
    reg("substitute", f_substitute, special=True)

    def f_eval(I, args, env):
        # eval(expr): expr is evaluated in the calling frame (envir = parent.frame()), which is `env` here
        named = {nm: a for nm, a in args if nm}
        pos = [a for nm, a in args if not nm]
        e = I.eval(named.get("expr", pos[0] if pos else None), env)
        if isinstance(e, Lang):
            return I.eval(e.ast, env)
        return e
    reg("eval", f_eval, special=True)

    def f_parse(I, pos, kw):
        # parse(text = "..."): an expression vector; eval() of it evaluates each element and returns the last value
        from . import rparse
        txt = kw.get("text", pos[0] if pos else None)
        nodes = rparse.parse("\n".join(str(s) for s in txt.v))
        return Lang(("block", list(nodes)))
    reg("parse", f_parse)

    def f_missing(I, args, env):
        nm = args[0][1][1]
        e, v = env.lookup(nm)
        return lgl(v is MISSING)
    reg("missing", f_missing, special=True)

    # exists / get / assign on the global environment (where the sourced files and the builtins live); the reference's
    # R code uses none of them -- r/patches.R does, to keep the reference bodies as <name>_R
    def _name_arg(pos, kw):
        return str((pos[0] if pos else kw.get("x", kw.get("name"))).v[0])

    def f_exists(I, pos, kw):
        return lgl(_name_arg(pos, kw) in I.globalenv.vars)
    reg("exists", f_exists)

    def f_get(I, pos, kw):
        nm = _name_arg(pos, kw)
        if nm not in I.globalenv.vars:
            raise RError("object '%s' not found" % nm)
        return I.force(I.globalenv.vars[nm])
    reg("get", f_get)

    def f_assign(I, pos, kw):
        value = pos[1] if len(pos) > 1 else kw["value"]
        I.globalenv.vars[_name_arg(pos, kw)] = value
        return value
    reg("assign", f_assign)
    reg("globalenv", lambda I, pos, kw: I.globalenv)

    def f_stop(I, pos, kw):
        raise RError(" ".join(str(p.v[0]) if isinstance(p, Vec) and len(p.v) else "" for p in pos))
    reg("stop", f_stop)

    def f_try(I, args, env):
        try:
            return I.eval(args[0][1], env)
        except RError as e:
            return TryError(str(e))
    reg("try", f_try, special=True)

    def f_invisible(I, pos, kw):
        return pos[0] if pos else None
    reg("invisible", f_invisible)
    reg("suppressWarnings", f_invisible)

    def f_print(I, pos, kw):
        return pos[0] if pos else None
    reg("print", f_print)
    reg("cat", lambda I, pos, kw: None)
    reg("warning", lambda I, pos, kw: None)
    reg("message", lambda I, pos, kw: None)
    reg("set.seed", lambda I, pos, kw: None)
    reg("Sys.time", lambda I, pos, kw: dbl(0.0))

    def f_class(I, pos, kw):
        x = pos[0]
        if isinstance(x, TryError):
            return chrv(["try-error"])
        if isinstance(x, RList):
            return chrv(["list"])
        if isinstance(x, (Closure, Builtin)):
            return chrv(["function"])
        if x is None:
            return chrv(["NULL"])
        if x.dim is not None:
            return chrv(["matrix", "array"])
        return chrv([{"f": "numeric", "i": "integer", "b": "logical", "O": "character"}[x.v.dtype.kind]])
    reg("class", f_class)

    def f_inherits(I, pos, kw):
        return lgl(isinstance(pos[0], TryError) and pos[1].v[0] == "try-error")
    reg("inherits", f_inherits)

    # ---------------------------------------------------------------- constructors
    reg("c", lambda I, pos, kw: _flatten_c(list(pos) + list(kw.values()), [None] * len(pos) + list(kw.keys())))

    def f_list(I, pos, kw):
        items = list(pos) + list(kw.values())
        names = [""] * len(pos) + list(kw.keys())
        return RList(items, names if any(names) else None)
    reg("list", f_list)

    def f_numeric(I, pos, kw):
        a = _args(pos, kw, ["length"], {"length": None})
        n = 0 if a["length"] is None else int(scalar(a["length"]))
        return Vec(np.zeros(n))
    reg("numeric", f_numeric)
    reg("double", f_numeric)
    reg("integer", lambda I, pos, kw: Vec(np.zeros(int(scalar((pos + list(kw.values()))[0])) if (pos or kw) else 0, dtype=np.int64)))
    reg("logical", lambda I, pos, kw: Vec(np.zeros(int(scalar((pos + list(kw.values()))[0])) if (pos or kw) else 0, dtype=bool)))
    reg("character", lambda I, pos, kw: chrv([""] * (int(scalar((pos + list(kw.values()))[0])) if (pos or kw) else 0)))

    def f_vector(I, pos, kw):
        a = _args(pos, kw, ["mode", "length"], {"mode": chrv(["logical"]), "length": dbl(0)})
        n = int(scalar(a["length"]))
        mode = a["mode"].v[0]
        if mode == "list":
            return RList([None] * n, None)
        if mode == "character":
            return chrv([""] * n)
        return Vec(np.zeros(n, dtype=bool if mode == "logical" else np.float64))
    reg("vector", f_vector)

    def f_rep(I, pos, kw):
        a = _args(pos, kw, ["x", "times", "each", "length.out"], {"times": None, "each": None, "length.out": None})
        x = a["x"]
        if isinstance(x, RList):
            t = int(scalar(a["times"])) if a["times"] is not None else 1
            return RList(x.items * t, (x.names * t) if x.names else None)
        v = x.v
        if a["each"] is not None:
            v = np.repeat(v, int(scalar(a["each"])))
        if a["times"] is not None:
            t = a["times"].v
            v = np.tile(v, int(t[0])) if len(t) == 1 else np.repeat(v, t.astype(np.int64))
        if a["length.out"] is not None:
            n = int(scalar(a["length.out"]))
            v = v[np.arange(n) % len(v)]
        return Vec(v.copy())
    reg("rep", f_rep)

    def f_seq(I, pos, kw):
        a = _args(pos, kw, ["from", "to", "by", "length.out"], {"from": None, "to": None, "by": None, "length.out": None})
        if a["from"] is not None and a["to"] is None and a["by"] is None and a["length.out"] is None:
            x = a["from"]
            n = int(scalar(x)) if len(x.v) == 1 else len(x.v)
            return Vec(np.arange(1, n + 1, dtype=np.int64))
        fr = float(scalar(a["from"])) if a["from"] is not None else 1.0
        if a["length.out"] is not None:
            n = int(scalar(a["length.out"]))
            if a["by"] is not None:
                return Vec(fr + float(scalar(a["by"])) * np.arange(n))
            return Vec(np.linspace(fr, float(scalar(a["to"])), n))
        to = float(scalar(a["to"]))
        by = float(scalar(a["by"])) if a["by"] is not None else (1.0 if to >= fr else -1.0)
        n = int(math.floor((to - fr) / by + 1e-10)) + 1
        r = fr + by * np.arange(max(n, 0))
        if fr == int(fr) and by == int(by):
            return Vec(r.astype(np.int64))
        return Vec(r)
    reg("seq", f_seq)
    reg("seq_len", lambda I, pos, kw: Vec(np.arange(1, int(scalar(pos[0])) + 1, dtype=np.int64)))
    reg("seq_along", lambda I, pos, kw: Vec(np.arange(1, len(pos[0]) + 1, dtype=np.int64)))

    def f_matrix(I, pos, kw):
        a = _args(pos, kw, ["data", "nrow", "ncol", "byrow"], {"data": None, "nrow": None, "ncol": None, "byrow": lgl(False)})
        if a["data"] is None:
            data = np.array([np.nan])          # matrix(): 1 x 1 logical NA
        else:
            data = a["data"].v
        nr = int(scalar(a["nrow"])) if a["nrow"] is not None else None
        nc = int(scalar(a["ncol"])) if a["ncol"] is not None else None
        if nr is None and nc is None:
            nr, nc = len(data), 1
        elif nr is None:
            nr = int(math.ceil(len(data) / nc)) if nc else 0
        elif nc is None:
            nc = int(math.ceil(len(data) / nr)) if nr else 0
        cnt = nr * nc
        flat = data[np.arange(cnt) % len(data)] if len(data) else np.zeros(0)
        if truthy(a["byrow"]):
            flat = flat.reshape((nr, nc)).reshape(-1, order="F")
        return Vec(np.ascontiguousarray(flat), dim=(nr, nc))
    reg("matrix", f_matrix)
    def f_abind(I, pos, kw):
        # abind::abind(a, b, along = 3) as the optimiser loops use it: stack m x d slices into an m x d x K history
        if int(scalar(kw.get("along", dbl(3)))) != 3:
            raise RError("abind: only along = 3 is implemented")
        parts, base = [], None
        for a in pos:
            dm = tuple(a.dim) if a.dim is not None else (len(a.v), 1)
            dm = dm + (1,) * (3 - len(dm))
            if base is None:
                base = dm[:2]
            elif dm[:2] != base:
                raise RError("abind: arg dimensions do not match")
            parts.append((as_float(a), dm[2]))
        return Vec(np.concatenate([p for p, _ in parts]), dim=base + (sum(k for _, k in parts),))
    reg("abind::abind", f_abind)

    # Matrix::Matrix(data = 0, nrow, ncol): the reference only uses it as a zero-initialised container whose column /
    # entries are then assigned and which enters %*% and `*` -- a dense double matrix has the same values
    reg("Matrix::Matrix", lambda I, pos, kw: (lambda r: Vec(r.v.astype(np.float64), dim=r.dim))(f_matrix(I, pos, kw)))

    def f_diag(I, pos, kw):
        a = _args(pos, kw, ["x", "nrow", "ncol"], {"nrow": None, "ncol": None})
        x = a["x"]
        if x.dim is not None:
            M = x.v.reshape(x.dim, order="F")
            return Vec(np.diagonal(M).copy())
        if len(x.v) == 1 and a["nrow"] is None:
            n = int(x.v[0])
            return from_matrix(np.eye(n))
        if a["nrow"] is not None:
            n = int(scalar(a["nrow"]))
            return from_matrix(np.diag(np.resize(as_float(x), n)))
        return from_matrix(np.diag(as_float(x)))
    reg("diag", f_diag)

    def f_diag_set(I, pos, kw):
        x, value = pos[0], kw["value"]
        nr, nc = x.dim
        M = as_float(x).reshape((nr, nc), order="F").copy(order="F")
        k = min(nr, nc)
        M[np.arange(k), np.arange(k)] = np.resize(as_float(value), k)
        return from_matrix(M)
    reg("diag<-", f_diag_set)

    def bind(pos, kw, axis):
        vals = [p for p in list(pos) + list(kw.values()) if p is not None]
        mats = []
        n = max((v.dim[0] if axis == 1 else v.dim[1]) if v.dim is not None else len(v.v) for v in vals)
        for v in vals:
            if v.dim is not None:
                mats.append(as_float(v).reshape(v.dim, order="F"))
            else:
                col = np.resize(as_float(v), n)
                mats.append(col.reshape(-1, 1) if axis == 1 else col.reshape(1, -1))
        return from_matrix(np.concatenate(mats, axis=axis))
    reg("cbind", lambda I, pos, kw: bind(pos, kw, 1))
    reg("rbind", lambda I, pos, kw: bind(pos, kw, 0))

    # ---------------------------------------------------------------- attributes
    reg("length", lambda I, pos, kw: Vec(np.array([0 if pos[0] is None else len(pos[0])], dtype=np.int64)))

    def f_names(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        if x is None or isinstance(x, (Closure, Builtin)) or x.names is None:
            return None
        return chrv(list(x.names))
    reg("names", f_names)

    def f_names_set(I, pos, kw):
        x, value = pos[0], kw["value"]
        nm = None if value is None else [str(s) for s in value.v]
        if isinstance(x, RList):
            return RList(list(x.items), nm)
        return Vec(x.v.copy(), dim=x.dim, names=nm)
    reg("names<-", f_names_set)

    def f_dim(I, pos, kw):
        x = pos[0]
        return None if not isinstance(x, Vec) or x.dim is None else Vec(np.array(x.dim, dtype=np.int64))
    reg("dim", f_dim)
    reg("dim<-", lambda I, pos, kw: Vec(pos[0].v.copy(), dim=None if kw["value"] is None else tuple(int(e) for e in kw["value"].v)))

    def f_nrow(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        return None if not isinstance(x, Vec) or x.dim is None else Vec(np.array([x.dim[0]], dtype=np.int64))

    def f_ncol(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        return None if not isinstance(x, Vec) or x.dim is None else Vec(np.array([x.dim[1]], dtype=np.int64))
    reg("nrow", f_nrow)
    reg("ncol", f_ncol)
    reg("NROW", lambda I, pos, kw: Vec(np.array([pos[0].dim[0] if pos[0].dim else len(pos[0].v)], dtype=np.int64)))
    reg("NCOL", lambda I, pos, kw: Vec(np.array([pos[0].dim[1] if pos[0].dim else 1], dtype=np.int64)))

    reg("[", lambda I, pos, kw: index_get(pos[0], pos[1:], False))
    reg("[[", lambda I, pos, kw: index_get(pos[0], pos[1:], True))
    reg("[<-", lambda I, pos, kw: index_set(pos[0], pos[1:], False, kw["value"]))
    reg("[[<-", lambda I, pos, kw: index_set(pos[0], pos[1:], True, kw["value"]))

    # ---------------------------------------------------------------- predicates / coercion
    reg("is.list", lambda I, pos, kw: lgl(isinstance(pos[0] if pos else list(kw.values())[0], RList)))
    reg("is.function", lambda I, pos, kw: lgl(isinstance(pos[0] if pos else list(kw.values())[0], (Closure, Builtin))))
    reg("is.null", lambda I, pos, kw: lgl((pos[0] if pos else list(kw.values())[0]) is None))
    reg("is.matrix", lambda I, pos, kw: lgl(isinstance(pos[0], Vec) and pos[0].dim is not None))
    reg("is.numeric", lambda I, pos, kw: lgl(isinstance(pos[0], Vec) and pos[0].v.dtype.kind in "fi"))
    reg("is.character", lambda I, pos, kw: lgl(is_chr(pos[0])))
    reg("is.logical", lambda I, pos, kw: lgl(isinstance(pos[0], Vec) and pos[0].v.dtype == bool))
    reg("is.vector", lambda I, pos, kw: lgl(isinstance(pos[0], (Vec, RList)) and getattr(pos[0], "dim", None) is None))

    def f_is_na(I, pos, kw):
        x = pos[0] if pos else list(kw.values())[0]
        if isinstance(x, RList):
            return Vec(np.array([isinstance(i, Vec) and len(i.v) == 1 and i.v.dtype == np.float64 and np.isnan(i.v[0])
                                 for i in x.items], dtype=bool))
        if x is None or isinstance(x, (Closure, Builtin)):
            return lgl(False) if x is not None else Vec(np.zeros(0, dtype=bool))
        if x.v.dtype == np.float64:
            return Vec(np.isnan(x.v), dim=x.dim)
        if x.v.dtype == object:
            return Vec(np.array([e is None for e in x.v], dtype=bool))
        return Vec(np.zeros(len(x.v), dtype=bool), dim=x.dim)
    reg("is.na", f_is_na)
    reg("is.nan", f_is_na)
    reg("is.finite", lambda I, pos, kw: Vec(np.isfinite(as_float(pos[0])), dim=pos[0].dim))
    reg("is.infinite", lambda I, pos, kw: Vec(np.isinf(as_float(pos[0])), dim=pos[0].dim))

    def f_as_numeric(I, pos, kw):
        x = pos[0] if pos else list(kw.values())[0]
        if x is None:
            return Vec(np.zeros(0))
        return Vec(as_float(x).copy())
    reg("as.numeric", f_as_numeric)
    reg("as.double", f_as_numeric)
    reg("as.vector", lambda I, pos, kw: pos[0] if isinstance(pos[0], RList) else Vec(pos[0].v.copy()))
    reg("as.integer", lambda I, pos, kw: Vec(np.trunc(as_float(pos[0])).astype(np.int64)))
    reg("as.logical", lambda I, pos, kw: Vec(as_float(pos[0]) != 0))
    reg("as.character", lambda I, pos, kw: chrv([e if isinstance(e, str) else _format_num(e) for e in pos[0].v]))

    def f_as_list(I, pos, kw):
        x = pos[0] if pos else list(kw.values())[0]
        if isinstance(x, RList):
            return x
        if x is None:
            return RList([], None)
        return RList([Vec(x.v[k:k + 1]) for k in range(len(x.v))], list(x.names) if x.names else None)
    reg("as.list", f_as_list)

    def f_as_matrix(I, pos, kw):
        x = pos[0] if pos else list(kw.values())[0]
        if x.dim is not None:
            return x
        return Vec(x.v.copy(), dim=(len(x.v), 1))
    reg("as.matrix", f_as_matrix)

    def f_unlist(I, pos, kw):
        x = pos[0]
        if not isinstance(x, RList):
            return x
        return _flatten_c(x.items, x.names or [None] * len(x.items))
    reg("unlist", f_unlist)

    # ---------------------------------------------------------------- math
    for nm, f in [("exp", np.exp), ("log", np.log), ("sqrt", np.sqrt), ("abs", np.abs), ("floor", np.floor),
                  ("ceiling", np.ceil), ("sign", np.sign), ("log1p", np.log1p), ("expm1", np.expm1), ("sin", np.sin),
                  ("cos", np.cos), ("tanh", np.tanh), ("log2", np.log2), ("log10", np.log10),
                  ("lgamma", np.vectorize(math.lgamma, otypes=[np.float64])),
                  ("gamma", np.vectorize(math.gamma, otypes=[np.float64])),
                  ("lfactorial", np.vectorize(lambda t: math.lgamma(t + 1.0), otypes=[np.float64])),
                  ("factorial", np.vectorize(lambda t: math.gamma(t + 1.0), otypes=[np.float64]))]:
        reg(nm, _num1(f))
    reg("round", lambda I, pos, kw: Vec(np.round(as_float(pos[0]), int(scalar(pos[1])) if len(pos) > 1 else
                                                 int(scalar(kw["digits"])) if "digits" in kw else 0), dim=pos[0].dim))
    def f_plogis(I, pos, kw):
        a = _args(pos, kw, ["q", "location", "scale", "lower.tail", "log.p"],
                  {"location": dbl(0.0), "scale": dbl(1.0), "lower.tail": lgl(True), "log.p": lgl(False)})
        x = (as_float(a["q"]) - scalar(a["location"])) / scalar(a["scale"])
        if not truthy(a["lower.tail"]):
            x = -x
        with np.errstate(all="ignore"):
            r = -np.logaddexp(0.0, -x) if truthy(a["log.p"]) else 1.0 / (1.0 + np.exp(-x))   # nmath/plogis.c
        return Vec(r, dim=a["q"].dim, names=a["q"].names)
    reg("plogis", f_plogis)

    def _all_values(pos):
        parts = [as_float(p) for p in pos if p is not None]
        return np.concatenate(parts) if parts else np.zeros(0)

    def f_sum(I, pos, kw):
        kw = {k: v for k, v in kw.items() if k != "na.rm"}
        v = _all_values(list(pos) + list(kw.values()))
        if all(isinstance(p, Vec) and p.v.dtype.kind in "ib" for p in pos):
            return Vec(np.array([int(v.sum())], dtype=np.int64))
        return dbl(float(np.sum(v, dtype=np.longdouble)))          # rsum(): LDOUBLE accumulator
    reg("sum", f_sum)
    reg("prod", lambda I, pos, kw: dbl(float(np.prod(_all_values(pos), dtype=np.longdouble))))
    reg("mean", lambda I, pos, kw: dbl(float(np.sum(as_float(pos[0] if pos else kw["x"]), dtype=np.longdouble) /
                                            max(len(pos[0] if pos else kw["x"]), 1))))
    reg("max", lambda I, pos, kw: dbl(np.max(_all_values(pos))) if len(_all_values(pos)) else dbl(-np.inf))
    reg("min", lambda I, pos, kw: dbl(np.min(_all_values(pos))) if len(_all_values(pos)) else dbl(np.inf))
    reg("which.max", lambda I, pos, kw: Vec(np.array([int(np.argmax(as_float(pos[0]))) + 1], dtype=np.int64)))
    reg("which.min", lambda I, pos, kw: Vec(np.array([int(np.argmin(as_float(pos[0]))) + 1], dtype=np.int64)))
    reg("any", lambda I, pos, kw: lgl(bool(np.any(_all_values(pos) != 0))))
    reg("all", lambda I, pos, kw: lgl(bool(np.all(_all_values(pos) != 0))))
    reg("cumsum", lambda I, pos, kw: Vec(np.cumsum(as_float(pos[0]))))
    reg("rev", lambda I, pos, kw: Vec(pos[0].v[::-1].copy()))
    reg("var", lambda I, pos, kw: dbl(float(np.var(as_float(pos[0]), ddof=1))))
    reg("sd", lambda I, pos, kw: dbl(float(np.std(as_float(pos[0]), ddof=1))))

    def f_which(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        return Vec((np.nonzero(x.v != 0)[0] + 1).astype(np.int64))
    reg("which", f_which)

    def f_ifelse(I, pos, kw):
        a = _args(pos, kw, ["test", "yes", "no"])
        t = a["test"].v != 0
        yes, no = a["yes"], a["no"]
        n = len(t)
        if yes.v.dtype == object or no.v.dtype == object:
            out = np.empty(n, dtype=object)
            for k in range(n):
                src = yes if t[k] else no
                out[k] = src.v[k % len(src.v)]
            return Vec(out)
        # R only touches `yes` where test is TRUE and `no` where it is FALSE (a zero-length branch that is never
        # selected is legal: vi_functions.R:811 reads obj_fun_vals[0] in the unused branch at iter == 1)
        out = np.full(n, np.nan)
        if np.any(t):
            out[t] = as_float(yes)[np.arange(n) % len(yes.v)][t]
        if np.any(~t):
            out[~t] = as_float(no)[np.arange(n) % len(no.v)][~t]
        return Vec(out, dim=a["test"].dim)
    reg("ifelse", f_ifelse)

    def f_duplicated(I, pos, kw):
        x = pos[0]
        seen, out = set(), []
        if x.dim is not None:
            M = x.v.reshape(x.dim, order="F")
            for r in M:
                key = tuple(r.tolist())
                out.append(key in seen)
                seen.add(key)
        else:
            for e in x.v.tolist():
                out.append(e in seen)
                seen.add(e)
        return Vec(np.array(out, dtype=bool))
    reg("duplicated", f_duplicated)

    def f_paste(I, pos, kw, default_sep=" "):
        sep = kw.pop("sep").v[0] if "sep" in kw else default_sep
        collapse = kw.pop("collapse", None)
        vals = [p for p in pos if p is not None]
        n = max((len(v.v) for v in vals), default=0)
        out = []
        for k in range(n):
            out.append(sep.join((v.v[k % len(v.v)] if v.v.dtype == object else _format_num(v.v[k % len(v.v)])) for v in vals))
        if collapse is not None:
            return chrv([collapse.v[0].join(out)])
        return chrv(out)
    reg("paste", f_paste)
    reg("paste0", lambda I, pos, kw: f_paste(I, pos, kw, ""))
    reg("nchar", lambda I, pos, kw: Vec(np.array([len(s) for s in pos[0].v], dtype=np.int64)))
    reg("identical", lambda I, pos, kw: lgl(pos[0] is pos[1] if isinstance(pos[0], (Closure, Builtin)) else
                                            isinstance(pos[0], Vec) and isinstance(pos[1], Vec) and
                                            len(pos[0].v) == len(pos[1].v) and bool(np.all(pos[0].v == pos[1].v))))

    # ---------------------------------------------------------------- linear algebra
    def f_t(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        if x.dim is None:
            return Vec(x.v.copy(), dim=(1, len(x.v)))
        M = x.v.reshape(x.dim, order="F")
        return Vec(np.ascontiguousarray(M.T.reshape(-1, order="F")), dim=(x.dim[1], x.dim[0]))
    reg("t", f_t)

    def f_solve(I, pos, kw):
        a = _args(pos, kw, ["a", "b"], {"b": None})
        A = matrix_of(a["a"])
        if A.shape[0] != A.shape[1]:
            raise RError("'a' must be a square matrix")
        try:
            if a["b"] is None:
                return from_matrix(np.linalg.solve(A, np.eye(A.shape[0])))      # solve.default: DGESV on I
            b = a["b"]
            Bm = matrix_of(b)
            if Bm.shape[0] != A.shape[0]:
                raise RError("'b' must be compatible with 'a'")
            X = np.linalg.solve(A, Bm)
        except np.linalg.LinAlgError as e:
            raise RError("Lapack routine dgesv: system is exactly singular") from e
        if b.dim is None:
            return Vec(np.ascontiguousarray(X.reshape(-1, order="F")))
        return from_matrix(X)
    reg("solve", f_solve)

    def f_chol(I, pos, kw):
        x = pos[0] if pos else kw["x"]
        try:
            L = np.linalg.cholesky(matrix_of(x))
        except np.linalg.LinAlgError as e:
            raise RError("the leading minor is not positive definite") from e
        return from_matrix(L.T)
    reg("chol", f_chol)
    reg("chol2inv", lambda I, pos, kw: from_matrix(np.linalg.inv(matrix_of(pos[0]).T @ matrix_of(pos[0]))))

    def f_det(I, pos, kw):
        sgn, logabs = np.linalg.slogdet(matrix_of(pos[0] if pos else kw["x"]))
        with np.errstate(all="ignore"):
            return dbl(float(sgn * np.exp(logabs)))                       # det.default: sign * exp(modulus)
    reg("det", f_det)

    def f_determinant(I, pos, kw):
        sgn, logabs = np.linalg.slogdet(matrix_of(pos[0]))
        return RList([dbl(logabs), Vec(np.array([int(sgn)], dtype=np.int64))], ["modulus", "sign"])
    reg("determinant", f_determinant)
    reg("crossprod", lambda I, pos, kw: from_matrix(matrix_of(pos[0]).T @ matrix_of(pos[1] if len(pos) > 1 else pos[0])))
    reg("tcrossprod", lambda I, pos, kw: from_matrix(matrix_of(pos[0]) @ matrix_of(pos[1] if len(pos) > 1 else pos[0]).T))
    reg("rowSums", lambda I, pos, kw: Vec(np.sum(matrix_of(pos[0]), axis=1, dtype=np.longdouble).astype(np.float64)))
    reg("colSums", lambda I, pos, kw: Vec(np.sum(matrix_of(pos[0]), axis=0, dtype=np.longdouble).astype(np.float64)))
    reg("outer", lambda I, pos, kw: from_matrix(np.outer(as_float(pos[0]), as_float(pos[1]))))

    # ---------------------------------------------------------------- apply family
    def call(f, *vals, **named):
        return I.apply_function(f, [(None, v) for v in vals] + [(k, v) for k, v in named.items()])

    def f_apply(I, pos, kw):
        a = _args(pos, kw, ["X", "MARGIN", "FUN"])
        X, margin, fun = a["X"], int(scalar(a["MARGIN"])), a["FUN"]
        extra = a.get("...", {})
        M = X.v.reshape(X.dim, order="F")
        res = []
        rng = range(X.dim[0]) if margin == 1 else range(X.dim[1])
        for k in rng:
            sl = M[k, :] if margin == 1 else M[:, k]
            res.append(call(fun, Vec(np.ascontiguousarray(sl)), **extra))
        if all(isinstance(r, Vec) and len(r.v) == 1 for r in res):
            return Vec(np.array([r.v[0] for r in res]))
        if all(isinstance(r, Vec) for r in res) and len({len(r.v) for r in res}) == 1:
            return from_matrix(np.stack([as_float(r) for r in res], axis=1))
        return RList(res, None)
    reg("apply", f_apply)

    def f_lapply(I, pos, kw, simplify=False):
        a = _args(pos, kw, ["X", "FUN"])
        X, fun = a["X"], a["FUN"]
        extra = a.get("...", {})
        items = X.items if isinstance(X, RList) else [Vec(X.v[k:k + 1]) for k in range(len(X.v))]
        names = X.names if isinstance(X, RList) else X.names
        res = [call(fun, it, **extra) for it in items]
        if simplify and res and all(isinstance(r, Vec) and len(r.v) == 1 for r in res):
            return Vec(np.array([r.v[0] for r in res]), names=list(names) if names else None)
        if simplify and res and all(isinstance(r, Vec) for r in res) and len({len(r.v) for r in res}) == 1:
            return from_matrix(np.stack([as_float(r) for r in res], axis=1))
        return RList(res, list(names) if names else None)
    reg("lapply", f_lapply)
    reg("sapply", lambda I, pos, kw: f_lapply(I, pos, kw, True))

    def f_do_call(I, pos, kw):
        a = _args(pos, kw, ["what", "args"])
        f = a["what"]
        if is_chr(f):
            f = I.get_fun(f.v[0], I.globalenv)
        lst = a["args"]
        return I.apply_function(f, [((nm or None) if lst.names else None, v)
                                    for nm, v in zip(lst.names or [None] * len(lst.items), lst.items)])
    reg("do.call", f_do_call)

    # ---------------------------------------------------------------- distributions used by the likelihood files
    from math import erf, sqrt as msqrt
    reg("pnorm", _num1(np.vectorize(lambda q: 0.5 * (1.0 + erf(q / msqrt(2.0))), otypes=[np.float64])))
    reg("dnorm", _num1(lambda q: np.exp(-0.5 * q * q) / math.sqrt(2.0 * math.pi)))


class TryError:
    def __init__(self, msg):
        self.msg = msg
