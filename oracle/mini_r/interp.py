"""Evaluator for the R subset parsed by rparse.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Value model (R Language Definition, sections 2-4):
  Vec    atomic vector: 1-D numpy array `v` (float64 "double", int64 "integer", bool "logical", object "character"),
         optional `dim` (column-major, as R stores matrices) and `names`.  NA is NaN for doubles; logical NA promotes
         the vector to double (only `matrix()` and `NA` defaults need it here).
  RList  generic vector with optional names.          None is NULL.
  Closure / Builtin   functions.   Lang   an unevaluated expression (substitute / quote).
Semantics implemented: lexical scoping, lazily evaluated default arguments, `...`, argument matching (exact, unique
partial, positional), replacement functions (`x[i] <- v`, `names(x) <- v`, `x$a[[2]] <- v`), recycling arithmetic,
1-based / negative / logical / character indexing with drop, column-major matrices, `%*%`, and the base functions
the reference's sparse-GP path calls.  Linear algebra goes to LAPACK through NumPy exactly where R goes to LAPACK:
solve -> DGESV, chol -> DPOTRF, det -> LU (sign * exp(log-modulus) like det.default).  sum() accumulates in long
double as R's rsum does.
"""
from __future__ import annotations

import math
import sys

import numpy as np

from . import rparse


class RError(Exception):
    pass


class _Break(Exception):
    pass


class _Next(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class Vec:
    __slots__ = ("v", "dim", "names")

    def __init__(self, v, dim=None, names=None):
        self.v = v
        self.dim = dim
        self.names = names

    def __len__(self):
        return len(self.v)

    def __repr__(self):
        return "Vec(%r, dim=%r, names=%r)" % (self.v, self.dim, self.names)


class RList:
    __slots__ = ("items", "names")

    def __init__(self, items, names=None):
        self.items = items
        self.names = names

    def __len__(self):
        return len(self.items)

    def get(self, name, partial=True):
        if self.names:
            for k, nm in enumerate(self.names):
                if nm == name:
                    return self.items[k]
            if partial:
                hits = [k for k, nm in enumerate(self.names) if nm and nm.startswith(name)]
                if len(hits) == 1:
                    return self.items[hits[0]]
        return None


class Lang:
    def __init__(self, ast):
        self.ast = ast


class Closure:
    def __init__(self, formals, body, env, name="<anonymous>"):
        self.formals, self.body, self.env, self.name = formals, body, env, name


class Builtin:
    def __init__(self, name, fn, special=False):
        self.name, self.fn, self.special = name, fn, special


class Promise:
    __slots__ = ("ast", "env", "value", "done")

    def __init__(self, ast, env):
        self.ast, self.env, self.value, self.done = ast, env, None, False


class Missing:
    pass


MISSING = Missing()


class Env:
    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e, e.vars[name]
            e = e.parent
        return None, None

    def set(self, name, value):
        self.vars[name] = value

    def set_super(self, name, value):
        e = self.parent
        while e is not None:
            if name in e.vars:
                e.vars[name] = value
                return
            if e.parent is None:
                e.vars[name] = value
                return
            e = e.parent


# ------------------------------------------------------------------------------------------------ helpers
def dbl(x):
    return Vec(np.array([x], dtype=np.float64))


def lgl(x):
    return Vec(np.array([bool(x)]))


def chrv(xs):
    a = np.empty(len(xs), dtype=object)
    for i, s in enumerate(xs):
        a[i] = s
    return Vec(a)


def is_chr(x):
    return isinstance(x, Vec) and x.v.dtype == object


def as_float(x):
    if isinstance(x, Vec):
        if x.v.dtype == object:
            return np.array([float(s) for s in x.v], dtype=np.float64)
        return x.v.astype(np.float64, copy=False)
    if x is None:
        return np.zeros(0)
    if isinstance(x, RList):
        return np.concatenate([as_float(i) for i in x.items]) if x.items else np.zeros(0)
    raise RError("cannot coerce %r to numeric" % type(x).__name__)


def scalar(x, what="value"):
    if isinstance(x, Vec) and len(x.v) >= 1:
        return x.v[0]
    raise RError("expected a length-one %s, got %r" % (what, x))


def truthy(x):
    if not isinstance(x, Vec) or len(x.v) == 0:
        raise RError("argument is of length zero / not interpretable as logical")
    v = x.v[0]
    if isinstance(v, (float, np.floating)) and math.isnan(v):
        raise RError("missing value where TRUE/FALSE needed")
    return bool(v)


def matrix_of(x):
    """-> 2-D numpy view (column-major semantics) of a Vec; a plain vector is a column."""
    a = as_float(x) if x.v.dtype != np.float64 else x.v
    if x.dim is not None:
        return a.reshape(x.dim, order="F")
    return a.reshape((len(a), 1), order="F")


def from_matrix(m, names=None):
    m = np.asarray(m, dtype=np.float64)
    return Vec(np.ascontiguousarray(m.reshape(-1, order="F")), dim=(m.shape[0], m.shape[1]))


def _recycle(a, n):
    if len(a) == n:
        return a
    if len(a) == 0:
        return a
    return a[np.arange(n) % len(a)]


def arith(op, x, y):
    if not isinstance(x, Vec) or not isinstance(y, Vec):
        raise RError("non-numeric argument to binary operator %s" % op)
    a, b = x.v, y.v
    if a.dtype == object or b.dtype == object:
        if op in ("==", "!="):
            n = max(len(a), len(b))
            aa, bb = _recycle(a, n), _recycle(b, n)
            r = np.array([str(p) == str(q) for p, q in zip(aa, bb)], dtype=bool)
            return Vec(r if op == "==" else ~r)
        raise RError("non-numeric argument to binary operator %s" % op)
    n = 0 if (len(a) == 0 or len(b) == 0) else max(len(a), len(b))
    a, b = _recycle(a, n), _recycle(b, n)
    both_int = a.dtype.kind in "ib" and b.dtype.kind in "ib"
    if a.dtype == bool:
        a = a.astype(np.int64)
    if b.dtype == bool:
        b = b.astype(np.int64)
    with np.errstate(all="ignore"):
        if op == "+":
            r = a + b
        elif op == "-":
            r = a - b
        elif op == "*":
            r = a * b
        elif op == "/":
            r = a.astype(np.float64) / b.astype(np.float64)
        elif op == "^":
            r = np.power(a.astype(np.float64), b.astype(np.float64))
        elif op == "%%":
            r = np.mod(a, b)
        elif op == "%/%":
            r = np.floor_divide(a, b)
        elif op == "==":
            r = a == b
        elif op == "!=":
            r = a != b
        elif op == "<":
            r = a < b
        elif op == ">":
            r = a > b
        elif op == "<=":
            r = a <= b
        elif op == ">=":
            r = a >= b
        elif op in ("&", "|"):
            r = (a != 0) & (b != 0) if op == "&" else (a != 0) | (b != 0)
        else:
            raise RError("unknown operator %s" % op)
    if both_int and op in ("+", "-", "*", "%%", "%/%") and r.dtype.kind != "i":
        r = r.astype(np.int64)
    dim = x.dim if x.dim is not None else y.dim
    if x.dim is not None and y.dim is not None and tuple(x.dim) != tuple(y.dim):
        raise RError("non-conformable arrays")
    if dim is not None and int(np.prod(dim)) != len(r):
        dim = None
    names = x.names if (x.names is not None and len(x.names) == len(r)) else (
        y.names if (y.names is not None and len(y.names) == len(r)) else None)
    return Vec(r, dim=dim, names=names)


def matmul(x, y):
    if not isinstance(x, Vec) or not isinstance(y, Vec):
        raise RError("requires numeric/complex matrix/vector arguments")
    if x.dim is None and y.dim is None:
        a, b = as_float(x), as_float(y)
        if len(a) == len(b):
            return from_matrix(np.array([[a @ b]]))
        if len(a) == 1:
            return from_matrix(a.reshape(1, 1) @ b.reshape(1, -1))
        if len(b) == 1:
            return from_matrix(a.reshape(-1, 1) @ b.reshape(1, 1))
        raise RError("non-conformable arguments")
    if x.dim is None:
        a, B = as_float(x), matrix_of(y)
        A = a.reshape(1, -1) if len(a) == B.shape[0] else a.reshape(-1, 1)
        if A.shape[1] != B.shape[0]:
            raise RError("non-conformable arguments")
        return from_matrix(A @ B)
    if y.dim is None:
        A, b = matrix_of(x), as_float(y)
        Bm = b.reshape(-1, 1) if len(b) == A.shape[1] else b.reshape(1, -1)
        if A.shape[1] != Bm.shape[0]:
            raise RError("non-conformable arguments")
        return from_matrix(A @ Bm)
    A, B = matrix_of(x), matrix_of(y)
    if A.shape[1] != B.shape[0]:
        raise RError("non-conformable arguments")
    return from_matrix(A @ B)


# ------------------------------------------------------------------------------------------------ indexing
def _resolve_index(idx, n, names, allow_extend=False):
    """R index -> 0-based integer numpy array (may exceed n - 1 when allow_extend)."""
    if idx is MISSING or idx is None:
        return np.arange(n)
    if not isinstance(idx, Vec):
        raise RError("invalid subscript type")
    v = idx.v
    if v.dtype == object:
        out = []
        for s in v:
            if names is not None and s in names:
                out.append(names.index(s))
            elif allow_extend:
                out.append(-1)          # caller appends
            else:
                raise RError("subscript out of bounds: %r" % s)
        return np.array(out, dtype=np.int64)
    if v.dtype == bool:
        m = _recycle(v, max(n, len(v)))
        return np.nonzero(m)[0]
    iv = v.astype(np.int64) if v.dtype.kind != "f" else np.trunc(v).astype(np.int64)
    if len(iv) and (iv < 0).any():
        if (iv > 0).any():
            raise RError("can't mix positive and negative subscripts")
        keep = np.ones(n, dtype=bool)
        keep[(-iv[iv != 0] - 1)[(-iv[iv != 0] - 1) < n]] = False
        return np.nonzero(keep)[0]
    iv = iv[iv != 0]
    if not allow_extend and len(iv) and iv.max() > n:
        return iv - 1          # out of range: caller decides (NA for vectors, error for lists / matrices)
    return iv - 1


def index_get(x, args, double):
    if isinstance(x, (Closure, Builtin)):
        raise RError("object of type 'closure' is not subsettable")
    if double:
        if len(args) != 1:
            if isinstance(x, Vec) and x.dim is not None and len(args) == 2:
                return index_get(x, args, False)
            raise RError("[[ ]] with %d subscripts" % len(args))
        i = args[0]
        if isinstance(x, RList):
            if is_chr(i):
                r = x.get(i.v[0], partial=False)
                return r
            k = int(scalar(i)) - 1
            if k < 0 or k >= len(x.items):
                raise RError("subscript out of bounds")
            return x.items[k]
        if x is None:
            return None
        if is_chr(i):
            if x.names is None or i.v[0] not in x.names:
                raise RError("subscript out of bounds")
            k = x.names.index(i.v[0])
        else:
            k = int(scalar(i)) - 1
        if k < 0 or k >= len(x.v):
            raise RError("subscript out of bounds")
        return Vec(x.v[k:k + 1].copy())
    if x is None:
        return None
    if isinstance(x, RList):
        ii = _resolve_index(args[0] if args else MISSING, len(x.items), x.names)
        return RList([x.items[k] if k < len(x.items) else None for k in ii],
                     [x.names[k] if k < len(x.items) else "" for k in ii] if x.names else None)
    drop = True
    pos = [a for (nm, a) in args if nm != "drop"] if args and isinstance(args[0], tuple) else args
    if args and isinstance(args[0], tuple):
        for nm, a in args:
            if nm == "drop":
                drop = truthy(a)
    if len(pos) == 1:
        i = pos[0]
        if isinstance(i, Vec) and i.dim is not None and x.dim is not None and i.v.dtype == bool:
            ii = np.nonzero(i.v)[0]
        else:
            ii = _resolve_index(i, len(x.v), x.names)
        if len(ii) and ii.max() >= len(x.v):
            vals = np.full(len(ii), np.nan)
            ok = ii < len(x.v)
            vals[ok] = x.v[ii[ok]]
            return Vec(vals)
        return Vec(x.v[ii], names=[x.names[k] for k in ii] if x.names else None)
    if len(pos) == 2:
        if x.dim is None:
            raise RError("incorrect number of dimensions")
        nr, nc = x.dim
        ri = _resolve_index(pos[0], nr, None)
        ci = _resolve_index(pos[1], nc, None)
        if (len(ri) and ri.max() >= nr) or (len(ci) and ci.max() >= nc):
            raise RError("subscript out of bounds")
        M = x.v.reshape((nr, nc), order="F")[np.ix_(ri, ci)]
        if drop and (len(ri) == 1 or len(ci) == 1):
            return Vec(np.ascontiguousarray(M.reshape(-1, order="F")))
        return Vec(np.ascontiguousarray(M.reshape(-1, order="F")), dim=(len(ri), len(ci)))
    raise RError("incorrect number of subscripts")


def _coerce_pair(x, value):
    """Common storage type for assignment of `value` into atomic `x`."""
    a, b = x.v, value.v
    if a.dtype == object or b.dtype == object:
        def to_s(arr):
            if arr.dtype == object:
                return arr
            out = np.empty(len(arr), dtype=object)
            for k, e in enumerate(arr):
                out[k] = _format_num(e)
            return out
        return to_s(a), to_s(b)
    if a.dtype == b.dtype:
        return a, b
    if a.dtype == np.float64 or b.dtype == np.float64:
        return a.astype(np.float64), b.astype(np.float64)
    return a.astype(np.int64), b.astype(np.int64)


def index_set(x, args, double, value):
    if x is None:
        x = RList([], None) if isinstance(value, (RList, Closure, Builtin)) or (double and not isinstance(value, Vec)) else \
            Vec(np.zeros(0, dtype=value.v.dtype if isinstance(value, Vec) else np.float64))
    if isinstance(x, RList) or (isinstance(x, Vec) and double and isinstance(value, (RList, Closure, Builtin))):
        if isinstance(x, Vec):
            x = RList([Vec(x.v[k:k + 1]) for k in range(len(x.v))], list(x.names) if x.names else None)
        items, names = list(x.items), (list(x.names) if x.names else None)
        i = args[0]
        if double:
            if is_chr(i):
                nm = i.v[0]
                if names and nm in names:
                    k = names.index(nm)
                else:
                    k = len(items)
                    if names is None:
                        names = [""] * len(items)
                    names.append(nm)
                    items.append(None)
            else:
                k = int(scalar(i)) - 1
                while len(items) <= k:
                    items.append(None)
                    if names is not None:
                        names.append("")
            if value is None:
                del items[k]
                if names is not None:
                    del names[k]
            else:
                items[k] = value
            return RList(items, names)
        ii = _resolve_index(i, len(items), names, allow_extend=True)
        vals = value.items if isinstance(value, RList) else [value] if not isinstance(value, Vec) else \
            [Vec(value.v[k:k + 1]) for k in range(len(value.v))]
        for q, k in enumerate(ii):
            if k == -1:
                k = len(items)
                if names is None:
                    names = [""] * len(items)
                names.append(i.v[q])
                items.append(None)
            while len(items) <= k:
                items.append(None)
                if names is not None:
                    names.append("")
            items[k] = vals[q % len(vals)]
        return RList(items, names)
    if not isinstance(value, Vec):
        raise RError("cannot assign a %s into an atomic vector" % type(value).__name__)
    pos = args
    if len(pos) == 1:
        i = pos[0]
        if isinstance(i, Vec) and i.v.dtype == bool and i.dim is not None:
            ii = np.nonzero(i.v)[0]
        else:
            ii = _resolve_index(i, len(x.v), x.names, allow_extend=True)
        a, b = _coerce_pair(x, value)
        a = a.copy()
        names = list(x.names) if x.names else None
        if len(ii) and (ii == -1).any():
            for q in np.nonzero(ii == -1)[0]:
                if names is None:
                    names = [""] * len(a)
                names.append(i.v[q])
                a = np.concatenate([a, np.array([np.nan if a.dtype == np.float64 else 0], dtype=a.dtype)])
                ii[q] = len(a) - 1
        if len(ii) and ii.max() >= len(a):
            ext = ii.max() + 1 - len(a)
            fill = np.full(ext, np.nan) if a.dtype == np.float64 else np.zeros(ext, dtype=a.dtype)
            if a.dtype == np.int64 or a.dtype == bool:
                a = a.astype(np.float64)
                b = b.astype(np.float64)
                fill = np.full(ext, np.nan)
            a = np.concatenate([a, fill])
            if names is not None:
                names += [""] * ext
            dim = None
        else:
            dim = x.dim
        if len(ii) == 0:
            return Vec(a, dim=dim, names=names)
        if len(b) == 0:
            raise RError("replacement has length zero")
        a[ii] = _recycle(b, len(ii))
        return Vec(a, dim=dim, names=names)
    if len(pos) == 2:
        if x.dim is None:
            raise RError("incorrect number of subscripts on matrix")
        nr, nc = x.dim
        ri = _resolve_index(pos[0], nr, None)
        ci = _resolve_index(pos[1], nc, None)
        a, b = _coerce_pair(x, value)
        M = a.reshape((nr, nc), order="F").copy(order="F")
        cnt = len(ri) * len(ci)
        if cnt:
            M[np.ix_(ri, ci)] = _recycle(b, cnt).reshape((len(ri), len(ci)), order="F")
        return Vec(np.ascontiguousarray(M.reshape(-1, order="F")), dim=(nr, nc), names=x.names)
    raise RError("incorrect number of subscripts")


def _format_num(e):
    if isinstance(e, (bool, np.bool_)):
        return "TRUE" if e else "FALSE"
    if isinstance(e, (int, np.integer)):
        return str(int(e))
    f = float(e)
    if math.isnan(f):
        return "NA"
    if f == int(f) and abs(f) < 1e15:
        return str(int(f))
    return repr(f) if len(repr(f)) <= 17 else "%.15g" % f


# ------------------------------------------------------------------------------------------------ interpreter
class Interp:
    def __init__(self):
        self.globalenv = Env()
        self.base = self.globalenv
        from . import builtins as B
        B.install(self)

    def run(self, src, env=None):
        env = env or self.globalenv
        out = None
        for node in rparse.parse(src):
            out = self.eval(node, env)
        return out

    def source(self, path):
        with open(path) as f:
            return self.run(f.read())

    # ---- evaluation
    def eval(self, n, env):
        k = n[0]
        if k == "num":
            return Vec(np.array([n[1]], dtype=np.float64))
        if k == "int":
            return Vec(np.array([n[1]], dtype=np.int64))
        if k == "str":
            return chrv([n[1]])
        if k == "const":
            v = n[1]
            if isinstance(v, bool):
                return Vec(np.array([v]))
            if v is None:
                return chrv([None])
            return Vec(np.array([v], dtype=np.float64))
        if k == "null":
            return None
        if k == "sym":
            return self.get_var(n[1], env)
        if k == "value":                      # a value spliced in by substitute()
            return n[1]
        if k == "call":
            return self.eval_call(n, env)
        if k == "binop":
            return self.eval_binop(n, env)
        if k == "unop":
            x = self.eval(n[2], env)
            if n[1] == "-":
                v = x.v.astype(np.int64) if x.v.dtype == bool else x.v
                return Vec(-v, dim=x.dim, names=x.names)
            if n[1] == "+":
                return x
            if n[1] == "!":
                if x.v.dtype == bool:
                    return Vec(~x.v, dim=x.dim)
                return Vec(as_float(x) == 0, dim=x.dim)
            raise RError("unary %s" % n[1])
        if k == "assign":
            v = self.eval(n[2], env)
            self.assign(n[1], v, env, n[3])
            return v
        if k == "block":
            out = None
            for s in n[1]:
                out = self.eval(s, env)
            return out
        if k == "if":
            if truthy(self.eval(n[1], env)):
                return self.eval(n[2], env)
            if n[3] is not None:
                return self.eval(n[3], env)
            return None
        if k == "for":
            seq = self.eval(n[2], env)
            items = seq.items if isinstance(seq, RList) else [] if seq is None else \
                [Vec(seq.v[i:i + 1]) for i in range(len(seq.v))]
            for it in items:
                env.set(n[1], it)
                try:
                    self.eval(n[3], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "while":
            while truthy(self.eval(n[1], env)):
                try:
                    self.eval(n[2], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "repeat":
            while True:
                try:
                    self.eval(n[1], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "break":
            raise _Break()
        if k == "next":
            raise _Next()
        if k == "function":
            return Closure(n[1], n[2], env)
        if k == "index":
            x = self.eval(n[1], env)
            args = self.eval_index_args(n[2], env)
            if n[3]:
                return index_get(x, [a for (_, a) in args], True)
            if isinstance(x, Vec) and any(nm == "drop" for nm, _ in args):
                return index_get(x, args, False)
            return index_get(x, [a for (_, a) in args], False)
        if k == "dollar":
            x = self.eval(n[1], env)
            if isinstance(x, RList):
                return x.get(n[2])
            if x is None:
                return None
            if isinstance(x, Vec):
                raise RError("$ operator is invalid for atomic vectors")
            raise RError("$ on %s" % type(x).__name__)
        if k == "dollar_expr":
            x = self.eval(n[1], env)
            nm = self.eval(n[2], env)
            return x.get(nm.v[0]) if isinstance(x, RList) else None
        if k == "ns":
            return self.get_var(n[1] + "::" + n[2], env)
        if k == "formula":
            return Lang(n)
        raise RError("cannot evaluate node %r" % (k,))

    def eval_index_args(self, args, env):
        out = []
        for nm, a in args:
            out.append((nm, MISSING if a is None else self.eval(a, env)))
        return out

    def force(self, p):
        if isinstance(p, Promise):
            if not p.done:
                p.value = self.eval(p.ast, p.env)
                p.done = True
            return p.value
        return p

    def get_var(self, name, env):
        e, v = env.lookup(name)
        if e is None:
            raise RError("object '%s' not found" % name)
        if v is MISSING:
            raise RError("argument \"%s\" is missing, with no default" % name)
        return self.force(v)

    def get_fun(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = self.force(e.vars[name]) if e.vars[name] is not MISSING else None
                if isinstance(v, (Closure, Builtin)):
                    return v
            e = e.parent
        raise RError("could not find function \"%s\"" % name)

    def eval_binop(self, n, env):
        op = n[1]
        if op == "&&":
            return lgl(truthy(self.eval(n[2], env)) and truthy(self.eval(n[3], env)))
        if op == "||":
            return lgl(truthy(self.eval(n[2], env)) or truthy(self.eval(n[3], env)))
        x = self.eval(n[2], env)
        y = self.eval(n[3], env)
        if op == "%*%":
            return matmul(x, y)
        if op == "%in%":
            xs = x.items if isinstance(x, RList) else list(x.v) if x is not None else []
            ys = set(y.v.tolist()) if isinstance(y, Vec) else set()
            return Vec(np.array([e in ys for e in xs], dtype=bool))
        if op == ":":
            a, b = float(scalar(x)), float(scalar(y))
            n_ = int(math.floor(abs(b - a) + 1e-10)) + 1
            step = 1 if b >= a else -1
            if a == int(a):
                return Vec(np.arange(int(a), int(a) + step * n_, step, dtype=np.int64))
            return Vec(a + step * np.arange(n_, dtype=np.float64))
        if op == "%o%":
            return from_matrix(np.outer(as_float(x), as_float(y)))
        if op.startswith("%") and op not in ("%%", "%/%"):
            f = self.get_fun(op, env)
            return self.apply_function(f, [(None, x), (None, y)], env)
        return arith(op, x, y)

    # ---- calls
    def eval_call(self, n, env):
        fn_ast = n[1]
        if fn_ast[0] == "sym":
            f = self.get_fun(fn_ast[1], env)
        else:
            f = self.eval(fn_ast, env)
        if not isinstance(f, (Closure, Builtin)):
            raise RError("attempt to apply non-function")
        if isinstance(f, Builtin) and f.special:
            return f.fn(self, n[2], env)
        args = []
        for nm, a in n[2]:
            if a is None:
                args.append((nm, MISSING))
            elif a[0] == "sym" and a[1] == "...":
                e, dots = env.lookup("...")
                if e is not None and dots:
                    for dn, dv in dots:
                        args.append((dn, self.force(dv)))
            else:
                args.append((nm, self.eval(a, env)))
        return self.apply_function(f, args, env)

    def apply_function(self, f, args, env=None):
        if isinstance(f, Builtin):
            pos = [a for (nm, a) in args if nm is None]
            kw = {nm: a for (nm, a) in args if nm is not None}
            return f.fn(self, pos, kw)
        formals = f.formals
        fenv = Env(f.env)
        names = [fm[0] for fm in formals]
        bound = {}
        rest = []
        used = [False] * len(args)
        # exact
        for k, (nm, a) in enumerate(args):
            if nm is not None and nm in names and nm != "..." and nm not in bound:
                bound[nm] = a
                used[k] = True
        # partial (formals before ... only)
        pre = names[:names.index("...")] if "..." in names else names
        for k, (nm, a) in enumerate(args):
            if used[k] or nm is None:
                continue
            hits = [fn for fn in pre if fn.startswith(nm) and fn not in bound]
            if len(hits) == 1:
                bound[hits[0]] = a
                used[k] = True
        # positional
        free = [fn for fn in names if fn not in bound and fn != "..."]
        if "..." in names:
            before = [fn for fn in names[:names.index("...")] if fn not in bound]
        else:
            before = free
        bi = 0
        for k, (nm, a) in enumerate(args):
            if used[k]:
                continue
            if nm is None and bi < len(before):
                bound[before[bi]] = a
                bi += 1
                used[k] = True
            elif "..." in names:
                rest.append((nm, a))
                used[k] = True
            else:
                raise RError("unused argument in call to %s" % f.name)
        for fn, default in formals:
            if fn == "...":
                fenv.vars["..."] = rest
            elif fn in bound and bound[fn] is not MISSING:
                fenv.vars[fn] = bound[fn]
            elif default is not None:
                fenv.vars[fn] = Promise(default, fenv)
            else:
                fenv.vars[fn] = MISSING
        try:
            return self.eval(f.body, fenv)
        except _Return as r:
            return r.value

    # ---- assignment
    def assign(self, target, value, env, superassign=False):
        k = target[0]
        if k == "str":
            target = ("sym", target[1])
            k = "sym"
        if k == "sym":
            if isinstance(value, Closure) and value.name == "<anonymous>":
                value.name = target[1]
            if superassign:
                env.set_super(target[1], value)
            else:
                env.set(target[1], value)
            return
        # replacement: f(x, args) <- value  ==>  x <- `f<-`(x, args, value)
        if k == "index":
            obj = self.eval_target(target[1], env)
            args = [a for (_, a) in self.eval_index_args(target[2], env)]
            new = index_set(obj, args, target[3], value)
            self.assign(target[1], new, env, superassign)
            return
        if k == "dollar":
            obj = self.eval_target(target[1], env)
            new = index_set(obj if obj is not None else RList([], None), [chrv([target[2]])], True, value)
            self.assign(target[1], new, env, superassign)
            return
        if k == "call" and target[1][0] == "sym":
            fname = target[1][1] + "<-"
            f = self.get_fun(fname, env)
            obj = self.eval_target(target[2][0][1], env)
            extra = [(nm, self.eval(a, env)) for nm, a in target[2][1:]]
            new = self.apply_function(f, [(None, obj)] + extra + [("value", value)], env)
            self.assign(target[2][0][1], new, env, superassign)
            return
        raise RError("invalid assignment target")

    def eval_target(self, t, env):
        try:
            return self.eval(t, env)
        except RError:
            if t[0] == "sym":
                return None
            raise


def to_py(x):
    """Vec/RList -> numpy / dict for the Python side."""
    if x is None:
        return None
    if isinstance(x, Vec):
        a = x.v
        if x.dim is not None:
            return a.reshape(x.dim, order="F").copy()
        if x.names is not None and len(a) > 0 and any(x.names):
            return {nm: a[k] for k, nm in enumerate(x.names)}
        return a.copy()
    if isinstance(x, RList):
        if x.names and all(x.names):
            return {nm: to_py(v) for nm, v in zip(x.names, x.items)}
        return [to_py(v) for v in x.items]
    return x


def from_py(x):
    """numpy / dict / str / list[str] / float -> R value."""
    if x is None:
        return None
    if isinstance(x, (Vec, RList, Closure, Builtin)):
        return x
    if isinstance(x, dict):
        return RList([from_py(v) for v in x.values()], list(x.keys()))
    if isinstance(x, str):
        return chrv([x])
    if isinstance(x, bool):
        return lgl(x)
    if isinstance(x, (list, tuple)) and len(x) and all(isinstance(s, str) for s in x):
        return chrv(list(x))
    a = np.asarray(x)
    if a.dtype == bool:
        return Vec(a.reshape(-1, order="F").copy(), dim=a.shape if a.ndim == 2 else None)
    if a.dtype.kind in "iu":
        a = a.astype(np.int64)
    else:
        a = a.astype(np.float64)
    if a.ndim == 2:
        return Vec(np.ascontiguousarray(a.reshape(-1, order="F")), dim=a.shape)
    return Vec(np.ascontiguousarray(a.reshape(-1)))
