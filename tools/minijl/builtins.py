"""Base / LinearAlgebra / DelimitedFiles / ArgParse / Logging functions the reference's MCMC path calls, for minijl.

`rand` is deliberately NOT provided: the fixture scripts define it themselves (a scripted stream of uniforms), and a
call that reaches an unscripted `rand` fails loudly instead of producing irreproducible numbers.
"""
from __future__ import annotations

import math
import os
import sys
import time as _time

import numpy as np

from .interp import (COLON, Env, Interp, JFunction, JList, JRange, JStruct, JType, JTypeApp, JlError, Method, ModuleNS, Sym,
                     getindex, iterate, jl_float, jl_repr, jl_str, make_vector, setindex)
from .parser import parse_expression

_interp: Interp = None


def _arr(x):
    if isinstance(x, JRange):
        return np.array(list(x))
    if isinstance(x, JList):
        return np.array(list(x), dtype=object)
    return x


def _unbox(x):
    return x.item() if isinstance(x, np.generic) else x


# ---- math -------------------------------------------------------------------------------------------------------------
def _real1(fn, name):
    def f(x):
        if isinstance(x, np.ndarray):
            raise JlError(f"MethodError: no method matching {name}(::Array); use broadcasting")
        try:
            return fn(x)
        except ValueError:
            raise JlError(f"DomainError: {name}({x})")
    f.__name__ = name
    return f


def jl_log(x):
    if isinstance(x, np.ndarray):
        raise JlError("MethodError: log of an array")
    if x == 0:
        return -math.inf
    if x < 0:
        raise JlError(f"DomainError with {x}: log will only return a complex result if called with a complex argument")
    return math.log(x)


def jl_exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def jl_sqrt(x):
    if x < 0:
        raise JlError(f"DomainError with {x}: sqrt")
    return math.sqrt(x)


def jl_min(*a):
    r = a[0]
    for x in a[1:]:
        if x != x or r != r:
            return math.nan
        if x < r:
            r = x
    if any(isinstance(x, float) for x in a):
        r = float(r)
    return r


def jl_max(*a):
    r = a[0]
    for x in a[1:]:
        if x != x or r != r:
            return math.nan
        if x > r:
            r = x
    if any(isinstance(x, float) for x in a):
        r = float(r)
    return r


def jl_round(*a):
    if len(a) == 2 and isinstance(a[0], (JType,)):
        return int(round(a[1]))
    x = a[0]
    return float(round(x)) if isinstance(x, float) else x


def jl_floor(*a):
    if len(a) == 2:
        return int(math.floor(a[1]))
    return float(math.floor(a[0])) if isinstance(a[0], float) else a[0]


def jl_ceil(*a):
    if len(a) == 2:
        return int(math.ceil(a[1]))
    return float(math.ceil(a[0])) if isinstance(a[0], float) else a[0]


def jl_abs(x):
    return abs(x)


# ---- arrays ------------------------------------------------------------------------------------------------------------
def zeros(*dims):
    if dims and isinstance(dims[0], (JType, JTypeApp)):
        dims = dims[1:]
    if len(dims) == 1 and isinstance(dims[0], tuple):
        dims = dims[0]
    return np.zeros(tuple(dims), dtype=np.float64)


def ones(*dims):
    if dims and isinstance(dims[0], (JType, JTypeApp)):
        dims = dims[1:]
    return np.ones(tuple(dims), dtype=np.float64)


def fill(v, *dims):
    if isinstance(v, (int, float)) and not isinstance(v, bool):
        return np.full(tuple(dims), v, dtype=np.float64 if isinstance(v, float) else np.int64)
    out = JList([v] * dims[0])
    return out


def length(x):
    if isinstance(x, np.ndarray):
        return int(x.size)
    if isinstance(x, (JList, JRange, tuple, str, dict, list)):
        return len(x)
    raise JlError(f"MethodError: no method matching length({x!r:.60})")


def size(x, d=None):
    x = _arr(x)
    if d is None:
        return tuple(int(s) for s in x.shape)
    return int(x.shape[d - 1]) if d <= x.ndim else 1


def jl_sum(*a, dims=None):
    if len(a) == 2:          # sum(f, xs)
        f, xs = a
        tot = None
        for x in iterate(xs):
            v = _interp.call(f, [x])
            tot = v if tot is None else tot + v
        return tot if tot is not None else 0
    x = _arr(a[0])
    if isinstance(x, tuple):
        return sum(x)
    if dims is None:
        if x.dtype == object:
            tot = None
            for v in x.flatten(order="F"):
                tot = v if tot is None else tot + v
            return tot
        return _unbox(x.sum())
    return x.sum(axis=dims - 1, keepdims=True)


def jl_prod(x):
    x = _arr(x)
    return _unbox(np.prod(x))


def cumsum(x, dims=None):
    x = _arr(x)
    if x.ndim == 1:
        return np.cumsum(x)
    if dims is None:
        raise JlError("cumsum of a matrix needs dims")
    return np.cumsum(x, axis=dims - 1)


def jl_map(f, *xs):
    its = [list(iterate(x)) for x in xs]
    n = len(its[0])
    for it in its:
        if len(it) != n:
            raise JlError("DimensionMismatch in map")
    out = [_interp.call(f, [it[k] for it in its]) for k in range(n)]
    if isinstance(xs[0], tuple):
        return tuple(out)
    return make_vector(out)


def foreach(f, xs):
    for x in iterate(xs):
        _interp.call(f, [x])
    return None


def dot(a, b):
    a, b = _arr(a), _arr(b)
    if a.shape != b.shape:
        raise JlError(f"DimensionMismatch: dot of {a.shape} and {b.shape}")
    # sequential like the generic fallback (3-vectors here); BLAS may reassociate, tolerances cover it
    s = 0.0
    for x, y in zip(a.tolist(), b.tolist()):
        s += x * y
    return s


def _as2d_for_hcat(x):
    if isinstance(x, np.ndarray):
        return x.reshape(-1, 1) if x.ndim == 1 else x
    if isinstance(x, JList):
        return np.array(list(x), dtype=object).reshape(-1, 1)
    if isinstance(x, JRange):
        return np.array(list(x)).reshape(-1, 1)
    return None


def hcat(*xs):
    if all(not isinstance(x, (np.ndarray, JList, JRange)) for x in xs):
        vals = list(xs)
        if all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in vals):
            dt = np.int64 if all(isinstance(v, int) for v in vals) else np.float64
            return np.array(vals, dtype=dt).reshape(1, -1)
        return np.array(vals, dtype=object).reshape(1, -1)
    mats = []
    rows = None
    for x in xs:
        m = _as2d_for_hcat(x)
        if m is not None:
            rows = m.shape[0] if rows is None else rows
            if m.shape[0] != rows:
                raise JlError("DimensionMismatch: hcat rows differ")
    for x in xs:
        m = _as2d_for_hcat(x)
        if m is None:
            if rows != 1:
                raise JlError("DimensionMismatch: hcat of a scalar with a multi-row array")
            m = np.array([[x]], dtype=object if isinstance(x, str) else None)
        mats.append(m)
    if any(m.dtype == object for m in mats):
        mats = [m.astype(object) for m in mats]
    elif any(m.dtype == np.float64 for m in mats):
        mats = [m.astype(np.float64) for m in mats]
    return np.hstack(mats)


def vcat(*xs):
    if all(not isinstance(x, (np.ndarray, JList, JRange)) for x in xs):
        return make_vector(list(xs))
    if any(isinstance(x, np.ndarray) and x.ndim == 2 for x in xs):
        mats = [x if isinstance(x, np.ndarray) and x.ndim == 2 else np.array(list(iterate(x)), dtype=object).reshape(1, -1)
                for x in xs]
        if any(m.dtype == object for m in mats):
            mats = [m.astype(object) for m in mats]
        return np.vstack(mats)
    out = []
    for x in xs:
        if isinstance(x, (np.ndarray, JList, JRange)):
            out.extend(iterate(x))
        else:
            out.append(x)
    return make_vector(out)


def reshape(x, *dims):
    if len(dims) == 1 and isinstance(dims[0], tuple):
        dims = dims[0]
    x = _arr(x)
    total = x.size
    known = 1
    for d in dims:
        if d is not COLON:
            known *= d
    shape = tuple((total // known) if d is COLON else d for d in dims)
    r = x.reshape(shape, order="F")
    return r if r.base is None else r.copy()     # (aliasing of reshape is not relied upon by the reference)


def transpose(x):
    if isinstance(x, (int, float)) and not isinstance(x, np.ndarray):
        return x                       # transpose of a number is the number
    if isinstance(x, (JList, list, tuple)):   # a tuple: a vector that served as a Dict key (interp.dict_key)
        x = np.array(list(x), dtype=object if any(isinstance(e, str) for e in x) else None)
    x = _arr(x)
    if x.ndim == 1:
        return x.reshape(1, -1)
    return x.T.copy()


def permutedims(x, perm=None):
    """permutedims(v) of a vector is the 1×n row matrix; of a matrix, its (non-recursive) transpose."""
    if isinstance(x, (JList, list)):
        x = np.array(list(x), dtype=object if any(isinstance(e, str) for e in x) else None)
    x = _arr(x)
    if perm is not None:
        return np.asfortranarray(np.transpose(x, [int(p) - 1 for p in iterate(perm)]))
    if x.ndim == 1:
        return x.reshape(1, -1)
    return np.asfortranarray(x.T)


def jl_copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, JList):
        out = JList(x)
        out.eltype = x.eltype
        return out
    if isinstance(x, dict):
        return dict(x)
    return x


def view(a, *idxs):
    if not isinstance(a, np.ndarray):
        raise JlError("view of a non-array")
    from .interp import _norm_index
    if a.ndim == 1:
        ix, _ = _norm_index(idxs[0], a.shape[0])
        return a[ix]
    i0, _ = _norm_index(idxs[0], a.shape[0])
    i1, _ = _norm_index(idxs[1], a.shape[1])
    return a[i0, i1]


def collect(x):
    if isinstance(x, JRange):
        return np.array(list(x), dtype=np.int64 if isinstance(x.start, int) else np.float64)
    if isinstance(x, np.ndarray):
        return x.copy()
    return make_vector(list(iterate(x)))


def findnext(pred, A, i):
    if isinstance(pred, str) and isinstance(A, str):     # findnext(pattern, string, start): the range of the match
        k = A.find(pred, i - 1)
        return None if k < 0 else JRange(k + 1, k + len(pred))
    items = list(iterate(A))
    for k in range(i - 1, len(items)):
        if _interp.call(pred, [items[k]]) is True:
            return k + 1
    return None


def findfirst(pred, A):
    return findnext(pred, A, 1)


def push_(v, *xs):
    if isinstance(v, JList):
        v.extend(xs)
        return v
    raise JlError("push! on a numeric vector is not supported by minijl (use a Vector{Any})")


def popfirst_(v):
    if isinstance(v, JList):
        if not v:
            raise JlError("ArgumentError: array must be non-empty (popfirst!)")
        return v.pop(0)
    raise JlError("popfirst! needs a Vector{Any}")


def pop_(v):
    return v.pop()


def append_(v, xs):
    v.extend(iterate(xs))
    return v


def empty_(v):
    if isinstance(v, (JList, dict)):
        v.clear()
        return v
    raise JlError("empty! unsupported")


def isempty(v):
    return length(v) == 0


def eigvals(M):
    return np.linalg.eigvalsh(np.asarray(M, dtype=np.float64))


def norm(x):
    return float(np.sqrt(np.sum(np.asarray(x, dtype=np.float64) ** 2)))


def jl_maximum(x):
    return _unbox(np.max(_arr(x)))


def jl_minimum(x):
    return _unbox(np.min(_arr(x)))


def jl_any(*a):
    if len(a) == 2:
        return any(_interp.call(a[0], [x]) is True for x in iterate(a[1]))
    return any(x is True or x == True for x in iterate(a[0]))  # noqa: E712


def jl_all(*a):
    if len(a) == 2:
        return all(_interp.call(a[0], [x]) is True for x in iterate(a[1]))
    return all(x is True or x == True for x in iterate(a[0]))  # noqa: E712


def first(x):
    return getindex(x, [1])


def last(x):
    return getindex(x, [length(x)])


# ---- strings / IO --------------------------------------------------------------------------------------------------------
def string(*a):
    return "".join(jl_str(x) for x in a)


def println(*a):
    io = _interp.stdout
    if a and hasattr(a[0], "write") and not isinstance(a[0], str):
        io, a = a[0], a[1:]
    io.write("".join(jl_str(x) for x in a) + "\n")
    return None


def jl_print(*a):
    io = _interp.stdout
    if a and hasattr(a[0], "write") and not isinstance(a[0], str):
        io, a = a[0], a[1:]
    io.write("".join(jl_str(x) for x in a))
    return None


def jl_parse(t, s):
    T = _interp.types
    if t is T["Float64"]:
        s2 = s.strip()
        if s2 in ("Inf", "+Inf"):
            return math.inf
        if s2 == "-Inf":
            return -math.inf
        if s2 == "NaN":
            return math.nan
        return float(s2)
    if t in (T["Int64"],):
        return int(s.strip())
    if t is T["Bool"]:
        s2 = s.strip()
        if s2 in ("true", "1"):
            return True
        if s2 in ("false", "0"):
            return False
        raise JlError(f"ArgumentError: invalid Bool representation: {s!r}")
    raise JlError(f"parse({t}, ...) unsupported")


def split(s, sep=None, limit=0):
    """split(s, sep; limit=0): at most `limit` pieces (0: no limit), like Base.split."""
    maxsplit = limit - 1 if limit and limit > 0 else -1
    parts = s.split(sep, maxsplit) if sep is not None else s.split(None, maxsplit)
    return JList(parts)


class GlobMatch:
    """Glob.GlobMatch(pattern) — what readdir(pattern, dir) of Glob.jl takes."""

    def __init__(self, pattern):
        self.pattern = pattern


def readdlm(path, delim=None, header=False):
    """readdlm(file, delim; header=false): numbers become Float64; a table with text or ragged rows becomes a matrix of
    Any with "" for the missing cells, as DelimitedFiles does.  With header=true returns (data, 1×n header matrix)."""
    with open(path, encoding="utf-8") as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip() != ""]
    rows = [ln.split(delim) if delim is not None else ln.split() for ln in lines]
    head = None
    if header:
        head, rows = rows[0], rows[1:]
    width = max([len(r) for r in rows] + ([len(head)] if head else [0]))

    def cell(x):
        x = x.strip()
        try:
            return float(x)
        except ValueError:
            return x
    data = [[cell(x) for x in r] + [""] * (width - len(r)) for r in rows]
    if all(isinstance(x, float) for r in data for x in r):
        arr = np.array(data, dtype=np.float64).reshape(len(data), width)
    else:
        arr = np.empty((len(data), width), dtype=object)
        for i, r in enumerate(data):
            for j, x in enumerate(r):
                arr[i, j] = x
    arr = np.asfortranarray(arr)
    if header:
        h = np.empty((1, len(head)), dtype=object)
        for j, x in enumerate(head):
            h[0, j] = x.strip()
        return (arr, h)
    return arr


def jl_filter(f, xs):
    keep = [x for x in iterate(xs) if _interp.call(f, [x]) is True]
    return make_vector(keep)


def jl_sort(xs, rev=False):
    return make_vector(sorted(iterate(xs), reverse=bool(rev)))


def jl_readdir(*a):
    """readdir(dir) (sorted names) and Glob's readdir(GlobMatch, dir) (sorted matching paths, joined with dir)."""
    import fnmatch
    if len(a) == 2 and isinstance(a[0], GlobMatch):
        pat, d = a[0].pattern, a[1]
        return JList([os.path.join(d, f) for f in sorted(os.listdir(d)) if fnmatch.fnmatchcase(f, pat)])
    d = a[0] if a else "."
    return JList(sorted(os.listdir(d)))


def jl_open(*a):
    """open(path[, mode]) or the do-block form open(f, path[, mode]): f(io), then close — also when f throws."""
    if a and not isinstance(a[0], str):
        f, rest = a[0], a[1:]
        io = open(rest[0], rest[1] if len(rest) > 1 else "r", encoding="utf-8")
        try:
            return _interp.call(f, [io])
        finally:
            io.close()
    return open(a[0], a[1] if len(a) > 1 else "r", encoding="utf-8")


def jl_close(io):
    io.close()
    return None


class Cmd:
    """A command literal.  `julia script.jl args…` is run by a nested interpreter (there is no julia here); anything else
    is refused — the interpreter is test infrastructure, not a shell."""

    def __init__(self, argv):
        self.argv = list(argv)

    def __repr__(self):
        return "`" + " ".join(self.argv) + "`"


def run_cmd_text(cmd: Cmd):
    """stdout of a command.  interp.cmd_runner(argv) -> str, when set, stands in for the process (tests)."""
    runner = getattr(_interp, "cmd_runner", None)
    if runner is not None:
        return runner(cmd.argv)
    return "\n".join(run_cmd_lines(cmd)) + "\n"


def jl_read(x, t=None):
    if isinstance(x, Cmd):
        return run_cmd_text(x)
    if isinstance(x, str):
        with open(x, encoding="utf-8") as f:
            return f.read()
    return x.read()


def run_cmd_lines(cmd: Cmd):
    import io as _io
    from .interp import Interp as _Interp
    runner = getattr(_interp, "cmd_runner", None)
    if runner is not None:
        return JList(runner(cmd.argv).splitlines())
    if len(cmd.argv) < 2 or cmd.argv[0] != "julia" or not cmd.argv[1].endswith(".jl"):
        raise JlError(f"minijl: cannot run {cmd!r}")
    child = _Interp(argv=cmd.argv[2:])
    out = _io.StringIO()
    child.stdout = out
    child.genv.vars["stdout"] = out
    child.run_main(os.path.abspath(cmd.argv[1]))
    return JList(out.getvalue().splitlines())


def readlines(x):
    if isinstance(x, Cmd):
        return run_cmd_lines(x)
    if isinstance(x, str):
        with open(x, encoding="utf-8") as f:
            return JList(f.read().splitlines())
    return JList(x.read().splitlines())


def _dlm_cell(v):
    v = _unbox(v)
    if isinstance(v, float):
        return jl_float(v)
    return jl_str(v)


def writedlm(io, A, delim="\t"):
    if isinstance(io, str):                      # writedlm(path, A, delim)
        with open(io, "w", encoding="utf-8") as f:
            return writedlm(f, A, delim)
    A = _arr(A)
    if isinstance(A, np.ndarray) and A.ndim == 2:
        for row in A:
            io.write(delim.join(_dlm_cell(v) for v in row) + "\n")
    elif isinstance(A, np.ndarray):
        for v in A:
            io.write(_dlm_cell(v) + "\n")
    else:
        raise JlError("writedlm: unsupported argument")
    return None


def jl_write(io, *a):
    if isinstance(io, str):                      # write(path, content)
        with open(io, "w", encoding="utf-8") as f:
            return jl_write(f, *a)
    for x in a:
        io.write(jl_str(x))
    return None


def joinpath(*a):
    return os.path.join(*a)


def jl_error(*a):
    raise JlError("".join(jl_str(x) for x in a))


def jl_exit(code=0):
    raise SystemExit(code)


def jl_typeof(x):
    return _interp.typeof(x)


def jl_isa(x, t):
    return _interp.isa(x, t)


def convert(t, x):
    return _interp.convert_to(t.base if isinstance(t, JTypeApp) else t, [x])


def jl_eval(ast):
    return _interp.comp(ast)(_interp.genv)


def haskey(d, k):
    from .interp import dict_key
    return dict_key(k) in d


def jl_get(d, k, default):
    return d.get(k, default)


def jl_keys(d):
    return JList(list(d.keys()))


def jl_values(d):
    return JList(list(d.values()))


def include(path):
    if not os.path.isabs(path):
        path = os.path.join(os.path.dirname(_interp.cur_file[-1]), path)
    return _interp.run_file(path)


def jl_rand(*a, **k):
    raise JlError("rand() called without a scripted stream: define `rand` in the fixture script (Main) before including "
                  "the reference sources")


def jl_isapprox(a, b, atol=0.0, rtol=None):
    if rtol is None:
        rtol = math.sqrt(2.220446049250313e-16) if atol == 0 else 0.0
    return abs(a - b) <= max(atol, rtol * max(abs(a), abs(b)))


# ---- Dict ---------------------------------------------------------------------------------------------------------------
def make_dict(*pairs):
    d = {}
    for p in pairs:
        if isinstance(p, tuple) and len(p) == 2:
            d[p[0]] = p[1]
        else:
            raise JlError("Dict(...) expects pairs")
    return d


# ---- ArgParse -------------------------------------------------------------------------------------------------------------
class ArgSettings:
    def __init__(self):
        self.entries = []


def compile_arg_table(interp: Interp, block):
    """The body of `@add_arg_table! s begin ... end`: option-name lines followed by `key = value` settings."""
    stmts = block[1] if block[0] == "block" else [block]
    entries = []
    for st in stmts:
        if st[0] == "str" or (st[0] == "tuple" and all(x[0] == "str" for x in st[1])):
            names = [st] if st[0] == "str" else st[1]
            entries.append({"names": [interp.comp(n) for n in names], "props": []})
        elif st[0] == "assign" and st[1][0] == "name":
            if not entries:
                raise JlError("@add_arg_table!: a setting before any option")
            entries[-1]["props"].append((st[1][1], interp.comp(st[2])))
        else:
            raise JlError(f"@add_arg_table!: unsupported line {st[0]}")
    return entries


def add_arg_table(settings: ArgSettings, entries, env):
    for e in entries:
        names = [n(env) for n in e["names"]]
        props = {k: v(env) for k, v in e["props"]}
        settings.entries.append((names, props))


def add_arg_table_fn(settings: ArgSettings, *rest):
    """The function form: add_arg_table!(s, "--opt" | ["--opt", "-o"], Dict(:arg_type => T, :default => v), …)."""
    k = 0
    while k < len(rest):
        names = rest[k]
        names = [names] if isinstance(names, str) else [n for n in iterate(names)]
        props = {}
        if k + 1 < len(rest) and isinstance(rest[k + 1], dict):
            props = {(key.name if isinstance(key, Sym) else key): v for key, v in rest[k + 1].items()}
            k += 1
        settings.entries.append((names, props))
        k += 1
    return settings


def parse_args(*a):
    settings = a[-1]
    argv = list(a[0]) if len(a) == 2 else list(_interp.genv.vars["ARGS"])
    T = _interp.types
    out = {}
    by_flag = {}
    for names, props in settings.entries:
        long = [n for n in names if n.startswith("--")]
        dest = (long[0][2:] if long else names[0].lstrip("-"))
        if props.get("action") == Sym("store_true"):
            out[dest] = False
        else:
            out[dest] = props.get("default", None)
        for n in names:
            by_flag[n] = (dest, props)
    k = 0
    while k < len(argv):
        tok = argv[k]
        val = None
        if tok.startswith("--") and "=" in tok:
            tok, val = tok.split("=", 1)
        if tok not in by_flag:
            raise JlError(f"ArgParse: unrecognized option {tok}")
        dest, props = by_flag[tok]
        if props.get("action") == Sym("store_true"):
            out[dest] = True
            k += 1
            continue
        if val is None:
            k += 1
            if k >= len(argv):
                raise JlError(f"ArgParse: option {tok} needs an argument")
            val = argv[k]
        t = props.get("arg_type", T["Any"])
        if t is T["Float64"]:
            val = float(val)
        elif t is T["Int64"]:
            val = int(val)
        out[dest] = val
        k += 1
    return out


# ---- Logging ------------------------------------------------------------------------------------------------------------
class LogLevel:
    def __init__(self, v):
        self.v = v


class Logger:
    def __init__(self, level):
        self.level = level


def console_logger(io=None, level=None):
    return Logger(level.v if level is not None else 0)


def null_logger():
    return Logger(10 ** 9)


def global_logger(lg):
    _interp.log_level = lg.level
    return lg


def install(interp: Interp):
    global _interp
    _interp = interp
    interp.B = sys.modules[__name__]
    g = interp.genv.vars
    T = interp.types
    g.update({
        "π": math.pi, "pi": math.pi, "Inf": math.inf, "NaN": math.nan, "ℯ": math.e, "nothing": None, "undef": None,
        "stdout": interp.stdout, "stderr": interp.stderr,
        "sin": _real1(math.sin, "sin"), "cos": _real1(math.cos, "cos"), "tan": _real1(math.tan, "tan"),
        "acos": _real1(math.acos, "acos"), "asin": _real1(math.asin, "asin"), "atan": math.atan2 if False else (lambda *a: math.atan(a[0]) if len(a) == 1 else math.atan2(a[0], a[1])),
        "exp": jl_exp, "log": jl_log, "sqrt": jl_sqrt, "abs": jl_abs, "min": jl_min, "max": jl_max,
        "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh, "coth": lambda x: 1.0 / math.tanh(x),
        "round": jl_round, "floor": jl_floor, "ceil": jl_ceil, "sign": lambda x: (x > 0) - (x < 0) if isinstance(x, int) else math.copysign(1.0, x) if x != 0 else 0.0,
        "isnan": lambda x: x != x, "isinf": lambda x: isinstance(x, float) and math.isinf(x),
        "isfinite": lambda x: not isinstance(x, float) or math.isfinite(x), "float": lambda x: float(x) if not isinstance(x, np.ndarray) else x.astype(np.float64),
        "mod": lambda a, b: a % b, "rem": lambda a, b: int(math.fmod(a, b)) if isinstance(a, int) and isinstance(b, int) else math.fmod(a, b),
        "div": lambda a, b: (abs(a) // abs(b)) * (1 if (a >= 0) == (b >= 0) else -1), "iseven": lambda x: x % 2 == 0, "isodd": lambda x: x % 2 == 1,
        "zeros": zeros, "ones": ones, "fill": fill, "length": length, "size": size, "sum": jl_sum, "prod": jl_prod,
        "cumsum": cumsum, "map": jl_map, "foreach": foreach, "dot": dot, "hcat": hcat, "vcat": vcat, "reshape": reshape,
        "transpose": transpose, "permutedims": permutedims, "copy": jl_copy, "deepcopy": jl_copy, "view": view, "collect": collect, "findnext": findnext,
        "findfirst": findfirst, "push!": push_, "popfirst!": popfirst_, "pop!": pop_, "append!": append_, "empty!": empty_,
        "isempty": isempty, "eigvals": eigvals, "norm": norm, "maximum": jl_maximum, "minimum": jl_minimum, "any": jl_any,
        "all": jl_all, "first": first, "last": last, "vec": lambda x: _arr(x).reshape(-1, order="F").copy(),
        "string": string, "println": println, "print": jl_print, "repr": jl_repr, "parse": jl_parse, "split": split,
        "strip": lambda s: s.strip(), "join": lambda xs, sep="": sep.join(jl_str(x) for x in iterate(xs)),
        "open": jl_open, "close": jl_close, "readlines": readlines, "writedlm": writedlm, "write": jl_write,
        "flush": lambda io: io.flush(), "isfile": os.path.isfile, "joinpath": joinpath, "dirname": os.path.dirname,
        "basename": os.path.basename, "abspath": os.path.abspath, "mkpath": lambda p: os.makedirs(p, exist_ok=True),
        "error": jl_error, "exit": jl_exit, "typeof": jl_typeof, "isa": jl_isa, "isnothing": lambda x: x is None,
        "convert": convert, "eval": jl_eval, "haskey": haskey, "get": jl_get, "keys": jl_keys, "values": jl_values,
        "include": include, "rand": jl_rand, "time": _time.time, "isapprox": jl_isapprox,
        "ArgParseSettings": lambda *a, **k: ArgSettings(), "parse_args": parse_args, "add_arg_table!": add_arg_table_fn,
        "global_logger": global_logger, "ConsoleLogger": console_logger,
        "Meta": ModuleNS("Meta", {"parse": parse_expression}),
        "Logging": ModuleNS("Logging", {"Info": LogLevel(0), "Warn": LogLevel(1000), "Error": LogLevel(2000),
                                        "Debug": LogLevel(-1000), "NullLogger": null_logger}),
        "Base": ModuleNS("Base", {}),
        "Glob": ModuleNS("Glob", {"GlobMatch": GlobMatch}), "readdir": jl_readdir, "readdlm": readdlm, "filter": jl_filter, "sort": jl_sort,
        "zip": lambda *xs: JList([tuple(t) for t in zip(*[list(iterate(x)) for x in xs])]),
        "read": jl_read, "pmap": jl_map, "run": lambda cmd: run_cmd_text(cmd) and None,
        "startswith": lambda s, p: s.startswith(p), "endswith": lambda s, p: s.endswith(p),
        "display": lambda x: println(x),
        "identity": lambda x: x, "tuple": lambda *a: tuple(a), "Pair": lambda a, b: (a, b),
        "xor": lambda a, b: a ^ b, "trunc": lambda *a: int(a[1]) if len(a) == 2 else float(int(a[0])),
        "abs2": lambda x: x * x, "sincos": lambda x: (math.sin(x), math.cos(x)),
    })
    # Dict is a type AND a constructor
    T["Dict"].ctors = None
    interp._dict_ctor = make_dict
    interp.builtin_names = set(g.keys())
