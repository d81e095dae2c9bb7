"""minijl.ffi — `ccall` for the Julia-subset interpreter: runs the repository's Julia hosts
(polymer-stats_b200/julia/*.jl) against the real shared library, through ctypes.

What the hosts use, and nothing more: `ccall((:sym, lib), Ret, (ArgTypes...,), args...)` with the C scalar types,
`Ptr{T}` of Float64 / Int32 arrays (column-major `Array{T}(undef, dims...)`), `Ptr{S}` of a vector of isbits structs,
`Ptr{Ptr{Cvoid}}` / `Ptr{Int32}` out-parameters as `Ref`, `Ptr{Cvoid}` handles, `Cstring` results, `C_NULL`.
A struct is marshalled field by field in declaration order with natural C alignment — the same rule Julia applies to
an isbits struct — so a wrong field order or type in the host's mirror of `pmc_case` shows up as wrong numbers here
exactly as it would under Julia.

`interp.ffi_libs[path] = obj` substitutes a Python object for a library path (CPU tests without a GPU: the host logic
runs against a mock whose attributes are the entry points, called with the marshalled ctypes arguments)."""
from __future__ import annotations

import ctypes as C
import time as _time

import numpy as np

from .interp import JlError, JList, JStruct, JType, JTypeApp, Sym


class JRef:
    """Ref{T}(x): a mutable cell; `r[]` reads, `r[] = v` writes."""

    def __init__(self, value=None, eltype=None):
        self.value = value
        self.eltype = eltype

    def __repr__(self):
        return f"Ref({self.value!r})"


class CPtr(int):
    """A raw address (Ptr{Cvoid} and friends).  C_NULL is CPtr(0)."""

    def __repr__(self):
        return f"Ptr @0x{int(self):016x}"


_SCALARS = {"Float64": C.c_double, "Float32": C.c_float, "Int64": C.c_int64, "Int32": C.c_int32, "UInt64": C.c_uint64,
            "UInt32": C.c_uint32, "Int8": C.c_int8, "UInt8": C.c_uint8, "Bool": C.c_bool}
_NP = {"Float64": np.float64, "Float32": np.float32, "Int64": np.int64, "Int32": np.int32, "UInt64": np.uint64,
       "UInt32": np.uint32, "Bool": np.bool_}


def _tname(t):
    return t.name if isinstance(t, JType) else None


def ctype_of(interp, t):
    """ctypes type of a Julia type object (scalar, Ptr{…}, Cstring, isbits struct)."""
    if isinstance(t, JTypeApp):
        if t.base.name in ("Ptr", "Ref"):
            return C.c_void_p
        raise JlError(f"ccall: unsupported argument type {t}")
    name = _tname(t)
    if name in _SCALARS:
        return _SCALARS[name]
    if name in ("Cstring", "Ptr"):
        return C.c_char_p if name == "Cstring" else C.c_void_p
    if name in ("Cvoid", "Nothing"):
        return None
    if isinstance(t, JType) and t.is_struct:
        return struct_ctype(interp, t)
    raise JlError(f"ccall: unsupported type {t}")


def struct_ctype(interp, t: JType):
    cache = interp.__dict__.setdefault("_ffi_structs", {})
    if t.name not in cache:
        fields = []
        for fname, ft in zip(t.fields, t.ftypes):
            ct = ctype_of(interp, ft)
            if ct is None:
                raise JlError(f"ccall: field {fname} of {t.name} has no C type")
            fields.append((fname, ct))
        cache[t.name] = type("C_" + t.name, (C.Structure,), {"_fields_": fields})
    return cache[t.name]


def _struct_value(interp, t: JType, v: JStruct):
    ct = struct_ctype(interp, t)
    out = ct()
    for fname, (_, fct) in zip(t.fields, ct._fields_):
        x = v.f[fname]
        setattr(out, fname, int(x) if fct not in (C.c_double, C.c_float) else float(x))
    return out


def _marshal(interp, t, v, keep):
    """One ccall argument → ctypes value.  `keep` collects (buffer, write-back) pairs that must outlive the call."""
    if isinstance(t, JTypeApp) and t.base.name in ("Ptr", "Ref"):
        el = t.params[0] if t.params else None
        if v is None or (isinstance(v, CPtr) and int(v) == 0):
            return C.c_void_p(None)
        if isinstance(v, CPtr):
            return C.c_void_p(int(v))
        if isinstance(v, JRef):        # out-parameter: a one-element buffer written back after the call
            ect = ctype_of(interp, el) if el is not None else C.c_void_p
            ect = ect or C.c_void_p
            cur = v.value
            buf = ect(int(cur) if cur is not None and ect not in (C.c_double, C.c_float) else (float(cur) if cur is not None else 0))

            def back(buf=buf, v=v, ect=ect):
                v.value = CPtr(buf.value or 0) if ect is C.c_void_p else buf.value
            keep.append((buf, back))
            return C.cast(C.pointer(buf), C.c_void_p)
        if isinstance(v, np.ndarray):
            want = _NP.get(_tname(el))
            if want is not None and v.dtype != want:
                raise JlError(f"ccall: array of {v.dtype} passed as Ptr{{{el}}}")
            if not (v.flags["F_CONTIGUOUS"] or v.flags["C_CONTIGUOUS"]):
                raise JlError("ccall: non-contiguous array")
            if v.ndim > 1 and not v.flags["F_CONTIGUOUS"]:
                raise JlError("ccall: a Julia array is column-major; this one is not (interpreter bug)")
            keep.append((v, None))
            return C.c_void_p(v.ctypes.data)
        if isinstance(v, (list, JList)):   # Vector of isbits structs
            if isinstance(el, JType) and el.is_struct:
                ct = struct_ctype(interp, el)
                arr = (ct * len(v))(*[_struct_value(interp, el, x) for x in v])
                keep.append((arr, None))
                return C.cast(arr, C.c_void_p)
            arr = np.array(list(v), dtype=_NP.get(_tname(el), np.float64))
            keep.append((arr, None))
            return C.c_void_p(arr.ctypes.data)
        raise JlError(f"ccall: cannot pass {type(v).__name__} as {t}")
    ct = ctype_of(interp, t)
    if ct in (C.c_double, C.c_float):
        return ct(float(v))
    if ct is C.c_void_p:
        return C.c_void_p(int(v) if v is not None else None)
    if ct is C.c_bool:
        return ct(bool(v))
    if isinstance(v, float):
        raise JlError(f"ccall: InexactError: {t}({v})")
    return ct(int(v))


def make_ccall(interp):
    libs = {}

    def ccall(target, ret, argtypes, *args):
        if not (isinstance(target, tuple) and len(target) == 2):
            raise JlError("ccall: the target must be (:symbol, library)")
        sym, path = target
        sym = sym.name if isinstance(sym, Sym) else str(sym)
        argtypes = tuple(argtypes) if isinstance(argtypes, (tuple, list, JList)) else (argtypes,)
        if len(argtypes) != len(args):
            raise JlError(f"ccall: {sym} declares {len(argtypes)} argument types but {len(args)} arguments were passed")
        lib = getattr(interp, "ffi_libs", {}).get(path)
        if lib is None:
            if path not in libs:
                try:
                    libs[path] = C.CDLL(path)
                except OSError as e:
                    raise JlError(f"could not load library {path!r}: {e}")
            lib = libs[path]
        try:
            fn = getattr(lib, sym)
        except AttributeError:
            raise JlError(f"ccall: could not load symbol {sym!r} from {path}")
        keep = []
        cargs = [_marshal(interp, t, v, keep) for t, v in zip(argtypes, args)]
        rct = ctype_of(interp, ret)
        if isinstance(lib, C.CDLL):
            fn.restype = rct
            fn.argtypes = [type(a) for a in cargs]
        res = fn(*cargs)
        for _, back in keep:
            if back:
                back()
        if rct is None:
            return None
        if rct is C.c_char_p:
            return res if isinstance(res, (bytes, type(None))) else bytes(res)
        if rct is C.c_void_p:
            return CPtr(res or 0)
        if rct in (C.c_double, C.c_float):
            return float(res)
        return int(res)
    return ccall


def install(interp):
    g = interp.genv.vars
    T = interp.types

    def mk(name, sup="Any", abstract=False):
        if name not in T:
            T[name] = JType(name, T[sup], abstract)
        g[name] = T[name]
        return T[name]
    for name, sup in (("Int32", "Signed"), ("Int8", "Signed"), ("UInt8", "Integer"), ("Cvoid", "Any"), ("Cstring", "Any"),
                      ("Ptr", "Any"), ("Ref", "Any")):
        mk(name, sup)
    for alias, name in (("Cdouble", "Float64"), ("Cfloat", "Float32"), ("Cint", "Int32"), ("Cuint", "UInt32"),
                        ("Clonglong", "Int64"), ("Culonglong", "UInt64"), ("Csize_t", "UInt64")):
        g[alias] = T[name]
    g["C_NULL"] = CPtr(0)
    g["ccall"] = make_ccall(interp)
    g["unsafe_string"] = lambda b: (b.decode("utf-8", "replace") if isinstance(b, bytes) else
                                    C.string_at(int(b)).decode("utf-8", "replace"))
    g["time_ns"] = lambda: _time.time_ns()
    g["ENV"] = interp.B.make_dict(*[(k, v) for k, v in __import__("os").environ.items()]) if hasattr(interp.B, "make_dict") else dict(__import__("os").environ)
    interp.ffi_libs = {}
    interp.JRef = JRef
    interp.CPtr = CPtr
