"""ctypes binding of libtaper_b200.so — the C ABI declared in include/taper_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.  There
is NO fallback: if the shared library is missing the import raises, and every call that returns a
non-zero status raises ``TaperError`` with the library's thread-local message (the reference panics
on the same conditions, e.g. src/ops.rs:11-15, 201-208).
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADERS = [os.path.join(ROOT, "include", "taper_b200.h"), os.path.join(ROOT, "include", "taper_b200_host.h")]
LIB_PATH = os.path.join(_HERE, "libtaper_b200.so")


class TaperError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"taper_b200 error {code}: {msg}")
        self.code = code


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n", "c_in", "h", "w", "c_out", "kh", "kw", "stride_h", "stride_w",
                                       "pad_h", "pad_w", "dil_h", "dil_w")]


class PoolDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n", "c", "h", "w", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w")]


class StepDesc(C.Structure):
    """tp_step_desc (include/taper_b200.h): a chain of Linear(+ReLU) layers + classifier head + optimizer."""
    MAX_LAYERS = 8
    _fields_ = [("n_layers", C.c_int), ("dims", C.c_int * 9), ("relu", C.c_int * 8), ("batch", C.c_int), ("optimizer", C.c_int),
                ("w_off", C.c_int64 * 8), ("b_off", C.c_int64 * 8), ("arena_len", C.c_int64), ("materialize_grads", C.c_int),
                ("data_parallel", C.c_int)]


_OPAQUE = ("tp_ctx", "tp_buf", "tp_graph", "tp_event", "tp_model", "tp_trainer", "tp_step", "tp_xchg", "tp_dataset", "tp_loader",
           "tp_scheduler", "tp_tensor", "tp_optimizer")
_BASE = {
    "int": C.c_int, "float": C.c_float, "double": C.c_double, "size_t": C.c_size_t, "uint64_t": C.c_uint64, "uint32_t": C.c_uint32,
    "int64_t": C.c_int64, "char": C.c_char, "void": None,
    "tp_conv_desc": ConvDesc, "tp_pool_desc": PoolDesc, "tp_step_desc": StepDesc,
}


def _ctype(decl: str):
    """Map one C parameter/return declaration (name already stripped) to a ctypes type."""
    d = decl.replace("const", " ").replace("struct", " ")
    stars = d.count("*")
    base = d.replace("*", " ").split()
    assert len(base) == 1, decl
    base = base[0]
    if base in _OPAQUE:
        assert stars >= 1, decl
        return C.c_void_p if stars == 1 else C.POINTER(C.c_void_p)
    if base == "void":
        if stars == 0:
            return None
        return C.c_void_p if stars == 1 else C.POINTER(C.c_void_p)
    if base == "char" and stars == 1:
        return C.c_char_p
    t = _BASE[base]
    for _ in range(stars):
        t = C.POINTER(t)
    return t


_PROTO = re.compile(r"([A-Za-z_][\w\s\*]*?)\b(tp_\w+)\s*\(([^()]*)\)\s*;")
_TYPE_WORDS = set(_BASE) | set(_OPAQUE) | {"const", "struct", "unsigned"}


def parse_header(path):
    """Return {function name: (restype decl, [param decls])} for every prototype in the header."""
    with open(path) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    src = re.sub(r"typedef\s+struct\s+\w+\s*\{[^}]*\}\s*\w+\s*;", " ", src, flags=re.S)
    src = re.sub(r"typedef[^;]*;", " ", src)
    src = re.sub(r"enum\s*\{[^}]*\}\s*;", " ", src, flags=re.S)
    out = {}
    for m in _PROTO.finditer(src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                toks = re.findall(r"\*|\w+", a)
                # drop the parameter name: a trailing identifier that is not a type word
                if len(toks) > 1 and toks[-1] != "*" and toks[-1] not in _TYPE_WORDS:
                    toks = toks[:-1]
                params.append(" ".join(toks))
        out[name] = (ret, params)
    return out


def declared_symbols():
    syms = {}
    for h in HEADERS:
        if os.path.exists(h):
            syms.update(parse_header(h))
    return syms


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "taper_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (ret, params) in declared_symbols().items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = _ctype(ret)
        fn.argtypes = [_ctype(p) for p in params]
    return lib


lib = _load()


def last_error() -> str:
    return lib.tp_last_error().decode()


def check(rc):
    if rc != 0:
        raise TaperError(rc, last_error())
    return rc


def _np_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Buf:
    """Owning handle of a tp_buf."""

    def __init__(self, ctx, handle, n):
        self.ctx, self.h, self.n = ctx, handle, n

    def __del__(self):
        try:
            if self.h:
                lib.tp_buf_release(self.h)
                self.h = None
        except Exception:
            pass

    def download(self, n=None, dtype=np.float32):
        n = self.n if n is None else n
        out = np.empty(n, dtype=dtype)
        check(lib.tp_buf_download(self.ctx.h, self.h, out.ctypes.data_as(C.c_void_p), n))
        return out

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.dtype.itemsize == 4 and arr.size <= self.n
        check(lib.tp_buf_upload(self.ctx.h, self.h, arr.ctypes.data_as(C.c_void_p), arr.size))
        return self

    def ptr(self):
        return lib.tp_buf_ptr(self.h)


class Ctx:
    """One device / one stream / one host thread (include/taper_b200.h conventions)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib.tp_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib.tp_ctx_destroy(self.h)
            self.h = None

    def alloc(self, n) -> Buf:
        b = C.c_void_p()
        check(lib.tp_buf_alloc(self.h, n, C.byref(b)))
        return Buf(self, b, n)

    def upload(self, arr) -> Buf:
        arr = np.ascontiguousarray(arr)
        if arr.dtype != np.int32:
            arr = _np_f32(arr)
        return self.alloc(max(arr.size, 1)).upload(arr.reshape(-1)) if arr.size else self.alloc(1)

    def zeros(self, n) -> Buf:
        b = self.alloc(max(n, 1))
        check(lib.tp_buf_fill(self.h, b.h, 0.0, n))
        return b

    def sync(self):
        check(lib.tp_sync(self.h))

    def launches(self) -> int:
        c = C.c_uint64()
        check(lib.tp_ctx_launch_count(self.h, C.byref(c)))
        return c.value

    def call(self, name, *args):
        """Invoke tp_<name>(ctx, *args) with Buf arguments unwrapped; raises on non-zero status."""
        fn = getattr(lib, "tp_" + name)
        conv = [a.h if isinstance(a, Buf) else (C.byref(a) if isinstance(a, (ConvDesc, PoolDesc, StepDesc)) else a) for a in args]
        return check(fn(self.h, *conv))
