"""ctypes binding of libaotb200.so (the C ABI declared in include/aotb200.h).

The signatures are parsed from the header itself so the binding cannot drift from the
declared ABI.  There is deliberately no fallback: if the shared library is missing or a kernel
reports an error the call raises -- nothing in this package computes on the CPU or through
PyTorch operators instead.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaotb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "aotb200.h")

_CTYPE = {
    "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
    "void": None, "const char*": ctypes.c_char_p, "unsigned long long": ctypes.c_ulonglong,
}


class AotbError(RuntimeError):
    pass


def parse_header(path: str = HEADER_PATH):
    """-> {name: (restype_str, [argtype_str, ...])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)          # preprocessor lines
    src = re.sub(r'extern\s+"C"\s*\{', "", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\s*\b(aotb_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = re.sub(r"\s*\*", "*", ret)
        argt = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\s*\*\s*", "* ", a.strip())
                t = a.rsplit(" ", 1)[0].strip() if " " in a else a
                argt.append(t)
        out[name] = (ret, argt)
    return out


def _to_ctype(t: str):
    if t.endswith("*") and t != "const char*":
        return ctypes.c_void_p
    return _CTYPE[t]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AotbError(
                f"{LIB_PATH} is missing: build it with `python -m aot_benchmark_b200.build` "
                "(or __graft_entry__.build()).  aot_benchmark_b200 has no CPU / PyTorch fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in parse_header().items():
            fn = getattr(h, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = _to_ctype(ret)
            fn.argtypes = [_to_ctype(a) for a in args]
        _lib = h
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().aotb_last_error_string()
        raise AotbError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
