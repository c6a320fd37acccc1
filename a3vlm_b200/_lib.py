"""ctypes binding of liba3vlm_b200.so (the C ABI declared in include/a3vlm_b200.h).

The prototypes are parsed from the header itself, so the Python side can never drift from the
C side.  There is no CPU fallback: if the shared library is missing the import of any op fails
loudly (`A3LibraryError`), and a compute call without an sm_100 device returns A3_ERR_CUDA.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_ROOT = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_ROOT), "include", "a3vlm_b200.h")
# A3_LIB_PATH: development aid for A/B-testing an alternative build of the same library (never a different backend)
LIB_PATH = os.environ.get("A3_LIB_PATH") or os.path.join(_ROOT, "lib", "liba3vlm_b200.so")


class A3LibraryError(RuntimeError):
    pass


class A3Error(RuntimeError):
    pass


_CTYPE = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
    "void": None,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """Returns {name: (return_type, [(ctype_string, arg_name), ...])} for every a3_* prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    text = text.replace('extern "C" {', "")
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(a3_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, arglist)
    return protos


def _to_ctype(t: str):
    if "*" in t:
        return ctypes.c_char_p if t.replace(" ", "") == "constchar*" else ctypes.c_void_p
    t = t.replace("const", "").strip()
    return _CTYPE[t]


_lib = None
PROTOS = parse_header()
MISSING: List[str] = []


def load(strict: bool = True) -> ctypes.CDLL:
    """Loads the library and binds every prototype of the header.  strict=False (bring-up /
    diagnostics only) tolerates symbols that are declared but not yet exported."""
    global _lib
    if _lib is not None:
        if strict and MISSING:
            raise A3LibraryError(f"liba3vlm_b200.so does not export {MISSING}")
        return _lib
    if not os.path.exists(LIB_PATH):
        raise A3LibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  a3vlm_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # header declares a symbol the library does not export
            if strict:
                raise A3LibraryError(f"liba3vlm_b200.so does not export {name}") from e
            MISSING.append(name)
            continue
        fn.restype = _to_ctype(ret)
        fn.argtypes = [_to_ctype(t) for t, _ in args]
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().a3_last_error()
        raise A3Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
