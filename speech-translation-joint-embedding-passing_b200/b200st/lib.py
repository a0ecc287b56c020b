"""ctypes binding of libb200st.so — the C ABI declared in include/b200st.h.

The prototypes are parsed from the header itself, so the Python binding cannot drift from the C
declarations.  There is NO fallback: if the shared library is missing or a symbol is absent, import of
the compute path fails loudly (the product never routes through PyTorch ops or the CPU oracle).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(_ROOT, 'include', 'b200st.h')
LIB_PATH = os.path.join(_HERE, 'libb200st.so')

_CTYPES = {
    'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'float': ctypes.c_float, 'double': ctypes.c_double,
    'b200st_stream_t': ctypes.c_void_p, 'void': None,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Return {symbol: (restype, [argtypes])} for every `b200st_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(b200st_\w+)\s*\(([^;{}]*?)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if name == 'b200st_stream_t':
            continue

        def conv(decl: str):
            decl = decl.strip()
            if decl in ('void', ''):
                return None
            if '*' in decl:
                return ctypes.c_char_p if decl.startswith('const char') else ctypes.c_void_p
            base = decl.replace('const', ' ').split()
            return _CTYPES[base[0]]
        argtypes = [conv(a) for a in args.split(',')] if args not in ('void', '') else []
        argtypes = [a for a in argtypes if a is not None]
        protos[name] = (conv(ret), argtypes)
    return protos


class MissingLibrary(RuntimeError):
    pass


_lib = None


def load(path: str = LIB_PATH):
    """dlopen the library and attach prototypes.  Raises MissingLibrary if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise MissingLibrary(
            f'{path} not found: build it with `python __graft_entry__.py build` (or `make -C '
            f'{os.path.join(_PKG, "csrc")}`).  There is no CPU/PyTorch fallback for the b200st kernels.')
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200st_last_error().decode()


def check(status: int, what: str):
    if status != 0:
        raise RuntimeError(f'b200st {what} failed: {last_error()}')
