"""ctypes binding of libopengenie_b200.so — the C-ABI boundary (include/opengenie_b200.h).

The prototypes are parsed from the header at import time, so the Python side cannot drift from the ABI.
There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised with
og_last_error()'s message. PyTorch is only used for device memory and streams; every call below passes
raw pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'opengenie_b200.h')
LIB_PATH = os.path.join(_HERE, 'csrc', 'libopengenie_b200.so')


class og_adamw_tensor(ctypes.Structure):
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p),
                ('p_bf16', ctypes.c_void_p), ('n', ctypes.c_int64), ('row_len', ctypes.c_int64),
                ('dst_ld', ctypes.c_int64)]


_SCALARS = {
    'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64, 'size_t': ctypes.c_size_t,
    'float': ctypes.c_float, 'double': ctypes.c_double, 'og_stream_t': ctypes.c_void_p,
}


def _ctype(decl: str):
    decl = decl.replace('const', '').strip()
    if decl.endswith('*') or '*' in decl:
        base = decl.replace('*', '').strip()
        return ctypes.c_char_p if base == 'char' else ctypes.c_void_p
    if decl == 'void':
        return None
    return _SCALARS[decl]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object], List[str]]]:
    """{name: (restype, argtypes, argnames)} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;', ' ', src, flags=re.S)
    src = re.sub(r'typedef\s+enum\s+\w+\s*\{.*?\}\s*\w+\s*;', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(og_\w+)\s*\(([^;{}]*?)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith('typedef'):
            continue
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                mm = re.match(r'(.*?)(\w+)$', a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


PROTOTYPES = parse_header()
_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError — never falls back — if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'open_genie_b200: {LIB_PATH} not found. Build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` or `make -C open_genie_b200/csrc`. There is no CPU / PyTorch fallback path.')
    try:
        import torch  # noqa: F401  (makes libcudart resolvable from torch's bundled copy when the toolkit is absent)
    except Exception:
        pass
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes, _) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().og_last_error().decode()


def launch_count() -> int:
    return int(load().og_launch_count())


# Optional per-launch timing of EVERY entry point (bench.py's per-kernel roofline table): when TIMING is a list,
# each call appends (name, args, start_event, end_event) — CUDA events on the launching stream, no syncs.
TIMING = None


def call(name: str, *args):
    """Call an int-returning entry point; raise on a non-zero status."""
    fn = getattr(load(), name)
    if TIMING is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        TIMING.append((name, args, e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed (status {rc}): {last_error()}')
    return rc
