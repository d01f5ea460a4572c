"""ctypes binding of libosb200.so (the C ABI declared in include/osb200.h).

The prototypes are parsed from the header itself, so the binding, the header and the
"every declared symbol is exported" test cannot drift apart.  There is NO fallback: if the
library is missing, or there is no CUDA device, callers get a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "osb200.h")
LIB_PATH = os.path.join(_HERE, "libosb200.so")

OSB_OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_BUFFER, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
FMT_PCM16, FMT_ULAW, FMT_ALAW, FMT_F32 = 0, 1, 2, 3

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "int32_t": ctypes.c_int32, "void": None,
}
_DECL = re.compile(r"^\s*(const\s+char\s*\*|uint64_t|int64_t|int|void)\s+(osb_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def parse_header(path: str = HEADER) -> dict[str, tuple[object, list[object]]]:
    """{symbol: (restype, [argtypes])} for every function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, args in _DECL.findall(text):
        ret = ret.strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    toks = [t for t in a.split() if t != "const"]
                    argtypes.append(_CTYPES[toks[0]])
        out[name] = (restype, argtypes)
    return out


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m open_speech_b200.build` "
                    "(open_speech_b200 has no CPU or PyTorch fallback)")
            L = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in parse_header().items():
                fn = getattr(L, name)  # AttributeError here = header/library drift
                fn.restype, fn.argtypes = restype, argtypes
            _lib = L
    return _lib


def last_error() -> str:
    return (lib().osb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a negative return code to the exception type the reference raises for it."""
    if rc == OSB_OK:
        return
    msg = last_error() or f"libosb200 error {rc}"
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_BUFFER:
        raise BufferError(msg)
    raise RuntimeError(msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


def ptr(arr) -> int:
    """Address of a numpy array's / torch tensor's first element."""
    if hasattr(arr, "data_ptr"):
        return arr.data_ptr()
    return arr.ctypes.data


def require_gpu() -> None:
    if lib().osb_device_count() <= 0:
        raise RuntimeError("open_speech_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
